"""Installing the B200 cell into the unmodified reference.

``STC_Encoder`` / ``STC_Decoder`` look the class name ``STC_Cell`` up in their module globals when they
are constructed (/root/reference/framework/STC_GNN.py:95,151), so rebinding that global before
``STCGNN(...)`` is built swaps the cell and nothing else; ``Model_Trainer.py`` and ``Main.py`` then run
unchanged.  (A shim module on PYTHONPATH does not work: the script directory is sys.path[0] and wins.)
"""
import os
import runpy
import sys

from .cell import STC_Cell


_NO_GROUP = object()


def install(stc_gnn_module=None, dp_group=_NO_GROUP, dp_average: bool = False):
    """Rebind ``STC_GNN.STC_Cell`` to the B200 cell. Returns the patched module.

    ``dp_group`` (a ``torch.distributed`` group, or None for the default group): additionally rebind
    ``MGP_Gen.forward`` to its batch-data-parallel form (``stc_gnn_b200/mgp.py``) so that the supports generated from a
    batch shard -- and every gradient of the generator -- equal the single-process global-batch values."""
    if stc_gnn_module is None:
        import STC_GNN as stc_gnn_module  # noqa: N813  (the reference's module name)
    if getattr(stc_gnn_module, "STC_Cell", None) is not STC_Cell:
        stc_gnn_module._reference_STC_Cell = stc_gnn_module.STC_Cell
        stc_gnn_module.STC_Cell = STC_Cell
    if dp_group is not _NO_GROUP:
        from .mgp import patch_generator
        patch_generator(stc_gnn_module, dp_group, dp_average)
    return stc_gnn_module


def uninstall(stc_gnn_module=None):
    if stc_gnn_module is None:
        import STC_GNN as stc_gnn_module  # noqa: N813
    ref = getattr(stc_gnn_module, "_reference_STC_Cell", None)
    if ref is not None:
        stc_gnn_module.STC_Cell = ref
    from .mgp import unpatch_generator
    unpatch_generator(stc_gnn_module)
    return stc_gnn_module


def run_main(framework_dir: str, argv):
    """Run the reference's Main.py unmodified with the B200 cell installed:
    ``run_main('/path/to/STC-GNN/framework', ['-city', 'SF', '-device', 'cuda:0'])``."""
    framework_dir = os.path.abspath(framework_dir)
    sys.dont_write_bytecode = True
    if framework_dir not in sys.path:
        sys.path.insert(0, framework_dir)
    old_cwd, old_argv = os.getcwd(), sys.argv
    os.chdir(framework_dir)
    try:
        install()
        sys.argv = [os.path.join(framework_dir, "Main.py")] + list(argv)
        runpy.run_path("Main.py", run_name="__main__")
    finally:
        sys.argv = old_argv
        os.chdir(old_cwd)


if __name__ == "__main__":
    run_main(sys.argv[1], sys.argv[2:])
