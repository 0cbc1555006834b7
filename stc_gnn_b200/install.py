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


class SlicedLoader:
    """Iterates a reference ``IncDataset`` (``Data_Container.py:43-66``) in contiguous batches by slicing its tensors.

    Same batches, in the same order, as the ``DataLoader(dataset, batch_size, shuffle=False)`` the reference builds
    (``Data_Container.py:102``: no shuffling, last partial batch kept) -- without default-collating ``batch_size``
    single-item device tensors per step."""

    def __init__(self, dataset, batch_size: int):
        self.dataset, self.batch_size = dataset, int(batch_size)
        self.x, self.y = dataset.inputs["x_seq"], dataset.output
        self.n = len(dataset)

    def __len__(self):
        return (self.n + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        for s in range(0, self.n, self.batch_size):
            e = min(self.n, s + self.batch_size)
            yield self.x[s:e], self.y[s:e]


def _patch_loop_hygiene(data_container_module):
    """Trainer-loop hygiene around the frozen files; returns an undo function.

    * ``DataGenerator.get_data_loader`` keeps the reference's own windowing and split and only swaps each DataLoader
      for a ``SlicedLoader`` over the same dataset;
    * ``torch.cuda.empty_cache()`` -- called after EVERY step by ``Model_Trainer.py:87``, which hands every cached block
      back to the driver and makes the next step cudaMalloc all of its buffers again -- becomes a no-op for the run."""
    import torch
    gen_cls = data_container_module.DataGenerator
    orig_get, orig_empty = gen_cls.get_data_loader, torch.cuda.empty_cache

    def get_data_loader(self, params, data):
        loaders = orig_get(self, params, data)
        return {mode: SlicedLoader(dl.dataset, dl.batch_size) for mode, dl in loaders.items()}

    gen_cls.get_data_loader = get_data_loader
    torch.cuda.empty_cache = lambda: None

    def undo():
        gen_cls.get_data_loader = orig_get
        torch.cuda.empty_cache = orig_empty
    return undo


def run_main(framework_dir: str, argv, hygiene: bool = False, install_cell: bool = True):
    """Run the reference's Main.py unmodified with the B200 cell installed:
    ``run_main('/path/to/STC-GNN/framework', ['-city', 'SF', '-device', 'cuda:0'])``.

    ``hygiene``: additionally feed the trainer contiguous batch slices instead of per-item collation and neutralise its
    per-step ``torch.cuda.empty_cache()`` (see ``_patch_loop_hygiene``); ``install_cell=False`` runs the stock cell (for
    A/B timing of the same script)."""
    framework_dir = os.path.abspath(framework_dir)
    sys.dont_write_bytecode = True
    if framework_dir not in sys.path:
        sys.path.insert(0, framework_dir)
    old_cwd, old_argv = os.getcwd(), sys.argv
    os.chdir(framework_dir)
    undo = None
    try:
        if install_cell:
            install()
        if hygiene:
            import Data_Container  # noqa: N813  (the reference's module)
            undo = _patch_loop_hygiene(Data_Container)
        sys.argv = [os.path.join(framework_dir, "Main.py")] + list(argv)
        runpy.run_path("Main.py", run_name="__main__")
    finally:
        if undo is not None:
            undo()
        sys.argv = old_argv
        os.chdir(old_cwd)


if __name__ == "__main__":
    _args = sys.argv[2:]      # python -m stc_gnn_b200.install <framework dir> [--hygiene] <Main.py arguments ...>
    _hyg = "--hygiene" in _args
    run_main(sys.argv[1], [x for x in _args if x != "--hygiene"], hygiene=_hyg)
