"""CPU tests of the boundary: the C-ABI library builds, loads and exports every symbol include/stc_b200.h
declares; host-only queries work; the Python cell mirrors the reference's surface and refuses to run
without CUDA (no fallback).  No kernel is launched here."""
import ctypes
import os
import re
import sys

import pytest
import torch

import stc_gnn_b200 as S
from stc_gnn_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.environ.get("STC_REF_DIR", "/root/reference/framework")


def header_functions():
    src = open(os.path.join(ROOT, "include", "stc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(stc_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = header_functions()
    assert set(names) == set(_lib.EXPORTED), (names, _lib.EXPORTED)
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), f"{n} not exported"
    assert lib.stc_abi_version() == _lib.ABI_VERSION


def test_buffer_size_queries_are_host_only():
    lib = _lib.load()
    d = _lib.StcDims(32, 100, 5, 16, 16, 2, 2, 0, 1)
    saved, scratch = lib.stc_cell_saved_bytes(d), lib.stc_cell_bwd_scratch_bytes(d)
    R = 32 * 100 * 5
    assert saved >= 4 * R * (3 * 16 + 2 * 16 + 16 + 16) and scratch >= 4 * R * 32
    bad = _lib.StcDims(1, 0, 5, 1, 16, 2, 2, 0, 1)
    assert lib.stc_cell_saved_bytes(bad) == 0
    assert b"bad dims" in lib.stc_last_error()


def test_cell_surface_matches_reference_contract():
    cell = S.STC_Cell(100, 5, 2, 2, 1, 16)
    sd = cell.state_dict()
    assert list(sd.keys()) == ["gates.W", "gates.b", "candi.W", "candi.b"]
    assert tuple(sd["gates.W"].shape) == (17 * 4, 32) and tuple(sd["candi.W"].shape) == (17 * 4, 16)
    assert float(sd["gates.b"].abs().sum()) == 0.0
    assert tuple(cell.init_hidden(3).shape) == (3, 100, 5, 16)
    nb = S.STC_Cell(10, 3, 3, 2, 4, 8, use_bias=False)
    assert list(nb.state_dict().keys()) == ["gates.W", "candi.W"]
    with pytest.raises(NotImplementedError):
        S.STC_Cell(10, 3, 2, 2, 1, 4, activation=torch.nn.Tanh)


def test_no_cpu_fallback():
    cell = S.STC_Cell(6, 3, 2, 2, 1, 4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        cell(Gs=torch.zeros(6, 6), Gc=torch.zeros(3, 3), Xt=torch.zeros(2, 6, 3, 1), Ht_1=torch.zeros(2, 6, 3, 4))
    with pytest.raises(AssertionError):
        cell(Gs=torch.zeros(6, 6), Gc=torch.zeros(3, 3), Xt=torch.zeros(6, 3, 1), Ht_1=torch.zeros(2, 6, 3, 4))


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "stc_gnn_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text, f"{f} mentions the oracle"


@pytest.mark.skipif(not os.path.isdir(REF_DIR), reason="reference tree not present (GPU box)")
def test_seeded_init_and_checkpoint_compat_with_live_reference():
    sys.dont_write_bytecode = True
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import STC_GNN as ref
    torch.manual_seed(123)
    a = ref.STC_Cell(20, 4, 3, 2, 2, 8)
    torch.manual_seed(123)
    b = S.STC_Cell(20, 4, 3, 2, 2, 8)
    for (ka, va), (kb, vb) in zip(a.state_dict().items(), b.state_dict().items()):
        assert ka == kb and torch.equal(va, vb), ka
    b.load_state_dict(a.state_dict(), strict=True)
    # whole-model: rebinding the module global swaps the cell inside STC_Encoder/STC_Decoder only
    try:
        S.install(ref)
        torch.manual_seed(5)
        enc = ref.STC_Encoder(12, 3, 2, 2, 1, 4, 2)
        assert all(isinstance(c, S.STC_Cell) for c in enc.cell_list)
        from stc_gnn_b200.install import uninstall
        uninstall(ref)
        torch.manual_seed(5)
        enc_ref = ref.STC_Encoder(12, 3, 2, 2, 1, 4, 2)
        assert list(enc.state_dict().keys()) == list(enc_ref.state_dict().keys())
        for k in enc.state_dict():
            assert torch.equal(enc.state_dict()[k], enc_ref.state_dict()[k]), k
        enc.load_state_dict(enc_ref.state_dict(), strict=True)
    finally:
        from stc_gnn_b200.install import uninstall
        uninstall(ref)
