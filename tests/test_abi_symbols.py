"""The C-ABI library loads on a machine without a GPU and exports exactly what include/ptk.h declares."""
import ctypes
import os
import subprocess

import ptk_b200
from ptk_b200 import _lib


def test_header_and_binding_in_sync():
    assert set(_lib.header_symbols()) == set(_lib.SIGNATURES)


def test_library_exports_every_declared_symbol():
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for name in _lib.header_symbols():
        assert hasattr(handle, name), name
    assert handle.ptk_version() == 1


def test_exported_symbols_are_only_the_abi():
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l and "ptk_" in l.split()[-1]}
    c_abi = {s for s in exported if s.startswith("ptk_")}
    assert c_abi == set(_lib.header_symbols())


def test_sass_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = {l.split(".")[-2] for l in out.splitlines() if "sm_" in l}
    assert archs == {"sm_100a"}, out


def test_error_codes_without_gpu():
    L = _lib.lib()
    # shape validation happens before any CUDA call, so it works without a device
    rc = L.ptk_chamfer_fwd(None, None, 1, 10, 10, None, None, None, None, None, None, 0, None)
    assert rc == _lib.PTK_ERR_SHAPE and "null" in _lib.last_error()
    # PairAux + padded SoA clouds (+ the pruned scan's staging copy) + its boxes (leaves + two levels: 7+1+1 and 4+1+1)
    # + keys / lists / flags
    # + rescue counters and bad-cloud flags
    assert L.ptk_chamfer_workspace_bytes(2, 100, 50) == (16 * 2 + 32 * 2 * (128 + 64) + 32 * 2 * (9 + 6) + 16 * 2 * 150
                                                         + 16 * 2)
    assert L.ptk_chamfer_workspace_bytes(0, 100, 50) == 0


def test_missing_extension_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    try:
        _lib.lib()
        raise AssertionError("expected ImportError")
    except ImportError as e:
        assert "no CPU fallback" in str(e)


def test_cpu_tensors_are_rejected():
    import pytest
    import torch
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ptk_b200.ops.chamfer(torch.rand(1, 8, 3), torch.rand(1, 8, 3))
