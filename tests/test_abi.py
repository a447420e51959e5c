"""The C-ABI library loads and exports every symbol include/mcg.h declares (no compute: runs without a GPU)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "mcg.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mcg_[a-z0-9_]+)\s*\(", src)))


def _ensure_built():
    from mocogan_chainer_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return _lib


def test_header_symbols_exported_and_bound():
    _lib = _ensure_built()
    names = _declared()
    assert len(names) >= 25
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, "declared in mcg.h but not exported: %s" % missing
    unbound = [n for n in names if n not in _lib.SIGNATURES]
    assert not unbound, "declared in mcg.h but not bound in _lib.SIGNATURES: %s" % unbound
    extra = [n for n in _lib.SIGNATURES if n not in names]
    assert not extra, "bound but not declared in mcg.h: %s" % extra


def test_version_and_error_string_without_gpu():
    _lib = _ensure_built()
    lib = _lib.load()
    assert lib.mcg_version() == 100
    assert isinstance(lib.mcg_last_error(), bytes)
    assert lib.mcg_colreduce_workspace_bytes(1000, 64) == 16 * 2 * 64 * 4 + 256    # 16 slot rows + the ticket counter


def test_shape_errors_are_reported_before_any_launch():
    """Host-side validation: inconsistent geometry is rejected with a message, no GPU needed."""
    _lib = _ensure_built()
    lib = _lib.load()
    g = _lib.ConvGeom(2, 64, 64, 1, 16, 16, 1, 9, 8, 1, 4, 4, 1, 2, 2, 0, 1, 1)  # Ho should be 8
    rc = lib.mcg_conv_fprop(ctypes.byref(g), 1, 1, None, 1, 0, 0, 0, None, 0, None)
    assert rc == -1 and b"inconsistent" in lib.mcg_last_error()
    rc = lib.mcg_loss_dis(None, None, None, None, 4, 1, 0, None, None, None, None)
    assert rc == -1


def test_no_cpu_fallback():
    import pytest
    import torch
    from mocogan_chainer_b200 import kernels
    from mocogan_chainer_b200._lib import McgError
    with pytest.raises(McgError):
        kernels.ptr(torch.zeros(4))   # a CPU tensor can never reach a kernel
