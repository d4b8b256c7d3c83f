"""The C-ABI library loads and exports every symbol include/unimp_b200.h declares (no compute)."""
import ctypes
import os
import re

from util import ROOT

from unimp_b200 import _lib


def header_symbols():
    src = open(os.path.join(ROOT, "include", "unimp_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(unimp_\w+)\s*\(", src)))


def test_library_exists_and_loads():
    assert os.path.exists(_lib.LIB_PATH), "run __graft_entry__.build() first"
    lib = _lib.load()
    assert lib.unimp_version() == 1


def test_every_header_symbol_is_exported_and_bound():
    syms = header_symbols()
    assert len(syms) >= 20
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(raw, s), f"{s} declared in the header but not exported"
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature in unimp_b200/_lib.py"
    for s in _lib.SIGNATURES:
        assert s in syms, f"{s} bound in _lib.py but not declared in the header"


def test_error_string_and_argument_validation_without_gpu():
    lib = _lib.load()
    # NULL pointers must be rejected before any launch (works on a GPU-less box)
    rc = lib.unimp_text_time(None, 1, 1, 1, 0, 1, None, None)
    assert rc == -1
    assert b"NULL" in lib.unimp_last_error_string()
    rc = lib.unimp_focal_ce_fwd(None, 0, None, None, 2.0, 1, None, None, None, None, None, 1, 2, 3, 1, 1, None)
    assert rc == -1


def test_ops_refuse_cpu_tensors():
    import pytest
    import torch

    from unimp_b200 import ops

    with pytest.raises(_lib.UnimpError):
        ops.layer_norm(torch.zeros(2, 8), torch.ones(8), torch.zeros(8))
