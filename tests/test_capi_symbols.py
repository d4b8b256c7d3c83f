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


def header_prototypes():
    """name -> list of parameter type strings, parsed from the header's prototypes."""
    src = open(os.path.join(ROOT, "include", "unimp_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    out = {}
    for m in re.finditer(r"\b(unimp_\w+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        params = " ".join(m.group(2).split())
        out[m.group(1)] = [] if params in ("", "void") else [p.strip() for p in params.split(",")]
    return out


def test_ctypes_signatures_match_the_header_prototypes():
    """Arity and pointer / integer / float class of every argument in unimp_b200/_lib.py agree with
    include/unimp_b200.h (a drifted binding corrupts arguments silently)."""
    protos = header_prototypes()
    assert set(protos) == set(_lib.SIGNATURES)

    def klass(ctype_decl):
        if "*" in ctype_decl:
            return "ptr"
        if re.search(r"\bunimp_view\b|\bUnimpView\b|\bView\b", ctype_decl):
            return "view"
        if re.search(r"\bfloat\b", ctype_decl):
            return "f32"
        if re.search(r"\bint64_t\b", ctype_decl):
            return "i64"
        if re.search(r"\bint\b", ctype_decl):
            return "i32"
        return "view"   # a by-value struct typedef

    def cklass(t):
        if t is _lib.View:
            return "view"
        if t in (ctypes.c_void_p, ctypes.c_char_p) or isinstance(t, type(ctypes.POINTER(ctypes.c_int64))) and t not in (
                ctypes.c_int, ctypes.c_int64, ctypes.c_float):
            return "ptr"
        return {ctypes.c_int: "i32", ctypes.c_int64: "i64", ctypes.c_float: "f32"}[t]

    for name, params in protos.items():
        _, argtypes = _lib.SIGNATURES[name]
        assert len(params) == len(argtypes), (name, params, argtypes)
        for i, (p, t) in enumerate(zip(params, argtypes)):
            assert klass(p) == cklass(t), (name, i, p, t)


def _entry_points():
    skip = {"unimp_version", "unimp_last_error_string", "unimp_device_ok"}
    return sorted(n for n in _lib.SIGNATURES
                  if n not in skip and not n.endswith("_workspace") and not n.endswith("_supported"))


def _null_args(name):
    vals = []
    for t in _lib.SIGNATURES[name][1]:
        if t is _lib.View:
            vals.append(_lib.View(None, 0, 0))
        elif t in (ctypes.c_int, ctypes.c_int64):
            vals.append(1)
        elif t is ctypes.c_float:
            vals.append(1.0)
        else:
            vals.append(None)
    return vals


import pytest  # noqa: E402


@pytest.mark.parametrize("name", _entry_points())
def test_every_entry_point_rejects_null_pointers_before_launching(name):
    """C-ABI error convention: invalid arguments return a negative code and leave a
    message in unimp_last_error_string(); nothing is launched (this runs on a GPU-less box)."""
    lib = _lib.load()
    rc = getattr(lib, name)(*_null_args(name))
    assert rc < 0, (name, rc)
    msg = lib.unimp_last_error_string().decode()
    assert name.replace("unimp__", "").replace("unimp_", "") in msg and "NULL" in msg, msg


def test_shape_dtype_and_alignment_are_validated_before_launching():
    lib = _lib.load()
    P = 4096   # a non-NULL, 16-byte aligned "pointer": validation must fail before it is ever read
    # K5 forward: D must be a multiple of the vector width; dtype must be known; pointers aligned
    f = lib.unimp_gate_residual_ln_fwd
    assert f(None, P, None, P, P, None, P, P, P, 4, 100, 1e-5, 1, None) < 0 and b"multiple" in lib.unimp_last_error_string()
    assert f(None, P, None, P, P, None, P, P, P, 4, 128, 1e-5, 7, None) < 0 and b"dtype" in lib.unimp_last_error_string()
    assert f(None, P + 8, None, P, P, None, P, P, P, 4, 128, 1e-5, 1, None) < 0 and b"aligned" in lib.unimp_last_error_string()
    # K5 backward: a gated branch needs somewhere to put d_branch
    b = lib.unimp_gate_residual_ln_bwd
    assert b(P, P, P, P, P, P, P, P, P, None, None, None, None, P, 4, 128, 0, 1, None) < 0
    assert b"d_branch" in lib.unimp_last_error_string()
    # GELU: element count must be whole vectors
    assert lib.unimp_gelu_fwd(P, P, 12, 1, None) < 0 and lib.unimp_gelu_bwd(P, P, P, 12, 1, None) < 0
    # rotary: rot/2 must be whole vectors
    assert lib.unimp_rotary_qkv_fwd(P, P + 4096, P, P, 1, 1, 2, 64, 24, 0, 1, None) < 0


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    """No CPU fallback: without the built .so every op raises (the product never routes around it)."""
    import torch

    from unimp_b200 import ops

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "libunimp_b200.so"))
    with pytest.raises(_lib.UnimpError, match="no CPU fallback"):
        _lib.load()
    with pytest.raises(_lib.UnimpError):
        ops.gelu(torch.zeros(8))


def test_lm_attn_rejects_unsupported_geometry_before_launching():
    """K4 covers bf16 with head dim 80 only; anything else must be refused with a message naming the
    constraint, before any CUDA call (the host keeps SDPA for those cases: flamingo_lm.fused_neox_layer)."""
    lib = _lib.load()
    buf = (ctypes.c_char * 4096)()
    p = ctypes.cast(buf, ctypes.c_void_p)          # 16-byte aligned dummy pointer, never dereferenced
    p = ctypes.c_void_p((p.value + 15) & ~15)
    assert lib.unimp_lm_attn_supported(256, 32, 80, _lib.BF16) == 1
    assert lib.unimp_lm_attn_supported(256, 32, 64, _lib.BF16) == 0
    assert lib.unimp_lm_attn_supported(256, 32, 80, _lib.F32) == 0
    for dh, dtype, stride, what in ((64, _lib.BF16, 240, "head dim"), (80, _lib.F32, 240, "bf16"),
                                    (80, _lib.BF16, 241, "strides")):
        rc = lib.unimp_lm_attn_fwd(p, p, p, 256 * 32 * 240, 32 * 240, stride, None, p, p, 1, 256, 32, dh, 0.1,
                                   dtype, None)
        assert rc < 0, (dh, dtype, stride, rc)
        assert what in lib.unimp_last_error_string().decode(), lib.unimp_last_error_string().decode()
        rc = lib.unimp_lm_attn_bwd(p, p, p, 256 * 32 * 240, 32 * 240, stride, None, p, p, p, p, p, p, p, 1, 256,
                                   32, dh, 0.1, dtype, None)
        assert rc < 0
    rc = lib.unimp_key_bits(p, 4, p, 1, 64, None)   # element size must be 1 or 8
    assert rc < 0 and "mask" in lib.unimp_last_error_string().decode()
