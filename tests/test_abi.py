"""The C-ABI library loads and exports every symbol include/ccvsq.h declares; argument validation
happens before any CUDA call (so these run without a GPU)."""
import ctypes
import os
import re

from ccvs_b200 import _lib


def _declared_symbols():
    text = open(_lib.HEADER_PATH).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ccvsq_[a-z_0-9]+)\s*\(", text)))


def test_library_exists_in_tree():
    assert os.path.exists(_lib.LIB_PATH), "run __graft_entry__.build() first"
    assert os.path.commonpath([_lib.LIB_PATH, os.path.dirname(os.path.dirname(_lib.__file__))]) != "/"


def test_every_declared_symbol_is_exported_and_bound():
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 14
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in ccvsq.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == declared


def test_version():
    assert _lib.load().ccvsq_version() == 202


def test_argument_validation_without_gpu():
    lib = _lib.load()
    null = ctypes.c_void_p(0)
    one = ctypes.c_void_p(16)
    assert lib.ccvsq_prepare_codebook(null, 4, 4, null, null, null, null) == -5    # NULL_POINTER
    assert b"non-null" in lib.ccvsq_last_error()
    assert lib.ccvsq_prepare_codebook(one, 0, 4, one, null, null, null) == -1       # BAD_SHAPE
    bad = _lib.Layout(4, 6, 4, 4)      # C not divisible by mult
    assert lib.ccvsq_search_exact(one, bad, one, one, 8, one, null) == -1
    odd = _lib.Layout(4, 100, 4, 1)    # D = 100: not a tensor-core shape
    good = _lib.Layout(4, 128, 4, 1)
    assert lib.ccvsq_screen(one, odd, one, one, 256, 1.0, 4, one, one, one, one, one, null) == -2    # UNSUPPORTED
    assert lib.ccvsq_screen(one, good, one, one, 256, 1.0, 9, one, one, one, one, one, null) == -1   # n_cand > 8
    assert lib.ccvsq_screen(ctypes.c_void_p(8), good, one, one, 256, 1.0, 4, one, one, one, one, one, null) == -3  # MISALIGNED
    assert lib.ccvsq_screen(one, good, one, one, 256, 1.0, 4, null, one, one, one, one, null) == -5  # NULL_POINTER
    assert lib.ccvsq_codebook_rows(1024) == 1056 and lib.ccvsq_codebook_rows(96) == 128 and lib.ccvsq_codebook_rows(192) == 256
    assert lib.ccvsq_finalize(null, null, null, null, 4, 4, 0.0, 1.0, 0.25, null, null, null, null) == -1


def test_composite_argument_validation_without_gpu():
    lib = _lib.load()
    null = ctypes.c_void_p(0)
    assert lib.ccvsq_quantize_forward(None, null) == -5
    a = _lib.ForwardArgs()
    assert a.struct_size == ctypes.sizeof(_lib.ForwardArgs)
    a.struct_size -= 8                                                        # a stale binding (one trailing field short)
    assert lib.ccvsq_quantize_forward(ctypes.byref(a), null) == -1
    assert b"struct_size" in lib.ccvsq_last_error()
    a.struct_size += 8
    a.flags = 4
    assert lib.ccvsq_quantize_forward(ctypes.byref(a), null) == -1            # reserved flag bits
    a.flags = 0
    assert lib.ccvsq_quantize_forward(ctypes.byref(a), null) == -5            # z / E / header / idx missing
    a.z = a.E = a.header = a.idx = 16
    a.lay = _lib.Layout(4, 64, 16, 1)
    a.K = 0
    assert lib.ccvsq_quantize_forward(ctypes.byref(a), null) == -1            # K <= 0
    a.K = 256
    a.search_mode = 7
    assert lib.ccvsq_quantize_forward(ctypes.byref(a), null) == -1            # unknown search mode
    a.search_mode = 1
    a.n_cand = 4
    assert lib.ccvsq_quantize_forward(ctypes.byref(a), null) == -1            # workspace too small
    assert b"workspace" in lib.ccvsq_last_error()
    need = lib.ccvsq_forward_workspace_bytes(64, 256, 64, 1, 4, 1)
    # e_sq + BF16 shadow (>= 288 x 80 x 2) + queue arrays, 256-byte aligned sections
    assert need >= 256 * 4 + 288 * 80 * 2 + 64 * 4 + 64 * 16 + 64 + 64 * 16
    assert lib.ccvsq_forward_workspace_bytes(64, 256, 64, 2, 4, 0) == 256     # exact search, cached codebook: nothing
    one = ctypes.c_void_p(16)
    assert lib.ccvsq_quantize_backward(null, a.lay, one, 8, one, null, one, 0.25, one, null, null, null) == -5
    assert lib.ccvsq_quantize_backward(one, a.lay, one, 8, one, null, one, 0.25, null, null, one, null) == -5   # dE without resid


def test_layout_struct_matches_header():
    assert ctypes.sizeof(_lib.ForwardArgs) == 192
    assert _lib.ForwardArgs.struct_size.offset == 0 and _lib.ForwardArgs.z.offset == 8
    assert _lib.ForwardArgs.resid.offset == 176 and _lib.ForwardArgs.counts_f32.offset == 184
    assert _lib.ForwardArgs.lay.offset == 16 and _lib.ForwardArgs.e_sq.offset == 80 and _lib.ForwardArgs.idx.offset == 128
    assert ctypes.sizeof(_lib.Layout) == 24     # int64 + 3 x int32 (+4 pad)
    assert _lib.Layout.G.offset == 0 and _lib.Layout.C.offset == 8 and _lib.Layout.mult.offset == 16


def integration_stub_namespace():
    """Execute the ctypes stub printed in INTEGRATION.md section 2 verbatim (cwd = repo root, as the text assumes)."""
    root = os.path.dirname(_lib.HEADER_PATH.rstrip("/")).rsplit("/include", 1)[0]
    text = open(os.path.join(root, "INTEGRATION.md")).read()
    m = re.search(r"```python\n(# --- ctypes stub.*?)```", text, flags=re.S)
    assert m, "INTEGRATION.md lost its ctypes stub block"
    ns = {"__name__": "integration_stub"}
    cwd = os.getcwd()
    os.chdir(root)
    try:
        exec(compile(m.group(1), "INTEGRATION.md:stub", "exec"), ns)
    finally:
        os.chdir(cwd)
    return ns


def test_integration_stub_matches_the_header():
    """The binding a maintainer would copy from INTEGRATION.md has the library's struct layout (round 1 shipped a stub
    one field short: the library would have read past the caller's struct)."""
    ns = integration_stub_namespace()
    stub, ours = ns["ForwardArgs"], _lib.ForwardArgs
    assert ctypes.sizeof(stub) == ctypes.sizeof(ours)
    assert [(n, getattr(stub, n).offset, getattr(stub, n).size) for n, _ in stub._fields_] == \
           [(n, getattr(ours, n).offset, getattr(ours, n).size) for n, _ in ours._fields_]
    assert ctypes.sizeof(ns["Layout"]) == ctypes.sizeof(_lib.Layout)
    # a zero-initialised stub struct without struct_size is rejected, not read out of bounds
    a = stub()
    assert ns["lib"].ccvsq_quantize_forward(ctypes.byref(a), None) == -1
