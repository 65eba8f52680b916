"""ctypes binding of libccvsq.so (the C ABI declared in include/ccvsq.h).

The library is built in-tree (ccvs_b200/lib/libccvsq.so) by `make -C ccvs_b200/csrc` — see
`build()`.  There is deliberately NO fallback: if the shared object is missing or a symbol does not
resolve, importing `ccvs_b200.ops` raises.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import POINTER, Structure, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_uint32, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CCVSQ_LIB") or os.path.join(_HERE, "lib", "libccvsq.so")   # (override: A/B runs of two builds)
CSRC_DIR = os.path.join(_HERE, "csrc")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "ccvsq.h")

MAX_CAND = 8


class Layout(Structure):
    """struct ccvsq_layout — see include/ccvsq.h."""

    _fields_ = [("G", c_int64), ("C", c_int32), ("S", c_int32), ("mult", c_int32)]

    def __repr__(self):
        return f"Layout(G={self.G}, C={self.C}, S={self.S}, mult={self.mult})"

    @property
    def positions(self) -> int:
        return self.G * self.S

    @property
    def rows(self) -> int:
        return self.G * self.S * self.mult

    @property
    def dim(self) -> int:
        return self.C // self.mult


class ForwardArgs(Structure):
    """struct ccvsq_forward_args — see include/ccvsq.h."""

    _fields_ = [
        ("struct_size", c_uint32), ("flags", c_uint32),
        ("z", c_void_p), ("lay", Layout), ("E", c_void_p), ("K", c_int32), ("beta", c_float),
        ("search_mode", c_int32), ("n_cand", c_int32), ("margin_tau", c_float), ("exact_fallback", c_int32),
        ("indices_only", c_int32), ("prepare", c_int32),
        ("e_sq", c_void_p), ("E_bf16", c_void_p), ("e_max", c_void_p),
        ("header", c_void_p), ("workspace", c_void_p), ("workspace_bytes", c_uint64),
        ("idx", c_void_p), ("zq", c_void_p), ("loss", c_void_p), ("perplexity", c_void_p),
        ("ev_search_begin", c_void_p), ("ev_search_end", c_void_p),
        ("resid", c_void_p), ("counts_f32", c_void_p),
    ]


    def __init__(self, *args, **kw):
        super().__init__(*args, **kw)
        self.struct_size = ctypes.sizeof(ForwardArgs)     # checked by the library (stale bindings are rejected)


ABI_MAJOR = 2
HEADER_INTS = 16
SEARCH_MODES = {"auto": 0, "tensor": 1, "exact": 2}

# name -> (restype, argtypes); every symbol include/ccvsq.h declares.
_P = c_void_p
SIGNATURES = {
    "ccvsq_version": (c_int, []),
    "ccvsq_last_error": (c_char_p, []),
    "ccvsq_codebook_rows": (c_int, [c_int]),
    "ccvsq_prepare_codebook": (c_int, [_P, c_int, c_int, _P, _P, _P, _P]),
    "ccvsq_search_exact": (c_int, [_P, Layout, _P, _P, c_int, _P, _P]),
    "ccvsq_screen": (c_int, [_P, Layout, _P, _P, c_int, c_float, c_int, _P, _P, _P, _P, _P, _P]),
    "ccvsq_screen_debug": (c_int, [_P, Layout, _P, _P, c_int, c_float, c_int, c_int, _P, _P, _P, _P, _P, _P, _P, _P, _P,
                                   _P, _P]),
    "ccvsq_screen_trace": (c_int, [_P, Layout, _P, _P, c_int, c_float, c_int, _P, _P, _P, _P, _P, _P, _P]),
    "ccvsq_rescore": (c_int, [_P, Layout, _P, _P, c_int, c_int, _P, _P, _P, _P, _P, _P, _P, c_int64, _P]),
    "ccvsq_search_exact_rows": (c_int, [_P, Layout, _P, _P, c_int, _P, _P, c_int64, _P, _P]),
    "ccvsq_assign": (c_int, [_P, Layout, _P, c_int, _P, _P, _P, _P, _P]),
    "ccvsq_assign_normalized": (c_int, [_P, Layout, _P, c_int, _P, _P, _P, _P, _P]),
    "ccvsq_backward_normalized": (c_int, [_P, Layout, _P, c_int, _P, _P, _P, _P, _P, _P]),
    "ccvsq_gather": (c_int, [_P, _P, c_int, Layout, _P, _P, _P]),
    "ccvsq_backward_dz": (c_int, [_P, Layout, _P, c_int, _P, _P, _P, _P, _P]),
    "ccvsq_code_stats": (c_int, [_P, Layout, _P, c_int, _P, c_float, _P, _P, _P]),
    "ccvsq_code_stats_fixed": (c_int, [_P, Layout, _P, c_int, _P, c_float, _P, _P, _P, _P, _P]),
    "ccvsq_finalize": (c_int, [_P, _P, _P, _P, c_int, c_int, c_double, c_double, c_float, _P, _P, _P, _P]),
    "ccvsq_ema_update": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_float, c_float, _P, _P]),
    "ccvsq_ema_update_packed": (c_int, [_P, _P, _P, _P, c_int, c_int, c_float, c_float, _P, _P]),
    "ccvsq_peer_exchange_bytes": (c_uint64, [c_int, c_int, c_int]),
    "ccvsq_peer_alloc": (c_int, [c_uint64, POINTER(c_void_p), _P]),
    "ccvsq_peer_open": (c_int, [_P, POINTER(c_void_p)]),
    "ccvsq_peer_close": (c_int, [_P]),
    "ccvsq_peer_free": (c_int, [_P]),
    "ccvsq_peer_publish": (c_int, [_P, c_int, c_int, POINTER(c_void_p), c_int, c_int, _P]),
    "ccvsq_peer_ema_update": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_float, c_float, _P]),
    "ccvsq_encoder_tail_prepare": (c_int, [_P, c_int, c_int, c_float, _P, _P]),
    "ccvsq_encoder_tail": (c_int, [_P, c_int64, c_int, c_int, _P, _P, c_int, c_float, c_int, _P, _P]),
    "ccvsq_gather_add": (c_int, [_P, _P, c_int, c_int, c_int64, _P, c_int64, _P, _P, _P]),
    "ccvsq_polyak": (c_int, [_P, _P, c_int64, c_double, _P]),
    "ccvsq_forward_workspace_bytes": (c_uint64, [c_int64, c_int, c_int, c_int, c_int, c_int]),
    "ccvsq_quantize_forward": (c_int, [POINTER(ForwardArgs), _P]),
    "ccvsq_quantize_backward": (c_int, [_P, Layout, _P, c_int, _P, _P, _P, c_float, _P, _P, _P, _P]),
}

_lib = None


def build(verbose: bool = False) -> str:
    """Compile libccvsq.so for sm_100a with nvcc (cross-compiles without a GPU)."""
    proc = subprocess.run(["make", "-C", CSRC_DIR], capture_output=True, text=True)
    if verbose or proc.returncode != 0:
        print(proc.stdout)
        print(proc.stderr)
    if proc.returncode != 0:
        raise RuntimeError("building libccvsq.so failed (see output above)")
    return LIB_PATH


def load() -> ctypes.CDLL:
    """Load the shared library and bind every entry point; raises if anything is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: the CUDA extension is not built. Run "
            "`python -c 'import __graft_entry__ as g; g.build()'` (or `make -C ccvs_b200/csrc`). "
            "There is no CPU / PyTorch fallback for the quantizer path."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        if os.environ.get("CCVSQ_LIB") and not hasattr(lib, name):
            continue             # an older build under A/B comparison may lack the newest diagnostics
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.ccvsq_version() // 100 != ABI_MAJOR:
        raise RuntimeError(f"libccvsq ABI version {lib.ccvsq_version()} is not {ABI_MAJOR}.x")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().ccvsq_last_error()
        raise RuntimeError(f"{what} failed with status {rc}: {msg.decode() if msg else '?'}")
