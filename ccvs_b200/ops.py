"""Tensor-level wrappers over the C ABI (include/ccvsq.h).

PyTorch is plumbing here: it owns device memory and the CUDA stream; every computation on the
quantizer path runs in libccvsq's hand-written sm_100a kernels.  Each function passes raw
`data_ptr()`s and the current stream and returns torch tensors allocated by the caching allocator.
There is no CPU path: tensors must live on a CUDA device.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Optional

import torch

from . import _lib
from ._lib import Layout

_L = _lib.load()  # fail loudly at import time if the extension is missing

# Screening margin (see ccvsq_screen).  A code survives the BF16 screen if its score is within
#     margin_n = tau * 2 * (||z_n - bf16(z_n)|| * max||e|| * (1 + 2^-8) + ||z_n|| * max_k||e_k - bf16(e_k)||) + 2^-13 ||z_n|| max||e||
# of the row maximum.  With tau = 1 this is a PROVEN bound on how far the FP32 winner can trail the BF16 maximum:
# score errors are sum_j (dz_j e_kj + z_j de_kj + dz_j de_kj), so two scores differ from their exact values by at most
# ||dz|| ||e_A - e_B|| + ||z|| (||de_A|| + ||de_B||) + second order (Cauchy-Schwarz), and the last term covers the FP32
# accumulation of the tensor core.  The rounding-error norms are MEASURED (per row by the loader warps, per codebook by
# ccvsq_prepare_codebook), so dense latents pay ~1.4 x 2^-8 ||z|| max||e|| while operands sitting on BF16 rounding
# midpoints (tests/test_gpu_hardening.py::test_adversarial_margin) automatically get up to 4 x 2^-8 ||z|| max||e||.
DEFAULT_MARGIN_TAU = 1.0

# kernels enqueued by each entry point (ccvsq_prepare_codebook: + one 4-byte memset node)
_KERNELS_PER_CALL = {
    "ccvsq_prepare_codebook": 1, "ccvsq_search_exact": 1, "ccvsq_screen": 1, "ccvsq_screen_trace": 1, "ccvsq_screen_debug": 1, "ccvsq_rescore": 1,
    "ccvsq_search_exact_rows": 1, "ccvsq_assign": 1, "ccvsq_gather": 1, "ccvsq_backward_dz": 1, "ccvsq_code_stats": 1,
    "ccvsq_finalize": 1, "ccvsq_ema_update": 2, "ccvsq_ema_update_packed": 2, "ccvsq_code_stats_fixed": 3,
    "ccvsq_assign_normalized": 1, "ccvsq_backward_normalized": 1,
    "ccvsq_gather_add": 1, "ccvsq_polyak": 1, "ccvsq_peer_publish": 1, "ccvsq_peer_ema_update": 2,
    "ccvsq_encoder_tail_prepare": 1, "ccvsq_encoder_tail": 1,
    "ccvsq_quantize_forward": 0, "ccvsq_quantize_backward": 0,   # composites: counted by their wrappers
}


class Profiler:
    """Launch counter + optional per-call CUDA-event timing (events recorded on the stream the
    kernels are launched on).  Used by bench.py for `gpu_launches` and the live roofline number."""

    def __init__(self):
        self.launches = 0
        self.timing = False
        self.records = []      # (entry point, start event, end event)

    def reset(self, timing=False):
        """timing: False | True (every entry point) | a set of entry-point names."""
        self.launches = 0
        self.timing = timing
        self.records = []

    def summary(self):
        """{entry point: (calls, total ms)} — call after torch.cuda.synchronize()."""
        out = {}
        for name, s, e in self.records:
            c, t = out.get(name, (0, 0.0))
            out[name] = (c + 1, t + s.elapsed_time(e))
        return out


PROFILER = Profiler()


class _StreamArg:
    """The stream argument of an entry point plus the device it belongs to: `_call` makes that device current for
    the duration of the C call when it is not already (kernel launches, cudaFuncSetAttribute and cudaGetDevice inside
    the library act on the thread's CURRENT device, while pointers and stream come from the tensor's device)."""

    __slots__ = ("ptr", "index")

    def __init__(self, ptr, index):
        self.ptr = ptr
        self.index = index


def _call(name: str, *args) -> None:
    fn = getattr(_L, name)
    if args and isinstance(args[-1], _StreamArg):
        st = args[-1]
        args = args[:-1] + (st.ptr,)
        if st.index is not None and st.index != torch.cuda.current_device():
            with torch.cuda.device(st.index):
                return _call(name, *args)
    if PROFILER.timing is True or (PROFILER.timing and name in PROFILER.timing):
        s = torch.cuda.Event(enable_timing=True)
        e = torch.cuda.Event(enable_timing=True)
        s.record()
        rc = fn(*args)
        e.record()
        PROFILER.records.append((name, s, e))
    else:
        rc = fn(*args)
    PROFILER.launches += _KERNELS_PER_CALL[name]
    _lib.check(rc, name)


def timed(name: str, fn):
    """Bracket a composite wrapper with events when the profiler is timing everything."""
    if PROFILER.timing is not True:
        return fn()
    s = torch.cuda.Event(enable_timing=True)
    e = torch.cuda.Event(enable_timing=True)
    s.record()
    out = fn()
    e.record()
    PROFILER.records.append((name, s, e))
    return out


def _ptr(t: Optional[torch.Tensor]) -> ctypes.c_void_p:
    if t is None:
        return ctypes.c_void_p(0)
    return ctypes.c_void_p(t.data_ptr())


def _stream(dev: torch.device) -> _StreamArg:
    dev = torch.device(dev)
    return _StreamArg(ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream), dev.index)


def _req(t: torch.Tensor, dtype: torch.dtype, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(
            f"{name} is on {t.device}: the B200 quantizer path has no CPU fallback (CUDA tensors only)"
        )
    if t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")
    return t


def layout_of(shape, e_dim: int, mult: int = 1) -> Layout:
    """Layout of the reference's flatten (quantize.py:40-42) without the copy.

    ndim >= 4: [..., C, h, w] -> G = prod(leading), C, S = h*w, rows ordered (g, h, w, m).
    ndim <  4: the reference does `z.view(-1, e_dim)` on the tensor as it is -> rows are contiguous.
    """
    shape = tuple(int(s) for s in shape)
    numel = 1
    for s in shape:
        numel *= s
    if len(shape) >= 4:
        C, S = shape[-3], shape[-2] * shape[-1]
        if C != e_dim * mult:
            raise ValueError(f"channel dim {C} != e_dim*mult = {e_dim}*{mult}")
        return Layout(numel // (C * S), C, S, mult)
    if numel % e_dim != 0:
        raise ValueError(f"numel {numel} not divisible by e_dim {e_dim}")
    return Layout(numel // e_dim, e_dim, 1, 1)


def rows_layout(n_rows: int, dim: int) -> Layout:
    return Layout(n_rows, dim, 1, 1)


# ------------------------------------------------------------------------------------------------
# codebook preparation
# ------------------------------------------------------------------------------------------------
@dataclass
class PreparedCodebook:
    """Per-codebook-version side data (quantize.py:46 recomputes ||e||^2 on every call)."""

    weight: torch.Tensor          # [K, D] fp32 (the live parameter's storage)
    e_sq: torch.Tensor            # [K] fp32
    e_bf16: Optional[torch.Tensor]  # [codebook_rows(K), D + 16] bf16: codes | bias split (hi, mid, lo) | zeros
    e_max: Optional[torch.Tensor]   # [2] fp32 = max ||e||, max ||e - bf16(e)||
    version: int
    ptr: int

    @property
    def K(self) -> int:
        return self.weight.shape[0]

    @property
    def D(self) -> int:
        return self.weight.shape[1]


def tensor_path_supported(K: int, D: int) -> bool:
    return D % 64 == 0 and 64 <= D <= 512


def codebook_rows(K: int) -> int:
    return int(_L.ccvsq_codebook_rows(int(K)))


def prepare_codebook(weight: torch.Tensor, with_bf16: Optional[bool] = None) -> PreparedCodebook:
    w = _req(weight.detach(), torch.float32, "codebook")
    K, D = w.shape
    if with_bf16 is None:
        with_bf16 = tensor_path_supported(K, D)
    dev = w.device
    e_sq = torch.empty(K, dtype=torch.float32, device=dev)
    e_bf16 = e_max = None
    if with_bf16:
        e_bf16 = torch.empty(codebook_rows(K), D + 16, dtype=torch.bfloat16, device=dev)
        e_max = torch.empty(2, dtype=torch.float32, device=dev)
    _call("ccvsq_prepare_codebook", _ptr(w), K, D, _ptr(e_sq), _ptr(e_bf16), _ptr(e_max), _stream(dev))
    return PreparedCodebook(w, e_sq, e_bf16, e_max, weight._version, w.data_ptr())


# ------------------------------------------------------------------------------------------------
# nearest-code search
# ------------------------------------------------------------------------------------------------
def search_exact(z: torch.Tensor, lay: Layout, cb: PreparedCodebook) -> torch.Tensor:
    """FP32 CUDA-core search (quantize.py:45-50). Returns int64 [N]."""
    _req(z, torch.float32, "z")
    idx = torch.empty(lay.rows, dtype=torch.int64, device=z.device)
    _call("ccvsq_search_exact", _ptr(z), lay, _ptr(cb.weight), _ptr(cb.e_sq), cb.K, _ptr(idx), _stream(z.device))
    return idx


@dataclass
class ScreenQueue:
    """Rows the screen could not decide (ccvsq.h): device-side count + per-entry row / candidates / flags."""

    count: torch.Tensor   # [1] int32
    rows: torch.Tensor    # [N] int32
    cand: torch.Tensor    # [N, n_cand] int32
    flags: torch.Tensor   # [N] uint8


def _new_queue(N: int, n_cand: int, dev) -> ScreenQueue:
    return ScreenQueue(torch.zeros(1, dtype=torch.int32, device=dev), torch.empty(N, dtype=torch.int32, device=dev),
                       torch.empty(N, n_cand, dtype=torch.int32, device=dev), torch.empty(N, dtype=torch.uint8, device=dev))


def screen(z: torch.Tensor, lay: Layout, cb: PreparedCodebook, n_cand: int = 4, margin_tau: float = DEFAULT_MARGIN_TAU):
    """tcgen05 screening GEMM with fused candidate selection.  Returns (idx int64 [N], ScreenQueue)."""
    dev = z.device
    N = lay.rows
    idx = torch.empty(N, dtype=torch.int64, device=dev)
    q = _new_queue(N, n_cand, dev)
    _call("ccvsq_screen", _ptr(z), lay, _ptr(cb.e_bf16), _ptr(cb.e_max), cb.K, float(margin_tau), n_cand, _ptr(idx),
          _ptr(q.count), _ptr(q.rows), _ptr(q.cand), _ptr(q.flags), _stream(dev))
    return idx, q


def screen_trace(z: torch.Tensor, lay: Layout, cb: PreparedCodebook, n_cand: int = 4, margin_tau: float = DEFAULT_MARGIN_TAU):
    """ccvsq_screen with the pipeline timeline (see ccvsq.h): returns (idx, queue, trace int64 [148, 32, 8])."""
    dev = z.device
    N = lay.rows
    idx = torch.empty(N, dtype=torch.int64, device=dev)
    q = _new_queue(N, n_cand, dev)
    trace = torch.zeros(148, 32, 8, dtype=torch.int64, device=dev)
    _call("ccvsq_screen_trace", _ptr(z), lay, _ptr(cb.e_bf16), _ptr(cb.e_max), cb.K, float(margin_tau), n_cand, _ptr(idx),
          _ptr(q.count), _ptr(q.rows), _ptr(q.cand), _ptr(q.flags), _ptr(trace), _stream(dev))
    return idx, q, trace


@dataclass
class ScreenDebug:
    idx: torch.Tensor         # [N] int64
    queue: ScreenQueue
    cand_idx: torch.Tensor    # [N, n_cand] int32, -1 padded, sorted by (score desc, code asc)
    cand_score: torch.Tensor  # [N, n_cand] fp32, -inf padded
    flags: torch.Tensor       # [N] uint8
    margin: torch.Tensor      # [N] fp32
    scores: Optional[torch.Tensor]   # [N, codebook_rows(K)] fp32


def screen_debug(z: torch.Tensor, lay: Layout, cb: PreparedCodebook, n_cand: int = 4, margin_tau: float = DEFAULT_MARGIN_TAU,
                 cta_group: int = 2, dump_scores: bool = False) -> ScreenDebug:
    """Diagnostic variant: candidate lists / margins for every row, optionally the full score matrix."""
    dev = z.device
    N = lay.rows
    idx = torch.empty(N, dtype=torch.int64, device=dev)
    q = _new_queue(N, n_cand, dev)
    cand_idx = torch.empty(N, n_cand, dtype=torch.int32, device=dev)
    cand_score = torch.empty(N, n_cand, dtype=torch.float32, device=dev)
    flags = torch.empty(N, dtype=torch.uint8, device=dev)
    margin = torch.empty(N, dtype=torch.float32, device=dev)
    scores = torch.full((N, cb.e_bf16.shape[0]), float("nan"), dtype=torch.float32, device=dev) if dump_scores else None
    _call("ccvsq_screen_debug", _ptr(z), lay, _ptr(cb.e_bf16), _ptr(cb.e_max), cb.K, float(margin_tau), n_cand, cta_group,
          _ptr(idx), _ptr(q.count), _ptr(q.rows), _ptr(q.cand), _ptr(q.flags), _ptr(cand_idx), _ptr(cand_score), _ptr(flags),
          _ptr(margin), _ptr(scores), _stream(dev))
    return ScreenDebug(idx, q, cand_idx, cand_score, flags, margin, scores)


def rescore(z: torch.Tensor, lay: Layout, cb: PreparedCodebook, idx: torch.Tensor, q: ScreenQueue,
            exact_fallback: bool = True, fallback_capacity: Optional[int] = None) -> torch.Tensor:
    """FP32 re-scoring of the queued rows (in place on idx); flagged rows go to the exact kernel."""
    dev = z.device
    N = lay.rows
    n_cand = q.cand.shape[1]
    fb_rows = fb_count = None
    cap = 0
    if exact_fallback:
        cap = N if fallback_capacity is None else min(N, fallback_capacity)   # default: room for every row
        fb_rows = torch.empty(2 * cap, dtype=torch.int64, device=dev)     # rows | packed keys
        fb_count = torch.zeros(2, dtype=torch.int32, device=dev)          # queued rows, scratch counter
    _call("ccvsq_rescore", _ptr(z), lay, _ptr(cb.weight), _ptr(cb.e_sq), cb.K, n_cand, _ptr(q.count), _ptr(q.rows),
          _ptr(q.cand), _ptr(q.flags), _ptr(idx), _ptr(fb_rows), _ptr(fb_count), cap, _stream(dev))
    if exact_fallback:
        # rows with more codes inside the margin than candidate slots: exact FP32 search, count read
        # on the device (no host sync)
        _call("ccvsq_search_exact_rows", _ptr(z), lay, _ptr(cb.weight), _ptr(cb.e_sq), cb.K, _ptr(fb_rows),
              _ptr(fb_count), cap, _ptr(idx), _stream(dev))
    return idx


def search(z: torch.Tensor, lay: Layout, cb: PreparedCodebook, mode: str = "auto", n_cand: int = 4,
           margin_tau: float = DEFAULT_MARGIN_TAU, exact_fallback: bool = True) -> torch.Tensor:
    """Nearest-code indices, int64 [N].  mode: 'auto' | 'tensor' | 'exact'."""
    _req(z, torch.float32, "z")
    if mode not in ("auto", "tensor", "exact"):
        raise ValueError(f"unknown search mode {mode!r}")
    use_tensor = mode == "tensor" or (
        mode == "auto" and cb.e_bf16 is not None and tensor_path_supported(cb.K, cb.D) and lay.rows >= 128 and cb.K >= 64
    )
    if not use_tensor:
        return search_exact(z, lay, cb)
    if cb.e_bf16 is None:
        raise RuntimeError("tensor search needs a codebook prepared with with_bf16=True")
    idx, q = screen(z, lay, cb, n_cand, margin_tau)
    return rescore(z, lay, cb, idx, q, exact_fallback)


# ------------------------------------------------------------------------------------------------
# assignment / decode / backward / statistics
# ------------------------------------------------------------------------------------------------
def assign(z: torch.Tensor, lay: Layout, weight: torch.Tensor, idx: torch.Tensor, want_zq: bool = True,
           want_counts: bool = True):
    """z_q (STE forward value), sum of squared errors (fp64 [1]) and per-code counts (int32 [K])."""
    _req(z, torch.float32, "z")
    _req(idx, torch.int64, "idx")
    dev = z.device
    K = weight.shape[0]
    zq = torch.empty_like(z) if want_zq else None
    sq = torch.zeros(1, dtype=torch.float64, device=dev)
    counts = torch.zeros(K, dtype=torch.int32, device=dev) if want_counts else None
    _call("ccvsq_assign", _ptr(z), lay, _ptr(weight), K, _ptr(idx), _ptr(zq), _ptr(sq), _ptr(counts), _stream(dev))
    return zq, sq, counts


def assign_normalized(z: torch.Tensor, lay: Layout, weight: torch.Tensor, idx: torch.Tensor):
    """`assign` for normalize=True (quantize.py:56-57, any mult): the quantized vector of a position is the concatenation of
    its code rows divided by its L2 norm over all channels.  Returns z_q, sum of squared errors (fp64 [1]), counts."""
    _req(z, torch.float32, "z")
    _req(idx, torch.int64, "idx")
    dev = z.device
    K = weight.shape[0]
    zq = torch.empty_like(z)
    sq = torch.zeros(1, dtype=torch.float64, device=dev)
    counts = torch.zeros(K, dtype=torch.int32, device=dev)
    _call("ccvsq_assign_normalized", _ptr(z), lay, _ptr(weight), K, _ptr(idx), _ptr(zq), _ptr(sq), _ptr(counts), _stream(dev))
    return zq, sq, counts


def backward_normalized(z: torch.Tensor, lay: Layout, weight: torch.Tensor, idx: torch.Tensor, g_zq: Optional[torch.Tensor],
                        g_loss: torch.Tensor, beta: float, want_dz: bool = True, want_dE: bool = True):
    """Backward of the normalize=True forward: dz = g_zq + (2 g/M)(z - u) and dE through the normalisation
    (ccvsq_backward_normalized + ccvsq_finalize)."""
    dev = z.device
    K, D = weight.shape
    dz = torch.empty_like(z) if want_dz else None
    resid = torch.empty(K, D, dtype=torch.float32, device=dev) if want_dE else None
    _call("ccvsq_backward_normalized", _ptr(z), lay, _ptr(weight), K, _ptr(idx), _ptr(g_zq), _ptr(g_loss), _ptr(dz), _ptr(resid),
          _stream(dev))
    dE = None
    if want_dE:
        dE, _, _ = finalize(K, D, float(z.numel()), float(lay.rows), beta, resid=resid, g_loss=g_loss, want_dE=True)
    return dz, dE


def gather(code: torch.Tensor, weight: torch.Tensor, out_lay: Optional[Layout] = None,
           err: Optional[torch.Tensor] = None) -> torch.Tensor:
    """embed_code (quantize.py:76-83): E[code].  With a channel-major `out_lay` (S>1) the result is
    written directly as [G, C, S] (the decoder's layout).  `err` (int32 [1], zeroed by the caller) receives a
    sticky 1 if a code lies outside [0, K); a fresh flag is allocated when it is not given."""
    _req(code, torch.int64, "code")
    w = _req(weight.detach(), torch.float32, "codebook")
    K, D = w.shape
    dev = w.device
    n = code.numel()
    if out_lay is None:
        out_lay = rows_layout(n, D)
        out = torch.empty(*code.shape, D, dtype=torch.float32, device=dev)
    else:
        if out_lay.rows != n or out_lay.dim != D:
            raise ValueError("out_lay does not match code/codebook shape")
        out = torch.empty(out_lay.G, out_lay.C, out_lay.S, dtype=torch.float32, device=dev)
    if err is None:
        err = torch.zeros(1, dtype=torch.int32, device=dev)
    if n == 0:
        return out, err
    _call("ccvsq_gather", _ptr(code), _ptr(w), K, out_lay, _ptr(out), _ptr(err), _stream(dev))
    return out, err


def gather_add(code: torch.Tensor, table: torch.Tensor, pos: torch.Tensor) -> torch.Tensor:
    """tok_emb(idx) + pos_emb of the prior (mingpt.py:234-236) in one pass: code [B, T] int64, table [K, D],
    pos [1, >=T, D] or [>=T, D] -> [B, T, D] with out[b, t] = table[code[b, t]] + pos[t]."""
    _req(code, torch.int64, "code")
    w = _req(table.detach(), torch.float32, "table")
    K, D = w.shape
    B, T = code.shape
    p2 = pos.detach().reshape(-1, D)[:T]
    p2 = _req(p2.contiguous(), torch.float32, "pos")
    if p2.shape[0] < T:
        raise ValueError(f"position table has {p2.shape[0]} rows, need {T}")
    out = torch.empty(B, T, D, dtype=torch.float32, device=w.device)
    err = torch.zeros(1, dtype=torch.int32, device=w.device)
    _call("ccvsq_gather_add", _ptr(code), _ptr(w), K, D, B * T, _ptr(p2), T, _ptr(out), _ptr(err), _stream(w.device))
    return out, err


def polyak(ema: torch.Tensor, live: torch.Tensor, decay: float) -> None:
    """ema <- decay*ema + (1-decay)*live in place (quantized_video_model.py:951-964), one launch."""
    _req(ema, torch.float32, "ema")
    _req(live, torch.float32, "live")
    if ema.shape != live.shape:
        raise ValueError("shape mismatch")
    _call("ccvsq_polyak", _ptr(ema), _ptr(live), ema.numel(), float(decay), _stream(ema.device))


def backward_dz(z, lay: Layout, weight, idx, g_zq: Optional[torch.Tensor], g_loss: torch.Tensor) -> torch.Tensor:
    dev = z.device
    dz = torch.empty_like(z)
    _call("ccvsq_backward_dz", _ptr(z), lay, _ptr(weight), weight.shape[0], _ptr(idx), _ptr(g_zq), _ptr(g_loss), _ptr(dz), _stream(dev))
    return dz


def code_stats(x: torch.Tensor, lay: Layout, weight: Optional[torch.Tensor], K: int, idx: torch.Tensor,
               sub: float = 1.0, want_counts: bool = True):
    """resid[k] = sum_{idx=k} (x - sub*E[k]) (fp32 [K, D]) and counts (int32 [K])."""
    dev = x.device
    resid = torch.zeros(K, lay.dim, dtype=torch.float32, device=dev)
    counts = torch.zeros(K, dtype=torch.int32, device=dev) if want_counts else None
    _call("ccvsq_code_stats", _ptr(x), lay, _ptr(weight), K, _ptr(idx), float(sub), _ptr(resid), _ptr(counts), _stream(dev))
    return resid, counts


def code_stats_fixed(x: torch.Tensor, lay: Layout, weight: Optional[torch.Tensor], K: int, idx: torch.Tensor,
                     sub: float = 1.0, want_counts: bool = False, out: Optional[torch.Tensor] = None):
    """Deterministic `code_stats`: 64-bit fixed-point accumulation (order-independent, run-to-run bit-identical)."""
    dev = x.device
    D = lay.dim
    acc = torch.empty(K * D, dtype=torch.int64, device=dev)
    amax = torch.empty(1, dtype=torch.float32, device=dev)
    resid = out if out is not None else torch.empty(K, D, dtype=torch.float32, device=dev)
    counts = torch.zeros(K, dtype=torch.int32, device=dev) if want_counts else None
    _call("ccvsq_code_stats_fixed", _ptr(x), lay, _ptr(weight), K, _ptr(idx), float(sub), _ptr(acc), _ptr(amax), _ptr(resid),
          _ptr(counts), _stream(dev))
    return resid, counts


def finalize(K: int, D: int, M: float, N: float, beta: float, resid=None, counts=None, sq_err=None, g_loss=None,
             want_dE: bool = False, want_loss: bool = False, want_perplexity: bool = False):
    ref = next(t for t in (resid, counts, sq_err, g_loss) if t is not None)
    dev = ref.device
    dE = torch.empty(K, D, dtype=torch.float32, device=dev) if want_dE else None
    loss = torch.empty((), dtype=torch.float32, device=dev) if want_loss else None
    perp = torch.empty((), dtype=torch.float32, device=dev) if want_perplexity else None
    _call("ccvsq_finalize", _ptr(resid), _ptr(counts), _ptr(sq_err), _ptr(g_loss), K, D, float(M), float(N), float(beta),
                          _ptr(dE), _ptr(loss), _ptr(perp), _stream(dev))
    return dE, loss, perp


# ------------------------------------------------------------------------------------------------
# whole-op composites: one FFI crossing per forward / backward
# ------------------------------------------------------------------------------------------------
def fast_stream_layout(lay: Layout) -> bool:
    """Mirror of stream_fast_supported() for torch-allocated (16-byte aligned) tensors."""
    D = lay.dim
    if D % 4:
        return False
    if lay.S == 1:
        return True
    if lay.S % 4 or lay.C % 32:
        return False
    nslab = (lay.C + 255) // 256
    return lay.C % nslab == 0 and (lay.C // nslab) % 32 == 0


def uses_tensor_path(mode: str, K: int, D: int, N: int) -> bool:
    return mode == "tensor" or (mode == "auto" and tensor_path_supported(K, D) and N >= 128 and K >= 64)


@dataclass
class ForwardOut:
    idx: torch.Tensor                     # int64 [N]
    zq: Optional[torch.Tensor]            # z's shape: fl(z + fl(E[idx] - z))
    loss: Optional[torch.Tensor]          # 0-d fp32
    perplexity: Optional[torch.Tensor]    # 0-d fp32
    counts: Optional[torch.Tensor]        # int32 [K]
    resid: Optional[torch.Tensor] = None  # fp32 [K, D]: sum_{idx=k} (z - E[k]) (want_resid)


def quantize_forward(z: torch.Tensor, lay: Layout, weight: torch.Tensor, beta: float, mode: str = "auto", n_cand: int = 4,
                     margin_tau: float = DEFAULT_MARGIN_TAU, exact_fallback: bool = True, cb: Optional[PreparedCodebook] = None,
                     indices_only: bool = False, want_resid: bool = False,
                     resid_out: Optional[torch.Tensor] = None, counts_f32_out: Optional[torch.Tensor] = None) -> ForwardOut:
    """The whole forward of quantize.py:32-74 in one call of the C ABI (ccvsq_quantize_forward).
    `cb` = cached codebook side data (frozen codebook); None rebuilds it inside the call."""
    _req(z, torch.float32, "z")
    w = _req(weight.detach(), torch.float32, "codebook")
    if mode not in _lib.SEARCH_MODES:
        raise ValueError(f"unknown search mode {mode!r}")
    dev = z.device
    K, D = w.shape
    N = lay.rows
    if lay.dim != D:
        raise ValueError(f"latent dim {lay.dim} != codebook dim {D}")
    m = _lib.SEARCH_MODES[mode]
    tensor = uses_tensor_path(mode, K, D, N)
    ws_bytes = int(_L.ccvsq_forward_workspace_bytes(N, K, D, m, n_cand, 0 if cb is not None else 1))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    hdr = torch.empty(_lib.HEADER_INTS + K, dtype=torch.int32, device=dev)
    idx = torch.empty(N, dtype=torch.int64, device=dev)
    zq = loss = perp = None
    if not indices_only:
        zq = torch.empty_like(z)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        perp = torch.empty((), dtype=torch.float32, device=dev)
    a = _lib.ForwardArgs()
    a.z, a.lay, a.E, a.K, a.beta = z.data_ptr(), lay, w.data_ptr(), K, float(beta)
    a.search_mode, a.n_cand, a.margin_tau = m, int(n_cand), float(margin_tau)
    a.exact_fallback, a.indices_only, a.prepare = int(exact_fallback), int(indices_only), 0
    if cb is not None:
        a.e_sq = cb.e_sq.data_ptr()
        a.E_bf16 = cb.e_bf16.data_ptr() if cb.e_bf16 is not None else None
        a.e_max = cb.e_max.data_ptr() if cb.e_max is not None else None
    a.header, a.workspace, a.workspace_bytes = hdr.data_ptr(), ws.data_ptr(), ws_bytes
    a.idx = idx.data_ptr()
    resid = None
    if not indices_only:
        a.zq, a.loss, a.perplexity = zq.data_ptr(), loss.data_ptr(), perp.data_ptr()
        if resid_out is not None:          # caller's buffer (e.g. a view of the packed all-reduce buffer)
            resid = _req(resid_out, torch.float32, "resid_out")
            if tuple(resid.shape) != (K, D):
                raise ValueError(f"resid_out must be [{K}, {D}]; got {tuple(resid.shape)}")
            a.resid = resid.data_ptr()
        elif want_resid:
            resid = torch.empty(K, D, dtype=torch.float32, device=dev)
            a.resid = resid.data_ptr()
        if counts_f32_out is not None:     # fp32 copy of the usage counts (tail of the packed all-reduce buffer)
            a.counts_f32 = _req(counts_f32_out, torch.float32, "counts_f32_out").data_ptr()
    name = "ccvsq_screen" if tensor else "ccvsq_search_exact"
    timed = PROFILER.timing is True or (PROFILER.timing and name in PROFILER.timing)
    if timed:   # the dominant search kernel is bracketed by events recorded inside the C call
        s = torch.cuda.Event(enable_timing=True)
        e = torch.cuda.Event(enable_timing=True)
        s.record()
        e.record()   # (instantiates the handles; re-recorded by the library around the kernel)
        a.ev_search_begin, a.ev_search_end = s.cuda_event, e.cuda_event
        PROFILER.records.append((name, s, e))
    _call("ccvsq_quantize_forward", ctypes.byref(a), _stream(dev))
    n = (1 if cb is None else 0) + ((2 + int(exact_fallback)) if tensor else 1)
    if not indices_only:
        n += 1 if fast_stream_layout(lay) else 2
    PROFILER.launches += n
    counts = None if indices_only else hdr[_lib.HEADER_INTS:]
    return ForwardOut(idx, zq, loss, perp, counts, resid)


def quantize_backward(z: torch.Tensor, lay: Layout, weight: torch.Tensor, idx: torch.Tensor,
                      g_zq: Optional[torch.Tensor], g_loss: torch.Tensor, beta: float, want_dz: bool = True,
                      want_dE: bool = True, deterministic: bool = False):
    """Autograd backward of quantize.py:55-64 in one pass over z (ccvsq_quantize_backward):
    dz = g_zq + (2 g/M)(z - E[idx]);  dE = -(2 beta g/M) sum_{idx=k}(z - E[k])."""
    dev = z.device
    K, D = weight.shape
    if deterministic and want_dE:
        # dz is a pure elementwise map; the per-code sums take the fixed-point route (bit-identical run to run)
        dz = backward_dz(z, lay, weight, idx, g_zq, g_loss) if want_dz else None
        resid, _ = code_stats_fixed(z, lay, weight, K, idx, sub=1.0)
        dE, _, _ = finalize(K, D, float(z.numel()), float(lay.rows), beta, resid=resid, g_loss=g_loss, want_dE=True)
        return dz, dE
    dz = torch.empty_like(z) if want_dz else None
    dE = torch.empty(K, D, dtype=torch.float32, device=dev) if want_dE else None   # doubles as the resid scratch
    _call("ccvsq_quantize_backward", _ptr(z), lay, _ptr(weight), K, _ptr(idx), _ptr(g_zq), _ptr(g_loss), float(beta),
          _ptr(dz), _ptr(dE), _ptr(dE), _stream(dev))
    fused = fast_stream_layout(lay)
    PROFILER.launches += (1 if fused or not (want_dz and want_dE) else 2) * int(want_dz or want_dE) + int(want_dE)
    return dz, dE


def ema_update_packed(weight: torch.Tensor, n_ema: torch.Tensor, sum_ema: torch.Tensor, packed: torch.Tensor,
                      decay: float, eps: float) -> None:
    """`ema_update` straight from the packed statistics buffer [resid: K*D | counts: K] fp32 (after the all-reduce)."""
    K, D = weight.shape
    scratch = torch.empty(1, dtype=torch.float32, device=weight.device)
    _call("ccvsq_ema_update_packed", _ptr(weight), _ptr(n_ema), _ptr(sum_ema), _ptr(packed), K, D, float(decay), float(eps),
          _ptr(scratch), _stream(weight.device))


def ema_update(weight: torch.Tensor, n_ema: torch.Tensor, sum_ema: torch.Tensor, resid: torch.Tensor,
               counts: torch.Tensor, decay: float, eps: float) -> None:
    K, D = weight.shape
    scratch = torch.empty(1, dtype=torch.float32, device=weight.device)
    _call("ccvsq_ema_update", _ptr(weight), _ptr(n_ema), _ptr(sum_ema), _ptr(resid), _ptr(counts), K, D, float(decay),
                            float(eps), _ptr(scratch), _stream(weight.device))


# ------------------------------------------------------------------------------------------------
# encoder tail: 1x1 EqualConv2d + bias + LeakyReLU (+ L2 normalisation) producing the latents
# ------------------------------------------------------------------------------------------------
def encoder_tail_supported(c_in: int, c_out: int) -> bool:
    return c_in >= 64 and c_in % 64 == 0 and c_out >= 16 and c_out % 16 == 0 and (c_out <= 256 or c_out % 256 == 0)


def encoder_tail_prepare(weight: torch.Tensor, scale: float) -> torch.Tensor:
    """weight [C_out, C_in(,1,1)] fp32 -> the three BF16 terms of fl(weight * scale), [3, C_out, C_in]."""
    w = _req(weight.detach().reshape(weight.shape[0], -1).contiguous(), torch.float32, "weight")
    c_out, c_in = w.shape
    terms = torch.empty(3, c_out, c_in, dtype=torch.bfloat16, device=w.device)
    _call("ccvsq_encoder_tail_prepare", _ptr(w), c_out, c_in, float(scale), _ptr(terms), _stream(w.device))
    return terms


def encoder_tail(x: torch.Tensor, terms: torch.Tensor, bias: Optional[torch.Tensor], negative_slope: float = 0.1,
                 normalize: bool = False) -> torch.Tensor:
    """x [G, C_in, h, w] (or [G, C_in, S]) fp32 -> z of the same leading / spatial shape with C_out channels."""
    _req(x, torch.float32, "x")
    _, c_out, c_in = terms.shape
    if x.shape[1] != c_in:
        raise ValueError(f"x has {x.shape[1]} channels, the weight expects {c_in}")
    G = x.shape[0]
    S = x[0, 0].numel()
    z = torch.empty((G, c_out) + tuple(x.shape[2:]), dtype=torch.float32, device=x.device)
    if x.numel() == 0:
        return z
    b = None if bias is None else _req(bias.detach(), torch.float32, "bias")
    _call("ccvsq_encoder_tail", _ptr(x), G, c_in, S, _ptr(terms), _ptr(b), c_out, float(negative_slope), int(normalize), _ptr(z),
          _stream(x.device))
    if normalize:
        PROFILER.launches += 1
    return z
