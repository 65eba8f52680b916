"""Tensor-level wrappers over the C ABI (include/ccvsq.h).

PyTorch is plumbing here: it owns device memory and the CUDA stream; every computation on the
quantizer path runs in libccvsq's hand-written sm_100a kernels.  Each function passes raw
`data_ptr()`s and the current stream and returns torch tensors allocated by the caching allocator.
There is no CPU path: tensors must live on a CUDA device.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import Layout

_L = _lib.load()  # fail loudly at import time if the extension is missing

# kernels enqueued by each entry point (ccvsq_prepare_codebook: + one 4-byte memset node)
_KERNELS_PER_CALL = {
    "ccvsq_prepare_codebook": 1, "ccvsq_search_exact": 1, "ccvsq_pack_latents": 1, "ccvsq_screen": 1,
    "ccvsq_screen_dump": 1, "ccvsq_screen_trace": 1, "ccvsq_rescore": 1, "ccvsq_search_exact_rows": 1, "ccvsq_assign": 1, "ccvsq_gather": 1,
    "ccvsq_backward_dz": 1, "ccvsq_code_stats": 1, "ccvsq_finalize": 1, "ccvsq_ema_update": 2,
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


def _call(name: str, *args) -> None:
    fn = getattr(_L, name)
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


def _ptr(t: Optional[torch.Tensor]) -> ctypes.c_void_p:
    if t is None:
        return ctypes.c_void_p(0)
    return ctypes.c_void_p(t.data_ptr())


def _stream(dev: torch.device) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


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
    e_bf16: Optional[torch.Tensor]  # [K_pad, D] bf16
    bias: Optional[torch.Tensor]    # [K_pad] fp32 = -0.5 ||e||^2 (-inf on padding)
    e_max: Optional[torch.Tensor]   # [1] fp32 = max ||e||
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


def prepare_codebook(weight: torch.Tensor, with_bf16: Optional[bool] = None) -> PreparedCodebook:
    w = _req(weight.detach(), torch.float32, "codebook")
    K, D = w.shape
    if with_bf16 is None:
        with_bf16 = tensor_path_supported(K, D)
    dev = w.device
    e_sq = torch.empty(K, dtype=torch.float32, device=dev)
    e_bf16 = bias = e_max = None
    if with_bf16:
        K_pad = (K + 255) // 256 * 256
        e_bf16 = torch.empty(K_pad, D, dtype=torch.bfloat16, device=dev)
        bias = torch.empty(K_pad, dtype=torch.float32, device=dev)
        e_max = torch.empty(1, dtype=torch.float32, device=dev)
    _call("ccvsq_prepare_codebook", _ptr(w), K, D, _ptr(e_sq), _ptr(e_bf16), _ptr(bias), _ptr(e_max), _stream(dev))
    return PreparedCodebook(w, e_sq, e_bf16, bias, e_max, weight._version, w.data_ptr())


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
class ScreenResult:
    cand_idx: torch.Tensor    # [N, 2, n_cand] int32, -1 padded (two epilogue halves, see ccvsq.h)
    cand_score: torch.Tensor  # [N, 2, n_cand] fp32, slot 0 of each half = that half's maximum
    flags: torch.Tensor       # [N, 2] uint8, bit0 = more than n_cand codes inside the margin
    margin: torch.Tensor      # [N_pad] fp32 row margins the screen used

    def merged(self):
        """(idx [N, 2*n_cand] int32 with -1 for dead entries, score, flag [N] bool) after applying
        the global per-row threshold — what ccvsq_rescore sees (tests / diagnostics)."""
        N = self.cand_idx.shape[0]
        sc = self.cand_score.view(N, -1)
        ci = self.cand_idx.view(N, -1)
        thr = torch.maximum(self.cand_score[:, 0, 0], self.cand_score[:, 1, 0]) - self.margin[:N]
        live = (ci >= 0) & (sc >= thr.unsqueeze(1))
        nc = self.cand_idx.shape[2]
        f = self.flags.view(N, 2)
        last_live = live.view(N, 2, nc)[:, :, -1]
        max_live = self.cand_score[:, :, 0] >= thr.unsqueeze(1)
        incomplete = (((f & 1) != 0) & last_live) | (((f & 2) != 0) & max_live)
        return torch.where(live, ci, torch.full_like(ci, -1)), torch.where(live, sc, torch.full_like(sc, float("-inf"))), \
            incomplete.any(dim=1)


def pack_latents(z: torch.Tensor, lay: Layout, cb: PreparedCodebook, margin_tau: float) -> Tuple[torch.Tensor, torch.Tensor]:
    N, D = lay.rows, lay.dim
    N_pad = (N + 127) // 128 * 128
    zb = torch.empty(N_pad, D, dtype=torch.bfloat16, device=z.device)
    margin = torch.empty(N_pad, dtype=torch.float32, device=z.device)
    _call("ccvsq_pack_latents", _ptr(z), lay, _ptr(zb), _ptr(margin), float(margin_tau) * 2.0 ** -8, _ptr(cb.e_max), _stream(z.device))
    return zb, margin


def screen(zb: torch.Tensor, margin: torch.Tensor, cb: PreparedCodebook, N: int, n_cand: int = 4) -> ScreenResult:
    dev = zb.device
    cand_idx = torch.empty(N, 2, n_cand, dtype=torch.int32, device=dev)
    cand_score = torch.empty(N, 2, n_cand, dtype=torch.float32, device=dev)
    flags = torch.empty(N, 2, dtype=torch.uint8, device=dev)
    _call("ccvsq_screen", _ptr(zb), _ptr(margin), _ptr(cb.e_bf16), _ptr(cb.bias), N, cb.K, cb.D, n_cand,
          _ptr(cand_idx), _ptr(cand_score), _ptr(flags), _stream(dev))
    return ScreenResult(cand_idx, cand_score, flags, margin)


def screen_dump(zb: torch.Tensor, margin: torch.Tensor, cb: PreparedCodebook, N: int, n_cand: int = 4):
    """Diagnostic: screen + the full fp32 score matrix [N_pad, K_pad] (tests / debugging only)."""
    dev = zb.device
    cand_idx = torch.empty(N, 2, n_cand, dtype=torch.int32, device=dev)
    cand_score = torch.empty(N, 2, n_cand, dtype=torch.float32, device=dev)
    flags = torch.empty(N, 2, dtype=torch.uint8, device=dev)
    scores = torch.full((zb.shape[0], cb.e_bf16.shape[0]), float("nan"), dtype=torch.float32, device=dev)
    _call("ccvsq_screen_dump", _ptr(zb), _ptr(margin), _ptr(cb.e_bf16), _ptr(cb.bias), N, cb.K, cb.D, n_cand,
          _ptr(cand_idx), _ptr(cand_score), _ptr(flags), _ptr(scores), _stream(dev))
    return ScreenResult(cand_idx, cand_score, flags, margin), scores


def screen_trace(zb: torch.Tensor, margin: torch.Tensor, cb: PreparedCodebook, N: int, n_cand: int = 4):
    """Diagnostic: per-role event timeline of CTA 0, int64 [4, 4000] (clock64 << 8 | event)."""
    dev = zb.device
    cand_idx = torch.empty(N, 2, n_cand, dtype=torch.int32, device=dev)
    cand_score = torch.empty(N, 2, n_cand, dtype=torch.float32, device=dev)
    flags = torch.empty(N, 2, dtype=torch.uint8, device=dev)
    trace = torch.zeros(4, 4000, dtype=torch.int64, device=dev)
    _call("ccvsq_screen_trace", _ptr(zb), _ptr(margin), _ptr(cb.e_bf16), _ptr(cb.bias), N, cb.K, cb.D, n_cand,
          _ptr(cand_idx), _ptr(cand_score), _ptr(flags), _ptr(trace), _stream(dev))
    return trace


def rescore(z: torch.Tensor, lay: Layout, cb: PreparedCodebook, sr: ScreenResult, exact_fallback: bool = True,
            fallback_capacity: int = 1 << 16) -> torch.Tensor:
    dev = z.device
    N = lay.rows
    idx = torch.empty(N, dtype=torch.int64, device=dev)
    n_cand = sr.cand_idx.shape[2]
    fb_rows = fb_count = None
    cap = 0
    if exact_fallback:
        cap = min(N, fallback_capacity)
        fb_rows = torch.empty(2 * cap, dtype=torch.int64, device=dev)     # rows | packed keys
        fb_count = torch.zeros(2, dtype=torch.int32, device=dev)          # queued rows, scratch counter
    _call("ccvsq_rescore", _ptr(z), lay, _ptr(cb.weight), _ptr(cb.e_sq), cb.K, _ptr(sr.cand_idx),
          _ptr(sr.cand_score), _ptr(sr.margin), n_cand,
                         _ptr(sr.flags), _ptr(idx), _ptr(fb_rows), _ptr(fb_count), cap, _stream(dev))
    if exact_fallback:
        # rows with more codes inside the margin than candidate slots: exact FP32 search, count read
        # on the device (no host sync)
        _call("ccvsq_search_exact_rows", _ptr(z), lay, _ptr(cb.weight), _ptr(cb.e_sq), cb.K, _ptr(fb_rows),
                                       _ptr(fb_count), cap, _ptr(idx), _stream(dev))
    return idx


def search(z: torch.Tensor, lay: Layout, cb: PreparedCodebook, mode: str = "auto", n_cand: int = 4,
           margin_tau: float = 1.0, exact_fallback: bool = True) -> torch.Tensor:
    """Nearest-code indices, int64 [N].  mode: 'auto' | 'tensor' | 'exact'."""
    _req(z, torch.float32, "z")
    use_tensor = mode == "tensor" or (
        mode == "auto" and cb.e_bf16 is not None and tensor_path_supported(cb.K, cb.D) and lay.rows >= 128 and cb.K >= 64
    )
    if mode not in ("auto", "tensor", "exact"):
        raise ValueError(f"unknown search mode {mode!r}")
    if not use_tensor:
        return search_exact(z, lay, cb)
    if cb.e_bf16 is None:
        raise RuntimeError("tensor search needs a codebook prepared with with_bf16=True")
    zb, margin = pack_latents(z, lay, cb, margin_tau)
    sr = screen(zb, margin, cb, lay.rows, n_cand)
    return rescore(z, lay, cb, sr, exact_fallback)


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


def gather(code: torch.Tensor, weight: torch.Tensor, out_lay: Optional[Layout] = None) -> torch.Tensor:
    """embed_code (quantize.py:76-83): E[code].  With a channel-major `out_lay` (S>1) the result is
    written directly as [G, C, S] (the decoder's layout)."""
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
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    if n == 0:
        return out, err
    _call("ccvsq_gather", _ptr(code), _ptr(w), K, out_lay, _ptr(out), _ptr(err), _stream(dev))
    return out, err


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


def ema_update(weight: torch.Tensor, n_ema: torch.Tensor, sum_ema: torch.Tensor, resid: torch.Tensor,
               counts: torch.Tensor, decay: float, eps: float) -> None:
    K, D = weight.shape
    scratch = torch.empty(1, dtype=torch.float32, device=weight.device)
    _call("ccvsq_ema_update", _ptr(weight), _ptr(n_ema), _ptr(sum_ema), _ptr(resid), _ptr(counts), K, D, float(decay),
                            float(eps), _ptr(scratch), _stream(weight.device))
