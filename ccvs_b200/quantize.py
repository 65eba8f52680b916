"""Drop-in `VectorQuantizer` for CCVS, B200-native.

Mirrors the reference module surface
    /root/reference/models/skip_vid_generator/modules/quantize.py:7-83
(same constructor, attributes, `forward` / `embed_code` signatures and return structure, the single
parameter `embedding.weight`) so `QVidModel`, `StateModel` and `StftModel`
(quantized_video_model.py:143, state_model.py:57, stft_model.py:59) can use it unmodified, while all
arithmetic runs in libccvsq's sm_100a kernels through the C ABI of include/ccvsq.h.

What differs by design:
  * no N x K distance matrix, no one-hot GEMM: a BF16 tcgen05 screening GEMM with a fused running
    candidate selection + FP32 re-scoring replaces quantize.py:45-55;
  * `min_encodings` (quantize.py:51-52, never consumed by any caller) is materialised lazily;
  * NCHW/NTCHW inputs are read and written in place: the two permute+contiguous copies of
    quantize.py:40-41,71-72 are folded into the kernels' addressing;
  * CUDA only.  A CPU tensor raises: there is no fallback path.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from . import dist as vq_dist
from . import ops
from ._lib import Layout


class LazyOneHot(torch.Tensor):
    """The reference's dense `min_encodings` [N, K] (quantize.py:51-52) as a tensor that is materialised on first use.

    No reference caller reads it and at the stress config it would be 128 GiB per GPU, so the forward
    returns this storage-less `torch.Tensor` subclass: `.shape`, `.dtype`, `.device`, `.size()`, `isinstance(x,
    torch.Tensor)` answer without allocating; `x.sum(0)` / `x.mean(0)` / `torch.mean(x, dim=0)` (what the
    reference itself does with it, quantize.py:67) are answered from a bincount of the indices; ANY other torch
    operation receives the real dense tensor (`.dense()`, built once by scatter and cached), so an unmodified
    caller sees exactly the reference's tensor.
    """

    @staticmethod
    def __new__(cls, indices: torch.Tensor, n_e: int, dtype: torch.dtype):
        return torch.Tensor._make_wrapper_subclass(cls, (indices.shape[0], n_e), dtype=dtype, device=indices.device,
                                                   requires_grad=False)

    def __init__(self, indices: torch.Tensor, n_e: int, dtype: torch.dtype):
        self.indices = indices
        self.n_e = n_e
        self._dense: Optional[torch.Tensor] = None

    def dense(self) -> torch.Tensor:
        if self._dense is None:
            out = torch.zeros(self.indices.shape[0], self.n_e, dtype=self.dtype, device=self.indices.device)
            out.scatter_(1, self.indices.view(-1, 1), 1)
            self._dense = out
        return self._dense

    def usage(self) -> torch.Tensor:
        """Column sums [K] (= min_encodings.sum(0)) without the dense matrix."""
        return torch.bincount(self.indices.view(-1), minlength=self.n_e).to(self.dtype)

    def __repr__(self):
        return f"LazyOneHot(shape={tuple(self.shape)}, dtype={self.dtype}, device={self.device})"

    _META = {"shape", "dtype", "device", "layout", "ndim", "requires_grad", "is_cuda", "is_leaf", "grad", "grad_fn",
             "names", "is_sparse", "is_quantized", "is_meta", "_version", "output_nr", "is_cpu"}

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        name = getattr(func, "__name__", "")
        owner = getattr(getattr(func, "__self__", None), "__name__", "")
        with torch._C.DisableTorchFunctionSubclass():
            if name == "__get__" and owner in cls._META:          # attribute reads: answered by the wrapper itself
                return func(*args, **kwargs)
            if name in ("size", "dim", "numel", "nelement", "element_size", "is_floating_point", "is_contiguous",
                        "__repr__", "__len__", "get_device") and not kwargs:
                return func(*args)
            if name in ("sum", "mean") and isinstance(args[0], LazyOneHot):
                rest = args[1:]
                dim = kwargs.get("dim", kwargs.get("axis", rest[0] if rest else None))
                extra = {k: v for k, v in kwargs.items() if k not in ("dim", "axis")}
                if dim in (0, (0,), [0], -2, (-2,)) and len(rest) <= 1 and not extra:
                    u = args[0].usage()
                    return u if name == "sum" else u / args[0].shape[0]

            def unwrap(x):
                if isinstance(x, LazyOneHot):
                    return x.dense()
                if isinstance(x, (list, tuple)):
                    return type(x)(unwrap(v) for v in x)
                return x

            return func(*unwrap(args), **{k: unwrap(v) for k, v in kwargs.items()})

    @classmethod
    def __torch_dispatch__(cls, func, types, args=(), kwargs=None):
        """Anything that reaches the dispatcher without passing __torch_function__ (C++ callers) gets the dense tensor."""
        def unwrap(x):
            if isinstance(x, LazyOneHot):
                return x.dense()
            if isinstance(x, (list, tuple)):
                return type(x)(unwrap(v) for v in x)
            return x

        return func(*unwrap(args), **{k: unwrap(v) for k, v in (kwargs or {}).items()})


class _QuantizeFn(torch.autograd.Function):
    """forward: search + assign + finalize; backward: dz kernel + per-code scatter-reduce.

    Gradients (autograd of quantize.py:60-64):
        dz = g_zq + (2 g_loss / M) (z - E[idx])
        dE = (2 beta g_loss / M) (n_k E_k - sum_{i in k} z_i)
    """

    @staticmethod
    def forward(ctx, z, weight, module):
        lay = ops.layout_of(z.shape, module.e_dim, module.mult)
        want_resid = getattr(module, "_want_resid", False)
        resid_out = getattr(module, "_resid_out", None)
        det = module.deterministic and (want_resid or resid_out is not None)
        out = ops.quantize_forward(z, lay, weight, module.beta, module.search_mode, module.n_cand, module.margin_tau,
                                   module.exact_fallback, cb=module._cb_cached(),
                                   want_resid=want_resid and not det, resid_out=None if det else resid_out,
                                   counts_f32_out=getattr(module, "_counts_f32_out", None))
        zq, loss, idx, perp, counts = out.zq, out.loss, out.idx, out.perplexity, out.counts
        if det:      # EMA statistics in fixed point (a second pass over z instead of riding on the assign pass)
            out.resid, _ = ops.code_stats_fixed(z, lay, weight.detach(), weight.shape[0], idx, sub=1.0, out=resid_out)
        module._last_resid = out.resid        # per-code residual sums for the EMA update (None unless asked for)
        ctx.lay = lay
        ctx.beta = module.beta
        ctx.module = module
        # the EMA variant rewrites the codebook in place right after forward: keep the version the
        # indices were computed with for backward
        ctx.save_for_backward(z, weight.detach().clone() if module._inplace_codebook_update else weight, idx)
        ctx.token = module._begin_deferred() if hasattr(module, "_begin_deferred") else None
        ctx.mark_non_differentiable(idx, perp, counts)
        return zq, loss, idx, perp, counts

    @staticmethod
    def backward(ctx, g_zq, g_loss, _gi, _gp, _gc):
        z, weight, idx = ctx.saved_tensors
        if ctx.token is not None:       # the codebook this forward used, if it was rewritten before this backward ran
            weight = getattr(ctx.module, "_stale_codebooks", {}).pop(ctx.token, weight)
        lay = ctx.lay
        dev = z.device
        if g_loss is None:
            g_loss = torch.zeros((), dtype=torch.float32, device=dev)
        g_loss = g_loss.to(torch.float32).contiguous()
        want_dz, want_dE = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if not (want_dz or want_dE):
            return None, None, None
        g = None if (g_zq is None or not want_dz) else g_zq.to(torch.float32).contiguous()
        dz, dE = ops.quantize_backward(z, lay, weight, idx, g, g_loss, ctx.beta, want_dz, want_dE,
                                       deterministic=ctx.module.deterministic)
        ctx.module._after_backward(ctx.token)
        return dz, dE, None


class _AssignFn(torch.autograd.Function):
    """Assign with GIVEN indices against a given table: z_q = fl(z + fl(T[idx] - z)), loss, perplexity, counts, and
    the same backward as _QuantizeFn (dz, dT).  Used by the `normalize=True` variant with mult == 1, where the
    search runs on the raw codebook (quantize.py:45-50) but everything after it sees the L2-normalised rows
    (quantize.py:56-64): T = E / ||E|| is K rows of torch ops, all N-sized work stays in the kernels."""

    @staticmethod
    def forward(ctx, z, table, idx, beta):
        lay = ops.layout_of(z.shape, table.shape[1], 1)
        K, D = table.shape
        zq, sq, counts = ops.assign(z, lay, table.detach(), idx)
        _, loss, perp = ops.finalize(K, D, float(z.numel()), float(lay.rows), beta, counts=counts, sq_err=sq,
                                     want_loss=True, want_perplexity=True)
        ctx.lay = lay
        ctx.beta = beta
        ctx.save_for_backward(z, table, idx)
        ctx.mark_non_differentiable(perp, counts)
        return zq, loss, perp, counts

    @staticmethod
    def backward(ctx, g_zq, g_loss, _gp, _gc):
        z, table, idx = ctx.saved_tensors
        dev = z.device
        if g_loss is None:
            g_loss = torch.zeros((), dtype=torch.float32, device=dev)
        g_loss = g_loss.to(torch.float32).contiguous()
        want_dz, want_dT = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if not (want_dz or want_dT):
            return None, None, None, None
        g = None if (g_zq is None or not want_dz) else g_zq.to(torch.float32).contiguous()
        dz, dT = ops.quantize_backward(z, ctx.lay, table.detach(), idx, g, g_loss, ctx.beta, want_dz, want_dT)
        return dz, dT, None, None


class _NormAssignFn(torch.autograd.Function):
    """normalize=True with mult > 1 (quantize.py:56-64): a position's quantized vector is the concatenation of its `mult`
    code rows divided by its L2 norm over ALL channels, so it is not a function of one code and no table can be
    pre-normalised.  Forward (z_q, loss, perplexity, counts) and backward (dz, dE through the normalisation) run in the
    position-tile kernels `ccvsq_assign_normalized` / `ccvsq_backward_normalized`; no N-sized torch op remains."""

    @staticmethod
    def forward(ctx, z, weight, idx, beta, mult):
        K, D = weight.shape
        if z.ndim >= 4:
            lay = ops.layout_of(z.shape, D, mult)
        else:
            # the reference normalises over the LAST dim of the tensor as it is (quantize.py:57 after `.view(z.shape)`):
            # a "position" is one row of that dim, holding last / D consecutive codes
            last = int(z.shape[-1])
            if last % D != 0:
                raise ValueError(f"normalize=True needs the last dim ({last}) to hold whole codes of {D} channels")
            lay = Layout(z.numel() // last, last, 1, last // D)
        zq, sq, counts = ops.assign_normalized(z, lay, weight.detach(), idx)
        _, loss, perp = ops.finalize(K, D, float(z.numel()), float(lay.rows), beta, counts=counts, sq_err=sq,
                                     want_loss=True, want_perplexity=True)
        ctx.lay = lay
        ctx.beta = beta
        ctx.save_for_backward(z, weight, idx)
        ctx.mark_non_differentiable(perp, counts)
        return zq, loss, perp, counts

    @staticmethod
    def backward(ctx, g_zq, g_loss, _gp, _gc):
        z, weight, idx = ctx.saved_tensors
        if g_loss is None:
            g_loss = torch.zeros((), dtype=torch.float32, device=z.device)
        g_loss = g_loss.to(torch.float32).contiguous()
        want_dz, want_dE = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if not (want_dz or want_dE):
            return None, None, None, None, None
        g = None if (g_zq is None or not want_dz) else g_zq.to(torch.float32).contiguous()
        dz, dE = ops.backward_normalized(z, ctx.lay, weight.detach(), idx, g, g_loss, ctx.beta, want_dz, want_dE)
        return dz, dE, None, None, None


class VectorQuantizer(nn.Module):
    """Discretization bottleneck of the VQ-VAE — same contract as the reference (quantize.py:7-30).

    Inputs:
    - n_e : number of embeddings
    - e_dim : dimension of embedding per position (divided by `mult` internally, quantize.py:21)
    - beta : weight of the codebook term, quantize.py:60-61 (the reference hard-codes 0.25)
    - mult : number of concatenated embeddings per position
    - normalize : L2-normalise z_q over the channel dim (quantize.py:56-57)

    B200-specific keyword-only knobs (defaults reproduce the reference's results):
    - search_mode: 'auto' (tensor-core screen + FP32 rescoring when the shape allows, exact FP32
      kernel otherwise), 'tensor', or 'exact'
    - n_cand: candidate slots per latent kept by the screen (<= 8)
    - margin_tau: multiplier of the screening margin; 1 (default) = the proven bound on the BF16 rounding error of a
      score difference, see ops.DEFAULT_MARGIN_TAU / ccvsq_screen
    - exact_fallback: rows with more than n_cand codes inside the margin are re-searched exactly
    - deterministic: accumulate the per-code sums behind `embedding.weight.grad` (and the EMA statistics) in 64-bit
      fixed point instead of FP32 atomics: run-to-run bit-identical gradients, at the cost of a second pass over z
    """

    def __init__(self, n_e, e_dim, beta, mult=1, normalize=False, *, search_mode: str = "auto", n_cand: int = 4,
                 margin_tau: float = ops.DEFAULT_MARGIN_TAU, exact_fallback: bool = True, deterministic: bool = False):
        super().__init__()
        self.n_e = n_e
        assert e_dim % mult == 0
        self.e_dim = e_dim // mult
        self.beta = beta
        self.mult = mult
        self.normalize = normalize
        self.search_mode = search_mode
        self.n_cand = n_cand
        self.margin_tau = margin_tau
        self.exact_fallback = exact_fallback
        self.deterministic = deterministic     # per-code sums (dE, EMA statistics) in 64-bit fixed point: bit-identical runs

        self.embedding = nn.Embedding(self.n_e, self.e_dim)
        if self.e_dim <= 1:
            self.embedding.weight.data.uniform_(0, 1.0)          # quantize.py:27-28
        else:
            self.embedding.weight.data.uniform_(-1.0 / self.n_e, 1.0 / self.n_e)   # quantize.py:30
        self._cb: Optional[ops.PreparedCodebook] = None
        self._inplace_codebook_update = False
        self.last_counts: Optional[torch.Tensor] = None   # int32 [K] usage of the last forward

    def _after_backward(self, token=None):
        """Called by `_QuantizeFn.backward` once its kernels are enqueued (the EMA variant completes its deferred
        statistics exchange here)."""

    # -- codebook side data (||e||^2, BF16 shadow, bias).  Rebuilt on EVERY call by default: the
    # reference's Polyak averaging writes `param.data` in place (quantized_video_model.py:962-964),
    # which does not bump the autograd version counter, so no cheap staleness test exists.  The
    # rebuild is one pass over K*D floats (microseconds).  `freeze_codebook()` pins it for
    # inference loops over a fixed codebook.
    def _cb_cached(self) -> Optional[ops.PreparedCodebook]:
        """The frozen side data if it still belongs to the live parameter, else None (= rebuild in-call)."""
        cb = self._cb
        w = self.embedding.weight
        if cb is not None and cb.ptr == w.data_ptr() and cb.weight.device == w.device and cb.version == w._version:
            return cb
        self._cb = None        # moved, or written in place through autograd-visible ops (optimizer.step, copy_, ...)
        return None

    def _prepared(self) -> ops.PreparedCodebook:
        return self._cb_cached() or ops.prepare_codebook(self.embedding.weight)

    def freeze_codebook(self):
        """Cache the codebook side data until `unfreeze_codebook()`.  The cache is dropped automatically when the
        parameter moves or its autograd version changes (optimizer steps, `copy_`, `load_state_dict`) and by
        `accumulate_from`; writes through `.data` / `.detach()` views made behind the module's back are NOT visible
        (no version bump, quantized_video_model.py:962-964): the caller promises not to do that while frozen."""
        self._cb = ops.prepare_codebook(self.embedding.weight)
        return self

    def unfreeze_codebook(self):
        self._cb = None
        return self

    def forward(self, z):
        """z [b, (t,) c, (h, w)] -> (z_q, loss, (perplexity, min_encodings, min_encoding_indices))
        exactly as quantize.py:32-74."""
        if not z.is_cuda:
            raise RuntimeError("ccvs_b200.VectorQuantizer runs on CUDA (sm_100a) only; there is no CPU fallback")
        in_dtype = z.dtype
        if in_dtype in (torch.float16, torch.bfloat16):
            # mixed-precision callers (the reference trains under apex amp, tools/engine.py:11-14): the latents are
            # upcast, the whole path runs in FP32 as specified, and z_q goes back in the caller's dtype
            z = z.float()
        elif in_dtype != torch.float32:
            raise TypeError(f"the reference quantizer is FP32 end to end; got {z.dtype}")
        z = z.contiguous()
        w = self.embedding.weight
        if z.numel() == 0:
            return self._forward_empty(z, w)
        if self.normalize:
            z_q, loss, idx, perp = self._forward_normalized(z, w)
        else:
            z_q, loss, idx, perp, counts = _QuantizeFn.apply(z, w, self)
            self.last_counts = counts
        if in_dtype != torch.float32:
            z_q = z_q.to(in_dtype)
        idx2 = idx.view(-1, 1)
        return z_q, loss, (perp, LazyOneHot(idx2, self.n_e, in_dtype), idx2)

    def _forward_empty(self, z, w):
        """Empty batch, as the reference handles it (quantize.py:40-74 on zero rows): empty z_q and indices, and the
        means over zero elements make loss and perplexity NaN; backward gives an empty dz and a zero dE."""
        nan = float("nan")
        z_q = z + torch.zeros_like(z)
        loss = (z.sum() + w.sum() * 0.0) + nan            # value NaN; d/dz empty, d/dE exactly zero
        perp = torch.full((), nan, dtype=torch.float32, device=z.device)
        idx2 = torch.empty(0, 1, dtype=torch.int64, device=z.device)
        self.last_counts = torch.zeros(self.n_e, dtype=torch.int32, device=z.device)
        return z_q, loss, (perp, LazyOneHot(idx2, self.n_e, z.dtype), idx2)

    def _forward_normalized(self, z, w):
        lay = ops.layout_of(z.shape, self.e_dim, self.mult)
        cb = self._prepared()
        with torch.no_grad():
            idx = ops.search(z.detach(), lay, cb, self.search_mode, self.n_cand, self.margin_tau, self.exact_fallback)
        if self.mult == 1:
            # a position carries ONE code, so its normalised vector is a function of the code: normalise the K rows
            # (autograd-tracked torch ops on [K, D]) and run assign / backward on that table
            table = (w / torch.norm(w, p=2, dim=1, keepdim=True)).contiguous()           # quantize.py:56-57
            zq, loss, perp, counts = _AssignFn.apply(z, table, idx, self.beta)
            self.last_counts = counts
            return zq, loss, idx, perp
        # mult > 1: the norm runs over the concatenation of several codes — assign / backward kernels that normalise per
        # position (every sub-row of a position sits in one position tile)
        zq, loss, perp, counts = _NormAssignFn.apply(z, w, idx, self.beta, self.mult)
        self.last_counts = counts
        return zq, loss, idx, perp

    @torch.no_grad()
    def encode_indices(self, z):
        """Indices only (what QVidModel.encode keeps, quantized_video_model.py:798-799): int64 [N]."""
        if not z.is_cuda:
            raise RuntimeError("CUDA only; there is no CPU fallback")
        if z.dtype in (torch.float16, torch.bfloat16):
            z = z.float()            # same upcast as forward()
        elif z.dtype != torch.float32:
            raise TypeError(f"the reference quantizer is FP32 end to end; got {z.dtype}")
        z = z.contiguous()
        if z.numel() == 0:
            return torch.empty(0, dtype=torch.int64, device=z.device)
        lay = ops.layout_of(z.shape, self.e_dim, self.mult)
        return ops.quantize_forward(z, lay, self.embedding.weight, self.beta, self.search_mode, self.n_cand,
                                    self.margin_tau, self.exact_fallback, cb=self._cb_cached(), indices_only=True).idx

    @torch.no_grad()
    def accumulate_from(self, live: "VectorQuantizer", decay: float = 0.999):
        """Polyak step of the reference's `accumulate()` for the quantizer (quantized_video_model.py:951-964,
        `acc(self.net_q_ema, self.net_q, decay)`): self.embedding.weight <- decay*self + (1-decay)*live, written
        through `.data` in place exactly like the reference (no autograd version bump)."""
        ops.polyak(self.embedding.weight.data, live.embedding.weight.data, decay)
        self._cb = None         # the `.data` write is invisible to the version check of a frozen codebook
        return self

    def embed_tokens(self, code: torch.Tensor, tok_emb: torch.Tensor, pos_emb: torch.Tensor):
        """Indices -> prior hand-off (mingpt.py:234-236): tok_emb(code) + pos_emb in one gather pass.
        code [B, T] int64 (what QVidModel.encode returns, quantized_video_model.py:799)."""
        out, err = ops.gather_add(code.contiguous(), tok_emb, pos_emb)
        self._last_gather_err = err
        return out

    def capture(self, z_static: torch.Tensor, decode: bool = False) -> "GraphedQuantizer":
        """CUDA-graph the inference call for inputs of `z_static`'s shape (see GraphedQuantizer)."""
        return GraphedQuantizer(self, z_static, decode=decode)

    @torch.no_grad()
    def fold_pointwise_head(self, head) -> torch.Tensor:
        """Decoder head folded into the codebook (decoder side; inference only).

        The decoder's first block is pointwise in space — `ConvLayer(z_size, block_in, 1)`: a 1x1 EqualConv2d and
        FusedLeakyReLU (skip_autoencoder.py:368,428; gan.py:379-421) — and its input at a position is a codebook
        row, so its output at that position is a function of the CODE alone.  `head` (any module or callable that
        maps `[K, C, 1, 1] -> [K, C_out, 1, 1]`, e.g. the reference's `net_dec.blocks[0]`) is evaluated once on the
        K codebook rows; `embed_code(code, channel_major_hw=(h, w), table=T)` then produces the head's output
        `[G, C_out, h, w]` with the decode gather alone: no per-latent GEMM, no `[G, C, h, w]` intermediate.
        Recompute the table whenever the codebook or the head's weights change."""
        if self.mult != 1:
            raise ValueError("fold_pointwise_head needs mult == 1 (with mult > 1 a position mixes several codes)")
        w = self.embedding.weight
        t = head(w.detach().view(self.n_e, self.e_dim, 1, 1))
        if t.ndim != 4 or t.shape[0] != self.n_e or t.shape[2] != 1 or t.shape[3] != 1:
            raise ValueError(f"head must map [K, C, 1, 1] to [K, C_out, 1, 1]; got {tuple(t.shape)}")
        return t.reshape(self.n_e, -1).to(torch.float32).contiguous()

    def embed_code(self, code, channel_major_hw=None, table=None):
        """E[code] (quantize.py:76-83).  `channel_major_hw=(h, w)` additionally fuses the caller's
        NHWC->NCHW copy (quantized_video_model.py:833): code [G, h, w] -> [G, C, h, w].
        `table` ([K, C_out] from `fold_pointwise_head`) gathers the folded decoder head instead of the codebook."""
        if table is not None:
            return self._embed_table(code, channel_major_hw, table)
        w = self.embedding.weight
        if not code.is_cuda:
            raise RuntimeError("CUDA only; there is no CPU fallback")
        code = code.contiguous()
        if code.dtype != torch.int64:
            code = code.to(torch.int64)
        flag = self._gather_flag(w.device)
        if channel_major_hw is not None:
            h, wd = channel_major_hw
            S = h * wd
            n_pos = code.numel() // self.mult
            lay = Layout(n_pos // S, self.e_dim * self.mult, S, self.mult)
            out, err = ops.gather(code, w, lay, err=flag)
            z = out.view(-1, self.e_dim * self.mult, h, wd)
        else:
            z, err = ops.gather(code, w, err=flag)
            if self.mult > 1:
                s = list(z.shape)
                s[-1] *= self.mult
                s[-2] //= self.mult
                z = z.view(s)
        self._last_gather_err = err   # device flag: nonzero if a code was outside [0, n_e)
        return z

    def _embed_table(self, code, channel_major_hw, table):
        if self.mult != 1:
            raise ValueError("a folded head table needs mult == 1")
        if not code.is_cuda or not table.is_cuda:
            raise RuntimeError("CUDA only; there is no CPU fallback")
        if table.ndim != 2 or table.shape[0] != self.n_e:
            raise ValueError(f"table must be [n_e, C_out]; got {tuple(table.shape)}")
        code = code.contiguous()
        if code.dtype != torch.int64:
            code = code.to(torch.int64)
        flag = self._gather_flag(table.device)
        c_out = table.shape[1]
        if channel_major_hw is not None:
            h, wd = channel_major_hw
            S = h * wd
            lay = Layout(code.numel() // S, c_out, S, 1)
            out, err = ops.gather(code, table, lay, err=flag)
            out = out.view(-1, c_out, h, wd)
        else:
            out, err = ops.gather(code, table, err=flag)
        self._last_gather_err = err
        return out

    def _gather_flag(self, dev) -> torch.Tensor:
        """Persistent device flag of embed_code (one int32, zeroed once): the gather kernels OR a 1 into it when a
        code is out of range, so a call costs no extra fill launch; `check_codes()` reads and clears it."""
        f = getattr(self, "_gather_err_buf", None)
        if f is None or f.device != dev:
            f = torch.zeros(1, dtype=torch.int32, device=dev)
            self._gather_err_buf = f
        return f

    def check_codes(self):
        """Synchronising check of the embed_code calls since the last check (nn.Embedding raises on out-of-range
        codes); clears the flag."""
        err = getattr(self, "_last_gather_err", None)
        if err is not None and int(err.item()) != 0:
            err.zero_()
            raise IndexError("embed_code: index out of range in codebook")


class GraphedQuantizer:
    """One CUDA graph for a whole inference call of the quantizer on fixed-shape inputs.

    The reference issues the quantizer on 1k-5k latents per call (scripts/*/train_frame_autoencoder.sh,
    SURVEY A.5) and once per generated frame in the autoregressive re-encode loop
    (quantized_video_model.py:870-904,939-947): at those sizes the six kernels of a forward take a few
    microseconds each and the cost is launch latency.  `VectorQuantizer.capture(z_static)` records
    forward (+ embed_code) once; `replay()` re-issues all kernels with a single driver call.  The caller
    writes new latents into `z_static` (or passes them to `__call__`, which copies) and reads the static
    outputs `z_q, loss, perplexity, indices[, decoded]`.  Inference only (no autograd through a graph).
    """

    def __init__(self, vq: "VectorQuantizer", z_static: torch.Tensor, decode: bool = False, warmup: int = 2):
        if not z_static.is_cuda:
            raise RuntimeError("CUDA only; there is no CPU fallback")
        self.vq = vq
        self.z = z_static
        self.decode = decode
        lead = z_static.shape[0] if z_static.ndim < 5 else z_static.shape[0] * z_static.shape[1]

        def run():
            z_q, loss, (perp, _, idx) = vq(self.z)
            dec = vq.embed_code(idx.view(lead, -1)) if decode else None
            return z_q, loss, perp, idx, dec

        side = torch.cuda.Stream(z_static.device)
        side.wait_stream(torch.cuda.current_stream(z_static.device))
        with torch.no_grad(), torch.cuda.stream(side):
            for _ in range(warmup):              # module loading, shared-memory opt-ins, allocator warm-up
                run()
        torch.cuda.current_stream(z_static.device).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.z_q, self.loss, self.perplexity, self.indices, self.decoded = run()

    def replay(self):
        self.graph.replay()
        return self.z_q, self.loss, (self.perplexity, None, self.indices)

    def __call__(self, z: Optional[torch.Tensor] = None):
        if z is not None and z.data_ptr() != self.z.data_ptr():
            self.z.copy_(z)
        return self.replay()


class GraphedTrainStep:
    """One CUDA graph for a whole TRAINING step of the quantizer on fixed-shape inputs: forward (search, assign, loss,
    perplexity, code statistics), the statistics exchange, the straight-through backward (dz, and dE for a codebook that
    takes gradients) and — for `EMAVectorQuantizer` — the EMA codebook update.

    Why: at the reference's batch sizes, and on 8 ranks sharing one host, the training step is bound by host issue time
    (0.3-0.5 ms of Python / autograd / launch overhead against 0.45 ms of kernels at the BASELINE c2 shape); the graph
    replays all of it with one driver call.  The caller writes new latents into `z_static` and the upstream gradient of
    z_q into `grad_zq_static`, calls `replay()`, and reads the static outputs.  With several ranks the exchange must be
    the NVLink peer exchange (its step counter lives in device memory; an NCCL collective is not captured here).
    Construct on every rank at the same point: the warm-up steps run the exchange for real."""

    def __init__(self, vq: "VectorQuantizer", z_static: torch.Tensor, grad_zq_static: torch.Tensor,
                 grad_loss: float = 1.0, warmup: int = 3):
        if not z_static.is_cuda:
            raise RuntimeError("CUDA only; there is no CPU fallback")
        if not vq.training:
            raise RuntimeError("GraphedTrainStep captures a training step: call vq.train() first")
        self.vq = vq
        self.z = z_static.detach().requires_grad_(True)
        self.grad_zq = grad_zq_static
        dev = z_static.device
        g_loss = self.grad_loss = torch.full((), float(grad_loss), dtype=torch.float32, device=dev)   # (static: the graph reads it)
        w = vq.embedding.weight

        def step():
            self.z.grad = None
            if w.requires_grad:
                w.grad = None
            z_q, loss, (perp, _, idx) = vq(self.z)
            torch.autograd.backward([z_q, loss], [self.grad_zq, g_loss])
            if hasattr(vq, "sync_codebook"):
                vq.sync_codebook()
            return z_q, loss, perp, idx

        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                step()
        torch.cuda.current_stream(dev).wait_stream(side)
        if getattr(vq, "sync", False) and vq_dist.world_info()[1] > 1 and getattr(vq, "_peer", None) is None:
            raise RuntimeError("GraphedTrainStep with several ranks needs the NVLink peer exchange (exchange='auto' / 'peer' "
                               "on one node); this module exchanges through NCCL")
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.z_q, self.loss, self.perplexity, self.indices = step()
        self.dz = self.z.grad
        self.dE = w.grad if w.requires_grad else None

    def replay(self):
        self.graph.replay()
        return self.z_q, self.loss, self.perplexity, self.indices, self.dz


class EMAVectorQuantizer(VectorQuantizer):
    """EXTENSION (not in the reference): codebook trained by exponential-moving-average statistics
    instead of the reference's Adam-on-codebook-loss (`--q_use_ema` in the reference is a
    Polyak average of all weights, not this).  Parity for this class is pinned only against the
    textbook restatement in oracle/vq_oracle.py::ema_update ("parity unpinned" w.r.t. the reference).

    Per training forward:  per-code counts n_k and sums S_k of the assigned latents (one fused
    scatter-reduce kernel) -> ONE packed all-reduce across ranks (ccvs_b200.dist) -> fused EMA kernel
        N_k <- g N_k + (1-g) n_k ;  m_k <- g m_k + (1-g) S_k ;  E_k <- m_k / smooth(N_k)
    """

    def __init__(self, n_e, e_dim, beta, mult=1, *, decay: float = 0.99, eps: float = 1e-5, sync: bool = True,
                 overlap: bool = True, exchange: str = "auto", **kw):
        super().__init__(n_e, e_dim, beta, mult=mult, normalize=False, **kw)
        if exchange not in ("auto", "peer", "nccl"):
            raise ValueError(f"exchange must be 'auto', 'peer' or 'nccl'; got {exchange!r}")
        self.decay = decay
        self.eps = eps
        self.sync = sync
        self.overlap = overlap            # hide the statistics exchange under the backward pass (see forward)
        # how the ranks exchange the statistics: 'peer' = pushed into every rank's inbox over NVLink and summed by the EMA
        # update kernel itself (ccvs_b200.peer, one node); 'nccl' = one all-reduce of the packed buffer; 'auto' = peer
        # when every rank can map every other rank's memory, else nccl (decided once, collectively, at the first step)
        self.exchange = exchange
        self._peer = None
        self._peer_decided = False
        self.embedding.weight.requires_grad_(False)
        self._inplace_codebook_update = True
        self._last_resid = None
        self._pending = None              # (work handle | None, packed buffer) of a deferred all-reduce + EMA update
        self._deferring = False           # set by forward() when this call defers its codebook update
        self._next_token = 0
        self._outstanding = None          # token of the forward whose backward has not run yet
        self._stale_codebooks = {}        # token -> codebook copy, for a backward that runs AFTER its update was flushed
        self.register_buffer("ema_count", torch.zeros(n_e))
        self.register_buffer("ema_sum", self.embedding.weight.detach().clone())

    def forward(self, z):
        """Training forward = the reference forward + EMA statistics.  The per-code residual sums ride on the assign pass
        (one read of z); with several ranks they are accumulated straight into the packed buffer of ONE all-reduce.

        Overlap (`overlap=True`, autograd on): the all-reduce (several ranks) is issued asynchronously right after the
        forward's kernels; the wait and the in-place EMA update of the codebook are deferred to the end of this
        module's backward (`_QuantizeFn.backward`) — so the collective runs on NCCL's stream under everything between
        this forward and that backward (decoder forward / backward, the dz kernel) instead of sitting serially in front
        of them, and the backward reads the live, not yet updated codebook (no per-step copy).  The update is flushed earlier by
        anything that reads the codebook through this module (next forward, embed_code, sync_codebook, state_dict)."""
        self.sync_codebook()
        if not self.training:
            return super().forward(z)
        # always the packed buffer [resid: K*D | counts: K] fp32: the forward fills both parts itself (assign pass and its
        # folded finalisation), the all-reduce (several ranks) and the EMA update read it as it is
        buf, self._resid_out = vq_dist.ema_stats_buffer(self.n_e, self.e_dim, z.device)
        self._counts_f32_out = buf[self.n_e * self.e_dim:]
        self._want_resid = True
        # deferred mode: the codebook is rewritten only after this module's backward has been enqueued, so the backward
        # can use the live parameter (no per-step copy of the codebook)
        will_defer = self.overlap and torch.is_grad_enabled() and z.requires_grad
        self._inplace_codebook_update = not will_defer
        self._deferring = will_defer
        try:
            out = super().forward(z)
        finally:
            self._want_resid = False
            self._resid_out = None
            self._counts_f32_out = None
        with torch.no_grad():
            self._last_resid = None
            if z.numel() == 0:
                if not (self.sync and vq_dist.world_info()[1] > 1):
                    return out                 # nothing was assigned: the codebook keeps its state
                buf.zero_()                    # an empty shard still joins the collective, with zero statistics
            work = None
            if self.sync and vq_dist.world_info()[1] > 1:
                peer = self._peer_exchange(z.device)
                if peer is not None:
                    peer.publish(buf, overlap=self.overlap)   # posted NVLink writes into every rank's inbox; nobody waits here
                    work = peer
                else:
                    work = vq_dist.start_reduce_ema_stats(buf, None, self.n_e, self.e_dim)
            self._pending = (work, buf)
            if not will_defer:
                self.sync_codebook()
        return out

    def _peer_exchange(self, device):
        """The NVLink peer exchange of this module, set up (collectively) at the first multi-rank training step; None when
        the ranks exchange through NCCL."""
        if not self._peer_decided:
            from .peer import PeerExchange
            self._peer_decided = True
            if self.exchange != "nccl":
                self._peer = PeerExchange.create(self.n_e, self.e_dim, device)
                if self._peer is None and self.exchange == "peer":
                    raise RuntimeError("exchange='peer': the ranks cannot map each other's memory (one node, one process per "
                                       "GPU, peer access and CUDA IPC are required)")
        return self._peer

    def _begin_deferred(self):
        """Token of a forward that keeps the LIVE codebook for its backward (deferred mode), else None."""
        if not self._deferring:
            return None
        self._deferring = False
        self._next_token += 1
        self._outstanding = self._next_token
        return self._next_token

    def _after_backward(self, token=None):
        if token is not None and token == self._outstanding:
            self._outstanding = None
        self.sync_codebook()

    @torch.no_grad()
    def sync_codebook(self):
        """Complete a deferred statistics exchange: the current stream waits for the all-reduce, then the EMA update
        rewrites the codebook in place.  No-op when nothing is pending.  If the backward of the forward that produced
        the statistics has not run yet (gradient accumulation, a second forward first, ...), the codebook it used is
        copied aside for it before the rewrite."""
        if self._pending is None:
            return
        work, buf = self._pending
        self._pending = None
        if self._outstanding is not None:
            self._stale_codebooks[self._outstanding] = self.embedding.weight.detach().clone()
            self._outstanding = None
        if work is not None and work is self._peer:
            # collective + update in one pass: wait for the ranks' pushes, sum the inbox slots in rank order, EMA update
            work.ema_update(self.embedding.weight, self.ema_count, self.ema_sum, self.decay, self.eps)
            return
        if work is not None:
            work.wait()            # the current STREAM waits for the collective (no host block with NCCL)
        ops.ema_update_packed(self.embedding.weight, self.ema_count, self.ema_sum, buf, self.decay, self.eps)

    def embed_code(self, code, channel_major_hw=None, table=None):
        self.sync_codebook()
        return super().embed_code(code, channel_major_hw=channel_major_hw, table=table)

    def state_dict(self, *args, **kw):
        self.sync_codebook()
        return super().state_dict(*args, **kw)
