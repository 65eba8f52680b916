"""Encoder tail of the CCVS frame autoencoder on B200 tensor cores.

The encoder's last block is  `ConvLayer(block_out, z_size, 1)`  — a 1x1 `EqualConv2d` (weight * 1/sqrt(C_in), bias) followed
by `nn.LeakyReLU(0.1)` — and, if `normalize_out`, an L2 normalisation over the channel dim:
    /root/reference/models/skip_vid_generator/models/skip_autoencoder.py:331,346-349  (EqualConv2d :40-58, ConvLayer :66-101)
Its output IS the latent tensor z the quantizer consumes.  `EncoderTail` keeps the reference block's parameters
(`weight [C_out, C_in, 1, 1]` ~ N(0,1), `bias [C_out]` zeros: a reference checkpoint's `blocks.<n>.0.weight / .bias` load
directly) and runs the forward as one tcgen05 GEMM kernel with FP32-level accuracy (three-term BF16 split of both
operands, six products per K step: `ccvsq_encoder_tail`), bias + LeakyReLU in the epilogue, z written once in the NCHW
layout the quantizer reads in place.  CUDA only; inference path (`QVidModel.encode` runs under `torch.no_grad()`,
quantized_video_model.py:773-799) — under autograd the backward falls to two plain library GEMMs.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import ops


class _TailFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, module):
        terms = module._terms_cached()
        z = ops.encoder_tail(x, terms, bias, module.negative_slope, normalize=False)
        ctx.save_for_backward(x, weight, z)
        ctx.scale, ctx.slope = module.scale, module.negative_slope
        return z

    @staticmethod
    def backward(ctx, gz):
        # training is not this kernel's job (the reference trains the whole encoder with autograd): plain library GEMMs
        x, weight, z = ctx.saved_tensors
        g = torch.where(z > 0, gz, gz * ctx.slope)                                  # LeakyReLU'
        w = weight.reshape(weight.shape[0], -1) * ctx.scale
        gx = torch.einsum("oc,go...->gc...", w, g) if ctx.needs_input_grad[0] else None
        gw = (torch.einsum("go...,gc...->oc", g, x) * ctx.scale).reshape(weight.shape) if ctx.needs_input_grad[1] else None
        gb = g.transpose(0, 1).reshape(g.shape[1], -1).sum(1) if ctx.needs_input_grad[2] else None
        return gx, gw, gb, None


class EncoderTail(nn.Module):
    """Drop-in for the encoder's last `ConvLayer(in_channel, out_channel, 1)` (+ optional output normalisation)."""

    def __init__(self, in_channel: int, out_channel: int, normalize_out: bool = False, negative_slope: float = 0.1):
        super().__init__()
        self.weight = nn.Parameter(torch.randn(out_channel, in_channel, 1, 1))     # EqualConv2d, skip_autoencoder.py:43
        self.bias = nn.Parameter(torch.zeros(out_channel))                          # :50
        self.scale = 1 / math.sqrt(in_channel)                                      # :44 (kernel_size = 1)
        self.negative_slope = negative_slope                                        # :99
        self.normalize_out = normalize_out                                          # :348
        self._terms = None
        if not ops.encoder_tail_supported(in_channel, out_channel):
            raise ValueError(f"EncoderTail needs C_in % 64 == 0 and C_out % 16 == 0 (<= 256 or a multiple of 256); got "
                             f"{in_channel} -> {out_channel}")

    def _terms_cached(self) -> torch.Tensor:
        w = self.weight
        t = self._terms
        if t is None or t[1] != w._version or t[2] != w.data_ptr():
            self._terms = t = (ops.encoder_tail_prepare(w, self.scale), w._version, w.data_ptr())
        return t[0]

    def refresh(self):
        """Drop the cached BF16 weight terms (after writing `weight.data` behind autograd's back)."""
        self._terms = None
        return self

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not x.is_cuda:
            raise RuntimeError("ccvs_b200.EncoderTail runs on CUDA (sm_100a) only; there is no CPU fallback")
        lead = x.shape[:-3]
        x4 = x.reshape((-1,) + tuple(x.shape[-3:])).contiguous().float()            # flatten_vid: [B, T, C, h, w] -> [B*T, C, h, w]
        if torch.is_grad_enabled() and (x4.requires_grad or self.weight.requires_grad):
            z = _TailFn.apply(x4, self.weight, self.bias, self)
            if self.normalize_out:
                z = z / torch.norm(z, p=2, dim=1, keepdim=True)
        else:
            z = ops.encoder_tail(x4, self._terms_cached(), self.bias, self.negative_slope, normalize=self.normalize_out)
        return z.reshape(tuple(lead) + tuple(z.shape[1:]))
