"""Host-fed quantization: overlap the PCIe copy of the latents with their quantization.

Rows (latents) are independent given the codebook, so a batch that arrives in HOST memory
is cut into frame chunks: chunk c+1 crosses PCIe (pinned memory, copy stream) while chunk c runs through
`VectorQuantizer.forward` (+ optionally `embed_code`) on the compute stream, and the indices of chunk
c-1 travel back on a third stream (PCIe is full duplex).  The device-side work per batch is 5-10 % of
the copy time at the BASELINE shapes, so the end-to-end rate is the PCIe rate instead of copy + compute.

Results are those of one whole-batch call: indices are concatenated; the loss is (1+beta)*sum of squared
errors / M with the per-chunk sums combined on the device (equal-sized chunks: the mean of the chunk
losses); the perplexity comes from the summed per-code counts (quantize.py:67-68) through the same
`ccvsq_finalize` kernel.
"""
from __future__ import annotations

from typing import Tuple

import torch

from . import ops


class HostQuantizePipeline:
    """Reusable staging (device buffer, streams, events) for batches of one shape.

    vq        ccvs_b200.VectorQuantizer on a CUDA device (eval / no_grad use)
    z_shape   shape of the host batch [G, (T,) C, h, w]; it is chunked along dim 0
    n_chunks  number of frame chunks (dim 0 must be divisible)
    decode    also run embed_code on every chunk (the decoded latents stay on the device)
    """

    def __init__(self, vq, z_shape, n_chunks: int = 8, decode: bool = True):
        self.vq = vq
        self.dev = vq.embedding.weight.device
        self.shape = tuple(int(s) for s in z_shape)
        if self.shape[0] % n_chunks:
            raise ValueError(f"leading dim {self.shape[0]} not divisible into {n_chunks} chunks")
        self.n_chunks = n_chunks
        self.decode = decode
        self.z_stage = torch.empty(self.shape, dtype=torch.float32, device=self.dev)
        self.rows_per_chunk = self.z_stage[: self.shape[0] // n_chunks].numel() // (vq.e_dim)
        n_rows = self.rows_per_chunk * n_chunks
        self.idx_dev = torch.empty(n_rows, dtype=torch.int64, device=self.dev)
        self.idx_host = torch.empty(n_rows, dtype=torch.int64, pin_memory=True)
        self.scalars_host = torch.empty(2, dtype=torch.float32, pin_memory=True)
        self.h2d = torch.cuda.Stream(self.dev)
        self.d2h = torch.cuda.Stream(self.dev)
        self.ev_in = [torch.cuda.Event() for _ in range(n_chunks)]
        self.ev_done = [torch.cuda.Event() for _ in range(n_chunks)]
        self._primed = False
        self.decoded = [None] * n_chunks

    @torch.no_grad()
    def run(self, z_host: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """z_host: pinned FP32 tensor of `z_shape`.  Enqueues everything and returns (idx_host int64 [N],
        scalars_host [loss, perplexity]) — valid after `torch.cuda.current_stream().synchronize()` and
        `self.d2h.synchronize()` (or any device-wide synchronisation)."""
        if tuple(z_host.shape) != self.shape or not z_host.is_pinned():
            raise ValueError("z_host must be a pinned tensor of the pipeline's shape")
        vq, n = self.vq, self.n_chunks
        cur = torch.cuda.current_stream(self.dev)
        g = self.shape[0] // n
        losses = []
        counts = None
        for c in range(n):
            zc = self.z_stage[c * g:(c + 1) * g]
            with torch.cuda.stream(self.h2d):
                if self._primed:
                    self.h2d.wait_event(self.ev_done[c])        # the previous batch has consumed this slot
                zc.copy_(z_host[c * g:(c + 1) * g], non_blocking=True)
                self.ev_in[c].record(self.h2d)
            cur.wait_event(self.ev_in[c])
            _, loss, (_, _, idx) = vq(zc)
            self.idx_dev[c * self.rows_per_chunk:(c + 1) * self.rows_per_chunk].copy_(idx.view(-1))
            if self.decode:
                self.decoded[c] = vq.embed_code(idx.view(g, -1))
            losses.append(loss)
            counts = vq.last_counts.clone() if counts is None else counts.add_(vq.last_counts)
            self.ev_done[c].record(cur)
            with torch.cuda.stream(self.d2h):
                self.d2h.wait_event(self.ev_done[c])
                sl = slice(c * self.rows_per_chunk, (c + 1) * self.rows_per_chunk)
                self.idx_host[sl].copy_(self.idx_dev[sl], non_blocking=True)
        self._primed = True
        loss = torch.stack(losses).mean()
        _, _, perp = ops.finalize(vq.n_e, vq.e_dim, float(self.z_stage.numel()), float(self.idx_dev.numel()), vq.beta,
                                  counts=counts, want_perplexity=True)
        self.scalars_host.copy_(torch.stack([loss, perp]), non_blocking=True)
        return self.idx_host, self.scalars_host

    def synchronize(self):
        torch.cuda.current_stream(self.dev).synchronize()
        self.d2h.synchronize()
