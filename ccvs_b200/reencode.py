"""The per-frame re-encode step of CCVS's autoregressive sampler as ONE CUDA graph.

Reference loop (models/skip_vid_generator/models/quantized_video_model.py:939-964 `vid_step_decode`, driven by
helpers/generator.py:142-159 once per generated frame):

    z    = net_q.embed_code(code.view(-1, h, w))                    # :947   codes -> embeddings      (this package)
    z    = z.view(...).transpose(-2, -1).transpose(-3, -2).contiguous()   # :948   NHWC -> NCHW           (fused into the gather)
    fake = net_g(z, inters, ...)                                    # :955   decoder                  (caller's module)
    new  = self.encode(fake, ...)                                   # :956   encoder trunk            (caller's module)
                                                                    #        encoder tail             (EncoderTail)
                                                                    #        quantizer -> new codes   (VectorQuantizer)

Every call works on B x 64 latents (8x8 latent frames): a dozen kernels of a few microseconds each, i.e. pure launch
latency.  `ReencodeStep` records the whole chain once — decode gather written channel-major, the caller's decoder +
encoder trunk (any CUDA-graph-capturable callable), the tcgen05 encoder tail and the indices-only quantizer forward on a
FROZEN codebook (side data and tensor maps are built once, not per frame) — and replays it with a single driver call
per generated frame.  The new codes land in a static buffer that can be fed straight back in.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch


class ReencodeStep:
    """codes [B, h*w] -> new codes [B, h*w] through decode -> trunk -> encoder tail -> quantizer, CUDA-graphed.

    vq      ccvs_b200.VectorQuantizer (eval; its codebook must not change while the step object lives)
    tail    ccvs_b200.EncoderTail producing the latents from the trunk's features
    trunk   callable: z [B, C, h, w] (decoded latents, NCHW) -> features [B, C_feat, h, w]; stands for the reference's
            decoder + encoder trunk (out of this package's scope: any capturable torch callable)
    code    int64 [B, h*w] example input (its storage becomes the static input buffer)
    """

    def __init__(self, vq, tail, trunk: Callable[[torch.Tensor], torch.Tensor], code: torch.Tensor, hw, warmup: int = 2):
        if not code.is_cuda:
            raise RuntimeError("CUDA only; there is no CPU fallback")
        self.vq, self.tail, self.trunk = vq, tail, trunk
        self.h, self.w = hw
        self.code = code.contiguous().clone()
        self.B = code.shape[0]
        vq.freeze_codebook()
        tail._terms_cached()

        def run():
            z = vq.embed_code(self.code.view(self.B, self.h, self.w), channel_major_hw=(self.h, self.w))    # [B, C, h, w]
            feats = trunk(z)
            lat = tail(feats)
            return vq.encode_indices(lat).view(self.B, self.h * self.w), z

        dev = code.device
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.no_grad(), torch.cuda.stream(side):
            for _ in range(warmup):
                run()
        torch.cuda.current_stream(dev).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.new_code, self.decoded = run()
        self._eager = run

    def eager(self, code: Optional[torch.Tensor] = None):
        """The same chain without the graph (checks, timing)."""
        if code is not None:
            self.code.copy_(code)
        with torch.no_grad():
            return self._eager()[0]

    def step(self, code: Optional[torch.Tensor] = None) -> torch.Tensor:
        """One generated frame: replay the graph on `code` (default: the static input buffer as it is)."""
        if code is not None and code.data_ptr() != self.code.data_ptr():
            self.code.copy_(code)
        self.graph.replay()
        return self.new_code

    def feed_back(self):
        """new codes -> input buffer (the next frame's call re-encodes what this frame produced)."""
        self.code.copy_(self.new_code)
        return self
