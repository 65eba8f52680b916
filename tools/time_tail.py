"""Device time of the encoder-tail kernel at the BASELINE shapes, next to the FP32 library convolution (cuDNN / cuBLAS,
TF32 disabled) of the same block.  usage: python tools/time_tail.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ccvs_b200 import EncoderTail
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda", 0)
for name, (G, ci, co, h, w) in {"c2 (BAIR-256: 1024 frames, 512 -> 256, 16x16)": (1024, 512, 256, 16, 16),
                                "c3d512 shard (1024 frames, 512 -> 512, 8x8)": (1024, 512, 512, 8, 8),
                                "c1 (16 frames, 512 -> 256, 16x16)": (16, 512, 256, 16, 16)}.items():
    x = torch.randn(G, ci, h, w, device=dev)
    m = EncoderTail(ci, co).to(dev)
    def ours():
        with torch.no_grad():
            return m(x)
    def lib():
        with torch.no_grad():
            return torch.nn.functional.leaky_relu(torch.nn.functional.conv2d(x, m.weight * m.scale, m.bias), 0.1)
    def t(fn):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10): fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / 10
    to, tl = t(ours), t(lib)
    flops = 2.0 * G * h * w * ci * co
    print(f"{name}: ours {to*1e3:8.1f} us ({flops/to/1e9:7.1f} TFLOP/s of FP32-accurate work, {6*flops/to/1e9:7.1f} BF16 tensor TFLOP/s)"
          f" | FP32 library conv + lrelu {tl*1e3:8.1f} us | max |diff| {float((ours()-lib()).abs().max()):.2e}", flush=True)
