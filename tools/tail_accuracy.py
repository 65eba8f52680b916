"""Accuracy of the encoder-tail kernel against float64: max and RMS error relative to the accumulated magnitude."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ccvs_b200 import EncoderTail
dev = torch.device("cuda", 0)
torch.backends.cudnn.allow_tf32 = False
for (G, ci, co, h, w) in ((64, 512, 256, 8, 8), (16, 256, 512, 16, 16), (32, 64, 64, 8, 8), (8, 1024, 256, 8, 8)):
    gen = torch.Generator().manual_seed(ci + co)
    x = torch.randn(G, ci, h, w, generator=gen) * 2
    m = EncoderTail(ci, co)
    with torch.no_grad():
        m.bias.copy_(torch.randn(co, generator=gen))
    ws = (m.weight * m.scale).double()[:, :, 0, 0]
    pre = torch.einsum("oc,gchw->gohw", ws, x.double()) + m.bias.double().view(1, -1, 1, 1)
    mag = torch.einsum("oc,gchw->gohw", ws.abs(), x.double().abs()) + m.bias.double().abs().view(1, -1, 1, 1)
    ref = torch.where(pre > 0, pre, pre * 0.1)
    with torch.no_grad():
        cpu32 = torch.nn.functional.leaky_relu(torch.nn.functional.conv2d(x, m.weight * m.scale, m.bias), 0.1)
        md = m.to(dev)
        ours = md(x.to(dev)).cpu()
        lib = torch.nn.functional.leaky_relu(torch.nn.functional.conv2d(x.to(dev), md.weight * md.scale, md.bias), 0.1).cpu()
    for name, t in (("ours (BF16x6)", ours), ("torch CPU FP32", cpu32), ("cuDNN FP32 (TF32 off)", lib)):
        e = (t.double() - ref).abs() / mag
        print(f"C_in {ci:4d} C_out {co:3d}: {name:22s} max err / magnitude {float(e.max()):.2e}   rms {float((e**2).mean().sqrt()):.2e}", flush=True)
