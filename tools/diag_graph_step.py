"""Eager vs GraphedTrainStep, per-output max abs difference over a few steps (diagnosis helper)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ccvs_b200.quantize import EMAVectorQuantizer, GraphedTrainStep, VectorQuantizer
dev = torch.device("cuda", 0)
K, D, shape = 512, 64, (4, 64, 16, 16)
for ema in (True, False):
    g = torch.Generator().manual_seed(11)
    cb = torch.randn(K, D, generator=g)
    zs = [(cb[torch.randint(0, K, (4 * 256,), generator=g)] + 0.3 * torch.randn(4 * 256, D, generator=g))
          .view(4, 16, 16, D).permute(0, 3, 1, 2).contiguous().to(dev) for _ in range(3)]
    gz = [torch.randn(shape, generator=g).to(dev) for _ in range(3)]
    def make():
        vq = (EMAVectorQuantizer(K, D, 0.25, decay=0.9, deterministic=True) if ema else VectorQuantizer(K, D, 0.25, deterministic=True))
        return vq.to(dev).train()
    def reset(vq):
        with torch.no_grad():
            if ema:
                vq.sync_codebook(); vq.ema_sum.copy_(cb.to(dev)); vq.ema_count.fill_(1.0)
            vq.embedding.weight.copy_(cb.to(dev))
    eager, outs_e = make(), []
    reset(eager)
    for z, gq in zip(zs, gz):
        zt = z.clone().requires_grad_(True)
        eager.embedding.weight.grad = None
        z_q, loss, (perp, _, idx) = eager(zt)
        torch.autograd.backward([z_q, loss], [gq, torch.ones_like(loss)])
        if ema: eager.sync_codebook()
        outs_e.append([t.detach().clone() for t in (z_q, loss, perp, idx.view(-1), zt.grad)] +
                      [eager.embedding.weight.detach().clone() if ema else eager.embedding.weight.grad.clone()])
    vq = make(); reset(vq)
    step = GraphedTrainStep(vq, zs[0].clone(), gz[0].clone())
    reset(vq)
    for i, (z, gq) in enumerate(zip(zs, gz)):
        step.z.data.copy_(z); step.grad_zq.copy_(gq)
        z_q, loss, perp, idx, dz = step.replay()
        torch.cuda.synchronize()
        last = vq.embedding.weight.detach() if ema else step.dE
        for name, a, b in zip(("z_q", "loss", "perp", "idx", "dz", "cb/dE"), outs_e[i], (z_q, loss, perp, idx.view(-1), dz, last)):
            d = (a.double() - b.double()).abs().max().item()
            print(f"ema={ema} step {i} {name:6s} max|diff| {d:.3e}  equal {torch.equal(a, b)}")
        # dz against the closed form with the eager step's codebook BEFORE its update
        print("   dz - g:", (dz - gq).abs().max().item(), " eager dz - g:", (outs_e[i][4] - gq).abs().max().item(),
              " dz - g_prev:", (dz - gz[i - 1]).abs().max().item() if i else None)
