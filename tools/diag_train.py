"""Where a training step of the EMA quantizer spends its time: device time vs host issue time, by variant."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ccvs_b200.quantize import EMAVectorQuantizer
dev = torch.device("cuda", 0)
z, cb, n = bench.make_inputs("c2", dev, 1234)
K, D = cb.shape
g_out = torch.randn_like(z)
zt = z.detach().clone().requires_grad_(True)
for name, kw in (("deferred (overlap=True)", dict(overlap=True)), ("immediate (overlap=False)", dict(overlap=False))):
    vq = EMAVectorQuantizer(K, D, 0.25, decay=0.99, sync=False, **kw).to(dev).train()
    with torch.no_grad():
        vq.embedding.weight.copy_(cb); vq.ema_sum.copy_(cb); vq.ema_count.fill_(1.0)
    def step():
        zt.grad = None
        z_q, loss, _ = vq(zt)
        torch.autograd.backward([z_q, loss], [g_out, torch.ones_like(loss)])
    for _ in range(5): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(50): step()
    e1.record(); host = (time.perf_counter() - t0) / 50 * 1e3
    torch.cuda.synchronize()
    print(f"{name:28s} device {e0.elapsed_time(e1)/50:.4f} ms/step   host issue {host:.4f} ms/step", flush=True)
    # forward only / backward only host cost
    t0 = time.perf_counter()
    for _ in range(50):
        z_q, loss, _ = vq(zt)
    torch.cuda.synchronize(); print(f"   forward only: {(time.perf_counter()-t0)/50*1e3:.4f} ms (wall, synced at end)")
