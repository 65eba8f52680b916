"""Diagnostic: cProfile of the bench step (module forward + embed_code) to find host-side time."""
import sys, time, os, cProfile, pstats
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ccvs_b200 import ops, VectorQuantizer

wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
dev = torch.device("cuda", 0)
(clips, frames), D, h, w_, K, desc = bench.WORKLOADS[wl]
z, cb, n = bench.make_inputs(wl, dev, 1234)
vq = VectorQuantizer(K, D, 0.25).to(dev).eval()
with torch.no_grad():
    vq.embedding.weight.copy_(cb)


def step():
    with torch.no_grad():
        z_q, loss, (perp, _, idx) = vq(z)
        dec = vq.embed_code(idx.view(clips * frames, h, w_))
    return idx, loss, perp, dec


for _ in range(5):
    out = step()
torch.cuda.synchronize()
for timing in (False, {"ccvsq_screen"}):
    ops.PROFILER.reset(timing=timing)
    t0 = time.perf_counter()
    for _ in range(20):
        out = step()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"timing={timing}: host issue {(t1 - t0) / 20 * 1e3:.3f} ms/step, with sync {(t2 - t0) / 20 * 1e3:.3f} ms/step", flush=True)
ops.PROFILER.reset(timing=False)
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    out = step()
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
