"""Same-box A/B of screen plans through the whole forward: distributions T and I at a bench workload.
usage: CCVSQ_SCREEN_PLAN=bn,nacc,abuf python tools/ab_plan.py [workload]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ccvs_b200 import VectorQuantizer

wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
dev = torch.device("cuda", 0)
(clips, frames), D, h, w_, K, desc = bench.WORKLOADS[wl]
out = []
for dist, mk in (("T", lambda: bench.make_inputs(wl, dev, 1234)), ("I", lambda: bench.make_inputs_I(wl, dev, 1234))):
    z, cb, n = mk()
    vq = VectorQuantizer(K, D, 0.25).to(dev).eval()
    with torch.no_grad():
        vq.embedding.weight.copy_(cb)
        for _ in range(5):
            vq(z)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ts = []
        for _ in range(20):
            e0.record(); vq(z); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort()
    out.append(f"{dist}: forward median {ts[len(ts)//2]*1e3:.1f} us")
print(f"plan {os.environ.get('CCVSQ_SCREEN_PLAN', 'auto'):10s} {wl}  " + "   ".join(out))
