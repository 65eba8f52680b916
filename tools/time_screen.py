"""Device time of the screen kernel alone (CUDA events, 30 launches after warm-up).  usage: time_screen.py [workload]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ccvs_b200 import ops

wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
tau = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
dev = torch.device("cuda", 0)
(clips, frames), D, h, w_, K, desc = bench.WORKLOADS[wl]
z, cb, n = bench.make_inputs(wl, dev, 1234)
lay = ops.layout_of(z.shape, D, 1)
pcb = ops.prepare_codebook(cb.contiguous())
for _ in range(5):
    ops.screen(z, lay, pcb, 4, tau)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ts = []
for _ in range(30):
    idx = torch.empty(n, dtype=torch.int64, device=dev)
    q = ops._new_queue(n, 4, dev)
    e0.record()
    ops._call("ccvsq_screen", ops._ptr(z), lay, ops._ptr(pcb.e_bf16), ops._ptr(pcb.e_max), pcb.K, tau, 4, ops._ptr(idx),
              ops._ptr(q.count), ops._ptr(q.rows), ops._ptr(q.cand), ops._ptr(q.flags), ops._stream(dev))
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ts.sort()
flops = 2.0 * n * K * D
print(f"{os.environ.get('CCVSQ_LIB', 'default'):60s} {wl} tau {tau} plan {os.environ.get('CCVSQ_SCREEN_PLAN', 'auto')}: screen median {ts[len(ts)//2]*1e3:.1f} us  best {ts[0]*1e3:.1f} us  "
      f"({flops / (ts[len(ts)//2] * 1e-3) / 1e12:.0f} TFLOP/s)")
