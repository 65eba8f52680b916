"""Where the time of the stress distribution (I) goes: screen / rescore / exact fallback, event-timed one by one, with
the queue and fallback counts.  usage: python tools/diag_stress.py [workload]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ccvs_b200 import ops

wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
dev = torch.device("cuda", 0)
(clips, frames), D, h, w_, K, desc = bench.WORKLOADS[wl]
for dist in ("T", "I"):
    z, cb, n = (bench.make_inputs if dist == "T" else bench.make_inputs_I)(wl, dev, 1234)
    lay = ops.layout_of(z.shape, D, 1)
    pcb = ops.prepare_codebook(cb.contiguous())
    ops.PROFILER.reset(timing=True)
    for it in range(4):
        if it == 1:
            torch.cuda.synchronize(); ops.PROFILER.reset(timing=True)
        idx, q = ops.screen(z, lay, pcb, 4, 1.0)
        idx = ops.rescore(z, lay, pcb, idx, q, True)
    torch.cuda.synchronize()
    s = ops.PROFILER.summary()
    ops.PROFILER.reset(timing=False)
    print(f"{wl} dist {dist}: queued {int(q.count)} of {n}  " + "  ".join(f"{k[6:]} {t / c * 1e3:.1f} us" for k, (c, t) in s.items()))
    # the rest of the step on the same indices: assign (z_q, loss, counts) and the decode gather
    ops.PROFILER.reset(timing=True)
    for it in range(4):
        if it == 1:
            torch.cuda.synchronize(); ops.PROFILER.reset(timing=True)
        ops.assign(z, lay, cb, idx)
        ops.gather(idx, cb)
    torch.cuda.synchronize()
    s = ops.PROFILER.summary()
    ops.PROFILER.reset(timing=False)
    used = int(torch.bincount(idx, minlength=K).ne(0).sum())
    print(f"   assign / gather on these indices: " + "  ".join(f"{k[6:]} {t / c * 1e3:.1f} us" for k, (c, t) in s.items()) + f"   codes in use {used} of {K}, max share {float(torch.bincount(idx, minlength=K).max()) / n:.4f}")
