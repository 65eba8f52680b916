"""Composite forward on the stress distribution: whole call vs pieces, with and without PDL (CCVSQ_NO_PDL=1)."""
import os, sys
import torch
sys.path.insert(0, "/root/repo")
import bench
from ccvs_b200 import ops
dev = torch.device("cuda", 0)
wl = "c2"
(clips, frames), D, h, w_, K, desc = bench.WORKLOADS[wl]
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
for dist in (os.environ.get("DIST", "T,I").split(",")):
    z, cb, n = (bench.make_inputs if dist == "T" else bench.make_inputs_I)(wl, dev, 1234)
    lay = ops.layout_of(z.shape, D, 1)
    pcb = ops.prepare_codebook(cb.contiguous())
    full = t(lambda: ops.quantize_forward(z, lay, cb, 0.25))
    full_cb = t(lambda: ops.quantize_forward(z, lay, cb, 0.25, cb=pcb))
    idx_only = t(lambda: ops.quantize_forward(z, lay, cb, 0.25, cb=pcb, indices_only=True))
    def pieces():
        idx, q = ops.screen(z, lay, pcb, 4, 1.0)
        ops.rescore(z, lay, pcb, idx, q, True)
    pc = t(pieces)
    print(f"PDL {'off' if os.environ.get('CCVSQ_NO_PDL') else 'on '} dist {dist}: composite {full:.1f} us | with cached codebook {full_cb:.1f} | indices only {idx_only:.1f} | screen+rescore+fallback as separate calls {pc:.1f}")
