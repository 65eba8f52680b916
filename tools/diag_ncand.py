"""Flag breakdown of the screen queue and forward time by candidate slots (n_cand) on the stress distribution I."""
import os, sys
import torch
sys.path.insert(0, "/root/repo")
import bench
from ccvs_b200 import VectorQuantizer, ops
dev = torch.device("cuda", 0)
wl = "c2"
(clips, frames), D, h, w_, K, desc = bench.WORKLOADS[wl]
z, cb, n = bench.make_inputs_I(wl, dev, 1234)
lay = ops.layout_of(z.shape, D, 1)
pcb = ops.prepare_codebook(cb.contiguous())
for nc in (4, 6, 8):
    idx, q = ops.screen(z, lay, pcb, nc, 1.0)
    torch.cuda.synchronize()
    cnt = int(q.count)
    fl = q.flags[:cnt]
    print(f"n_cand {nc}: queued {cnt}  flag&1 (more codes than slots) {int((fl & 1).ne(0).sum())}  flag&2 (dropped chunk) {int((fl & 2).ne(0).sum())}  any {int(fl.ne(0).sum())}")
    vq = VectorQuantizer(K, D, 0.25, n_cand=nc).to(dev).eval()
    with torch.no_grad():
        vq.embedding.weight.copy_(cb)
        for _ in range(3): vq(z)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10): vq(z)
        b.record(); torch.cuda.synchronize()
    print(f"   forward {a.elapsed_time(b) / 10 * 1e3:.1f} us")
