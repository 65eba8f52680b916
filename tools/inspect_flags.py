"""GPU diagnostic: candidate statistics of the screen on the bench workloads (flags, #candidates)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from ccvs_b200 import ops

for wl in sys.argv[1:] or ["c2", "c3"]:
    z, cb, n = bench.make_inputs(wl, torch.device("cuda:0"), 1234)
    (clips, frames), D, h, w, K, _ = bench.WORKLOADS[wl]
    lay = ops.layout_of(z.shape, D, 1)
    pcb = ops.prepare_codebook(cb)
    zb, margin = ops.pack_latents(z, lay, pcb, 1.0)
    sr = ops.screen(zb, margin, pcb, lay.rows, 4)
    torch.cuda.synchronize()
    ci, sc, flag = sr.merged()
    nc = (ci >= 0).sum(1)
    print(wl, "rows", n, "flags", int(flag.sum()), "ncand hist", torch.bincount(nc, minlength=5).tolist(),
          "margin mean", float(margin[:n].mean()))
    for r in flag.nonzero().view(-1)[:5].tolist():
        print("  row", r, sr.cand_idx[r].tolist(), sr.cand_score[r].tolist(), float(margin[r]))
