"""GPU diagnostic: find flagged rows at c3, dump their score rows and emulate the candidate-list
algorithm on the host to see why the overflow flag was raised."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from ccvs_b200 import ops

wl = "c3"
z, cb, n = bench.make_inputs(wl, torch.device("cuda:0"), 1234)
(clips, frames), D, h, w, K, _ = bench.WORKLOADS[wl]
lay = ops.layout_of(z.shape, D, 1)
pcb = ops.prepare_codebook(cb)
zb, margin = ops.pack_latents(z, lay, pcb, 1.0)
sr = ops.screen(zb, margin, pcb, lay.rows, 4)
torch.cuda.synchronize()
rows = sr.flags.nonzero().view(-1).tolist()
print("flagged rows", rows)
LCAP = 12
for r in rows[:3]:
    t0 = r // 128 * 128
    sr2, scores = ops.screen_dump(zb[t0:t0 + 128].contiguous(), margin[t0:t0 + 128].contiguous(), pcb, 128, 4)
    torch.cuda.synchronize()
    print(" re-run tile: flag", int(sr2.flags[r - t0]), sr2.cand_idx[r - t0].tolist())
    s = scores[r - t0].cpu()
    mg = float(margin[r])
    for g in (0, 1):
        runmax = -3.4e38; lst = []; ovf = 0; ncomp = 0
        for j in range(g, K // 256, 2):
            for c in range(8):
                v = s[j * 256 + c * 32: j * 256 + c * 32 + 32]
                m = float(v.max())
                if m >= runmax - mg:
                    runmax = max(runmax, m); thr = runmax - mg
                    for i in range(8):
                        if float(v[4 * i:4 * i + 4].max()) >= thr:
                            if len(lst) > LCAP - 4:
                                lst = [e for e in lst if e[0] >= thr]; ncomp += 1
                                if len(lst) > LCAP - 4:
                                    print("   OVERFLOW at tile", j, "chunk", c, "grp", i, "thr", thr, "runmax", runmax, "list", lst)
                                    ovf = 1; lst = lst[:LCAP - 4]
                            for u in range(4):
                                if float(v[4 * i + u]) >= thr:
                                    lst.append((float(v[4 * i + u]), j * 256 + c * 32 + 4 * i + u))
        print("  group", g, "runmax", runmax, "n", len(lst), "ovf", ovf, "compactions", ncomp, "margin", mg)
