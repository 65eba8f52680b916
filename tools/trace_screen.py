"""GPU diagnostic: event timeline of CTA 0 of the screening kernel on a bench workload."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from ccvs_b200 import ops

wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
z, cb, n = bench.make_inputs(wl, torch.device("cuda:0"), 1234)
(clips, frames), D, h, w, K, _ = bench.WORKLOADS[wl]
lay = ops.layout_of(z.shape, D, 1)
pcb = ops.prepare_codebook(cb)
zb, margin = ops.pack_latents(z, lay, pcb, 1.0)
for _ in range(2):
    tr = ops.screen_trace(zb, margin, pcb, lay.rows, 4)
torch.cuda.synchronize()
tr = tr.cpu()
names = {0: "producer", 1: "mma", 2: "epi0", 3: "epi1"}
ev = []
for role in range(4):
    for x in tr[role].tolist():
        if x == 0:
            break
        ev.append((x >> 8, role, x & 255))
ev.sort()
t0 = ev[0][0]
print("events", len(ev))
# print the timeline of the 3rd..5th row tiles (steady state): use mma 'tile issued' events to delimit
lim = int(sys.argv[2]) if len(sys.argv) > 2 else 260
for t, role, code in ev[:lim]:
    print(f"{t - t0:9d}  {names[role]:9s} {code}")
# per-role statistics
import collections
for role in range(4):
    seq = [(t, c) for t, r, c in ev if r == role]
    d = collections.defaultdict(list)
    for (ta, ca), (tb, cb_) in zip(seq, seq[1:]):
        d[(ca, cb_)].append(tb - ta)
    print(names[role], {k: (len(v), sum(v) // len(v)) for k, v in sorted(d.items())})
