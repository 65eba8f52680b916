"""Timeline of the screen kernel's pipeline hand-offs (ccvsq_screen_trace) at a bench workload."""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ccvs_b200 import ops

wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
dev = torch.device("cuda", 0)
(clips, frames), D, h, w_, K, desc = bench.WORKLOADS[wl]
z, cb, n = bench.make_inputs(wl, dev, 1234)
lay = ops.layout_of(z.shape, D, 1)
pcb = ops.prepare_codebook(cb.contiguous())
tau = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0    # e.g. -1e28: the candidate slow path is never taken
for _ in range(3):
    idx, q, tr = ops.screen_trace(z, lay, pcb, margin_tau=tau)
torch.cuda.synchronize()
tr = tr.cpu()
names = ["mma:start", "mma:A ready", "mma:last issued", "ld:ready", "ld:buf free", "ld:stored", "ep:start", "ep:done"]
for cta in (0, 1, 74, 147):
    t = tr[cta]
    t0 = int(t[0][t[0] > 0].min()) if (t[0] > 0).any() else 0
    print(f"--- CTA {cta} (cycles since its first stamp)")
    for sw in range(min(16, t.shape[0])):
        row = [int(v) - t0 if int(v) else -1 for v in t[sw]]
        if all(v == -1 for v in row):
            break
        print(f"sweep {sw:2d}: " + "  ".join(f"{nm}={v:7d}" for nm, v in zip(names, row)))
# aggregate: per-sweep durations on leader CTAs (even blockIdx)
lead = tr[0::2].double()
ok = (lead[:, 1:13, 0] > 0) & (lead[:, 1:13, 2] > 0)
wait_a = (lead[:, 1:13, 1] - lead[:, 1:13, 0])[ok]
mma = (lead[:, 1:13, 2] - lead[:, 1:13, 1])[ok]
per = (lead[:, 2:13, 0] - lead[:, 1:12, 0])[ok[:, 1:] & ok[:, :-1]]
print(f"leader CTAs, sweeps 1-12: wait for A {wait_a.mean():.0f} cyc, A ready -> last MMA issued {mma.mean():.0f} cyc, sweep period {per.mean():.0f} cyc")
ld = tr.double()
okl = (ld[:, 2:13, 3] > 0) & (ld[:, 2:13, 5] > 0)
print(f"loaders, sweeps 2-12: wait for free buffer {(ld[:, 2:13, 4] - ld[:, 2:13, 3])[okl].mean():.0f} cyc, fill {(ld[:, 2:13, 5] - ld[:, 2:13, 4])[okl].mean():.0f} cyc")
ep = tr.double()
oke = (ep[:, 1:13, 6] > 0) & (ep[:, 1:13, 7] > 0)
print(f"epilogue, sweeps 1-12: sweep {(ep[:, 1:13, 7] - ep[:, 1:13, 6])[oke].mean():.0f} cyc; gap to next sweep start {(ep[:, 2:13, 6] - ep[:, 1:12, 7])[oke[:, 1:] & oke[:, :-1]].mean():.0f} cyc")

# per-tile stamps of the fourth sweep (trace[cta][16 + tile][slot]); leader CTA 0 and its peer CTA 1
tn = ["mma:acc free", "mma:B landed", "mma:issued", "ep:acc full", "ep:released", "ep:tile done", "tma:slot free", "tma:issued"]
for cta in (0, 1):
    t = tr[cta][16:32]
    vals = t[t > 0]
    if vals.numel() == 0:
        continue
    t0 = int(vals.min())
    print(f"--- CTA {cta}, sweep 3, per tile (cycles since the first stamp of the sweep)")
    for j in range(16):
        row = [int(v) - t0 if int(v) else -1 for v in t[j]]
        print(f"tile {j:2d}: " + "  ".join(f"{nm}={v:6d}" for nm, v in zip(tn, row)))
