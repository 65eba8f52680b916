"""One launch of every HBM-bound kernel at a workload's shape (for `ncu --set full -k regex:cm4|rows4`)."""
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
w = cb.contiguous()
idx = ops.quantize_forward(z, lay, w, 0.25, indices_only=True).idx
g = torch.randn_like(z)
one = torch.ones((), device=dev)
for _ in range(2):
    ops.assign(z, lay, w, idx)
    ops.gather(idx, w)
    ops.gather(idx, w, lay)
    ops.quantize_backward(z, lay, w, idx, g, one, 0.25)
    ops.quantize_backward(z, lay, w, idx, g, one, 0.25, want_dE=False)
    ops.code_stats(z, lay, w, K, idx, 1.0)
torch.cuda.synchronize()
