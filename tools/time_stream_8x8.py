"""Per-launch medians of the HBM-bound kernels at an 8x8-layout workload (assign, row-major / channel-major decode
gather) next to a device copy / fill of the same bytes.  usage: [CCVSQ_LIB=...] python tools/time_stream_8x8.py [workload]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ccvs_b200 import ops

wl = sys.argv[1] if len(sys.argv) > 1 else "c3d512"
dev = torch.device("cuda", 0)
(clips, frames), D, h, w_, K, desc = bench.WORKLOADS[wl]
z, cb, n = bench.make_inputs(wl, dev, 1234)
lay = ops.layout_of(z.shape, D, 1)
idx = ops.quantize_forward(z, lay, cb, 0.25, indices_only=True).idx
def med(fn, reps=25):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); ts.append((a, b))
    torch.cuda.synchronize()
    v = sorted(x.elapsed_time(y) for x, y in ts)
    return v[len(v) // 2] * 1e3
zq = torch.empty_like(z); hdr = torch.zeros(K + 4, dtype=torch.int32, device=dev); sq = torch.zeros(1, dtype=torch.float64, device=dev)
dec = torch.empty(n, D, device=dev); err = torch.zeros(1, dtype=torch.int32, device=dev); rl = ops.rows_layout(n, D)
t_a = med(lambda: ops._call("ccvsq_assign", ops._ptr(z), lay, ops._ptr(cb), K, ops._ptr(idx), ops._ptr(zq), ops._ptr(sq), ops._ptr(hdr), ops._stream(dev)))
t_r = med(lambda: ops._call("ccvsq_gather", ops._ptr(idx), ops._ptr(cb), K, rl, ops._ptr(dec), ops._ptr(err), ops._stream(dev)))
t_c = med(lambda: ops._call("ccvsq_gather", ops._ptr(idx), ops._ptr(cb), K, lay, ops._ptr(dec), ops._ptr(err), ops._stream(dev)))
t_copy = med(lambda: zq.copy_(z)); t_fill = med(lambda: dec.zero_())
print(f"{os.environ.get('CCVSQ_LIB', 'default'):28s} {wl}: assign {t_a:.1f} us | gather rows {t_r:.1f} | gather channel-major {t_c:.1f} | copy {t_copy:.1f} | fill {t_fill:.1f}")
