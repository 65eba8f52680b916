"""Diagnostic: device / host time of the composite forward vs the step-by-step entry points."""
import sys, time, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ccvs_b200 import ops, VectorQuantizer

wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
dev = torch.device("cuda", 0)
(clips, frames), D, h, w_, K, desc = bench.WORKLOADS[wl]
z, cb, n = bench.make_inputs(wl, dev, 1234)
lay = ops.layout_of(z.shape, D, 1)
w = cb.contiguous()


def timeit(name, fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    host = (time.perf_counter() - t0) / reps * 1e3
    torch.cuda.synchronize()
    print(f"{name:40s} device {e0.elapsed_time(e1) / reps:8.4f} ms   host-issue {host:8.4f} ms", flush=True)


pcb = ops.prepare_codebook(w)
timeit("composite full (own cb)", lambda: ops.quantize_forward(z, lay, w, 0.25))
timeit("composite full (cached cb)", lambda: ops.quantize_forward(z, lay, w, 0.25, cb=pcb))
timeit("composite indices_only (cached cb)", lambda: ops.quantize_forward(z, lay, w, 0.25, cb=pcb, indices_only=True))
idx = ops.quantize_forward(z, lay, w, 0.25, cb=pcb, indices_only=True).idx
timeit("stepwise search", lambda: ops.search(z, lay, pcb))
timeit("stepwise assign", lambda: ops.assign(z, lay, w, idx))
timeit("prepare", lambda: ops.prepare_codebook(w))
timeit("gather rows", lambda: ops.gather(idx, w))
timeit("gather cm", lambda: ops.gather(idx, w, lay))
g = torch.randn_like(z)
one = torch.ones((), device=dev)
timeit("backward fused dz+dE", lambda: ops.quantize_backward(z, lay, w, idx, g, one, 0.25))
timeit("backward dz only", lambda: ops.quantize_backward(z, lay, w, idx, g, one, 0.25, want_dE=False))
timeit("backward dE only", lambda: ops.quantize_backward(z, lay, w, idx, None, one, 0.25, want_dz=False))
timeit("code_stats", lambda: ops.code_stats(z, lay, w, K, idx, 1.0))
timeit("empty ws", lambda: torch.empty(60 << 20, dtype=torch.uint8, device=dev))
z2 = torch.empty_like(z)
timeit("yardstick: torch copy (r+w)", lambda: z2.copy_(z))
timeit("yardstick: torch fill (w)", lambda: z2.fill_(1.0))
timeit("assign, no zq write", lambda: ops.assign(z, lay, w, idx, want_zq=False))
idx0 = torch.zeros_like(idx)
timeit("assign, all codes = 0 (E from L1)", lambda: ops.assign(z, lay, w, idx0))
timeit("assign, codes = 0, no zq write", lambda: ops.assign(z, lay, w, idx0, want_zq=False))
timeit("gather cm, all codes = 0", lambda: ops.gather(idx0, w, lay))
timeit("gather rows, all codes = 0", lambda: ops.gather(idx0, w))
timeit("assign, no counts", lambda: ops.assign(z, lay, w, idx, want_counts=False))
timeit("assign, no counts, no zq write", lambda: ops.assign(z, lay, w, idx, want_zq=False, want_counts=False))
