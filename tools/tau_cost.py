"""Cost of the screening margin: forward time, queue length and fallback rows by margin_tau, distributions T and I.
usage: python tools/tau_cost.py [workload]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import bench
from ccvs_b200 import ops

wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
dev = torch.device("cuda", 0)
(clips, frames), D, h, w_, K, desc = bench.WORKLOADS[wl]
for dist in ("T", "I"):
    if dist == "T":
        z, cb, n = bench.make_inputs(wl, dev, 1234)
    else:
        g = torch.Generator(device=dev).manual_seed(7)
        cb = (torch.rand(K, D, generator=g, device=dev) * 2 - 1) / K
        z = torch.randn(clips, frames, D, h, w_, generator=g, device=dev)
        n = clips * frames * h * w_
    lay = ops.layout_of(z.shape, D, 1)
    pcb = ops.prepare_codebook(cb.contiguous())
    ref = ops.search_exact(z, lay, pcb)
    for tau in (1.0, 2.0, 3.0, 4.0):
        idx, q = ops.screen(z, lay, pcb, 4, tau)
        nq = int(q.count.item())
        nflag = int((q.flags[:nq] != 0).sum())
        def fwd():
            return ops.quantize_forward(z, lay, cb, 0.25, margin_tau=tau, cb=pcb, indices_only=True)
        for _ in range(3):
            o = fwd()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            o = fwd()
        e1.record()
        torch.cuda.synchronize()
        same = float((o.idx == ref).float().mean())
        print(f"{wl} dist {dist} tau {tau}: search {e0.elapsed_time(e1) / 10 * 1e3:9.1f} us   queued {nq:8d} ({100.0 * nq / n:6.2f} %)  "
              f"flagged (exact fallback) {nflag:8d} ({100.0 * nflag / n:6.2f} %)   equal to the FP32 CUDA-core search: {same:.6f}", flush=True)
