"""Which of our kernels makes a following NCCL all-reduce slow?  (torchrun, 2 ranks)"""
import os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ccvs_b200 import ops
rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
def p(*a):
    if rank == 0: print(*a, flush=True)
buf = torch.randn(1024 * 256 + 1024, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
def timed(name, fn, reps=30):
    for _ in range(5): fn()
    torch.cuda.synchronize(); dist.barrier()
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    p(f"{name:60s} {e0.elapsed_time(e1)/reps*1e3:8.1f} us")
def ar():
    w = dist.all_reduce(buf, async_op=True); w.wait()
z, cb, n = bench.make_inputs("c2", dev, 1234 + rank)
K, D = cb.shape
lay = ops.layout_of(z.shape, D, 1)
pcb = ops.prepare_codebook(cb)
idx, q = ops.screen(z, lay, pcb)
idx = ops.rescore(z, lay, pcb, idx, q)
zq = torch.empty_like(z); sq = torch.zeros(1, dtype=torch.float64, device=dev); cnt = torch.zeros(K, dtype=torch.int32, device=dev)
qq = ops._new_queue(n, 4, dev); idx2 = torch.empty(n, dtype=torch.int64, device=dev)
def k_prepare(): ops.prepare_codebook(cb)
def k_screen():
    qq.count.zero_()
    ops._call("ccvsq_screen", ops._ptr(z), lay, ops._ptr(pcb.e_bf16), ops._ptr(pcb.e_max), K, 1.0, 4, ops._ptr(idx2), ops._ptr(qq.count), ops._ptr(qq.rows), ops._ptr(qq.cand), ops._ptr(qq.flags), ops._stream(dev))
def k_exact_small(): ops.search_exact(z[:1], ops.layout_of(z[:1].shape, D, 1), pcb)
def k_assign(): ops._call("ccvsq_assign", ops._ptr(z), lay, ops._ptr(cb), K, ops._ptr(idx), ops._ptr(zq), ops._ptr(sq), ops._ptr(cnt), ops._stream(dev))
def k_gather(): ops.gather(idx, cb)
def k_torch(): zq.add_(1.0)
for name, k in (("nothing", lambda: None), ("torch add_ 268 MB", k_torch), ("prepare_codebook", k_prepare), ("screen", k_screen), ("exact search (256 rows)", k_exact_small),
                ("assign", k_assign), ("gather", k_gather)):
    timed(f"{name}: kernel only", k)
    timed(f"{name}: kernel + all_reduce + wait", lambda: (k(), ar()))
    timed(f"{name}: kernel + tiny torch kernel + all_reduce + wait", lambda: (k(), sq.zero_(), ar()))

p("---- composite forward variants")
keep = []
def comp(**kw):
    return ops.quantize_forward(z, lay, cb, 0.25, **kw)
for name, k in (("composite indices_only", lambda: comp(indices_only=True)),
                ("composite indices_only, cached codebook", lambda: comp(indices_only=True, cb=pcb)),
                ("composite indices_only, no exact fallback", lambda: comp(indices_only=True, exact_fallback=False)),
                ("composite full", lambda: comp()),
                ("composite exact mode, 4096 rows", lambda: ops.quantize_forward(z[:1], ops.layout_of(z[:1].shape, D, 1), cb, 0.25, mode="exact")),
                ("composite full, outputs kept alive", lambda: keep.append(comp()) or (len(keep) > 4 and keep.pop(0)))):
    timed(f"{name}: only", k)
    timed(f"{name}: + all_reduce + wait", lambda: (k(), ar()))
    timed(f"{name}: + tiny torch kernel + all_reduce + wait", lambda: (k(), sq.zero_(), ar()))
    timed(f"{name}: + torch.cuda.current_stream().synchronize() + all_reduce + wait", lambda: (k(), torch.cuda.current_stream().synchronize(), ar()))
dist.destroy_process_group()
