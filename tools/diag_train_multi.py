"""Multi-rank training-step diagnosis (torchrun): all-reduce latency alone, and the EMA training step by variant with
device time (CUDA events) and host issue time."""
import os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ccvs_b200.quantize import EMAVectorQuantizer
rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
def p(*a):
    if rank == 0: print(*a, flush=True)
buf = torch.randn(1024 * 256 + 1024, device=dev)
for _ in range(10): dist.all_reduce(buf)
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
for _ in range(100): dist.all_reduce(buf)
e1.record(); host = (time.perf_counter() - t0) / 100 * 1e3; torch.cuda.synchronize()
p(f"all_reduce 1.05 MB back to back: device {e0.elapsed_time(e1)/100*1e3:.1f} us each, host issue {host*1e3:.1f} us each")
t0 = time.perf_counter(); e0.record()
for _ in range(100):
    w = dist.all_reduce(buf, async_op=True); w.wait()
e1.record(); host = (time.perf_counter() - t0) / 100 * 1e3; torch.cuda.synchronize()
p(f"all_reduce async + wait:         device {e0.elapsed_time(e1)/100*1e3:.1f} us each, host issue {host*1e3:.1f} us each")
def timed(name, fn, reps=100):
    for _ in range(10): fn()
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter(); e0.record()
    for _ in range(reps): fn()
    e1.record(); host = (time.perf_counter() - t0) / reps * 1e6; torch.cuda.synchronize()
    p(f"{name:70s} device {e0.elapsed_time(e1)/reps*1e3:7.1f} us  host {host:7.1f} us")
def fresh():
    b = torch.empty(1024 * 256 + 1024, device=dev); b.fill_(1.0)
    w = dist.all_reduce(b, async_op=True); w.wait()
timed("fresh torch.empty buffer each time: fill + all_reduce async + wait", fresh)
def fresh_zero_tail():
    b = torch.empty(1024 * 256 + 1024, device=dev); b[:1024 * 256].zero_(); b[1024 * 256:].fill_(3.0)
    w = dist.all_reduce(b, async_op=True); w.wait()
timed("fresh buffer, head zeroed by memset-like op, tail filled", fresh_zero_tail)
big = torch.zeros(64 << 20, device=dev)
def after_big_kernel():
    big.add_(1.0)
    w = dist.all_reduce(buf, async_op=True); w.wait()
timed("persistent buffer after a 256 MB elementwise kernel", after_big_kernel)
z, cb, n = bench.make_inputs("c2", dev, 1234 + rank, cb_seed=1234)
K, D = cb.shape
g_out = torch.randn_like(z); zt = z.detach().clone().requires_grad_(True)
for name, kw in (("overlapped", dict(sync=True, overlap=True)), ("serial", dict(sync=True, overlap=False)), ("no all-reduce", dict(sync=False))):
    vq = EMAVectorQuantizer(K, D, 0.25, decay=0.99, **kw).to(dev).train()
    with torch.no_grad():
        vq.embedding.weight.copy_(cb); vq.ema_sum.copy_(cb); vq.ema_count.fill_(1.0)
    def step():
        zt.grad = None
        z_q, loss, _ = vq(zt)
        torch.autograd.backward([z_q, loss], [g_out, torch.ones_like(loss)])
    for _ in range(5): step()
    vq.sync_codebook(); torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter(); e0.record()
    for _ in range(50): step()
    vq.sync_codebook(); e1.record(); host = (time.perf_counter() - t0) / 50 * 1e3
    torch.cuda.synchronize()
    print(f"[rank {rank}, cpus {sorted(os.sched_getaffinity(0))[:4]}..x{len(os.sched_getaffinity(0))}] {name:14s} device {e0.elapsed_time(e1)/50:.4f} ms/step   host issue {host:.4f} ms/step", flush=True)

# all-reduce right after a forward of the module (no internal collective): persistent buffer vs an inference forward
vq0 = EMAVectorQuantizer(K, D, 0.25, decay=0.99, sync=False).to(dev).train()
with torch.no_grad():
    vq0.embedding.weight.copy_(cb); vq0.ema_sum.copy_(cb); vq0.ema_count.fill_(1.0)
def fwd_then_ar(train_mode):
    vq0.train(train_mode)
    with torch.no_grad():
        vq0(z)
    w = dist.all_reduce(buf, async_op=True); w.wait()
def fwd_only(train_mode):
    vq0.train(train_mode)
    with torch.no_grad():
        vq0(z)
timed("eval forward only", lambda: fwd_only(False), 30)
timed("eval forward + all_reduce(persistent buffer) + wait", lambda: fwd_then_ar(False), 30)
timed("train forward (no_grad: immediate EMA update) only", lambda: fwd_only(True), 30)
timed("train forward + all_reduce(persistent buffer) + wait", lambda: fwd_then_ar(True), 30)
from ccvs_b200 import ops
lay = ops.layout_of(z.shape, D, 1)
def screen_then_ar():
    ops.quantize_forward(z, lay, cb, 0.25, indices_only=True)
    w = dist.all_reduce(buf, async_op=True); w.wait()
timed("search only (screen+rescore+fallback) + all_reduce + wait", screen_then_ar, 30)
timed("search only", lambda: ops.quantize_forward(z, lay, cb, 0.25, indices_only=True), 30)
# phase timing of the overlapped variant: events on the current stream
vq = EMAVectorQuantizer(K, D, 0.25, decay=0.99, sync=True, overlap=True).to(dev).train()
with torch.no_grad():
    vq.embedding.weight.copy_(cb); vq.ema_sum.copy_(cb); vq.ema_count.fill_(1.0)
evs = []
for it in range(30):
    zt.grad = None
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True); c = torch.cuda.Event(enable_timing=True); d = torch.cuda.Event(enable_timing=True)
    a.record()
    z_q, loss, _ = vq(zt)                      # sync of the previous step + forward kernels + all-reduce issue
    b.record()
    work, pbuf = vq._pending
    work.wait()                                # current stream waits for the collective
    c.record()
    torch.autograd.backward([z_q, loss], [g_out, torch.ones_like(loss)])   # dz + (wait: no-op now) + ema update
    d.record()
    evs.append((a, b, c, d))
torch.cuda.synchronize()
f = sum(x[0].elapsed_time(x[1]) for x in evs[5:]) / 25; w = sum(x[1].elapsed_time(x[2]) for x in evs[5:]) / 25; bk = sum(x[2].elapsed_time(x[3]) for x in evs[5:]) / 25
p(f"phases (ms): forward {f:.4f} | wait for all-reduce right after forward {w:.4f} | backward + ema {bk:.4f}")
# the same with a dummy all-reduce on a persistent buffer instead of the module's packed buffer
dist.destroy_process_group()
