"""Peer exchange, piece by piece (torchrun, >= 2 ranks): device time of the publish kernel alone, of the fused EMA update
with all flags already in place, of the plain packed update, and of the graphed training step by variant."""
import os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ccvs_b200 import ops
from ccvs_b200.peer import PeerExchange
from ccvs_b200.quantize import EMAVectorQuantizer, GraphedTrainStep
rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
def p(*a):
    if rank == 0: print(*a, flush=True)
K, D = 1024, 256
px = PeerExchange.create(K, D, dev)
assert px is not None
E = torch.randn(K, D, device=dev); n_ema = torch.ones(K, device=dev); s_ema = E.clone()
stats = torch.randn(K * D + K, device=dev).abs()
def ev(): return torch.cuda.Event(enable_timing=True)
tp, tu, tk = [], [], []
for it in range(30):
    torch.cuda.synchronize(); dist.barrier()
    a, b, c, d, e, f = ev(), ev(), ev(), ev(), ev(), ev()
    a.record(); px.publish(stats, overlap=False); b.record()
    torch.cuda.synchronize(); dist.barrier()           # every rank has published: no waiting inside the update
    c.record(); px.ema_update(E, n_ema, s_ema, 0.99, 1e-5); d.record()
    e.record(); ops.ema_update_packed(E, n_ema, s_ema, stats, 0.99, 1e-5); f.record()
    torch.cuda.synchronize()
    if it >= 5:
        tp.append(a.elapsed_time(b)); tu.append(c.elapsed_time(d)); tk.append(e.elapsed_time(f))
med = lambda v: sorted(v)[len(v) // 2] * 1e3
p(f"world {world}: publish alone {med(tp):.1f} us | fused peer update (flags present) {med(tu):.1f} us | packed update (1 rank's stats) {med(tk):.1f} us")
# back-to-back publish+update pairs in lock step (no other work): the pure exchange cycle
torch.cuda.synchronize(); dist.barrier()
a, b = ev(), ev(); a.record()
for _ in range(200):
    px.publish(stats, overlap=False); px.ema_update(E, n_ema, s_ema, 0.99, 1e-5)
b.record(); torch.cuda.synchronize()
p(f"publish + update back to back, lock step: {a.elapsed_time(b) / 200 * 1e3:.1f} us per pair")
z, cb, n = bench.make_inputs("c2", dev, 1234 + rank, cb_seed=1234)
g_out = torch.randn_like(z)
for name, kw in (("graphed peer overlap", dict(sync=True, overlap=True)), ("graphed peer, publish on the main stream", dict(sync=True, overlap=True)),
                 ("graphed no exchange", dict(sync=False))):
    vq = EMAVectorQuantizer(K, D, 0.25, decay=0.99, **kw).to(dev).train()
    with torch.no_grad():
        vq.embedding.weight.copy_(cb); vq.ema_sum.copy_(cb); vq.ema_count.fill_(1.0)
    if "main stream" in name:
        import ccvs_b200.peer as peer_mod
        orig = peer_mod.PeerExchange.publish
        peer_mod.PeerExchange.publish = lambda self, stats, overlap=True: orig(self, stats, overlap=False)
    gs = GraphedTrainStep(vq, z.detach(), g_out, warmup=3)
    if "main stream" in name:
        peer_mod.PeerExchange.publish = orig
    for _ in range(5): gs.replay()
    torch.cuda.synchronize(); dist.barrier()
    a, b = ev(), ev(); a.record()
    for _ in range(100): gs.replay()
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / 100], device=dev); allt = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(allt, t)
    p(f"{name:45s} ms/step by rank: " + " ".join(f"{float(x):.4f}" for x in allt))
    del gs, vq
dist.barrier(); dist.destroy_process_group()
