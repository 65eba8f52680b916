"""Multi-GPU (NCCL) correctness of the training collective and of the reference-faithful DDP mode.

Needs >= 2 CUDA devices (skipped otherwise; the CPU/gloo versions of the packing logic are in test_host_logic.py).
Reference semantics: one process per GPU, the batch split across ranks (tools/engine.py:56-57,86-89), gradients
averaged by the parent's DDP (tools/engine.py:71-74), scalars all-reduced for logging (tools/engine.py:127-132).
"""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _need_two_gpus():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 CUDA devices")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _spawn(fn, world, *args):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=fn, args=(r, world, port, q) + args) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    return sorted(out, key=lambda t: t[0])


def _init(rank, world, port):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "oracle"))
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    return dist


SHAPE, K, D, STEPS = (8, 256, 16, 16), 1024, 256, 3


def _ema_worker(rank, world, port, q, overlap, exchange="nccl", graphed=False):
    dist = _init(rank, world, port)
    import vq_oracle
    from ccvs_b200.quantize import EMAVectorQuantizer
    dev = torch.device("cuda", rank)
    z_all, cb = vq_oracle.synth(SHAPE, K, D, "T", seed=77)
    g_all = torch.randn(SHAPE, generator=torch.Generator().manual_seed(78))
    per = SHAPE[0] // world
    z = z_all[rank * per:(rank + 1) * per].to(dev).requires_grad_(True)
    g = g_all[rank * per:(rank + 1) * per].to(dev)
    vq = EMAVectorQuantizer(K, D, 0.25, decay=0.9, sync=True, overlap=overlap, exchange=exchange).to(dev).train()

    def reset():
        with torch.no_grad():
            vq.sync_codebook()
            vq.embedding.weight.copy_(cb.to(dev))
            vq.ema_sum.copy_(cb.to(dev))
            vq.ema_count.fill_(1.0)

    reset()
    grads = []
    if graphed:
        from ccvs_b200.quantize import GraphedTrainStep
        gs = GraphedTrainStep(vq, z.detach().clone(), g.clone(), warmup=3)   # (its warm-up steps move the codebook:
        assert vq._peer is not None and vq._peer.steps >= 3                  #  start again from the initial state)
        reset()
        for _ in range(STEPS):
            gs.replay()
            grads.append(gs.dz.detach().cpu().clone())
    else:
        for _ in range(STEPS):
            z.grad = None
            z_q, loss, _ = vq(z)
            torch.autograd.backward([z_q, loss], [g, torch.ones_like(loss)])
            grads.append(z.grad.detach().cpu().clone())
        vq.sync_codebook()
    assert (vq._peer is not None) == (exchange == "peer")
    torch.cuda.synchronize()
    q.put((rank, vq.embedding.weight.detach().cpu(), vq.ema_count.cpu(), vq.ema_sum.cpu(), torch.stack(grads)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("overlap,exchange,graphed", [(True, "nccl", False), (False, "nccl", False), (True, "peer", False),
                                                      (False, "peer", False), (True, "peer", True)])
def test_ema_training_two_ranks_equal_one_gpu_on_the_whole_batch(overlap, exchange, graphed):
    """The training collective — NCCL all-reduce of the packed statistics, or the NVLink peer exchange fused into the EMA
    update (eager and replayed from one CUDA graph per step): after the same steps every rank holds the SAME codebook /
    EMA state (bit for bit — they apply the same summed statistics), and it equals the single-GPU run on the
    concatenated batch up to the summation order of the FP32 statistics."""
    _need_two_gpus()
    got = _spawn(_ema_worker, 2, overlap, exchange, graphed)
    (_, w0, n0, s0, dz0), (_, w1, n1, s1, dz1) = got
    assert torch.equal(w0, w1) and torch.equal(n0, n1) and torch.equal(s0, s1)

    import vq_oracle
    from ccvs_b200.quantize import EMAVectorQuantizer
    dev = torch.device("cuda", 0)
    z_all, cb = vq_oracle.synth(SHAPE, K, D, "T", seed=77)
    g_all = torch.randn(SHAPE, generator=torch.Generator().manual_seed(78)).to(dev)
    z = z_all.to(dev).requires_grad_(True)
    vq = EMAVectorQuantizer(K, D, 0.25, decay=0.9, sync=False).to(dev).train()
    with torch.no_grad():
        vq.embedding.weight.copy_(cb.to(dev))
        vq.ema_sum.copy_(cb.to(dev))
        vq.ema_count.fill_(1.0)
    for step in range(STEPS):
        z.grad = None
        z_q, loss, _ = vq(z)
        torch.autograd.backward([z_q, loss], [g_all, torch.ones_like(loss)])
        # loss of a rank is the mean over ITS shard: dz of the 2-rank run carries 2/M_shard = world x 2/M_whole
        per = SHAPE[0] // 2
        full = z.grad.detach().cpu()
        for r, dz in ((0, dz0), (1, dz1)):
            ste = g_all[r * per:(r + 1) * per].cpu()
            # (the commitment term is ~1e-6 next to an O(1) straight-through term: compare it in float64, to the FP32
            #  rounding of the sums it was extracted from)
            torch.testing.assert_close((dz[step].double() - ste.double()) / 2, full[r * per:(r + 1) * per].double() - ste.double(),
                                       rtol=1e-3, atol=5e-7)
    torch.testing.assert_close(vq.embedding.weight.detach().cpu(), w0, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(vq.ema_count.cpu(), n0, rtol=1e-6, atol=0)
    torch.testing.assert_close(vq.ema_sum.cpu(), s0, rtol=1e-5, atol=1e-6)


class _Parent(torch.nn.Module):
    """A stand-in for the reference's autoencoder around the quantizer: a 1x1 conv producing z, the quantizer, a 1x1
    conv consuming z_q (skip_autoencoder.py:331,368); loss = reconstruction + quantizer loss."""

    def __init__(self, vq):
        super().__init__()
        self.enc = torch.nn.Conv2d(D, D, 1)
        self.net_q = vq
        self.dec = torch.nn.Conv2d(D, 8, 1)

    def forward(self, x):
        z = self.enc(x)
        z_q, q_loss, _ = self.net_q(z)
        return self.dec(z_q).pow(2).mean() + q_loss


def _ddp_worker(rank, world, port, q):
    dist = _init(rank, world, port)
    import vq_oracle
    from ccvs_b200 import VectorQuantizer
    dev = torch.device("cuda", rank)
    torch.manual_seed(5)
    x_all, cb = vq_oracle.synth(SHAPE, K, D, "T", seed=91)
    vq = VectorQuantizer(K, D, 0.25)
    model = _Parent(vq)
    with torch.no_grad():
        vq.embedding.weight.copy_(cb)
        model.enc.weight.copy_(torch.eye(D).view(D, D, 1, 1))
        model.enc.bias.zero_()
    model = model.to(dev)
    ddp = torch.nn.parallel.DistributedDataParallel(model, device_ids=[rank])
    per = SHAPE[0] // world
    loss = ddp(x_all[rank * per:(rank + 1) * per].to(dev))
    loss.backward()
    torch.cuda.synchronize()
    q.put((rank, model.net_q.embedding.weight.grad.cpu(), model.enc.weight.grad.cpu(), float(loss)))
    dist.barrier()
    dist.destroy_process_group()


def test_reference_faithful_ddp_mode_two_ranks():
    """`embedding.weight.grad` is a dense FP32 [K, D] tensor produced by our backward, so the parent's DDP averages it
    like any parameter (tools/engine.py:71-74): the averaged gradient equals the single-GPU gradient of the mean of
    the two shard losses (== the whole-batch loss for equal shards)."""
    _need_two_gpus()
    got = _spawn(_ddp_worker, 2)
    (_, dE0, dW0, l0), (_, dE1, dW1, l1) = got
    assert torch.equal(dE0, dE1) and torch.equal(dW0, dW1)       # DDP leaves identical gradients on every rank

    import vq_oracle
    from ccvs_b200 import VectorQuantizer
    dev = torch.device("cuda", 0)
    x_all, cb = vq_oracle.synth(SHAPE, K, D, "T", seed=91)
    # same decoder initialisation as the workers: they seed before building the model
    torch.manual_seed(5)
    vq2 = VectorQuantizer(K, D, 0.25)
    model = _Parent(vq2)
    with torch.no_grad():
        vq2.embedding.weight.copy_(cb)
        model.enc.weight.copy_(torch.eye(D).view(D, D, 1, 1))
        model.enc.bias.zero_()
    model = model.to(dev)
    per = SHAPE[0] // 2
    loss = 0.5 * (model(x_all[:per].to(dev)) + model(x_all[per:].to(dev)))
    loss.backward()
    assert abs(float(loss) - 0.5 * (l0 + l1)) <= 1e-5 * abs(float(loss))
    torch.testing.assert_close(model.net_q.embedding.weight.grad.cpu(), dE0, rtol=1e-4, atol=1e-7)
    torch.testing.assert_close(model.enc.weight.grad.cpu(), dW0, rtol=1e-3, atol=1e-6)


def _empty_shard_worker(rank, world, port, q, exchange):
    dist = _init(rank, world, port)
    import vq_oracle
    from ccvs_b200.quantize import EMAVectorQuantizer
    dev = torch.device("cuda", rank)
    z_all, cb = vq_oracle.synth(SHAPE, K, D, "T", seed=79)
    vq = EMAVectorQuantizer(K, D, 0.25, decay=0.9, sync=True, exchange=exchange).to(dev).train()
    with torch.no_grad():
        vq.embedding.weight.copy_(cb.to(dev))
        vq.ema_sum.copy_(cb.to(dev))
        vq.ema_count.fill_(1.0)
    for step in range(3):
        # rank 1 has nothing in step 1 (a ragged last batch): it must still join the exchange, with zero statistics
        z = (z_all if not (rank == 1 and step == 1) else z_all[:0]).to(dev).requires_grad_(True)
        z_q, loss, _ = vq(z)
        if z.numel():
            torch.autograd.backward([z_q, loss], [torch.ones_like(z_q), torch.ones_like(loss)])
    vq.sync_codebook()
    torch.cuda.synchronize()
    q.put((rank, vq.embedding.weight.detach().cpu(), vq.ema_count.cpu()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("exchange", ["peer", "nccl"])
def test_an_empty_shard_still_joins_the_exchange(exchange):
    """A rank whose shard is empty in one step publishes zero statistics: the ranks stay in lock step and end with the
    same finite codebook (round-1 advisor finding: the exchange used to sum uninitialised memory)."""
    _need_two_gpus()
    (_, w0, n0), (_, w1, n1) = _spawn(_empty_shard_worker, 2, exchange)
    assert torch.equal(w0, w1) and torch.equal(n0, n1)
    assert bool(torch.isfinite(w0).all()) and bool(torch.isfinite(n0).all())
