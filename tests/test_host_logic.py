"""Host-side logic that needs no GPU: layout derivation, module surface, state-dict compatibility,
frame sharding, packed statistics all-reduce (world_size-2 gloo)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import vq_oracle
from ccvs_b200 import VectorQuantizer
from ccvs_b200 import dist as vqd
from ccvs_b200 import ops
from ccvs_b200.quantize import EMAVectorQuantizer, LazyOneHot


def test_layout_of_matches_reference_flatten():
    # 4-D / 5-D: rows ordered (g, h, w[, m]); address of (g, c, s) = (g*C + c)*S + s
    for shape, e_dim, mult in [((2, 8, 3, 5), 8, 1), ((2, 3, 8, 4, 4), 8, 1), ((2, 8, 3, 3), 2, 4)]:
        lay = ops.layout_of(shape, e_dim, mult)
        z = torch.arange(float(torch.tensor(shape).prod())).view(shape)
        rows = vq_oracle.to_channel_last(z).reshape(-1, e_dim)
        flat = z.reshape(-1)
        assert lay.rows == rows.shape[0] and lay.dim == e_dim
        for n in (0, 1, lay.rows // 2, lay.rows - 1):
            p, m = divmod(n, lay.mult)
            g, s = divmod(p, lay.S)
            for j in (0, e_dim - 1):
                c = m * lay.dim + j
                assert flat[(g * lay.C + c) * lay.S + s] == rows[n, j]
    lay = ops.layout_of((4, 16, 2), 1, 1)       # state quantizer (state_model.py:57): no transpose
    assert (lay.G, lay.C, lay.S, lay.mult) == (128, 1, 1, 1)
    with pytest.raises(ValueError):
        ops.layout_of((2, 7, 3, 3), 8, 1)


def test_module_surface_and_state_dict():
    vq = VectorQuantizer(1024, 256, 0.25)
    assert (vq.n_e, vq.e_dim, vq.beta, vq.mult, vq.normalize) == (1024, 256, 0.25, 1, False)
    assert [n for n, _ in vq.named_parameters()] == ["embedding.weight"]
    assert list(vq.state_dict().keys()) == ["embedding.weight"]       # models/__init__.py:21 saves this
    assert vq.embedding.weight.abs().max() <= 1.0 / 1024               # quantize.py:30
    sd = {"embedding.weight": torch.randn(1024, 256)}                  # a reference-saved checkpoint
    vq.load_state_dict(sd, strict=True)
    assert torch.equal(vq.embedding.weight, sd["embedding.weight"])
    vq4 = VectorQuantizer(32, 16, 0.25, mult=4)
    assert vq4.e_dim == 4 and vq4.embedding.weight.shape == (32, 4)
    st = VectorQuantizer(128, 1, 0.25)
    assert st.embedding.weight.min() >= 0 and st.embedding.weight.max() <= 1   # quantize.py:27-28


def test_cpu_tensor_raises_no_fallback():
    vq = VectorQuantizer(16, 8, 0.25)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        vq(torch.randn(2, 8, 4, 4))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        vq.embed_code(torch.zeros(2, 4, 4, dtype=torch.int64))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.prepare_codebook(torch.randn(4, 4))


def test_lazy_one_hot():
    idx = torch.tensor([[2], [0], [2]])
    oh = LazyOneHot(idx, 4, torch.float32)
    assert tuple(oh.shape) == (3, 4)
    assert torch.equal(oh.dense(), torch.tensor([[0., 0, 1, 0], [1, 0, 0, 0], [0, 0, 1, 0]]))


def test_ema_module_surface():
    m = EMAVectorQuantizer(32, 8, 0.25, decay=0.9)
    assert not m.embedding.weight.requires_grad
    assert set(m.state_dict()) == {"embedding.weight", "ema_count", "ema_sum"}


def test_frame_shard_partitions():
    for total, world in [(1024 * 16, 8), (10, 4), (3, 8), (0, 2)]:
        blocks = [vqd.frame_shard(total, r, world) for r in range(world)]
        assert blocks[0][0] == 0 and blocks[-1][1] == total
        assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
        sizes = [b - a for a, b in blocks]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        vqd.frame_shard(8, 2, 2)


def test_pack_unpack_roundtrip():
    resid = torch.randn(6, 3)
    counts = torch.tensor([0, 5, 7, 1, 0, 16_000_000], dtype=torch.int32)
    sq = torch.tensor([123456.789012345], dtype=torch.float64)
    r, c, s = vqd.unpack_stats(vqd.pack_stats(resid, counts, sq), 6, 3)
    assert torch.equal(r, resid) and torch.equal(c, counts)
    assert abs(float(s) - float(sq)) < 1e-6


def _shard_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        z, cb = vq_oracle.synth((8, 16, 4, 4), 32, 16, "T", seed=3)
        a, b = vqd.frame_shard(z.shape[0], rank, world)
        zs = z[a:b]
        rows = vq_oracle.to_channel_last(zs).reshape(-1, 16)
        idx = vq_oracle.nearest(rows, cb)
        resid = torch.zeros(32, 16).index_add_(0, idx, rows - cb[idx])
        counts = torch.bincount(idx, minlength=32).to(torch.int32)
        sq = ((cb[idx] - rows) ** 2).sum().double().view(1)
        resid, counts, sq = vqd.all_reduce_stats(resid, counts, sq)
        if rank == 0:
            ret.put((resid.clone(), counts.clone(), sq.clone()))
    finally:
        dist.destroy_process_group()


def test_sharded_statistics_equal_global_gloo():
    """2 ranks, frames sharded, ONE packed all-reduce: the reduced statistics give the same loss,
    perplexity and codebook gradient as the oracle run on the whole batch."""
    ctx = mp.get_context("spawn")
    ret = ctx.SimpleQueue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_shard_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    resid, counts, sq = ret.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    z, cb = vq_oracle.synth((8, 16, 4, 4), 32, 16, "T", seed=3)
    g_zq = torch.zeros_like(z)
    res, _, dE = vq_oracle.forward_backward(z, cb, 0.25, g_zq, 1.0)
    M, N = z.numel(), z.numel() // 16
    loss = (1 + 0.25) * float(sq) / M
    assert abs(loss - float(res.loss)) <= 1e-5 * float(res.loss)
    torch.testing.assert_close(-(2 * 0.25 / M) * resid, dE, rtol=1e-4, atol=1e-7)
    p = counts.float() / N
    perp = torch.exp(-(p * torch.log(p + 1e-10)).sum())
    torch.testing.assert_close(perp, res.perplexity, rtol=1e-5, atol=0)
    assert int(counts.sum()) == N


def _ema_stats_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        K, D = 32, 16
        z, cb = vq_oracle.synth((8, 16, 4, 4), K, D, "T", seed=5)
        a, b = vqd.frame_shard(z.shape[0], rank, world)
        rows = vq_oracle.to_channel_last(z[a:b]).reshape(-1, D)
        idx = vq_oracle.nearest(rows, cb)
        buf, resid_view = vqd.ema_stats_buffer(K, D, "cpu")
        resid_view.zero_().index_add_(0, idx, rows - cb[idx])       # what the assign pass accumulates in place
        counts = torch.bincount(idx, minlength=K).to(torch.int32)
        resid, counts = vqd.reduce_ema_stats(buf, counts, K, D)
        if rank == 0:
            ret.put((resid.clone(), counts.clone()))
    finally:
        dist.destroy_process_group()


def test_ema_statistics_one_all_reduce_gloo():
    """2 ranks: residual sums accumulated straight into the packed buffer + counts, ONE all-reduce; the result equals
    the statistics of the whole batch and drives the textbook EMA update to the same codebook."""
    ctx = mp.get_context("spawn")
    ret = ctx.SimpleQueue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_ema_stats_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    resid, counts = ret.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    K, D = 32, 16
    z, cb = vq_oracle.synth((8, 16, 4, 4), K, D, "T", seed=5)
    rows = vq_oracle.to_channel_last(z).reshape(-1, D)
    idx = vq_oracle.nearest(rows, cb)
    assert counts.dtype == torch.int32 and torch.equal(counts.long(), torch.bincount(idx, minlength=K))
    torch.testing.assert_close(resid, torch.zeros(K, D).index_add_(0, idx, rows - cb[idx]), rtol=1e-5, atol=1e-5)


def _ema_split_worker(rank, world, port, q):
    """start_reduce_ema_stats / finish_reduce_ema_stats (the deferred form used to hide the all-reduce under the
    backward) give the same sums as the one-shot form, and an empty shard joins with zero statistics."""
    import torch.distributed as dist
    from ccvs_b200 import dist as vqd
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    K, D = 6, 4
    buf, resid_view = vqd.ema_stats_buffer(K, D, "cpu")
    g = torch.Generator().manual_seed(100 + rank)
    if rank == 0:
        resid_view.copy_(torch.randn(K, D, generator=g))
        counts = torch.randint(0, 5, (K,), generator=g, dtype=torch.int32)
    else:                      # empty shard: all-zero packed buffer
        buf.zero_()
        counts = torch.zeros(K, dtype=torch.int32)
    mine = (resid_view.clone(), counts.clone())
    work = vqd.start_reduce_ema_stats(buf, counts, K, D)
    resid, cnt = vqd.finish_reduce_ema_stats(work, buf, K, D)
    q.put((rank, mine[0], mine[1], resid.clone(), cnt.clone()))
    dist.barrier()
    dist.destroy_process_group()


def test_deferred_ema_statistics_gloo():
    import socket
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_ema_split_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    got.sort(key=lambda t: t[0])
    want_resid = got[0][1] + got[1][1]
    want_counts = got[0][2] + got[1][2]
    for _, _, _, resid, cnt in got:
        assert torch.equal(resid, want_resid) and torch.equal(cnt, want_counts)
    assert torch.equal(got[1][1], torch.zeros(6, 4))


# ------------------------------------------------------------------------------------------------
# peer exchange: set-up decisions that need no GPU
# ------------------------------------------------------------------------------------------------
def test_peer_exchange_is_off_without_a_process_group_and_validates_the_mode():
    from ccvs_b200.peer import PeerExchange
    assert PeerExchange.create(1024, 256, torch.device("cpu")) is None      # no process group: nothing to exchange
    with pytest.raises(ValueError):
        EMAVectorQuantizer(64, 16, 0.25, exchange="smoke-signals")
    vq = EMAVectorQuantizer(64, 16, 0.25, exchange="nccl")
    assert vq.exchange == "nccl" and vq._peer is None
    from ccvs_b200 import _lib
    L = _lib.load()
    K, D, W = 1024, 256, 8
    per_slot = (K * D + K + 63) // 64 * 64 * 4
    assert L.ccvsq_peer_exchange_bytes(K, D, W) == 1024 + 2 * W * per_slot      # header + 2 parities x W inbox slots
    assert L.ccvsq_peer_exchange_bytes(K, D, 17) == 0 and L.ccvsq_peer_exchange_bytes(0, D, 2) == 0


def _peer_gloo_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ccvs_b200.peer import PeerExchange
    # a CPU device cannot take the peer path: every rank must come back with None WITHOUT entering a collective alone
    ret[rank] = PeerExchange.create(64, 16, torch.device("cpu")) is None
    dist.barrier()
    dist.destroy_process_group()


def test_peer_exchange_declines_on_cpu_ranks_gloo():
    world, port = 2, 33500 + os.getpid() % 2000
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_peer_gloo_worker, args=(world, port, ret), nprocs=world, join=True)
    assert ret[0] and ret[1]
