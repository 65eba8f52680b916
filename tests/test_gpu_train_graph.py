"""Single-GPU checks of the training-step plumbing added for the multi-GPU path: the peer-exchange kernels through the
C ABI with a world of one rank (publish + fused EMA update == ccvsq_ema_update_packed), and `GraphedTrainStep` (one CUDA
graph per training step) against the eager step.  The 2-rank versions are in test_gpu_multi.py."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


def _state(K, D, dev, seed):
    g = torch.Generator().manual_seed(seed)
    E = torch.randn(K, D, generator=g).to(dev)
    n_ema = (torch.rand(K, generator=g) * 5 + 0.1).to(dev)
    sum_ema = (E.cpu() * n_ema.cpu()[:, None]).to(dev)
    packed = torch.cat([torch.randn(K * D, generator=g), torch.randint(0, 9, (K,), generator=g).float()]).to(dev)
    return E, n_ema, sum_ema, packed


@pytest.mark.parametrize("K,D", [(1024, 256), (96, 64), (7, 4)])
def test_peer_exchange_world_of_one_equals_packed_update(K, D):
    from ccvs_b200 import _lib, ops
    L = _lib.load()
    dev = torch.device("cuda", 0)
    nbytes = int(L.ccvsq_peer_exchange_bytes(K, D, 1))
    assert nbytes >= 1024 + 2 * (K * D + K) * 4
    area = ctypes.c_void_p(0)
    handle = (ctypes.c_ubyte * 64)()
    _lib.check(L.ccvsq_peer_alloc(nbytes, ctypes.byref(area), handle), "peer_alloc")
    assert any(handle)
    try:
        areas = (ctypes.c_void_p * 1)(area.value)
        ref = _state(K, D, dev, 3)
        got = [t.clone() for t in ref]
        for step in range(3):          # both parities, and the device-side step counter
            ops.ema_update_packed(ref[0], ref[1], ref[2], ref[3], 0.9, 1e-5)
            ops._call("ccvsq_peer_publish", ops._ptr(got[3]), K, D, areas, 0, 1, ops._stream(dev))
            ops._call("ccvsq_peer_ema_update", ops._ptr(got[0]), ops._ptr(got[1]), ops._ptr(got[2]), area, K, D, 1, 0.9, 1e-5,
                      ops._stream(dev))
            torch.cuda.synchronize()
            for a, b in zip(ref[:3], got[:3]):
                torch.testing.assert_close(a, b, rtol=2e-6, atol=1e-7)
            ref[3].mul_(0.5).add_(1.0)
            got[3].copy_(ref[3])
    finally:
        torch.cuda.synchronize()
        _lib.check(L.ccvsq_peer_free(area), "peer_free")


def test_peer_exchange_rejects_bad_arguments():
    from ccvs_b200 import _lib
    L = _lib.load()
    assert L.ccvsq_peer_exchange_bytes(1024, 256, 0) == 0
    assert L.ccvsq_peer_exchange_bytes(1024, 256, 17) == 0
    areas = (ctypes.c_void_p * 1)(None)
    buf = torch.zeros(1024 * 256 + 1024, device="cuda")
    assert L.ccvsq_peer_publish(buf.data_ptr(), 1024, 256, areas, 0, 1, None) != 0          # null area
    assert L.ccvsq_peer_publish(buf.data_ptr(), 1024, 256, areas, 1, 1, None) != 0          # rank outside the world


@pytest.mark.parametrize("ema", [True, False])
def test_graphed_train_step_equals_eager(ema):
    """Replays of `GraphedTrainStep` give the eager training step's outputs: z_q, loss, perplexity, indices, dz, and the
    codebook trajectory (EMA variant) or dE (reference-faithful variant, codebook trained by the caller's optimizer)."""
    from ccvs_b200.quantize import EMAVectorQuantizer, GraphedTrainStep, VectorQuantizer
    dev = torch.device("cuda", 0)
    K, D, shape = 512, 64, (4, 64, 16, 16)
    g = torch.Generator().manual_seed(11)
    cb = torch.randn(K, D, generator=g)
    zs = [(cb[torch.randint(0, K, (4 * 256,), generator=g)] + 0.3 * torch.randn(4 * 256, D, generator=g))
          .view(4, 16, 16, D).permute(0, 3, 1, 2).contiguous().to(dev) for _ in range(3)]
    gz = [torch.randn(shape, generator=g).to(dev) for _ in range(3)]

    def make():
        vq = (EMAVectorQuantizer(K, D, 0.25, decay=0.9, deterministic=True) if ema else VectorQuantizer(K, D, 0.25, deterministic=True))
        vq = vq.to(dev).train()
        return vq

    def reset(vq):
        with torch.no_grad():
            if ema:
                vq.sync_codebook()
                vq.ema_sum.copy_(cb.to(dev))
                vq.ema_count.fill_(1.0)
            vq.embedding.weight.copy_(cb.to(dev))

    eager, outs_e = make(), []
    reset(eager)
    for z, gq in zip(zs, gz):
        zt = z.clone().requires_grad_(True)
        eager.embedding.weight.grad = None
        z_q, loss, (perp, _, idx) = eager(zt)
        torch.autograd.backward([z_q, loss], [gq, torch.ones_like(loss)])
        if ema:
            eager.sync_codebook()
        outs_e.append([t.detach().clone() for t in (z_q, loss, perp, idx.view(-1), zt.grad)] +
                      [eager.embedding.weight.detach().clone() if ema else eager.embedding.weight.grad.clone()])

    vq = make()
    reset(vq)
    step = GraphedTrainStep(vq, zs[0].clone(), gz[0].clone())
    reset(vq)
    for i, (z, gq) in enumerate(zip(zs, gz)):
        step.z.data.copy_(z)
        step.grad_zq.copy_(gq)
        z_q, loss, perp, idx, dz = step.replay()
        torch.cuda.synchronize()
        last = vq.embedding.weight.detach() if ema else step.dE
        for name, a, b in zip(("z_q", "loss", "perplexity", "indices", "dz", "codebook / dE"), outs_e[i],
                              (z_q, loss, perp, idx.view(-1), dz, last)):
            assert torch.equal(a, b), f"step {i}: {name} differs between the eager and the graphed step"
