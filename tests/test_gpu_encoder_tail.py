"""GPU parity of the encoder tail: the tcgen05 BF16x6 1x1-convolution kernel against fixtures produced by
the reference's own ConvLayer source, against a float64 evaluation (accuracy class: FP32), and chained into the
quantizer against the oracle of the same chain.

Tolerance (floating point, written here as the task requires): every output within 4e-7 * (sum_c |Ws[o,c] x[c]| + |b|)
of the float64 value — about three FP32 ulps of the accumulated magnitude, what torch-CPU and cuDNN FP32 convolutions
reach on the same inputs (tools/tail_accuracy.py: 1.2e-7 ... 2.7e-7 for all three) — and within rtol 1e-5 / atol 1e-5 of
the reference's own FP32 output."""
import pytest
import torch

import vq_oracle
from ccvs_b200 import EncoderTail, VectorQuantizer, ops

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _f64(x, weight, bias, slope=0.1):
    ws = (weight * (1 / weight.shape[1] ** 0.5)).double()[:, :, 0, 0]              # fl32(W * scale), then exact
    pre = torch.einsum("oc,gchw->gohw", ws, x.double()) + bias.double().view(1, -1, 1, 1)
    mag = torch.einsum("oc,gchw->gohw", ws.abs(), x.double().abs()) + bias.double().abs().view(1, -1, 1, 1)
    return torch.where(pre > 0, pre, pre * slope), mag


def _module(g, normalize=False):
    m = EncoderTail(g.weight.shape[1], g.weight.shape[0], normalize_out=normalize).to(DEV)
    with torch.no_grad():
        m.weight.copy_(g.weight.to(DEV))
        m.bias.copy_(g.bias.to(DEV))
    return m


def test_encoder_tail_matches_reference_fixture(tail_golden):
    g = tail_golden
    m = _module(g)
    with torch.no_grad():
        z = m(g.x.to(DEV)).cpu()
    assert z.shape == g.out.shape
    torch.testing.assert_close(z, g.out, rtol=1e-5, atol=1e-5)
    ref64, mag = _f64(g.x, g.weight, g.bias)
    err = (z.double() - ref64).abs()
    assert bool((err <= 4e-7 * mag + 1e-12).all()), float((err / mag).max())
    mn = _module(g, normalize=True)
    with torch.no_grad():
        zn = mn(g.x.to(DEV)).cpu()
    torch.testing.assert_close(zn, g.out_normalized, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("G,ci,co,h,w", [(5, 64, 16, 3, 3), (64, 512, 256, 8, 8), (16, 256, 512, 16, 16), (7, 192, 48, 4, 6)])
def test_encoder_tail_vs_oracle_shapes(G, ci, co, h, w):
    """Ragged / multi-tile shapes (several row tiles per CTA pair, two channel tiles, fewer positions than a tile)."""
    gen = torch.Generator().manual_seed(G * 1000 + ci + co)
    x = torch.randn(G, ci, h, w, generator=gen) * 2
    weight = torch.randn(co, ci, 1, 1, generator=gen)
    bias = torch.randn(co, generator=gen)
    m = EncoderTail(ci, co).to(DEV)
    with torch.no_grad():
        m.weight.copy_(weight.to(DEV))
        m.bias.copy_(bias.to(DEV))
        z = m(x.to(DEV)).cpu()
    ref64, mag = _f64(x, weight, bias)
    err = (z.double() - ref64).abs()
    assert bool((err <= 4e-7 * mag + 1e-12).all()), float((err / mag).max())
    torch.testing.assert_close(z, vq_oracle.encoder_tail(x, weight, bias), rtol=1e-5, atol=1e-5)
    # 5-D input [B, T, C, h, w] is flattened like the reference's flatten_vid
    if G % 2 == 0:
        with torch.no_grad():
            z5 = m(x.view(2, G // 2, ci, h, w).to(DEV)).cpu()
        assert z5.shape == (2, G // 2, co, h, w) and torch.equal(z5.view(G, co, h, w), z)


def test_encoder_tail_into_quantizer_matches_the_oracle_chain():
    """encoder tail -> quantizer, both on our kernels, against oracle(encoder_tail) -> oracle(quantizer): indices equal
    except documented near-ties (the latents differ from the FP32 reference by a few ulps)."""
    gen = torch.Generator().manual_seed(77)
    G, ci, D, h, w, K = 16, 512, 256, 8, 8, 1024
    x = torch.randn(G, ci, h, w, generator=gen)
    weight = torch.randn(D, ci, 1, 1, generator=gen)
    bias = 0.1 * torch.randn(D, generator=gen)
    z_ref = vq_oracle.encoder_tail(x, weight, bias)
    rows = vq_oracle.to_channel_last(z_ref).reshape(-1, D)
    cb = rows[torch.randperm(rows.shape[0], generator=gen)[:K]] + 0.05 * torch.randn(K, D, generator=gen)   # codes near the data
    tail = EncoderTail(ci, D).to(DEV)
    vq = VectorQuantizer(K, D, 0.25).to(DEV).eval()
    with torch.no_grad():
        tail.weight.copy_(weight.to(DEV)); tail.bias.copy_(bias.to(DEV)); vq.embedding.weight.copy_(cb.to(DEV))
        z = tail(x.to(DEV))
        z_q, loss, (_, _, idx) = vq(z)
    par = vq_oracle.classify_indices(idx, rows, cb)
    assert par.mismatch == 0, par
    res = vq_oracle.forward(z_ref, cb, 0.25)
    torch.testing.assert_close(loss.cpu(), res.loss, rtol=1e-5, atol=0)


def test_encoder_tail_backward_and_errors():
    gen = torch.Generator().manual_seed(5)
    x = torch.randn(3, 64, 4, 4, generator=gen)
    m = EncoderTail(64, 32).to(DEV)
    xr = x.to(DEV).requires_grad_(True)
    z = m(xr)
    (z * torch.randn(z.shape, generator=gen).to(DEV)).sum().backward()
    assert xr.grad is not None and m.weight.grad is not None and m.bias.grad is not None
    # against autograd of the restatement
    x2 = x.clone().requires_grad_(True)
    w2 = m.weight.detach().cpu().clone().requires_grad_(True)
    b2 = m.bias.detach().cpu().clone().requires_grad_(True)
    gen = torch.Generator().manual_seed(5); torch.randn(3, 64, 4, 4, generator=gen)
    gz = torch.randn(z.shape, generator=gen)
    (vq_oracle.encoder_tail(x2, w2, b2) * gz).sum().backward()
    torch.testing.assert_close(xr.grad.cpu(), x2.grad, rtol=1e-3, atol=1e-4)
    torch.testing.assert_close(m.weight.grad.cpu(), w2.grad, rtol=1e-3, atol=1e-4)
    with pytest.raises(RuntimeError):
        m.cpu()(x)                                  # no CPU fallback
    with pytest.raises(ValueError):
        EncoderTail(100, 64)                        # C_in not a multiple of 64
    assert not ops.encoder_tail_supported(64, 24 + 2)


def test_reencode_step_graph_equals_eager_over_a_rollout():
    """The autoregressive re-encode step (vid_step_decode, quantized_video_model.py:939-964) as one CUDA graph: same codes
    as the eager chain over a rollout of several frames, and the decode half equals the oracle's embed_code."""
    from ccvs_b200.reencode import ReencodeStep
    gen = torch.Generator().manual_seed(9)
    B, h, w, D, K, cf = 16, 8, 8, 256, 1024, 512
    cb = torch.randn(K, D, generator=gen)
    vq = VectorQuantizer(K, D, 0.25).to(DEV).eval()
    tail = EncoderTail(cf, D).to(DEV)
    trunk = torch.nn.Sequential(torch.nn.Conv2d(D, cf, 1), torch.nn.Tanh()).to(DEV)       # stand-in for decoder + encoder trunk
    with torch.no_grad():
        vq.embedding.weight.copy_(cb.to(DEV))
    code0 = torch.randint(0, K, (B, h * w), generator=gen).to(DEV)
    stepper = ReencodeStep(vq, tail, trunk, code0, (h, w))
    code = code0.clone()
    for _ in range(4):
        want = stepper.eager(code).clone()
        got = stepper.step(code).clone()
        assert torch.equal(got, want)
        dec_ref = vq_oracle.embed_code(code.view(B, h, w).cpu(), cb).permute(0, 3, 1, 2)
        assert torch.equal(stepper.decoded.cpu(), dec_ref)
        assert int(got.min()) >= 0 and int(got.max()) < K
        code = got
    stepper.feed_back()
    assert torch.equal(stepper.code, stepper.new_code)
