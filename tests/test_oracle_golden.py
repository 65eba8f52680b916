"""Pin the oracle (oracle/vq_oracle.py) to the reference: every golden fixture was produced by the
unmodified reference file (oracle/gen_golden.py); the restatement must reproduce it."""
import torch

import vq_oracle
from conftest import GOLDEN_NAMES


def test_fixtures_present():
    assert len(GOLDEN_NAMES) >= 12


def test_forward_matches_reference(golden):
    g = golden
    res = vq_oracle.forward(g.z, g.codebook, g.beta, g.mult, g.normalize)
    # indices: exact except rows the reference itself resolves inside its FP32 noise
    same = res.indices == g.indices
    if not bool(same.all()):
        bad = (~same.view(-1)) & (g.top2_rel_gap > 1e-6)
        assert not bool(bad.any()), f"{int(bad.sum())} index mismatches outside near-ties"
    else:
        assert torch.equal(res.z_q, g.z_q), "z_q must be bitwise fl(z + fl(e - z))"
    torch.testing.assert_close(res.loss, g.loss, rtol=1e-6, atol=0)
    torch.testing.assert_close(res.perplexity, g.perplexity, rtol=1e-5, atol=0)
    assert torch.equal(res.one_hot.sum(0), g.one_hot_sum) or not bool(same.all())


def test_lazy_path_equals_dense(golden):
    g = golden
    a = vq_oracle.forward(g.z, g.codebook, g.beta, g.mult, g.normalize, dense=True)
    b = vq_oracle.forward(g.z, g.codebook, g.beta, g.mult, g.normalize, dense=False)
    assert torch.equal(a.indices, b.indices)
    assert torch.equal(a.z_q, b.z_q)            # onehot @ E == E[idx] bitwise
    torch.testing.assert_close(a.loss, b.loss, rtol=1e-6, atol=0)
    torch.testing.assert_close(a.perplexity, b.perplexity, rtol=1e-5, atol=0)


def test_backward_matches_reference(golden):
    g = golden
    _, dz, dE = vq_oracle.forward_backward(g.z, g.codebook, g.beta, g.g_zq, g.g_loss, g.mult, g.normalize)
    torch.testing.assert_close(dz, g.dz, rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(dE, g.dE, rtol=1e-5, atol=1e-7)


def test_closed_form_gradients(golden):
    """dz = g_zq + 2 g (z - E[idx]) / M ; dE = (2 beta g / M)(n_k E_k - sum z) — the
    formulas the CUDA backward implements, checked against the reference's autograd output."""
    g = golden
    if g.normalize:
        return
    rows = vq_oracle.to_channel_last(g.z).reshape(-1, g.e_dim)
    idx = g.indices.view(-1)
    M = g.z.numel()
    e = g.codebook[idx]
    dz_rows = 2 * g.g_loss * (rows - e) / M
    dz = vq_oracle.to_channel_first(dz_rows.view(vq_oracle.to_channel_last(g.z).shape)) + g.g_zq
    torch.testing.assert_close(dz, g.dz, rtol=1e-5, atol=1e-7)
    K = g.n_e
    n = torch.bincount(idx, minlength=K).to(torch.float32)
    s = torch.zeros(K, g.e_dim).index_add_(0, idx, rows)
    dE = (2 * g.beta * g.g_loss / M) * (n.unsqueeze(1) * g.codebook - s)
    torch.testing.assert_close(dE, g.dE, rtol=1e-4, atol=1e-7)


def test_embed_code_matches_reference(golden):
    g = golden
    if g.code is None:
        return
    out = vq_oracle.embed_code(g.code, g.codebook, g.mult)
    assert out.shape == g.embedded.shape
    assert torch.equal(out, g.embedded)


def test_duplicate_rows_resolve_to_lowest_index():
    from conftest import Golden
    g = Golden("dup_rows_ties")
    idx = g.indices.view(-1)
    pick = torch.arange(idx.numel()) % g.n_e
    expect = torch.where((pick == 5) | (pick == 9), torch.tensor(2), pick)
    assert torch.equal(idx, expect)     # the reference's own behaviour (first occurrence)
    assert torch.equal(vq_oracle.nearest(vq_oracle.to_channel_last(g.z).view(-1, g.e_dim), g.codebook), expect)


def test_chunked_nearest_equals_unchunked():
    z, cb = vq_oracle.synth((4, 64, 8, 8), 512, 64, "T", seed=5)
    rows = vq_oracle.to_channel_last(z).view(-1, 64)
    assert torch.equal(vq_oracle.nearest(rows, cb), vq_oracle.nearest(rows, cb, chunk=37))


def test_classifier_counts():
    z, cb = vq_oracle.synth((2, 32, 8, 8), 128, 32, "T", seed=6)
    rows = vq_oracle.to_channel_last(z).view(-1, 32)
    ref = vq_oracle.nearest(rows, cb)
    par = vq_oracle.classify_indices(ref, rows, cb)
    assert par.exact == par.n and par.mismatch == 0 and par.agreement == 1.0
    wrong = ref.clone()
    wrong[:5] = (wrong[:5] + 1) % 128
    par = vq_oracle.classify_indices(wrong, rows, cb)
    assert par.exact == par.n - 5 and par.mismatch == 5


def test_ema_restatement_fixed_point():
    # a codebook already at the cluster means with full EMA history stays put
    torch.manual_seed(0)
    cb = torch.randn(8, 4)
    idx = torch.arange(64) % 8
    rows = cb[idx]
    new_cb, n, s = vq_oracle.ema_update(cb, torch.full((8,), 8.0), cb * 8.0, rows, idx, decay=0.9, eps=0.0)
    torch.testing.assert_close(new_cb, cb, rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(n, torch.full((8,), 8.0))


def test_oracle_polyak_and_embed_tokens_shapes():
    """The checker's restatements of the caller-side rows (quantized_video_model.py:951-964, mingpt.py:234-236)."""
    import torch
    import vq_oracle
    ema, live = torch.ones(4, 3), torch.zeros(4, 3)
    out = vq_oracle.polyak(ema, live, 0.999)
    assert torch.allclose(out, torch.full((4, 3), 0.999)) and torch.equal(ema, torch.ones(4, 3))
    tok, pos = torch.arange(12.).view(4, 3), torch.ones(1, 5, 3)
    e = vq_oracle.embed_tokens(torch.tensor([[0, 3], [1, 1]]), tok, pos)
    assert e.shape == (2, 2, 3) and torch.equal(e[0, 1], tok[3] + 1)


def test_oracle_decoder_head_is_a_function_of_the_code():
    """The property the folded decoder head relies on: the reference chain embed_code -> transposes -> 1x1 conv ->
    bias + LeakyReLU (vq_oracle.decoder_head) gives, at every position, the same vector as the chain applied to the
    single codebook row of that position's code."""
    torch.manual_seed(3)
    K, C, C_out, G, h, w = 64, 32, 48, 3, 4, 5
    cb, weight, bias = torch.randn(K, C), torch.randn(C_out, C, 1, 1), torch.randn(C_out)
    code = torch.randint(0, K, (G, h * w))
    full = vq_oracle.decoder_head(code, cb, weight, bias, (h, w))                      # [G, C_out, h, w]
    per_code = vq_oracle.decoder_head(torch.arange(K).view(K, 1), cb, weight, bias, (1, 1)).view(K, C_out)
    torch.testing.assert_close(full.permute(0, 2, 3, 1).reshape(-1, C_out), per_code[code.view(-1)], rtol=1e-6, atol=1e-6)


def test_oracle_empty_batch_matches_reference_behaviour():
    """What the unmodified reference does on zero latents (checked by running quantize.py on CPU): empty z_q and
    indices, NaN loss and perplexity.  The oracle restates it; the CUDA module is tested against the same facts."""
    z, cb = torch.zeros(0, 8, 4, 4), torch.randn(16, 8)
    r = vq_oracle.forward(z, cb, 0.25)
    assert r.z_q.shape == z.shape and tuple(r.indices.shape) == (0, 1)
    assert torch.isnan(r.loss) and torch.isnan(r.perplexity)


def test_oracle_encoder_tail_matches_reference_classes(tail_golden):
    """The restated encoder tail (conv2d with weight * scale + bias, LeakyReLU(0.1), optional L2 normalisation) against
    outputs of the reference's own EqualConv2d / ConvLayer source (oracle/gen_golden_tail.py)."""
    g = tail_golden
    out = vq_oracle.encoder_tail(g.x, g.weight, g.bias)
    torch.testing.assert_close(out, g.out, rtol=1e-6, atol=1e-6)
    out_n = vq_oracle.encoder_tail(g.x, g.weight, g.bias, normalize_out=True)
    torch.testing.assert_close(out_n, g.out_normalized, rtol=1e-6, atol=1e-6)
    assert float((out < 0).float().mean()) > 0.2          # the LeakyReLU branch is exercised
