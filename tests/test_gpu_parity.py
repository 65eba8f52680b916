"""GPU parity: the CUDA path (through the C ABI) against the golden fixtures and the CPU oracle.

Tolerances (BASELINE.json north_star): indices bit-exact except documented near-ties (oracle FP32
distance gap < 1e-6 relative); gathered embeddings bit-exact; z_q bit-exact given equal indices;
loss / perplexity within 1e-5 relative; dz rtol 1e-5 / atol 1e-7; dE rtol 1e-4 / atol 1e-7
(gradient tolerances are ours — north_star does not state them)."""
import pytest
import torch

import vq_oracle
from ccvs_b200 import VectorQuantizer, ops
from ccvs_b200.quantize import EMAVectorQuantizer

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _modes(g):
    modes = ["exact"]
    if ops.tensor_path_supported(g.n_e, g.e_dim):
        modes.append("tensor")
    return modes


def _build(g, mode):
    vq = VectorQuantizer(g.n_e, g.e_dim_total, g.beta, mult=g.mult, normalize=g.normalize, search_mode=mode).to(DEV)
    with torch.no_grad():
        vq.embedding.weight.copy_(g.codebook.to(DEV))
    return vq


def _check_indices(idx_ours, g):
    ours = idx_ours.view(-1).cpu()
    ref = g.indices.view(-1)
    diff = ours != ref
    if bool(diff.any()):
        rows = vq_oracle.to_channel_last(g.z).reshape(-1, g.e_dim)
        par = vq_oracle.classify_indices(ours, rows, g.codebook)
        assert par.mismatch == 0, f"{par.mismatch} rows differ outside documented near-ties ({par})"
    return not bool(diff.any())


def test_golden_forward_backward(golden):
    g = golden
    for mode in _modes(g):
        vq = _build(g, mode)
        z = g.z.to(DEV).requires_grad_(True)
        z_q, loss, (perp, one_hot, idx) = vq(z)
        assert z_q.shape == z.shape and idx.shape == (g.indices.shape[0], 1) and idx.dtype == torch.int64
        assert tuple(one_hot.shape) == (g.indices.shape[0], g.n_e)
        ((z_q * g.g_zq.to(DEV)).sum() + loss * g.g_loss).backward()
        same = _check_indices(idx, g)
        if same:
            if g.normalize:
                torch.testing.assert_close(z_q.detach().cpu(), g.z_q, rtol=1e-6, atol=1e-7)
            else:
                assert torch.equal(z_q.detach().cpu(), g.z_q), f"{g.name}/{mode}: z_q not bitwise fl(z+fl(e-z))"
            torch.testing.assert_close(loss.detach().cpu(), g.loss, rtol=1e-5, atol=0)
            torch.testing.assert_close(perp.cpu(), g.perplexity, rtol=1e-5, atol=0)
            assert torch.equal(one_hot.dense().sum(0).cpu(), g.one_hot_sum)
            torch.testing.assert_close(z.grad.cpu(), g.dz, rtol=1e-5, atol=1e-7)
            torch.testing.assert_close(vq.embedding.weight.grad.cpu(), g.dE, rtol=1e-4, atol=1e-7)
        else:  # only near-ties differ (fresh-init fixture): loss still within tolerance
            torch.testing.assert_close(loss.detach().cpu(), g.loss, rtol=1e-5, atol=0)


def test_golden_embed_code(golden):
    g = golden
    if g.code is None:
        return
    vq = _build(g, "exact")
    out = vq.embed_code(g.code.to(DEV))
    vq.check_codes()
    assert out.shape == g.embedded.shape
    assert torch.equal(out.cpu(), g.embedded)        # pure gather: bit-exact


def test_no_grad_and_eval_paths(golden):
    g = golden
    vq = _build(g, "auto").eval()
    with torch.no_grad():
        z_q, loss, (perp, _, idx) = vq(g.z.to(DEV))
    assert not z_q.requires_grad and not loss.requires_grad
    _check_indices(idx, g)
    idx2 = vq.encode_indices(g.z.to(DEV))
    assert torch.equal(idx2, idx.view(-1))


# ------------------------------------------------------------------------------------------------
# C-ABI level checks on seeded inputs (oracle finishes in well under a second at these sizes)
# ------------------------------------------------------------------------------------------------
CASES = [
    # shape,                K,    D,   dist
    ((16, 256, 16, 16), 1024, 256, "T"),     # BASELINE config 1
    ((16, 256, 16, 16), 1024, 256, "I"),     # fresh-init stress: near-ties everywhere
    ((4, 16, 512, 8, 8), 1024, 512, "T"),    # reference BAIR script shape
    ((3, 64, 5, 7), 200, 64, "T"),           # ragged N and K
    ((2, 128, 8, 8), 16384, 128, "T"),       # large codebook, few rows
]


@pytest.mark.parametrize("shape,K,D,dist", CASES)
@pytest.mark.parametrize("mode", ["exact", "tensor"])
def test_search_vs_oracle(shape, K, D, dist, mode):
    z, cb = vq_oracle.synth(shape, K, D, dist, seed=1234)
    lay = ops.layout_of(shape, D, 1)
    pcb = ops.prepare_codebook(cb.to(DEV))
    idx = ops.search(z.to(DEV), lay, pcb, mode=mode)
    rows = vq_oracle.to_channel_last(z).reshape(-1, D)
    par = vq_oracle.classify_indices(idx, rows, cb)
    assert par.mismatch == 0, par
    if dist == "T":
        assert par.agreement >= 0.9999, par
    torch.testing.assert_close(pcb.e_sq.cpu(), (cb ** 2).sum(1), rtol=1e-6, atol=0)


@pytest.mark.parametrize("cta_group", [1, 2])
@pytest.mark.parametrize("shape,K,D", [((8, 256, 8, 8), 1024, 256), ((5, 128, 7, 9), 300, 128), ((200, 512), 200, 512),
                                       ((3, 64, 16, 16), 96, 64)])
def test_screen_scores_and_candidates(shape, K, D, cta_group):
    """The tcgen05 GEMM itself: scores equal <bf16(z), bf16(e)> - 0.5|e|^2 (FP32 accumulate; the bias
    enters through the K-extension of the codebook shadow) and every code inside the margin is reported."""
    z, cb = vq_oracle.synth(shape, K, D, "T", seed=7)
    z = torch.randn_like(z)        # unit-variance latents against an N(0,1) codebook: real competition
    lay = ops.layout_of(shape, D, 1)
    pcb = ops.prepare_codebook(cb.to(DEV))
    N = lay.rows
    rows = vq_oracle.to_channel_last(z).reshape(-1, D).to(DEV)
    # codebook shadow: rounded codes + a bias split that reproduces -0.5|e|^2 to FP32 accuracy
    assert pcb.e_bf16.shape == (ops.codebook_rows(K), D + 16)
    assert torch.equal(pcb.e_bf16[:K, :D].float(), cb.to(DEV).to(torch.bfloat16).float())
    bias = pcb.e_bf16[:K, D:].float().sum(1)
    torch.testing.assert_close(bias, -0.5 * (cb.to(DEV) ** 2).sum(1), rtol=1e-6, atol=0)
    assert bool((pcb.e_bf16[K:, D] < -1e38).all()) and bool((pcb.e_bf16[K:, :D] == 0).all())

    tau = 3.0
    sd = ops.screen_debug(z.to(DEV), lay, pcb, n_cand=8, margin_tau=tau, cta_group=cta_group, dump_scores=True)
    # margin = tau * 2 (||dz|| max||e|| (1 + 2^-8) + ||z|| max||de||) + 2^-13 ||z|| max||e||, d. = the BF16 rounding error
    cbd = cb.to(DEV)
    emax = cbd.norm(dim=1).max()
    demax = (cbd - cbd.to(torch.bfloat16).float()).norm(dim=1).max()
    torch.testing.assert_close(pcb.e_max, torch.stack([emax, demax]), rtol=1e-5, atol=0)
    dz = (rows - rows.to(torch.bfloat16).float()).norm(dim=1)
    want = tau * 2 * (dz * emax * (1 + 2.0 ** -8) + rows.norm(dim=1) * demax) + 2.0 ** -13 * rows.norm(dim=1) * emax
    torch.testing.assert_close(sd.margin, want, rtol=2e-5, atol=0)
    assert float((sd.margin / (2.0 ** -8 * rows.norm(dim=1) * emax)).median()) < 2.0 * tau     # dense data: ~1.7 x 2^-8 ||z|| max||e|| per tau
    s = rows.to(torch.bfloat16).float() @ pcb.e_bf16[:K, :D].float().t() + bias
    # FP32 accumulation order differs between the tensor core and torch.matmul: |s| ~ 1e2 -> 2e-2
    torch.testing.assert_close(sd.scores[:, :K], s, rtol=1e-4, atol=2e-2)
    pad = sd.scores[:, K:]                                     # padding codes: bias -3e38, or never swept (NaN fill)
    assert bool(((pad < -1e38) | pad.isnan()).all())
    top = s.max(dim=1)
    ci, sc = sd.cand_idx, sd.cand_score
    torch.testing.assert_close(sc[:, 0], top.values, rtol=1e-4, atol=2e-2)
    picked = s.gather(1, ci[:, :1].long()).squeeze(1)
    assert bool((top.values - picked <= 2e-2).all())          # slot 0 is an argmax up to accumulation noise
    assert torch.equal(sd.idx, ci[:, 0].long())
    live = ci >= 0
    rep = s.gather(1, ci.clamp_min(0).long())
    assert bool(((rep - sc).abs()[live] <= 2e-2).all())       # reported scores belong to the reported codes
    # completeness: codes clearly inside the margin must be listed (unless the row overflowed)
    inside = s >= (top.values - sd.margin + 5e-2).unsqueeze(1)
    ok_rows = sd.flags == 0
    assert float(ok_rows.float().mean()) > 0.5
    listed = torch.zeros(N, K + 1, dtype=torch.bool, device=DEV)
    listed.scatter_(1, torch.where(live, ci, torch.full_like(ci, K)).long(), True)
    listed = listed[:, :K]
    assert bool((listed[ok_rows] | ~inside[ok_rows]).all())
    # nothing clearly outside the margin is listed
    outside = s < (top.values - sd.margin - 5e-2).unsqueeze(1)
    assert not bool((listed & outside).any())
    if K >= 300:
        assert float((live.sum(1) > 1).float().mean()) > 0.05     # the margin really admits rivals here
    assert bool((sc[:, :-1] >= sc[:, 1:]).all())              # sorted by score descending
    srt = torch.where(live, ci, torch.arange(-1, -1 - ci.shape[1], -1, device=DEV).expand_as(ci)).sort(dim=1).values
    assert bool((srt[:, 1:] != srt[:, :-1]).all())            # no duplicates among live candidates
    # the queue holds exactly the rows with more than one candidate or a flag
    nq = int(sd.queue.count)
    expect = ((live.sum(1) > 1) | (sd.flags != 0)).nonzero().squeeze(1)
    assert torch.equal(sd.queue.rows[:nq].long().sort().values, expect)
    order = sd.queue.rows[:nq].long()
    assert torch.equal(sd.queue.cand[:nq], ci[order]) and torch.equal(sd.queue.flags[:nq], sd.flags[order])


def test_overflow_rows_fall_back_to_exact():
    """More identical codes than candidate slots: the screen flags the row, the exact kernel resolves
    it to the lowest index (first-occurrence argmin, quantize.py:50)."""
    K, D = 512, 64
    torch.manual_seed(3)
    cb = torch.randn(K, D)
    cb[100:140] = cb[100]                         # 40 identical rows
    z = cb[torch.tensor([100, 7, 120, 139])].repeat(64, 1) + 0.0
    z = z.view(256, D)
    lay = ops.rows_layout(256, D)
    pcb = ops.prepare_codebook(cb.to(DEV))
    idx = ops.search(z.to(DEV), lay, pcb, mode="tensor", n_cand=4).cpu()
    expect = torch.tensor([100, 7, 100, 100]).repeat(64)
    assert torch.equal(idx, expect)
    sd = ops.screen_debug(z.to(DEV), lay, pcb, n_cand=4, margin_tau=1.0)
    flag = (sd.flags != 0).cpu().view(64, 4)
    assert bool(flag[:, [0, 2, 3]].all()) and not bool(flag[:, 1].any())


def test_assign_gather_backward_stats_finalize():
    shape, K, D = (6, 128, 5, 9), 300, 128
    z, cb = vq_oracle.synth(shape, K, D, "T", seed=11)
    zc, cbc = z.to(DEV), cb.to(DEV)
    lay = ops.layout_of(shape, D, 1)
    rows = vq_oracle.to_channel_last(z).reshape(-1, D)
    idx = vq_oracle.nearest(rows, cb)
    idc = idx.to(DEV)
    N, M = rows.shape[0], z.numel()

    zq, sq, counts = ops.assign(zc, lay, cbc, idc)
    e = cb[idx]
    zq_ref = vq_oracle.to_channel_first((rows + (e - rows)).view(vq_oracle.to_channel_last(z).shape))
    assert torch.equal(zq.cpu(), zq_ref)
    sq_ref = float(((e - rows).double() ** 2).sum())
    assert abs(float(sq) - sq_ref) <= 1e-5 * sq_ref
    assert torch.equal(counts.cpu(), torch.bincount(idx, minlength=K).to(torch.int32))

    out, err = ops.gather(idc, cbc)
    assert int(err) == 0 and torch.equal(out.cpu(), e)
    out_cm, _ = ops.gather(idc, cbc, lay)
    assert torch.equal(out_cm.view(shape).cpu(), vq_oracle.to_channel_first(e.view(vq_oracle.to_channel_last(z).shape)))
    bad = idc.clone()
    bad[5] = K
    _, err = ops.gather(bad, cbc)
    assert int(err) == 1                       # nn.Embedding would raise; we flag, never read out of bounds

    torch.manual_seed(5)
    g_zq = torch.randn(shape)
    g_loss = torch.tensor(0.37)
    dz = ops.backward_dz(zc, lay, cbc, idc, g_zq.to(DEV), g_loss.to(DEV))
    dz_rows = 2 * 0.37 * (rows - e) / M
    dz_ref = vq_oracle.to_channel_first(dz_rows.view(vq_oracle.to_channel_last(z).shape)) + g_zq
    torch.testing.assert_close(dz.cpu(), dz_ref, rtol=1e-5, atol=1e-7)
    dz0 = ops.backward_dz(zc, lay, cbc, idc, None, g_loss.to(DEV))
    dz0_ref = vq_oracle.to_channel_first(dz_rows.view(vq_oracle.to_channel_last(z).shape))
    torch.testing.assert_close(dz0.cpu(), dz0_ref, rtol=1e-5, atol=1e-7)

    resid, cnt = ops.code_stats(zc, lay, cbc, K, idc, sub=1.0)
    resid_ref = torch.zeros(K, D, dtype=torch.float64).index_add_(0, idx, (rows - e).double()).float()
    torch.testing.assert_close(resid.cpu(), resid_ref, rtol=1e-4, atol=1e-5)
    assert torch.equal(cnt.cpu(), counts.cpu())
    sums, _ = ops.code_stats(zc, lay, None, K, idc, sub=0.0, want_counts=False)
    sums_ref = torch.zeros(K, D, dtype=torch.float64).index_add_(0, idx, rows.double()).float()
    torch.testing.assert_close(sums.cpu(), sums_ref, rtol=1e-4, atol=1e-5)

    dE, loss, perp = ops.finalize(K, D, M, N, 0.25, resid=resid, counts=counts, sq_err=sq, g_loss=g_loss.to(DEV),
                                  want_dE=True, want_loss=True, want_perplexity=True)
    torch.testing.assert_close(dE.cpu(), -(2 * 0.25 * 0.37 / M) * resid_ref, rtol=1e-4, atol=1e-8)
    res = vq_oracle.forward(z, cb, 0.25)
    torch.testing.assert_close(loss.cpu(), res.loss, rtol=1e-5, atol=0)
    torch.testing.assert_close(perp.cpu(), res.perplexity, rtol=1e-5, atol=0)


# 128-bit fast paths (stream_fast.cu): partial last tile, two channel slabs, mult > 1, row-major rows,
# S not a multiple of the 32-position tile; each against plain torch math on the CPU
FAST_CASES = [((3, 256, 6, 6), 300, 256, 1), ((2, 512, 8, 8), 128, 512, 1), ((5, 256, 4, 4), 64, 128, 2),
              ((70, 256), 100, 256, 1), ((3, 64, 2, 10), 50, 64, 1), ((2, 3, 128, 8, 8), 1024, 128, 1),
              # enough 32-position tiles per SM for the persistent cp.async pipeline (incl. two channel slabs, C < 256)
              ((40, 256, 16, 16), 300, 256, 1), ((20, 512, 16, 16), 128, 512, 1), ((3, 80, 128, 8, 8), 64, 128, 1),
              ((160, 64, 8, 8), 40, 64, 1)]


@pytest.mark.parametrize("shape,K,D,mult", FAST_CASES)
def test_fast_stream_kernels(shape, K, D, mult):
    z, cb = vq_oracle.synth(shape, K, D, "T", seed=17)
    lay = ops.layout_of(shape, D, mult)
    assert ops.fast_stream_layout(lay)
    zl = vq_oracle.to_channel_last(z)
    rows = zl.reshape(-1, D)
    idx = vq_oracle.nearest(rows, cb)
    zc, cbc, idc = z.to(DEV), cb.to(DEV), idx.to(DEV)
    e = cb[idx]
    N, M = rows.shape[0], z.numel()
    back = (lambda r: vq_oracle.to_channel_first(r.view(zl.shape))) if len(shape) >= 4 else (lambda r: r.view(shape))

    zq, sq, counts = ops.assign(zc, lay, cbc, idc)
    assert torch.equal(zq.cpu(), back(rows + (e - rows)))
    sq_ref = float(((e - rows).double() ** 2).sum())
    assert abs(float(sq) - sq_ref) <= 1e-5 * sq_ref
    assert torch.equal(counts.cpu(), torch.bincount(idx, minlength=K).to(torch.int32))

    out, err = ops.gather(idc, cbc)
    assert int(err) == 0 and torch.equal(out.cpu(), e)
    if lay.S > 1:
        out_cm, _ = ops.gather(idc, cbc, lay)
        assert torch.equal(out_cm.view(shape).cpu(), back(e))
    bad = idc.clone()
    bad[N // 2] = -3
    _, err = ops.gather(bad, cbc, lay if lay.S > 1 else None)
    assert int(err) == 1

    torch.manual_seed(5)
    g_zq, g_loss = torch.randn(shape), torch.tensor(0.37)
    dz_rows = 2 * 0.37 * (rows - e) / M
    resid_ref = torch.zeros(K, D, dtype=torch.float64).index_add_(0, idx, (rows - e).double()).float()
    dz, dE = ops.quantize_backward(zc, lay, cbc, idc, g_zq.to(DEV), g_loss.to(DEV), 0.25)
    torch.testing.assert_close(dz.cpu(), back(dz_rows) + g_zq, rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(dE.cpu(), -(2 * 0.25 * 0.37 / M) * resid_ref, rtol=1e-4, atol=1e-8)
    dz0, none = ops.quantize_backward(zc, lay, cbc, idc, None, g_loss.to(DEV), 0.25, want_dE=False)
    assert none is None
    torch.testing.assert_close(dz0.cpu(), back(dz_rows), rtol=1e-5, atol=1e-7)
    none, dE1 = ops.quantize_backward(zc, lay, cbc, idc, None, g_loss.to(DEV), 0.25, want_dz=False)
    assert none is None
    torch.testing.assert_close(dE1.cpu(), dE.cpu(), rtol=1e-4, atol=1e-8)
    torch.testing.assert_close(ops.backward_dz(zc, lay, cbc, idc, g_zq.to(DEV), g_loss.to(DEV)).cpu(), dz.cpu(), rtol=0, atol=0)

    resid, cnt = ops.code_stats(zc, lay, cbc, K, idc, sub=1.0)
    torch.testing.assert_close(resid.cpu(), resid_ref, rtol=1e-4, atol=1e-5)
    assert torch.equal(cnt.cpu(), counts.cpu())
    sums, _ = ops.code_stats(zc, lay, None, K, idc, sub=0.0, want_counts=False)
    sums_ref = torch.zeros(K, D, dtype=torch.float64).index_add_(0, idx, rows.double()).float()
    torch.testing.assert_close(sums.cpu(), sums_ref, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("shape,K,D,mult,mode", [((16, 256, 16, 16), 1024, 256, 1, "auto"), ((3, 64, 5, 7), 200, 64, 1, "exact"),
                                                  ((5, 256, 4, 4), 64, 128, 2, "tensor"), ((4, 16, 2), 128, 1, 1, "auto"),
                                                  ((300, 64), 96, 64, 1, "auto")])
def test_composite_forward_equals_stepwise(shape, K, D, mult, mode):
    """ccvsq_quantize_forward (one call, folded finalize, in-call codebook side data) against the
    step-by-step entry points, and with a cached (frozen) codebook."""
    z, cb = vq_oracle.synth(shape, K, D, "T" if D > 1 else "I", seed=23)
    if D == 1:
        cb = torch.rand(K, 1)
    zc, cbc = z.to(DEV), cb.to(DEV)
    lay = ops.layout_of(shape, D, mult)
    pcb = ops.prepare_codebook(cbc)
    idx = ops.search(zc, lay, pcb, mode=mode)
    zq, sq, counts = ops.assign(zc, lay, cbc, idx)
    _, loss, perp = ops.finalize(K, D, float(z.numel()), float(lay.rows), 0.25, counts=counts, sq_err=sq, want_loss=True,
                                 want_perplexity=True)
    for cached in (None, pcb):
        out = ops.quantize_forward(zc, lay, cbc, 0.25, mode, cb=cached)
        assert torch.equal(out.idx, idx) and torch.equal(out.zq, zq) and torch.equal(out.counts, counts)
        torch.testing.assert_close(out.loss, loss, rtol=1e-6, atol=0)
        torch.testing.assert_close(out.perplexity, perp, rtol=1e-6, atol=0)
        only = ops.quantize_forward(zc, lay, cbc, 0.25, mode, cb=cached, indices_only=True)
        assert torch.equal(only.idx, idx) and only.zq is None


def test_mult_and_flat_layouts_through_abi():
    # mult > 1 on a channel-major tensor, and the ndim < 4 contiguous-row case incl. e_dim = 1
    for shape, K, D, mult in [((3, 32, 4, 5), 64, 8, 4), ((50, 24), 16, 24, 1), ((4, 16, 2), 128, 1, 1)]:
        z, cb = vq_oracle.synth(shape, K, D, "T" if D > 1 else "I", seed=21)
        if D == 1:
            cb = torch.rand(K, 1)
        lay = ops.layout_of(shape, D, mult)
        pcb = ops.prepare_codebook(cb.to(DEV))
        idx = ops.search(z.to(DEV), lay, pcb, mode="exact")
        res = vq_oracle.forward(z, cb, 0.25, mult)
        par = vq_oracle.classify_indices(idx, vq_oracle.to_channel_last(z).reshape(-1, D), cb)
        assert par.mismatch == 0, (shape, par)
        if par.exact == par.n:
            zq, sq, _ = ops.assign(z.to(DEV), lay, cb.to(DEV), idx)
            assert torch.equal(zq.cpu(), res.z_q.detach())


def test_single_code_and_tiny_inputs():
    cb = torch.randn(1, 8)
    z = torch.randn(1, 8, 1, 1)
    lay = ops.layout_of(z.shape, 8, 1)
    idx = ops.search(z.to(DEV), lay, ops.prepare_codebook(cb.to(DEV)), mode="auto")
    assert idx.tolist() == [0]


def test_ema_update_matches_textbook():
    K, D = 64, 32
    z, cb = vq_oracle.synth((4, 32, 6, 6), K, D, "T", seed=31)
    rows = vq_oracle.to_channel_last(z).reshape(-1, D)
    m = EMAVectorQuantizer(K, D, 0.25, decay=0.9, eps=1e-5, sync=False).to(DEV).train()
    with torch.no_grad():
        m.embedding.weight.copy_(cb.to(DEV))
        m.ema_sum.copy_(cb.to(DEV))
        m.ema_count.fill_(1.0)
    idx_ref = vq_oracle.nearest(rows, cb)
    new_cb, n, s = vq_oracle.ema_update(cb, torch.ones(K), cb.clone(), rows, idx_ref, 0.9, 1e-5)
    zc = z.to(DEV).requires_grad_(True)
    z_q, loss, (_, _, idx) = m(zc)
    assert torch.equal(idx.view(-1).cpu(), idx_ref)
    assert torch.equal(m.embedding.weight.detach().cpu(), cb)     # deferred: the update waits for the backward ...
    m.sync_codebook()                                             # ... or for an explicit flush
    torch.testing.assert_close(m.ema_count.cpu(), n, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(m.ema_sum.cpu(), s, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(m.embedding.weight.cpu(), new_cb, rtol=1e-4, atol=1e-5)
    loss.backward()       # backward uses the pre-update codebook
    e = cb[idx_ref]
    dz_ref = vq_oracle.to_channel_first((2 * (rows - e) / z.numel()).view(vq_oracle.to_channel_last(z).shape))
    torch.testing.assert_close(zc.grad.cpu(), dz_ref, rtol=1e-5, atol=1e-7)


def test_frozen_codebook_cache_and_polyak_update():
    """Polyak averaging writes param.data in place (quantized_video_model.py:962-964): the default
    (unfrozen) module must see the new codebook."""
    z, cb = vq_oracle.synth((2, 64, 8, 8), 256, 64, "T", seed=41)
    vq = VectorQuantizer(256, 64, 0.25).to(DEV)
    with torch.no_grad():
        vq.embedding.weight.copy_(cb.to(DEV))
        i0 = vq(z.to(DEV))[2][2].clone()
        perm = torch.randperm(256)
        vq.embedding.weight.data.copy_(cb[perm].to(DEV))        # .data write: no version bump
        i1 = vq(z.to(DEV))[2][2]
    inv = torch.empty(256, dtype=torch.int64)
    inv[perm] = torch.arange(256)
    assert torch.equal(i1.view(-1).cpu(), inv[i0.view(-1).cpu()])   # permuting rows permutes indices


def test_host_pipeline_equals_whole_batch():
    """ccvs_b200.pipeline: frame-chunked, copy/compute-overlapped quantization of a host batch returns what one
    whole-batch forward returns (indices bit-exact; loss / perplexity to rounding)."""
    from ccvs_b200.pipeline import HostQuantizePipeline
    shape, K, D = (8, 4, 64, 8, 8), 256, 64
    z, cb = vq_oracle.synth(shape, K, D, "T", seed=51)
    vq = VectorQuantizer(K, D, 0.25).to(DEV).eval()
    with torch.no_grad():
        vq.embedding.weight.copy_(cb.to(DEV))
        _, loss, (perp, _, idx) = vq(z.to(DEV))
    pipe = HostQuantizePipeline(vq, shape, n_chunks=4, decode=True)
    zh = z.pin_memory()
    for _ in range(2):                      # second run exercises the slot-reuse events
        idx_h, sc_h = pipe.run(zh)
        pipe.synchronize()
        torch.cuda.synchronize()
    assert torch.equal(idx_h, idx.view(-1).cpu())
    torch.testing.assert_close(sc_h[0], loss.cpu(), rtol=1e-5, atol=0)
    torch.testing.assert_close(sc_h[1], perp.cpu(), rtol=1e-5, atol=0)
    dec = torch.cat([d.reshape(-1, D) for d in pipe.decoded])
    assert torch.equal(dec.cpu(), cb[idx.view(-1).cpu()])


def test_cuda_graph_replay_matches_eager():
    """VectorQuantizer.capture: one graph launch per call (small-batch / per-frame re-encode regime)."""
    shape, K, D = (16, 256, 16, 16), 1024, 256            # BASELINE config 1
    z, cb = vq_oracle.synth(shape, K, D, "T", seed=61)
    z2, _ = vq_oracle.synth(shape, K, D, "T", seed=62)
    vq = VectorQuantizer(K, D, 0.25).to(DEV).eval()
    with torch.no_grad():
        vq.embedding.weight.copy_(cb.to(DEV))
    zs = z.to(DEV).clone()
    g = vq.capture(zs, decode=True)
    for zin in (z, z2, z):
        zq_g, loss_g, (perp_g, _, idx_g) = g(zin.to(DEV))
        with torch.no_grad():
            zq, loss, (perp, _, idx) = vq(zin.to(DEV))
            dec = vq.embed_code(idx.view(shape[0], -1))
        assert torch.equal(idx_g, idx) and torch.equal(zq_g, zq) and torch.equal(g.decoded, dec)
        torch.testing.assert_close(loss_g, loss, rtol=1e-6, atol=0)
        torch.testing.assert_close(perp_g, perp, rtol=1e-6, atol=0)
    # the codebook is re-read on every replay (Polyak averaging writes it in place)
    with torch.no_grad():
        vq.embedding.weight.data.copy_(cb.flip(0).to(DEV))
    _, _, (_, _, idx_f) = g(z.to(DEV))
    res = vq_oracle.forward(z, cb.flip(0), 0.25)
    par = vq_oracle.classify_indices(idx_f.view(-1).cpu(), vq_oracle.to_channel_last(z).reshape(-1, D), cb.flip(0))
    assert par.mismatch == 0 and par.agreement >= 0.9999, par


def test_polyak_and_prior_handoff():
    """SURVEY 8a row a15 (Polyak average of the codebook) and 8f N2 (tok_emb gather + positional add)."""
    torch.manual_seed(71)
    K, D = 1024, 256
    a, b = VectorQuantizer(K, D, 0.25).to(DEV), VectorQuantizer(K, D, 0.25).to(DEV)
    with torch.no_grad():
        a.embedding.weight.copy_(torch.randn(K, D))
        b.embedding.weight.copy_(torch.randn(K, D))
    ref = a.embedding.weight.detach().cpu().clone()
    for _ in range(3):
        ref = vq_oracle.polyak(ref, b.embedding.weight.detach().cpu(), 0.999)
        v0 = a.embedding.weight._version
        a.accumulate_from(b, 0.999)
        assert a.embedding.weight._version == v0              # .data write, like the reference
    torch.testing.assert_close(a.embedding.weight.detach().cpu(), ref, rtol=1e-6, atol=1e-7)

    n_embd, T, B = 1024, 320, 6                                # mingpt: n_embd 1024, a 5-frame 8x8 window
    tok = torch.randn(K, n_embd)
    pos = torch.randn(1, 1280, n_embd) * 0.02
    code = torch.randint(0, K, (B, T))
    out = a.embed_tokens(code.to(DEV), tok.to(DEV), pos.to(DEV))
    a.check_codes()
    assert torch.equal(out.cpu(), vq_oracle.embed_tokens(code, tok, pos))       # one rounding: bit-exact
    bad = code.clone()
    bad[2, 5] = K + 3
    a.embed_tokens(bad.to(DEV), tok.to(DEV), pos.to(DEV))
    with pytest.raises(IndexError):
        a.check_codes()
    # a table whose width is not a multiple of 128 takes the flat 16-byte kernel
    tok2, pos2 = torch.randn(50, 72), torch.randn(1, 40, 72)
    code2 = torch.randint(0, 50, (3, 40))
    out2, _ = ops.gather_add(code2.to(DEV), tok2.to(DEV), pos2.to(DEV))
    assert torch.equal(out2.cpu(), vq_oracle.embed_tokens(code2, tok2, pos2))


def test_collapsed_codebook_every_row_falls_back():
    """A collapsed codebook (all codes within 1e-6 of one vector: what k-means-free VQ training produces right
    after a bad init) puts every code of every row inside the screening margin, so ALL rows go to the exact FP32
    fallback — more than the 65 536 rows the fallback list used to hold.  The result must still be the FP32 argmin."""
    torch.manual_seed(5)
    K, D, N = 256, 64, 1 << 17
    base = torch.randn(1, D)
    cb = (base + 1e-6 * torch.randn(K, D)).contiguous()
    z = (base + 0.5 * torch.randn(N, D)).contiguous()
    lay = ops.layout_of(z.shape, D, 1)
    pcb = ops.prepare_codebook(cb.to(DEV))
    zt = z.to(DEV)
    idx_t = ops.search(zt, lay, pcb, mode="tensor")
    idx_e = ops.search(zt, lay, pcb, mode="exact")
    assert torch.equal(idx_t, idx_e)
    out = ops.quantize_forward(zt, lay, cb.to(DEV), 0.25, mode="tensor")
    assert torch.equal(out.idx.view(-1), idx_e)
    # against the CPU oracle only the tie band can be checked here: all K distances of a row lie within ~1e-6
    # relative of each other, and d = (||z||^2 + ||e||^2) - 2 z.e cancels from ~150 down to ~16, so one ulp of the
    # sums is already 1e-6 of d: any two FP32 evaluations (MKL vs CUDA cores) pick different winners inside it
    par = vq_oracle.classify_indices(idx_t[: 1 << 14], z[: 1 << 14], cb, rel_tol=1e-5)
    assert par.mismatch == 0, par


def test_decoder_head_folded_into_the_decode_gather():
    """SURVEY 8f N3 (decoder side): embed_code + NHWC->NCHW + ConvLayer(z_size, block_in, 1) as ONE gather of a
    K-row table.  Checked against the oracle's restatement of the reference chain (embedding -> transposes -> 1x1
    equalised conv -> bias + LeakyReLU(0.2) * sqrt(2)) evaluated in FP32 on the CPU."""
    torch.manual_seed(11)
    K, C, C_out, G, h, w = 1024, 256, 512, 12, 16, 16
    cb = torch.randn(K, C)
    weight = torch.randn(C_out, C, 1, 1)
    bias = 0.1 * torch.randn(C_out)
    code = torch.randint(0, K, (G, h * w))
    ref = vq_oracle.decoder_head(code, cb, weight, bias, (h, w))

    vq = VectorQuantizer(K, C, 0.25).to(DEV).eval()
    with torch.no_grad():
        vq.embedding.weight.copy_(cb.to(DEV))
    wd, bd = weight.to(DEV), bias.to(DEV)

    def head(x):   # the reference's ConvLayer(C, C_out, 1) in plain torch ops (gan.py:108-115, fused_act.py:105-114)
        out = torch.nn.functional.conv2d(x, wd * (1.0 / C ** 0.5))
        return torch.nn.functional.leaky_relu(out + bd.view(1, -1, 1, 1), 0.2) * 2 ** 0.5

    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False           # the table must be FP32 like the CPU oracle
    try:
        table = vq.fold_pointwise_head(head)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    assert table.shape == (K, C_out)
    out = vq.embed_code(code.to(DEV), channel_major_hw=(h, w), table=table)
    vq.check_codes()
    assert out.shape == (G, C_out, h, w)
    torch.testing.assert_close(out.cpu(), ref, rtol=1e-5, atol=1e-5)
    # the gather itself is exact: every position carries its code's table row bit for bit
    assert torch.equal(out.permute(0, 2, 3, 1).reshape(-1, C_out), table[code.view(-1).to(DEV)])
    # channel-last variant ([G, h*w, C_out])
    out_cl = vq.embed_code(code.to(DEV), table=table)
    assert torch.equal(out_cl.view(-1, C_out), table[code.view(-1).to(DEV)])


@pytest.mark.parametrize("shape", [(6, 256, 16, 16), (3, 64, 5, 8), (700, 128)])
def test_assign_with_skewed_code_usage(shape):
    """Codebook collapse: most latents on one or two codes (the usage counts combine equal codes per warp before the
    atomics, and the codebook-row gather must not depend on rows being distinct)."""
    torch.manual_seed(17)
    D = shape[1] if len(shape) > 2 else shape[-1]
    K = 37
    z = torch.randn(shape, device=DEV)
    cb = torch.randn(K, D, device=DEV)
    lay = ops.layout_of(shape, D, 1)
    N = lay.rows
    for idx in (torch.zeros(N, dtype=torch.int64, device=DEV),
                torch.where(torch.rand(N, device=DEV) < 0.9, 5, 31).to(torch.int64),
                torch.randint(0, K, (N,), device=DEV)):
        zq, sq, counts = ops.assign(z, lay, cb, idx)
        assert torch.equal(counts.to(torch.int64), torch.bincount(idx, minlength=K))
        rows = vq_oracle.to_channel_last(z.cpu()).reshape(-1, D) if len(shape) >= 4 else z.cpu().reshape(-1, D)
        e = cb.cpu()[idx.cpu()]
        ref = rows + (e - rows)
        ours = vq_oracle.to_channel_last(zq.cpu()).reshape(-1, D) if len(shape) >= 4 else zq.cpu().reshape(-1, D)
        assert torch.equal(ours, ref)
        ref_sq = float(((e - rows).double() ** 2).sum())
        assert abs(float(sq) - ref_sq) <= 1e-5 * ref_sq


@pytest.mark.parametrize("shape,K,D", [((8, 256, 16, 16), 1024, 256), ((3, 64, 5, 7), 200, 64), ((300, 128), 96, 128)])
def test_forward_with_fused_code_statistics(shape, K, D):
    """ccvsq_forward_args.resid: the per-code residual sums of the EMA update accumulated by the assign pass equal
    the stand-alone scatter-reduce (ccvsq_code_stats) and the oracle's sum_{idx=k} (z - E[k]); everything else
    the forward returns is unchanged."""
    z, cb = vq_oracle.synth(shape, K, D, "T", seed=99)
    zd, cbd = z.to(DEV), cb.to(DEV)
    lay = ops.layout_of(shape, D, 1)
    plain = ops.quantize_forward(zd, lay, cbd, 0.25)
    fused = ops.quantize_forward(zd, lay, cbd, 0.25, want_resid=True)
    assert torch.equal(plain.idx, fused.idx) and torch.equal(plain.zq, fused.zq)
    assert torch.equal(plain.counts, fused.counts)
    torch.testing.assert_close(fused.loss, plain.loss, rtol=1e-6, atol=0)
    alone, _ = ops.code_stats(zd, lay, cbd, K, plain.idx, sub=1.0, want_counts=False)
    torch.testing.assert_close(fused.resid, alone, rtol=1e-4, atol=1e-4)
    rows = (vq_oracle.to_channel_last(z) if len(shape) >= 4 else z).reshape(-1, D).double()
    idx = plain.idx.cpu()
    ref = torch.zeros(K, D, dtype=torch.float64).index_add_(0, idx, rows - cb.double()[idx])
    torch.testing.assert_close(fused.resid.cpu().double(), ref, rtol=1e-4, atol=1e-4)


def test_empty_batch_behaves_like_the_reference():
    """Zero latents (an empty shard): the reference returns empty z_q / indices and NaN loss / perplexity (means over
    zero elements), backward gives an empty dz and a zero dE; embed_code of no codes is an empty tensor."""
    vq = VectorQuantizer(16, 8, 0.25).to(DEV)
    z = torch.zeros(0, 8, 4, 4, device=DEV, requires_grad=True)
    z_q, loss, (perp, enc, idx) = vq(z)
    assert z_q.shape == z.shape and tuple(idx.shape) == (0, 1) and idx.dtype == torch.int64
    assert tuple(enc.shape) == (0, 16)
    assert torch.isnan(loss) and torch.isnan(perp)
    (z_q.sum() + loss).backward()
    assert z.grad.shape == z.shape
    assert vq.embedding.weight.grad.shape == (16, 8) and float(vq.embedding.weight.grad.abs().sum()) == 0.0
    assert vq.encode_indices(z.detach()).shape == (0,)
    assert vq.embed_code(torch.zeros(0, 4, 4, dtype=torch.int64, device=DEV)).shape == (0, 4, 4, 8)


def test_half_precision_latents_are_upcast():
    """Mixed-precision callers: FP16 / BF16 latents are upcast, the path runs in FP32 (indices equal those of the
    FP32 call on the upcast values), z_q comes back in the caller's dtype and gradients flow through the cast."""
    z, cb = vq_oracle.synth((4, 256, 16, 16), 1024, 256, "T", seed=77)
    vq = VectorQuantizer(1024, 256, 0.25).to(DEV)
    with torch.no_grad():
        vq.embedding.weight.copy_(cb.to(DEV))
    for dt in (torch.float16, torch.bfloat16):
        zh = z.to(DEV).to(dt).requires_grad_(True)
        z_q, loss, (_, _, idx) = vq(zh)
        assert z_q.dtype == dt and z_q.shape == zh.shape and loss.dtype == torch.float32
        with torch.no_grad():
            ref_q, ref_loss, (_, _, ref_idx) = vq(zh.detach().float())
        assert torch.equal(idx, ref_idx)
        assert torch.equal(z_q.detach(), ref_q.to(dt))
        torch.testing.assert_close(loss.detach(), ref_loss, rtol=1e-6, atol=0)
        (z_q.float().sum() + loss).backward()
        assert zh.grad is not None and zh.grad.dtype == dt


def test_smoke_without_programmatic_dependent_launch():
    """The kernels call griddepcontrol.wait / launch_dependents unconditionally; CCVSQ_NO_PDL=1 drops the launch
    attribute (ordinary stream serialisation).  Same results either way: run the smoke check in a child process."""
    import os
    import subprocess
    import sys

    from conftest import ROOT
    env = dict(os.environ, CCVSQ_NO_PDL="1")
    r = subprocess.run([sys.executable, "-c", "import __graft_entry__ as g; g.smoke()"], cwd=ROOT, env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "smoke ok" in r.stdout


def test_ema_deferred_update_keeps_the_codebook_for_a_late_backward():
    """Deferred mode (the EMA update waits for the module's backward): a second forward before the first backward
    flushes the pending update — the first backward must still see the codebook ITS forward used."""
    shape, K, D = (4, 64, 8, 8), 96, 64
    z1, cb = vq_oracle.synth(shape, K, D, "T", seed=31)
    z2, _ = vq_oracle.synth(shape, K, D, "T", seed=32)
    vq = EMAVectorQuantizer(K, D, 0.25, decay=0.5, sync=False, overlap=True).to(DEV).train()
    with torch.no_grad():
        vq.embedding.weight.copy_(cb.to(DEV))
        vq.ema_sum.copy_(cb.to(DEV))
        vq.ema_count.fill_(1.0)
    a = z1.to(DEV).requires_grad_(True)
    b = z2.to(DEV).requires_grad_(True)
    w0 = vq.embedding.weight.detach().clone()
    zq1, l1, (_, _, i1) = vq(a)
    assert torch.equal(vq.embedding.weight.detach(), w0)          # update deferred: still the codebook of forward 1
    zq2, l2, (_, _, i2) = vq(b)                                   # flushes update 1 (then defers update 2)
    w1 = vq.embedding.weight.detach().clone()
    assert not torch.equal(w1, w0)
    l2.backward()
    l1.backward()
    vq.sync_codebook()
    M = z1.numel()
    # dz = (2 / M) (z - E_used[idx]) with the codebook each forward saw
    torch.testing.assert_close(a.grad.cpu(), (2.0 / M) * (z1 - vq_oracle.to_channel_first(w0.cpu()[i1.view(-1).cpu()].view(4, 8, 8, D))),
                               rtol=1e-5, atol=1e-9)
    torch.testing.assert_close(b.grad.cpu(), (2.0 / M) * (z2 - vq_oracle.to_channel_first(w1.cpu()[i2.view(-1).cpu()].view(4, 8, 8, D))),
                               rtol=1e-5, atol=1e-9)
    assert not vq._stale_codebooks and vq._pending is None
