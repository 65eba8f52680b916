"""GPU parity hardening: adversarial inputs for the BF16 screening margin, property tests (hypothesis) and
non-finite latents.  Everything goes through the C ABI (ccvs_b200.ops / VectorQuantizer) and is checked against
the CPU oracle (oracle/vq_oracle.py) on the same inputs.

Parity rules as in tests/test_gpu_parity.py: indices equal to the oracle's except documented near-ties (oracle
FP32 distance gap <= 1e-6 relative); z_q / embed_code bit-exact given equal indices; loss within 1e-5 relative."""
import numpy as np
import pytest
import torch
from hypothesis import HealthCheck, given, settings, strategies as st

import vq_oracle
from ccvs_b200 import VectorQuantizer, ops

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


# ------------------------------------------------------------------------------------------------
# adversarial margin: operands at BF16 rounding midpoints, energy concentrated in two channels
# ------------------------------------------------------------------------------------------------
def _bf16(x):
    return torch.as_tensor(x, dtype=torch.float32).to(torch.bfloat16).to(torch.float64)


def _midpoints(direction, rng, n):
    """FP32 values exactly halfway between two neighbouring BF16 numbers in [1.5, 2) whose round-to-nearest-even
    goes `direction` (+1 up, -1 down): the largest rounding error BF16 can make (2^-8 relative)."""
    out = []
    while len(out) < n:
        m = int(rng.integers(64, 127))               # 7-bit mantissa of the lower neighbour
        goes_up = (m & 1) == 1                       # ties to even: an odd lower neighbour rounds up
        if goes_up == (direction > 0):
            out.append(1.0 + (2 * m + 1) / 256.0)    # lower neighbour 1 + m/128, midpoint adds 2^-8
    return np.array(out)


def _round_up_midpoint_below(x):
    """Largest BF16 midpoint <= x that rounds UP (1 + m/256 with m = 3 mod 4, times the binade of x)."""
    e = int(np.floor(np.log2(x)))
    m = int(np.floor((x / 2.0 ** e - 1.0) * 256.0))
    while m % 4 != 3:
        m -= 1
    return None if m < 0 else (1.0 + m / 256.0) * 2.0 ** e


def adversarial_case(D=64, rows_per_pair=8, seed=5):
    """Rows and codes built so that the FP32 winner A of every row LOSES the BF16-rounded comparison against a
    rival B by more than the tau = 1 margin: row = (x1 at channel 2p, x2 at channel 2p+1), A_p = a on channel 2p,
    B_p = b on channel 2p+1.  x1 and a round DOWN (the score of A shrinks by ~2^-7 x1 a), x2 and b round UP (the
    score of B grows by ~2^-7 x2 b); x2 is solved so that in exact arithmetic A still wins by a gap of at least
    2e-4 of its distance (200 x the documented near-tie tolerance).
    Returns z [N, D], codebook [K, D], and per row (BF16 deficit of the FP32 winner) / (margin at tau = 1)."""
    rng = np.random.default_rng(seed)
    pairs = D // 2
    cb = np.zeros((2 * pairs, D), dtype=np.float64)
    picks = []
    for p in range(pairs):
        rows = []
        while len(rows) < rows_per_pair:                      # resample the pair until it yields enough rows
            a, b = float(_midpoints(-1, rng, 1)[0]), float(_midpoints(+1, rng, 1)[0])
            rows = []
            for _ in range(400):
                x1 = float(_midpoints(-1, rng, 1)[0])
                sA = x1 * a - 0.5 * a * a
                x2 = _round_up_midpoint_below((sA + 0.5 * b * b) / b)     # s_B just below s_A
                if x2 is None:
                    continue
                gap = sA - (x2 * b - 0.5 * b * b)                          # exact-arithmetic lead of A
                hA = float(_bf16(x1) * _bf16(a)) - 0.5 * a * a             # what the screen sees
                hB = float(_bf16(x2) * _bf16(b)) - 0.5 * b * b
                d_best = (x1 - a) ** 2 + x2 * x2                           # oracle distance of the winner
                rows.append((x1, x2, gap, hB - hA, d_best))
            rows = [r for r in rows if r[2] > 2e-4 * r[4]]
            rows = sorted(set(rows), key=lambda r: -r[3])[:rows_per_pair]
        cb[2 * p, 2 * p], cb[2 * p + 1, 2 * p + 1] = a, b
        picks.append(rows)
    emax = float(np.sqrt((cb ** 2).sum(1)).max())
    z_rows, ratio = [], []
    for p, rows in enumerate(picks):
        for x1, x2, gap, deficit, _ in rows:
            row = np.zeros(D)
            row[2 * p], row[2 * p + 1] = x1, x2
            z_rows.append(row)
            ratio.append(deficit / (2.0 ** -8 * np.hypot(x1, x2) * emax))
    z = torch.tensor(np.stack(z_rows), dtype=torch.float32)
    return z, torch.tensor(cb, dtype=torch.float32), np.array(ratio)


def test_adversarial_margin():
    """The round-1 margin (2^-8 ||z|| max||e||: a first-order bound for vectors with evenly spread energy) prunes the FP32
    winner on these inputs; the shipped margin (a proven bound for the difference of two scores whose operands are
    BOTH rounded to BF16, built from the measured rounding-error norms) must not."""
    z, cb, ratio = adversarial_case()
    assert np.median(ratio) > 1.15 and ratio.max() < 4.0      # inside the proven bound, mostly outside tau = 1
    assert z.shape[0] >= 128                                    # the tensor path needs a full row tile
    D = z.shape[1]
    lay = ops.layout_of(z.shape, D, 1)
    pcb = ops.prepare_codebook(cb.to(DEV))
    zc = z.to(DEV)
    ref = vq_oracle.nearest(z, cb)
    exact = ops.search(zc, lay, pcb, mode="exact").cpu()
    assert torch.equal(exact, ref)
    assert bool((ref % 2 == 0).all())                           # the A codes win in FP32
    # a margin of 2^-8 ||z|| max||e|| (round 1's default) prunes the FP32 winner: simulate that screen on the host ...
    s_bf = (z.to(torch.bfloat16).double() @ cb.to(torch.bfloat16).double().t()) - 0.5 * (cb.double() ** 2).sum(1)
    old_margin = 2.0 ** -8 * z.norm(dim=1).double() * cb.norm(dim=1).max().double()
    pruned = s_bf.gather(1, ref.view(-1, 1)).squeeze(1) < s_bf.max(1).values - old_margin
    assert int(pruned.sum()) > 100, "construction no longer defeats the first-order margin: strengthen it"
    # ... the shipped margin is built from the MEASURED rounding errors of these operands and keeps it
    sd = ops.screen_debug(zc, lay, pcb, n_cand=4)
    assert bool((sd.margin.cpu().double() > 1.6 * old_margin).all())
    idx = ops.search(zc, lay, pcb, mode="tensor").cpu()
    par = vq_oracle.classify_indices(idx, z, cb)
    assert par.mismatch == 0 and par.exact == par.n, par
    vq = VectorQuantizer(cb.shape[0], D, 0.25, search_mode="tensor").to(DEV)
    assert vq.margin_tau == ops.DEFAULT_MARGIN_TAU == 1.0
    with torch.no_grad():
        vq.embedding.weight.copy_(cb.to(DEV))
        _, _, (_, _, idm) = vq(zc)
    assert torch.equal(idm.view(-1).cpu(), ref)


@pytest.mark.parametrize("kind", ["one_hot_energy", "leaky_relu", "sparse_codes"])
def test_concentrated_energy_latents(kind):
    """Latents whose energy sits in a few channels (what a LeakyReLU encoder tail produces,
    skip_autoencoder.py:346-349) against dense and sparse codebooks: zero mismatches outside documented
    near-ties at the default margin, through the tensor-core path."""
    g = torch.Generator().manual_seed(11)
    N, K, D = 4096, 1024, 256
    cb = torch.randn(K, D, generator=g)
    if kind == "one_hot_energy":
        z = 0.01 * torch.randn(N, D, generator=g)
        z[torch.arange(N), torch.randint(0, D, (N,), generator=g)] = 8.0
    elif kind == "leaky_relu":
        z = torch.nn.functional.leaky_relu(4.0 * torch.randn(N, D, generator=g), 0.2) * 2 ** 0.5
        z = z * (torch.rand(N, D, generator=g) < 0.1)            # 10 % of the channels carry the energy
    else:
        keep = torch.rand(K, D, generator=g) < 0.03
        cb = cb * keep * 6.0
        z = cb[torch.randint(0, K, (N,), generator=g)] + 0.05 * torch.randn(N, D, generator=g)
    z = z.contiguous()
    lay = ops.layout_of(z.shape, D, 1)
    pcb = ops.prepare_codebook(cb.to(DEV))
    idx = ops.search(z.to(DEV), lay, pcb, mode="tensor")
    par = vq_oracle.classify_indices(idx, z, cb)
    assert par.mismatch == 0, par
    exact = ops.search(z.to(DEV), lay, pcb, mode="exact")
    par2 = vq_oracle.classify_indices(exact, z, cb)
    assert par2.mismatch == 0, par2


def test_non_finite_latents_give_in_range_indices():
    """Rows containing NaN / Inf: torch.argmin still returns an index in [0, K); so must we."""
    z, cb = vq_oracle.synth((256, 64), 128, 64, "T", seed=3)
    z[5, 3] = float("nan")
    z[17, :] = float("inf")
    z[40, 0] = float("-inf")
    for mode in ("tensor", "exact"):
        vq = VectorQuantizer(128, 64, 0.25, search_mode=mode).to(DEV).eval()
        with torch.no_grad():
            vq.embedding.weight.copy_(cb.to(DEV))
            _, _, (_, _, idx) = vq(z.to(DEV))
            enc = vq.encode_indices(z.to(DEV))
        for t in (idx.view(-1), enc):
            assert int(t.min()) >= 0 and int(t.max()) < 128, mode
        ok = torch.ones(256, dtype=torch.bool)
        ok[[5, 17, 40]] = False
        ref = vq_oracle.nearest(z[ok], cb)
        assert torch.equal(idx.view(-1).cpu()[ok], ref), mode


# ------------------------------------------------------------------------------------------------
# property tests
# ------------------------------------------------------------------------------------------------
_dims = st.sampled_from([1, 4, 64, 256, 512])
_fast = settings(max_examples=20, deadline=None, derandomize=True,
                 suppress_health_check=[HealthCheck.too_slow, HealthCheck.data_too_large])


def _search_all_modes(z, cb):
    D = cb.shape[1]
    lay = ops.layout_of(z.shape, D, 1)
    pcb = ops.prepare_codebook(cb.to(DEV))
    out = {"exact": ops.search(z.to(DEV), lay, pcb, mode="exact").cpu()}
    if ops.tensor_path_supported(cb.shape[0], D):
        out["tensor"] = ops.search(z.to(DEV), lay, pcb, mode="tensor").cpu()
    return out


@_fast
@given(D=_dims, K=st.integers(1, 700), N=st.integers(1, 900), seed=st.integers(0, 2 ** 16), dist=st.sampled_from("TI"))
def test_property_ragged_shapes_match_oracle(D, K, N, seed, dist):
    """Any N, K (ragged, tiny, K = 1) and every supported D: same indices as the oracle."""
    z, cb = vq_oracle.synth((N, D), K, D, dist, seed=seed)
    for mode, idx in _search_all_modes(z, cb).items():
        par = vq_oracle.classify_indices(idx, z, cb)
        assert par.mismatch == 0, (mode, par)
        if dist == "T" and D >= 64:
            assert par.near_tie == 0, (mode, par)


@_fast
@given(D=st.sampled_from([4, 64, 256]), K=st.integers(2, 300), seed=st.integers(0, 2 ** 16), ndup=st.integers(1, 6))
def test_property_duplicate_rows_take_the_lowest_index(D, K, seed, ndup):
    """quantize.py:50 (torch.argmin) returns the first occurrence: latents equal to a duplicated code resolve to the
    lowest index among the copies, on both search paths."""
    g = torch.Generator().manual_seed(seed)
    z, cb = vq_oracle.synth((256, D), K, D, "T", seed=seed)
    src = torch.randint(0, K, (ndup,), generator=g)
    dst = torch.randint(0, K, (ndup,), generator=g)
    cb[dst] = cb[src]                                    # dst rows become copies of src rows
    z[:ndup] = cb[src]                                   # latents sitting exactly on duplicated codes
    ref = vq_oracle.nearest(z, cb)
    for mode, idx in _search_all_modes(z, cb).items():
        par = vq_oracle.classify_indices(idx, z, cb)
        assert par.mismatch == 0, (mode, par)
        # exact ties: the lowest index among identical rows
        for n in range(ndup):
            same = (cb == cb[ref[n]]).all(1).nonzero().view(-1)
            assert int(idx[n]) == int(same.min()) == int(ref[n]), (mode, n)


@_fast
@given(D=st.sampled_from([4, 64, 256]), K=st.integers(2, 400), seed=st.integers(0, 2 ** 16))
def test_property_codebook_permutation_equivariance(D, K, seed):
    """Permuting the codebook rows permutes the answer: perm[idx'] == idx (distribution T: no ties)."""
    z, cb = vq_oracle.synth((300, D), K, D, "T", seed=seed)
    perm = torch.randperm(K, generator=torch.Generator().manual_seed(seed + 1))
    a = _search_all_modes(z, cb)
    b = _search_all_modes(z, cb[perm].contiguous())
    for mode in a:
        assert torch.equal(perm[b[mode]], a[mode]), mode


@_fast
@given(D=st.sampled_from([4, 64, 256]), K=st.integers(2, 400), seed=st.integers(0, 2 ** 16), e=st.integers(-12, 12))
def test_property_power_of_two_scale_invariance(D, K, seed, e):
    """Scaling latents AND codebook by 2^e is exact in FP32: indices are unchanged, z_q scales by 2^e bit for bit and
    the loss by 4^e."""
    z, cb = vq_oracle.synth((4, D, 4, 8) if D >= 4 else (128, D), K, D, "T", seed=seed)
    s = 2.0 ** e
    outs = []
    for zz, cc in ((z, cb), (z * s, cb * s)):
        vq = VectorQuantizer(K, D, 0.25).to(DEV).eval()
        with torch.no_grad():
            vq.embedding.weight.copy_(cc.to(DEV))
            z_q, loss, (_, _, idx) = vq(zz.to(DEV))
        outs.append((z_q.cpu(), float(loss), idx.view(-1).cpu()))
    assert torch.equal(outs[0][2], outs[1][2])
    assert torch.equal(outs[0][0] * s, outs[1][0])
    assert abs(outs[1][1] - outs[0][1] * s * s) <= 1e-6 * abs(outs[1][1])


@_fast
@given(D=st.sampled_from([4, 64, 256]), K=st.integers(1, 300), g=st.integers(1, 5), hw=st.sampled_from([(1, 1), (2, 3), (4, 4), (5, 7), (8, 8)]),
       seed=st.integers(0, 2 ** 16))
def test_property_forward_matches_oracle(D, K, g, hw, seed):
    """Whole forward through the module on ragged NCHW shapes: z_q bit-exact, loss / perplexity within 1e-5."""
    shape = (g, D, hw[0], hw[1])
    z, cb = vq_oracle.synth(shape, K, D, "T", seed=seed)
    vq = VectorQuantizer(K, D, 0.25).to(DEV).eval()
    with torch.no_grad():
        vq.embedding.weight.copy_(cb.to(DEV))
        z_q, loss, (perp, onehot, idx) = vq(z.to(DEV))
    res = vq_oracle.forward(z, cb, 0.25)
    rows = vq_oracle.to_channel_last(z).reshape(-1, D)
    par = vq_oracle.classify_indices(idx, rows, cb)
    assert par.mismatch == 0, par
    if par.exact == par.n:
        assert torch.equal(z_q.cpu(), res.z_q)
        torch.testing.assert_close(loss.cpu(), res.loss, rtol=1e-5, atol=1e-12)
        torch.testing.assert_close(perp.cpu(), res.perplexity, rtol=1e-5, atol=0)
        assert torch.equal(onehot.sum(0).cpu(), res.one_hot.sum(0))      # an unmodified caller's use of min_encodings


# ------------------------------------------------------------------------------------------------
# deterministic per-code sums (fixed-point accumulation)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape,K,D,mult", [((8, 256, 16, 16), 48, 256, 1), ((6, 64, 5, 7), 20, 32, 2), ((700, 128), 33, 128, 1)])
def test_deterministic_code_sums_are_bit_identical_and_accurate(shape, K, D, mult):
    """`deterministic=True`: embedding.weight.grad is bit-identical from run to run (64-bit fixed-point accumulation
    with integer atomics) and agrees with the closed form evaluated in float64; the default FP32-atomic path agrees
    with the same closed form within the documented tolerance (rtol 1e-4 / atol 1e-7) but not necessarily bit for bit."""
    z, cb = vq_oracle.synth(shape, K, D, "T", seed=21)
    g_zq = torch.randn(shape, generator=torch.Generator().manual_seed(22))
    beta, g_loss = 0.25, 0.7
    lay = ops.layout_of(shape, D, mult)
    rows = (vq_oracle.to_channel_last(z) if len(shape) >= 4 else z).reshape(-1, D)
    idx = vq_oracle.nearest(rows, cb)
    M = z.numel()
    # closed form in float64: dE = (2 beta g / M) (n_k E_k - sum_{i in k} z_i)
    onehot = torch.zeros(rows.shape[0], K, dtype=torch.float64).scatter_(1, idx.view(-1, 1), 1)
    resid64 = onehot.t() @ rows.double() - onehot.sum(0).unsqueeze(1) * cb.double()
    dE64 = -(2 * beta * g_loss / M) * resid64

    def run(det):
        vq = VectorQuantizer(K, D * mult, beta, mult=mult, deterministic=det, search_mode="exact").to(DEV)
        with torch.no_grad():
            vq.embedding.weight.copy_(cb.to(DEV))
        zc = z.to(DEV).requires_grad_(True)
        z_q, loss, _ = vq(zc)
        ((z_q * g_zq.to(DEV)).sum() + loss * g_loss).backward()
        return vq.embedding.weight.grad.detach().cpu().clone(), zc.grad.detach().cpu().clone()

    det = [run(True) for _ in range(4)]
    for dE, dz in det[1:]:
        assert torch.equal(dE, det[0][0]) and torch.equal(dz, det[0][1])
    torch.testing.assert_close(det[0][0].double(), dE64, rtol=1e-5, atol=1e-9)
    fast = run(False)
    torch.testing.assert_close(fast[0].double(), dE64, rtol=1e-4, atol=1e-7)
    assert torch.equal(fast[1], det[0][1])                      # dz is elementwise: identical on both routes
    # the raw statistic through the C ABI: exact to the fixed-point grid
    resid, counts = ops.code_stats_fixed(z.to(DEV), lay, cb.to(DEV), K, idx.to(DEV), sub=1.0, want_counts=True)
    torch.testing.assert_close(resid.cpu().double(), resid64, rtol=2e-6, atol=1e-6)
    assert torch.equal(counts.cpu().long(), onehot.sum(0).long())
    plain, _ = ops.code_stats_fixed(z.to(DEV), lay, None, K, idx.to(DEV), sub=0.0)
    torch.testing.assert_close(plain.cpu().double(), onehot.t() @ rows.double(), rtol=2e-6, atol=1e-6)


# ------------------------------------------------------------------------------------------------
# normalize=True with mult > 1 (quantize.py:56-57): the norm spans the concatenation of several codes
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape,K,C,mult", [((4, 64, 8, 8), 256, 64, 2), ((3, 128, 5, 7), 200, 128, 4), ((2, 3, 32, 4, 4), 64, 32, 2),
                                            ((37, 48), 96, 48, 3)])
def test_normalized_concatenation_forward_backward_match_oracle(shape, K, C, mult):
    """Forward (z_q, loss, perplexity) and backward (dz, dE through the normalisation) of `ccvsq_assign_normalized` /
    `ccvsq_backward_normalized` against the oracle's autograd on the reference op sequence; 4-D, 5-D and 2-D inputs."""
    D = C // mult
    g = torch.Generator().manual_seed(17)
    cb = torch.randn(K, D, generator=g)
    n_pos = int(np.prod(shape)) // C
    idx0 = torch.randint(0, K, (n_pos * mult,), generator=g)
    rows = cb[idx0].view(n_pos, C)
    rows = rows / rows.norm(dim=1, keepdim=True) + 0.05 * torch.randn(n_pos, C, generator=g)     # near normalised codes
    if len(shape) >= 4:
        zl = rows.view(*shape[:-3], shape[-2], shape[-1], C)
        z = zl.transpose(-3, -1).transpose(-2, -1).contiguous()        # channel-last -> the reference's input layout
    else:
        z = rows.view(shape).contiguous()
    g_zq = torch.randn(shape, generator=g)
    # oracle
    zc, cbc = z.clone().requires_grad_(True), cb.clone().requires_grad_(True)
    ref = vq_oracle.forward(zc, cbc, 0.25, mult=mult, normalize=True)
    ((ref.z_q * g_zq).sum() + ref.loss * 1.5).backward()
    # ours
    vq = VectorQuantizer(K, C, 0.25, mult=mult, normalize=True, search_mode="exact").to(DEV)
    with torch.no_grad():
        vq.embedding.weight.copy_(cb.to(DEV))
    zd = z.to(DEV).requires_grad_(True)
    z_q, loss, (perp, _, idx) = vq(zd)
    ((z_q * g_zq.to(DEV)).sum() + loss * 1.5).backward()
    assert torch.equal(idx.view(-1).cpu(), ref.indices.view(-1))
    torch.testing.assert_close(z_q.detach().cpu(), ref.z_q.detach(), rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(loss.detach().cpu(), ref.loss.detach(), rtol=1e-5, atol=0)
    torch.testing.assert_close(perp.cpu(), ref.perplexity.detach(), rtol=1e-5, atol=0)
    torch.testing.assert_close(zd.grad.cpu(), zc.grad, rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(vq.embedding.weight.grad.cpu(), cbc.grad, rtol=1e-4, atol=1e-7)


@_fast
@given(D=st.sampled_from([2, 8, 32]), mult=st.sampled_from([1, 2, 4]), K=st.integers(2, 200), g=st.integers(1, 4),
       hw=st.sampled_from([(1, 1), (2, 3), (4, 4), (5, 7)]), seed=st.integers(0, 2 ** 16))
def test_property_normalized_forward_backward_match_oracle(D, mult, K, g, hw, seed):
    """normalize=True on ragged shapes, one and several codes per position: forward value, loss and both gradients against
    the oracle's autograd on the reference op sequence (rows whose winner is a documented near tie are skipped)."""
    C = D * mult
    shape = (g, C, hw[0], hw[1])
    gen = torch.Generator().manual_seed(seed)
    cb = torch.randn(K, D, generator=gen)
    z = torch.randn(shape, generator=gen)
    g_zq = torch.randn(shape, generator=gen)
    zc, cbc = z.clone().requires_grad_(True), cb.clone().requires_grad_(True)
    ref = vq_oracle.forward(zc, cbc, 0.25, mult=mult, normalize=True)
    ((ref.z_q * g_zq).sum() + ref.loss * 0.75).backward()
    vq = VectorQuantizer(K, C, 0.25, mult=mult, normalize=True).to(DEV)
    with torch.no_grad():
        vq.embedding.weight.copy_(cb.to(DEV))
    zd = z.to(DEV).requires_grad_(True)
    z_q, loss, (perp, _, idx) = vq(zd)
    ((z_q * g_zq.to(DEV)).sum() + loss * 0.75).backward()
    if not torch.equal(idx.view(-1).cpu(), ref.indices.view(-1)):
        rows = vq_oracle.to_channel_last(z).reshape(-1, D)
        assert vq_oracle.classify_indices(idx, rows, cb).mismatch == 0
        return                                               # a near tie resolved the other way: values differ legitimately
    torch.testing.assert_close(z_q.detach().cpu(), ref.z_q.detach(), rtol=2e-6, atol=2e-7)
    torch.testing.assert_close(loss.detach().cpu(), ref.loss.detach(), rtol=1e-5, atol=1e-12)
    torch.testing.assert_close(zd.grad.cpu(), zc.grad, rtol=1e-5, atol=2e-7)
    torch.testing.assert_close(vq.embedding.weight.grad.cpu(), cbc.grad, rtol=1e-4, atol=2e-7)
