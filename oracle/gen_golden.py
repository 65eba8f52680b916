"""Generate tests/golden/*.npz from the UNMODIFIED reference quantizer.   *** TEST INFRASTRUCTURE ***

Runs only in the build container, where /root/reference exists.  The reference file
    /root/reference/models/skip_vid_generator/modules/quantize.py
is imported BY PATH (importing it through its package pulls JIT CUDA extensions and cupy,
modules/__init__.py:12-14) and run on CPU with seeded inputs.  Inputs and outputs are stored so
the fixtures are self-contained on the GPU box, where the reference does not exist.

    python oracle/gen_golden.py            # rewrites tests/golden/
"""
from __future__ import annotations

import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import vq_oracle  # noqa: E402

REF = "/root/reference/models/skip_vid_generator/modules/quantize.py"
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def load_reference():
    spec = importlib.util.spec_from_file_location("ref_quantize", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.VectorQuantizer


def run_case(VQ, name, shape, n_e, e_dim_total, mult=1, normalize=False, dist="T", seed=0, cb_override=None,
             z_override=None, code_shape=None, beta=0.25, g_loss=0.37):
    e_dim = e_dim_total // mult
    z, cb = vq_oracle.synth(shape, n_e, e_dim, dist=dist, seed=seed)
    if cb_override is not None:
        cb = cb_override(cb)
    if z_override is not None:
        z = z_override(z, cb)
    torch.manual_seed(seed + 99)
    g_zq = torch.randn(shape)
    vq = VQ(n_e, e_dim_total, beta, mult=mult, normalize=normalize)
    with torch.no_grad():
        vq.embedding.weight.copy_(cb)
    zr = z.clone().requires_grad_(True)
    z_q, loss, (perp, one_hot, idx) = vq(zr)
    ((z_q * g_zq).sum() + loss * g_loss).backward()
    rows = vq_oracle.to_channel_last(z).view(-1, e_dim)
    d = vq_oracle.distances(rows, cb)
    top2 = torch.topk(d, k=min(2, n_e), dim=1, largest=False).values
    rel_gap = ((top2[:, -1] - top2[:, 0]).abs() / top2[:, 0].abs().clamp_min(1e-30)) if n_e > 1 else torch.ones(len(d))
    out = dict(
        z=z.numpy(), codebook=cb.numpy(), beta=np.float32(beta), mult=np.int32(mult), normalize=np.bool_(normalize),
        n_e=np.int32(n_e), e_dim_total=np.int32(e_dim_total),
        z_q=z_q.detach().numpy(), loss=loss.detach().numpy(), perplexity=perp.detach().numpy(),
        indices=idx.numpy(), one_hot_sum=one_hot.sum(0).numpy(),
        g_zq=g_zq.numpy(), g_loss=np.float32(g_loss), dz=zr.grad.numpy(), dE=vq.embedding.weight.grad.numpy(),
        top2_rel_gap=rel_gap.numpy(),
    )
    if code_shape is not None:
        torch.manual_seed(seed + 7)
        code = torch.randint(0, n_e, code_shape)
        out["code"] = code.numpy()
        out["embedded"] = vq.embed_code(code).detach().numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(f"{name:28s} shape={tuple(shape)} K={n_e} D={e_dim} mult={mult} norm={normalize} dist={dist} "
          f"loss={float(loss.detach()):.6f} perp={float(perp):.3f} min_rel_gap={float(rel_gap.min()):.2e}")


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)   # deterministic summation order for the stored outputs
    VQ = load_reference()

    def dup_rows(cb):
        cb = cb.clone()
        cb[5] = cb[2]
        cb[9] = cb[2]
        return cb

    def z_on_dups(z, cb):
        # every latent sits exactly on a codebook row; rows 2, 5, 9 are identical -> index must be 2
        rows = vq_oracle.to_channel_last(z).view(-1, cb.shape[1]).clone()
        pick = torch.arange(rows.shape[0]) % cb.shape[0]
        rows.copy_(cb[pick])
        cl = vq_oracle.to_channel_last(z).shape
        return vq_oracle.to_channel_first(rows.view(cl)).contiguous()

    run_case(VQ, "img_4d_T", (2, 32, 4, 4), 64, 32, seed=1, code_shape=(2, 4, 4))
    run_case(VQ, "vid_5d_T", (2, 3, 32, 4, 4), 64, 32, seed=2, code_shape=(6, 4, 4))
    run_case(VQ, "state_3d_edim1", (4, 16, 2), 128, 1, seed=3, dist="I", code_shape=(4, 16, 2),
             cb_override=lambda cb: torch.rand(128, 1, generator=torch.Generator().manual_seed(33)))
    run_case(VQ, "mult4_4d_T", (2, 16, 3, 3), 32, 16, mult=4, seed=4, code_shape=(2, 3, 12))
    run_case(VQ, "mult2_norm_4d_T", (2, 16, 3, 3), 32, 16, mult=2, normalize=True, seed=5, code_shape=(2, 3, 6))
    run_case(VQ, "norm_4d_T", (3, 32, 4, 4), 64, 32, normalize=True, seed=6)
    run_case(VQ, "dup_rows_ties", (2, 16, 4, 4), 16, 16, seed=7, cb_override=dup_rows, z_override=z_on_dups)
    run_case(VQ, "fresh_init_I", (2, 64, 8, 8), 256, 64, seed=8, dist="I")
    run_case(VQ, "flat_2d_T", (96, 64), 128, 64, seed=9, code_shape=(96,))
    run_case(VQ, "tensor_k1024_d64_T", (2, 64, 16, 16), 1024, 64, seed=10)
    run_case(VQ, "tensor_k256_d256_T", (3, 256, 8, 8), 256, 256, seed=11, code_shape=(3, 8, 8))
    run_case(VQ, "tensor_k300_d128_ragged_T", (3, 128, 7, 9), 300, 128, seed=12)
    run_case(VQ, "tensor_k256_d512_T", (1, 2, 512, 8, 8), 256, 512, seed=13)


if __name__ == "__main__":
    main()
