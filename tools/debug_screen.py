"""GPU diagnostic for the tcgen05 screening kernel: dumps the raw score matrix and compares it with
<bf16(z), bf16(e)> - 0.5|e|^2 computed by torch (fp32 accumulate).  Prints error statistics per
row block / column block so a wrong smem descriptor, swizzle or TMEM mapping is visible at once."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from ccvs_b200 import ops  # noqa: E402

DEV = "cuda:0"


def run(N, K, D, seed=0):
    torch.manual_seed(seed)
    cb = torch.randn(K, D, device=DEV)
    z = torch.randn(N, D, device=DEV)
    lay = ops.rows_layout(N, D)
    pcb = ops.prepare_codebook(cb)
    zb, margin = ops.pack_latents(z, lay, pcb, margin_tau=1.0)
    sr, scores = ops.screen_dump(zb, margin, pcb, N, n_cand=4)
    torch.cuda.synchronize()
    ref = zb.float() @ pcb.e_bf16.float().t() + pcb.bias
    got = scores
    K_pad = pcb.e_bf16.shape[0]
    valid = torch.zeros_like(ref, dtype=torch.bool)
    valid[:N, :K] = True
    err = (got - ref).abs()
    err[~valid] = 0
    nan = torch.isnan(got) & valid
    print(f"[N={N} K={K} D={D}] max|err|={float(err[~nan].max()) if (~nan).any() else float('nan'):.4g} "
          f"nan(unwritten)={int(nan.sum())} of {int(valid.sum())}")
    if float(err[~nan].max()) > 0.05 or int(nan.sum()):
        # localise: error by 32-row block x 64-col block
        Np = got.shape[0]
        e = torch.where(nan, torch.full_like(err, 1e9), err)
        blk = e.view(Np // 32, 32, K_pad // 64, 64).amax(dim=(1, 3))
        print("  per (32-row, 64-col) block max err (first 8x8):")
        print(blk[:8, :8].cpu())
        print("  got[0,:8]", got[0, :8].cpu().tolist())
        print("  ref[0,:8]", ref[0, :8].cpu().tolist())
        print("  got[1,:8]", got[1, :8].cpu().tolist())
        print("  ref[1,:8]", ref[1, :8].cpu().tolist())
    top = ref[:N, :K].max(dim=1)
    ci, sc, flag = sr.merged()
    best = ci.gather(1, sc.argmax(1, keepdim=True)).squeeze(1).long()
    ok = (best == top.indices)
    print(f"  top-1 candidate == torch argmax on {float(ok.float().mean()) * 100:.3f}% of rows; "
          f"flags set on {int(flag.sum())} rows; mean #cands {float((ci >= 0).sum(1).float().mean()):.2f}")
    return float(err[~nan].max()) if (~nan).any() else float("inf")


if __name__ == "__main__":
    worst = 0.0
    for (N, K, D) in [(128, 256, 64), (128, 256, 256), (256, 512, 128), (1000, 1000, 256), (384, 1024, 512), (4096, 16384, 256)]:
        worst = max(worst, run(N, K, D))
    print("WORST", worst)
