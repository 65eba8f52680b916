"""GPU diagnostic for the screen kernel: dumps the score matrix and compares it with a torch BF16
reference, for the single-CTA and the 2-CTA MMA path.  usage: python tools/debug_screen.py [K D N]"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ccvs_b200 import ops

K, D, N = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (300, 128, 500)
dev = "cuda:0"
torch.manual_seed(0)
cb = torch.randn(K, D, device=dev)
z = torch.randn(N, D, device=dev)
lay = ops.rows_layout(N, D)
pcb = ops.prepare_codebook(cb)
torch.cuda.synchronize()
bias = pcb.e_bf16[:K, D:].float().sum(1)
print("bias err", float((bias + 0.5 * (cb ** 2).sum(1)).abs().max()))
ref = z.to(torch.bfloat16).float() @ pcb.e_bf16[:K, :D].float().t() + bias
for cg in (1, 2):
    try:
        sd = ops.screen_debug(z, lay, pcb, n_cand=4, margin_tau=1.0, cta_group=cg, dump_scores=True)
        torch.cuda.synchronize()
    except Exception as e:
        print(f"cta_group={cg}: FAILED {e}")
        break
    sc = sd.scores[:, :K]
    err = (sc - ref).abs()
    print(f"cta_group={cg}: max err {float(err.max()):.4g}  nan {int(sc.isnan().sum())}  idx agree "
          f"{float((sd.idx == ref.argmax(1)).float().mean()):.4f}  queued {int(sd.queue.count)}")
    if float(err.nan_to_num(1e9).max()) > 0.05:
        bad = (err.nan_to_num(1e9) > 0.05)
        print("  bad rows:", bad.any(1).nonzero().flatten()[:16].tolist(), " bad cols:", bad.any(0).nonzero().flatten()[:32].tolist())
        print("  ours[0,:8]", sc[0, :8].tolist()); print("  ref [0,:8]", ref[0, :8].tolist())
        print("  ours-ref-bias? [0,:4]", (sc[0, :4] - bias[:4]).tolist(), (ref[0, :4] - bias[:4]).tolist())
