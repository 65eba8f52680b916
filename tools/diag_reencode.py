"""Per-kernel device time of one ReencodeStep (eager), to see what a generated frame costs."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ccvs_b200 import VectorQuantizer, EncoderTail, ops
from ccvs_b200.reencode import ReencodeStep
dev = torch.device("cuda", 0)
K, D, cf = 1024, 256, 512
vq = VectorQuantizer(K, D, 0.25).to(dev).eval()
tail = EncoderTail(cf, D).to(dev)
trunk = torch.nn.Sequential(torch.nn.Conv2d(D, cf, 1), torch.nn.Tanh()).to(dev)
code0 = torch.randint(0, K, (16, 64), device=dev)
with torch.no_grad():
    for _ in range(3):
        lat = tail(trunk(vq.embed_code(code0.view(16, 8, 8), channel_major_hw=(8, 8))))
        rows = lat.permute(0, 2, 3, 1).reshape(-1, D)
        pick = rows[torch.randint(0, rows.shape[0], (K,), device=dev)]
        vq.embedding.weight.copy_(pick + 0.02 * rows.std() * torch.randn_like(pick))
st = ReencodeStep(vq, tail, trunk, code0, (8, 8))
def t(fn, n=100):
    for _ in range(10): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
print(f"eager {t(st.eager):.1f} us   graph {t(st.step):.1f} us")
with torch.no_grad():
    z = vq.embed_code(code0.view(16, 8, 8), channel_major_hw=(8, 8)); f = trunk(z); lat = tail(f)
    print(f"embed_code {t(lambda: vq.embed_code(code0.view(16, 8, 8), channel_major_hw=(8, 8))):.1f} us | trunk {t(lambda: trunk(z)):.1f} us | "
          f"tail {t(lambda: tail(f)):.1f} us | encode_indices {t(lambda: vq.encode_indices(lat)):.1f} us")
    lay = ops.layout_of(lat.shape, D, 1)
    pcb = ops.prepare_codebook(vq.embedding.weight)
    idx, q = ops.screen(lat, lay, pcb)
    print("queued rows", int(q.count), "of", lay.rows, "flagged", int((q.flags[:int(q.count)] != 0).sum()))
    sd = ops.screen_debug(lat, lay, pcb, n_cand=8, dump_scores=True)
    s = sd.scores[:, :K]
    top2 = s.topk(2, dim=1).values
    print("rows", tuple(rows.shape), "row norm mean", float(rows.norm(dim=1).mean()), "spread (std over rows per channel, mean)", float(rows.std(0).mean()),
          "mean vector norm", float(rows.mean(0).norm()))
    print("margin mean", float(sd.margin.mean()), "top1-top2 gap: median", float((top2[:, 0] - top2[:, 1]).median()), "min", float((top2[:, 0] - top2[:, 1]).min()))
    print("e_max", pcb.e_max.tolist(), "lat finite", bool(torch.isfinite(lat).all()), "lat absmax", float(lat.abs().max()))
    print("flags hist", torch.bincount(sd.flags.long(), minlength=4).tolist(), "cands>=0 per row mean", float((sd.cand_idx >= 0).sum(1).float().mean()))
