"""bench.py — latents quantized per second on the CCVS vector-quantizer path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c1|c3|c4|train] [--impl ours|reference]

A *step* is one pass of the hot path over one batch of synthetic latents:
    c2 (default, BASELINE configs[1]): BAIR-256 encode+decode, 64 clips x 16 frames x 16x16 latents,
        K=1024, D=256, FP32, per GPU:  forward(z) -> (z_q, loss, (perplexity, _, indices)) and
        embed_code(indices) through the reference-facing VectorQuantizer API.
    c1: configs[0] geometry (16 frames), c3: Kinetics 8x8 / K=16384 shard, c4: large-codebook stress
        shard (K=16384, D=256, 2^21 latents per GPU), train: c2 shape forward+backward+EMA update.
`value`  : whole-job latents/s with the inputs resident in HBM (CUDA events, max over ranks).
`e2e`    : same metric through the same API with HOST buffers (pinned H2D of z, D2H of indices/loss
           inside the timed region).
`roofline`: the dominant kernel (tcgen05 screening GEMM), 2*N*K*D FLOPs per launch / its mean
           launch duration measured live with CUDA events on the launching stream.
`cpu_baseline` / `--impl reference`: the oracle port of the reference's CPU path (torch-CPU FP32, all
           host threads), timed on a bounded sample of the same workload.
Multi-GPU (torchrun, one rank per GPU): latents are sharded by frame, codebook replicated, no
data-path collective (weak scaling: per-GPU work fixed).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (leading dims (clips, frames), C=D, h, w, K, description)
    "c1": ((1, 16), 256, 16, 16, 1024, "BAIR-256 quantizer forward, 16 frames x 16x16 latents, K=1024, D=256"),
    "c2": ((64, 16), 256, 16, 16, 1024, "BAIR-256 encode+decode, 64 clips x 16 frames x 16x16 latents, K=1024, D=256"),
    "c3": ((128, 16), 256, 8, 8, 16384, "Kinetics-600 64p shard, 128 clips x 16 frames x 8x8 latents, K=16384, D=256"),
    "c3d512": ((64, 16), 512, 8, 8, 16384, "Kinetics-600 shard at the reference script's D=512 (SURVEY F7), 64 clips x 16 frames x 8x8, K=16384"),
    "c4": ((2048, 16), 256, 8, 8, 16384, "large-codebook stress shard, 2^21 latents, K=16384, D=256"),
    "train": ((64, 16), 256, 16, 16, 1024, "training-mode quantizer (fwd+bwd+EMA update), c2 shape"),
}
print_result = print      # replaced in main(): writes to the process's original stdout
METRIC = "latents_quantized_per_sec"
UNIT = "latents/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), d.get("hbm_gbs", 6650.0), "measured"
    return 1590.0, 1400.0, 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons = [], [], set()
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        hi = sorted(sm)[len(sm) // 2:]          # upper half = samples under load
        return {"sm_mhz": statistics.median(hi), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def make_inputs(workload: str, device, seed: int, cb_seed=None):
    """Distribution T: E ~ N(0,1); z = E[randint] + 0.5 N(0,1) laid out [clips, frames, C, h, w].
    `cb_seed`: seed of the codebook when it differs from the latents' (multi-GPU: the codebook is replicated, every
    rank draws its own latents around the SAME codes)."""
    (clips, frames), D, h, w, K, _ = WORKLOADS[workload]
    g = torch.Generator(device=device).manual_seed(seed)
    if cb_seed is None or cb_seed == seed:
        cb = torch.randn(K, D, generator=g, device=device)
    else:
        cb = torch.randn(K, D, generator=torch.Generator(device=device).manual_seed(cb_seed), device=device)
    n = clips * frames * h * w
    z = torch.empty(clips, frames, D, h, w, device=device)
    # build frame by frame blocks to bound temporary memory
    zf = z.view(clips * frames, D, h * w)
    blk = max(1, (1 << 22) // (h * w * D))
    for s in range(0, clips * frames, blk):
        e = min(clips * frames, s + blk)
        m = (e - s) * h * w
        pick = torch.randint(0, K, (m,), generator=g, device=device)
        rows = cb[pick] + 0.5 * torch.randn(m, D, generator=g, device=device)
        zf[s:e] = rows.view(e - s, h * w, D).transpose(1, 2)
    return z, cb, n


def make_inputs_I(workload: str, device, seed: int):
    """Distribution I (stress): E ~ U(-1/K, 1/K) exactly as quantize.py:30, z ~ N(0,1)."""
    (clips, frames), D, h, w, K, _ = WORKLOADS[workload]
    g = torch.Generator(device=device).manual_seed(seed)
    cb = (torch.rand(K, D, generator=g, device=device) * 2 - 1) / K
    z = torch.randn(clips, frames, D, h, w, generator=g, device=device)
    return z, cb, clips * frames * h * w


def stress_record(dev, search_mode: str):
    """SURVEY 8d asks for BOTH distributions: the step of the default workload on distribution I (fresh-init codebook
    U(-1/K, 1/K), z ~ N(0,1): 4-5 % of the rows are near ties by construction, so the screen queues / flags far more rows
    and the FP32 re-scoring and exact fallback carry real work).  Same step as the headline: forward + embed_code."""
    from ccvs_b200 import VectorQuantizer
    (clips, frames), D, h, w, K, _ = WORKLOADS["c2"]
    z, cb, n = make_inputs_I("c2", dev, 1234)
    vq = VectorQuantizer(K, D, 0.25, search_mode=search_mode).to(dev).eval()
    with torch.no_grad():
        vq.embedding.weight.copy_(cb)
        for _ in range(3):
            _, _, (_, _, idx) = vq(z)
            vq.embed_code(idx.view(clips * frames, h, w))
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        a.record()
        for _ in range(reps):
            _, _, (_, _, idx) = vq(z)
            vq.embed_code(idx.view(clips * frames, h, w))
        b.record()
        torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    return {"workload": "c2 shape, distribution I (fresh-init codebook, z ~ N(0,1))", "ms_per_step": ms, "value": n / (ms * 1e-3),
            "unit": UNIT, "note": "stress case: near ties everywhere; the headline distribution is T"}


def rows_on_cpu(z):
    """[clips, frames, C, h, w] on the device -> [N, C] rows in the reference's flatten order (quantize.py:40-42)."""
    clips, frames, D, h, w = z.shape
    return z.view(clips * frames, D, h * w).transpose(1, 2).reshape(-1, D).cpu()


def index_match_record(dev, search_mode: str):
    """Index agreement of the product path with the CPU oracle (SURVEY 8d / BASELINE metric "index match %"):
    2^20 rows of the c2 geometry (four seeded batches) and 2^16 rows at K = 16384 (c3 geometry), distributions T and I.
    The oracle's own FP32 distances classify every differing row: documented near-tie (gap <= 1e-6 relative) or
    mismatch.  The oracle is the checker here, never the thing measured."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import vq_oracle
    from ccvs_b200 import VectorQuantizer
    use_all_host_threads()
    out = {}
    t_start = time.perf_counter()
    plans = {"c2": 4, "c3": 1}          # batches: 4 x 262144 = 2^20 rows ; c3 is cut to 2^16 rows below
    for wl, batches in plans.items():
        (clips, frames), D, h, w, K, _ = WORKLOADS[wl]
        for dist in ("T", "I"):
            tot = {"rows": 0, "exact": 0, "near_ties": 0, "mismatches": 0, "worst_rel_gap": 0.0}
            for b in range(batches):
                z, cb, n = (make_inputs if dist == "T" else make_inputs_I)(wl, dev, 9000 + 17 * b)
                if wl == "c3":
                    z = z[: (1 << 16) // (frames * h * w)].contiguous()
                vq = VectorQuantizer(K, D, 0.25, search_mode=search_mode).to(dev).eval()
                with torch.no_grad():
                    vq.embedding.weight.copy_(cb)
                idx = vq.encode_indices(z).cpu()
                par = vq_oracle.classify_indices(idx, rows_on_cpu(z), cb.cpu())
                tot["rows"] += par.n
                tot["exact"] += par.exact
                tot["near_ties"] += par.near_tie
                tot["mismatches"] += par.mismatch
                tot["worst_rel_gap"] = max(tot["worst_rel_gap"], par.worst_rel_gap)
                del z, cb, vq, idx
            tot["raw_pct"] = 100.0 * tot["exact"] / tot["rows"]
            tot["excl_near_ties_pct"] = 100.0 * (tot["exact"] + tot["near_ties"]) / tot["rows"]
            out[f"{'c2' if wl == 'c2' else 'k16384'}_{dist}"] = dict(tot, K=K, D=D)
    out["oracle"] = ("oracle/vq_oracle.py (torch-CPU FP32 restatement of quantize.py:45-50, row-chunked); near tie = oracle "
                     "distance gap <= 1e-6 relative; T = trained-like, I = fresh-init codebook (4-61 % near ties by construction)")
    out["cpu_seconds"] = time.perf_counter() - t_start
    return out


def k16384_records(dev, search_mode: str, peaks_):
    """The north_star's GEMM target: the screen kernel on the c4 shard (2^21 latents, K = 16384, D = 256), event-timed
    per launch inside the composite call, with clocks sampled while it runs; plus the HBM-bound kernels on the
    reference's real 8x8 layouts (c4, c3d512): per-launch medians."""
    from ccvs_b200 import ops
    bf16_peak, bf16_sust, hbm_peak, peak_src = peaks_
    rec, hbm = None, {}
    for wl in ("c4", "c3d512"):
        (clips, frames), D, h, w, K, desc = WORKLOADS[wl]
        z, cb, n = make_inputs(wl, dev, 1234)
        lay = ops.layout_of(z.shape, D, 1)
        pcb = ops.prepare_codebook(cb)
        for _ in range(2):
            o = ops.quantize_forward(z, lay, cb, 0.25, mode=search_mode, cb=pcb, indices_only=True)
        torch.cuda.synchronize()
        if wl == "c4":
            sampler = ClockSampler(dev.index or 0)
            sampler.start()
            ops.PROFILER.reset(timing={"ccvsq_screen"})
            reps = 50
            for _ in range(reps):
                o = ops.quantize_forward(z, lay, cb, 0.25, mode=search_mode, cb=pcb, indices_only=True)
            torch.cuda.synchronize()
            clocks = sampler.stop()
            calls, tms = ops.PROFILER.summary().get("ccvsq_screen", (0, 0.0))
            ops.PROFILER.reset(timing=False)
            if calls:
                flops = 2.0 * n * K * D
                ach = flops / (tms / calls * 1e-3) / 1e12
                rec = {"kernel": "screen_kernel", "workload": "c4 shard: 2^21 latents, K=16384, D=256 (BASELINE configs[3] per GPU)",
                       "bound": "tensor", "achieved": ach, "peak": bf16_peak, "unit": "TFLOP/s", "frac": ach / bf16_peak,
                       "frac_of_sustained": ach / bf16_sust, "peak_source": f"{peak_src} (burst)", "launches": calls,
                       "ms_per_launch": tms / calls, "algorithmic_flops_per_launch": flops, "clocks": clocks,
                       "latents_per_sec_search_only": n / (tms / calls * 1e-3)}
        idx = o.idx

        def med(fn, reps=15):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            ts = []
            for _ in range(reps):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); fn(); b.record()
                ts.append((a, b))
            torch.cuda.synchronize()
            v = sorted(x.elapsed_time(y) for x, y in ts)
            return v[len(v) // 2]

        zq = torch.empty_like(z)
        hdr = torch.zeros(K + 4, dtype=torch.int32, device=dev)
        sq = torch.zeros(1, dtype=torch.float64, device=dev)
        dec = torch.empty(n, D, device=dev)
        err = torch.zeros(1, dtype=torch.int32, device=dev)
        rl = ops.rows_layout(n, D)
        t_assign = med(lambda: ops._call("ccvsq_assign", ops._ptr(z), lay, ops._ptr(cb), K, ops._ptr(idx), ops._ptr(zq), ops._ptr(sq),
                                         ops._ptr(hdr), ops._stream(dev)))
        t_rows = med(lambda: ops._call("ccvsq_gather", ops._ptr(idx), ops._ptr(cb), K, rl, ops._ptr(dec), ops._ptr(err), ops._stream(dev)))
        t_cm = med(lambda: ops._call("ccvsq_gather", ops._ptr(idx), ops._ptr(cb), K, lay, ops._ptr(dec), ops._ptr(err), ops._stream(dev)))
        # measuring sticks at the same size: a device copy moves assign's bytes (read 4D + write 4D per latent), a device fill
        # writes the decode's bytes — at c3d512 the whole working set (134 MB) is about the size of L2 and a launch is ~30 us,
        # so neither reaches the large-copy peak the fractions below are quoted against
        t_copy = med(lambda: zq.copy_(z))
        t_fill = med(lambda: dec.zero_())
        gb = lambda bytes_, ms: bytes_ / (ms * 1e-3) / 1e9
        hbm[wl] = {
            "layout": [clips, frames, D, h, w], "K": K,
            "assign": {"GBs": gb(n * (8 * D + 8), t_assign), "frac_of_hbm_peak": gb(n * (8 * D + 8), t_assign) / hbm_peak, "ms": t_assign},
            "decode_gather": {"GBs": gb(n * (4 * D + 8), t_rows), "frac_of_hbm_peak": gb(n * (4 * D + 8), t_rows) / hbm_peak, "ms": t_rows},
            "decode_gather_channel_major": {"GBs": gb(n * (4 * D + 8), t_cm), "frac_of_hbm_peak": gb(n * (4 * D + 8), t_cm) / hbm_peak, "ms": t_cm},
            "device_copy_of_assign_bytes_ms": t_copy, "assign_vs_device_copy": t_copy / t_assign,
            "device_fill_of_decode_bytes_ms": t_fill, "decode_gather_vs_device_fill": t_fill / t_rows,
        }
        del z, cb, pcb, o, idx, zq, dec
        torch.cuda.empty_cache()
    hbm["note"] = "per-launch medians (CUDA events around one launch, 15 launches back to back); bytes = SURVEY 8d algorithmic bytes"
    return rec, hbm


def encoder_tail_record(dev, peaks_):
    """the encoder's last 1x1 ConvLayer that produces the c2 latents (1024 frames, 512 -> 256 channels,
    16x16), on the BF16x6 tcgen05 kernel, next to the FP32 library convolution (TF32 off) of the same block."""
    from ccvs_b200 import EncoderTail
    bf16_peak = peaks_[0]
    G, ci, co, h, w = 1024, 512, 256, 16, 16
    x = torch.randn(G, ci, h, w, device=dev)
    m = EncoderTail(ci, co).to(dev)
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False

    def ours():
        with torch.no_grad():
            return m(x)

    def lib():
        with torch.no_grad():
            return torch.nn.functional.leaky_relu(torch.nn.functional.conv2d(x, m.weight * m.scale, m.bias), 0.1)

    def t(fn, reps=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    to, tl = t(ours), t(lib)
    diff = float((ours() - lib()).abs().max())
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    flops = 2.0 * G * h * w * ci * co
    return {"workload": "encoder tail of the c2 batch: ConvLayer(512, 256, 1) + LeakyReLU(0.1) on 262144 positions (skip_autoencoder.py:331)",
            "ms": to, "positions_per_sec": G * h * w / (to * 1e-3), "fp32_accurate_TFLOPs": flops / to / 1e9,
            "bf16_tensor_TFLOPs": 6 * flops / to / 1e9, "frac_of_bf16_peak": 6 * flops / to / 1e9 / bf16_peak,
            "fp32_library_conv_ms": tl, "speedup_vs_fp32_library_conv": tl / to, "max_abs_diff_vs_library": diff,
            "scheme": "three-term BF16 split of both operands, six tcgen05.mma per K step, leading / correction products in "
                      "separate tensor-memory accumulators (FP32-class accuracy: tools/tail_accuracy.py)"}


def train_record(args, dev, world, cb, z, barrier):
    """Training-mode quantizer (BASELINE configs[4]) at the c2 shape on every rank: forward + straight-through backward
    + commitment loss + EMA codebook update, the code statistics exchanged between the ranks once per step.  Variants,
    same steps: `graphed` = the whole step as one CUDA graph with the NVLink peer exchange fused into the EMA update
    (the product path on one node); `peer` = the same exchange issued eagerly; `nccl_overlapped` / `nccl_serial` = one
    NCCL all-reduce of the packed statistics under / in front of the backward; `no_exchange` = sync off -> how much of
    the collective is exposed, and how much of the step is host issue time."""
    import torch.distributed as dist
    from ccvs_b200.quantize import EMAVectorQuantizer, GraphedTrainStep
    K, D = cb.shape
    n_lat = z.numel() // D
    cb = cb.clone()                        # (replicated codebook: make_inputs draws it from the same seed on every rank)
    g_out = torch.randn_like(z)
    zt = z.detach().clone().requires_grad_(True)
    res, notes = {}, {}
    variants = [("graphed", dict(sync=True, overlap=True), True), ("peer", dict(sync=True, overlap=True, exchange="auto"), False),
                ("nccl_overlapped", dict(sync=True, overlap=True, exchange="nccl"), False),
                ("nccl_serial", dict(sync=True, overlap=False, exchange="nccl"), False),
                ("no_exchange", dict(sync=False), False), ("no_exchange_graphed", dict(sync=False), True)]
    for name, kw, graphed in variants:
        if world == 1 and name not in ("no_exchange", "no_exchange_graphed"):
            continue
        vq = EMAVectorQuantizer(K, D, 0.25, decay=0.99, search_mode=args.search_mode, **kw).to(dev).train()
        with torch.no_grad():
            vq.embedding.weight.copy_(cb)
            vq.ema_sum.copy_(cb)
            vq.ema_count.fill_(1.0)

        def tstep():
            zt.grad = None
            z_q, loss, _ = vq(zt)
            torch.autograd.backward([z_q, loss], [g_out, torch.ones_like(loss)])

        try:
            if graphed:
                gs = GraphedTrainStep(vq, zt.detach(), g_out, warmup=max(3, args.warmup))
                tstep = gs.replay                                                            # noqa: F811
            else:
                for _ in range(max(3, args.warmup)):
                    tstep()
                vq.sync_codebook()
        except RuntimeError as e:      # (no peer access on this box: the NCCL variants still run; every rank takes this branch)
            notes[name] = str(e)[:160]
            del vq
            continue
        if name == "peer" and vq._peer is None:
            notes[name] = "peer exchange unavailable: NCCL was used"
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            tstep()
        vq.sync_codebook()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[name] = float(t)
        if world > 1 and name in ("graphed", "peer", "nccl_overlapped"):   # every rank must hold the same codebook after the same steps
            w = vq.embedding.weight.detach()
            lo, hi = w.clone(), w.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            res["codebooks_identical_across_ranks_" + name] = bool(torch.equal(lo, hi)) and bool(torch.isfinite(w).all())
        del vq
    order = ("graphed", "peer", "nccl_overlapped") if world > 1 else ("no_exchange_graphed", "no_exchange")
    main_name = next(n for n in order if n in res)
    main = res[main_name]
    rec = {"workload": "training-mode quantizer at the c2 shape per GPU: fwd + STE backward (dz) + loss + EMA codebook update",
           "ms_per_step": main, "variant": main_name, "value": n_lat * world / (main * 1e-3), "unit": UNIT, "n_gpus": world,
           "ms_per_step_by_variant": res, "exchange_bytes_per_rank": (K * D + K) * 4 if world > 1 else 0,
           "collective": ("NVLink peer exchange: every rank pushes its packed statistics [resid K*D | counts K] into an inbox on "
                          "every peer after its forward; the EMA update kernel waits for the flags, sums the inboxes in rank "
                          "order and rewrites the codebook (csrc/peer_exchange.cu); whole step = one CUDA graph"
                          if main_name == "graphed" else
                          "NVLink peer exchange fused into the EMA update (eager launches)" if main_name == "peer" else
                          "one NCCL all-reduce (SUM, FP32) of [resid K*D | counts K] per step, issued async after the forward, "
                          "waited for at the end of the quantizer's backward" if world > 1 else "none (one rank)")}
    if notes:
        rec["notes"] = notes
    if world > 1 and "no_exchange_graphed" in res and "graphed" in res:
        rec["exposed_exchange_us"] = (res["graphed"] - res["no_exchange_graphed"]) * 1e3
    if world > 1 and "no_exchange" in res:
        for n in ("peer", "nccl_overlapped", "nccl_serial"):
            if n in res:
                rec["exposed_exchange_us_" + n] = (res[n] - res["no_exchange"]) * 1e3
    return rec


_ORIGINAL_AFFINITY = None


def bind_near_gpu(cuda_index: int):
    """Pin this rank's host thread (and with it the pages of the pinned buffers it allocates) to the CPUs NVML
    reports as local to its GPU: with 8 ranks each pulling ~55 GB/s over PCIe, host buffers on the far socket would
    cross the inter-socket link.  Returns a short description for the JSON line; never fatal."""
    global _ORIGINAL_AFFINITY
    try:
        import pynvml
        pynvml.nvmlInit()
        try:
            uuid = str(torch.cuda.get_device_properties(cuda_index).uuid)
            h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
        except Exception:  # noqa: BLE001
            h = pynvml.nvmlDeviceGetHandleByIndex(cuda_index)
        before = os.sched_getaffinity(0)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        after = os.sched_getaffinity(0)
        if not after:
            os.sched_setaffinity(0, before)
            return "unchanged (empty NVML affinity)"
        _ORIGINAL_AFFINITY = before
        return f"{len(after)} of {len(before)} CPUs (NVML affinity of the rank's GPU)"
    except Exception as e:  # noqa: BLE001
        return f"unchanged ({type(e).__name__})"


def use_all_host_threads() -> int:
    """The CPU legs run on every host core this process may use.  torchrun exports OMP_NUM_THREADS=1 to its workers,
    which would silently time the reference on ONE thread: override it explicitly."""
    try:
        if _ORIGINAL_AFFINITY:
            os.sched_setaffinity(0, _ORIGINAL_AFFINITY)      # undo the GPU-local binding for the CPU legs
        n = len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        n = os.cpu_count() or 1
    torch.set_num_threads(max(1, n))
    return torch.get_num_threads()


def run_reference(args):
    """The reference's own CPU path (oracle port, torch-CPU FP32, all host threads) on a bounded
    sample of the workload.  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import vq_oracle
    (clips, frames), D, h, w, K, desc = WORKLOADS[args.workload]
    # bounded sample: at most 16384 latents per step (the N x K fp32 one-hot / distance matrices of the
    # full batch do not fit the reference's dense algorithm), same geometry and distribution
    sample_frames = max(1, min(clips * frames, 16384 // (h * w)))
    if K >= 16384:
        sample_frames = max(1, min(sample_frames, 4096 // (h * w)))
    z, cb = vq_oracle.synth((sample_frames, D, h, w), K, D, "T", seed=1234)
    n = sample_frames * h * w
    threads = use_all_host_threads()

    def step():
        with torch.no_grad():
            res = vq_oracle.forward(z, cb, 0.25)
            vq_oracle.embed_code(res.indices.view(sample_frames, h, w), cb)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    val = n / dt
    sample = f"{n} latents ({sample_frames} frames x {h}x{w}) per step, dense N x K algorithm, distribution T"
    print_result(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "description": desc, "K": K, "D": D, "sample": sample},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def cpu_baseline(workload: str, budget_s: float = 10.0):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import vq_oracle
    (clips, frames), D, h, w, K, _ = WORKLOADS[workload]
    use_all_host_threads()
    sample_frames = max(1, min(clips * frames, (16384 if K < 16384 else 4096) // (h * w)))
    z, cb = vq_oracle.synth((sample_frames, D, h, w), K, D, "T", seed=1234)
    n = sample_frames * h * w

    def step():
        with torch.no_grad():
            res = vq_oracle.forward(z, cb, 0.25)
            vq_oracle.embed_code(res.indices.view(sample_frames, h, w), cb)

    step()
    t0 = time.perf_counter()
    reps = 0
    while reps < 3 or (time.perf_counter() - t0 < budget_s and reps < 200):
        step()
        reps += 1
    dt = (time.perf_counter() - t0) / reps
    return {"value": n / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{reps} passes over {n} latents ({sample_frames} frames x {h}x{w}), dense reference algorithm, "
                      f"torch-CPU FP32, {dt * 1e3:.1f} ms/pass"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--search-mode", default="auto", choices=["auto", "tensor", "exact"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the sub-records of the default line (train, roofline_k16384, index_match, 8x8-layout HBM kernels)")
    ap.add_argument("--no-numa-bind", action="store_true", help="do not pin the rank to its GPU's local CPUs")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    # stdout carries exactly ONE JSON line: anything a library prints there (NCCL's version banner, ...) goes to
    # stderr instead — fd 1 is pointed at fd 2 for the run and the result is written to the saved descriptor
    sys.stdout.flush()
    result_fd = os.dup(1)
    os.dup2(2, 1)
    result_out = os.fdopen(result_fd, "w")
    global print_result
    def print_result(line: str):
        result_out.write(line + "\n")
        result_out.flush()

    if args.impl == "reference":
        run_reference(args)
        return

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback for the product path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    host_affinity = "not requested" if args.no_numa_bind else bind_near_gpu(local_rank)
    import torch.distributed as dist
    if world > 1:
        if os.environ.get("CCVSQ_BENCH_NCCL_HIGH_PRIORITY"):       # (A/B switch: measured no better, see profiles/)
            os.environ.setdefault("TORCH_NCCL_HIGH_PRIORITY", "1")
        dist.init_process_group("nccl", device_id=dev)

    from ccvs_b200 import VectorQuantizer, ops
    from ccvs_b200.quantize import EMAVectorQuantizer

    (clips, frames), D, h, w_, K, desc = WORKLOADS[args.workload]
    z, cb, n_lat = make_inputs(args.workload, dev, 1234 + rank, cb_seed=1234)
    train = args.workload == "train"
    if train:
        vq = EMAVectorQuantizer(K, D, 0.25, decay=0.99, search_mode=args.search_mode).to(dev).train()
        g_out = torch.randn_like(z)
        z.requires_grad_(True)
    else:
        vq = VectorQuantizer(K, D, 0.25, search_mode=args.search_mode).to(dev).eval()
    with torch.no_grad():
        vq.embedding.weight.copy_(cb)
        if train:
            vq.ema_sum.copy_(cb)
            vq.ema_count.fill_(1.0)

    def step(zin):
        if train:
            zin.grad = None
            z_q, loss, (perp, _, idx) = vq(zin)
            torch.autograd.backward([z_q, loss], [g_out, torch.ones_like(loss)])
            return idx, loss, perp, zin.grad
        with torch.no_grad():
            z_q, loss, (perp, _, idx) = vq(zin)
            dec = vq.embed_code(idx.view(clips * frames, h, w_))
        return idx, loss, perp, dec

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident timing ----------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        out = step(z)      # (same object lifetime pattern as the timed loop: the caching allocator reaches
                           #  its steady state here, not inside the timed region)
    barrier()
    # timed region, pass A: exactly the product's launch chain (no events inside it) -> `value`
    ops.PROFILER.reset(timing=False)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_host0 = time.perf_counter()
    ev0.record()
    for _ in range(args.steps):
        out = step(z)
    ev1.record()
    host_issue_ms = (time.perf_counter() - t_host0) * 1e3 / args.steps
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    launches = ops.PROFILER.launches
    # pass B: the same K steps again with the dominant kernel (screen) bracketed by its own CUDA events, recorded
    # inside the composite call on the launching stream -> `roofline` (the two cudaEventRecords sit inside the
    # programmatic-dependent-launch chain, so this pass is NOT the one `value` comes from)
    ops.PROFILER.reset(timing={"ccvsq_screen", "ccvsq_search_exact"})
    evb0, evb1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    evb0.record()
    for _ in range(args.steps):
        out = step(z)
    evb1.record()
    barrier()
    ms_total_b = evb0.elapsed_time(evb1)
    prof_main = ops.PROFILER.summary()
    # keep the GPU under the same load until nvidia-smi has produced a few samples (its period is 100 ms,
    # a short timed region can end before the first one)
    # (every rank runs the same number of extra steps: the training workload contains a collective)
    extra_steps = 0
    if ms_total < 600.0:
        extra_steps = min(5000, max(1, int(800.0 / max(ms_total / args.steps, 1e-3))))
    if world > 1:
        t = torch.tensor([extra_steps], device=dev, dtype=torch.int64)
        dist.broadcast(t, src=0)
        extra_steps = int(t)
    ops.PROFILER.reset(timing=False)
    for _ in range(extra_steps):
        step(z)
    torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["note"] = f"sampled from warm-up through the timed region plus {extra_steps} identical untimed steps"
    # per-kernel breakdown: a separate, untimed-for-`value` pass that issues the same kernels through the
    # step-by-step entry points of the C ABI, each bracketed by events (the product path above makes one
    # composite call per forward / backward, which cannot be bracketed kernel by kernel)
    lay = ops.layout_of(z.shape, D, 1)
    w = vq.embedding.weight.detach()
    zd = z.detach()
    g_one = torch.ones((), device=dev)

    def breakdown_step():
        pcb = ops.prepare_codebook(w)
        idx = ops.search(zd, lay, pcb, mode=args.search_mode)
        _, sq, counts = ops.assign(zd, lay, w, idx)
        ops.finalize(K, D, float(zd.numel()), float(n_lat), 0.25, counts=counts, sq_err=sq, want_loss=True, want_perplexity=True)
        if train:
            # what the EMA training step runs besides the forward: dz only (the codebook takes no gradient) and the
            # EMA update; the per-code residual sums ride on the forward's assign pass (ccvsq_forward_args.resid)
            ops.quantize_backward(zd, lay, w, idx, g_out, g_one, 0.25, want_dE=False)
            ops.ema_update(w_scratch, n_scratch, s_scratch, resid_fixed, counts, 0.99, 1e-5)
        else:
            ops.gather(idx.view(clips * frames, h, w_), w)

    if train:
        resid_fixed = ops.quantize_forward(zd, lay, w, 0.25, want_resid=True).resid
        w_scratch, n_scratch, s_scratch = w.clone(), vq.ema_count.clone(), vq.ema_sum.clone()
    ops.PROFILER.reset(timing=True)
    for _ in range(args.steps):
        breakdown_step()
    torch.cuda.synchronize()
    prof = ops.PROFILER.summary()
    ops.PROFILER.reset(timing=False)
    for k, v in prof_main.items():
        prof[k] = v            # the dominant kernel's figure comes from the timed region itself
    ms_step = ms_total / args.steps
    t = torch.tensor([ms_step], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step_max = float(t)
    value = n_lat * world / (ms_step_max * 1e-3)

    # ---------------- end-to-end with host buffers ----------------
    z_host = torch.empty(z.shape, dtype=z.dtype, pin_memory=True)
    z_host.copy_(z.detach())
    idx_host = torch.empty(n_lat, dtype=torch.int64, pin_memory=True)
    sc_host = torch.empty(2, dtype=torch.float32, pin_memory=True)
    z_stage = torch.empty(z.shape, dtype=z.dtype, device=dev)

    pipe = None
    if not train:
        # frame-chunked host pipeline (ccvs_b200.pipeline): chunk c+1 crosses PCIe while chunk c is quantized
        from ccvs_b200.pipeline import HostQuantizePipeline
        n_chunks = max(d for d in (8, 4, 2, 1) if clips % d == 0)
        pipe = HostQuantizePipeline(vq, z.shape, n_chunks=n_chunks, decode=True)
        idx_host = pipe.idx_host

    def e2e_step():
        if pipe is not None:
            pipe.run(z_host)
            return
        z_stage.copy_(z_host, non_blocking=True)
        zin = z_stage.detach().requires_grad_(True) if train else z_stage
        idx, loss, perp, _ = step(zin)
        idx_host.copy_(idx.view(-1), non_blocking=True)
        sc_host.copy_(torch.stack([loss.detach(), perp]), non_blocking=True)

    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        e2e_step()
    barrier()
    ev0.record()
    for _ in range(e2e_steps):
        e2e_step()
    if pipe is not None:
        torch.cuda.current_stream().wait_stream(pipe.d2h)     # the indices of the last chunk have reached the host
    ev1.record()
    barrier()
    if pipe is not None:   # the pipelined path returns what the whole-batch call returns
        assert torch.equal(pipe.idx_host, out[0].view(-1).cpu()), "pipelined indices differ from the whole-batch call"
    t = torch.tensor([ev0.elapsed_time(ev1) / e2e_steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t)
    e2e = {"value": n_lat * world / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
           "h2d_bytes_per_step": z.numel() * 4, "d2h_bytes_per_step": n_lat * 8 + 8,
           "note": "forward+embed_code through ccvs_b200.VectorQuantizer; z from pinned host memory, indices+loss+"
                   "perplexity read back; decoded latents stay on the device (they feed the decoder there)"
                   + ("; frame-chunked x%d (ccvs_b200.pipeline.HostQuantizePipeline): H2D of chunk c+1 overlaps the "
                      "quantization of chunk c" % pipe.n_chunks if pipe is not None else "")}

    # ---------------- training-mode sub-record (every rank: it contains the design's only collective) ----------------
    # It runs LAST on rank 0 (after every other measurement of the line, right before printing) and behind a guard: a
    # failure inside the collective must not cost the line its headline numbers.  The other ranks enter it right away and
    # wait in its first collective until rank 0 arrives.
    want_train = args.workload == "c2" and not args.no_extras

    def guarded_train_record():
        try:
            return train_record(args, dev, world, cb, z.detach(), barrier)
        except Exception as e:                                   # noqa: BLE001 — reported in the record
            return {"error": f"{type(e).__name__}: {str(e)[:300]}"}

    if rank != 0:
        if want_train:
            guarded_train_record()
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- roofline of the dominant kernel ----------------
    bf16_peak, bf16_sust, hbm_peak, peak_src = peaks()
    breakdown = {k: {"calls": c, "ms_per_step": tms / args.steps} for k, (c, tms) in prof.items()}
    roof = None
    if "ccvsq_screen" in prof:
        calls, tms = prof["ccvsq_screen"]
        flops = 2.0 * n_lat * K * D
        ach = flops / (tms / calls * 1e-3) / 1e12
        traffic = None
        tp = os.path.join(ROOT, "profiles", "screen_traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get(args.workload)
        roof = {"kernel": "screen_kernel (tcgen05 BF16 distance GEMM + fused candidate selection)", "bound": "tensor",
                "achieved": ach, "peak": bf16_peak, "unit": "TFLOP/s", "frac": ach / bf16_peak,
                "frac_of_sustained": ach / bf16_sust, "peak_source": f"{peak_src} (burst; kernel timed alone per launch)",
                "ms_per_launch": tms / calls, "algorithmic_flops_per_launch": flops, "traffic": traffic}
    else:
        name = max(prof, key=lambda k: prof[k][1])
        roof = {"kernel": name, "bound": "hbm", "achieved": None, "peak": hbm_peak, "unit": "GB/s", "frac": None,
                "traffic": None}
    # secondary: decode gather bandwidth (bytes = N*(4D+8))
    extra = {}
    if "ccvsq_gather" in prof:
        calls, tms = prof["ccvsq_gather"]
        gbs = n_lat * (4 * D + 8) / (tms / calls * 1e-3) / 1e9
        extra["decode_gather"] = {"achieved_GBs": gbs, "frac_of_hbm_peak": gbs / hbm_peak, "ms_per_launch": tms / calls}
    if "ccvsq_assign" in prof:
        calls, tms = prof["ccvsq_assign"]
        gbs = n_lat * (8 * D + 8) / (tms / calls * 1e-3) / 1e9
        extra["assign"] = {"achieved_GBs": gbs, "frac_of_hbm_peak": gbs / hbm_peak, "ms_per_launch": tms / calls}
    if "ccvsq_quantize_backward" in prof:
        calls, tms = prof["ccvsq_quantize_backward"]
        gbs = n_lat * (12 * D + 8) / (tms / calls * 1e-3) / 1e9
        extra["backward_dz"] = {"achieved_GBs": gbs, "frac_of_hbm_peak": gbs / hbm_peak, "ms_per_launch": tms / calls,
                                "note": "dz = g_out + (2 g_loss / M)(z - E[idx]): reads z and g_out, writes dz"}

    cpu = None if args.no_cpu_baseline or world > 1 else cpu_baseline(args.workload)

    # ---------------- small-batch latency (what the reference trainer actually issues) ----------------
    small = None
    if not train:
        from ccvs_b200 import VectorQuantizer as _VQ
        zs = z.view(clips * frames, D, h, w_)[:16].contiguous()      # 16 frames: BASELINE configs[0] geometry when h = w = 16
        vqs = _VQ(K, D, 0.25).to(dev).eval()
        with torch.no_grad():
            vqs.embedding.weight.copy_(cb)

        def eager():
            with torch.no_grad():
                _, _, (_, _, i) = vqs(zs)
                vqs.embed_code(i.view(16, -1))

        graphed = vqs.capture(zs, decode=True)

        def time_calls(fn, n=200):
            for _ in range(10):
                fn()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(n):
                fn()
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) / n

        small = {"latents_per_call": int(zs.numel() // D), "layout": list(zs.shape), "K": K, "D": D,
                 "ms_per_call_eager": time_calls(eager), "ms_per_call_cuda_graph": time_calls(graphed.replay),
                 "note": "forward + embed_code; cuda_graph = VectorQuantizer.capture() replay (one driver call per step)"}
        # the autoregressive sampler's per-frame step (quantized_video_model.py:939-964): decode gather -> [decoder +
        # encoder trunk: a 1x1-conv stand-in] -> encoder tail -> quantizer, 16 latent frames of 8x8, one CUDA graph
        from ccvs_b200 import EncoderTail
        from ccvs_b200.reencode import ReencodeStep
        cf = 512
        tail_s = EncoderTail(cf, D).to(dev)
        conv = torch.nn.Conv2d(D, cf, 1).to(dev)
        feats0 = torch.randn(16, cf, 8, 8, device=dev)

        def trunk(zdec):            # stand-in for decoder + encoder trunk: a 1x1 conv on the decoded latents + fixed diverse features
            return feats0 + 1e-3 * torch.tanh(conv(zdec))

        code0 = torch.randint(0, K, (16, 64), device=dev)
        # a codebook that lies NEAR the latents this chain produces (a trained model's regime; against an unrelated random
        # codebook every row is a near-tie and takes the exact fallback)
        vqr = _VQ(K, D, 0.25).to(dev).eval()
        with torch.no_grad():
            rows = tail_s(feats0).permute(0, 2, 3, 1).reshape(-1, D)
            pick = rows[torch.randint(0, rows.shape[0], (K,), device=dev)]      # (K may exceed the 1 024 positions: with replacement)
            vqr.embedding.weight.copy_(pick + 0.02 * rows.std() * torch.randn_like(pick))
        stepper = ReencodeStep(vqr, tail_s, trunk, code0, (8, 8))
        small["reencode_step"] = {"latents_per_call": 16 * 64, "ms_per_call_eager": time_calls(stepper.eager),
                                  "ms_per_call_cuda_graph": time_calls(stepper.step),
                                  "note": "ccvs_b200.reencode.ReencodeStep: embed_code (NCHW) -> stand-in trunk -> EncoderTail -> "
                                          "encode_indices on a frozen codebook; graph = one driver call per generated frame"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step_max, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": args.workload, "description": desc, "K": K, "D": D, "latents_per_gpu": n_lat,
                   "layout": [clips, frames, D, h, w_], "distribution": "T (E~N(0,1), z=E[randint]+0.5N(0,1))",
                   "search_mode": args.search_mode, "screen_operands": "bf16 (fp32 accumulate), fp32 rescoring",
                   "l2": f"inputs larger than L2 ({z.numel() * 4 / 2**20:.0f} MiB of latents per step)",
                   "parallelism": f"frame-sharded x{world}, codebook replicated, no data-path collective",
                   "host_affinity": host_affinity,
                   "timing": "value: K steps of the product launch chain between two CUDA events (no events inside); "
                             "roofline: a second pass of the same K steps with the screen kernel bracketed by events"},
        "e2e": e2e, "gpu_launches": launches, "host_issue_ms_per_step": host_issue_ms, "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
        "kernel_breakdown": breakdown, "hbm_kernels": extra, "small_batch": small,
        "ms_per_step_with_kernel_events": ms_total_b / args.steps,
        "train": None,
    }
    if args.workload == "c2" and world == 1 and not args.no_extras:
        # north_star targets that the default workload does not exercise (rank 0, one GPU)
        torch.cuda.empty_cache()
        line["roofline_k16384"], line["hbm_kernels_8x8"] = k16384_records(dev, args.search_mode, peaks())
        line["encoder_tail"] = encoder_tail_record(dev, peaks())
        line["stress_distribution_I"] = stress_record(dev, args.search_mode)
        if not args.no_cpu_baseline:
            line["index_match"] = index_match_record(dev, args.search_mode)
    if want_train:
        line["train"] = guarded_train_record()
    print_result(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
