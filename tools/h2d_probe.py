"""Bare host->device copy ceiling of the box, next to which bench.py's end-to-end numbers are read.

    python tools/h2d_probe.py                                   # one rank
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/h2d_probe.py [--bind]

Every rank copies a pinned FP32 buffer of c2's size (268 MB) to its GPU with cudaMemcpyAsync, all ranks at the
same time (barrier), timed with CUDA events on the copy stream; the aggregate is sum(bytes) / max-over-ranks
time.  Variants: one whole-buffer copy, 8 chunk copies on one stream (what HostQuantizePipeline issues), the
chunks alternating over two streams, and a simultaneous D2H of 2 MB (indices) per chunk.  --bind pins the rank
to the CPUs NVML reports as local to its GPU before the pinned buffer is allocated (bench.py's default).
Prints one JSON line on rank 0.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bind", action="store_true")
    ap.add_argument("--mb", type=int, default=256)
    ap.add_argument("--reps", type=int, default=20)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    aff = "unbound"
    if args.bind:
        import bench
        aff = bench.bind_near_gpu(local)
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = args.mb * (1 << 20) // 4
    host = torch.empty(n, dtype=torch.float32, pin_memory=True)
    host.normal_()
    devbuf = torch.empty(n, dtype=torch.float32, device=dev)
    back_d = torch.empty(1 << 18, dtype=torch.int64, device=dev)      # 2 MB
    back_h = torch.empty(1 << 18, dtype=torch.int64, pin_memory=True)
    s1, s2, s3 = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn):
        for _ in range(3):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s1)
        for _ in range(args.reps):
            fn()
        s1.wait_stream(s2)
        s1.wait_stream(s3)
        e1.record(s1)
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) / args.reps], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    def whole():
        with torch.cuda.stream(s1):
            devbuf.copy_(host, non_blocking=True)

    def chunks(streams, with_d2h=False):
        c = n // 8
        for i in range(8):
            st = streams[i % len(streams)]
            with torch.cuda.stream(st):
                devbuf[i * c:(i + 1) * c].copy_(host[i * c:(i + 1) * c], non_blocking=True)
            if with_d2h:
                with torch.cuda.stream(s3):
                    back_h[: (1 << 18) // 8].copy_(back_d[: (1 << 18) // 8], non_blocking=True)

    out = {}
    for name, fn in (("whole_buffer", whole), ("8_chunks_one_stream", lambda: chunks([s1])),
                     ("8_chunks_two_streams", lambda: chunks([s1, s2])),
                     ("8_chunks_plus_d2h", lambda: chunks([s1], True))):
        ms = timed(fn)
        out[name] = {"ms": ms, "GBs_per_rank": n * 4 / ms / 1e6, "GBs_aggregate": world * n * 4 / ms / 1e6}
    if rank == 0:
        print(json.dumps({"probe": "h2d", "ranks": world, "bytes_per_rank": n * 4, "affinity": aff,
                          "cpus_visible": len(os.sched_getaffinity(0)), "variants": out}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
