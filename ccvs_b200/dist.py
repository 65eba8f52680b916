"""Multi-GPU plumbing for the quantizer path: one process per GPU, latents sharded by frame,
codebook replicated (the reference shards the batch the same way, tools/engine.py:86-89).

Forward / encode / decode need NO collective: indices and z_q stay on the owning rank.  Training
statistics (per-code residual sums, counts, squared error) are exchanged with ONE all-reduce over a
packed FP32 buffer  [resid: K*D | counts: K | sq_err hi, lo], after which every rank finalises
identical dE / loss / perplexity / EMA state.  The reference-faithful mode instead leaves
`embedding.weight.grad` local and lets the parent's DDP average it (tools/engine.py:71-74).

Device-agnostic on purpose (pure torch ops + torch.distributed) so the packing/reduction logic is
covered by world_size-2 gloo tests on CPU; on the GPU box the backend is NCCL over NVLink.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def frame_shard(total_frames: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, near-equal [start, stop) block of frames for `rank` (first `rem` ranks get one more)."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of {world}")
    base, rem = divmod(total_frames, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def pack_stats(resid: torch.Tensor, counts: torch.Tensor, sq_err: torch.Tensor) -> torch.Tensor:
    """[K, D] fp32, [K] int32, [1] fp64 -> one fp32 buffer.  Counts are exact in fp32 below 2^24 per
    code; the fp64 squared error travels as a (hi, lo) fp32 pair."""
    K, D = resid.shape
    buf = torch.empty(K * D + K + 2, dtype=torch.float32, device=resid.device)
    buf[: K * D] = resid.reshape(-1)
    buf[K * D: K * D + K] = counts.to(torch.float32)
    hi = sq_err.to(torch.float32)
    lo = (sq_err - hi.to(torch.float64)).to(torch.float32)
    buf[K * D + K] = hi.reshape(())
    buf[K * D + K + 1] = lo.reshape(())
    return buf


def unpack_stats(buf: torch.Tensor, K: int, D: int):
    resid = buf[: K * D].view(K, D)
    counts = buf[K * D: K * D + K].round().to(torch.int32)
    sq = buf[K * D + K].to(torch.float64) + buf[K * D + K + 1].to(torch.float64)
    return resid, counts, sq.reshape(1)


def all_reduce_stats(resid: torch.Tensor, counts: torch.Tensor, sq_err: torch.Tensor,
                     group: Optional[dist.ProcessGroup] = None):
    """Sum the code statistics over all ranks with a single all-reduce; returns (resid, counts, sq_err).
    A no-op when torch.distributed is not initialised or the world has one rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return resid, counts, sq_err
    K, D = resid.shape
    buf = pack_stats(resid, counts, sq_err)
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return unpack_stats(buf, K, D)


def ema_stats_buffer(K: int, D: int, device) -> Tuple[torch.Tensor, torch.Tensor]:
    """Packed buffer for the EMA statistics exchange, [resid: K*D | counts: K] fp32, and the [K, D] view the assign
    pass accumulates the residual sums into directly (no pack copy)."""
    buf = torch.empty(K * D + K, dtype=torch.float32, device=device)   # (the forward zeroes the resid part; an empty
    return buf, buf[: K * D].view(K, D)                                 #  shard zeroes all of it before the collective)


def reduce_ema_stats(buf: torch.Tensor, counts: torch.Tensor, K: int, D: int,
                     group: Optional[dist.ProcessGroup] = None):
    """`buf[:K*D]` already holds this rank's residual sums; the usage counts join it (exact in fp32 below 2^24 per
    code) and ONE all-reduce sums both over the ranks.  Returns (resid [K, D] view, counts int32 [K])."""
    buf[K * D:].copy_(counts)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return buf[: K * D].view(K, D), buf[K * D:].round().to(torch.int32)


def start_reduce_ema_stats(buf: torch.Tensor, counts: torch.Tensor, K: int, D: int,
                           group: Optional[dist.ProcessGroup] = None):
    """First half of `reduce_ema_stats`: pack the counts and ISSUE the all-reduce without making the current stream
    wait for it (NCCL runs it on its own stream, ordered after the work already enqueued on the current one).
    Returns the work handle (None when there is nothing to reduce)."""
    if counts is not None:          # (None: the forward already wrote the fp32 counts into the buffer's tail)
        buf[K * D:].copy_(counts)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        return dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group, async_op=True)
    return None


def finish_reduce_ema_stats(work, buf: torch.Tensor, K: int, D: int):
    """Second half: the current stream waits for the collective (no host block with NCCL); returns
    (resid [K, D] view, counts int32 [K])."""
    if work is not None:
        work.wait()
    return buf[: K * D].view(K, D), buf[K * D:].round().to(torch.int32)


def world_info(group: Optional[dist.ProcessGroup] = None) -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1
