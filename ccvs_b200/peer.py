"""Training collective over NVLink peer memory (csrc/peer_exchange.cu; C ABI: ccvsq_peer_* in include/ccvsq.h).

The reference moves training state between ranks with NCCL through DDP / apex (tools/engine.py:71-74,127-132).  The EMA
statistics of the quantizer are ONE 1 MB buffer per step — latency, not bandwidth — so on one NVSwitch node they take a
shorter way: every rank pushes its buffer into an inbox on every peer right after its forward, and the EMA update kernel
sums the inboxes in rank order while it rewrites the codebook (`PeerExchange.publish` / `.ema_update`).  No NCCL call on
the step, two ctypes calls of host work, replayable from a CUDA graph.

`torch.distributed` is used once, at set-up, to hand the 64-byte IPC handles around and to agree on whether the
peer path is usable at all (same host, peer access, handles open on every rank); otherwise the caller keeps NCCL.
"""
from __future__ import annotations

import ctypes
import os
import socket
from typing import Optional

import torch
import torch.distributed as dist

from . import _lib, ops

HANDLE_BYTES = 64
MAX_WORLD = 16


class PeerExchange:
    """Exchange areas of all ranks of `group`, mapped into this process.  Construct on every rank at the same point
    (collective); `PeerExchange.create` returns None on every rank when the peer path cannot be used."""

    def __init__(self, K: int, D: int, device: torch.device, rank: int, world: int, own_ptr: int, ptrs, opened):
        self.K, self.D, self.device, self.rank, self.world = K, D, torch.device(device), rank, world
        self._own = own_ptr
        self._opened = opened                      # peer mappings to close
        self.areas = (ctypes.c_void_p * world)(*ptrs)
        self.steps = 0
        # the pushes run here, next to the caller's backward; high priority: their few CTAs are placed as soon as SM slots free
        # up instead of queueing behind the whole backward grid (the peers are waiting for them)
        self._side = torch.cuda.Stream(self.device, priority=-1)
        self._forked = False

    # -- set-up (collective) ---------------------------------------------------------------------------------------
    @staticmethod
    def create(K: int, D: int, device, group: Optional[dist.ProcessGroup] = None) -> Optional["PeerExchange"]:
        if not (dist.is_available() and dist.is_initialized()):
            return None
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        device = torch.device(device)
        if world < 2 or world > MAX_WORLD or device.type != "cuda" or D % 4 != 0:
            return None
        if os.environ.get("CCVSQ_EMA_EXCHANGE", "").lower() == "nccl":
            return None
        L = _lib.load()

        def all_agree(ok: bool) -> bool:
            t = torch.tensor([1 if ok else 0], dtype=torch.int32, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
            return bool(int(t.item()))

        # one node, one process per GPU
        host = socket.gethostname().encode()[:63].ljust(64, b"\0")
        mine = torch.tensor(list(host) + [device.index if device.index is not None else torch.cuda.current_device()],
                            dtype=torch.int32, device=device)
        every = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(every, mine, group=group)
        every = [e.tolist() for e in every]
        same_host = all(e[:64] == every[0][:64] for e in every)
        devs = [e[64] for e in every]
        ok = same_host and len(set(devs)) == world
        if ok:
            me = devs[rank]
            ok = all(d == me or torch.cuda.can_device_access_peer(me, d) for d in devs)
        if not all_agree(ok):
            return None

        nbytes = int(L.ccvsq_peer_exchange_bytes(K, D, world))
        own = ctypes.c_void_p(0)
        handle = (ctypes.c_ubyte * HANDLE_BYTES)()
        with torch.cuda.device(device):
            rc = L.ccvsq_peer_alloc(nbytes, ctypes.byref(own), handle)
        if not all_agree(rc == 0):
            if rc == 0:
                L.ccvsq_peer_free(own)
            return None
        h_mine = torch.tensor(list(handle), dtype=torch.uint8, device=device)
        h_all = [torch.empty_like(h_mine) for _ in range(world)]
        dist.all_gather(h_all, h_mine, group=group)
        ptrs, opened, ok = [], [], True
        with torch.cuda.device(device):
            for r in range(world):
                if r == rank:
                    ptrs.append(own.value)
                    continue
                raw = (ctypes.c_ubyte * HANDLE_BYTES)(*h_all[r].tolist())
                p = ctypes.c_void_p(0)
                if L.ccvsq_peer_open(raw, ctypes.byref(p)) != 0:
                    ok = False
                    break
                ptrs.append(p.value)
                opened.append(p.value)
        if not all_agree(ok):
            with torch.cuda.device(device):
                for p in opened:
                    L.ccvsq_peer_close(ctypes.c_void_p(p))
                L.ccvsq_peer_free(own)
            return None
        return PeerExchange(K, D, device, rank, world, own.value, ptrs, opened)

    # -- per step ------------------------------------------------------------------------------------------------------
    def publish(self, stats: torch.Tensor, overlap: bool = True) -> None:
        """Push this rank's packed statistics [resid: K*D | counts: K] fp32 to every rank.  `overlap`: on a side stream
        that forks from the current one here and is joined by `ema_update` — the NVLink writes (W x 1 MB) then run under
        whatever the caller enqueues in between (the backward pass); the caller keeps `stats` alive until `ema_update`."""
        if stats.numel() != self.K * self.D + self.K or stats.dtype != torch.float32 or not stats.is_contiguous():
            raise ValueError("publish: stats must be the packed fp32 buffer [K*D + K]")
        if overlap:
            self._side.wait_stream(torch.cuda.current_stream(self.device))
            self._forked = True
            with torch.cuda.stream(self._side):
                ops._call("ccvsq_peer_publish", ops._ptr(stats), self.K, self.D, self.areas, self.rank, self.world,
                          ops._stream(self.device))
        else:
            ops._call("ccvsq_peer_publish", ops._ptr(stats), self.K, self.D, self.areas, self.rank, self.world,
                      ops._stream(self.device))

    def ema_update(self, weight: torch.Tensor, n_ema: torch.Tensor, sum_ema: torch.Tensor, decay: float, eps: float) -> None:
        """Wait for every rank's statistics of this step, sum them in rank order and apply the EMA update in place."""
        if self._forked:
            torch.cuda.current_stream(self.device).wait_stream(self._side)     # join: the own push precedes the own update
            self._forked = False
        ops._call("ccvsq_peer_ema_update", ops._ptr(weight), ops._ptr(n_ema), ops._ptr(sum_ema),
                  ctypes.c_void_p(self._own), self.K, self.D, self.world, float(decay), float(eps), ops._stream(self.device))
        self.steps += 1

    def close(self) -> None:
        """Unmap the peers' areas and free the own one (call on every rank after a barrier: a peer may still push)."""
        if self._own is None:
            return
        L = _lib.load()
        with torch.cuda.device(self.device):
            torch.cuda.synchronize(self.device)
            for p in self._opened:
                L.ccvsq_peer_close(ctypes.c_void_p(p))
            L.ccvsq_peer_free(ctypes.c_void_p(self._own))
        self._own, self._opened = None, []
