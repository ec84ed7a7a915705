"""Multi-GPU plumbing: one process per GPU, pairs sharded by rank, results gathered to rank 0.

The marginal path has no data-path exchange (pairs are independent), so the only
collective is the gather of per-rank result buffers (rows + per-pair records).  `torch.distributed` is
used as plumbing: NCCL on GPUs, gloo in the CPU tests (tests/test_dist_gloo.py)."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.distributed as dist


def shard_range(n_total: int, rank: int, world: int):
    """Contiguous shard [first, last) of a stream of n_total pairs for `rank` (sizes differ by <= 1)."""
    base, rem = divmod(n_total, world)
    first = rank * base + min(rank, rem)
    return first, first + base + (1 if rank < rem else 0)


class DeviceBytes:
    """Zero-copy view of a raw device allocation as a torch uint8 tensor (for NCCL send/recv)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1",
                                         "data": (int(ptr), False), "version": 3}

    def tensor(self, device) -> torch.Tensor:
        return torch.as_tensor(self, device=device)


def gather_to_root(tensors: Dict[str, torch.Tensor], root: int = 0,
                   recv_cache: Optional[dict] = None) -> Optional[List[Dict[str, torch.Tensor]]]:
    """Gather a dict of 1-D uint8 tensors of rank-dependent sizes to `root`.

    Sizes are exchanged with one all_gather; payloads move with batched isend/irecv (NCCL P2P over
    NVLink on GPUs).  Returns on root a list (one dict per rank, root's own tensors by reference);
    None elsewhere.  `recv_cache` lets the caller reuse receive buffers across steps."""
    rank, world = dist.get_rank(), dist.get_world_size()
    names = sorted(tensors)
    dev = tensors[names[0]].device
    # sizes are exchanged once per (cache, local payload sizes): no host sync per step.  Every rank must call
    # with the same sequence of payload shapes; a rank whose sizes changed re-runs the exchange, and the
    # all_gather then fails loudly on the others instead of posting receives with stale byte counts.
    local = tuple(int(tensors[n].numel()) for n in names)
    cached = None if recv_cache is None else recv_cache.get("__sizes__")
    all_sizes = cached[1] if cached is not None and cached[0] == local else None
    if all_sizes is None:
        sizes = torch.tensor(local, dtype=torch.int64, device=dev)
        gathered = [torch.empty_like(sizes) for _ in range(world)]
        dist.all_gather(gathered, sizes)
        all_sizes = [[int(v) for v in g.tolist()] for g in gathered]
        if recv_cache is not None:
            recv_cache["__sizes__"] = (local, all_sizes)
    ops, out = [], None
    if rank == root:
        out = []
        for r in range(world):
            if r == root:
                out.append({n: tensors[n] for n in names})
                continue
            got = {}
            for i, n in enumerate(names):
                nbytes = int(all_sizes[r][i])
                key = (r, n)
                buf = None if recv_cache is None else recv_cache.get(key)
                if buf is None or buf.numel() < nbytes:
                    buf = torch.empty(nbytes, dtype=torch.uint8, device=dev)
                    if recv_cache is not None:
                        recv_cache[key] = buf
                got[n] = buf[:nbytes]
                if nbytes:
                    ops.append(dist.P2POp(dist.irecv, got[n], r))
            out.append(got)
    else:
        for n in names:
            if tensors[n].numel():
                ops.append(dist.P2POp(dist.isend, tensors[n], root))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return out


class OverlappedGather:
    """Per-step gather of per-rank device buffers to `root`, off the compute stream.

    step(): the payload is copied device-to-device into one of two staging sets on the compute stream (the
    producer may overwrite its buffers as soon as the next step starts), then sent from there on a side stream
    (NCCL send/recv over NVLink), so the transfer overlaps the next step's kernels.  A staging set is reused only
    after its previous transfer has finished (event fence).  finish() makes the compute stream wait for every
    transfer in flight: put it inside the timed region.  bytes_to_root = bytes arriving at the root per step."""

    def __init__(self, payload: Dict[str, torch.Tensor], compute_stream, root: int = 0):
        self.names = sorted(payload)
        self.payload = payload
        self.cs = compute_stream
        self.root = root
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        dev = payload[self.names[0]].device
        self.side = torch.cuda.Stream(device=dev)
        self.stage = [{n: torch.empty_like(payload[n]) for n in self.names} for _ in range(2)]
        self.done = [None, None]
        self.count = 0
        sizes = torch.tensor([payload[n].numel() for n in self.names], dtype=torch.int64, device=dev)
        gathered = [torch.empty_like(sizes) for _ in range(self.world)]
        dist.all_gather(gathered, sizes)
        self.sizes = [[int(v) for v in g.tolist()] for g in gathered]
        self.bytes_to_root = sum(sum(s) for r, s in enumerate(self.sizes) if r != root)
        self.recv = None
        if self.rank == root:
            self.recv = [None if r == root else {n: torch.empty(self.sizes[r][i], dtype=torch.uint8, device=dev)
                                                 for i, n in enumerate(self.names)} for r in range(self.world)]

    def step(self):
        k = self.count % 2
        with torch.cuda.stream(self.cs):
            if self.done[k] is not None:
                self.cs.wait_event(self.done[k])
            for n in self.names:
                self.stage[k][n].copy_(self.payload[n], non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(self.cs)
        with torch.cuda.stream(self.side):
            self.side.wait_event(ready)
            ops = []
            if self.rank == self.root:
                for r in range(self.world):
                    if r != self.root:
                        ops += [dist.P2POp(dist.irecv, self.recv[r][n], r) for n in self.names if self.recv[r][n].numel()]
            else:
                ops = [dist.P2POp(dist.isend, self.stage[k][n], self.root) for n in self.names
                       if self.stage[k][n].numel()]
            if ops:
                for wk in dist.batch_isend_irecv(ops):
                    wk.wait()
            ev = torch.cuda.Event()
            ev.record(self.side)
            self.done[k] = ev
        self.count += 1

    def finish(self):
        for ev in self.done:
            if ev is not None:
                self.cs.wait_event(ev)

    def root_buffers(self):
        """On the root: one dict per rank (the root's own entry is its latest staging set)."""
        assert self.rank == self.root
        own = self.stage[(self.count - 1) % 2]
        return [own if r == self.root else self.recv[r] for r in range(self.world)]
