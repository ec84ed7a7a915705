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
    all_sizes = None if recv_cache is None else recv_cache.get("__sizes__")
    if all_sizes is None:  # sizes are exchanged once per (cache, payload shape): no host sync per step
        sizes = torch.tensor([tensors[n].numel() for n in names], dtype=torch.int64, device=dev)
        gathered = [torch.empty_like(sizes) for _ in range(world)]
        dist.all_gather(gathered, sizes)
        all_sizes = [[int(v) for v in g.tolist()] for g in gathered]
        if recv_cache is not None:
            recv_cache["__sizes__"] = all_sizes
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
