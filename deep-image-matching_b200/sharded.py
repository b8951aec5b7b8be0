"""Multi-GPU driver of the hot path: one process per GPU (torch.distributed), independent pairs sharded
across ranks with no data-path collective; the only exchange is the final gather of the variable-length match
tables to rank 0.  Backend nccl on GPUs, gloo in the CPU tests."""
from __future__ import annotations

import numpy as np


def shard_pairs(n_pairs: int, world: int, rank: int, costs=None) -> list:
    """Indices of the pairs rank `rank` processes.  Round-robin by default (pairs_from_bruteforce order,
    pairs_generator.py:37-38); with per-pair costs (e.g. N0*N1) a longest-processing-time deal balances the
    early-exit variance.  Deterministic on every rank."""
    if costs is None:
        return list(range(rank, n_pairs, world))
    order = sorted(range(n_pairs), key=lambda i: (-float(costs[i]), i))
    load = [0.0] * world
    mine = []
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        load[r] += float(costs[i])
        if r == rank:
            mine.append(i)
    return sorted(mine)


def gather_match_tables(local_ids, local_matches, n_pairs: int, dist=None, device=None):
    """Gather {pair id -> int64 (S,2)} from all ranks to rank 0 (counts all_gather + padded all_gather).
    Returns the full list on rank 0 (None elsewhere)."""
    import torch

    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        out = [None] * n_pairs
        for i, m in zip(local_ids, local_matches):
            out[i] = np.asarray(m, np.int64).reshape(-1, 2)
        return out
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = device if device is not None else torch.device("cpu")
    k = len(local_ids)
    cnt = torch.tensor([k, max([len(m) for m in local_matches], default=0)], dtype=torch.int64, device=dev)
    cnts = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(cnts, cnt)
    kmax, smax = int(max(c[0] for c in cnts)), int(max(c[1] for c in cnts))
    buf = torch.full((max(kmax, 1), 2 + 2 * max(smax, 1)), -1, dtype=torch.int64, device=dev)
    for j, (i, m) in enumerate(zip(local_ids, local_matches)):
        m = torch.as_tensor(np.asarray(m, np.int64).reshape(-1, 2), device=dev)
        buf[j, 0], buf[j, 1] = i, m.shape[0]
        if m.shape[0]:
            buf[j, 2:2 + 2 * m.shape[0]] = m.reshape(-1)
    bufs = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(bufs, buf)
    if rank != 0:
        return None
    out = [None] * n_pairs
    for r in range(world):
        b = bufs[r].cpu().numpy()
        for j in range(int(cnts[r][0])):
            i, s = int(b[j, 0]), int(b[j, 1])
            out[i] = b[j, 2:2 + 2 * s].reshape(-1, 2).copy()
    return out
