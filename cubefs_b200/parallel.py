"""Multi-GPU host logic: stripes are independent, so a batch is partitioned
contiguously over the ranks (one process per GPU) and the only shared state is the m x k coding
matrix, broadcast from rank 0 with torch.distributed (NCCL on GPUs, gloo in the CPU tests).
No data-path collective exists: weak scaling."""
from __future__ import annotations

from typing import Tuple

import numpy as np


def partition(n_stripes: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous slice [first, last) of the stripe batch owned by `rank` (stripe s -> rank s*world//n)."""
    return n_stripes * rank // world, n_stripes * (rank + 1) // world


def parity_rows_host(k: int, m: int) -> np.ndarray:
    """The m x k parity rows of reedsolomon.New(k, m) (default options), computed on the host by the
    product-side generator (cubefs_b200/csrc/gen_bitslice.py) -- no GPU, no oracle."""
    from .csrc import gen_bitslice
    return np.array(gen_bitslice.parity_rows(k, m), dtype=np.uint8)


def broadcast_matrix(rows, k: int, m: int, src: int = 0, device=None):
    """Broadcast the coding matrix from `src`; every rank returns the same m x k uint8 array."""
    import torch
    import torch.distributed as dist
    t = torch.zeros((m, k), dtype=torch.uint8, device=device)
    if dist.get_rank() == src:
        t.copy_(torch.as_tensor(np.ascontiguousarray(rows, dtype=np.uint8)))
    dist.broadcast(t, src=src)
    return t.cpu().numpy()


def max_over_ranks(value: float, device=None) -> float:
    """Timing rule: a multi-GPU step takes as long as its slowest rank."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
