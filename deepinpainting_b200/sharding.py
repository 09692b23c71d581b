"""Multi-GPU partitioning of the shift layer (one process per GPU, torch.distributed).

* batch sharding -- images are independent (models/IPSRFunction.py:46 ``for idx in range(bz)``), so
  each rank runs the layer on its slice of the batch with NO collective;
* bank sharding  -- for large / reference-guided inputs every rank holds the same images but
  correlates against its own slice of the bank columns; the one exchange step is an all-reduce MAX
  of order-preserving int64 (score, index) keys (NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_batch(batch: int, world_size: int, rank: int) -> Tuple[int, int]:
    """[begin, end) of the images rank ``rank`` owns; sizes differ by at most one."""
    if not (0 <= rank < world_size):
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    base, rem = divmod(batch, world_size)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_bank(n_cols: int, world_size: int, rank: int, align: int = 128) -> Tuple[int, int]:
    """[begin, end) of the bank columns rank ``rank`` correlates against, in units of ``align``
    columns (the tcgen05 path consumes 128-column tiles).  Ranks beyond the number of tiles get an
    empty range and contribute the identity key."""
    if n_cols % align != 0:
        raise ValueError("bank of %d columns is not a multiple of %d" % (n_cols, align))
    b, e = shard_batch(n_cols // align, world_size, rank)
    return b * align, e * align


KEY_IDENTITY = -(1 << 63)


def allreduce_max_keys(keys: torch.Tensor, group=None) -> torch.Tensor:
    """The exchange step of the bank-sharded mode: in-place all-reduce MAX of int64 keys
    (high 32 bits = orderable fp32 score, low 32 bits = ~index, so the largest score and, on ties,
    the LOWEST index win -- torch.max semantics of util/MaxCoord.py:22)."""
    if keys.dtype != torch.int64:
        raise TypeError("keys must be int64")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(keys, op=dist.ReduceOp.MAX, group=group)
    return keys
