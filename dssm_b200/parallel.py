"""Data-parallel training (new capability -- the reference is single-process).

One process per GPU (torchrun); every rank holds the full parameters and runs the reference graph on its own
query groups with per-replica BatchNorm moments.  After the backward the flat gradient buffer is summed with
one NCCL all-reduce over NVLink and Adam runs with grad_scale = 1/world_size, which equals one Adam step on the
mean gradient (oracle/dssm_oracle.py:DPOracle).  The EMA shadows are averaged the same way (linear in the batch
statistics, so this equals averaging the statistics first).
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import scipy.sparse as sp
import torch
import torch.distributed as dist

from .batch import StackedBatch


def shard_stacked_batch(b: StackedBatch, query_BS_global: int, NEG: int, rank: int, world: int) -> StackedBatch:
    """Rank r of n takes query groups [r*B/n, (r+1)*B/n) of a global batch: their query, positive and negative
    rows, re-stacked [q ; pos ; neg] (neg rows of group j stay contiguous at j*NEG)."""
    if query_BS_global % world:
        raise ValueError("global query_BS must divide evenly over the ranks")
    Bl = query_BS_global // world
    X = b.to_scipy()
    B = query_BS_global
    lo, hi = rank * Bl, (rank + 1) * Bl
    parts = [X[lo:hi], X[B + lo:B + hi], X[2 * B + lo * NEG:2 * B + hi * NEG]]
    Y = sp.vstack(parts, format="csr")
    Y.sort_indices()
    return StackedBatch(Y.indptr.astype(np.int32), Y.indices.astype(np.int32), Y.data.astype(np.float32), b.n_cols)


def allreduce_mean_(tensors: List[torch.Tensor], group=None) -> None:
    """In-place mean over ranks (sum all-reduce, then scale).  Works on NCCL (GPU) and gloo (CPU tests)."""
    world = dist.get_world_size(group)
    for t in tensors:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        t.mul_(1.0 / world)


class DataParallelTower:
    """Wraps a DSSMTower: forward+backward locally, all-reduce of the flat grads (+ EMA), Adam with 1/world."""

    def __init__(self, tower, group=None):
        self.tower = tower
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        # identical starting parameters on every rank
        if self.world > 1:
            dist.broadcast(tower.params, src=0, group=group)

    def train_step(self, x) -> torch.Tensor:
        t = self.tower
        loss = t.forward(x, on_train=True)
        t.backward()
        if self.world > 1:
            dist.all_reduce(t.grads, op=dist.ReduceOp.SUM, group=self.group)
            if t.conf.use_bn:
                dist.all_reduce(t.ema, op=dist.ReduceOp.SUM, group=self.group)
                t.ema.mul_(1.0 / self.world)
        t.adam(grad_scale=1.0 / self.world)
        return loss
