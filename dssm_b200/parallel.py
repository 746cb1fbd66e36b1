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
    """Wraps a DSSMTower.  Per step: local forward and dense backward, then the dW1 gather is issued in `n_chunks`
    column chunks; the NCCL all-reduce (AVG) of chunk k -- a contiguous slice of the flat gradient buffer -- runs on
    NCCL's stream while chunk k+1 is being gathered, and Adam on chunk k runs while chunk k+1 is still on the wire.
    The small gradients (FC2.., BN) and the EMA shadows live behind W1 in the same allocation and go out first as ONE
    collective.  Equals one Adam step on the mean gradient (oracle/dssm_oracle.py:DPOracle).
    capture_graph() records the whole step -- kernels and collectives -- into one CUDA graph, so the ~60 launches
    cost no CPU time per step."""

    def __init__(self, tower, group=None, n_chunks: int = 2):
        self.tower = tower
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.n_chunks = max(1, min(int(n_chunks), 64))
        off, rows, cols = tower._layout[0]["W1"]
        self.w1_end = off + ((rows * cols + 3) // 4) * 4
        self.pipelined = tower.conf.layers[0] % 4 == 0 and off == 0
        self.graph = None
        self.graphed = False
        # identical starting parameters on every rank
        if self.world > 1:
            dist.broadcast(tower.params, src=0, group=group)

    # ---- one step on the staging CSR -------------------------------------------------------------------
    def _step_staged(self) -> None:
        t, n = self.tower, self.n_chunks
        t.fwd_bwd_begin_staged()
        if self.world == 1:
            for k in range(n):
                t.backward_w1(k, n)
            t.adam(grad_scale=1.0)
            return
        works = []
        for k in range(n):
            t.backward_w1(k, n)
            off, cnt = t.w1_chunk(k, n)
            if cnt:
                works.append((off, cnt, dist.all_reduce(t.grads[off:off + cnt], op=dist.ReduceOp.AVG, group=self.group, async_op=True)))
        # [grads beyond W1 | EMA shadows] are contiguous in tower.comm: one small collective, queued BEHIND the W1
        # chunks so that its latency does not delay the large transfers (it is only needed by the last Adam call)
        w_rest = dist.all_reduce(t.comm[self.w1_end:], op=dist.ReduceOp.AVG, group=self.group, async_op=True)
        for off, cnt, w in works:
            w.wait()
            t.adam_range(off, cnt, 1.0)
        w_rest.wait()
        t.adam_range(self.w1_end, t.P - self.w1_end, 1.0)
        t.adam_advance()

    def capture_graph(self, warmup: int = 3) -> None:
        """Record the staged step (compute + collectives) into one CUDA graph; falls back to graphing only the compute
        half (C side) if the collectives cannot be captured."""
        if not self.pipelined:
            return
        t = self.tower
        try:
            s = torch.cuda.Stream(device=t.device)
            s.wait_stream(torch.cuda.current_stream(t.device))
            with torch.cuda.stream(s):
                for _ in range(warmup):  # communicators, attribute calls, allocator warm-up outside the capture
                    self._step_staged()
            torch.cuda.current_stream(t.device).wait_stream(s)
            torch.cuda.synchronize(t.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s):
                self._step_staged()
            self.graph = g
        except Exception as e:  # pragma: no cover - depends on the NCCL / driver combination
            import sys

            print(f"[dssm_b200] whole-step graph capture failed ({type(e).__name__}: {e}); graphing the compute half only",
                  file=sys.stderr)
            self.graph = None
            torch.cuda.synchronize(t.device)
            t.capture_graph_dp()
        self.graphed = True

    def train_step(self, x=None) -> torch.Tensor:
        """x: a DeviceCSR, or None to run on whatever is in the tower's staging CSR (tower.stage / staging_views)."""
        t = self.tower
        if not self.pipelined:
            loss = t.forward(x, on_train=True)
            t.backward()
            if self.world > 1:
                dist.all_reduce(t.comm, op=dist.ReduceOp.AVG, group=self.group)
            t.adam(grad_scale=1.0)
            return loss
        if x is not None:
            t.stage(x)
        if self.graph is not None:
            self.graph.replay()
        else:
            self._step_staged()
        return t.tensor("loss")
