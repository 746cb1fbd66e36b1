"""TEST INFRASTRUCTURE -- specification of the SyncBN option of data-parallel training.

The reference is single-process: a batch of n*B query groups goes through ONE batch_normalization per instance
(new_dssm.py:62-88,129-132,151-154).  Data-parallel training with per-replica moments (DPOracle) is a different
model; with SyncBN the n replicas reproduce the single-process step exactly.  This file states which quantities must
cross replicas, as a DSSMOracle whose two BN coupling points call an all-reduce(mean) supplied by the caller:

  forward   mean = avg_r(mean_r);  var = avg_r(var_r + (mean_r - mean)^2)       (equal row counts: Chan's merge)
  backward  dbeta, dgamma = avg_r(sum_rows g_r), avg_r(sum_rows g_r * xhat_r)   with every replica's loss still divided
            by its LOCAL query_BS: g_r is then n times the global-batch g, the average restores the global sums, and the
            kernel formula dh = gamma*rstd*(g - dbeta/n_local - xhat*dgamma/n_local) is unchanged
  update    gradients averaged over replicas, one Adam step -- as without SyncBN

tests/test_syncbn_oracle.py checks that n such replicas (threads with a barrier all-reduce, and two gloo processes)
land on DSSMOracle run on the re-stacked global batch.  Per BN layer that is two small collectives forward
([mean | second moment], 2 x 2L floats) and one backward ([dbeta | dgamma]).
"""
from __future__ import annotations

import threading
from typing import Callable, Dict, List, Sequence

import numpy as np
import scipy.sparse as sp

from .dssm_oracle import DSSMOracle, OracleConfig


class SyncBNReplica(DSSMOracle):
    """One replica; `allreduce_mean(array) -> array` returns the mean of the argument over all replicas."""

    def __init__(self, cfg: OracleConfig, params: Dict[str, np.ndarray], allreduce_mean: Callable[[np.ndarray], np.ndarray],
                 dtype=np.float32):
        super().__init__(cfg, params, dtype)
        self._avg = allreduce_mean

    def _bn_moments(self, x: np.ndarray):
        dt = self.dtype
        mean_l, var_l = super()._bn_moments(x)
        mean = self._avg(mean_l).astype(dt)
        var = self._avg(var_l + (mean_l - mean) ** 2).astype(dt)
        return mean, var

    def _bn_bwd_sums(self, g: np.ndarray, xhat: np.ndarray):
        dt = self.dtype
        b, gm = super()._bn_bwd_sums(g, xhat)
        both = self._avg(np.stack([b, gm])).astype(dt)
        return both[0], both[1]

    def train_step(self, X: sp.csr_matrix) -> float:
        cache = self.forward(X, on_train=True)  # the EMA shadows move with the GLOBAL moments: identical on every replica
        grads = self.backward(cache)
        grads = {k: self._avg(v).astype(self.dtype) for k, v in grads.items()}
        self.adam_update(grads)
        return float(cache["loss"])


class ThreadAllReduce:
    """all-reduce(mean) among n threads of one process (rank order summation in float64: every rank gets the same bits)."""

    def __init__(self, n: int):
        self.n = n
        self._slots: List = [None] * n
        self._bar = threading.Barrier(n)

    def for_rank(self, r: int) -> Callable[[np.ndarray], np.ndarray]:
        def avg(a: np.ndarray) -> np.ndarray:
            a = np.asarray(a)
            self._slots[r] = a
            self._bar.wait()
            out = sum(self._slots[i].astype(np.float64) for i in range(self.n)) / self.n
            self._bar.wait()
            return out.astype(a.dtype)

        return avg


def restack_global(shards: Sequence[sp.csr_matrix], B_local: int, NEG: int) -> sp.csr_matrix:
    """[q ; pos ; neg] of the union of the replicas' query groups, replica 0's groups first."""
    q = [s[:B_local] for s in shards]
    p = [s[B_local:2 * B_local] for s in shards]
    n = [s[2 * B_local:] for s in shards]
    out = sp.vstack(q + p + n, format="csr")
    out.sort_indices()
    return out


def run_syncbn_threads(cfg: OracleConfig, params: Dict[str, np.ndarray], steps: Sequence[Sequence[sp.csr_matrix]],
                       dtype=np.float32) -> List[SyncBNReplica]:
    """n replicas in n threads; steps[t][r] is replica r's batch at step t.  Returns the replicas (all identical)."""
    n = len(steps[0])
    ar = ThreadAllReduce(n)
    reps = [SyncBNReplica(cfg, params, ar.for_rank(r), dtype) for r in range(n)]
    errs: List = []

    def work(r: int):
        try:
            for batches in steps:
                reps[r].train_step(batches[r])
        except BaseException as e:  # pragma: no cover
            errs.append(e)
            ar._bar.abort()

    ths = [threading.Thread(target=work, args=(r,)) for r in range(n)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    if errs:
        raise errs[0]
    return reps
