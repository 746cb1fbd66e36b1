"""Data-parallel training (new capability -- the reference is single-process).

One process per GPU (torchrun); every rank holds the full parameters and runs the reference graph on its own
query groups with per-replica BatchNorm moments.  After the backward the flat gradient buffer is summed with
one NCCL all-reduce over NVLink and Adam runs with grad_scale = 1/world_size, which equals one Adam step on the
mean gradient (oracle/dssm_oracle.py:DPOracle).  The EMA shadows are averaged the same way (linear in the batch
statistics, so this equals averaging the statistics first).
"""
from __future__ import annotations

import os
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


def _all_ranks_agree(flag: bool, device, group=None) -> bool:
    """True iff `flag` is true on EVERY rank (all-reduce MIN): mode decisions must be taken together."""
    t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    return bool(t.item())


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

    def __init__(self, tower, group=None, n_chunks: int = 2, comm: str = "nccl", multicast: Optional[bool] = None,
                 sync_bn: bool = False):
        if comm not in ("nccl", "nvlink"):
            raise ValueError("comm must be 'nccl' or 'nvlink'")
        self.tower = tower
        self.group = group
        self.comm = comm
        self.multicast = multicast  # None: use NVSwitch multicast (NVLS) when the symmetric-memory handle offers it
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.n_chunks = max(1, min(int(n_chunks), 64))
        off, rows, cols = tower._layout[0]["W1"]
        self.w1_end = off + ((rows * cols + 3) // 4) * 4
        self.pipelined = tower.conf.layers[0] % 4 == 0 and off == 0
        self.graph = None
        self.graphed = False
        self.launches_per_step = 0
        self.replays = 0
        self.sync_bn = False
        # identical starting state on every rank: parameters AND optimizer state (Adam m, v, beta powers) AND the EMA
        # shadows -- rank 0 may have restored a checkpoint; replicas whose slots differ would apply different updates to
        # the same averaged gradient and drift apart silently
        if self.world > 1:
            for buf in (tower.params, tower.m, tower.v, tower.beta_pow, tower.ema):
                dist.broadcast(buf, src=0, group=group)
        if comm == "nvlink" and self.world > 1:
            # the choice of exchange must be COLLECTIVE: a rank that fell back to NCCL alone would wait in an all-reduce
            # while the others wait in the symmetric-memory rendezvous
            ok, why = True, ""
            if not (self.pipelined and getattr(tower, "symmetric", False)):
                ok, why = False, "needs DSSMTower(..., symmetric=True) and an FC1 width that is a multiple of 4"
            if self.world > 16:
                ok, why = False, "addresses at most 16 peers"
            if not _all_ranks_agree(ok, tower.device, group):
                ok = False
            if ok:
                try:
                    self._setup_nvlink()
                except (RuntimeError, ImportError) as e:  # rendezvous itself is collective: it fails on every rank or none
                    ok, why = False, f"{type(e).__name__}: {e}"
                if not _all_ranks_agree(ok, tower.device, group):
                    ok = False
            if not ok:
                import sys

                print(f"[dssm_b200] rank {self.rank}: NVLink peer-memory exchange unavailable ({why or 'another rank cannot use it'}); "
                      "all ranks use the NCCL all-reduce", file=sys.stderr)
                self.comm = "nccl"
        if sync_bn and self.world > 1:
            self._setup_syncbn()

    # ---- NVLink peer-memory exchange of dW1 (csrc/nvlink.cu) --------------------------------------------------
    def _setup_nvlink(self) -> None:
        """Rendezvous of the tower's symmetric params / grads buffers: afterwards every rank holds device pointers to
        every other rank's W1 and dW1.  Rank r owns the W1 rows [r*D/n, (r+1)*D/n): it alone reduces their gradient, runs
        Adam on them and writes the result into all replicas."""
        import ctypes as C

        import torch.distributed._symmetric_memory as symm_mem

        t = self.tower
        if not (self.pipelined and getattr(t, "symmetric", False)):
            raise ValueError("comm='nvlink' needs DSSMTower(..., symmetric=True) and an FC1 width that is a multiple of 4")
        if self.world > 16:
            raise ValueError("comm='nvlink' addresses at most 16 peers")
        grp = self.group if self.group is not None else dist.group.WORLD
        self._h_params = symm_mem.rendezvous(t.params, grp)
        self._h_comm = symm_mem.rendezvous(t.comm, grp)
        w1_off = t._layout[0]["W1"][0]  # 0 (checked by self.pipelined)
        arr = C.c_void_p * self.world
        self._peer_w = arr(*[int(p) + 4 * w1_off for p in self._h_params.buffer_ptrs])
        self._peer_dw = arr(*[int(p) + 4 * w1_off for p in self._h_comm.buffer_ptrs])  # grads = comm[:P]
        has_mc = bool(getattr(self._h_params, "has_multicast_support", False)) and bool(self._h_params.multicast_ptr) \
            and bool(self._h_comm.multicast_ptr)
        # measured (C2, profiles/r1_bench_c2_n{2,8}_*): at n = 2 the multicast path moves the same bytes as plain peer
        # loads/stores and its multimem instructions are ~15 % slower (0.557 vs 0.484 ms per step); at n = 8 it moves
        # 59 instead of 104 MB per direction and wins (0.543 vs 0.591 ms).  Default: NVLS at 8 ranks and up -- the sizes
        # it was measured at; n = 4 ran with peer loads/stores only (0.532 ms) and stays there until NVLS is measured too.
        env = os.environ.get("DSSM_NVLINK_MULTICAST")
        want = self.multicast if self.multicast is not None else (env != "0" and (env == "1" or self.world >= 8))
        self.use_multicast = bool(want) and has_mc
        if self.use_multicast:
            self._mc_w = int(self._h_params.multicast_ptr) + 4 * w1_off
            self._mc_dw = int(self._h_comm.multicast_ptr) + 4 * w1_off
        # flag block for the chunked exchange (csrc/nvlink.cu: peer_signal / peer_wait kernels)
        from ._lib import lib

        nb = int(lib.dssm_peer_flags_bytes())
        self._flag_buf = symm_mem.empty(nb // 4, dtype=torch.int32, device=t.device).zero_()
        self._h_flags = symm_mem.rendezvous(self._flag_buf, grp)
        torch.cuda.synchronize(t.device)
        dist.barrier(group=self.group)  # every block is zero before anybody's first flag can land in it
        self._peer_flags = arr(*[int(p) for p in self._h_flags.buffer_ptrs])
        self._xstream = torch.cuda.Stream(device=t.device)
        # chunks of the exchange: the dW1 gather is issued in `x_chunks` column chunks, every chunk's rows are split over
        # the ranks (rank r owns the r-th n-th of EVERY chunk), and chunk k is exchanged on a second stream while chunk
        # k+1 is still being gathered
        # exchange = "push" (default): the gather writes every gradient row into its owner's slot buffer (posted peer
        # stores under the gather itself, absent rows cost nothing), the owner pass reads local memory only;
        # exchange = "pull": local dense dW1, owners pull the rows over NVLink (round 1's scheme, in column chunks)
        self.exchange = os.environ.get("DSSM_DP_EXCHANGE", "push")
        if self.exchange not in ("push", "pull"):
            raise ValueError("DSSM_DP_EXCHANGE must be 'push' or 'pull'")
        self.x_chunks = 1 if self.exchange == "push" else max(1, min(int(os.environ.get("DSSM_DP_CHUNKS", 2)), 32))
        self._owned = [self._owned_rows(k) for k in range(self.x_chunks)]
        if self.exchange == "push":
            import ctypes as C

            from ._lib import check

            D, L1 = t.conf.TRIGRAM_D, t.conf.layers[0]
            self._per = (D + self.world - 1) // self.world
            self._slots = symm_mem.empty(self.world * self._per * L1, dtype=torch.float32, device=t.device)
            self._valid = symm_mem.empty(self.world * self._per, dtype=torch.int32, device=t.device).zero_()
            self._h_slots = symm_mem.rendezvous(self._slots, grp)
            self._h_valid = symm_mem.rendezvous(self._valid, grp)
            torch.cuda.synchronize(t.device)
            dist.barrier(group=self.group)
            self._peer_slots = arr(*[int(p) for p in self._h_slots.buffer_ptrs])
            self._peer_valid = arr(*[int(p) for p in self._h_valid.buffer_ptrs])
            self._epoch_ptr = C.c_void_p(self._flag_buf.data_ptr() + 4 * 16)  # the epoch word behind the DSSM_MAX_PEERS flags
            check(lib.dssm_tower_set_w1_push(t._h, 1))

    def _chunk_cols(self, k: int):
        """Column (= W1 row) range of chunk k of x_chunks -- the same split as dssm_tower_w1_chunk."""
        D, n = self.tower.conf.TRIGRAM_D, self.x_chunks
        per = ((D + n - 1) // n + 3) // 4 * 4
        c0 = min(k * per, D)
        return c0, min(c0 + per, D)

    def _owned_rows(self, k: int, rank: Optional[int] = None):
        """Rows of chunk k whose gradient this rank reduces, whose Adam state it keeps and whose new weights it pushes."""
        r = self.rank if rank is None else rank
        if getattr(self, "exchange", "pull") == "push":  # one contiguous block per rank
            D = self.tower.conf.TRIGRAM_D
            per = (D + self.world - 1) // self.world
            return min(r * per, D), min((r + 1) * per, D)
        c0, c1 = self._chunk_cols(k)
        per = (c1 - c0 + self.world - 1) // self.world
        return min(c0 + r * per, c1), min(c0 + (r + 1) * per, c1)

    # ---- SyncBN: global-batch BatchNorm moments (csrc/nvlink.cu: syncbn_fwd_kernel / syncbn_bwd_kernel) ------------------
    def _setup_syncbn(self) -> None:
        """A small peer-mapped exchange buffer per rank (flags + one slot per BN layer, direction and rank); the tower's
        training-mode BN then uses the moments of the GLOBAL batch (what the single-process reference computes over n*B
        groups, new_dssm.py:77) and the backward averages [dbeta | dgamma] over the replicas: n replicas at B groups each
        reproduce the reference step at n*B (oracle/syncbn.py is the specification)."""
        import ctypes as C

        import torch.distributed._symmetric_memory as symm_mem

        from ._lib import check, lib

        t = self.tower
        if not t.conf.use_bn:
            return
        grp = self.group if self.group is not None else dist.group.WORLD
        nbytes = lib.dssm_tower_syncbn_bytes(t._h, self.world)
        self._sync_buf = symm_mem.empty(nbytes // 4, dtype=torch.float32, device=t.device).zero_()
        self._h_sync = symm_mem.rendezvous(self._sync_buf, grp)
        torch.cuda.synchronize(t.device)
        dist.barrier(group=self.group)  # every rank's buffer is zero before anybody's first flag can land in it
        arr = (C.c_void_p * self.world)(*[int(p) for p in self._h_sync.buffer_ptrs])
        check(lib.dssm_tower_set_syncbn(t._h, self.world, self.rank, arr))
        self.sync_bn = True

    # ---- checkpointing under sharded Adam ---------------------------------------------------------------------
    def state_dict(self):
        """tower.state_dict() made whole: with comm='nvlink' each rank runs Adam only on the W1 rows it owns, so its m / v
        are current for those rows only.  Every rank gathers the owners' row blocks first (all ranks must call this)."""
        t = self.tower
        if self.comm == "nvlink" and self.world > 1:
            L1 = t.conf.layers[0]
            D = t.conf.TRIGRAM_D
            for buf in (t.m, t.v):
                w = buf[:D * L1].view(D, L1)
                for k in range(self.x_chunks):
                    for r in range(self.world):
                        lo, hi = self._owned_rows(k, r)
                        if hi > lo:
                            dist.broadcast(w[lo:hi], src=dist.get_global_rank(self.group, r) if self.group is not None else r,
                                           group=self.group)
        return t.state_dict()

    def save(self, path: str, vocabulary=None):
        """Collective: gathers the sharded optimizer state on every rank, rank 0 writes the file."""
        sd = self.state_dict()
        out = None
        if self.rank == 0:
            out = self.tower.save(path, vocabulary=vocabulary, state=sd)
        if self.world > 1:
            dist.barrier(group=self.group)
        return out

    def _exchange_rows(self, lo: int, hi: int) -> None:
        """Owner pass over the W1 rows [lo, hi): pull / reduce / Adam / push (csrc/nvlink.cu), on the current stream."""
        from ._lib import check, lib, ptr, stream_ptr

        t, c = self.tower, self.tower.conf
        if hi <= lo:
            return
        if self.use_multicast:  # NVLS: in-switch reduction of the gradient rows, in-switch replication of the weight rows
            check(lib.dssm_w1_shard_reduce_adam_mc(self._mc_dw, self._mc_w, ptr(t.params), self.world, c.TRIGRAM_D, c.layers[0],
                                                   lo, hi, ptr(t.m), ptr(t.v), ptr(t.beta_pow), c.learning_rate, c.beta1, c.beta2,
                                                   c.adam_eps, stream_ptr()))
        else:
            check(lib.dssm_w1_shard_reduce_adam(self._peer_dw, self._peer_w, self.world, self.rank, c.TRIGRAM_D, c.layers[0],
                                                lo, hi, ptr(t.m), ptr(t.v), ptr(t.beta_pow), c.learning_rate, c.beta1, c.beta2,
                                                c.adam_eps, stream_ptr()))

    def _step_staged_push(self) -> None:
        """forward + dense backward + CSC (one C call / graph); the dW1 gather PUSHES every finished gradient row into the
        slot buffer of the rank that owns the row (csrc/spmm.cu: PushTarget) -- the gradient's trip over NVLink happens
        under the gather, as posted stores, and only for rows the batch touches; one flag round; the owner pass
        (csrc/nvlink.cu: w1_slots_reduce_adam_kernel) sums its LOCAL slots in rank order, applies Adam and replicates the
        new rows into every replica of W1 (peer stores, or multimem.st through the NVSwitch); one more flag round."""
        from ._lib import check, lib, ptr, stream_ptr

        t, c, n = self.tower, self.tower.conf, self.world
        main = torch.cuda.current_stream(t.device)
        t.fwd_bwd_begin_staged()
        # [grads beyond W1 | EMA shadows]: 0.5 MB, stays on NCCL.  Its all-reduce AND the Adam step on it run on the exchange
        # stream beside the gather / owner pass (they share nothing with W1; the beta powers are only read until
        # adam_advance) -- on the main stream they were ~10 us of serial tail per step.
        self._xstream.wait_stream(main)
        with torch.cuda.stream(self._xstream):
            w_rest = dist.all_reduce(t.comm[self.w1_end:], op=dist.ReduceOp.AVG, group=self.group, async_op=True)
            w_rest.wait()
            t.adam_range(self.w1_end, t.P - self.w1_end, 1.0)
        check(lib.dssm_tower_backward_w1_push(t._h, self._peer_slots, self._peer_valid, self._epoch_ptr, n, self.rank, self._per,
                                              stream_ptr(main)))
        # every rank's rows for my shard have landed (signal + wait in one launch)
        check(lib.dssm_peer_barrier(self._peer_flags, n, self.rank, 0, 2, 0, stream_ptr(main)))
        check(lib.dssm_w1_slots_reduce_adam(ptr(self._slots), ptr(self._valid), self._epoch_ptr, self._peer_w,
                                            self._mc_w if self.use_multicast else None, n, self.rank, c.TRIGRAM_D, c.layers[0], self._per,
                                            ptr(t.m), ptr(t.v), ptr(t.beta_pow), c.learning_rate, c.beta1, c.beta2, c.adam_eps,
                                            stream_ptr(main)))
        main.wait_stream(self._xstream)
        t.adam_advance()
        # every owner's rows have landed in every replica of W1 before anybody's next forward reads it; ends the step (epoch)
        check(lib.dssm_peer_barrier(self._peer_flags, n, self.rank, 1, 2, 1, stream_ptr(main)))

    def _step_staged_nvlink(self) -> None:
        """forward + dense backward + CSC (one C call / graph), then the dW1 gather in x_chunks column chunks on the main
        stream; after chunk k a flag tells every peer that this rank's rows of the chunk are in (symmetric) memory, and the
        owner pass of chunk k -- gated by a kernel that spins until ALL ranks have raised that flag -- runs on the exchange
        stream under the gather of chunk k+1.  No host-side barrier: the two cross-device barriers of the un-chunked
        exchange (27 + 12 us) become flag waits inside the streams, and the step stays one CUDA graph."""
        from ._lib import check, lib, ptr, stream_ptr

        t, K, n = self.tower, self.x_chunks, self.world
        stride = K + 1
        main = torch.cuda.current_stream(t.device)
        own_flags = ptr(self._flag_buf)
        t.fwd_bwd_begin_staged()
        # [grads beyond W1 | EMA shadows]: small, stays on NCCL; queued now so that it runs beside the dW1 gather
        w_rest = dist.all_reduce(t.comm[self.w1_end:], op=dist.ReduceOp.AVG, group=self.group, async_op=True)
        self._xstream.wait_stream(main)  # fork (also what puts the exchange stream inside a graph capture)
        for k in range(K):
            t.backward_w1(k, K)  # this rank's dense dW1 rows of chunk k, in symmetric memory
            check(lib.dssm_peer_signal(self._peer_flags, n, self.rank, k, stride, stream_ptr(main)))
            with torch.cuda.stream(self._xstream):
                check(lib.dssm_peer_wait(own_flags, n, k, stride, stream_ptr(self._xstream)))  # every rank's chunk k is complete
                self._exchange_rows(*self._owned[k])
        w_rest.wait()
        t.adam_range(self.w1_end, t.P - self.w1_end, 1.0)
        main.wait_stream(self._xstream)  # join: this rank's owner passes are done (beta powers were only read)
        t.adam_advance()
        # every owner's rows have landed in every replica of W1 before anybody's next forward reads it
        check(lib.dssm_peer_signal(self._peer_flags, n, self.rank, K, stride, stream_ptr(main)))
        check(lib.dssm_peer_wait(own_flags, n, K, stride, stream_ptr(main)))
        check(lib.dssm_peer_epoch_advance(own_flags, stream_ptr(main)))

    # ---- one step on the staging CSR -------------------------------------------------------------------
    def _step_staged(self) -> None:
        if self.comm == "nvlink" and self.world > 1:
            return self._step_staged_push() if self.exchange == "push" else self._step_staged_nvlink()
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
            # most urgent priority: the main chain's kernel nodes then outrank the tower's side streams (csrc/tower.cu: bind)
            s = torch.cuda.Stream(device=t.device, priority=-1 if os.environ.get("DSSM_SIDE_PRIORITY") != "0" else 0)
            s.wait_stream(torch.cuda.current_stream(t.device))
            with torch.cuda.stream(s):
                for _ in range(warmup):  # communicators, attribute calls, allocator warm-up outside the capture
                    n0 = t.launch_count
                    self._step_staged()
                    self.launches_per_step = t.launch_count - n0  # kernels of OUR library in one step (replays are not counted by the C side)
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
            self.replays += 1
        else:
            self._step_staged()
        return t.tensor("loss")

    # ---- pipelined host feed (tower.train_step_host_async for the data-parallel step) ------------------------------
    def train_step_host_async(self, pinned) -> int:
        """Upload this rank's CSR from pinned host memory on the tower's copy stream (overlapping the previous step), run
        the data-parallel step on it, copy the loss back; returns the step id for tower.feed_loss / feed_wait."""
        t = self.tower
        k = t.feed_upload_async(pinned)
        self.train_step(None)
        t.feed_step_done(k)
        return k

