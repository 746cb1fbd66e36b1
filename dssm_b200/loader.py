"""Host-side batch pipeline for the training loop.

The reference rebuilds its feed for every `sess.run`: three scipy CSR row slices, three `tocoo()` conversions and an
`np.mat` transpose (`utils/utils.py:20-24,45-61`, driven by `new_dssm.py:261-269`), all on the one Python thread that
also drives TensorFlow.  Here a worker thread assembles the stacked CSR `[query ; doc_pos ; doc_neg]` of the coming
batches straight into a ring of pinned host buffers: because `pull_batch` takes *contiguous* row ranges, each of the
three parts is one contiguous run of the epoch matrices' `indices` / `data` arrays, so a batch is three `memcpy`s plus
an offset `indptr` -- no intermediate scipy objects.  The consumer hands the buffers to
`DSSMTower.train_step_host_async` (upload on the copy stream under the previous step's kernels).

Batch count per epoch follows the reference: `int(n_queries / query_BS) - 1` steps (`new_dssm.py:46`).
"""
from __future__ import annotations

import queue
import threading
from typing import Iterator, Optional, Sequence, Tuple

import numpy as np
import scipy.sparse as sp
import torch


def _canonical(x, n_cols: Optional[int] = None) -> sp.csr_matrix:
    m = sp.csr_matrix(x)
    if not m.has_sorted_indices:
        m = m.copy()
        m.sort_indices()
    if n_cols is not None and m.shape[1] != n_cols:
        raise ValueError(f"expected {n_cols} columns, got {m.shape[1]}")
    return m


def reference_epoch_steps(n_queries: int, query_BS: int) -> int:
    """train_epoch_steps of new_dssm.py:46."""
    return max(int(n_queries / query_BS) - 1, 0)


def fill_stacked(parts: Sequence[Tuple[sp.csr_matrix, int, int]], indptr: np.ndarray, indices: np.ndarray,
                 values: np.ndarray) -> int:
    """Write the rows [lo, hi) of every (matrix, lo, hi) in `parts`, stacked in order, as one CSR into the given
    arrays (int32 indptr[R+1], int32 indices[>=nnz], float32 values[>=nnz]); returns nnz."""
    o, r = 0, 0
    indptr[0] = 0
    for m, lo, hi in parts:
        s, e = int(m.indptr[lo]), int(m.indptr[hi])
        n = e - s
        if o + n > indices.shape[0]:
            raise ValueError(f"batch has more than the {indices.shape[0]} non-zeros the buffers were sized for")
        indices[o:o + n] = m.indices[s:e]
        values[o:o + n] = m.data[s:e]  # int64 counts / float64 tf-idf are cast to float32 here, as the feed does
        indptr[r + 1:r + 1 + (hi - lo)] = m.indptr[lo + 1:hi + 1] - (s - o)
        o += n
        r += hi - lo
    return o


class HostBatchLoader:
    """Iterates one epoch of (query_data, doc_data, doc_neg_data) -- the three scipy matrices `pull_batch` slices -- as
    pinned stacked batches `(indptr, indices, values, nnz)` ready for `DSSMTower.train_step_host_async`.

    The worker thread runs up to `depth - 3` batches ahead of the one being consumed, and a buffer handed out at
    iteration i is overwritten while the consumer works on iteration i + 3 at the earliest -- with the double-buffered
    feed (at most two steps in flight, the loss of step k-1 read before step k+1 is issued) step i is complete by then.
    Iterating the loader walks `range(reference_epoch_steps(...))`; `iterate(ids)` takes any order, e.g. the shuffled
    batch ids of `new_dssm.py:262-265`."""

    def __init__(self, query_data, doc_data, doc_neg_data, query_BS: int, NEG: int, max_nnz: Optional[int] = None,
                 depth: int = 5, pin: Optional[bool] = None, native: bool = True, copy_threads: int = 4):
        if depth < 4:
            raise ValueError("depth must be at least 4 (one batch of look-ahead)")
        self.q = _canonical(query_data)
        self.p = _canonical(doc_data, self.q.shape[1])
        self.n = _canonical(doc_neg_data, self.q.shape[1])
        if self.p.shape[0] != self.q.shape[0] or self.n.shape[0] != self.q.shape[0] * NEG:
            raise ValueError("need one positive and NEG negatives per query (utils/utils.py:49-51)")
        self.native, self.copy_threads, self._tab = bool(native), int(copy_threads), None  # assembly by dssm_host_stack_csr (C, GIL-free)
        self.B, self.NEG, self.depth = int(query_BS), int(NEG), int(depth)
        self.R = (2 + self.NEG) * self.B
        self.steps = reference_epoch_steps(self.q.shape[0], self.B)
        self.max_nnz = int(max_nnz) if max_nnz is not None else self.max_batch_nnz()
        pin = torch.cuda.is_available() if pin is None else pin
        mk = (lambda n, dt: torch.empty(n, dtype=dt).pin_memory()) if pin else (lambda n, dt: torch.empty(n, dtype=dt))
        self._bufs = [(mk(self.R + 1, torch.int32), mk(max(self.max_nnz, 1), torch.int32), mk(max(self.max_nnz, 1), torch.float32))
                      for _ in range(self.depth)]

    @classmethod
    def from_texts(cls, vectorizer, queries: Sequence[str], docs: Sequence[str], neg_docs: Sequence[str], query_BS: int, NEG: int,
                   **kw) -> "HostBatchLoader":
        """The reference's path from text to batches (new_dssm.py:33-43): the fitted vectorizer -- a
        dssm_b200.vectorizer.CountVectorizerCompat or sklearn's CountVectorizer -- transforms the three text lists into the
        epoch matrices this loader slices; int64 term counts are cast to float32 when a batch is assembled, as the feed does."""
        return cls(vectorizer.transform(queries), vectorizer.transform(docs), vectorizer.transform(neg_docs), query_BS, NEG, **kw)

    def _parts(self, b: int):
        B, N = self.B, self.NEG
        return ((self.q, b * B, (b + 1) * B), (self.p, b * B, (b + 1) * B), (self.n, b * B * N, (b + 1) * B * N))

    def batch_nnz(self, b: int) -> int:
        return sum(int(m.indptr[hi] - m.indptr[lo]) for m, lo, hi in self._parts(b))

    def max_batch_nnz(self) -> int:
        """Largest stacked batch of the epoch (size DSSMTower's max_nnz with it)."""
        return max((self.batch_nnz(b) for b in range(max(self.steps, 1)) if (b + 1) * self.B <= self.q.shape[0]), default=0)

    _IK = {np.dtype(np.int32): 0, np.dtype(np.int64): 1}
    _VK = {np.dtype(np.float32): 0, np.dtype(np.float64): 1, np.dtype(np.int64): 2, np.dtype(np.int32): 3}

    def _native_tables(self):
        """ctypes tables of the three epoch matrices for dssm_host_stack_csr (built once)."""
        import ctypes as C

        mats = (self.q, self.p, self.n)
        for m in mats:
            if m.indptr.dtype != m.indices.dtype or m.indptr.dtype not in self._IK or m.data.dtype not in self._VK:
                return None
        arr = C.c_void_p * 3
        return dict(ip=arr(*[m.indptr.ctypes.data for m in mats]), ix=arr(*[m.indices.ctypes.data for m in mats]),
                    vl=arr(*[m.data.ctypes.data for m in mats]),
                    ik=(C.c_int32 * 3)(*[self._IK[m.indptr.dtype] for m in mats]),
                    vk=(C.c_int32 * 3)(*[self._VK[m.data.dtype] for m in mats]))

    def _fill(self, slot: int, b: int):
        ip, ix, vl = self._bufs[slot]
        if self.native:
            import ctypes as C

            from ._lib import DssmError, last_error, lib

            if self._tab is None:
                self._tab = self._native_tables() or False
            if self._tab:
                t = self._tab
                parts = self._parts(b)
                lo = (C.c_int64 * 3)(*[p_[1] for p_ in parts])
                hi = (C.c_int64 * 3)(*[p_[2] for p_ in parts])
                # one C call (the GIL is released for its duration): three contiguous copies + the offset indptr
                nnz = lib.dssm_host_stack_csr(3, t["ip"], t["ix"], t["vl"], t["ik"], t["vk"], lo, hi, ip.data_ptr(), ix.data_ptr(),
                                              vl.data_ptr(), ix.shape[0], self.copy_threads)
                if nnz < 0:
                    raise DssmError(int(nnz), last_error())
                return ip, ix, vl, int(nnz)
        nnz = fill_stacked(self._parts(b), ip.numpy(), ix.numpy(), vl.numpy())
        return ip, ix, vl, nnz

    def __len__(self) -> int:
        return self.steps

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, torch.Tensor, torch.Tensor, int]]:
        return self.iterate(range(self.steps))

    def iterate(self, batch_ids: Sequence[int]):
        """Generator over the given batch ids; a worker thread stays up to depth - 3 batches ahead of the consumer."""
        ids = list(batch_ids)
        for b in ids:
            if b < 0 or (b + 1) * self.B > self.q.shape[0]:
                raise IndexError(f"batch {b} is outside the epoch")
        ready: "queue.Queue" = queue.Queue()
        # fills allowed = depth - 1 + (batches consumed - 1): while batch j is being consumed the worker may be as far as
        # batch j + depth - 3, whose slot last held batch j - 3
        free = threading.Semaphore(self.depth - 1)
        stop = threading.Event()

        def work():
            try:
                for i, b in enumerate(ids):
                    while not free.acquire(timeout=0.1):
                        if stop.is_set():
                            return
                    ready.put(self._fill(i % self.depth, b))
                ready.put(None)
            except BaseException as e:  # surfaced in the consumer
                ready.put(e)

        th = threading.Thread(target=work, name="dssm-batch-loader", daemon=True)
        th.start()
        try:
            handed = 0
            while True:
                item = ready.get()
                if item is None:
                    return
                if isinstance(item, BaseException):
                    raise item
                yield item
                handed += 1
                if handed >= 2:
                    free.release()  # the batch handed out two iterations ago is no longer in flight
        finally:
            stop.set()
            th.join(timeout=5)
