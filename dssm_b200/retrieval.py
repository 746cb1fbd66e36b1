"""Corpus cosine top-k (north_star item 5; no reference function exists -- SURVEY.md section 8 row a13).

Scores follow new_dssm.py:185-197 (dot / (||q||*||d||), no epsilon); ordering follows
tf.nn.top_k(sorted=True) (utils/tf_ranking_utils.py:47): score descending, ties to the lower doc id.
Document-sharded: every rank scores its contiguous id range, local top-k lists are all-gathered and merged.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist

from ._lib import check, lib, ptr, stream_ptr


LAST_CALL = {"method": None, "fallback": False}  # what the most recent corpus_topk did (tests / bench read it)
TC_MIN_DOCS = 32768  # below this the exact kernels are as fast as the filter + rescore pipeline


class CorpusIndex:
    """Per-corpus index of the bf16 filter (csrc/topk_bf16.cu), built ONCE and reused by every query batch: the fp32 row
    norms plus the rows normalised, rounded to bf16 and stored as ready-made SWIZZLE_128B tensor-core tiles (256 B per
    document).  The fp32 rows stay the source of truth: survivors of the filter are re-scored from them exactly."""

    def __init__(self, docs: torch.Tensor):
        if not docs.is_cuda or docs.dtype != torch.float32 or docs.dim() != 2 or docs.shape[1] != 128:
            raise ValueError("CorpusIndex takes a CUDA float32 [n_docs, 128] matrix")
        self.docs = docs.contiguous()
        self.nd = int(docs.shape[0])
        nb = int(lib.dssm_corpus_index_bytes(self.nd, 128))
        buf = torch.empty(nb + 1024, dtype=torch.uint8, device=docs.device)
        off = (-buf.data_ptr()) % 1024
        self._buf = buf
        self.data = buf[off:off + nb]
        check(lib.dssm_corpus_index_build(ptr(self.docs), self.nd, 128, ptr(self.data), nb, stream_ptr()))

    @property
    def nbytes(self) -> int:
        return int(self.data.numel())


def corpus_topk(Q: torch.Tensor, docs, k: int, id_offset: int = 0, method: str = "auto") -> Tuple[torch.Tensor, torch.Tensor]:
    """(scores [nq,k] fp32, ids [nq,k] int32), sorted; ids = id_offset + local row.
    docs: the fp32 corpus matrix, or a CorpusIndex built over it (then the bf16 filter is used).
    method: "exact" (SIMT, the oracle's arithmetic for every pair), "tc" (tcgen05 tf32 filter straight from the fp32 rows +
    exact rescoring of the survivors), "bf16" (tcgen05 bf16 filter over a CorpusIndex + the same exact rescoring); all three
    return the same ids and scores bit for bit.  "auto": bf16 when given an index, else tc when it applies and the corpus is
    large, else exact."""
    index = docs if isinstance(docs, CorpusIndex) else None
    if index is not None:
        docs = index.docs
    if not (Q.is_cuda and docs.is_cuda):
        raise ValueError("corpus_topk takes CUDA tensors (there is no CPU path)")
    nq, d = Q.shape
    nd = docs.shape[0]
    k = min(k, nd)
    if method == "auto":
        method = "bf16" if index is not None else ("tc" if (d == 128 and nd >= TC_MIN_DOCS) else "exact")
    if method == "bf16":
        if index is None:
            index = CorpusIndex(docs)  # callers that query repeatedly should build it once and pass it in
        nb = lib.dssm_corpus_topk_indexed_workspace_bytes(nq, k)
        ws = torch.empty(nb, dtype=torch.uint8, device=Q.device)
        s = torch.empty((nq, k), dtype=torch.float32, device=Q.device)
        i = torch.empty((nq, k), dtype=torch.int32, device=Q.device)
        flag = torch.zeros(1, dtype=torch.int32, device=Q.device)
        check(lib.dssm_corpus_topk_indexed(ptr(Q), nq, ptr(docs), ptr(index.data), nd, d, k, id_offset, ptr(s), ptr(i), ptr(flag), ptr(ws), nb,
                                           stream_ptr()))
        # overflow of a candidate list, or a query with some but fewer than k docs of finite cosine (the filters do not
        # collect the -inf fillers behind them; a zero-norm query, all -inf, is handled by the kernels)
        fell_back = int(flag.item()) != 0 or bool((torch.isinf(s[:, k - 1]) & ~torch.isinf(s[:, 0])).any())
        LAST_CALL.update(method="bf16", fallback=fell_back)
        if not fell_back:
            return s, i
        method = "exact-after-bf16"  # LAST_CALL keeps {"bf16", fallback=True}; the exact kernels below produce the result
    if method == "tc":
        nb = lib.dssm_corpus_topk_tc_workspace_bytes(nq, nd, d, k)
        ws = torch.empty(nb, dtype=torch.uint8, device=Q.device)
        s = torch.empty((nq, k), dtype=torch.float32, device=Q.device)
        i = torch.empty((nq, k), dtype=torch.int32, device=Q.device)
        flag = torch.zeros(1, dtype=torch.int32, device=Q.device)
        check(lib.dssm_corpus_topk_tc(ptr(Q), nq, ptr(docs), nd, d, k, id_offset, ptr(s), ptr(i), ptr(flag), ptr(ws), nb, stream_ptr()))
        fell_back = int(flag.item()) != 0 or bool((torch.isinf(s[:, k - 1]) & ~torch.isinf(s[:, 0])).any())
        LAST_CALL.update(method="tc", fallback=fell_back)
        if not fell_back:
            return s, i
        # a candidate list overflowed (thresholds too loose for this data): the exact kernels always work
    elif method == "exact":
        LAST_CALL.update(method="exact", fallback=False)
    nb = lib.dssm_corpus_topk_workspace_bytes(nq, nd, d, k)
    ws = torch.empty(nb, dtype=torch.uint8, device=Q.device)
    s = torch.empty((nq, k), dtype=torch.float32, device=Q.device)
    i = torch.empty((nq, k), dtype=torch.int32, device=Q.device)
    check(lib.dssm_corpus_topk(ptr(Q), nq, ptr(docs), nd, d, k, id_offset, ptr(s), ptr(i), ptr(ws), nb, stream_ptr()))
    return s, i


def topk_merge(part_scores: torch.Tensor, part_ids: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """part_* [n_parts, nq, k] (each part sorted) -> global (scores, ids) [nq, k]."""
    n_parts, nq, k = part_scores.shape
    s = torch.empty((nq, k), dtype=torch.float32, device=part_scores.device)
    i = torch.empty((nq, k), dtype=torch.int32, device=part_scores.device)
    check(lib.dssm_topk_merge(ptr(part_scores.contiguous()), ptr(part_ids.contiguous()), n_parts, nq, k, ptr(s), ptr(i), stream_ptr()))
    return s, i


def shard_range(n_docs: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous id range [lo, hi) owned by `rank`."""
    per = (n_docs + world - 1) // world
    lo = min(n_docs, rank * per)
    return lo, min(n_docs, lo + per)


PAD_ID = 0x7FFFFFFF  # id of a padding entry in a local list: loses every (score desc, id asc) comparison against a real doc


def sharded_corpus_topk(Q: torch.Tensor, local_docs: torch.Tensor, k: int, id_offset: int, group=None,
                        total_docs: Optional[int] = None, method: str = "auto"):
    """Every rank: local top-k over its shard, NCCL all-gather of the (scores, ids) lists, merge on every rank.
    The local stage ALWAYS yields [nq, k]: a shard with fewer than k documents (shard_range gives the last rank a short or
    even empty one) pads its lists with (-inf, PAD_ID), which the merge ranks behind every real document, so the gathered
    shapes agree on all ranks; the result is cut to min(k, total_docs) columns (total_docs: corpus size over all shards;
    None = all-reduced here)."""
    nq = Q.shape[0]
    nd_local = int(local_docs.nd if isinstance(local_docs, CorpusIndex) else local_docs.shape[0])
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return corpus_topk(Q, local_docs, k, id_offset, method=method)
    s = torch.full((nq, k), float("-inf"), dtype=torch.float32, device=Q.device)
    i = torch.full((nq, k), PAD_ID, dtype=torch.int32, device=Q.device)
    if nd_local > 0:
        ls, li = corpus_topk(Q, local_docs, min(k, nd_local), id_offset, method=method)  # local_docs may be a CorpusIndex
        s[:, :ls.shape[1]] = ls
        i[:, :li.shape[1]] = li
    if total_docs is None:
        t = torch.tensor([nd_local], dtype=torch.int64, device=Q.device)
        dist.all_reduce(t, group=group)
        total_docs = int(t.item())
    gs = torch.empty((world, nq, k), dtype=s.dtype, device=s.device)
    gi = torch.empty((world, nq, k), dtype=i.dtype, device=i.device)
    dist.all_gather_into_tensor(gs, s, group=group)
    dist.all_gather_into_tensor(gi, i, group=group)
    ms, mi = topk_merge(gs, gi)
    kk = min(k, total_docs)
    return (ms[:, :kk].contiguous(), mi[:, :kk].contiguous()) if kk < k else (ms, mi)
