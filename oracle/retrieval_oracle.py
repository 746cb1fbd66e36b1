"""Corpus cosine top-k oracle (TEST INFRASTRUCTURE ONLY).

The reference has NO retrieval path; the nearest precedents are the
cosine of new_dssm.py:185-197 (dot / (||q||*||d||), no epsilon) and the ordering contract of
``tf.nn.top_k(sorted=True)`` used at utils/tf_ranking_utils.py:47 (descending score, ties to the
lower index).  This file *defines* the arithmetic both sides must share so that ids are bit-exact:

  dot_seq(a,b)  = (((a0*b0) + a1*b1) + a2*b2) + ...   every * and + individually rounded to fp32
                  (no FMA contraction), t = 0..d-1 in order
  score(q,d)    = dot_seq(q,d) / (sqrt(dot_seq(q,q)) * sqrt(dot_seq(d,d)))    IEEE fp32 sqrt/mul/div
  key           = score, NaN (zero-norm row, 0/0) ranks as -inf, -0.0 ranks equal to +0.0
  order         = key descending, then doc id ascending
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def _dot_seq(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """a [m,d] fp32, b [n,d] fp32 -> [m,n] with strictly sequential fp32 multiply-then-add."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    acc = np.zeros((a.shape[0], b.shape[0]), dtype=np.float32)
    for t in range(a.shape[1]):
        acc = acc + a[:, t : t + 1] * b[None, :, t]
    return acc


def _sqnorm_seq(a: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float32)
    acc = np.zeros(a.shape[0], dtype=np.float32)
    for t in range(a.shape[1]):
        acc = acc + a[:, t] * a[:, t]
    return acc


def exact_cosine_scores(Q: np.ndarray, docs: np.ndarray, chunk: int = 65536) -> np.ndarray:
    qn = np.sqrt(_sqnorm_seq(Q))
    out = np.empty((Q.shape[0], docs.shape[0]), dtype=np.float32)
    for s in range(0, docs.shape[0], chunk):
        d = docs[s : s + chunk]
        dn = np.sqrt(_sqnorm_seq(d))
        with np.errstate(invalid="ignore", divide="ignore"):
            out[:, s : s + chunk] = _dot_seq(Q, d) / (qn[:, None] * dn[None, :])
    return out


def _rank_key(scores: np.ndarray) -> np.ndarray:
    key = np.where(np.isnan(scores), -np.inf, scores).astype(np.float32)
    return key + np.float32(0.0)  # -0.0 -> +0.0


def _topk_rows(key: np.ndarray, ids: np.ndarray, k: int) -> Tuple[np.ndarray, np.ndarray]:
    nq = key.shape[0]
    k = min(k, key.shape[1])
    out_s = np.empty((nq, k), np.float32)
    out_i = np.empty((nq, k), np.int32)
    for r in range(nq):
        order = np.lexsort((ids[r], -key[r]))[:k]
        out_s[r] = key[r, order]
        out_i[r] = ids[r, order]
    return out_s, out_i


def corpus_topk_oracle(Q: np.ndarray, docs: np.ndarray, k: int, id_offset: int = 0):
    """Returns (scores [nq,k] fp32, ids [nq,k] int32), ids are global (id_offset + local row)."""
    scores = exact_cosine_scores(Q, docs)
    key = _rank_key(scores)
    ids = np.broadcast_to(np.arange(docs.shape[0], dtype=np.int64)[None, :] + id_offset, key.shape)
    return _topk_rows(key, ids, k)


def merge_topk_oracle(parts: Sequence[Tuple[np.ndarray, np.ndarray]], k: int):
    """Merge per-shard (scores, ids) lists (each already a local top-k) into the global top-k."""
    s = np.concatenate([p[0] for p in parts], axis=1)
    i = np.concatenate([p[1] for p in parts], axis=1).astype(np.int64)
    return _topk_rows(_rank_key(s), i, k)
