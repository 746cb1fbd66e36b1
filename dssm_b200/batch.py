"""Batch format of the reference (utils/utils.py:20-24,45-61; utils/data_input.py:53-60,121-161) and its
device form: the three CSR slices stacked [query ; doc_pos ; doc_neg] as one canonical int32/fp32 CSR."""
from __future__ import annotations

from collections import namedtuple
from dataclasses import dataclass
from typing import Dict, Optional, Sequence

import numpy as np
import scipy.sparse as sp

# stand-in for tf.SparseTensorValue (utils/utils.py:24): indices [nnz,2] int64 row-major, values, dense_shape
SparseTensorValue = namedtuple("SparseTensorValue", ["indices", "values", "dense_shape"])

QUERY_BATCH, DOC_POS_BATCH, DOC_NEG_BATCH, ON_TRAIN = "query_batch", "doc_positive_batch", "doc_negative_batch", "on_train"


def convert_sparse_matrix_to_sparse_tensor(X) -> SparseTensorValue:
    """utils/utils.py:20-24 (np.mat([row, col]).transpose() -> [nnz,2])."""
    coo = sp.csr_matrix(X).tocoo()
    indices = np.stack([coo.row.astype(np.int64), coo.col.astype(np.int64)], axis=1)
    return SparseTensorValue(indices, coo.data, tuple(coo.shape))


def pull_batch(on_training, query_data, doc_data, doc_neg_data, batch_idx, BS, query_batch=QUERY_BATCH,
               doc_pos_batch=DOC_POS_BATCH, doc_neg_batch=DOC_NEG_BATCH, on_train_batch=ON_TRAIN, conf=None):
    """Same signature and slicing as utils/utils.py:45-61; the placeholder arguments default to their
    graph names (new_dssm.py:111-115) and are used as the feed-dict keys."""
    query_in = query_data[batch_idx * BS:(batch_idx + 1) * BS, :]
    doc_pos_in = doc_data[batch_idx * BS:(batch_idx + 1) * BS, :]
    doc_neg_in = doc_neg_data[batch_idx * BS * conf.NEG:(batch_idx + 1) * BS * conf.NEG, :]
    return {query_batch: convert_sparse_matrix_to_sparse_tensor(query_in),
            doc_pos_batch: convert_sparse_matrix_to_sparse_tensor(doc_pos_in),
            doc_neg_batch: convert_sparse_matrix_to_sparse_tensor(doc_neg_in),
            on_train_batch: on_training}


@dataclass
class StackedBatch:
    """Host CSR of one step: rows [0,B) query, [B,2B) positives, [2B,2B+B*NEG) negatives."""

    indptr: np.ndarray  # int32 [R+1]
    indices: np.ndarray  # int32 [nnz]
    values: np.ndarray  # float32 [nnz]
    n_cols: int

    @property
    def nnz(self) -> int:
        return int(self.indptr[-1])

    @property
    def rows(self) -> int:
        return int(self.indptr.shape[0] - 1)

    def to_scipy(self) -> sp.csr_matrix:
        return sp.csr_matrix((self.values, self.indices, self.indptr), shape=(self.rows, self.n_cols))


def _as_csr(x, n_cols: Optional[int] = None) -> sp.csr_matrix:
    if isinstance(x, SparseTensorValue):
        ind = np.asarray(x.indices)
        m = sp.coo_matrix((np.asarray(x.values, dtype=np.float32), (ind[:, 0], ind[:, 1])), shape=tuple(x.dense_shape))
        return m.tocsr()
    return sp.csr_matrix(x)


def stack_csr(query_in, doc_pos_in, doc_neg_in, query_BS: int, NEG: int) -> StackedBatch:
    """Stack the three feeds.  Raises on anything other than exactly B / B / B*NEG rows -- the reference graph
    bakes query_BS into its slices (new_dssm.py:131,199) and fails the same way on a short batch."""
    mats = [_as_csr(m) for m in (query_in, doc_pos_in, doc_neg_in)]
    want = (query_BS, query_BS, query_BS * NEG)
    for m, w, nm in zip(mats, want, ("query", "doc_positive", "doc_negative")):
        if m.shape[0] != w:
            raise ValueError(f"{nm}_batch has {m.shape[0]} rows, the graph needs exactly {w} (query_BS={query_BS}, NEG={NEG})")
    if len({m.shape[1] for m in mats}) != 1:
        raise ValueError("query/doc feeds disagree on TRIGRAM_D")
    X = sp.vstack(mats, format="csr")
    X.sum_duplicates()
    X.sort_indices()
    return StackedBatch(X.indptr.astype(np.int32), X.indices.astype(np.int32), X.data.astype(np.float32), X.shape[1])


def stack_feed(feed: Dict, conf) -> StackedBatch:
    return stack_csr(feed[QUERY_BATCH], feed[DOC_POS_BATCH], feed[DOC_NEG_BATCH], conf.query_BS, conf.NEG)


# ---- utils/data_input.py vocabulary-file bag-of-words format --------------------------------------
def load_vocab(path: str) -> Dict[str, int]:
    """data/vocab.txt: one token per line, id = line number ([PAD]=0, [UNK]=100)."""
    vocab = {}
    with open(path, encoding="utf8") as f:
        for i, line in enumerate(f):
            vocab[line.rstrip("\n")] = i
    return vocab


def convert_seq2bow(query: Sequence[str], vocab_map: Dict[str, int], nwords: Optional[int] = None, unk: str = "[UNK]"):
    """utils/data_input.py:53-60: count vector over the vocabulary, OOV tokens counted on [UNK]."""
    nwords = nwords or len(vocab_map)
    bow = np.zeros(nwords, dtype=np.float32)
    for w in query:
        bow[vocab_map[w] if w in vocab_map else vocab_map[unk]] += 1
    return bow


def bow_csr(texts: Sequence[Sequence[str]], vocab_map: Dict[str, int], nwords: Optional[int] = None, unk: str = "[UNK]"):
    """csr_matrix float32 [n, nwords] as get_data_by_dssm2 builds it (utils/data_input.py:157-159), without
    going through the dense [n, nwords] intermediate."""
    nwords = nwords or len(vocab_map)
    rows, cols = [], []
    for r, q in enumerate(texts):
        for w in q:
            rows.append(r)
            cols.append(vocab_map[w] if w in vocab_map else vocab_map[unk])
    m = sp.coo_matrix((np.ones(len(rows), np.float32), (rows, cols)), shape=(len(texts), nwords)).tocsr()
    m.sum_duplicates()
    m.sort_indices()
    return m
