"""Callers on either side of the hot path, host-side Python like the reference:

  * StreamingAUC          tf.metrics.auc(labels, cos_sim_raw, num_thresholds=2000) as new_dssm.py:226-231 uses it --
                          including the reference's behaviour of never resetting the accumulators (:252,281-286)
  * accuracy              new_dssm.py:220-221
  * write_mid_vectors     the "text \\t idx:val,idx:val" dump of load_model_and_save_vector.py:108-152
  * embed_docs / embed_queries   eval-mode (EMA statistics) embeddings of arbitrary rows through the tower, i.e. the
                          producer of the doc-embedding matrix that corpus_topk consumes
"""
from __future__ import annotations

from typing import Iterable, List, Sequence

import numpy as np
import scipy.sparse as sp

from .batch import StackedBatch


class StreamingAUC:
    """tf.metrics.auc, TF-1.x defaults: curve='ROC', summation_method='trapezoidal', thresholds
    [-1e-7, 1/(n-1), ..., (n-2)/(n-1), 1+1e-7]; a prediction is positive at threshold t iff prediction > t; the four
    confusion counters accumulate over every update() (the reference never re-initialises its local variables)."""

    def __init__(self, num_thresholds: int = 2000):
        eps = 1e-7
        inner = [(i + 1) * 1.0 / (num_thresholds - 1) for i in range(num_thresholds - 2)]
        self.thresholds = np.asarray([0.0 - eps] + inner + [1.0 + eps], dtype=np.float32)
        z = lambda: np.zeros(num_thresholds, dtype=np.float64)
        self.tp, self.fn, self.tn, self.fp = z(), z(), z(), z()

    def update(self, labels: np.ndarray, predictions: np.ndarray) -> float:
        labels = np.asarray(labels).reshape(-1).astype(bool)
        pred = np.asarray(predictions, dtype=np.float32).reshape(-1)
        pos = np.sort(pred[labels])
        neg = np.sort(pred[~labels])
        # count of predictions > t for every threshold (NaN predictions compare false, as in TF)
        pos, neg = pos[~np.isnan(pos)], neg[~np.isnan(neg)]
        n_pos_nan, n_neg_nan = int(labels.sum()) - pos.size, int((~labels).sum()) - neg.size
        tp = pos.size - np.searchsorted(pos, self.thresholds, side="right")
        fp = neg.size - np.searchsorted(neg, self.thresholds, side="right")
        self.tp += tp
        self.fn += pos.size - tp + n_pos_nan
        self.fp += fp
        self.tn += neg.size - fp + n_neg_nan
        return self.result()

    def result(self) -> float:
        eps = 1e-6
        tpr = (self.tp + eps) / (self.tp + self.fn + eps)
        fpr = self.fp / (self.fp + self.tn + eps)
        return float(np.sum((fpr[:-1] - fpr[1:]) * (tpr[:-1] + tpr[1:]) / 2.0))


def labels_for(query_BS: int, NEG: int) -> np.ndarray:
    """label = [1]*query_BS + [0]*query_BS*NEG  (new_dssm.py:163-165), aligned with cos_sim_raw's reference order."""
    return np.asarray([1] * query_BS + [0] * query_BS * NEG, dtype=np.int32)


def accuracy(prob: np.ndarray) -> float:
    """mean(argmax(prob, 1) == 0)  (new_dssm.py:220-221)."""
    return float(np.mean(np.argmax(np.asarray(prob), axis=1) == 0))


def format_mid_vector(vec: Sequence[float]) -> str:
    """One embedding as 'idx:val,idx:val': entries with str(v) != '0.0' and v > 1e-4, value cut to 6 characters
    (load_model_and_save_vector.py:111-118)."""
    s = []
    for index, j in enumerate(np.asarray(vec).tolist()):
        j_s = str(j)
        if j_s != "0.0" and j > 0.0001:
            s.append(str(index) + ":" + j_s[0:6])
    return ",".join(s)


def write_mid_vectors(path: str, texts: Sequence[str], Y: np.ndarray) -> None:
    """Appends 'text-without-spaces \\t idx:val,...' lines (files are opened 'a+' like the reference, :109,124,138)."""
    with open(path, "a+") as f:
        for text, row in zip(texts, np.asarray(Y)):
            f.write(text.replace(" ", "") + "\t" + format_mid_vector(row) + "\n")


# ---- embeddings of arbitrary rows through the tower (eval mode) --------------------------------------------------
def _pack(rows: sp.csr_matrix, n_slots: int) -> sp.csr_matrix:
    """rows padded (by repeating row 0) to exactly n_slots rows."""
    if rows.shape[0] == n_slots:
        return rows
    pad = sp.vstack([rows[0]] * (n_slots - rows.shape[0]), format="csr") if rows.shape[0] < n_slots else None
    return sp.vstack([rows, pad], format="csr")


def _embed(tower, X: sp.csr_matrix, segment: str) -> np.ndarray:
    import torch

    conf = tower.conf
    B, NEG = conf.query_BS, conf.NEG
    X = sp.csr_matrix(X, dtype=np.float32)
    per = B if segment == "q" else (1 + NEG) * B
    out = np.empty((X.shape[0], conf.layers[-1]), dtype=np.float32)
    filler = X[:1]
    for lo in range(0, X.shape[0], per):
        chunk = _pack(X[lo:lo + per], per)
        if segment == "q":
            q, docs = chunk, _pack(filler, (1 + NEG) * B)
        else:
            q, docs = _pack(filler, B), chunk
        stacked = sp.vstack([q, docs], format="csr")
        stacked.sort_indices()
        sb = StackedBatch(stacked.indptr.astype(np.int32), stacked.indices.astype(np.int32), stacked.data.astype(np.float32), X.shape[1])
        if sb.nnz > tower.max_nnz:
            raise ValueError(f"embedding batch has {sb.nnz} non-zeros, the tower was sized for {tower.max_nnz}")
        tower.forward(tower.to_device(sb), on_train=False)
        Y = tower.tensor("Y")
        part = Y[:B] if segment == "q" else Y[B:]
        n = min(per, X.shape[0] - lo)
        out[lo:lo + n] = part[:n].cpu().numpy()
    return out


def embed_queries(tower, X: sp.csr_matrix) -> np.ndarray:
    """embedding_query_y of arbitrary rows with on_train=False (query BN instance, EMA statistics)."""
    return _embed(tower, X, "q")


def embed_docs(tower, X: sp.csr_matrix) -> np.ndarray:
    """embedding_doc_*_y of arbitrary rows with on_train=False (doc BN instance, EMA statistics): the corpus matrix
    for corpus_topk.  In eval mode every row is embedded independently, so packing rows into the positive/negative
    slots of fixed-shape batches does not change their values."""
    return _embed(tower, X, "d")
