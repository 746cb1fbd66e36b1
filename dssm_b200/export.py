"""Callers on either side of the hot path, host-side Python like the reference:

  * DeviceStreamingAUC    the same metric accumulated on the GPU (csrc/metrics.cu), fed straight from the forward's cos_sim_raw
  * StreamingAUC          host restatement (checker of the device kernel): tf.metrics.auc(labels, cos_sim_raw, num_thresholds=2000) as new_dssm.py:226-231 uses it --
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
        self.thresholds = auc_thresholds(num_thresholds)
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


def auc_thresholds(num_thresholds: int = 2000) -> np.ndarray:
    """tf.metrics.auc's threshold table (TF 1.x): [-1e-7, 1/(n-1), ..., (n-2)/(n-1), 1+1e-7] as float32."""
    eps = 1e-7
    inner = [(i + 1) * 1.0 / (num_thresholds - 1) for i in range(num_thresholds - 2)]
    return np.asarray([0.0 - eps] + inner + [1.0 + eps], dtype=np.float32)


class DeviceStreamingAUC:
    """StreamingAUC on the device (csrc/metrics.cu): the confusion counters live in HBM as two never-reset 64-bit
    histograms, update() is one kernel over cos_sim_raw as the forward left it (no device-to-host copy of the
    predictions), result() one more.  Same thresholds, same `prediction > t` rule, same never-reset behaviour
    (new_dssm.py:226-231,252) -- tests/test_gpu_metrics.py checks it against StreamingAUC."""

    def __init__(self, device, num_thresholds: int = 2000):
        import torch

        self.T = int(num_thresholds)
        self.device = torch.device(device)
        self.thresholds = torch.from_numpy(auc_thresholds(self.T)).to(self.device)
        self.pos_hist = torch.zeros(self.T + 1, dtype=torch.int64, device=self.device)
        self.neg_hist = torch.zeros(self.T + 1, dtype=torch.int64, device=self.device)
        self._out = torch.zeros(3, dtype=torch.float64, device=self.device)

    def update(self, predictions, n_pos: int):
        """predictions: CUDA fp32 tensor whose first n_pos entries carry label 1 and the rest label 0 (cos_sim_raw's
        order: label = [1]*query_BS + [0]*query_BS*NEG, new_dssm.py:163-165).  Returns the running AUC (device scalar)."""
        from ._lib import check, lib, ptr, stream_ptr

        p = predictions.reshape(-1)
        if not p.is_cuda or p.dtype != self.thresholds.dtype:
            raise ValueError("DeviceStreamingAUC.update takes a CUDA float32 tensor")
        check(lib.dssm_auc_update(ptr(p), int(n_pos), p.numel(), ptr(self.thresholds), self.T, ptr(self.pos_hist), ptr(self.neg_hist),
                                  stream_ptr()))
        return self.result()

    def result(self):
        from ._lib import check, lib, ptr, stream_ptr

        check(lib.dssm_auc_result(ptr(self.pos_hist), ptr(self.neg_hist), self.T, ptr(self._out), stream_ptr()))
        return self._out[0]


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
def _embed(tower, X: sp.csr_matrix, segment: str, return_device: bool = False):
    """Rows of X through the eval-mode forward, per-batch work kept off the host: a batch is one contiguous run of X's
    index / value arrays copied into one of two pinned staging buffers (the slots of the other BN instance, and the tail of
    a short last batch, stay EMPTY rows -- in eval mode every row is embedded independently), uploaded asynchronously, and
    its embeddings are copied device-to-device into the output matrix.  The host waits only when it is about to overwrite
    a staging buffer that an upload two batches back may still be reading, and once at the end."""
    import torch

    from .ops import DeviceCSR

    conf = tower.conf
    B, NEG, R = conf.query_BS, conf.NEG, conf.rows
    X = sp.csr_matrix(X, dtype=np.float32)
    if not X.has_sorted_indices:
        X.sort_indices()
    if X.shape[1] != conf.TRIGRAM_D:
        raise ValueError("embedding input TRIGRAM_D mismatch")
    per = B if segment == "q" else (1 + NEG) * B
    row0 = 0 if segment == "q" else B
    n = X.shape[0]
    dev = tower.device
    out = torch.empty((n, conf.layers[-1]), dtype=torch.float32, device=dev)
    cap = max(tower.max_nnz, 1)
    bufs = []
    for _ in range(2):
        bufs.append(dict(ip=torch.zeros(R + 1, dtype=torch.int32).pin_memory(), ix=torch.zeros(cap, dtype=torch.int32).pin_memory(),
                         vl=torch.zeros(cap, dtype=torch.float32).pin_memory(),
                         dip=torch.zeros(R + 1, dtype=torch.int32, device=dev), dix=torch.zeros(cap, dtype=torch.int32, device=dev),
                         dvl=torch.zeros(cap, dtype=torch.float32, device=dev), ev=None))
    xip = X.indptr
    for ci, lo in enumerate(range(0, n, per)):
        hi = min(n, lo + per)
        b = bufs[ci & 1]
        if b["ev"] is not None:
            b["ev"].synchronize()  # the upload that last read this staging buffer is done
        s, e = int(xip[lo]), int(xip[hi])
        nnz = e - s
        if nnz > tower.max_nnz:
            raise ValueError(f"embedding batch has {nnz} non-zeros, the tower was sized for {tower.max_nnz}")
        ip = b["ip"].numpy()
        ip[:row0 + 1] = 0
        ip[row0 + 1:row0 + 1 + (hi - lo)] = xip[lo + 1:hi + 1] - s
        ip[row0 + 1 + (hi - lo):] = nnz
        b["ix"].numpy()[:nnz] = X.indices[s:e]
        b["vl"].numpy()[:nnz] = X.data[s:e]
        b["dip"].copy_(b["ip"], non_blocking=True)
        b["dix"][:nnz].copy_(b["ix"][:nnz], non_blocking=True)
        b["dvl"][:nnz].copy_(b["vl"][:nnz], non_blocking=True)
        b["ev"] = torch.cuda.Event()
        b["ev"].record(torch.cuda.current_stream(dev))
        tower.forward(DeviceCSR(b["dip"], b["dix"], b["dvl"], R, conf.TRIGRAM_D, nnz), on_train=False)
        out[lo:hi].copy_(tower.tensor("Y")[row0:row0 + (hi - lo)])
    if return_device:
        return out
    return out.cpu().numpy()


def embed_queries(tower, X: sp.csr_matrix, return_device: bool = False):
    """embedding_query_y of arbitrary rows with on_train=False (query BN instance, EMA statistics)."""
    return _embed(tower, X, "q", return_device)


def embed_docs(tower, X: sp.csr_matrix, return_device: bool = False):
    """embedding_doc_*_y of arbitrary rows with on_train=False (doc BN instance, EMA statistics): the corpus matrix
    for corpus_topk (return_device=True hands it over without leaving HBM).  In eval mode every row is embedded
    independently, so packing rows into the positive/negative slots of fixed-shape batches does not change their values."""
    return _embed(tower, X, "d", return_device)
