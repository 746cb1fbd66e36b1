"""Synthetic batches of SURVEY.md section 8(d): nnz per row = clip(Poisson(lam),1,64); column ids ~ Zipf(1.1)
over [0,D) mapped through a fixed random permutation; duplicates merged into counts; sorted columns."""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from .batch import StackedBatch

_ZIPF_CACHE = {}


def _zipf_cdf(D: int, s: float):
    key = (D, s)
    if key not in _ZIPF_CACHE:
        w = np.arange(1, D + 1, dtype=np.float64) ** (-s)
        cdf = np.cumsum(w)
        cdf /= cdf[-1]
        perm = np.random.Generator(np.random.PCG64(12345)).permutation(D).astype(np.int32)
        _ZIPF_CACHE[key] = (cdf, perm)
    return _ZIPF_CACHE[key]


def sparse_rows(rng: np.random.Generator, n_rows: int, D: int, lam: float, zipf_s: float = 1.1, value_mode: str = "count"):
    cdf, perm = _zipf_cdf(D, zipf_s)
    per_row = np.clip(rng.poisson(lam, size=n_rows), 1, 64)
    total = int(per_row.sum())
    cols = perm[np.minimum(np.searchsorted(cdf, rng.random(total)), D - 1)]
    rows = np.repeat(np.arange(n_rows, dtype=np.int64), per_row)
    m = sp.coo_matrix((np.ones(total, np.float32), (rows, cols)), shape=(n_rows, D)).tocsr()
    m.sum_duplicates()
    m.sort_indices()
    if value_mode == "tfidf":  # dssm_no_bn/dssm_tf_idf.py:37 style real-valued features
        m.data = rng.random(m.nnz).astype(np.float32)
        norms = np.sqrt(np.asarray(m.multiply(m).sum(axis=1)).ravel())
        m = sp.diags((1.0 / np.maximum(norms, 1e-12)).astype(np.float32)) @ m
        m = sp.csr_matrix(m, dtype=np.float32)
        m.sort_indices()
    return m


def make_batch(conf, seed: int = 0, lam_query: float = 18.0, lam_doc: float = 48.0, value_mode: str = "count") -> StackedBatch:
    """One stacked batch [query ; doc_pos ; doc_neg] for conf (TRIGRAM_D, query_BS, NEG)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    B, N, D = conf.query_BS, conf.NEG, conf.TRIGRAM_D
    q = sparse_rows(rng, B, D, lam_query, value_mode=value_mode)
    p = sparse_rows(rng, B, D, lam_doc, value_mode=value_mode)
    n = sparse_rows(rng, B * N, D, lam_doc, value_mode=value_mode)
    X = sp.vstack([q, p, n], format="csr")
    X.sort_indices()
    return StackedBatch(X.indptr.astype(np.int32), X.indices.astype(np.int32), X.data.astype(np.float32), D)


def lambdas_for(conf):
    """C1 (unigram vocabulary) uses lam 12/24, the letter-trigram configs 18/48."""
    return (12.0, 24.0) if conf.TRIGRAM_D <= 30000 else (18.0, 48.0)


def init_params(conf, seed: int = 0):
    """add_layer initialiser (archive/dssm_v3.py:44-53, new_dssm.py:118-120): W and b ~ U(+-sqrt(6/(in+out)));
    BN beta=0, gamma=1 (new_dssm.py:75-76).  Keys: W{l}, b{l}, bn{l}_{q|d}_{beta|gamma}."""
    rng = np.random.Generator(np.random.PCG64(seed))
    p = {}
    for l, (d_in, d_out) in enumerate(conf.layer_dims(), start=1):
        lim = np.sqrt(6.0 / (d_in + d_out))
        p[f"W{l}"] = rng.uniform(-lim, lim, size=(d_in, d_out)).astype(np.float32)
        p[f"b{l}"] = rng.uniform(-lim, lim, size=(d_out,)).astype(np.float32)
    if conf.use_bn:
        for l, (_, d_out) in enumerate(conf.layer_dims(), start=1):
            for seg in ("q", "d"):
                p[f"bn{l}_{seg}_beta"] = np.zeros(d_out, np.float32)
                p[f"bn{l}_{seg}_gamma"] = np.ones(d_out, np.float32)
    return p
