"""Shared test helpers: small random batches, oracle construction, comparison with a scale-aware tolerance."""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from oracle import DSSMOracle, OracleConfig, init_params


def oracle_config(conf) -> OracleConfig:
    return OracleConfig(TRIGRAM_D=conf.TRIGRAM_D, layers=tuple(conf.layers), NEG=conf.NEG, query_BS=conf.query_BS,
                        learning_rate=conf.learning_rate, use_bn=conf.use_bn, act=conf.act, bn_eps=conf.bn_eps,
                        ema_decay=conf.ema_decay, gamma=conf.gamma, loss_eps=conf.loss_eps, loss_div_bs=conf.loss_div_bs,
                        beta1=conf.beta1, beta2=conf.beta2, adam_eps=conf.adam_eps)


def random_csr(rng, rows, D, max_nnz_row=6, value_mode="count", allow_empty=True):
    lo = 0 if allow_empty else 1
    data, idx, ptr = [], [], [0]
    for _ in range(rows):
        k = int(rng.integers(lo, max_nnz_row + 1))
        cols = np.sort(rng.choice(D, size=min(k, D), replace=False))
        idx.extend(cols.tolist())
        if value_mode == "count":
            data.extend(rng.integers(1, 4, size=len(cols)).astype(np.float32).tolist())
        else:
            data.extend(rng.random(len(cols)).astype(np.float32).tolist())
        ptr.append(len(idx))
    return sp.csr_matrix((np.asarray(data, np.float32), np.asarray(idx, np.int32), np.asarray(ptr, np.int32)), shape=(rows, D))


def rel_err(a, b):
    """max |a-b| / max(|b|) -- error relative to the tensor's scale (the 1e-5 bar of north_star)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    if a.size == 0:
        return 0.0
    nan_a, nan_b = np.isnan(a), np.isnan(b)
    assert np.array_equal(nan_a, nan_b), "NaN pattern differs"
    scale = max(np.nanmax(np.abs(b)) if np.any(~nan_b) else 0.0, 1e-30)
    d = np.abs(a - b)
    return float(np.nanmax(d) / scale) if np.any(~nan_b) else 0.0


def assert_close(a, b, tol, what=""):
    e = rel_err(a, b)
    assert e <= tol, f"{what}: relative-to-scale error {e:.3e} > {tol:.1e}"


# ---- golden fixtures -----------------------------------------------------------------------------------
import os

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_CASES = ["tiny_bn_relu", "tiny_nobn_eps", "tiny_tanh_3layer_sum", "odd_shapes_neg1"]


def load_golden(name):
    """Returns (Config, [csr per step], params0 dict, raw npz)."""
    from dssm_b200 import Config

    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    conf = Config(TRIGRAM_D=int(z["cfg_TRIGRAM_D"]), query_BS=int(z["cfg_query_BS"]), NEG=int(z["cfg_NEG"]),
                  layers=tuple(int(x) for x in z["cfg_layers"]), use_bn=bool(z["cfg_use_bn"]), act=str(z["cfg_act"]),
                  loss_eps=float(z["cfg_loss_eps"]), loss_div_bs=bool(z["cfg_loss_div_bs"]),
                  learning_rate=float(z["cfg_learning_rate"]))
    Xs = []
    for s in range(int(z["steps"])):
        Xs.append(sp.csr_matrix((z[f"x{s}/values"], z[f"x{s}/indices"], z[f"x{s}/indptr"]), shape=(conf.rows, conf.TRIGRAM_D)))
    params = {k[len("param0/"):]: z[k] for k in z.files if k.startswith("param0/")}
    return conf, Xs, params, z


def to_stacked(X):
    from dssm_b200.batch import StackedBatch

    X = sp.csr_matrix(X)
    return StackedBatch(X.indptr.astype(np.int32), X.indices.astype(np.int32), X.data.astype(np.float32), X.shape[1])
