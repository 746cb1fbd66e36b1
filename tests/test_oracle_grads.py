"""The oracle's hand-derived backward, Adam and EMA against an independent torch float64 autograd model of the
same graph (the second opinion SURVEY.md section 8c asks for, since TensorFlow cannot run here)."""
import numpy as np
import pytest
import torch

from oracle import DSSMOracle, DPOracle, OracleConfig, init_params
from tests.helpers import random_csr


def torch_model_loss(cfg, params, X_dense, on_train=True):
    """Independent restatement with torch ops (float64), returns loss and the param tensors."""
    P = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in params.items()}
    B, N = cfg.query_BS, cfg.NEG
    a = torch.tensor(X_dense, dtype=torch.float64)
    for l in range(1, len(cfg.layers) + 1):
        h = a @ P[f"W{l}"] + P[f"b{l}"]
        if cfg.use_bn:
            outs = []
            for seg, sl in (("q", slice(0, B)), ("d", slice(B, None))):
                x = h[sl]
                mean = x.mean(0)
                var = ((x - mean) ** 2).mean(0)
                outs.append((x - mean) / torch.sqrt(var + cfg.bn_eps) * P[f"bn{l}_{seg}_gamma"] + P[f"bn{l}_{seg}_beta"])
            h = torch.cat(outs, 0)
        a = torch.relu(h) if cfg.act == "relu" else torch.tanh(h)
    q, pos, neg = a[:B], a[B:2 * B], a[2 * B:]
    docs = torch.cat([pos[:, None, :], neg.reshape(B, N, -1)], 1)
    cos = (q[:, None, :] * docs).sum(-1) / (q.norm(dim=1)[:, None] * docs.norm(dim=2))
    prob = torch.softmax(cos * cfg.gamma, dim=1)
    loss = -torch.log(prob[:, 0] + cfg.loss_eps).sum()
    if cfg.loss_div_bs:
        loss = loss / B
    return loss, P


VARIANTS = [
    dict(),
    dict(use_bn=False, loss_eps=1e-8),
    dict(act="tanh"),
    dict(loss_div_bs=False),
    dict(layers=(7, 6, 5)),
]


@pytest.mark.parametrize("kw", VARIANTS)
def test_backward_matches_autograd(kw):
    rng = np.random.default_rng(0)
    base = dict(TRIGRAM_D=40, layers=(8, 6), NEG=3, query_BS=5)
    base.update(kw)
    cfg = OracleConfig(**base)
    params = init_params(cfg, 1)
    for k in params:  # move BN params off their trivial init so their grads are exercised
        if "gamma" in k:
            params[k] = (1 + 0.3 * rng.standard_normal(params[k].shape)).astype(np.float32)
        if "beta" in k:
            params[k] = (0.2 * rng.standard_normal(params[k].shape)).astype(np.float32)
    X = random_csr(rng, cfg.rows, cfg.TRIGRAM_D, allow_empty=False)
    orc = DSSMOracle(cfg, params, dtype=np.float64)
    cache = orc.forward(X, on_train=True)
    grads = orc.backward(cache)
    loss, P = torch_model_loss(cfg, params, X.toarray())
    loss.backward()
    assert np.isclose(float(cache["loss"]), loss.item(), rtol=1e-10)
    for k, g in grads.items():
        ref = P[k].grad.numpy()
        scale = max(np.abs(ref).max(), 1e-12)
        assert np.abs(g - ref).max() / scale < 1e-7 or np.abs(g - ref).max() < 1e-12, k


def test_fp32_oracle_close_to_fp64():
    rng = np.random.default_rng(3)
    cfg = OracleConfig(TRIGRAM_D=64, layers=(12, 16), NEG=3, query_BS=5)  # wide enough that no relu row is all-zero
    params = init_params(cfg, 0)
    X = random_csr(rng, cfg.rows, cfg.TRIGRAM_D, allow_empty=False)
    a = DSSMOracle(cfg, params, np.float32)
    b = DSSMOracle(cfg, params, np.float64)
    ca, cb = a.forward(X, True), b.forward(X, True)
    assert np.isfinite(float(cb["loss"]))
    assert abs(float(ca["loss"]) - float(cb["loss"])) < 1e-5 * abs(float(cb["loss"]))
    ga, gb = a.backward(ca), b.backward(cb)
    for k in ga:
        if k[0] == "b" and k[1:].isdigit():  # analytically zero under BN (the batch mean is removed): fp32 value is rounding noise
            assert np.abs(ga[k] - gb[k]).max() <= 1e-5 * np.abs(ca["dh" + k[1:]]).sum(axis=0).max(), k
            continue
        assert np.abs(ga[k] - gb[k]).max() <= 2e-4 * max(np.abs(gb[k]).max(), 1e-6), k


def test_adam_matches_tf_formula_and_moves_every_row():
    """TF Adam: epsilon outside the bias correction, dense update (rows with zero gradient still see m,v decay)."""
    cfg = OracleConfig(TRIGRAM_D=30, layers=(4, 3), NEG=2, query_BS=3, learning_rate=0.01, act="tanh")  # tanh: no all-zero rows
    params = init_params(cfg, 0)
    orc = DSSMOracle(cfg, params, np.float64)
    rng = np.random.default_rng(0)
    X = random_csr(rng, cfg.rows, cfg.TRIGRAM_D, max_nnz_row=3, allow_empty=False)
    w0 = orc.p["W1"].copy()
    cache = orc.forward(X, True)
    g = orc.backward(cache)
    orc.adam_update(g)
    lr_t = 0.01 * np.sqrt(1 - 0.999) / (1 - 0.9)
    m, v = 0.1 * g["W1"], 0.001 * g["W1"] ** 2
    np.testing.assert_allclose(orc.p["W1"], w0 - lr_t * m / (np.sqrt(v) + 1e-8), rtol=1e-12)
    absent = np.setdiff1d(np.arange(cfg.TRIGRAM_D), np.unique(X.indices))
    assert absent.size and np.array_equal(orc.p["W1"][absent], w0[absent])  # zero grad, zero state: unchanged
    # second step: rows touched in step 1 but not now still move (m != 0)
    X2 = random_csr(np.random.default_rng(5), cfg.rows, cfg.TRIGRAM_D, max_nnz_row=2, allow_empty=False)
    w1 = orc.p["W1"].copy()
    orc.adam_update(orc.backward(orc.forward(X2, True)))
    only_first = np.setdiff1d(np.unique(X.indices), np.unique(X2.indices))
    assert only_first.size and np.all(np.abs(orc.p["W1"][only_first] - w1[only_first]).max(axis=1) > 0)


def test_ema_semantics():
    """Shadows start at zero, decay 0.5, move only when on_train; eval uses the shadows (new_dssm.py:78-86)."""
    cfg = OracleConfig(TRIGRAM_D=20, layers=(4, 3), NEG=2, query_BS=4)
    orc = DSSMOracle(cfg, init_params(cfg, 0))
    X = random_csr(np.random.default_rng(0), cfg.rows, 20, allow_empty=False)
    c1 = orc.forward(X, on_train=True)
    np.testing.assert_allclose(orc.ema["bn1_q_ema_mean"], 0.5 * c1["bn1_q_mean"], rtol=1e-6)
    np.testing.assert_allclose(orc.ema["bn2_d_ema_var"], 0.5 * c1["bn2_d_var"], rtol=1e-6)
    before = {k: v.copy() for k, v in orc.ema.items()}
    ce = orc.forward(X, on_train=False)
    for k in before:
        assert np.array_equal(before[k], orc.ema[k])
    np.testing.assert_allclose(ce["bn1_q_mean"], before["bn1_q_ema_mean"])
    orc.forward(X, on_train=True)
    np.testing.assert_allclose(orc.ema["bn1_q_ema_mean"], 0.75 * c1["bn1_q_mean"], rtol=1e-5)


def test_dp_oracle_single_replica_equals_plain():
    cfg = OracleConfig(TRIGRAM_D=30, layers=(5, 4), NEG=2, query_BS=4)
    p = init_params(cfg, 0)
    X = random_csr(np.random.default_rng(1), cfg.rows, 30, allow_empty=False)
    a, b = DSSMOracle(cfg, p), DPOracle(cfg, p)
    la, lb = a.train_step(X), b.train_step([X])
    assert np.isclose(la, lb)
    for k in a.p:
        np.testing.assert_allclose(a.p[k], b.model.p[k], rtol=1e-6, atol=1e-8)
    for k in a.ema:
        np.testing.assert_allclose(a.ema[k], b.model.ema[k], rtol=1e-6, atol=1e-8)
