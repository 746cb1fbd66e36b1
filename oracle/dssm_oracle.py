"""NumPy restatement of the reference DSSM tower graph (TEST INFRASTRUCTURE ONLY).

Follows, stage by stage, ``semantic_matching/dssm/new_dssm.py`` of MC-Zealot/dssm:

  * ``batch_normalization``            new_dssm.py:62-88
  * FC1 (sparse) / add_layer init      new_dssm.py:117-126, archive/dssm_v3.py:44-53
  * BN1 + relu                         new_dssm.py:128-136
  * FC2                                new_dssm.py:138-148
  * BN2 + relu -> embeddings           new_dssm.py:150-158
  * Merge_Negative_Doc                 new_dssm.py:160-180  (closed form; the literal loop is in literal_replay.py)
  * Cosine_Similarity                  new_dssm.py:182-201
  * Loss                               new_dssm.py:203-213  (+1e-8 variant dssm_no_bn/my_dssm.py:169,
                                                             un-normalised variant archive/dssm_v2.py:184)
  * Training (TF AdamOptimizer)        new_dssm.py:215-217

PARITY UNPINNED (see oracle/__init__.py): the reference has no tests or golden
vectors and TensorFlow is not installable here.  The TF-1.x semantics this file
assumes are written next to the code that encodes them (A1..A7 below); gradients
are hand-derived and are cross-checked in tests/ against torch float64 autograd.

Assumptions about TF 1.x (recalled, not verifiable offline):
  A1 tf.nn.moments(x,[0]) = (mean, mean((x-mean)^2))  -- biased, two-pass.
  A2 tf.nn.batch_normalization: inv = rsqrt(var+eps)*gamma; y = x*inv + (beta - mean*inv).
  A3 ExponentialMovingAverage(decay).apply on tensors: zero-initialised shadow,
     shadow -= (1-decay)*(shadow - value); no zero_debias, no num_updates.
  A4 tf.cond(on_train): EMA update only when training; training uses batch stats,
     inference uses the shadows.
  A5 tf.nn.softmax subtracts the row max; tf.truediv is IEEE divide (0/0 = NaN).
  A6 AdamOptimizer: m,v zero-init; beta powers start at beta1,beta2 and are multiplied
     after every apply; lr_t = lr*sqrt(1-b2p)/(1-b1p); w -= lr_t*m/(sqrt(v)+eps);
     applied densely to every trainable variable (all of W1 moves every step).
  A7 grad of sparse_tensor_dense_matmul wrt the dense operand is the dense X^T dY;
     the three tower applications (query/pos/neg) share W,b so their grads add.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import scipy.sparse as sp


@dataclass
class OracleConfig:
    """Hyper-parameters, names as in semantic_matching/dssm/config.py:19-28."""

    TRIGRAM_D: int
    layers: Tuple[int, ...] = (100, 100)  # (L1_N, L2_N[, L3_N ...]); reference graph has exactly two
    NEG: int = 4
    query_BS: int = 400
    learning_rate: float = 0.01
    use_bn: bool = True  # True: semantic_matching/dssm ; False: semantic_matching/dssm_no_bn
    act: str = "relu"  # reference always relu; north_star allows tanh
    bn_eps: float = 1e-3  # new_dssm.py:87
    ema_decay: float = 0.5  # new_dssm.py:78
    gamma: float = 20.0  # new_dssm.py:199
    loss_eps: float = 0.0  # 1e-8 in dssm_no_bn/my_dssm.py:169
    loss_div_bs: bool = True  # False in archive/dssm_v2.py:184
    beta1: float = 0.9
    beta2: float = 0.999
    adam_eps: float = 1e-8

    @property
    def rows(self) -> int:
        return (2 + self.NEG) * self.query_BS

    def layer_dims(self) -> List[Tuple[int, int]]:
        dims, d_in = [], self.TRIGRAM_D
        for n in self.layers:
            dims.append((d_in, n))
            d_in = n
        return dims


def init_params(cfg: OracleConfig, seed: int = 0, dtype=np.float32) -> Dict[str, np.ndarray]:
    """add_layer rule (archive/dssm_v3.py:44-53, new_dssm.py:118-120,139-142):
    W and b ~ U(+-sqrt(6/(in+out))); BN beta=0, gamma=1 (new_dssm.py:75-76)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    p: Dict[str, np.ndarray] = {}
    for l, (d_in, d_out) in enumerate(cfg.layer_dims(), start=1):
        lim = np.sqrt(6.0 / (d_in + d_out))
        p[f"W{l}"] = rng.uniform(-lim, lim, size=(d_in, d_out)).astype(dtype)
        p[f"b{l}"] = rng.uniform(-lim, lim, size=(d_out,)).astype(dtype)
    if cfg.use_bn:
        for l, (_, d_out) in enumerate(cfg.layer_dims(), start=1):
            for seg in ("q", "d"):
                p[f"bn{l}_{seg}_beta"] = np.zeros(d_out, dtype)
                p[f"bn{l}_{seg}_gamma"] = np.ones(d_out, dtype)
    return p


def _act(x, kind):
    if kind == "relu":
        return np.maximum(x, 0)
    if kind == "tanh":
        return np.tanh(x)
    raise ValueError(kind)


def _act_grad_from_out(a, kind):
    if kind == "relu":
        return (a > 0).astype(a.dtype)
    return 1 - a * a


class DSSMOracle:
    """One replica of the reference graph.  Rows of X are [query(B); doc_pos(B); doc_neg(B*NEG)],
    negatives of query j at neg rows j*NEG..j*NEG+NEG-1 (utils/utils.py:49-51)."""

    def __init__(self, cfg: OracleConfig, params: Dict[str, np.ndarray], dtype=np.float32):
        self.cfg = cfg
        self.dtype = np.dtype(dtype)
        self.p = {k: np.array(v, dtype=self.dtype) for k, v in params.items()}
        self.ema: Dict[str, np.ndarray] = {}
        if cfg.use_bn:
            for l, (_, d_out) in enumerate(cfg.layer_dims(), start=1):
                for seg in ("q", "d"):
                    self.ema[f"bn{l}_{seg}_ema_mean"] = np.zeros(d_out, self.dtype)  # A3
                    self.ema[f"bn{l}_{seg}_ema_var"] = np.zeros(d_out, self.dtype)
        self.m = {k: np.zeros_like(v) for k, v in self.p.items()}  # A6
        self.v = {k: np.zeros_like(v) for k, v in self.p.items()}
        self.beta1_power = self.dtype.type(cfg.beta1)
        self.beta2_power = self.dtype.type(cfg.beta2)
        self.n_layers = len(cfg.layers)

    # ------------------------------------------------------------------ forward
    def _segments(self, n_rows: int):
        B = self.cfg.query_BS
        return (("q", slice(0, B)), ("d", slice(B, n_rows)))

    def _bn_forward(self, l: int, h: np.ndarray, on_train: bool, cache: dict, update_ema: bool):
        """batch_normalization, new_dssm.py:62-88; query BN over B rows (:129,151), doc BN over
        concat([pos,neg]) = (1+NEG)*B rows (:130,152)."""
        cfg, dt = self.cfg, self.dtype
        y = np.empty_like(h)
        for seg, sl in self._segments(h.shape[0]):
            x = h[sl]
            if on_train:
                mean, var = self._bn_moments(x)  # A1
                if update_ema:  # A3/A4
                    for nm, val in (("mean", mean), ("var", var)):
                        key = f"bn{l}_{seg}_ema_{nm}"
                        self.ema[key] = (self.ema[key] - dt.type(1 - cfg.ema_decay) * (self.ema[key] - val)).astype(dt)
            else:
                mean = self.ema[f"bn{l}_{seg}_ema_mean"]
                var = self.ema[f"bn{l}_{seg}_ema_var"]
            rstd = (1 / np.sqrt(var + dt.type(cfg.bn_eps))).astype(dt)
            inv = rstd * self.p[f"bn{l}_{seg}_gamma"]  # A2
            y[sl] = x * inv + (self.p[f"bn{l}_{seg}_beta"] - mean * inv)
            cache[f"bn{l}_{seg}_mean"] = mean
            cache[f"bn{l}_{seg}_var"] = var
            cache[f"bn{l}_{seg}_rstd"] = rstd
        return y

    def tower(self, X: sp.csr_matrix, on_train: bool, update_ema: bool = True) -> dict:
        """FC1..BNn: new_dssm.py:117-158.  Returns every intermediate tensor."""
        cfg, dt = self.cfg, self.dtype
        cache: dict = {"X": X}
        a = None
        for l in range(1, self.n_layers + 1):
            W, b = self.p[f"W{l}"], self.p[f"b{l}"]
            if l == 1:
                h = self._spmm(X.astype(dt), W) + b  # sparse_tensor_dense_matmul :124-126
            else:
                h = a @ W + b  # tf.matmul :146-148
            cache[f"h{l}"] = h
            y = self._bn_forward(l, h, on_train, cache, update_ema) if cfg.use_bn else h
            a = _act(y, cfg.act).astype(dt)
            cache[f"a{l}"] = a
        cache["Y"] = a
        return cache

    def embeddings(self, cache: dict):
        """embedding_query_y / embedding_doc_positive_y / embedding_doc_negative_y, new_dssm.py:156-158."""
        B = self.cfg.query_BS
        Y = cache["Y"]
        return Y[:B], Y[B : 2 * B], Y[2 * B :]

    def cosine_loss(self, cache: dict) -> dict:
        """Merge_Negative_Doc + Cosine_Similarity + Loss in closed form (new_dssm.py:160-213).
        cos_sim[j,0] = positive of query j, cos_sim[j,k>=1] = negative row j*NEG+k-1."""
        cfg, dt = self.cfg, self.dtype
        B, N = cfg.query_BS, cfg.NEG
        q, pos, neg = self.embeddings(cache)
        L = q.shape[1]
        docs = np.concatenate([pos[:, None, :], neg.reshape(B, N, L)], axis=1)  # [B,1+N,L]
        dot = np.einsum("jl,jkl->jk", q, docs).astype(dt)
        qn = np.sqrt(np.sum(q * q, axis=1, dtype=dt))  # query_norm_single :187
        dn = np.sqrt(np.sum(docs * docs, axis=2, dtype=dt))  # doc_norm :190
        with np.errstate(invalid="ignore", divide="ignore"):
            raw = dot / (qn[:, None] * dn)  # cos_sim_raw :197 (no epsilon)
        cos_sim = raw * dt.type(cfg.gamma)  # :199
        z = cos_sim - cos_sim.max(axis=1, keepdims=True)  # A5
        e = np.exp(z)
        prob = e / e.sum(axis=1, keepdims=True, dtype=dt)  # :206
        hit = prob[:, 0]  # :208
        denom = dt.type(B if cfg.loss_div_bs else 1)
        loss = -np.sum(np.log(hit + dt.type(cfg.loss_eps)), dtype=dt) / denom  # :209
        out = dict(docs=docs, dot=dot, query_norm_single=qn, doc_norm_grouped=dn, raw=raw,
                   cos_sim=cos_sim, prob=prob, hit_prob=hit, loss=dt.type(loss))
        # reference memory order of cos_sim_raw / doc_norm: row k*B + j  (new_dssm.py:162-199)
        out["cos_sim_raw"] = raw.T.reshape(-1).copy()
        out["doc_norm"] = dn.T.reshape(-1).copy()
        out["accuracy"] = dt.type(np.mean(np.argmax(prob, axis=1) == 0))  # :220-221
        return out

    def forward(self, X: sp.csr_matrix, on_train: bool, update_ema: bool = True) -> dict:
        cache = self.tower(X, on_train, update_ema)
        cache.update(self.cosine_loss(cache))
        return cache

    # ----------------------------------------------------------------- backward
    def backward(self, cache: dict) -> Dict[str, np.ndarray]:
        """Reverse-mode gradient of `loss` wrt the trainables of new_dssm.py:217 (training mode)."""
        cfg, dt = self.cfg, self.dtype
        B, N = cfg.query_BS, cfg.NEG
        q, _, _ = self.embeddings(cache)
        docs, raw, prob = cache["docs"], cache["raw"], cache["prob"]
        qn, dn = cache["query_norm_single"], cache["doc_norm_grouped"]
        L = q.shape[1]
        denom = dt.type(B if cfg.loss_div_bs else 1)
        onehot = np.zeros_like(prob)
        onehot[:, 0] = 1
        w = (prob[:, 0] / (prob[:, 0] + dt.type(cfg.loss_eps)))[:, None]
        dlogit = w * (prob - onehot) / denom
        dcos = (dlogit * dt.type(cfg.gamma)).astype(dt)
        with np.errstate(invalid="ignore", divide="ignore"):
            inv_qd = 1 / (qn[:, None] * dn)  # [B,1+N]
            dq = np.einsum("jk,jkl->jl", dcos * inv_qd, docs) - (np.sum(dcos * raw, axis=1) / (qn * qn))[:, None] * q
            ddocs = (dcos * inv_qd)[:, :, None] * q[:, None, :] - (dcos * raw / (dn * dn))[:, :, None] * docs
        dY = np.empty_like(cache["Y"])
        dY[:B] = dq
        dY[B : 2 * B] = ddocs[:, 0, :]
        dY[2 * B :] = ddocs[:, 1:, :].reshape(B * N, L)

        grads: Dict[str, np.ndarray] = {}
        dA = dY.astype(dt)
        for l in range(self.n_layers, 0, -1):
            a, h = cache[f"a{l}"], cache[f"h{l}"]
            g = (dA * _act_grad_from_out(a, cfg.act)).astype(dt)
            if cfg.use_bn:
                dh = np.empty_like(g)
                for seg, sl in self._segments(h.shape[0]):
                    n = dt.type(h[sl].shape[0])
                    mean, rstd = cache[f"bn{l}_{seg}_mean"], cache[f"bn{l}_{seg}_rstd"]
                    gamma = self.p[f"bn{l}_{seg}_gamma"]
                    xhat = (h[sl] - mean) * rstd
                    dbeta, dgamma = self._bn_bwd_sums(g[sl], xhat)
                    dh[sl] = (gamma * rstd) * (g[sl] - dbeta / n - xhat * (dgamma / n))
                    grads[f"bn{l}_{seg}_beta"] = dbeta
                    grads[f"bn{l}_{seg}_gamma"] = dgamma
            else:
                dh = g
            grads[f"b{l}"] = dh.sum(axis=0, dtype=dt)
            if l > 1:
                grads[f"W{l}"] = (cache[f"a{l-1}"].T @ dh).astype(dt)
                dA = (dh @ self.p[f"W{l}"].T).astype(dt)
            else:
                X = cache["X"].astype(dt)
                grads["W1"] = self._spmm_t(X, dh)  # A7: dense [D, L1]
            cache[f"dh{l}"] = dh
        cache["dY"] = dY
        return grads

    # --------------------------------------------------------------------- Adam
    def adam_update(self, grads: Dict[str, np.ndarray]) -> None:
        """tf.train.AdamOptimizer(lr).minimize, new_dssm.py:217 (A6)."""
        cfg, dt = self.cfg, self.dtype
        one = dt.type(1)
        lr_t = dt.type(cfg.learning_rate) * np.sqrt(one - self.beta2_power) / (one - self.beta1_power)
        b1, b2, eps = dt.type(cfg.beta1), dt.type(cfg.beta2), dt.type(cfg.adam_eps)
        for k, g in grads.items():
            self._adam_tensor(k, g, lr_t, b1, b2, eps)
        self.beta1_power = dt.type(self.beta1_power * b1)
        self.beta2_power = dt.type(self.beta2_power * b2)

    # The three places where the work is proportional to nnz or to the parameter count.  They are separate methods so that
    # bench.py's CPU legs can run the SAME arithmetic on all host threads (bench.py:ThreadedPort); tests use these.
    # The two places where BatchNorm couples the rows of a batch.  Separate methods so that oracle/syncbn.py can turn them
    # into cross-replica reductions (SyncBN) without restating the rest of the graph.
    def _bn_moments(self, x: np.ndarray):
        dt = self.dtype
        mean = x.mean(axis=0, dtype=dt)
        var = np.mean((x - mean) ** 2, axis=0, dtype=dt)
        return mean, var

    def _bn_bwd_sums(self, g: np.ndarray, xhat: np.ndarray):
        dt = self.dtype
        return g.sum(axis=0, dtype=dt), (g * xhat).sum(axis=0, dtype=dt)

    def _spmm(self, X: sp.csr_matrix, W: np.ndarray) -> np.ndarray:
        return np.asarray(X @ W, dtype=self.dtype)

    def _spmm_t(self, X: sp.csr_matrix, dh: np.ndarray) -> np.ndarray:
        return np.asarray(X.T @ dh, dtype=self.dtype)

    def _adam_tensor(self, k: str, g: np.ndarray, lr_t, b1, b2, eps) -> None:
        dt, one = self.dtype, self.dtype.type(1)
        self.m[k] = (b1 * self.m[k] + (one - b1) * g).astype(dt)
        self.v[k] = (b2 * self.v[k] + (one - b2) * (g * g)).astype(dt)
        self.p[k] = (self.p[k] - lr_t * self.m[k] / (np.sqrt(self.v[k]) + eps)).astype(dt)

    def train_step(self, X: sp.csr_matrix) -> float:
        """sess.run(train_step, feed_dict=pull_batch(True, ...)), new_dssm.py:267-269."""
        cache = self.forward(X, on_train=True)
        grads = self.backward(cache)
        self.adam_update(grads)
        return float(cache["loss"])


class DPOracle:
    """Data-parallel semantics defined by this project (the reference is single-process,
    SURVEY.md section 8e): n replicas share parameters; each runs the reference graph on its
    own B query groups with per-replica BN moments; gradients and the batch statistics fed to
    the EMA shadows are averaged over replicas; one Adam update."""

    def __init__(self, cfg: OracleConfig, params: Dict[str, np.ndarray], dtype=np.float32):
        self.model = DSSMOracle(cfg, params, dtype)

    def train_step(self, batches: Sequence[sp.csr_matrix]) -> float:
        mdl, dt = self.model, self.model.dtype
        n = len(batches)
        grads_sum: Optional[Dict[str, np.ndarray]] = None
        stats_sum: Dict[str, np.ndarray] = {}
        losses = []
        for X in batches:
            cache = mdl.forward(X, on_train=True, update_ema=False)
            g = mdl.backward(cache)
            losses.append(float(cache["loss"]))
            grads_sum = g if grads_sum is None else {k: grads_sum[k] + g[k] for k in g}
            if mdl.cfg.use_bn:
                for l in range(1, mdl.n_layers + 1):
                    for seg in ("q", "d"):
                        for nm in ("mean", "var"):
                            key = f"bn{l}_{seg}_{nm}"
                            stats_sum[key] = stats_sum.get(key, 0) + cache[key]
        grads = {k: (v / dt.type(n)).astype(dt) for k, v in grads_sum.items()}
        if mdl.cfg.use_bn:
            dec = dt.type(1 - mdl.cfg.ema_decay)
            for key, s in stats_sum.items():
                l_seg, nm = key.rsplit("_", 1)
                ek = f"{l_seg}_ema_{nm}"
                mdl.ema[ek] = (mdl.ema[ek] - dec * (mdl.ema[ek] - (s / dt.type(n)).astype(dt))).astype(dt)
        mdl.adam_update(grads)
        return float(np.mean(losses))
