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


def tf_adam_reference(p0, g, m0, v0, b1p, b2p, lr=0.01, beta1=0.9, beta2=0.999, eps=1e-8, grad_scale=1.0):
    """tf.train.AdamOptimizer's ApplyAdam in float64 with TF's fp32 scalars ((1-beta) is 1-float32(beta))."""
    f = lambda x: float(np.float32(x))
    b1, b2, lr, eps = f(beta1), f(beta2), f(lr), f(eps)
    omb1, omb2 = float(np.float32(1) - np.float32(beta1)), float(np.float32(1) - np.float32(beta2))
    g = np.asarray(g, np.float64) * grad_scale
    m = b1 * np.asarray(m0, np.float64) + omb1 * g
    v = b2 * np.asarray(v0, np.float64) + omb2 * g * g
    lr_t = lr * np.sqrt(1 - float(b2p)) / (1 - float(b1p))
    return np.asarray(p0, np.float64) - lr_t * m / (np.sqrt(v) + eps), m, v


def assert_update_close(got, ref, init, use_bn, what="", l2_tol=2e-2):
    """Parameters after Adam steps, compared as the relative L2 error of the *update* (got-ref vs ref-init).
    An element-wise max-norm check is ill-posed here for ANY two fp32 implementations (TF-CPU vs TF-GPU included):
    Adam divides by sqrt(v)+1e-8, so for the many W entries whose gradient is below ~1e-6 the update is
    lr*g/3e-7 -- fp32 summation noise of 1e-9 in g moves the weight by 3e-5, i.e. 0.2 % of the weight scale --
    and under BN the pre-BN biases b{l} have an analytically ZERO gradient, so their update is pure rounding noise
    (they are skipped; BN removes them from the function).  The Adam arithmetic itself is pinned separately, to
    1e-6, by applying the TF formula on the host to the GPU's own gradients (test_gpu_tower.py)."""
    for k in ref:
        if use_bn and k[0] == "b" and k[1:].isdigit():
            continue
        g, r, i0 = np.asarray(got[k], np.float64), np.asarray(ref[k], np.float64), np.asarray(init[k], np.float64)
        assert g.shape == r.shape, k
        upd = np.linalg.norm((r - i0).ravel())
        if upd > 0:
            rel = np.linalg.norm((g - r).ravel()) / upd
            if os.environ.get("DSSM_TEST_REPORT"):  # measured values behind the tolerances (appended to the named file)
                with open(os.environ["DSSM_TEST_REPORT"], "a") as f:
                    f.write(f"{what} {k} rel_l2_of_update {rel:.3e} (tol {l2_tol})\n")
            assert rel <= l2_tol, f"{what} {k}: update differs by {rel:.2e} in relative L2"


def grad_tolerances(conf, X, params, tol=5e-5, masks=None):
    """Reference gradients in float64 plus, per tensor, the error the NumPy fp32 port itself makes against them.
    A GPU gradient passes if it is within `tol` of the tensor scale OR within 4x the fp32 port's own error
    (reductions with cancellation, e.g. BN beta/gamma sums over thousands of rows, are noise-limited for every
    fp32 implementation)."""
    from oracle import DSSMOracle

    o64 = DSSMOracle(oracle_config(conf), params, dtype=np.float64)
    c64 = o64.forward(X, on_train=True, update_ema=False)
    o32 = DSSMOracle(oracle_config(conf), params, dtype=np.float32)
    c32 = o32.forward(X, on_train=True, update_ema=False)
    if masks is not None:  # differentiate with the device's active set (align_relu_masks)
        c64["kink_flips"] = align_relu_masks(c64, masks)
        align_relu_masks(c32, masks)
    g64 = o64.backward(c64)
    g32 = o32.backward(c32)
    allow = {}
    for k in g64:
        scale = max(np.abs(g64[k]).max(), 1e-30)
        allow[k] = max(tol * scale, 4 * np.abs(g32[k].astype(np.float64) - g64[k]).max())
    return g64, allow, c64


def assert_grads_close(conf, got, g64, allow, c64):
    for k, ref in g64.items():
        if conf.use_bn and k[0] == "b" and k[1:].isdigit():
            # analytically zero under BN; what is left is summation noise of dh{l}'s columns
            bound = 1e-5 * float(np.abs(c64["dh" + k[1:]]).sum(axis=0).max())
            assert np.abs(got[k] - ref).max() <= bound, f"grad {k} (zero under BN): {np.abs(got[k] - ref).max():.3e} > {bound:.3e}"
            continue
        err = np.abs(np.asarray(got[k], np.float64) - ref).max()
        assert err <= allow[k], f"grad {k}: abs error {err:.3e} > allowed {allow[k]:.3e} (scale {np.abs(ref).max():.3e})"


# normalised pre-activations are O(1); BN's mean subtraction amplifies the rounding of h (|h - mean| << |h| for the
# count-valued features of C4), so they carry an absolute error of ~1e-5 in any fp32 implementation
KINK_TOL = 1e-4


def gpu_relu_masks(conf, t):
    """Which units the GPU step treated as active, per layer, rebuilt from its stored tensors.  The kernels evaluate
    act(fmaf(h, scale, shift)) everywhere (forward producers and backward alike); float64 holds h*scale exactly and
    rounds the sum correctly, so the sign computed here is the sign the device saw."""
    B, n = conf.query_BS, len(conf.layers)
    out = {}
    for l in range(1, n + 1):
        h = t.tensor(f"h{l}").cpu().numpy().astype(np.float64)
        if conf.use_bn:
            sc = t.tensor(f"bn{l}_scale").cpu().numpy().astype(np.float64)
            sh = t.tensor(f"bn{l}_shift").cpu().numpy().astype(np.float64)
            y = np.concatenate([h[:B] * sc[0] + sh[0], h[B:] * sc[1] + sh[1]])
        else:
            y = h
        out[l] = (y > 0, np.abs(y))
    return out


def align_relu_masks(cache, masks, kink_tol=KINK_TOL):
    """relu has a kink at 0: a pre-activation within fp32 rounding of zero gets derivative 1 in one implementation and 0
    in another, and every gradient fed by that unit then moves by its full upstream value -- for ANY two fp32
    implementations.  With R*L pre-activations ~ N(0,1) that hits ~1e-6 of them: never in the small tests, about one unit
    per step at C2, several at C4.  So the oracle differentiates with the device's active set, after checking that the
    two sets differ only at units whose pre-activation is within `kink_tol` of zero on both sides.  Returns the number
    of such units."""
    flips = 0
    for l, (m, absy) in masks.items():
        a = cache[f"a{l}"]
        diff = m != (a > 0)
        if diff.any():
            worst = max(float(np.abs(a[diff]).max()), float(absy[diff].max()))
            assert worst < kink_tol, f"layer {l}: active sets differ at a unit {worst:.3e} away from the relu kink"
            flips += int(diff.sum())
            cache[f"a{l}"] = np.where(m, np.maximum(a, np.finfo(a.dtype).tiny), 0).astype(a.dtype)
    return flips
