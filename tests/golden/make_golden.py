"""Generates tests/golden/*.npz from the oracle (run here: `python tests/golden/make_golden.py`).

PARITY UNPINNED: the reference ships no golden vectors and TensorFlow cannot run in this container, so these
fixtures pin the *oracle* (a restatement of new_dssm.py under the TF-1.x semantics documented in
oracle/dssm_oracle.py), not outputs of the reference itself.  They exist so that (a) the oracle cannot drift
silently and (b) the GPU tests have committed expected values that do not depend on the oracle code of the day.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import DSSMOracle, OracleConfig, init_params  # noqa: E402
from tests.helpers import random_csr  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # name: (config kwargs, value_mode, steps)
    "tiny_bn_relu": (dict(TRIGRAM_D=64, layers=(8, 16), NEG=3, query_BS=5), "count", 3),
    "tiny_nobn_eps": (dict(TRIGRAM_D=64, layers=(8, 16), NEG=3, query_BS=5, use_bn=False, loss_eps=1e-8), "tfidf", 3),
    "tiny_tanh_3layer_sum": (dict(TRIGRAM_D=96, layers=(12, 8, 4), NEG=2, query_BS=7, act="tanh", loss_div_bs=False), "count", 3),
    "odd_shapes_neg1": (dict(TRIGRAM_D=77, layers=(20, 12), NEG=1, query_BS=33), "count", 2),
}


def run_case(name, kw, value_mode, steps):
    cfg = OracleConfig(**kw)
    rng = np.random.default_rng(abs(hash(name)) % (2 ** 31) if False else sum(map(ord, name)))
    params = init_params(cfg, seed=7)
    Xs = [random_csr(rng, cfg.rows, cfg.TRIGRAM_D, max_nnz_row=6, value_mode=value_mode, allow_empty=True) for _ in range(steps)]
    orc = DSSMOracle(cfg, params)
    out = {"cfg_TRIGRAM_D": cfg.TRIGRAM_D, "cfg_layers": np.asarray(cfg.layers), "cfg_NEG": cfg.NEG, "cfg_query_BS": cfg.query_BS,
           "cfg_use_bn": int(cfg.use_bn), "cfg_act": cfg.act, "cfg_loss_eps": cfg.loss_eps, "cfg_loss_div_bs": int(cfg.loss_div_bs),
           "cfg_learning_rate": cfg.learning_rate, "steps": steps}
    for k, v in params.items():
        out[f"param0/{k}"] = v
    for s, X in enumerate(Xs):
        out[f"x{s}/indptr"], out[f"x{s}/indices"], out[f"x{s}/values"] = X.indptr.astype(np.int32), X.indices.astype(np.int32), X.data.astype(np.float32)
        cache = orc.forward(X, on_train=True)
        grads = orc.backward(cache)
        if s == 0:
            for l in range(1, len(cfg.layers) + 1):
                out[f"fwd/h{l}"] = cache[f"h{l}"]
            for k in ("Y", "cos_sim_raw", "cos_sim", "prob", "query_norm_single", "doc_norm", "dY"):
                out[f"fwd/{k}"] = cache[k]
            for k, g in grads.items():
                out[f"grad0/{k}"] = g
        out[f"loss{s}"] = np.float32(cache["loss"])
        orc.adam_update(grads)
        if s in (0, steps - 1):
            for k, v in orc.p.items():
                out[f"param{s + 1}/{k}"] = v
    for k, v in orc.ema.items():
        out[f"ema{steps}/{k}"] = v
    ev = orc.forward(Xs[0], on_train=False)
    out["eval/Y"], out["eval/loss"], out["eval/cos_sim_raw"] = ev["Y"], np.float32(ev["loss"]), ev["cos_sim_raw"]
    assert np.isfinite(out["loss0"]), name
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "loss", [float(out[f"loss{s}"]) for s in range(steps)])


if __name__ == "__main__":
    for name, (kw, vm, steps) in CASES.items():
        run_case(name, kw, vm, steps)
