"""The oracle against the committed golden fixtures (tests/golden/make_golden.py): pins the oracle so that the
GPU parity target cannot drift silently.  PARITY UNPINNED w.r.t. the reference itself -- see make_golden.py."""
import numpy as np
import pytest

from oracle import DSSMOracle, init_params
from tests.helpers import GOLDEN_CASES, load_golden, oracle_config


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_oracle_reproduces_golden(name):
    conf, Xs, params, z = load_golden(name)
    orc = DSSMOracle(oracle_config(conf), params)
    for s, X in enumerate(Xs):
        cache = orc.forward(X, on_train=True)
        grads = orc.backward(cache)
        if s == 0:
            for k in ("Y", "cos_sim_raw", "cos_sim", "prob", "dY"):
                np.testing.assert_allclose(cache[k], z[f"fwd/{k}"], rtol=1e-6, atol=1e-7, err_msg=k)
            for k, g in grads.items():
                np.testing.assert_allclose(g, z[f"grad0/{k}"], rtol=1e-5, atol=1e-7, err_msg=k)
        assert np.isclose(float(cache["loss"]), float(z[f"loss{s}"]), rtol=1e-6)
        orc.adam_update(grads)
    last = len(Xs)
    for k, v in orc.p.items():
        np.testing.assert_allclose(v, z[f"param{last}/{k}"], rtol=1e-5, atol=1e-7, err_msg=k)
    for k, v in orc.ema.items():
        np.testing.assert_allclose(v, z[f"ema{last}/{k}"], rtol=1e-6, atol=1e-8, err_msg=k)


def test_golden_params_follow_add_layer_rule():
    conf, _, params, _ = load_golden("tiny_bn_relu")
    ref = init_params(oracle_config(conf), seed=7)
    for k in ref:
        assert np.array_equal(ref[k], params[k])
