"""bench.py's CPU legs time the oracle with its sparse products and Adam spread over host threads: that port must be
the oracle's arithmetic, element for element."""
import numpy as np

import bench
from dssm_b200 import Config
from dssm_b200.synthetic import init_params, make_batch
from oracle import DSSMOracle
from tests.helpers import oracle_config


def test_threaded_port_is_bit_identical_to_the_oracle():
    conf = Config(TRIGRAM_D=9000, query_BS=96, NEG=4, layers=(128, 64, 32))  # W1 > 2^20 elements: threaded Adam path
    X = make_batch(conf, seed=4, lam_query=8, lam_doc=16).to_scipy()
    params = init_params(conf, 0)
    a = DSSMOracle(oracle_config(conf), params)
    t = bench.threaded_port(oracle_config(conf), params, threads=3)
    for _ in range(3):
        la, lt = a.train_step(X), t.train_step(X)
        assert la == lt
    for k in a.p:
        assert np.array_equal(a.p[k], t.p[k]) and np.array_equal(a.m[k], t.m[k]) and np.array_equal(a.v[k], t.v[k]), k
    for k in a.ema:
        assert np.array_equal(a.ema[k], t.ema[k]), k
