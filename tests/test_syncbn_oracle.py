"""SyncBN specification (oracle/syncbn.py): n replicas that all-reduce the BN moments forward and the two BN sums
backward reproduce the single-process reference step on the global batch (new_dssm.py:62-88 over n*B groups)."""
import os
import socket
from dataclasses import replace

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dssm_b200 import Config
from dssm_b200.synthetic import init_params, make_batch
from oracle import DSSMOracle
from oracle.syncbn import SyncBNReplica, restack_global, run_syncbn_threads
from tests.helpers import oracle_config


def _setup(n, act="relu", layers=(24, 16, 12)):
    conf = Config(TRIGRAM_D=300, query_BS=6, NEG=3, layers=layers, act=act)
    steps = [[make_batch(conf, seed=10 * t + r, lam_query=5, lam_doc=9).to_scipy() for r in range(n)] for t in range(3)]
    return conf, steps, init_params(conf, 0)


@pytest.mark.parametrize("n,act", [(2, "relu"), (3, "tanh"), (4, "relu")])
def test_syncbn_replicas_equal_the_global_batch_reference(n, act):
    conf, steps, params = _setup(n, act)
    ocfg = oracle_config(conf)
    gcfg = replace(ocfg, query_BS=n * conf.query_BS)
    for dtype, tol in ((np.float64, 1e-10), (np.float32, 2e-4)):
        reps = run_syncbn_threads(ocfg, params, steps, dtype)
        ref = DSSMOracle(gcfg, params, dtype)
        for batches in steps:
            ref.train_step(restack_global(batches, conf.query_BS, conf.NEG))
        for r in reps[1:]:
            for k in reps[0].p:
                assert np.array_equal(r.p[k], reps[0].p[k]), f"replicas diverged on {k}"
        for k in ref.p:
            if k[0] == "b" and k[1:].isdigit():
                continue  # analytically zero gradient under BN: Adam random-walks it on rounding noise
            scale = max(np.abs(ref.p[k] - params[k].astype(dtype)).max(), 1e-30)
            assert np.abs(reps[0].p[k] - ref.p[k]).max() <= tol * max(scale, np.abs(ref.p[k]).max()), k
        for k in ref.ema:
            if k.endswith("ema_var"):
                np.testing.assert_allclose(reps[0].ema[k], ref.ema[k], rtol=1e-6 if dtype == np.float64 else 1e-3)


def test_syncbn_differs_from_per_replica_moments():
    """Sanity: the per-replica-BN model (DPOracle) is a different model -- SyncBN is not a no-op."""
    from oracle import DPOracle

    conf, steps, params = _setup(2)
    ocfg = oracle_config(conf)
    reps = run_syncbn_threads(ocfg, params, steps[:1], np.float64)
    dp = DPOracle(ocfg, params, np.float64)
    dp.train_step(steps[0])
    assert np.abs(reps[0].p["W2"] - dp.model.p["W2"]).max() > 1e-6


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        conf, steps, params = _setup(world)

        def avg(a):
            t = torch.from_numpy(np.array(a, dtype=np.float64, copy=True))  # all_reduce works in place: never on the caller's array
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            return (t.numpy() / world).astype(a.dtype)

        rep = SyncBNReplica(oracle_config(conf), params, avg, np.float64)
        for batches in steps:
            rep.train_step(batches[rank])
        if rank == 1:
            np.savez(out, **rep.p)
    finally:
        dist.destroy_process_group()


def test_syncbn_over_two_gloo_ranks(tmp_path):
    out = str(tmp_path / "rank1.npz")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = np.load(out)
    conf, steps, params = _setup(2)
    gcfg = replace(oracle_config(conf), query_BS=2 * conf.query_BS)
    ref = DSSMOracle(gcfg, params, np.float64)
    for batches in steps:
        ref.train_step(restack_global(batches, conf.query_BS, conf.NEG))
    for k in ref.p:
        if k[0] == "b" and k[1:].isdigit():
            continue
        assert np.abs(got[k] - ref.p[k]).max() <= 1e-10 * max(np.abs(ref.p[k]).max(), 1e-30), k
