"""N>1 host logic on CPU: two gloo ranks shard a global batch, compute per-replica gradients (oracle as the
per-rank compute, since the kernels need a GPU), average them with the same all-reduce helper the GPU path uses,
and must land on DPOracle's parameters.  Also the document-shard ranges and top-k merge of the retrieval path."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dssm_b200 import Config
from dssm_b200.batch import stack_csr
from dssm_b200.parallel import allreduce_mean_, shard_stacked_batch
from dssm_b200.retrieval import shard_range
from oracle import DPOracle, DSSMOracle, corpus_topk_oracle, init_params, merge_topk_oracle
from tests.helpers import oracle_config, random_csr


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
        conf = Config(TRIGRAM_D=60, query_BS=8, NEG=3, layers=(10, 16))
        rng = np.random.default_rng(0)
        sb = stack_csr(random_csr(rng, 8, 60, allow_empty=False), random_csr(rng, 8, 60, allow_empty=False),
                       random_csr(rng, 24, 60, allow_empty=False), 8, 3)
        local_conf = Config(TRIGRAM_D=60, query_BS=8 // world, NEG=3, layers=(10, 16))
        mine = shard_stacked_batch(sb, 8, 3, rank, world).to_scipy()
        orc = DSSMOracle(oracle_config(local_conf), init_params(oracle_config(local_conf), 0))
        cache = orc.forward(mine, on_train=True, update_ema=True)
        grads = orc.backward(cache)
        keys = sorted(grads)
        flat = torch.from_numpy(np.concatenate([grads[k].ravel() for k in keys]).astype(np.float32))
        ema_keys = sorted(orc.ema)
        ema = torch.from_numpy(np.concatenate([orc.ema[k].ravel() for k in ema_keys]).astype(np.float32))
        allreduce_mean_([flat, ema])
        off = 0
        avg = {}
        for k in keys:
            n = grads[k].size
            avg[k] = flat[off:off + n].numpy().reshape(grads[k].shape)
            off += n
        orc.adam_update(avg)
        if rank == 0:
            np.savez(out, ema=ema.numpy(), **{k: orc.p[k] for k in orc.p})
    finally:
        dist.destroy_process_group()


def test_two_rank_data_parallel_matches_dp_oracle(tmp_path):
    out = str(tmp_path / "rank0.npz")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = np.load(out)
    conf = Config(TRIGRAM_D=60, query_BS=8, NEG=3, layers=(10, 16))
    rng = np.random.default_rng(0)
    sb = stack_csr(random_csr(rng, 8, 60, allow_empty=False), random_csr(rng, 8, 60, allow_empty=False),
                   random_csr(rng, 24, 60, allow_empty=False), 8, 3)
    local_conf = Config(TRIGRAM_D=60, query_BS=4, NEG=3, layers=(10, 16))
    ocfg = oracle_config(local_conf)
    dp = DPOracle(ocfg, init_params(ocfg, 0))
    dp.train_step([shard_stacked_batch(sb, 8, 3, r, 2).to_scipy() for r in range(2)])
    for k in dp.model.p:
        np.testing.assert_allclose(got[k], dp.model.p[k], rtol=2e-5, atol=1e-7)
    ema = np.concatenate([dp.model.ema[k].ravel() for k in sorted(dp.model.ema)])
    np.testing.assert_allclose(got["ema"], ema, rtol=1e-5, atol=1e-7)


def test_doc_shard_ranges_cover_corpus_and_merge_is_exact():
    n = 1003
    ranges = [shard_range(n, r, 8) for r in range(8)]
    assert ranges[0][0] == 0 and ranges[-1][1] == n
    assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
    rng = np.random.default_rng(0)
    Q = np.maximum(rng.standard_normal((5, 8)), 0).astype(np.float32)
    docs = np.maximum(rng.standard_normal((n, 8)), 0).astype(np.float32)
    parts = [corpus_topk_oracle(Q, docs[lo:hi], 20, id_offset=lo) for lo, hi in ranges]
    whole = corpus_topk_oracle(Q, docs, 20)
    merged = merge_topk_oracle(parts, 20)
    assert np.array_equal(merged[1], whole[1])
