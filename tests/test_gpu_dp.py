"""Data-parallel training on real GPUs (needs >= 2; skipped otherwise): two NCCL ranks, each with its own query
groups, pipelined gather/all-reduce/Adam (and its CUDA-graph form) against DPOracle (per-replica BN moments, mean
gradient, one Adam step); the same with dW1 exchanged over NVLink peer memory (comm="nvlink": pull / Adam / push kernel
between two symmetric-memory barriers) instead of NCCL."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _join_all(ctx, seconds, what):
    """mp.spawn(join=False).join() returns as soon as ONE process is done: loop until all are, with a deadline."""
    import time

    deadline = time.time() + seconds
    while not ctx.join(timeout=5):
        if time.time() > deadline:
            for p in ctx.processes:
                if p.is_alive():
                    p.kill()
            pytest.fail(f"{what} did not finish in {seconds} s")


def _worker(rank, world, port, out, use_graph, comm="nccl"):
    import torch.distributed as dist

    from dssm_b200 import Config, DSSMTower
    from dssm_b200.parallel import DataParallelTower
    from dssm_b200.synthetic import init_params, make_batch

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    conf = Config(TRIGRAM_D=21128, query_BS=100, NEG=4, layers=(300, 300, 128), gemm_mode="tc_3xtf32")
    batches = [make_batch(conf, seed=10 * s + rank, lam_query=12, lam_doc=24) for s in range(2)]
    t = DSSMTower(conf, max_nnz=max(b.nnz for b in batches), device=f"cuda:{rank}", params=init_params(conf, 0),
                  symmetric=(comm == "nvlink"))
    dp = DataParallelTower(t, n_chunks=3, comm=comm)
    losses = []
    if use_graph:
        # capture on a scratch copy of the state, then restore it so the comparison starts from the initial parameters
        snap = {k: getattr(t, k).clone() for k in ("params", "m", "v", "comm", "beta_pow")}
        t.stage(t.to_device(batches[0]))
        dp.capture_graph(warmup=2)
        for k, v in snap.items():
            getattr(t, k).copy_(v)
    for b in batches:
        losses.append(dp.train_step(t.to_device(b)).item())
    torch.cuda.synchronize()
    if rank == 0:
        np.savez(out, losses=np.asarray(losses), **{"p_" + k: v for k, v in t.export_params().items()},
                 **{"e_" + k: v for k, v in t.export_ema().items()})
    dp.graph = None
    dist.barrier()
    torch.cuda.synchronize()
    os._exit(0)  # skip the NCCL destructor (can block with captured collectives)


@pytest.mark.parametrize("use_graph,comm", [(False, "nccl"), (True, "nccl"), (False, "nvlink"), (True, "nvlink")])
def test_two_gpu_data_parallel_matches_dp_oracle(tmp_path, use_graph, comm):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    from dssm_b200 import Config
    from dssm_b200.synthetic import init_params, make_batch
    from oracle import DPOracle
    from tests.helpers import assert_close, assert_update_close, oracle_config

    out = str(tmp_path / "rank0.npz")
    ctx = mp.spawn(_worker, args=(2, _free_port(), out, use_graph, comm), nprocs=2, join=False)
    _join_all(ctx, 240, "data-parallel workers")
    got = np.load(out)
    conf = Config(TRIGRAM_D=21128, query_BS=100, NEG=4, layers=(300, 300, 128))
    params = init_params(conf, 0)
    dp = DPOracle(oracle_config(conf), params)
    ref_losses = []
    for s in range(2):
        mats = [make_batch(conf, seed=10 * s + r, lam_query=12, lam_doc=24).to_scipy() for r in range(2)]
        cache0 = dp.model.forward(mats[0], on_train=True, update_ema=False)  # rank 0's local loss is what rank 0 reports
        ref_losses.append(float(cache0["loss"]))
        dp.train_step(mats)
    assert abs(got["losses"][0] - ref_losses[0]) <= 1e-5 * abs(ref_losses[0])
    assert abs(got["losses"][1] - ref_losses[1]) <= 2e-3 * abs(ref_losses[1])
    # measured maximum (DSSM_TEST_REPORT, round 2): 5.6e-2 on bn1_q_beta, 1.3e-2 on W1; tolerance = measured + 40 %
    assert_update_close({k: got["p_" + k] for k in params}, dp.model.p, params, conf.use_bn, "dp params", l2_tol=0.08)
    for k, v in dp.model.ema.items():
        if k.endswith("ema_var"):
            assert_close(got["e_" + k], v, 2e-2, f"dp ema {k}")


def _retrieval_worker(rank, world, port, out):
    import torch.distributed as dist

    from dssm_b200.retrieval import shard_range, sharded_corpus_topk

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    rng = np.random.default_rng(0)
    Q = np.maximum(rng.standard_normal((96, 128)), 0).astype(np.float32)
    D = np.maximum(rng.standard_normal((90001, 128)), 0).astype(np.float32)
    D[70000] = D[11]  # tie across shards
    lo, hi = shard_range(D.shape[0], rank, world)
    s, i = sharded_corpus_topk(torch.from_numpy(Q).cuda(), torch.from_numpy(D[lo:hi]).cuda(), 100, id_offset=lo)
    torch.cuda.synchronize()
    if rank == 1:  # every rank holds the merged result; check a non-zero rank
        np.savez(out, s=s.cpu().numpy(), i=i.cpu().numpy())
    dist.barrier()
    torch.cuda.synchronize()
    os._exit(0)


def test_two_gpu_sharded_retrieval_is_bit_exact(tmp_path):
    """Corpus sharded by doc id over two GPUs, local tensor-core top-k, NCCL all-gather, merge kernel == oracle."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    from oracle import corpus_topk_oracle

    out = str(tmp_path / "r.npz")
    ctx = mp.spawn(_retrieval_worker, args=(2, _free_port(), out), nprocs=2, join=False)
    _join_all(ctx, 240, "retrieval workers")
    got = np.load(out)
    rng = np.random.default_rng(0)
    Q = np.maximum(rng.standard_normal((96, 128)), 0).astype(np.float32)
    D = np.maximum(rng.standard_normal((90001, 128)), 0).astype(np.float32)
    D[70000] = D[11]
    rs, ri = corpus_topk_oracle(Q, D, 100)
    assert np.array_equal(got["i"], ri) and np.array_equal(got["s"], rs)


def _syncbn_worker(rank, world, port, out, comm):
    import torch.distributed as dist

    from dssm_b200 import Config, DSSMTower
    from dssm_b200.parallel import DataParallelTower
    from dssm_b200.synthetic import init_params, make_batch

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    conf = Config(TRIGRAM_D=21128, query_BS=100, NEG=4, layers=(300, 300, 128), gemm_mode="tc_3xtf32")
    batches = [make_batch(conf, seed=10 * s + rank, lam_query=12, lam_doc=24) for s in range(2)]
    t = DSSMTower(conf, max_nnz=max(b.nnz for b in batches), device=f"cuda:{rank}", params=init_params(conf, 0),
                  symmetric=(comm == "nvlink"))
    dp = DataParallelTower(t, n_chunks=2, comm=comm, sync_bn=True)
    assert dp.sync_bn
    losses, grads0 = [], None
    for s, b in enumerate(batches):
        losses.append(dp.train_step(t.to_device(b)).item())
        if s == 0:
            torch.cuda.synchronize()
            grads0 = {k: v for k, v in t.export_grads().items() if k != "W1"}  # the small gradients after their all-reduce
            stats0 = {f"bn{l}_{nm}": t.tensor(f"bn{l}_{nm}").cpu().numpy() for l in (1, 2, 3) for nm in ("mean", "var")}
    torch.cuda.synchronize()
    np.savez(out + f".{rank}", losses=np.asarray(losses), **{"p_" + k: v for k, v in t.export_params().items()},
             **{"e_" + k: v for k, v in t.export_ema().items()}, **{"g_" + k: v for k, v in grads0.items()},
             **{"s_" + k: v for k, v in stats0.items()})
    # checkpoint under sharded Adam: every rank gathers the owners' m / v rows; continuing from the file is bit-identical
    path = out + ".ckpt"
    dp.save(path)
    sd = dp.state_dict()
    full_m = torch.from_numpy(sd["adam_m/W1"]).to(t.device)
    gathered = [torch.empty_like(full_m) for _ in range(world)]
    dist.all_gather(gathered, full_m)
    assert all(torch.equal(g, gathered[0]) for g in gathered), "gathered Adam state differs between ranks"
    dist.barrier()
    torch.cuda.synchronize()
    os._exit(0)


@pytest.mark.parametrize("comm", ["nccl", "nvlink"])
def test_two_gpu_syncbn_reproduces_the_single_process_reference(tmp_path, comm):
    """SyncBN (csrc/nvlink.cu syncbn_*_kernel): two replicas at B groups each == DSSMOracle on the re-stacked global
    batch of 2B groups -- the single-process reference graph, whose BN moments span the whole batch (new_dssm.py:77)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    from dssm_b200 import Config
    from dssm_b200.synthetic import init_params, make_batch
    from oracle import DSSMOracle
    from oracle.syncbn import restack_global
    from tests.helpers import assert_close, assert_update_close, oracle_config

    out = str(tmp_path / "r")
    ctx = mp.spawn(_syncbn_worker, args=(2, _free_port(), out, comm), nprocs=2, join=False)
    _join_all(ctx, 300, "syncbn workers")
    got = [np.load(out + f".{r}.npz") for r in range(2)]
    conf = Config(TRIGRAM_D=21128, query_BS=100, NEG=4, layers=(300, 300, 128))
    gconf = Config(TRIGRAM_D=21128, query_BS=200, NEG=4, layers=(300, 300, 128))
    params = init_params(conf, 0)
    orc = DSSMOracle(oracle_config(gconf), params)
    for s in range(2):
        shards = [make_batch(conf, seed=10 * s + r, lam_query=12, lam_doc=24).to_scipy() for r in range(2)]
        X = restack_global(shards, conf.query_BS, conf.NEG)
        cache = orc.forward(X, on_train=True)
        if s == 0:
            # global moments on every replica, bit-identical between replicas
            for l in (1, 2, 3):
                for nm, key in (("mean", "mean"), ("var", "var")):
                    want = np.stack([cache[f"bn{l}_q_{key}"], cache[f"bn{l}_d_{key}"]])
                    assert_close(got[0][f"s_bn{l}_{nm}"], want, 1e-5, f"global bn{l} {nm}")
                    assert np.array_equal(got[0][f"s_bn{l}_{nm}"], got[1][f"s_bn{l}_{nm}"])
            # the mean of the replicas' local losses is the global-batch loss
            assert abs(0.5 * (got[0]["losses"][0] + got[1]["losses"][0]) - float(cache["loss"])) <= 1e-5 * abs(float(cache["loss"]))
            grads = orc.backward(cache)
            for k in ("W2", "W3", "bn1_q_gamma", "bn2_d_beta", "bn3_d_gamma"):
                assert_close(got[0]["g_" + k], grads[k], 1e-4, f"syncbn grad {k}")
            orc.adam_update(grads)
        else:
            orc.adam_update(orc.backward(cache))
    for r in range(2):
        # measured maximum (DSSM_TEST_REPORT, round 2): 1.3e-2; tolerance = measured x 2.3
        assert_update_close({k: got[r]["p_" + k] for k in params}, orc.p, params, conf.use_bn, f"syncbn params rank {r}", l2_tol=0.03)
    for k in params:
        assert np.array_equal(got[0]["p_" + k], got[1]["p_" + k]), f"replicas differ in {k}"
    for k, v in orc.ema.items():
        assert np.array_equal(got[0]["e_" + k], got[1]["e_" + k])
        if k.endswith("ema_var"):
            assert_close(got[0]["e_" + k], v, 2e-2, f"syncbn ema {k}")


def _short_shard_worker(rank, world, port, out):
    import torch.distributed as dist

    from dssm_b200.retrieval import shard_range, sharded_corpus_topk

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    rng = np.random.default_rng(1)
    Q = np.maximum(rng.standard_normal((7, 128)), 0).astype(np.float32)
    res = {}
    for nd in (1, 37, 130):  # k = 100: rank 1's shard is empty / shorter than k / the corpus itself is shorter than k
        D = np.maximum(rng.standard_normal((nd, 128)), 0).astype(np.float32)
        lo, hi = shard_range(nd, rank, world)
        s, i = sharded_corpus_topk(torch.from_numpy(Q).cuda(), torch.from_numpy(D[lo:hi]).cuda().reshape(hi - lo, 128), 100, id_offset=lo)
        res[f"s{nd}"], res[f"i{nd}"] = s.cpu().numpy(), i.cpu().numpy()
    torch.cuda.synchronize()
    if rank == 1:
        np.savez(out, **res)
    dist.barrier()
    torch.cuda.synchronize()
    os._exit(0)


def test_two_gpu_sharded_retrieval_with_short_and_empty_shards(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    from oracle import corpus_topk_oracle

    out = str(tmp_path / "r.npz")
    ctx = mp.spawn(_short_shard_worker, args=(2, _free_port(), out), nprocs=2, join=False)
    _join_all(ctx, 240, "retrieval workers")
    got = np.load(out)
    rng = np.random.default_rng(1)
    Q = np.maximum(rng.standard_normal((7, 128)), 0).astype(np.float32)
    for nd in (1, 37, 130):
        D = np.maximum(rng.standard_normal((nd, 128)), 0).astype(np.float32)
        rs, ri = corpus_topk_oracle(Q, D, min(100, nd))
        assert got[f"i{nd}"].shape == (7, min(100, nd))
        assert np.array_equal(got[f"i{nd}"], ri) and np.array_equal(got[f"s{nd}"], rs)
