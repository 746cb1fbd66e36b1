"""Whole-graph parity on the GPU: DSSMTower (one C call per step) against the committed golden fixtures and the
live oracle, on the reference variants (BN / no-BN, loss epsilon, un-normalised loss, tanh, 2 and 3 layers) and
the BASELINE configs.  fp32 tolerance 1e-5 relative to tensor scale on forward tensors, 5e-5 on gradients and on
parameters after Adam (fp32 summation order differs from NumPy's pairwise sums)."""
import numpy as np
import pytest
import torch

from tests.helpers import GOLDEN_CASES, assert_close, load_golden, oracle_config, rel_err, to_stacked

pytestmark = pytest.mark.gpu

FWD_TOL, GRAD_TOL = 1e-5, 5e-5


def bias_grad_tol(conf, cache, l):
    """Bias gradients are analytically zero under BN; compare against the scale of what was summed."""
    return 1e-5 * float(np.abs(cache[f"dh{l}"]).sum(axis=0).max())


def compare_grads(conf, tower, grads, cache, tol=GRAD_TOL):
    got = tower.export_grads()
    for k, g in grads.items():
        if conf.use_bn and k[0] == "b" and k[1:].isdigit():
            assert np.abs(got[k] - g).max() <= bias_grad_tol(conf, cache, int(k[1:])), k
        else:
            assert_close(got[k], g, tol, f"grad {k}")


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_tower_against_golden(name):
    from dssm_b200 import DSSMTower

    conf, Xs, params, z = load_golden(name)
    t = DSSMTower(conf, max_nnz=max(X.nnz for X in Xs) + 8, params=params)
    x0 = t.to_device(to_stacked(Xs[0]))
    loss = t.forward(x0, on_train=True, update_ema=False)
    for l in range(1, len(conf.layers) + 1):
        assert_close(t.tensor(f"h{l}").cpu().numpy(), z[f"fwd/h{l}"], FWD_TOL, f"h{l}")
    assert_close(t.tensor("Y").cpu().numpy(), z["fwd/Y"], FWD_TOL, "Y")
    assert_close(t.tensor("cos_sim_raw").cpu().numpy().ravel(), z["fwd/cos_sim_raw"], FWD_TOL, "cos_sim_raw")
    assert_close(t.tensor("cos_sim").cpu().numpy(), z["fwd/cos_sim"], FWD_TOL, "cos_sim")
    assert_close(t.tensor("prob").cpu().numpy(), z["fwd/prob"], FWD_TOL, "prob")
    assert abs(loss.item() - float(z["loss0"])) <= FWD_TOL * abs(float(z["loss0"]))
    assert_close(t.tensor(f"dh{len(conf.layers)}").cpu().numpy(), z["fwd/dY"], GRAD_TOL, "dY")
    t.backward()
    got = t.export_grads()
    for k in got:
        ref = z[f"grad0/{k}"]
        if conf.use_bn and k[0] == "b" and k[1:].isdigit():
            assert np.abs(got[k] - ref).max() <= 1e-5 * max(np.abs(z["fwd/dY"]).sum(), 1e-6), k
        else:
            assert_close(got[k], ref, GRAD_TOL, f"grad {k}")
    # full training steps from the initial state
    t2 = DSSMTower(conf, max_nnz=max(X.nnz for X in Xs) + 8, params=params)
    for s, X in enumerate(Xs):
        l = t2.train_step(t2.to_device(to_stacked(X)))
        assert abs(l.item() - float(z[f"loss{s}"])) <= 1e-4 * abs(float(z[f"loss{s}"])), f"loss step {s}"
        if s == 0:
            p1 = t2.export_params()
            for k in p1:
                assert_close(p1[k], z[f"param1/{k}"], GRAD_TOL, f"param after 1 step {k}")
    last = len(Xs)
    pl = t2.export_params()
    for k in pl:
        assert_close(pl[k], z[f"param{last}/{k}"], 2e-4, f"param after {last} steps {k}")
    for k, v in t2.export_ema().items():
        assert_close(v, z[f"ema{last}/{k}"], 1e-4, f"ema {k}")
    # eval-mode forward (EMA statistics), embeddings by reference tensor name
    t2.forward(t2.to_device(to_stacked(Xs[0])), on_train=False)
    B = conf.query_BS
    assert_close(t2.tensor("BN2/embedding_query_y:0").cpu().numpy(), z["eval/Y"][:B], 2e-4, "eval embedding_query_y")
    assert_close(t2.tensor("embedding_doc_negative_y").cpu().numpy(), z["eval/Y"][2 * B:], 2e-4, "eval embedding_doc_negative_y")


CONFIGS = {
    "C1": dict(TRIGRAM_D=21128, query_BS=100, NEG=4, layers=(300, 300, 128)),
    "ref2layer": dict(TRIGRAM_D=6231, query_BS=100, NEG=4, layers=(400, 120)),  # archive/dssm_v2.py:28-36
    "C2_small": dict(TRIGRAM_D=49284, query_BS=256, NEG=4, layers=(300, 300, 128)),
    "C4_small_bn": dict(TRIGRAM_D=49284, query_BS=64, NEG=50, layers=(300, 300, 128)),
    "C4_small_nobn": dict(TRIGRAM_D=49284, query_BS=64, NEG=50, layers=(300, 300, 128), use_bn=False, loss_eps=1e-8),
    "tanh_sumloss": dict(TRIGRAM_D=5000, query_BS=37, NEG=3, layers=(64, 32), act="tanh", loss_div_bs=False),
}


@pytest.mark.parametrize("name", list(CONFIGS))
def test_tower_against_live_oracle(name):
    from dssm_b200 import Config, DSSMTower
    from dssm_b200.synthetic import init_params, lambdas_for, make_batch
    from oracle import DSSMOracle

    conf = Config(**CONFIGS[name])
    lq, ld = lambdas_for(conf)
    vm = "tfidf" if name == "C4_small_nobn" else "count"
    batches = [make_batch(conf, seed=s, lam_query=lq, lam_doc=ld, value_mode=vm) for s in range(2)]
    params = init_params(conf, 0)
    orc = DSSMOracle(oracle_config(conf), params)
    t = DSSMTower(conf, max_nnz=max(b.nnz for b in batches), params=params)
    cache = orc.forward(batches[0].to_scipy(), on_train=True)
    grads = orc.backward(cache)
    assert np.isfinite(cache["loss"])
    loss = t.forward(t.to_device(batches[0]), on_train=True)
    t.backward()
    for l in range(1, len(conf.layers) + 1):
        assert_close(t.tensor(f"h{l}").cpu().numpy(), cache[f"h{l}"], FWD_TOL, f"h{l}")
    B = conf.query_BS
    q, pos, neg = orc.embeddings(cache)
    assert_close(t.tensor("embedding_query_y").cpu().numpy(), q, FWD_TOL, "embedding_query_y")
    assert_close(t.tensor("embedding_doc_positive_y").cpu().numpy(), pos, FWD_TOL, "embedding_doc_positive_y")
    assert_close(t.tensor("embedding_doc_negative_y").cpu().numpy(), neg, FWD_TOL, "embedding_doc_negative_y")
    assert_close(t.tensor("cos_sim_raw").cpu().numpy().ravel(), cache["cos_sim_raw"], FWD_TOL, "cos_sim_raw")
    assert_close(t.tensor("query_norm_single").cpu().numpy().ravel(), cache["query_norm_single"], FWD_TOL, "query_norm_single")
    assert abs(loss.item() - float(cache["loss"])) <= FWD_TOL * abs(float(cache["loss"]))
    assert abs(t.tensor("accuracy").item() - float(cache["accuracy"])) < 1e-6
    compare_grads(conf, t, grads, cache)
    # two full steps (Adam + EMA) from scratch on both sides
    orc2 = DSSMOracle(oracle_config(conf), params)
    t2 = DSSMTower(conf, max_nnz=max(b.nnz for b in batches), params=params)
    for b in batches:
        lo = orc2.train_step(b.to_scipy())
        lg = t2.train_step(t2.to_device(b)).item()
        assert abs(lg - lo) <= 1e-4 * abs(lo)
    got = t2.export_params()
    for k, v in orc2.p.items():
        assert_close(got[k], v, 2e-4, f"param {k} after 2 steps")
    for k, v in t2.export_ema().items():
        assert_close(v, orc2.ema[k], 1e-4, f"ema {k}")


def test_host_step_graph_and_sess_run_shim():
    """train_step_host (H2D + step + D2H, CUDA-graph replay) == train_step on device buffers; tower.run keeps the
    sess.run(fetch, feed_dict=pull_batch(...)) call shape of new_dssm.py:267-285."""
    from dssm_b200 import Config, DSSMTower, pull_batch
    from dssm_b200.synthetic import init_params, make_batch

    conf = Config(TRIGRAM_D=21128, query_BS=100, NEG=4, layers=(300, 300, 128))
    batches = [make_batch(conf, seed=s, lam_query=12, lam_doc=24) for s in range(3)]
    params = init_params(conf, 0)
    mx = max(b.nnz for b in batches)
    a, b_, c = (DSSMTower(conf, max_nnz=mx, params=params) for _ in range(3))
    c.capture_graph()
    for bt in batches:
        la = a.train_step(a.to_device(bt)).item()
        lb = b_.train_step_host(b_.pin(bt))
        lc = c.train_step_host(c.pin(bt))
        assert abs(la - lb) <= 1e-5 * abs(la) and abs(la - lc) <= 1e-5 * abs(la)
    pa, pb, pc = a.export_params(), b_.export_params(), c.export_params()
    for k in pa:
        assert_close(pb[k], pa[k], 1e-5, f"host path {k}")
        assert_close(pc[k], pa[k], 1e-5, f"graph path {k}")
    assert c.launch_count > 0 and a.launch_count > 0
    # sess.run shim on reference-style feeds
    X = batches[0].to_scipy()
    B, N = conf.query_BS, conf.NEG
    feed = pull_batch(False, X[:B], X[B:2 * B], X[2 * B:], 0, B, conf=conf)
    d = DSSMTower(conf, max_nnz=mx, params=params)
    loss, qy = d.run(["loss", "BN2/embedding_query_y:0"], feed)
    assert qy.shape == (B, 128) and np.isfinite(loss).all()
    feed_t = pull_batch(True, X[:B], X[B:2 * B], X[2 * B:], 0, B, conf=conf)
    d.run("train_step", feed_t)
    e = DSSMTower(conf, max_nnz=mx, params=params)
    e.train_step(e.to_device(batches[0]))
    assert_close(d.export_params()["W2"], e.export_params()["W2"], 1e-5, "run('train_step')")


def test_wrong_batch_shape_is_rejected():
    from dssm_b200 import Config, DSSMTower
    from dssm_b200.synthetic import make_batch

    conf = Config(TRIGRAM_D=500, query_BS=8, NEG=2, layers=(16, 8))
    t = DSSMTower(conf, max_nnz=4096)
    other = Config(TRIGRAM_D=500, query_BS=7, NEG=2, layers=(16, 8))
    with pytest.raises(ValueError):
        t.train_step(t.to_device(make_batch(other, 0)))
    with pytest.raises(Exception):
        t.backward()  # no preceding training forward


def test_checkpoint_roundtrip():
    from dssm_b200 import Config, DSSMTower
    from dssm_b200.synthetic import make_batch

    conf = Config(TRIGRAM_D=2000, query_BS=16, NEG=3, layers=(32, 16))
    b = [make_batch(conf, s, 6, 10) for s in range(3)]
    mx = max(x.nnz for x in b)
    t = DSSMTower(conf, max_nnz=mx, seed=1)
    t.train_step(t.to_device(b[0]))
    sd = t.state_dict()
    u = DSSMTower(conf, max_nnz=mx, seed=2)
    u.load_state_dict(sd)
    for x in b[1:]:
        lt, lu = t.train_step(t.to_device(x)).item(), u.train_step(u.to_device(x)).item()
        assert abs(lt - lu) <= 1e-5 * abs(lt)
    pt, pu = t.export_params(), u.export_params()
    for k in pt:
        assert_close(pu[k], pt[k], 1e-5, k)


def test_full_size_c2_properties():
    """BASELINE C2 shape (B=1024): size-independent checks -- loss falls over steps, everything stays finite, the
    dense-Adam contract holds (rows of W1 absent from the batch keep moving once they have momentum), and the EMA
    shadows equal the batch statistics' recursion."""
    from dssm_b200 import DSSMTower, baseline_config
    from dssm_b200.synthetic import make_batch

    conf = baseline_config("C2")
    batches = [make_batch(conf, s) for s in range(2)]
    t = DSSMTower(conf, max_nnz=max(b.nnz for b in batches))
    x0, x1 = (t.to_device(b) for b in batches)
    losses = [t.train_step(x0).item() for _ in range(4)]
    assert all(np.isfinite(losses)) and losses[-1] < losses[0]
    w_before = t.param("W1").clone()
    t.train_step(x1)
    w_after = t.param("W1")
    only0 = np.setdiff1d(np.unique(batches[0].indices), np.unique(batches[1].indices))
    absent = np.setdiff1d(np.arange(conf.TRIGRAM_D), np.union1d(batches[0].indices, batches[1].indices))
    assert only0.size and (w_after[only0] != w_before[only0]).any(dim=1).all()  # momentum moves them with zero grad
    assert absent.size and torch.equal(w_after[absent], w_before[absent])  # never touched: zero state, zero update
    assert torch.isfinite(t.params).all() and torch.isfinite(t.ema).all()
