"""Whole-graph parity on the GPU: DSSMTower (one C call per step) against the committed golden fixtures and the
live oracle, on the reference variants (BN / no-BN, loss epsilon, un-normalised loss, tanh, 2 and 3 layers) and
the BASELINE configs.  fp32 tolerance 1e-5 relative to tensor scale on forward tensors, 5e-5 on gradients and on
parameters after Adam (fp32 summation order differs from NumPy's pairwise sums)."""
import numpy as np
import pytest
import torch

from tests.helpers import (GOLDEN_CASES, assert_close, assert_grads_close, assert_update_close, gpu_relu_masks,
                           grad_tolerances, load_golden,
                           oracle_config, rel_err, tf_adam_reference, to_stacked)

pytestmark = pytest.mark.gpu

FWD_TOL, GRAD_TOL = 1e-5, 5e-5


def check_adam_on_own_grads(tower, conf, p_before, m_before, v_before, bp_before, grad_scale=1.0):
    """Adam arithmetic pinned on the GPU's own gradients: host TF formula (float64) vs the kernel, element-wise."""
    g = tower.export_grads()
    p_after, m_after, v_after = tower.export_params(), tower._export("m"), tower._export("v")
    for k in g:
        rp, rm, rv = tf_adam_reference(p_before[k], g[k], m_before[k], v_before[k], bp_before[0], bp_before[1],
                                       lr=conf.learning_rate, grad_scale=grad_scale)
        assert_close(m_after[k], rm, 1e-6, f"adam m {k}")
        assert_close(v_after[k], rv, 1e-6, f"adam v {k}")
        # the quotient m/(sqrt(v)+eps) amplifies fp32 rounding of m,v where v is tiny: compare the step against lr
        assert np.abs(p_after[k] - rp).max() <= 2e-5 * conf.learning_rate + 1e-6 * np.abs(rp).max(), f"adam param {k}"


def one_step_checks(conf, tower, X, params):
    """forward tensors, gradients, EMA after the first update and the Adam arithmetic for one training step."""
    from oracle import DSSMOracle

    g64, allow, c64 = grad_tolerances(conf, X, params, GRAD_TOL)
    orc = DSSMOracle(oracle_config(conf), params)
    cache = orc.forward(X, on_train=True)
    assert np.isfinite(cache["loss"])
    x = tower.to_device(to_stacked(X))
    loss = tower.forward(x, on_train=True)
    n = len(conf.layers)
    for l in range(1, n + 1):
        assert_close(tower.tensor(f"h{l}").cpu().numpy(), cache[f"h{l}"], FWD_TOL, f"h{l}")
    q, pos, neg = orc.embeddings(cache)
    assert_close(tower.tensor("embedding_query_y").cpu().numpy(), q, FWD_TOL, "embedding_query_y")
    assert_close(tower.tensor("embedding_doc_positive_y").cpu().numpy(), pos, FWD_TOL, "embedding_doc_positive_y")
    assert_close(tower.tensor("embedding_doc_negative_y").cpu().numpy(), neg, FWD_TOL, "embedding_doc_negative_y")
    for k in ("cos_sim_raw", "query_norm_single", "doc_norm"):
        assert_close(tower.tensor(k).cpu().numpy().ravel(), cache[k], FWD_TOL, k)
    assert_close(tower.tensor("cos_sim").cpu().numpy(), cache["cos_sim"], FWD_TOL, "cos_sim")
    assert_close(tower.tensor("prob").cpu().numpy(), cache["prob"], FWD_TOL, "prob")
    assert abs(loss.item() - float(c64["loss"])) <= FWD_TOL * abs(float(c64["loss"]))
    assert abs(tower.tensor("accuracy").item() - float(cache["accuracy"])) < 1e-6
    assert_close(tower.tensor(f"dh{n}").cpu().numpy(), c64["dY"], GRAD_TOL, "dY")
    if conf.use_bn:  # shadows after the first update: 0.5 * batch statistics (identical parameters on both sides)
        for k, v in tower.export_ema().items():
            assert_close(v, orc.ema[k], FWD_TOL, f"ema {k}")
    tower.backward()
    assert_grads_close(conf, tower.export_grads(), g64, allow, c64)
    pb, mb, vb = tower.export_params(), tower._export("m"), tower._export("v")
    bp = tower.beta_pow.cpu().numpy().copy()
    tower.adam()
    check_adam_on_own_grads(tower, conf, pb, mb, vb, bp)
    orc.adam_update(orc.backward(cache))
    assert_update_close(tower.export_params(), orc.p, params, conf.use_bn, "after 1 step")
    return orc


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_tower_against_golden(name):
    from dssm_b200 import DSSMTower

    conf, Xs, params, z = load_golden(name)
    mx = max(X.nnz for X in Xs) + 8
    t = DSSMTower(conf, max_nnz=mx, params=params)
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
    g64, allow, c64 = grad_tolerances(conf, Xs[0], params, GRAD_TOL)
    assert_grads_close(conf, got, g64, allow, c64)
    for k in got:  # and directly against the committed fp32 values, at the looser of the two bounds
        if conf.use_bn and k[0] == "b" and k[1:].isdigit():
            continue
        ref = z[f"grad0/{k}"]
        assert np.abs(got[k] - ref).max() <= 2 * allow[k], f"golden grad {k}"
    # full training steps from the initial state: the loss trajectory is the well-conditioned observable
    t2 = DSSMTower(conf, max_nnz=mx, params=params)
    for s, X in enumerate(Xs):
        l = t2.train_step(t2.to_device(to_stacked(X)))
        assert abs(l.item() - float(z[f"loss{s}"])) <= 2e-4 * abs(float(z[f"loss{s}"])), f"loss step {s}"
        if s == 0:
            assert_update_close(t2.export_params(), {k: z[f"param1/{k}"] for k in params}, params, conf.use_bn, "after 1 step")
    last = len(Xs)
    assert_update_close(t2.export_params(), {k: z[f"param{last}/{k}"] for k in params}, params, conf.use_bn,
                        f"after {last} steps", l2_tol=2e-2)
    for k, v in t2.export_ema().items():
        if k.endswith("ema_var"):  # ema_mean carries the noise-driven pre-BN biases (see helpers.assert_update_close)
            assert_close(v, z[f"ema{last}/{k}"], 2e-3, f"ema {k}")
    # eval-mode forward with shadows produced by two training-mode forwards and NO parameter update (well-posed)
    from oracle import DSSMOracle

    t3 = DSSMTower(conf, max_nnz=mx, params=params)
    orc = DSSMOracle(oracle_config(conf), params)
    for X in Xs[:2]:
        t3.forward(t3.to_device(to_stacked(X)), on_train=True)
        orc.forward(X, on_train=True)
    ev = orc.forward(Xs[0], on_train=False)
    le = t3.forward(t3.to_device(to_stacked(Xs[0])), on_train=False)
    B = conf.query_BS
    assert_close(t3.tensor("BN2/embedding_query_y:0").cpu().numpy(), ev["Y"][:B], FWD_TOL, "eval embedding_query_y")
    assert_close(t3.tensor("embedding_doc_negative_y").cpu().numpy(), ev["Y"][2 * B:], FWD_TOL, "eval embedding_doc_negative_y")
    assert_close(t3.tensor("query_norm_single").cpu().numpy().ravel(), ev["query_norm_single"], FWD_TOL, "eval query_norm_single")
    assert abs(le.item() - float(ev["loss"])) <= 2e-5 * abs(float(ev["loss"]))
    for k, v in t3.export_ema().items():
        assert_close(v, orc.ema[k], FWD_TOL, f"ema {k} (no Adam)")


@pytest.mark.parametrize("gemm_mode", ["fp32", "tc_3xtf32"])
def test_golden_one_step_details(gemm_mode):
    from dssm_b200 import DSSMTower

    for name in GOLDEN_CASES:
        conf, Xs, params, _ = load_golden(name)
        conf.gemm_mode = gemm_mode
        t = DSSMTower(conf, max_nnz=Xs[0].nnz + 8, params=params)
        one_step_checks(conf, t, Xs[0], params)


CONFIGS = {
    "C1": dict(TRIGRAM_D=21128, query_BS=100, NEG=4, layers=(300, 300, 128)),
    "ref2layer": dict(TRIGRAM_D=6231, query_BS=100, NEG=4, layers=(400, 120)),  # archive/dssm_v2.py:28-36
    "C2_small": dict(TRIGRAM_D=49284, query_BS=256, NEG=4, layers=(300, 300, 128)),
    "C4_small_bn": dict(TRIGRAM_D=49284, query_BS=64, NEG=50, layers=(300, 300, 128)),
    "C4_small_nobn": dict(TRIGRAM_D=49284, query_BS=64, NEG=50, layers=(300, 300, 128), use_bn=False, loss_eps=1e-8),
    "tanh_sumloss": dict(TRIGRAM_D=5000, query_BS=37, NEG=3, layers=(64, 32), act="tanh", loss_div_bs=False),
}


@pytest.mark.parametrize("gemm_mode", ["fp32", "tc_3xtf32"])
@pytest.mark.parametrize("name", list(CONFIGS))
def test_tower_against_live_oracle(name, gemm_mode):
    from dssm_b200 import Config, DSSMTower
    from dssm_b200.synthetic import init_params, lambdas_for, make_batch

    conf = Config(gemm_mode=gemm_mode, **CONFIGS[name])
    lq, ld = lambdas_for(conf)
    vm = "tfidf" if name == "C4_small_nobn" else "count"
    batches = [make_batch(conf, seed=s, lam_query=lq, lam_doc=ld, value_mode=vm) for s in range(2)]
    params = init_params(conf, 0)
    t = DSSMTower(conf, max_nnz=max(b.nnz for b in batches), params=params)
    orc = one_step_checks(conf, t, batches[0].to_scipy(), params)
    # second step on both sides: loss is the well-conditioned observable, parameters in relative L2 of the update
    lo = orc.train_step(batches[1].to_scipy())
    lg = t.train_step(t.to_device(batches[1])).item()
    assert abs(lg - lo) <= 2e-3 * abs(lo)  # one violent Adam step (lr*sign(g) on every touched weight) amplifies noise
    # measured maximum over the parametrisations (DSSM_TEST_REPORT, round 2): 0.103 -- on a BN beta vector, where a handful of
    # noise-gradient entries whose first Adam step is +-lr weigh most; W1..W3 stay below 2e-2.  Tolerance = measured + 25 %.
    assert_update_close(t.export_params(), orc.p, params, conf.use_bn, "after 2 steps", l2_tol=0.13)
    for k, v in t.export_ema().items():
        if k.endswith("ema_var"):
            assert_close(v, orc.ema[k], 2e-2, f"ema {k}")  # after a violent first Adam step (see above)


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
        # same kernels, same fixed summation orders (row-ordered CSC columns, chunk-ordered BN / split-K merges): the three
        # paths are BIT-identical
        assert la == lb == lc, (la, lb, lc)
    pa, pb, pc = a.export_params(), b_.export_params(), c.export_params()
    for k in pa:
        assert np.array_equal(pa[k], pb[k]), f"host path differs in {k}"
        assert np.array_equal(pa[k], pc[k]), f"graph path differs in {k}"
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
    assert_update_close({"W2": d.export_params()["W2"]}, {"W2": e.export_params()["W2"]}, params, conf.use_bn, "run('train_step')")


def test_pipelined_host_feed_matches_synchronous_host_steps():
    """dssm_tower_train_step_host_async (upload of step k+1 on the copy stream under step k, losses read one step late)
    trains exactly like the synchronous host path: same losses step by step, same parameters; a batch reusing the pinned
    buffers of step k-2 is safe; a malformed batch is rejected before anything is enqueued."""
    from dssm_b200 import Config, DSSMTower
    from dssm_b200._lib import DssmError
    from dssm_b200.synthetic import init_params, make_batch

    conf = Config(TRIGRAM_D=21128, query_BS=100, NEG=4, layers=(300, 300, 128))
    batches = [make_batch(conf, seed=s, lam_query=12, lam_doc=24) for s in range(3)]
    params = init_params(conf, 0)
    mx = max(b.nnz for b in batches)
    a, b_ = DSSMTower(conf, max_nnz=mx, params=params), DSSMTower(conf, max_nnz=mx, params=params)
    a.capture_graph()
    b_.capture_graph()
    seq = [batches[i % 3] for i in range(7)]
    ref = [a.train_step_host(a.pin(bt)) for bt in seq]
    pinned = [b_.pin(bt) for bt in batches]  # 3 pinned buffers cycled over 7 steps
    got = b_.train_epoch_host([pinned[i % 3] for i in range(7)])
    assert len(got) == len(ref)
    for i, (x, y) in enumerate(zip(ref, got)):
        assert x == y, f"step {i}: {x} vs {y}"
    pa, pb = a.export_params(), b_.export_params()
    for k in pa:
        assert np.array_equal(pa[k], pb[k]), f"pipelined feed differs in {k}"
    with pytest.raises(DssmError):
        b_.feed_wait(0)  # only the last two steps are waitable
    ip, ix, vl, nnz = pinned[0]
    with pytest.raises(DssmError):
        b_.train_step_host_async((ip, ix, vl, nnz - 1))  # indptr[R] != nnz


def test_host_batch_loader_drives_the_pipelined_feed():
    """An epoch through HostBatchLoader + train_epoch_host (worker-thread batch assembly, pinned ring, double-buffered
    upload) trains like the reference-shaped loop over pull_batch feeds (new_dssm.py:261-269)."""
    import scipy.sparse as sp

    from dssm_b200 import Config, DSSMTower, HostBatchLoader, pull_batch, stack_feed
    from dssm_b200.synthetic import init_params, sparse_rows

    conf = Config(TRIGRAM_D=2000, query_BS=32, NEG=3, layers=(64, 32))
    rng = np.random.default_rng(5)
    nq = 32 * 7
    mk = lambda n, lam: sp.csr_matrix(sparse_rows(rng, n, conf.TRIGRAM_D, lam), shape=(n, conf.TRIGRAM_D))
    q, p, n = mk(nq, 8), mk(nq, 16), mk(nq * conf.NEG, 16)
    loader = HostBatchLoader(q, p, n, conf.query_BS, conf.NEG)
    assert len(loader) == 6
    params = init_params(conf, 0)
    a, b_ = DSSMTower(conf, max_nnz=loader.max_nnz, params=params), DSSMTower(conf, max_nnz=loader.max_nnz, params=params)
    b_.capture_graph()
    ref = [a.train_step(a.to_device(stack_feed(pull_batch(True, q, p, n, i, conf.query_BS, conf=conf), conf))).item()
           for i in range(len(loader))]
    got = b_.train_epoch_host(loader)
    assert len(got) == len(ref)
    for i, (x, y) in enumerate(zip(ref, got)):
        assert x == y, f"step {i}: {x} vs {y}"


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


def test_checkpoint_roundtrip(tmp_path):
    """state_dict -> load_state_dict, and save(path) -> restore(path) / from_checkpoint(path) on disk with the vocabulary
    (new_dssm.py:248,331; utils/utils.py:241-261): training continues BIT-identically from the restored state."""
    from dssm_b200 import Config, DSSMTower
    from dssm_b200.synthetic import make_batch

    conf = Config(TRIGRAM_D=2000, query_BS=16, NEG=3, layers=(32, 16))
    b = [make_batch(conf, s, 6, 10) for s in range(3)]
    mx = max(x.nnz for x in b)
    t = DSSMTower(conf, max_nnz=mx, seed=1)
    t.train_step(t.to_device(b[0]))
    sd = t.state_dict()
    vocab = {f"tok{i}": i for i in range(conf.TRIGRAM_D)}
    path = t.save(str(tmp_path / "model_1.ckpt"), vocabulary=vocab)
    assert path.endswith(".npz")
    u = DSSMTower(conf, max_nnz=mx, seed=2)
    u.load_state_dict(sd)
    w = DSSMTower(conf, max_nnz=mx, seed=3)
    assert w.restore(path) == vocab
    x_ = DSSMTower.from_checkpoint(path, max_nnz=mx)
    assert x_.vocabulary == vocab and tuple(x_.conf.layers) == (32, 16)
    for x in b[1:]:
        lt = t.train_step(t.to_device(x)).item()
        for other in (u, w, x_):
            assert other.train_step(other.to_device(x)).item() == lt
    pt = t.export_params()
    for other in (u, w, x_):
        po = other.export_params()
        for k in pt:
            assert np.array_equal(pt[k], po[k]), k
        assert torch.equal(other.m, t.m) and torch.equal(other.v, t.v) and torch.equal(other.ema, t.ema)
        assert torch.equal(other.beta_pow, t.beta_pow)
    bad = DSSMTower(Config(TRIGRAM_D=2000, query_BS=16, NEG=3, layers=(32, 8)), max_nnz=mx)
    with pytest.raises(ValueError):
        bad.restore(path)


def test_train_step_is_bit_reproducible():
    """Two towers fed the same batches end on the same bits (the reference on TF-CPU is deterministic too): the dW1
    gather sums every column in row order (csc_sort_*), all other reductions merge in a fixed order."""
    from dssm_b200 import DSSMTower, baseline_config
    from dssm_b200.synthetic import init_params, make_batch

    for name, steps in (("C1", 3), ("C2", 2)):
        conf = baseline_config(name)
        batches = [make_batch(conf, seed=s) for s in range(steps)]
        params = init_params(conf, 0)
        runs = []
        for _ in range(2):
            t = DSSMTower(conf, max_nnz=max(b.nnz for b in batches), params=params)
            losses = [t.train_step(t.to_device(b)).item() for b in batches]
            runs.append((losses, t.params.clone(), t.m.clone(), t.v.clone()))
        assert runs[0][0] == runs[1][0]
        for a_, b_ in zip(runs[0][1:], runs[1][1:]):
            assert torch.equal(a_, b_)


@pytest.mark.parametrize("name", ["C1", "C2"])
def test_tf32_mode_tolerance(name):
    """DSSM_GEMM_TC_TF32 (one tf32 MMA per product) against the float64 oracle: the tolerance quoted in include/dssm_b200.h
    and DESIGN.md is measured HERE.  Forward tensors are compared relative to their scale."""
    from dssm_b200 import DSSMTower, baseline_config
    from dssm_b200.synthetic import init_params, lambdas_for, make_batch
    from oracle import DSSMOracle

    conf = baseline_config(name, "tc_tf32")
    lq, ld = lambdas_for(conf)
    b = make_batch(conf, seed=3, lam_query=lq, lam_doc=ld)
    params = init_params(conf, 0)
    t = DSSMTower(conf, max_nnz=b.nnz, params=params)
    loss = t.forward(t.to_device(b), on_train=True).item()
    o64 = DSSMOracle(oracle_config(conf), params, dtype=np.float64)
    c = o64.forward(b.to_scipy(), on_train=True, update_ema=False)
    e_loss = abs(loss - float(c["loss"])) / abs(float(c["loss"]))
    e_y = rel_err(t.tensor("Y").cpu().numpy(), c["Y"])
    e_cos = rel_err(t.tensor("cos_sim_raw").cpu().numpy().ravel(), c["cos_sim_raw"])
    print(f"tf32 single-pass {name}: loss rel {e_loss:.2e}, embeddings {e_y:.2e}, cosines {e_cos:.2e}")
    assert e_loss <= 5e-3 and e_y <= 5e-3 and e_cos <= 5e-3
    t.backward()
    g = t.export_grads()
    g64 = o64.backward(c)
    # gradients in relative L2: a tf32-sized perturbation of a pre-activation next to the relu kink flips that unit's
    # derivative, which moves single gradient entries by their full upstream value (max-norm errors of 2e-2 .. 2e-1 were
    # measured) while the tensor as a whole stays within a percent
    for k in ("W1", "W2", "W3"):
        d = np.linalg.norm((g[k].astype(np.float64) - g64[k]).ravel()) / np.linalg.norm(g64[k].ravel())
        print(f"  grad {k}: rel L2 {d:.2e}, max-norm {rel_err(g[k], g64[k]):.2e}")
        assert d <= 5e-2, k


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


def test_chunked_backward_equals_monolithic():
    """The data-parallel pipeline pieces (backward_begin + dW1 column chunks + ranged Adam) reproduce backward()+adam()
    on one GPU (world_size 1 semantics: no exchange), for several chunk counts."""
    from dssm_b200 import Config, DSSMTower
    from dssm_b200.parallel import DataParallelTower
    from dssm_b200.synthetic import init_params, make_batch

    conf = Config(TRIGRAM_D=21128, query_BS=100, NEG=4, layers=(300, 300, 128), gemm_mode="tc_3xtf32")
    b = make_batch(conf, 0, 12, 24)
    params = init_params(conf, 0)
    ref = DSSMTower(conf, max_nnz=b.nnz, params=params)
    ref.forward(ref.to_device(b), on_train=True)
    ref.backward()
    g_ref = ref.export_grads()
    ref.adam()
    for n in (1, 3, 7):
        t = DSSMTower(conf, max_nnz=b.nnz, params=params)
        t.forward(t.to_device(b), on_train=True)
        t.backward_begin()
        for k in range(n):
            t.backward_w1(k, n)
        g = t.export_grads()
        for k in g_ref:
            assert np.array_equal(g[k], g_ref[k]), f"chunked grad {k} (n={n})"
        dp = DataParallelTower(t, n_chunks=n)
        spans = [t.w1_chunk(k, n) for k in range(n)]
        assert spans[0][0] == 0 and sum(c for _, c in spans) == 21128 * 300
        for off, cnt in spans:
            t.adam_range(off, cnt, 1.0)
        t.adam_range(dp.w1_end, t.P - dp.w1_end, 1.0)
        t.adam_advance()
        # the chunked gather sums every column in the same (row) order as the monolithic one: bit-identical parameters
        pr, pt = ref.export_params(), t.export_params()
        for k in pr:
            assert np.array_equal(pr[k], pt[k]), f"chunked adam (n={n}) differs in {k}"
        assert torch.equal(t.beta_pow, ref.beta_pow)


def test_embed_docs_feeds_retrieval():
    """Eval-mode embeddings of arbitrary rows (the corpus builder) equal the oracle's eval forward, and the corpus
    matrix they produce goes through corpus_topk with the oracle's ids."""
    from dssm_b200 import Config, DSSMTower, corpus_topk
    from dssm_b200.export import embed_docs, embed_queries
    from dssm_b200.synthetic import init_params, make_batch, sparse_rows
    from oracle import DSSMOracle, corpus_topk_oracle

    conf = Config(TRIGRAM_D=5000, query_BS=16, NEG=3, layers=(64, 128), gemm_mode="tc_3xtf32")
    params = init_params(conf, 0)
    b = [make_batch(conf, s, 6, 12) for s in range(2)]
    t = DSSMTower(conf, max_nnz=4096, params=params)
    orc = DSSMOracle(oracle_config(conf), params)
    for x in b:  # populate the EMA shadows identically (forward only, no parameter update)
        t.forward(t.to_device(x), on_train=True)
        orc.forward(x.to_scipy(), on_train=True)
    rng = np.random.Generator(np.random.PCG64(5))
    docs = sparse_rows(rng, 150, conf.TRIGRAM_D, 12.0)  # 150 rows: not a multiple of the 64 doc slots per batch
    queries = sparse_rows(rng, 21, conf.TRIGRAM_D, 6.0)
    E = embed_docs(t, docs)
    Qe = embed_queries(t, queries)
    # oracle: eval forward of batches holding the same rows in the doc / query slots
    import scipy.sparse as sp

    def oracle_embed(rows, seg):
        out = []
        per = conf.query_BS if seg == "q" else 4 * conf.query_BS
        for lo in range(0, rows.shape[0], per):
            chunk = rows[lo:lo + per]
            pad = per - chunk.shape[0]
            if pad:
                chunk = sp.vstack([chunk] + [rows[:1]] * pad, format="csr")
            other = sp.vstack([rows[:1]] * (4 * conf.query_BS if seg == "q" else conf.query_BS), format="csr")
            X = sp.vstack([chunk, other] if seg == "q" else [other, chunk], format="csr")
            Y = orc.forward(X, on_train=False)["Y"]
            part = Y[:conf.query_BS] if seg == "q" else Y[conf.query_BS:]
            out.append(part[:per - pad])
        return np.concatenate(out)

    assert_close(E, oracle_embed(docs, "d"), FWD_TOL, "embed_docs")
    assert_close(Qe, oracle_embed(queries, "q"), FWD_TOL, "embed_queries")
    s, i = corpus_topk(torch.from_numpy(Qe).cuda(), torch.from_numpy(E).cuda(), 10)
    rs, ri = corpus_topk_oracle(Qe, E, 10)
    assert np.array_equal(i.cpu().numpy(), ri) and np.array_equal(s.cpu().numpy(), rs)


@pytest.mark.parametrize("name", ["C2", "C3", "C4", "C4_NOBN"])
def test_full_size_baseline_configs_against_oracle(name):
    """BASELINE.json configs at their full sizes (C2: B=1024, NEG=4; C3: B=8192 -- the per-GPU batch of the data-parallel
    config; C4: B=1024, NEG=50 with and without BN): one
    training step against the oracle -- loss and cosines at 1e-5, every gradient at the bounds of grad_tolerances,
    the oracle differentiating with the device's relu active set (helpers.align_relu_masks: the two sets may differ only
    at units within helpers.KINK_TOL of the kink)."""
    from dssm_b200 import DSSMTower, baseline_config
    from dssm_b200.synthetic import init_params, make_batch

    conf = baseline_config(name)
    vm = "tfidf" if name == "C4_NOBN" else "count"  # dssm_no_bn/dssm_tf_idf.py:37 feeds real-valued features
    b = make_batch(conf, seed=3, value_mode=vm)
    params = init_params(conf, 0)
    t = DSSMTower(conf, max_nnz=b.nnz, params=params)
    X = b.to_scipy()
    loss = t.forward(t.to_device(b), on_train=True)
    g64, allow, c64 = grad_tolerances(conf, X, params, GRAD_TOL, masks=gpu_relu_masks(conf, t))
    assert c64["kink_flips"] <= 1e-5 * sum(c64[f"a{l}"].size for l in range(1, len(conf.layers) + 1)) + 4
    assert np.isfinite(float(c64["loss"]))
    assert abs(loss.item() - float(c64["loss"])) <= FWD_TOL * abs(float(c64["loss"]))
    assert_close(t.tensor("cos_sim_raw").cpu().numpy().ravel(), c64["cos_sim_raw"], FWD_TOL, "cos_sim_raw")
    assert_close(t.tensor("Y").cpu().numpy(), c64["Y"], FWD_TOL, "Y")
    t.backward()
    assert_grads_close(conf, t.export_grads(), g64, allow, c64)


def test_fused_bn_epilogue_matches_separate_bn_kernels(monkeypatch):
    """DSSM_FUSED_BN=1: the tcgen05 GEMM epilogue takes the batch moments of its own output and its last CTAs finalize them
    (csrc/fc_tc.cu, FusedBnStats).  Same oracle bounds as the default path on C2 (B % 128 == 0 is the fused path's condition),
    and the two paths agree with each other to fp32 summation-order noise."""
    from dssm_b200 import DSSMTower, baseline_config
    from dssm_b200.synthetic import init_params, make_batch

    conf = baseline_config("C2")
    b = make_batch(conf, seed=5)
    params = init_params(conf, 0)
    X = b.to_scipy()
    ref = DSSMTower(conf, max_nnz=b.nnz, params=params)
    ref.forward(ref.to_device(b), on_train=True)
    stats_ref = {k: ref.tensor(k).clone() for k in ("bn2_mean", "bn2_var", "bn3_mean", "bn3_var", "bn3_scale", "bn3_shift")}
    n_ref = ref.launch_count
    monkeypatch.setenv("DSSM_FUSED_BN", "1")
    t = DSSMTower(conf, max_nnz=b.nnz, params=params)
    loss = t.forward(t.to_device(b), on_train=True)
    assert t.launch_count == n_ref - 2  # the bn_stats launches of layers 2 and 3 are gone
    for k, v in stats_ref.items():
        assert_close(t.tensor(k).cpu().numpy(), v.cpu().numpy(), 1e-5, f"fused {k}")
    for k, v in t.export_ema().items():
        assert_close(v, ref.export_ema()[k], 1e-5, f"fused ema {k}")
    g64, allow, c64 = grad_tolerances(conf, X, params, GRAD_TOL, masks=gpu_relu_masks(conf, t))
    assert abs(loss.item() - float(c64["loss"])) <= FWD_TOL * abs(float(c64["loss"]))
    assert_close(t.tensor("Y").cpu().numpy(), c64["Y"], FWD_TOL, "Y")
    t.backward()
    assert_grads_close(conf, t.export_grads(), g64, allow, c64)
