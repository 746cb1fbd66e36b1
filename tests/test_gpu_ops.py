"""Per-kernel parity on the GPU, through the C ABI (dssm_b200.ops -> libdssm_b200.so), against NumPy/SciPy/torch-fp64
restatements and the oracle.  Tolerance: 1e-5 relative to the tensor's scale (north_star) unless stated; index
outputs are bit-exact."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

from tests.helpers import assert_close, random_csr, rel_err, to_stacked

pytestmark = pytest.mark.gpu

TOL = 1e-5


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def csr_dev(X):
    from dssm_b200.ops import DeviceCSR

    return DeviceCSR.from_host(to_stacked(X), "cuda")


# ---- FC1 forward -------------------------------------------------------------------------------------------
@pytest.mark.parametrize("R,D,L1,mode", [(600, 21128, 300, "count"), (257, 1000, 128, "tfidf"), (33, 50, 8, "count"),
                                          (64, 200, 7, "count"), (40, 300, 516, "tfidf"), (5, 64, 1028, "count")])
def test_spmm_fwd(R, D, L1, mode):
    from dssm_b200 import ops

    rng = np.random.default_rng(R + L1)
    X = random_csr(rng, R, D, max_nnz_row=40, value_mode=mode, allow_empty=True)
    W = rng.uniform(-0.1, 0.1, (D, L1)).astype(np.float32)
    b = rng.uniform(-0.1, 0.1, (L1,)).astype(np.float32)
    want = np.asarray(X @ W, dtype=np.float32) + b
    got = ops.spmm_fwd(csr_dev(X), dev(W), dev(b)).cpu().numpy()
    assert_close(got, want, TOL, "spmm_fwd")
    got_nb = ops.spmm_fwd(csr_dev(X), dev(W), None).cpu().numpy()
    assert_close(got_nb, np.asarray(X @ W, dtype=np.float32), TOL, "spmm_fwd no bias")
    empty = np.flatnonzero(np.diff(X.indptr) == 0)
    if empty.size:
        assert np.array_equal(got[empty], np.broadcast_to(b, (empty.size, L1)))  # empty row = bias exactly


def test_spmm_fwd_linearity_full_size():
    """Size-independent property at the C2 shape: spmm(X, W+V) == spmm(X, W) + spmm(X, V)."""
    from dssm_b200 import Config, ops
    from dssm_b200.synthetic import make_batch

    conf = Config(TRIGRAM_D=49284, query_BS=1024, NEG=4, layers=(300, 300, 128))
    x = ops.DeviceCSR.from_host(make_batch(conf, 0), "cuda")
    g = torch.Generator(device="cpu").manual_seed(0)
    W = (torch.rand((49284, 300), generator=g) - 0.5).cuda()
    V = (torch.rand((49284, 300), generator=g) - 0.5).cuda()
    a = ops.spmm_fwd(x, W + V, None)
    b = ops.spmm_fwd(x, W, None) + ops.spmm_fwd(x, V, None)
    assert rel_err(a.cpu().numpy(), b.cpu().numpy()) < 1e-5
    # and against torch's own sparse matmul
    Xt = torch.sparse_csr_tensor(x.indptr.long(), x.indices.long(), x.values, size=(conf.rows, 49284))
    ref = torch.sparse.mm(Xt.double(), W.double()).float()
    assert rel_err(ops.spmm_fwd(x, W, None).cpu().numpy(), ref.cpu().numpy()) < 1e-5


# ---- FC1 backward ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("method", [0, 1])
@pytest.mark.parametrize("R,D,L1", [(600, 2000, 300), (1500, 400, 128), (33, 50, 8), (64, 200, 7)])
def test_spmm_bwd_dw(method, R, D, L1):
    from dssm_b200 import ops

    rng = np.random.default_rng(R + L1 + method)
    X = random_csr(rng, R, D, max_nnz_row=30, value_mode="tfidf", allow_empty=True).tolil()
    X[:, 3] = rng.random((R, 1)).astype(np.float32) + 0.1  # a hot column present in every row (multi-chunk path)
    X = sp.csr_matrix(X, dtype=np.float32)
    X.sort_indices()
    dH = rng.standard_normal((R, L1)).astype(np.float32)
    want = np.asarray(X.T.astype(np.float64) @ dH.astype(np.float64))
    got = ops.spmm_bwd_dw(csr_dev(X), dev(dH), D, method=method).cpu().numpy()
    assert_close(got, want, TOL, f"spmm_bwd_dw method {method}")
    absent = np.setdiff1d(np.arange(D), np.unique(X.indices))
    if absent.size:
        assert not got[absent].any()  # rows of absent columns are exactly zero (dense TF gradient)


def test_spmm_bwd_methods_agree_full_size():
    from dssm_b200 import Config, ops
    from dssm_b200.synthetic import make_batch

    conf = Config(TRIGRAM_D=49284, query_BS=1024, NEG=4, layers=(300, 300, 128))
    x = ops.DeviceCSR.from_host(make_batch(conf, 1), "cuda")
    dH = torch.randn((conf.rows, 300), device="cuda")
    a = ops.spmm_bwd_dw(x, dH, 49284, method=0)
    b = ops.spmm_bwd_dw(x, dH, 49284, method=1)
    assert rel_err(a.cpu().numpy(), b.cpu().numpy()) < 1e-5
    # checksum of checksums: sum over all of dW1 == sum_p value_p * rowsum(dH[row_p])
    rows = torch.repeat_interleave(torch.arange(conf.rows, device="cuda"), torch.diff(x.indptr.long()))
    want = (x.values.double() * dH.double().sum(1)[rows]).sum().item()
    assert abs(a.double().sum().item() - want) < 1e-6 * (x.values.double().abs() * dH.double().abs().sum(1)[rows]).sum().item()


# ---- batch_normalization -----------------------------------------------------------------------------------
def np_bn_stats(x):
    m = x.astype(np.float64).mean(0)
    return m, ((x.astype(np.float64) - m) ** 2).mean(0)


@pytest.mark.parametrize("R,B,L", [(600, 100, 300), (1030, 257, 128), (12, 5, 7), (300, 300, 64)])
def test_bn_forward_train_eval_and_ema(R, B, L):
    from dssm_b200 import ops

    rng = np.random.default_rng(R + L)
    x = (rng.standard_normal((R, L)) * 3 + 5).astype(np.float32)  # large mean: punishes E[x^2]-E[x]^2
    st = ops.BNState(L, "cuda")
    st.gamma.copy_(dev(rng.uniform(0.5, 1.5, (2, L)).astype(np.float32)))
    st.beta.copy_(dev(rng.uniform(-0.5, 0.5, (2, L)).astype(np.float32)))
    ops.bn_forward(dev(x), B, st, on_train=True)
    segs = [(0, slice(0, B))] + ([(1, slice(B, R))] if B < R else [])
    for s, sl in segs:
        m, v = np_bn_stats(x[sl])
        assert_close(st.mean[s].cpu().numpy(), m, TOL, "mean")
        assert_close(st.var[s].cpu().numpy(), v, TOL, "var")
        assert_close(st.ema_mean[s].cpu().numpy(), 0.5 * m, TOL, "ema_mean")
        assert_close(st.ema_var[s].cpu().numpy(), 0.5 * v, TOL, "ema_var")
        inv = 1 / np.sqrt(v + 1e-3) * st.gamma[s].cpu().numpy()
        assert_close(st.scale[s].cpu().numpy(), inv, TOL, "scale")
        assert_close(st.shift[s].cpu().numpy(), st.beta[s].cpu().numpy() - m * inv, TOL, "shift")
    y = ops.bn_act_apply(dev(x), st.scale, st.shift, "relu", B).cpu().numpy()
    for s, sl in segs:
        m, v = np_bn_stats(x[sl])
        want = np.maximum((x[sl] - m) / np.sqrt(v + 1e-3) * st.gamma[s].cpu().numpy() + st.beta[s].cpu().numpy(), 0)
        assert_close(y[sl], want, TOL, "bn+relu")
    # second training pass: shadows move again; eval pass: shadows are used and do not move
    ops.bn_forward(dev(x), B, st, on_train=True)
    m, _ = np_bn_stats(x[:B])
    assert_close(st.ema_mean[0].cpu().numpy(), 0.75 * m, TOL, "ema second step")
    before = st.ema_mean.clone()
    ops.bn_forward(dev(x * 2), B, st, on_train=False)
    assert torch.equal(before, st.ema_mean)
    assert_close(st.mean[0].cpu().numpy(), before[0].cpu().numpy(), 1e-7, "eval uses shadows")


def test_batch_normalization_reference_call_shape():
    """normed = batch_normalization(x, on_train, out_size) on a single tensor, as new_dssm.py:129 calls it."""
    from dssm_b200 import batch_normalization

    x = np.random.default_rng(0).standard_normal((100, 32)).astype(np.float32)
    normed, st = batch_normalization(dev(x), True, 32)
    m, v = np_bn_stats(x)
    assert_close(normed.cpu().numpy(), (x - m) / np.sqrt(v + 1e-3), TOL, "normed")


@pytest.mark.parametrize("act", ["relu", "tanh"])
@pytest.mark.parametrize("R,B,L", [(600, 100, 300), (70, 13, 9)])
def test_bn_act_backward_vs_autograd(act, R, B, L):
    from dssm_b200 import ops

    rng = np.random.default_rng(L)
    h = rng.standard_normal((R, L)).astype(np.float32)
    dA = rng.standard_normal((R, L)).astype(np.float32)
    gam = rng.uniform(0.5, 1.5, (2, L)).astype(np.float32)
    bet = rng.uniform(-0.5, 0.5, (2, L)).astype(np.float32)
    ht = torch.tensor(h, dtype=torch.float64, requires_grad=True)
    gt = torch.tensor(gam, dtype=torch.float64, requires_grad=True)
    bt = torch.tensor(bet, dtype=torch.float64, requires_grad=True)
    outs = []
    for s, sl in ((0, slice(0, B)), (1, slice(B, R))):
        x = ht[sl]
        m = x.mean(0)
        v = ((x - m) ** 2).mean(0)
        y = (x - m) / torch.sqrt(v + 1e-3) * gt[s] + bt[s]
        outs.append(torch.relu(y) if act == "relu" else torch.tanh(y))
    (torch.cat(outs) * torch.tensor(dA, dtype=torch.float64)).sum().backward()
    st = ops.BNState(L, "cuda")
    st.gamma.copy_(dev(gam))
    st.beta.copy_(dev(bet))
    ops.bn_forward(dev(h), B, st, on_train=True)
    d = dev(dA).clone()
    dgamma, dbeta = ops.bn_act_backward(d, dev(h), B, act, st)
    assert_close(d.cpu().numpy(), ht.grad.numpy(), 2e-5, "dH")
    assert_close(dgamma.cpu().numpy(), gt.grad.numpy(), 2e-5, "dgamma")
    assert_close(dbeta.cpu().numpy(), bt.grad.numpy(), 2e-5, "dbeta")
    # no-BN mode: activation derivative only
    d2 = dev(dA).clone()
    ops.bn_act_backward(d2, dev(h), B, act, None)
    a = np.maximum(h, 0) if act == "relu" else np.tanh(h)
    want = dA * ((a > 0) if act == "relu" else (1 - a * a))
    assert_close(d2.cpu().numpy(), want, TOL, "act backward")


# ---- dense layers ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("R,K,N,B", [(600, 300, 300, 100), (6144, 300, 128, 1024), (77, 13, 7, 11), (130, 64, 65, 130)])
@pytest.mark.parametrize("with_bn", [True, False])
def test_fc_fwd_bwd(R, K, N, B, with_bn):
    from dssm_b200 import ops

    rng = np.random.default_rng(R + K + N)
    h = rng.standard_normal((R, K)).astype(np.float32)
    W = rng.uniform(-0.1, 0.1, (K, N)).astype(np.float32)
    b = rng.uniform(-0.1, 0.1, (N,)).astype(np.float32)
    scale = rng.uniform(0.5, 1.5, (2, K)).astype(np.float32) if with_bn else None
    shift = rng.uniform(-0.5, 0.5, (2, K)).astype(np.float32) if with_bn else None
    a = h.astype(np.float64)
    if with_bn:
        a = np.concatenate([a[:B] * scale[0] + shift[0], a[B:] * scale[1] + shift[1]])
    a = np.maximum(a, 0)
    want = a @ W.astype(np.float64) + b
    got = ops.fc_fwd(dev(h), dev(W), dev(b), dev(scale) if with_bn else None, dev(shift) if with_bn else None, "relu", B)
    assert_close(got.cpu().numpy(), want, TOL, "fc_fwd")
    dH = rng.standard_normal((R, N)).astype(np.float32)
    dA = ops.fc_bwd_dx(dev(dH), dev(W)).cpu().numpy()
    assert_close(dA, dH.astype(np.float64) @ W.T.astype(np.float64), TOL, "fc_bwd_dx")
    dW, db = ops.fc_bwd_dw(dev(h), dev(dH), dev(scale) if with_bn else None, dev(shift) if with_bn else None, "relu", B)
    assert_close(dW.cpu().numpy(), a.T @ dH.astype(np.float64), TOL, "fc_bwd_dw")
    assert_close(db.cpu().numpy(), dH.astype(np.float64).sum(0), TOL, "db")
    assert_close(ops.colsum(dev(dH)).cpu().numpy(), dH.astype(np.float64).sum(0), TOL, "colsum")


@pytest.mark.parametrize("R,K,N,B", [(600, 300, 300, 100), (6144, 300, 128, 1024), (130, 64, 68, 130), (257, 300, 300, 100),
                                      (49152, 300, 300, 8192)])
@pytest.mark.parametrize("with_bn", [True, False])
def test_fc_tensor_core_3xtf32(R, K, N, B, with_bn):
    """tcgen05 kind::tf32 with error-compensated operands (3 MMAs per product): same 1e-5 bar as the FFMA path."""
    from dssm_b200 import ops

    rng = np.random.default_rng(R + K + N)
    h = rng.standard_normal((R, K)).astype(np.float32)
    W = rng.uniform(-0.1, 0.1, (K, N)).astype(np.float32)
    b = rng.uniform(-0.1, 0.1, (N,)).astype(np.float32)
    scale = rng.uniform(0.5, 1.5, (2, K)).astype(np.float32) if with_bn else None
    shift = rng.uniform(-0.5, 0.5, (2, K)).astype(np.float32) if with_bn else None
    a = h.astype(np.float64)
    if with_bn:
        a = np.concatenate([a[:B] * scale[0] + shift[0], a[B:] * scale[1] + shift[1]])
    a = np.maximum(a, 0)
    want = a @ W.astype(np.float64) + b
    got = ops.fc_fwd(dev(h), dev(W), dev(b), dev(scale) if with_bn else None, dev(shift) if with_bn else None, "relu", B,
                     gemm_mode="tc_3xtf32")
    assert_close(got.cpu().numpy(), want, TOL, "fc_fwd tc")
    dH = rng.standard_normal((R, N)).astype(np.float32)
    dA = ops.fc_bwd_dx(dev(dH), dev(W), gemm_mode="tc_3xtf32").cpu().numpy()
    assert_close(dA, dH.astype(np.float64) @ W.T.astype(np.float64), TOL, "fc_bwd_dx tc")
    dW, db = ops.fc_bwd_dw(dev(h), dev(dH), dev(scale) if with_bn else None, dev(shift) if with_bn else None, "relu", B,
                           gemm_mode="tc_3xtf32")
    assert_close(dW.cpu().numpy(), a.T @ dH.astype(np.float64), TOL, "fc_bwd_dw tc")
    assert_close(db.cpu().numpy(), dH.astype(np.float64).sum(0), TOL, "db tc")


def test_add_layer_reference_call_shape():
    """add_layer(inputs, in_size, out_size, activation_function) creates Xavier-uniform W and b (dssm_v3.py:44-53)."""
    from dssm_b200 import add_layer

    x = torch.randn((50, 40), device="cuda")
    out, W, b = add_layer(x, 40, 30, activation_function="relu")
    lim = np.sqrt(6.0 / 70)
    assert W.abs().max().item() <= lim and b.abs().max().item() <= lim and W.abs().max().item() > 0.5 * lim
    want = torch.relu(x.double() @ W.double() + b.double())
    assert rel_err(out.cpu().numpy(), want.cpu().numpy()) < TOL
    X = random_csr(np.random.default_rng(0), 20, 40)
    out2, _, _ = add_layer(csr_dev(X), 40, 30, None, Weights=W, biases=b)
    assert rel_err(out2.cpu().numpy(), np.asarray(X @ W.cpu().numpy()) + b.cpu().numpy()) < TOL


# ---- Merge_Negative_Doc ------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,NEG,L", [(5, 3, 4), (100, 4, 128), (33, 1, 7), (16, 50, 12)])
def test_merge_negative_doc_bit_exact(B, NEG, L):
    from dssm_b200 import Merge_Negative_Doc, merge_negative_doc_index
    from oracle import merge_negative_doc_literal

    rng = np.random.default_rng(B)
    pos = rng.standard_normal((B, L)).astype(np.float32)
    neg = rng.standard_normal((B * NEG, L)).astype(np.float32)
    doc_y, _, src = merge_negative_doc_literal(pos, neg, B, NEG)
    got = Merge_Negative_Doc(dev(pos), dev(neg), B, NEG).cpu().numpy()
    assert np.array_equal(got, doc_y)
    assert np.array_equal(merge_negative_doc_index(B, NEG).cpu().numpy(), src)


def test_merge_negative_doc_full_size_closed_form():
    from dssm_b200 import Merge_Negative_Doc, merge_negative_doc_index

    B, NEG, L = 1024, 50, 128
    pos = torch.randn((B, L), device="cuda")
    neg = torch.randn((B * NEG, L), device="cuda")
    got = Merge_Negative_Doc(pos, neg, B, NEG)
    want = torch.cat([pos, neg.view(B, NEG, L).transpose(0, 1).reshape(B * NEG, L)])
    assert torch.equal(got, want)
    r = torch.arange(B * NEG, device="cuda")
    src = merge_negative_doc_index(B, NEG)
    assert torch.equal(src[B:].long(), B + (r % B) * NEG + r // B) and torch.equal(src[:B].long(), torch.arange(B, device="cuda"))


# ---- Cosine_Similarity + Loss ------------------------------------------------------------------------------
@pytest.mark.parametrize("B,NEG,L,eps,div", [(100, 4, 128, 0.0, True), (7, 3, 5, 1e-8, True), (33, 50, 128, 0.0, False), (4, 1, 300, 0.0, True)])
def test_cos_softmax_loss_vs_oracle(B, NEG, L, eps, div):
    from dssm_b200 import ops
    from oracle import DSSMOracle, OracleConfig, cosine_similarity_literal, init_params, merge_negative_doc_literal

    rng = np.random.default_rng(B + NEG)
    Y = (np.maximum(rng.standard_normal(((2 + NEG) * B, L)), 0) + 0.01).astype(np.float32)
    cfg = OracleConfig(TRIGRAM_D=8, layers=(4, L), NEG=NEG, query_BS=B, loss_eps=eps, loss_div_bs=div)
    orc = DSSMOracle(cfg, init_params(cfg, 0), dtype=np.float64)
    ref = orc.cosine_loss({"Y": Y.astype(np.float64)})
    out = ops.cos_softmax_loss(dev(Y), B, NEG, 20.0, eps, div, want_grad=True)
    for k in ("cos_sim_raw", "doc_norm", "query_norm_single", "cos_sim", "prob"):
        assert_close(out[k].cpu().numpy().reshape(ref[k].shape), ref[k], TOL, k)
    assert abs(out["loss"].item() - float(ref["loss"])) <= TOL * abs(float(ref["loss"]))
    # gradient vs torch fp64 autograd
    Yt = torch.tensor(Y, dtype=torch.float64, requires_grad=True)
    q, pos, neg = Yt[:B], Yt[B:2 * B], Yt[2 * B:]
    docs = torch.cat([pos[:, None, :], neg.reshape(B, NEG, L)], 1)
    cos = (q[:, None, :] * docs).sum(-1) / (q.norm(dim=1)[:, None] * docs.norm(dim=2))
    loss = -torch.log(torch.softmax(cos * 20.0, 1)[:, 0] + eps).sum() / (B if div else 1)
    loss.backward()
    assert_close(out["dY"].cpu().numpy(), Yt.grad.numpy(), 2e-5, "dY")
    # reference memory order of cos_sim_raw (row k*B + j) against the literal replay
    doc_y, _, _ = merge_negative_doc_literal(Y[B:2 * B], Y[2 * B:], B, NEG)
    lit = cosine_similarity_literal(Y[:B], doc_y, B, NEG)
    assert_close(out["cos_sim_raw"].cpu().numpy(), lit["cos_sim_raw"], TOL, "cos_sim_raw order")


def test_zero_embedding_row_propagates_nan():
    from dssm_b200 import ops

    B, NEG, L = 6, 2, 16
    Y = np.ones(((2 + NEG) * B, L), np.float32)
    Y[B + 1] = 0  # positive of query 1 -> 0/0
    Y[2 * B + 3 * NEG + 1] = 0  # a negative of query 3
    out = ops.cos_softmax_loss(dev(Y), B, NEG, want_grad=True)
    cs = out["cos_sim"].cpu().numpy()
    assert np.isnan(cs[1, 0]) and np.isnan(cs[3, 2]) and np.isfinite(cs[[0, 2, 4, 5]]).all()
    lt = out["loss_terms"].cpu().numpy()
    assert np.isnan(lt[[1, 3]]).all() and np.isfinite(lt[[0, 2, 4, 5]]).all()
    assert np.isnan(out["loss"].item())
    assert np.isnan(out["prob"].cpu().numpy()[[1, 3]]).all()


# ---- Adam --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [4, 1003, 300 * 300 + 300])
def test_adam_three_steps_vs_tf_formula(n):
    from dssm_b200 import ops

    rng = np.random.default_rng(n)
    p = rng.standard_normal(n).astype(np.float32)
    m = np.zeros(n, np.float64)
    v = np.zeros(n, np.float64)
    ref = p.astype(np.float64)
    P, M, V = dev(p), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    bp = torch.tensor([0.9, 0.999], device="cuda")
    # TF's ApplyAdam works on fp32 scalars: (1 - beta2) is 1 - float32(0.999), 1.3e-5 away from 0.001
    b1, b2, lr, eps = (float(np.float32(x)) for x in (0.9, 0.999, 0.01, 1e-8))
    omb1, omb2 = float(np.float32(1) - np.float32(0.9)), float(np.float32(1) - np.float32(0.999))
    b1p, b2p = b1, b2
    for step in range(3):
        g = rng.standard_normal(n).astype(np.float32)
        g[::7] = 0
        ops.adam_step(P, dev(g), M, V, bp, 0.01, grad_scale=0.5)
        ge = 0.5 * g.astype(np.float64)
        m = b1 * m + omb1 * ge
        v = b2 * v + omb2 * ge * ge
        ref = ref - lr * np.sqrt(1 - b2p) / (1 - b1p) * m / (np.sqrt(v) + eps)
        b1p, b2p = float(np.float32(b1p * b1)), float(np.float32(b2p * b2))
    assert_close(P.cpu().numpy(), ref, TOL, "adam params")
    assert_close(M.cpu().numpy(), m, TOL, "adam m")
    assert_close(V.cpu().numpy(), v, TOL, "adam v")
    assert_close(bp.cpu().numpy(), np.array([b1p, b2p]), 1e-6, "beta powers")


def test_errors_are_loud():
    from dssm_b200 import DssmError, ops

    with pytest.raises(DssmError):
        ops.cos_softmax_loss(torch.ones((8, 4), device="cuda"), 2, 0)  # NEG must be positive
    with pytest.raises(ValueError):
        ops.spmm_fwd(csr_dev(random_csr(np.random.default_rng(0), 4, 8)), torch.ones((8, 4)), None)  # CPU tensor
