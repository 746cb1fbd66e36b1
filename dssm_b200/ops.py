"""Op-level Python surface: the reference's model-building names over the C ABI.

    add_layer               archive/dssm_v3.py:44-53, README.md:66-84   (dense input: tf.matmul; sparse input:
                            tf.sparse_tensor_dense_matmul as in new_dssm.py:124-126)
    batch_normalization     semantic_matching/dssm/new_dssm.py:62-88
    Merge_Negative_Doc      new_dssm.py:160-180
    Cosine_Similarity       new_dssm.py:182-201
    Loss                    new_dssm.py:203-213

torch tensors are device buffers only; every function launches hand-written sm_100a kernels through
libdssm_b200.so and raises DssmError on failure.  Nothing here computes on the host.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import ACT, GEMM, check, lib, ptr, stream_ptr
from .batch import StackedBatch


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise ValueError("dssm_b200 ops take CUDA tensors (there is no CPU path)")


def _f32(shape, device):
    return torch.empty(shape, dtype=torch.float32, device=device)


def _ws(nbytes: int, device):
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def _ws_zero(nbytes: int, device):
    """Workspaces that start with self-resetting ticket counters (BN reductions, split-K dW) must begin zeroed."""
    return torch.zeros(max(int(nbytes), 256), dtype=torch.uint8, device=device)


@dataclass
class DeviceCSR:
    """Stacked batch on the device (int32 indptr/indices, fp32 values)."""

    indptr: torch.Tensor
    indices: torch.Tensor
    values: torch.Tensor
    rows: int
    n_cols: int
    nnz: int

    @staticmethod
    def from_host(b: StackedBatch, device="cuda") -> "DeviceCSR":
        dev = torch.device(device)
        return DeviceCSR(torch.from_numpy(b.indptr).to(dev), torch.from_numpy(b.indices).to(dev),
                         torch.from_numpy(b.values).to(dev), b.rows, b.n_cols, b.nnz)


# ---- FC1 ---------------------------------------------------------------------------------------------
def spmm_fwd(x: DeviceCSR, W1: torch.Tensor, b1: Optional[torch.Tensor]) -> torch.Tensor:
    _require_cuda(x.indptr, W1, b1)
    D, L1 = W1.shape
    assert D == x.n_cols, "W1 rows must equal TRIGRAM_D"
    Y = _f32((x.rows, L1), W1.device)
    check(lib.dssm_spmm_fwd(ptr(x.indptr), ptr(x.indices), ptr(x.values), x.rows, D, ptr(W1), ptr(b1), L1, ptr(Y), stream_ptr()))
    return Y


def spmm_bwd_dw(x: DeviceCSR, dH: torch.Tensor, D: int, method: int = 0) -> torch.Tensor:
    _require_cuda(x.indptr, dH)
    R, L1 = dH.shape
    dW = _f32((D, L1), dH.device)
    nb = lib.dssm_spmm_bwd_dw_workspace_bytes(R, D, L1, max(x.nnz, 1))
    ws = _ws(nb, dH.device)
    check(lib.dssm_spmm_bwd_dw(ptr(x.indptr), ptr(x.indices), ptr(x.values), R, D, ptr(dH), L1, ptr(dW), method, ptr(ws),
                               ws.numel(), stream_ptr()))
    return dW


# ---- add_layer ---------------------------------------------------------------------------------------
def xavier_uniform(in_size: int, out_size: int, device, generator: Optional[torch.Generator] = None):
    """tf.random_uniform([in,out], -wlimit, wlimit) for W and for b (archive/dssm_v3.py:45-47)."""
    wlimit = float(np.sqrt(6.0 / (in_size + out_size)))
    W = (torch.rand((in_size, out_size), generator=generator, dtype=torch.float32) * 2 - 1) * wlimit
    b = (torch.rand((out_size,), generator=generator, dtype=torch.float32) * 2 - 1) * wlimit
    return W.to(device), b.to(device)


def add_layer(inputs, in_size: int, out_size: int, activation_function: Optional[str] = None,
              Weights: Optional[torch.Tensor] = None, biases: Optional[torch.Tensor] = None, gemm_mode: str = "fp32",
              generator: Optional[torch.Generator] = None):
    """outputs = activation_function(inputs @ Weights + biases); fresh Xavier-uniform Weights/biases per call as
    in the reference unless given.  `inputs` is a DeviceCSR (FC1) or a dense [rows, in_size] tensor.
    Returns (outputs, Weights, biases)."""
    device = inputs.indptr.device if isinstance(inputs, DeviceCSR) else inputs.device
    if Weights is None:
        Weights, biases = xavier_uniform(in_size, out_size, device, generator)
    assert tuple(Weights.shape) == (in_size, out_size)
    if isinstance(inputs, DeviceCSR):
        wx_plus_b = spmm_fwd(inputs, Weights, biases)
    else:
        wx_plus_b = fc_fwd(inputs, Weights, biases, gemm_mode=gemm_mode)
    if activation_function is None:
        return wx_plus_b, Weights, biases
    out = bn_act_apply(wx_plus_b, None, None, activation_function, wx_plus_b.shape[0])
    return out, Weights, biases


def fc_fwd(Hprev, W, bias, scale=None, shift=None, act=None, B: int = 0, gemm_mode: str = "fp32"):
    _require_cuda(Hprev, W, bias, scale, shift)
    R, K = Hprev.shape
    N = W.shape[1]
    out = _f32((R, N), Hprev.device)
    ws = _ws(lib.dssm_fc_fwd_workspace_bytes(K, N, GEMM[gemm_mode]), Hprev.device)
    check(lib.dssm_fc_fwd(ptr(Hprev), R, K, B, ptr(scale), ptr(shift), ACT[act], ptr(W), ptr(bias), N, ptr(out),
                          GEMM[gemm_mode], ptr(ws), ws.numel(), stream_ptr()))
    return out


def fc_bwd_dx(dH, W, gemm_mode: str = "fp32"):
    R, N = dH.shape
    K = W.shape[0]
    dA = _f32((R, K), dH.device)
    ws = _ws(lib.dssm_fc_fwd_workspace_bytes(K, N, GEMM[gemm_mode]), dH.device)
    check(lib.dssm_fc_bwd_dx(ptr(dH), R, N, ptr(W), K, ptr(dA), GEMM[gemm_mode], ptr(ws), ws.numel(), stream_ptr()))
    return dA


def fc_bwd_dw(Hprev, dH, scale=None, shift=None, act=None, B: int = 0, gemm_mode: str = "fp32"):
    R, K = Hprev.shape
    N = dH.shape[1]
    dW, db = _f32((K, N), dH.device), _f32((N,), dH.device)
    ws = _ws(lib.dssm_fc_bwd_dw_workspace_bytes(R, K, N), dH.device)
    check(lib.dssm_fc_bwd_dw(ptr(Hprev), R, K, B, ptr(scale), ptr(shift), ACT[act], ptr(dH), N, ptr(dW), ptr(db),
                             GEMM[gemm_mode], ptr(ws), ws.numel(), stream_ptr()))
    return dW, db


def colsum(X):
    R, N = X.shape
    out = _f32((N,), X.device)
    ws = _ws(lib.dssm_colsum_workspace_bytes(R, N), X.device)
    check(lib.dssm_colsum(ptr(X), R, N, ptr(out), ptr(ws), ws.numel(), stream_ptr()))
    return out


# ---- batch_normalization -----------------------------------------------------------------------------
class BNState:
    """Variables of the two batch_normalization instances of one layer ([2,L]: 0 = query, 1 = doc):
    beta/gamma trainable (new_dssm.py:75-76), EMA shadows zero-initialised (ExponentialMovingAverage on
    tensors, :78-81)."""

    def __init__(self, out_size: int, device):
        self.gamma = torch.ones((2, out_size), dtype=torch.float32, device=device)
        self.beta = torch.zeros((2, out_size), dtype=torch.float32, device=device)
        self.ema_mean = torch.zeros((2, out_size), dtype=torch.float32, device=device)
        self.ema_var = torch.zeros((2, out_size), dtype=torch.float32, device=device)
        z = lambda: torch.zeros((2, out_size), dtype=torch.float32, device=device)
        self.mean, self.var, self.rstd, self.scale, self.shift = z(), z(), z(), z(), z()


def bn_forward(X, B: int, state: BNState, on_train: bool, update_ema: bool = True, eps: float = 1e-3, decay: float = 0.5):
    """Both instances of a layer over the stacked rows (query rows [0,B), doc rows [B,R)); fills
    state.mean/var/rstd/scale/shift (and moves the shadows when training)."""
    _require_cuda(X)
    R, L = X.shape
    ws = _ws_zero(lib.dssm_bn_workspace_bytes(R, L), X.device)
    check(lib.dssm_bn_forward(ptr(X), R, L, B, int(on_train), int(update_ema), ptr(state.gamma), ptr(state.beta),
                              ptr(state.ema_mean), ptr(state.ema_var), eps, decay, ptr(state.mean), ptr(state.var),
                              ptr(state.rstd), ptr(state.scale), ptr(state.shift), ptr(ws), ws.numel(), stream_ptr()))
    return state


def bn_act_apply(X, scale, shift, act, B: int):
    R, L = X.shape
    Y = _f32((R, L), X.device)
    check(lib.dssm_bn_act_apply(ptr(X), R, L, B, ptr(scale), ptr(shift), ACT[act], ptr(Y), stream_ptr()))
    return Y


def batch_normalization(x, phase_train: bool, out_size: int, state: Optional[BNState] = None, query_BS: Optional[int] = None):
    """normed = batch_normalization(x, phase_train, out_size)  (new_dssm.py:62-88).
    x is either the stacked [query ; docs] activations (pass query_BS: both instances at once, as the tower
    does) or a single tensor (one instance, exactly the reference call).  Returns (normed, state)."""
    R, L = x.shape
    assert L == out_size
    if state is None:
        state = BNState(out_size, x.device)
    B = R if query_BS is None else query_BS
    bn_forward(x, B, state, phase_train)
    return bn_act_apply(x, state.scale, state.shift, None, B), state


def bn_act_backward(dA, H, B: int, act, state: Optional[BNState], want_db: bool = False):
    """In place: dA becomes dLoss/dH.  Returns (dgamma, dbeta) [2,L] (None without BN), plus db [L] when want_db."""
    R, L = H.shape
    if state is None:
        check(lib.dssm_bn_act_backward(ptr(dA), ptr(H), R, L, B, ACT[act], None, None, None, None, None, None, None, None, None,
                                       0, stream_ptr()))
        return (None, None, None) if want_db else (None, None)
    dgamma, dbeta = _f32((2, L), H.device), _f32((2, L), H.device)
    db = _f32((L,), H.device) if want_db else None
    ws = _ws_zero(lib.dssm_bn_workspace_bytes(R, L), H.device)
    check(lib.dssm_bn_act_backward(ptr(dA), ptr(H), R, L, B, ACT[act], ptr(state.gamma), ptr(state.mean), ptr(state.rstd),
                                   ptr(state.scale), ptr(state.shift), ptr(dgamma), ptr(dbeta), ptr(db), ptr(ws), ws.numel(),
                                   stream_ptr()))
    return (dgamma, dbeta, db) if want_db else (dgamma, dbeta)


# ---- Merge_Negative_Doc / Cosine_Similarity / Loss -----------------------------------------------------
def Merge_Negative_Doc(doc_positive_y, doc_negative_y, query_BS: int, NEG: int):
    """doc_y [(1+NEG)*B, L] in the reference's concat order (new_dssm.py:162-179)."""
    _require_cuda(doc_positive_y, doc_negative_y)
    L = doc_positive_y.shape[1]
    doc_y = _f32(((1 + NEG) * query_BS, L), doc_positive_y.device)
    check(lib.dssm_merge_negative_doc(ptr(doc_positive_y), ptr(doc_negative_y), query_BS, NEG, L, ptr(doc_y), stream_ptr()))
    return doc_y


def merge_negative_doc_index(query_BS: int, NEG: int, device="cuda"):
    src = torch.empty(((1 + NEG) * query_BS,), dtype=torch.int32, device=device)
    check(lib.dssm_merge_negative_doc_index(query_BS, NEG, ptr(src), stream_ptr()))
    return src


def cos_softmax_loss(Y, query_BS: int, NEG: int, gamma: float = 20.0, loss_eps: float = 0.0, loss_div_bs: bool = True,
                     want_grad: bool = False):
    """Cosine_Similarity + Loss over the stacked embeddings Y = [query_y ; doc_positive_y ; doc_negative_y]."""
    _require_cuda(Y)
    B, K1, L = query_BS, NEG + 1, Y.shape[1]
    dev = Y.device
    out = dict(query_norm_single=_f32((B, 1), dev), doc_norm=_f32((K1 * B, 1), dev), cos_sim_raw=_f32((K1 * B, 1), dev),
               cos_sim=_f32((B, K1), dev), prob=_f32((B, K1), dev), loss_terms=_f32((B,), dev), loss=_f32((1,), dev))
    dY = _f32(tuple(Y.shape), dev) if want_grad else None
    check(lib.dssm_cos_softmax_loss(ptr(Y), B, NEG, L, gamma, loss_eps, int(loss_div_bs), ptr(out["query_norm_single"]),
                                    ptr(out["doc_norm"]), ptr(out["cos_sim_raw"]), ptr(out["cos_sim"]), ptr(out["prob"]),
                                    ptr(out["loss_terms"]), ptr(out["loss"]), ptr(dY), stream_ptr()))
    out["hit_prob"] = out["prob"][:, 0:1]
    if want_grad:
        out["dY"] = dY
    return out


def Cosine_Similarity(query_y, doc_positive_y, doc_negative_y, NEG: int, gamma: float = 20.0):
    """Returns the tensors the reference scope defines (new_dssm.py:185-199): query_norm_single, doc_norm,
    cos_sim_raw, cos_sim."""
    B = query_y.shape[0]
    Y = torch.cat([query_y, doc_positive_y, doc_negative_y], dim=0).contiguous()
    return cos_softmax_loss(Y, B, NEG, gamma)


def Loss(query_y, doc_positive_y, doc_negative_y, NEG: int, gamma: float = 20.0, loss_eps: float = 0.0, loss_div_bs: bool = True):
    B = query_y.shape[0]
    Y = torch.cat([query_y, doc_positive_y, doc_negative_y], dim=0).contiguous()
    return cos_softmax_loss(Y, B, NEG, gamma, loss_eps, loss_div_bs)["loss"]


# ---- Training ----------------------------------------------------------------------------------------
def adam_step(params, grads, m, v, beta_pow, lr: float, beta1=0.9, beta2=0.999, eps=1e-8, grad_scale: float = 1.0, advance: bool = True):
    _require_cuda(params, grads, m, v, beta_pow)
    check(lib.dssm_adam_step(ptr(params), ptr(grads), ptr(m), ptr(v), params.numel(), ptr(beta_pow), lr, beta1, beta2, eps,
                             grad_scale, stream_ptr()))
    if advance:
        check(lib.dssm_adam_advance(ptr(beta_pow), beta1, beta2, stream_ptr()))
