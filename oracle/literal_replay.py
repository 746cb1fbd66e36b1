"""Literal, op-by-op NumPy replay of the reference's Python-level graph construction
(TEST INFRASTRUCTURE ONLY; small sizes only -- the concat chain is O((B*NEG)^2)).

It exists to pin the *ordering* semantics that the closed forms in dssm_oracle.py and the
CUDA kernels must reproduce bit-exactly:

  merge_negative_doc_literal  <- new_dssm.py:160-180 (tf.tile + B*NEG x (tf.slice, tf.concat))
  cosine_similarity_literal   <- new_dssm.py:182-201 (tile / reduce_sum / truediv / transpose / reshape)
  loss_literal                <- new_dssm.py:203-213

Each numpy call below stands for the TF op on the cited line; nothing is simplified.
"""
from __future__ import annotations

import numpy as np


def merge_negative_doc_literal(doc_positive_y: np.ndarray, doc_negative_y: np.ndarray, query_BS: int, NEG: int):
    """Returns (doc_y, label, src_index) where src_index[r] names the source row of doc_y[r]:
    r < B -> positive row r (encoded r), else B + (negative row index)."""
    doc_y = np.tile(doc_positive_y, [1, 1])  # :162
    src = list(range(query_BS))
    label_pos = [1] * query_BS  # :163
    label_neg = [0] * query_BS * NEG  # :164
    label = label_pos + label_neg  # :165
    for i in range(NEG):  # :169
        for j in range(query_BS):  # :171
            row = j * NEG + i
            sl = doc_negative_y[row : row + 1, :]  # tf.slice(doc_negative_y, [j*NEG+i, 0], [1,-1]) :174-178
            doc_y = np.concatenate([doc_y, sl], axis=0)  # tf.concat(..., 0) :173-179
            src.append(query_BS + row)
    return doc_y, np.asarray(label, dtype=np.int32), np.asarray(src, dtype=np.int32)


def cosine_similarity_literal(query_y: np.ndarray, doc_y: np.ndarray, query_BS: int, NEG: int, gamma: float = 20.0):
    dt = query_y.dtype
    query_norm = np.tile(np.sqrt(np.sum(np.square(query_y), 1, keepdims=True)), [NEG + 1, 1])  # :185
    query_norm_single = np.sqrt(np.sum(np.square(query_y), 1, keepdims=True))  # :187
    doc_norm = np.sqrt(np.sum(np.square(doc_y), 1, keepdims=True))  # :190
    prod = np.sum(np.multiply(np.tile(query_y, [NEG + 1, 1]), doc_y), 1, keepdims=True)  # :193
    norm_prod = np.multiply(query_norm, doc_norm)  # :194
    with np.errstate(invalid="ignore", divide="ignore"):
        cos_sim_raw = np.true_divide(prod, norm_prod)  # :197
    cos_sim = np.transpose(np.reshape(np.transpose(cos_sim_raw), [NEG + 1, query_BS])) * dt.type(gamma)  # :199
    return dict(query_norm=query_norm, query_norm_single=query_norm_single, doc_norm=doc_norm, prod=prod,
                cos_sim_raw=cos_sim_raw, cos_sim=cos_sim)


def loss_literal(cos_sim: np.ndarray, query_BS: int, loss_eps: float = 0.0, loss_div_bs: bool = True):
    dt = cos_sim.dtype
    z = cos_sim - cos_sim.max(axis=1, keepdims=True)
    e = np.exp(z)
    prob = e / e.sum(axis=1, keepdims=True)  # tf.nn.softmax :206
    hit_prob = prob[:, 0:1]  # tf.slice(prob,[0,0],[-1,1]) :208
    loss = -np.sum(np.log(hit_prob + dt.type(loss_eps)))  # :209 / my_dssm.py:169
    if loss_div_bs:
        loss = loss / dt.type(query_BS)
    return prob, hit_prob, dt.type(loss)
