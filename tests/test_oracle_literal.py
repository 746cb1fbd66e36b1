"""The oracle's closed forms against the literal op-by-op replay of the reference's Python loops
(new_dssm.py:160-213): ordering must be bit-exact, values identical up to summation order."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import DSSMOracle, OracleConfig, cosine_similarity_literal, init_params, loss_literal, merge_negative_doc_literal
from tests.helpers import random_csr


@given(B=st.integers(1, 9), NEG=st.integers(1, 6))
@settings(max_examples=40, deadline=None)
def test_merge_order_closed_form(B, NEG):
    L = 3
    pos = np.arange(B * L, dtype=np.float32).reshape(B, L)
    neg = 1000 + np.arange(B * NEG * L, dtype=np.float32).reshape(B * NEG, L)
    doc_y, label, src = merge_negative_doc_literal(pos, neg, B, NEG)
    # closed form of SURVEY.md 3.2: doc_y[(i+1)*B + j] = neg[j*NEG + i]
    want = np.concatenate([pos, neg.reshape(B, NEG, L).transpose(1, 0, 2).reshape(B * NEG, L)])
    assert np.array_equal(doc_y, want)
    r = np.arange(B * NEG)
    assert np.array_equal(src[B:], B + (r % B) * NEG + r // B)
    assert np.array_equal(src[:B], np.arange(B))
    assert label.tolist() == [1] * B + [0] * (B * NEG)


@pytest.mark.parametrize("B,NEG,L", [(5, 3, 4), (7, 1, 8), (4, 6, 5)])
def test_cosine_loss_closed_form_vs_literal(B, NEG, L):
    rng = np.random.default_rng(1)
    cfg = OracleConfig(TRIGRAM_D=16, layers=(6, L), NEG=NEG, query_BS=B)
    orc = DSSMOracle(cfg, init_params(cfg, 0))
    Y = np.maximum(rng.standard_normal(((2 + NEG) * B, L)).astype(np.float32), 0) + np.float32(0.01)
    cache = {"Y": Y}
    out = orc.cosine_loss(cache)
    q, pos, neg = Y[:B], Y[B:2 * B], Y[2 * B:]
    doc_y, _, _ = merge_negative_doc_literal(pos, neg, B, NEG)
    lit = cosine_similarity_literal(q, doc_y, B, NEG)
    np.testing.assert_allclose(out["cos_sim_raw"], lit["cos_sim_raw"].ravel(), rtol=2e-6, atol=1e-7)
    np.testing.assert_allclose(out["cos_sim"], lit["cos_sim"], rtol=2e-6, atol=1e-6)
    np.testing.assert_allclose(out["query_norm_single"], lit["query_norm_single"].ravel(), rtol=1e-6)
    np.testing.assert_allclose(out["doc_norm"], lit["doc_norm"].ravel(), rtol=1e-6)
    prob, hit, loss = loss_literal(lit["cos_sim"], B)
    np.testing.assert_allclose(out["prob"], prob, rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(out["loss"], loss, rtol=1e-5)


def test_loss_variants_literal():
    rng = np.random.default_rng(2)
    cs = (rng.standard_normal((6, 5)) * 5).astype(np.float32)
    _, _, l0 = loss_literal(cs, 6)
    _, _, l1 = loss_literal(cs, 6, loss_div_bs=False)
    _, _, l2 = loss_literal(cs, 6, loss_eps=1e-8)
    assert np.isclose(l1, l0 * 6, rtol=1e-6)
    assert abs(l2 - l0) < 1e-4


def test_zero_embedding_row_gives_nan_like_reference():
    """new_dssm.py:197 has no epsilon: an all-zero relu embedding makes 0/0 = NaN and poisons that group's loss."""
    B, NEG, L = 3, 2, 4
    cfg = OracleConfig(TRIGRAM_D=8, layers=(4, L), NEG=NEG, query_BS=B)
    orc = DSSMOracle(cfg, init_params(cfg, 0))
    Y = np.ones(((2 + NEG) * B, L), np.float32)
    Y[B + 1] = 0  # positive doc of query 1
    out = orc.cosine_loss({"Y": Y})
    assert np.isnan(out["cos_sim"][1, 0]) and np.isnan(out["loss"])
    assert not np.isnan(out["cos_sim"][0]).any()
