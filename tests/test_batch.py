"""Batch format (utils/utils.py:20-24,45-61; utils/data_input.py:53-60,157-159) and the synthetic generator."""
import numpy as np
import pytest
import scipy.sparse as sp

from dssm_b200 import Config
from dssm_b200.batch import (bow_csr, convert_seq2bow, convert_sparse_matrix_to_sparse_tensor, pull_batch, stack_csr,
                             stack_feed)
from dssm_b200.parallel import shard_stacked_batch
from dssm_b200.synthetic import init_params, make_batch
from tests.helpers import random_csr


def test_convert_sparse_matrix_is_row_major_coo():
    X = random_csr(np.random.default_rng(0), 6, 20)
    st = convert_sparse_matrix_to_sparse_tensor(X)
    assert st.indices.shape == (X.nnz, 2) and st.dense_shape == (6, 20)
    order = np.lexsort((st.indices[:, 1], st.indices[:, 0]))
    assert np.array_equal(order, np.arange(X.nnz))  # already sorted row-major, columns ascending


def test_pull_batch_slices_like_reference():
    rng = np.random.default_rng(1)
    conf = Config(TRIGRAM_D=30, query_BS=4, NEG=3)
    q, p, n = random_csr(rng, 12, 30), random_csr(rng, 12, 30), random_csr(rng, 36, 30)
    feed = pull_batch(True, q, p, n, 1, conf.query_BS, conf=conf)
    sb = stack_feed(feed, conf)
    X = sb.to_scipy()
    assert feed["on_train"] is True
    assert (X[:4] != q[4:8]).nnz == 0 and (X[4:8] != p[4:8]).nnz == 0 and (X[8:] != n[12:24]).nnz == 0
    assert sb.indptr.dtype == np.int32 and sb.indices.dtype == np.int32 and sb.values.dtype == np.float32


def test_short_batch_is_rejected_like_static_query_bs():
    rng = np.random.default_rng(2)
    with pytest.raises(ValueError):
        stack_csr(random_csr(rng, 3, 10), random_csr(rng, 4, 10), random_csr(rng, 8, 10), 4, 2)


def test_int64_counts_are_cast_to_float32():
    X = sp.csr_matrix(np.array([[0, 2, 0], [1, 0, 3]], dtype=np.int64))
    sb = stack_csr(X[:1], X[1:], X[:1], 1, 1)
    assert sb.values.dtype == np.float32 and sb.values.tolist() == [2.0, 1.0, 3.0, 2.0]


def test_bow_encoding_matches_convert_seq2bow():
    vocab = {"[PAD]": 0, "a": 1, "b": 2, "[UNK]": 3, "c": 4}
    texts = ["abca", "zzb", ""]
    m = bow_csr(texts, vocab)
    dense = np.stack([convert_seq2bow(t, vocab) for t in texts])
    assert np.array_equal(m.toarray(), dense)
    assert dense[1, 3] == 2  # two OOV characters counted on [UNK]


def test_synthetic_batch_spec():
    conf = Config(TRIGRAM_D=49284, query_BS=64, NEG=4, layers=(300, 300, 128))
    b = make_batch(conf, seed=0)
    X = b.to_scipy()
    assert X.shape == (conf.rows, 49284) and X.has_sorted_indices
    per_row = np.diff(b.indptr)
    assert per_row.min() >= 1 and per_row.max() <= 64
    assert per_row[:64].mean() < per_row[64:].mean()  # queries shorter than docs
    assert (b.values >= 1).all() and np.array_equal(b.values, np.round(b.values))
    b2 = make_batch(conf, seed=0)
    assert np.array_equal(b.indices, b2.indices)
    t = make_batch(conf, seed=1, value_mode="tfidf").to_scipy()
    np.testing.assert_allclose(np.sqrt(np.asarray(t.multiply(t).sum(1)).ravel()), 1.0, rtol=1e-5)


def test_init_params_matches_oracle_rule():
    from oracle import OracleConfig, init_params as oracle_init

    conf = Config(TRIGRAM_D=100, query_BS=4, NEG=2, layers=(12, 8))
    a = init_params(conf, 3)
    b = oracle_init(OracleConfig(TRIGRAM_D=100, layers=(12, 8), NEG=2, query_BS=4), 3)
    assert a.keys() == b.keys()
    for k in a:
        assert np.array_equal(a[k], b[k])
    lim = np.sqrt(6.0 / (100 + 12))
    assert np.abs(a["W1"]).max() <= lim and np.abs(a["b1"]).max() <= lim


def test_shard_stacked_batch_partitions_groups():
    conf = Config(TRIGRAM_D=50, query_BS=8, NEG=3)
    rng = np.random.default_rng(0)
    sb = stack_csr(random_csr(rng, 8, 50), random_csr(rng, 8, 50), random_csr(rng, 24, 50), 8, 3)
    X = sb.to_scipy()
    parts = [shard_stacked_batch(sb, 8, 3, r, 2).to_scipy() for r in range(2)]
    for r, P in enumerate(parts):
        assert P.shape[0] == (2 + 3) * 4
        assert (P[:4] != X[4 * r:4 * r + 4]).nnz == 0
        assert (P[4:8] != X[8 + 4 * r:8 + 4 * r + 4]).nnz == 0
        assert (P[8:] != X[16 + 12 * r:16 + 12 * r + 12]).nnz == 0
