"""Retrieval oracle: ordering contract (score desc, id asc; NaN last) and shard-merge == whole corpus."""
import numpy as np

from oracle import corpus_topk_oracle, exact_cosine_scores, merge_topk_oracle


def test_ties_break_to_lower_id_and_nan_ranks_last():
    Q = np.array([[1.0, 0.0], [0.0, 1.0]], np.float32)
    docs = np.array([[2, 0], [1, 0], [0, 0], [1, 1], [3, 0]], np.float32)  # docs 0,1,4 tie for q0; doc 2 is 0/0
    s, i = corpus_topk_oracle(Q, docs, 5)
    assert i[0].tolist() == [0, 1, 4, 3, 2]
    assert s[0, -1] == -np.inf
    assert i[1].tolist()[:1] == [3]


def test_shard_merge_equals_whole():
    rng = np.random.default_rng(0)
    Q = np.maximum(rng.standard_normal((7, 16)), 0).astype(np.float32)
    docs = np.maximum(rng.standard_normal((300, 16)), 0).astype(np.float32)
    docs[17] = docs[5]  # exact duplicate -> tie
    whole = corpus_topk_oracle(Q, docs, 10)
    parts = [corpus_topk_oracle(Q, docs[lo:hi], 10, id_offset=lo) for lo, hi in ((0, 100), (100, 200), (200, 300))]
    merged = merge_topk_oracle(parts, 10)
    assert np.array_equal(whole[1], merged[1]) and np.array_equal(whole[0], merged[0])


def test_scores_match_float64_cosine():
    rng = np.random.default_rng(1)
    Q = rng.standard_normal((4, 128)).astype(np.float32)
    D = rng.standard_normal((50, 128)).astype(np.float32)
    s = exact_cosine_scores(Q, D)
    ref = (Q.astype(np.float64) @ D.T.astype(np.float64)) / (np.linalg.norm(Q, axis=1)[:, None] * np.linalg.norm(D, axis=1)[None])
    assert np.abs(s - ref).max() < 2e-6
