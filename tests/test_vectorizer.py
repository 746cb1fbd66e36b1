"""CountVectorizerCompat against scikit-learn's CountVectorizer as the reference configures it (new_dssm.py:33-44):
same vocabulary, same int64 CSR (indptr / sorted indices / counts), OOV dropped, feeding HostBatchLoader."""
import numpy as np
import pytest

from dssm_b200.vectorizer import TOKEN_PATTERN, CountVectorizerCompat, char_split

sklearn_text = pytest.importorskip("sklearn.feature_extraction.text")


def _corpus(rng, n, alphabet, lo=1, hi=12):
    docs = []
    for _ in range(n):
        k = int(rng.integers(lo, hi))
        docs.append(char_split("".join(rng.choice(alphabet, size=k))))
    return docs


ALPHABET = list("手机壳苹果华为小米充电器数据线耳机abcXYZ019_-，。 ")


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_matches_sklearn_token_for_token(seed):
    rng = np.random.default_rng(seed)
    fit_docs = _corpus(rng, 300, ALPHABET)
    new_docs = _corpus(rng, 100, ALPHABET + list("未见字Q")) + ["", "   ", "，。"]
    ref = sklearn_text.CountVectorizer(token_pattern=TOKEN_PATTERN).fit(fit_docs)
    mine = CountVectorizerCompat().fit(fit_docs)
    assert mine.vocabulary_ == {k: int(v) for k, v in ref.vocabulary_.items()}
    assert list(mine.get_feature_names_out()) == list(ref.get_feature_names_out())
    for docs in (fit_docs, new_docs):
        a, b = mine.transform(docs), ref.transform(docs)
        assert a.shape == b.shape and a.dtype == b.dtype == np.int64
        b.sort_indices()
        assert np.array_equal(a.indptr, b.indptr) and np.array_equal(a.indices, b.indices) and np.array_equal(a.data, b.data)
        assert a.has_sorted_indices


def test_max_features_keeps_the_most_frequent_terms():
    docs = ["a a a b b c", "a b d", "e"]
    v = CountVectorizerCompat(max_features=2).fit(docs)
    assert v.vocabulary_ == {"a": 0, "b": 1}
    ref = sklearn_text.CountVectorizer(token_pattern=TOKEN_PATTERN, max_features=2).fit(docs)
    assert v.vocabulary_ == {k: int(x) for k, x in ref.vocabulary_.items()}


def test_loader_from_texts_feeds_reference_shaped_batches():
    from dssm_b200.loader import HostBatchLoader

    rng = np.random.default_rng(0)
    B, NEG, n = 4, 2, 16
    q, d, neg = _corpus(rng, n, ALPHABET), _corpus(rng, n, ALPHABET), _corpus(rng, n * NEG, ALPHABET)
    vec = CountVectorizerCompat().fit(q + d + neg)
    ld = HostBatchLoader.from_texts(vec, q, d, neg, B, NEG, pin=False)
    assert len(ld) == n // B - 1  # new_dssm.py:46
    Xq, Xd, Xn = vec.transform(q), vec.transform(d), vec.transform(neg)
    for b, (ip, ix, vl, nnz) in enumerate(ld):
        import scipy.sparse as sp

        want = sp.vstack([Xq[b * B:(b + 1) * B], Xd[b * B:(b + 1) * B], Xn[b * B * NEG:(b + 1) * B * NEG]], format="csr")
        want.sort_indices()
        assert nnz == want.nnz
        assert np.array_equal(ip.numpy(), want.indptr) and np.array_equal(ix.numpy()[:nnz], want.indices)
        assert np.array_equal(vl.numpy()[:nnz], want.data.astype(np.float32))
