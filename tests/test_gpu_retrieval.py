"""Corpus cosine top-k on the GPU: bit-exact ids (and scores) against the oracle, ties, zero-norm rows, chunk
boundaries, shard merge."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def make(nq, nd, d, seed):
    rng = np.random.default_rng(seed)
    Q = np.maximum(rng.standard_normal((nq, d)), 0).astype(np.float32)
    D = np.maximum(rng.standard_normal((nd, d)), 0).astype(np.float32)
    return Q, D


@pytest.mark.parametrize("nq,nd,d,k", [(64, 3000, 128, 100), (5, 70, 16, 10), (130, 20000, 128, 100), (3, 40, 7, 40)])
def test_topk_bit_exact(nq, nd, d, k):
    from dssm_b200 import corpus_topk
    from oracle import corpus_topk_oracle

    Q, D = make(nq, nd, d, nq + nd)
    D[nd // 2] = D[3]  # duplicate doc -> exact tie
    D[7] = 0  # zero-norm doc -> NaN -> ranks last
    if nq > 4:
        Q[4] = 0  # zero-norm query: every score NaN -> ids 0..k-1
    s, i = corpus_topk(torch.from_numpy(Q).cuda(), torch.from_numpy(D).cuda(), k, id_offset=1000)
    rs, ri = corpus_topk_oracle(Q, D, k, id_offset=1000)
    assert np.array_equal(i.cpu().numpy(), ri)
    assert np.array_equal(s.cpu().numpy(), rs)


def test_topk_merge_equals_whole_corpus():
    from dssm_b200 import corpus_topk, topk_merge
    from dssm_b200.retrieval import shard_range
    from oracle import corpus_topk_oracle

    Q, D = make(33, 5003, 128, 0)
    Qd, Dd = torch.from_numpy(Q).cuda(), torch.from_numpy(D).cuda()
    parts = []
    for r in range(8):
        lo, hi = shard_range(5003, r, 8)
        parts.append(corpus_topk(Qd, Dd[lo:hi].contiguous(), 100, id_offset=lo))
    s, i = topk_merge(torch.stack([p[0] for p in parts]), torch.stack([p[1] for p in parts]))
    rs, ri = corpus_topk_oracle(Q, D, 100)
    assert np.array_equal(i.cpu().numpy(), ri) and np.array_equal(s.cpu().numpy(), rs)
    ws, wi = corpus_topk(Qd, Dd, 100)
    assert torch.equal(wi, i) and torch.equal(ws, s)


@pytest.mark.parametrize("nq,nd,k", [(64, 40000, 100), (130, 70001, 10)])
def test_topk_tensor_core_filter_bit_exact_vs_oracle(nq, nd, k):
    """tcgen05 tf32 filter + exact rescoring returns the oracle's ids and scores bit for bit (ties, zero rows, a doc
    range that is not a multiple of the 128-doc tile)."""
    from dssm_b200 import corpus_topk
    from oracle import corpus_topk_oracle

    Q, D = make(nq, nd, 128, nq + nd)
    D[nd - 5] = D[3]  # duplicate far apart (one in the exact seed pass, one in the tensor-core pass)
    D[30000] = D[20000]
    D[17000] = 0  # zero-norm doc in the tensor-core range
    Q[1] = Q[0]
    from dssm_b200 import retrieval

    s, i = corpus_topk(torch.from_numpy(Q).cuda(), torch.from_numpy(D).cuda(), k, id_offset=7, method="tc")
    assert retrieval.LAST_CALL == {"method": "tc", "fallback": False}  # the tensor-core path itself produced this
    rs, ri = corpus_topk_oracle(Q, D, k, id_offset=7)
    assert np.array_equal(i.cpu().numpy(), ri)
    assert np.array_equal(s.cpu().numpy(), rs)


def test_topk_tensor_core_equals_exact_path_large():
    """300k docs, several filter passes with growing chunks: same result as the all-pairs exact kernels."""
    from dssm_b200 import corpus_topk

    g = torch.Generator(device="cuda").manual_seed(0)
    Q = torch.relu(torch.randn((256, 128), generator=g, device="cuda"))
    D = torch.relu(torch.randn((300000, 128), generator=g, device="cuda"))
    s1, i1 = corpus_topk(Q, D, 100, method="exact")
    from dssm_b200 import retrieval

    s2, i2 = corpus_topk(Q, D, 100, method="tc")
    assert retrieval.LAST_CALL == {"method": "tc", "fallback": False}
    assert torch.equal(i1, i2) and torch.equal(s1, s2)
    # signed (not post-relu) embeddings: thresholds stay positive at the top of the ranking, the filter still applies
    Qs, Ds = torch.randn((128, 128), generator=g, device="cuda"), torch.randn((100000, 128), generator=g, device="cuda")
    s3, i3 = corpus_topk(Qs, Ds, 50, method="exact")
    s4, i4 = corpus_topk(Qs, Ds, 50, method="tc")
    assert retrieval.LAST_CALL["method"] == "tc"  # may fall back on signed data; the result must be right either way
    assert torch.equal(i3, i4) and torch.equal(s3, s4)


@pytest.mark.parametrize("nq,nd,k", [(64, 40000, 100), (130, 70001, 10), (300, 5000, 100), (7, 129, 50)])
def test_topk_bf16_index_bit_exact_vs_oracle(nq, nd, k):
    """CorpusIndex (normalised bf16 tile image, built once) + tcgen05 bf16 filter with two query tiles per CTA + exact fp32
    rescoring: the oracle's ids and scores bit for bit -- ties across the seed / filter boundary, zero-norm docs and
    queries, a corpus that is not a multiple of the 128-doc tile, a query count that is not a multiple of 256."""
    from dssm_b200 import CorpusIndex, corpus_topk, retrieval
    from oracle import corpus_topk_oracle

    Q, D = make(nq, nd, 128, nq + nd)
    D[nd - 5] = D[3]
    if nd > 30000:
        D[30000] = D[20000]
        D[17000] = 0
    D[100] = 0
    Q[1] = Q[0]
    if nq > 4:
        Q[4] = 0
    Dd = torch.from_numpy(D).cuda()
    index = CorpusIndex(Dd)
    assert index.nbytes >= 256 * nd
    for rep in range(2):  # the index is reusable across query batches
        s, i = corpus_topk(torch.from_numpy(Q).cuda(), index, k, id_offset=7)
        assert retrieval.LAST_CALL == {"method": "bf16", "fallback": False}
        rs, ri = corpus_topk_oracle(Q, D, k, id_offset=7)
        assert np.array_equal(i.cpu().numpy(), ri)
        assert np.array_equal(s.cpu().numpy(), rs)


def test_topk_bf16_index_equals_exact_path_large_and_signed():
    from dssm_b200 import CorpusIndex, corpus_topk, retrieval

    g = torch.Generator(device="cuda").manual_seed(1)
    Q = torch.relu(torch.randn((512, 128), generator=g, device="cuda"))
    D = torch.relu(torch.randn((300000, 128), generator=g, device="cuda"))
    s1, i1 = corpus_topk(Q, D, 100, method="exact")
    s2, i2 = corpus_topk(Q, CorpusIndex(D), 100)
    assert retrieval.LAST_CALL == {"method": "bf16", "fallback": False}
    assert torch.equal(i1, i2) and torch.equal(s1, s2)
    Qs, Ds = torch.randn((128, 128), generator=g, device="cuda"), torch.randn((100000, 128), generator=g, device="cuda")
    s3, i3 = corpus_topk(Qs, Ds, 50, method="exact")
    s4, i4 = corpus_topk(Qs, CorpusIndex(Ds), 50)
    assert retrieval.LAST_CALL["method"] == "bf16"
    assert torch.equal(i3, i4) and torch.equal(s3, s4)


@pytest.mark.parametrize("k", [1, 300, 700])
def test_topk_filters_other_k(k):
    """k = 1, and k beyond 160 (chunks then grow x2, the per-query bag is larger; k = 700 may overflow the candidate lists
    and fall back -- the result must be the exact path's either way), for both filters."""
    from dssm_b200 import CorpusIndex, corpus_topk, retrieval

    g = torch.Generator(device="cuda").manual_seed(10 + k)
    Q = torch.relu(torch.randn((96, 128), generator=g, device="cuda"))
    D = torch.relu(torch.randn((90000, 128), generator=g, device="cuda"))
    D[12345] = 0
    D[77] = D[60000]
    s1, i1 = corpus_topk(Q, D, k, method="exact")
    s2, i2 = corpus_topk(Q, D, k, method="tc")
    if k <= 300:
        assert retrieval.LAST_CALL == {"method": "tc", "fallback": False}
    assert torch.equal(i1, i2) and torch.equal(s1, s2)
    s3, i3 = corpus_topk(Q, CorpusIndex(D), k)
    if k <= 300:
        assert retrieval.LAST_CALL == {"method": "bf16", "fallback": False}
    assert torch.equal(i1, i3) and torch.equal(s1, s3)


def test_topk_filters_fewer_finite_docs_than_k():
    """Most docs have zero norm (cosine NaN = -inf): fewer than k finite scores per query.  The filters do not collect the
    -inf fillers, corpus_topk notices and falls back to the exact kernels: ids and scores equal the exact path."""
    from dssm_b200 import CorpusIndex, corpus_topk, retrieval

    g = torch.Generator(device="cuda").manual_seed(3)
    Q = torch.relu(torch.randn((40, 128), generator=g, device="cuda"))
    D = torch.zeros((40000, 128), device="cuda")
    D[::997] = torch.relu(torch.randn((len(range(0, 40000, 997)), 128), generator=g, device="cuda"))  # 41 real docs
    s1, i1 = corpus_topk(Q, D, 100, method="exact")
    for docs, method in ((D, "tc"), (CorpusIndex(D), "auto")):
        s2, i2 = corpus_topk(Q, docs, 100, method=method)
        assert retrieval.LAST_CALL["fallback"] is True
        assert torch.equal(i1, i2) and torch.equal(s1, s2)
