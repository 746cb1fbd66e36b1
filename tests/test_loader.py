"""Host batch pipeline (dssm_b200/loader.py): the direct three-memcpy assembly equals pull_batch + stacking
(utils/utils.py:45-61), the epoch length follows new_dssm.py:46, and the ring never overwrites a buffer that may still be
in flight."""
import numpy as np
import pytest
import scipy.sparse as sp

from dssm_b200 import Config, pull_batch, stack_feed
from dssm_b200.loader import HostBatchLoader, fill_stacked, reference_epoch_steps


def _epoch(n_q, D, NEG, seed=0, dtype=np.int64):
    rng = np.random.default_rng(seed)
    mk = lambda n: sp.random(n, D, density=0.03, format="csr", random_state=rng, data_rvs=lambda k: rng.integers(1, 4, k)).astype(dtype)
    return mk(n_q), mk(n_q), mk(n_q * NEG)


def test_epoch_steps_follow_the_reference():
    assert reference_epoch_steps(1000, 100) == 9  # int(n / BS) - 1
    assert reference_epoch_steps(150, 100) == 0


@pytest.mark.parametrize("dtype", [np.int64, np.float64])
def test_loader_equals_pull_batch(dtype):
    B, NEG, D = 8, 3, 200
    q, p, n = _epoch(50, D, NEG, dtype=dtype)
    p[3] = 0  # an empty row
    p.eliminate_zeros()
    conf = Config(TRIGRAM_D=D, query_BS=B, NEG=NEG, layers=(16, 8))
    ld = HostBatchLoader(q, p, n, B, NEG, pin=False)
    assert len(ld) == int(50 / B) - 1
    got = [(ip.numpy().copy(), ix.numpy()[:nnz].copy(), vl.numpy()[:nnz].copy(), nnz) for ip, ix, vl, nnz in ld]
    assert len(got) == len(ld)
    for b, (ip, ix, vl, nnz) in enumerate(got):
        ref = stack_feed(pull_batch(True, q, p, n, b, B, conf=conf), conf)
        assert nnz == ref.nnz and ip.dtype == np.int32 and vl.dtype == np.float32
        assert np.array_equal(ip, ref.indptr) and np.array_equal(ix, ref.indices) and np.array_equal(vl, ref.values)
    assert ld.max_nnz == max(g[3] for g in got)


def test_loader_shuffled_ids_and_bounds():
    B, NEG, D = 4, 2, 64
    q, p, n = _epoch(30, D, NEG, seed=1)
    conf = Config(TRIGRAM_D=D, query_BS=B, NEG=NEG, layers=(8, 4))
    ld = HostBatchLoader(q, p, n, B, NEG, pin=False, depth=4)
    order = [5, 0, 3, 3, 1]
    for b, (ip, ix, vl, nnz) in zip(order, ld.iterate(order)):
        ref = stack_feed(pull_batch(True, q, p, n, b, B, conf=conf), conf)
        assert np.array_equal(ix.numpy()[:nnz], ref.indices)
    with pytest.raises(IndexError):
        next(iter(ld.iterate([99])))
    with pytest.raises(ValueError):
        HostBatchLoader(q, p, n[:-1], B, NEG, pin=False)


def test_ring_does_not_overwrite_batches_in_flight():
    """A consumer that looks at a batch again while holding the next two (two steps in flight + the one being issued) must
    still see it intact."""
    B, NEG, D = 4, 2, 64
    q, p, n = _epoch(80, D, NEG, seed=2)
    ld = HostBatchLoader(q, p, n, B, NEG, pin=False, depth=4)
    held = []
    for i, item in enumerate(ld):
        held.append((item, item[1].numpy()[:item[3]].copy()))
        for (ip, ix, vl, nnz), snap in held[-3:]:
            assert np.array_equal(ix.numpy()[:nnz], snap)
    assert len(held) == len(ld)


def test_fill_rejects_oversized_batch():
    q, p, n = _epoch(10, 50, 1, seed=3)
    ip, ix, vl = np.zeros(31, np.int32), np.zeros(2, np.int32), np.zeros(2, np.float32)
    with pytest.raises(ValueError):
        fill_stacked(((q, 0, 10), (p, 0, 10), (n, 0, 10)), ip, ix, vl)


def test_fill_stacked_equals_vstack_on_ragged_inputs():
    """Property: for any row ranges (empty rows, empty ranges, duplicates across parts) the three-run copy equals
    scipy's vstack of the same slices."""
    from hypothesis import given, settings, strategies as st

    rng = np.random.default_rng(7)
    mats = [sp.random(40, 30, density=d, format="csr", random_state=rng, dtype=np.float64) for d in (0.0, 0.05, 0.3)]
    for m in mats:
        m.sort_indices()

    @settings(max_examples=60, deadline=None)
    @given(st.lists(st.tuples(st.integers(0, 2), st.integers(0, 40), st.integers(0, 40)), min_size=1, max_size=4))
    def check(spec):
        parts = [(mats[i], min(a, b), max(a, b)) for i, a, b in spec]
        R = sum(hi - lo for _, lo, hi in parts)
        nnz_cap = sum(int(m.indptr[hi] - m.indptr[lo]) for m, lo, hi in parts)
        ip, ix, vl = np.full(R + 1, -1, np.int32), np.zeros(max(nnz_cap, 1), np.int32), np.zeros(max(nnz_cap, 1), np.float32)
        nnz = fill_stacked(parts, ip, ix, vl)
        ref = sp.vstack([m[lo:hi] for m, lo, hi in parts], format="csr") if R else sp.csr_matrix((0, 30))
        assert nnz == ref.nnz and np.array_equal(ip, ref.indptr.astype(np.int32))
        assert np.array_equal(ix[:nnz], ref.indices) and np.array_equal(vl[:nnz], ref.data.astype(np.float32))

    check()


def test_native_assembly_equals_python_assembly_for_every_value_type():
    """dssm_host_stack_csr (C, one call per batch, GIL released) writes the same stacked CSR as the NumPy path, for the
    value types the reference feeds: int64 counts (CountVectorizer), float64 tf-idf (dssm_tf_idf.py:37), float32
    (data_input.py:157-159); and refuses a batch larger than the buffers."""
    import scipy.sparse as sp

    from dssm_b200._lib import DssmError
    from dssm_b200.loader import HostBatchLoader
    from tests.helpers import random_csr

    rng = np.random.default_rng(3)
    B, NEG, nb, D = 16, 3, 5, 300
    base = [random_csr(rng, B * nb, D), random_csr(rng, B * nb, D), random_csr(rng, B * nb * NEG, D)]
    for dt in (np.int64, np.float64, np.float32):
        mats = [sp.csr_matrix(m, dtype=dt) for m in base]
        a = HostBatchLoader(*mats, B, NEG, pin=False, native=True, copy_threads=3)
        b = HostBatchLoader(*mats, B, NEG, pin=False, native=False)
        for (ip1, ix1, vl1, n1), (ip2, ix2, vl2, n2) in zip(a, b):
            assert n1 == n2
            assert np.array_equal(ip1.numpy(), ip2.numpy())
            assert np.array_equal(ix1.numpy()[:n1], ix2.numpy()[:n2]) and np.array_equal(vl1.numpy()[:n1], vl2.numpy()[:n2])
    small = HostBatchLoader(*base, B, NEG, pin=False, max_nnz=5)
    with pytest.raises(DssmError):
        list(small)
