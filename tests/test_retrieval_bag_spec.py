"""The selection rule of the tensor-core corpus top-k (csrc/topk_tc.cu: topk_approx_select_kernel), as a NumPy model.

Between filter passes the kernels know only APPROXIMATE cosines (|approx - exact| <= m).  Per query they keep a bag of
candidates, take tau = any lower bound of the k-th largest approximate score in the bag, keep the entries with
approx >= tau - 2m and let the next pass admit docs with approx >= tau - 2m.  Claim: whatever the chunking
and however loose the lower bounds, the final bag contains the exact top-k, so ONE exact rescoring of the bag at the end
yields the oracle's answer.  This test states the claim executable; the GPU tests check the kernels against the oracle."""
import numpy as np
from hypothesis import given, settings, strategies as st


def bag_after_passes(approx, k, m, chunks, slack_rng):
    tau, bag = -np.inf, np.zeros(0, dtype=np.int64)
    lo = 0
    for size in chunks:
        ids = np.arange(lo, min(lo + size, approx.size))
        lo += size
        bag = np.concatenate([bag, ids[approx[ids] >= tau - 2 * m]])  # the filter pass
        if bag.size >= k:  # the select kernel: a LOWER BOUND of the k-th largest approximate score of the bag ...
            kth = np.sort(approx[bag])[-k]
            tau = kth - slack_rng.uniform(0.0, 3 * m)  # ... as loose as the bucket width makes it (even below the last tau)
            bag = bag[approx[bag] >= tau - 2 * m]
        if lo >= approx.size:
            break
    return bag


@settings(max_examples=200, deadline=None)
@given(seed=st.integers(0, 2**31 - 1), nd=st.integers(1, 600), k=st.integers(1, 40), m=st.floats(1e-4, 5e-2),
       first=st.integers(1, 64), growth=st.integers(1, 5))
def test_bag_contains_the_exact_topk(seed, nd, k, m, first, growth):
    rng = np.random.default_rng(seed)
    k = min(k, nd)
    exact = np.round(rng.uniform(0.5, 1.0, nd), 2).astype(np.float64)  # coarse values: plenty of exact ties
    approx = exact + rng.uniform(-m, m, nd)
    chunks, size, tot = [], first, 0
    while tot < nd:
        chunks.append(size)
        tot += size
        size = max(1, size * growth)
    bag = bag_after_passes(approx, k, m, chunks, rng)
    order = np.lexsort((np.arange(nd), -exact))[:k]  # (score desc, id asc): tf.nn.top_k(sorted=True)
    assert np.isin(order, bag).all()
    # and the exact rescoring of the bag alone reproduces that list
    b = np.sort(bag)
    again = b[np.lexsort((b, -exact[b]))[:k]]
    assert np.array_equal(again, order)
