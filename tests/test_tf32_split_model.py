"""The operand split of the 3xTF32 dense path (csrc/tc_common.cuh: tf32_rna; csrc/fc_tc.cu), as a NumPy model.

hi = rna_tf32(x) is computed on the device with two integer instructions, (bits + 0x1000) & 0xFFFFE000, and
lo = rna_tf32(x - hi); a product is accumulated as a_hi*b_lo + a_lo*b_hi + a_hi*b_hi.  Checked here: (1) the integer trick
equals round-to-nearest (ties away from zero) to a 10-bit mantissa, (2) hi and lo are exactly representable in tf32 and
|x - hi - lo| <= 2^-22 |x|, (3) the three-term product is within 2^-20 of the exact product -- the bound behind the 1e-5
parity of gemm_mode='tc_3xtf32' (DESIGN.md 4a); the single-pass mode keeps hi only: 2^-11 per operand."""
import numpy as np


def tf32_rna(x):
    b = np.asarray(x, np.float32).view(np.uint32)
    return ((b + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)


def test_integer_trick_is_round_to_nearest_ties_away():
    rng = np.random.default_rng(0)
    x = (rng.standard_normal(200000) * np.exp(rng.uniform(-20, 20, 200000))).astype(np.float32)
    hi = tf32_rna(x)
    # reference: scale the mantissa to 11 significant bits and round half away from zero in float64
    m, e = np.frexp(x.astype(np.float64))
    ref = np.ldexp(np.sign(m) * np.floor(np.abs(m) * 2048 + 0.5) / 2048, e)
    assert np.array_equal(hi.astype(np.float64), ref)
    assert np.all((hi.view(np.uint32) & np.uint32(0x1FFF)) == 0)  # 13 low mantissa bits clear: a tf32 value


def test_hi_lo_split_and_three_term_product():
    rng = np.random.default_rng(1)
    a = (rng.standard_normal(100000) * np.exp(rng.uniform(-8, 8, 100000))).astype(np.float32)
    b = (rng.standard_normal(100000) * np.exp(rng.uniform(-8, 8, 100000))).astype(np.float32)
    a_hi, b_hi = tf32_rna(a), tf32_rna(b)
    a_lo, b_lo = tf32_rna(a - a_hi), tf32_rna(b - b_hi)
    for x, hi, lo in ((a, a_hi, a_lo), (b, b_hi, b_lo)):
        assert np.all(np.abs(x.astype(np.float64) - hi.astype(np.float64) - lo.astype(np.float64)) <= 2.0**-22 * np.abs(x))
    exact = a.astype(np.float64) * b.astype(np.float64)
    f = np.float64
    three = a_hi.astype(f) * b_lo.astype(f) + a_lo.astype(f) * b_hi.astype(f) + a_hi.astype(f) * b_hi.astype(f)
    one = a_hi.astype(f) * b_hi.astype(f)
    assert np.all(np.abs(three - exact) <= 2.0**-20 * np.abs(exact))
    assert np.all(np.abs(one - exact) <= 2.0**-10 * np.abs(exact))
    assert np.abs(one - exact).max() / np.abs(exact).max() > 2.0**-16  # and the single pass really is that much coarser
