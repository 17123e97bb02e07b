"""CPU check of the split-precision scheme the tcgen05 RealNVP kernel computes with (DESIGN.md 4.1,
harmonic_b200/csrc/hb_realnvp_tc.cu at H2_SCALE): numpy float16 stands in for the tensor core's exact FP16 products.

  activation a -> hs = fp16(a * 2^-11) (round to nearest), lo = fp16(a - 2^11 * hs)   (compensated split)
  weight     w -> w_hi = top 11 significant bits, w_lo = w - w_hi, stored as fp16(w_hi), fp16(w_hi * 2^11), fp16(w_lo * 2^11)
  a * w ~= hs * (w_hi 2^11) + hs * (w_lo 2^11) + lo * w_hi        (lo * w_lo, <= 2^-22 relative, dropped)

The test pins the claims the design rests on: float32-equivalent products over the stated range, graceful behaviour
in the half subnormal range (absolute error <= 2^-22 |w| for |a| < 0.125), and what a cheaper split would cost."""
import numpy as np


def split_activation(a):
    a = a.astype(np.float32)
    hs = (a * np.float32(2.0 ** -11)).astype(np.float16)
    lo = (a - hs.astype(np.float32) * np.float32(2048.0)).astype(np.float16)  # exact difference in float32
    return hs.astype(np.float64), lo.astype(np.float64)


def split_weight(w):
    w = w.astype(np.float32)
    hi = (w.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)  # 11 significant bits
    lo = w - hi
    b_main = (hi * np.float32(2048.0)).astype(np.float16)
    b_lo = (lo * np.float32(2048.0)).astype(np.float16)
    b_hi = hi.astype(np.float16)
    big = np.abs(hi) >= np.float32(2.0 ** -14)
    assert np.array_equal(b_hi.astype(np.float32)[big], hi[big])  # an 11-bit significand is exactly a (normal) half
    return b_main.astype(np.float64), b_lo.astype(np.float64), b_hi.astype(np.float64)


def three_products(a, w):
    hs, lo = split_activation(a)
    b_main, b_lo, b_hi = split_weight(w)
    return hs * b_main + hs * b_lo + lo * b_hi


def test_products_are_float32_equivalent_in_the_normal_range():
    rng = np.random.default_rng(0)
    a = (rng.uniform(0.125, 1.0, 200000) * 2.0 ** rng.integers(0, 27, 200000) * rng.choice([-1.0, 1.0], 200000)).astype(np.float32)
    w = (rng.uniform(0.5, 1.0, 200000) * 2.0 ** rng.integers(-9, 5, 200000) * rng.choice([-1.0, 1.0], 200000)).astype(np.float32)
    exact = a.astype(np.float64) * w.astype(np.float64)
    rel = np.abs(three_products(a, w) - exact) / np.abs(exact)
    assert rel.max() < 2.0 ** -20, rel.max()         # worst case: dropped lo*w_lo (2^-22) + rounding of w_lo and lo to halves
    assert np.sqrt(np.mean(rel ** 2)) < 2.0 ** -22.5  # rms below a float32 rounding (2^-24 .. 2^-23 relative)


def test_small_activations_keep_an_absolute_error_floor():
    """Below |a| = 0.125 hs is a subnormal half; lo carries what hs lost, down to the half subnormal spacing."""
    rng = np.random.default_rng(1)
    a = (rng.uniform(-0.125, 0.125, 100000)).astype(np.float32)
    w = rng.uniform(-2.0, 2.0, 100000).astype(np.float32)
    exact = a.astype(np.float64) * w.astype(np.float64)
    err = np.abs(three_products(a, w) - exact)
    assert (err <= 2.0 ** -22 * np.abs(w)).all()     # i.e. 2^-22 of a unit-sized contribution (measured: 0.8 of that)


def test_range_limit_of_the_scaled_halves():
    a = np.array([1.3e8, -1.3e8, 6.0e4], dtype=np.float32)
    hs, _ = split_activation(a)
    assert np.isfinite(hs).all()                      # 1.3e8 * 2^-11 = 63477 < 65504
    with np.errstate(over="ignore", invalid="ignore"):
        assert np.isinf(split_activation(np.array([1.4e8], dtype=np.float32))[0]).all()  # the kernel saturates here
    b_main, _, _ = split_weight(np.array([15.9, -15.9], dtype=np.float32))
    assert np.isfinite(b_main).all()                  # |w| < 16 is what hb_flow_create checks for the tensor path


def test_cheaper_splits_are_not_float32_equivalent():
    """Dropping either correction term leaves a 2^-12-level error: three products are the minimum."""
    rng = np.random.default_rng(2)
    a = rng.uniform(0.5, 4.0, 50000).astype(np.float32)
    w = rng.uniform(-1.0, 1.0, 50000).astype(np.float32)
    exact = a.astype(np.float64) * w.astype(np.float64)
    hs, lo = split_activation(a)
    b_main, b_lo, b_hi = split_weight(w)
    for approx in (hs * b_main + hs * b_lo, hs * b_main + lo * b_hi):
        rel = np.abs(approx - exact) / np.abs(exact)
        assert rel.max() > 2.0 ** -13
