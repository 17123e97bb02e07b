"""CPU check of the arithmetic the tcgen05 RealNVP kernel computes with (the header comment of
harmonic_b200/csrc/hb_realnvp_tc.cu): numpy float16 stands in for the tensor core's exact FP16 products.

  activation a -> hs = fp16(a) (round to nearest), lo = fp16(a - hs)          (the difference is exact in float32)
  weight     w -> w_hi = top 11 significant bits, w_lo = w - w_hi, stored as fp16(w_hi * S), fp16(w_lo * S)
  S * a * w ~= hs * (w_hi S) + lo * (w_hi S) + hs * (w_lo S)                   (lo * w_lo, <= 2^-22 relative, dropped)
  leaky_relu(h) = a h + b relu(h) (a = 0.01, b = 0.99)  =>  leaky_relu(x' W0') W2 + b2 = x' V + relu(x' W0') U,
      V = a W0' W2 (+ b2), U = b W2
  hidden activation r = relu(h) -> hs = fp16(r) rounded TOWARDS ZERO, lo = fp16(max(h - hs, 0))   (the clamp is a modifier
      of both conversion instructions: h < 0 gives hs = 0 and a negative "remainder" h that the second one clamps)

The tests pin the claims the design rests on: float32-equivalent products over the guarded range, graceful behaviour
in the half subnormal range, what a cheaper split would cost, and that the linear / relu decomposition of a
coupling MLP evaluated with these products stays at the float32 level against the float64 MLP."""
import numpy as np

S = np.float32(32.0)  # SH = SU in the kernel


def split_activation(a):
    a = a.astype(np.float32)
    hs = a.astype(np.float16)
    lo = (a - hs.astype(np.float32)).astype(np.float16)  # exact difference in float32
    return hs.astype(np.float64), lo.astype(np.float64)


def split_weight(w, scale=S):
    w = w.astype(np.float32)
    hi = (w.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)  # 11 significant bits
    lo = w - hi
    b_hi = (hi * scale).astype(np.float16)
    b_lo = (lo * scale).astype(np.float16)
    big = np.abs(hi * scale) >= np.float32(2.0 ** -14)
    assert np.array_equal(b_hi.astype(np.float32)[big], (hi * scale)[big])  # an 11-bit significand is exactly a (normal) half
    return b_hi.astype(np.float64), b_lo.astype(np.float64)


def split_relu_rz(h):
    """E1 of the kernel: cvt.rz.relu.f16x2.f32, h - hs by one mixed-precision FMA (exact), cvt.rn.relu.f16x2.f32."""
    h = h.astype(np.float32)
    r = np.maximum(h, np.float32(0))
    hs = r.astype(np.float16)
    over = hs.astype(np.float32) > r
    hs = np.where(over, np.nextafter(hs, np.float16(0)), hs).astype(np.float16)
    rest = h - hs.astype(np.float32)
    assert (rest[h > 0] >= 0).all()  # rounded towards zero: the remainder of a positive value is never negative
    lo = np.maximum(rest, np.float32(0)).astype(np.float16)
    return hs.astype(np.float64), lo.astype(np.float64)


def three_products(a, w):
    hs, lo = split_activation(a)
    b_hi, b_lo = split_weight(w)
    return (hs * b_hi + lo * b_hi + hs * b_lo) / float(S)


def test_products_are_float32_equivalent_in_the_guarded_range():
    rng = np.random.default_rng(0)
    n = 200000
    a = (rng.uniform(0.5, 1.0, n) * 2.0 ** rng.integers(-3, 11, n) * rng.choice([-1.0, 1.0], n)).astype(np.float32)
    w = (rng.uniform(0.5, 1.0, n) * 2.0 ** rng.integers(-6, 11, n) * rng.choice([-1.0, 1.0], n)).astype(np.float32)
    exact = a.astype(np.float64) * w.astype(np.float64)
    rel = np.abs(three_products(a, w) - exact) / np.abs(exact)
    assert rel.max() < 2.0 ** -20, rel.max()         # worst case: dropped lo*w_lo (2^-22) + rounding of w_lo and lo to halves
    assert np.sqrt(np.mean(rel ** 2)) < 2.0 ** -22.5  # rms below a float32 rounding (2^-24 .. 2^-23 relative)


def test_small_activations_keep_an_absolute_error_floor():
    """Below |a| = 0.25 lo is a subnormal half (spacing 2^-24): the error is absolute, ~2^-25 |w| from rounding lo
    plus the weight-side floor of the next test and the 2^-21 |a w| of the normal range."""
    rng = np.random.default_rng(1)
    a = (rng.uniform(-0.25, 0.25, 100000)).astype(np.float32)
    w = rng.uniform(-2.0, 2.0, 100000).astype(np.float32)
    exact = a.astype(np.float64) * w.astype(np.float64)
    err = np.abs(three_products(a, w) - exact)
    assert (err <= 2.0 ** -24 * np.abs(w) + 2.0 ** -29 * np.abs(a) + 2.0 ** -21 * np.abs(exact)).all()


def test_small_weights_keep_an_absolute_error_floor():
    """Below |w| ~ 2^-6 the scaled w_lo is a subnormal half: the error is absolute, 2^-25 / 32 of a unit-sized
    activation -- far below the rounding of the sums such a weight takes part in."""
    rng = np.random.default_rng(4)
    a = rng.uniform(-4.0, 4.0, 100000).astype(np.float32)
    w = rng.uniform(-2.0 ** -6, 2.0 ** -6, 100000).astype(np.float32)
    exact = a.astype(np.float64) * w.astype(np.float64)
    err = np.abs(three_products(a, w) - exact)
    assert (err <= 2.0 ** -28 * np.maximum(np.abs(a), 0.25) + 2.0 ** -21 * np.abs(exact)).all()


def test_range_limits_of_the_halves():
    hs, _ = split_activation(np.array([65000.0, -65000.0], dtype=np.float32))
    assert np.isfinite(hs).all()
    with np.errstate(over="ignore", invalid="ignore"):
        assert np.isinf(split_activation(np.array([7.0e4], dtype=np.float32))[0]).all()  # the range guard's job
    b_hi, _ = split_weight(np.array([2040.0, -2040.0], dtype=np.float32))
    assert np.isfinite(b_hi).all()                    # |w| * 32 < 65504: checked when the weight image is built


def test_cheaper_splits_are_not_float32_equivalent():
    """Dropping either correction term leaves a 2^-12-level error: three products are the minimum."""
    rng = np.random.default_rng(2)
    a = rng.uniform(0.5, 4.0, 50000).astype(np.float32)
    w = rng.uniform(-1.0, 1.0, 50000).astype(np.float32)
    exact = a.astype(np.float64) * w.astype(np.float64)
    hs, lo = split_activation(a)
    b_hi, b_lo = split_weight(w)
    for approx in ((hs * b_hi + hs * b_lo) / float(S), (hs * b_hi + lo * b_hi) / float(S)):
        rel = np.abs(approx - exact) / np.abs(exact)
        assert rel.max() > 2.0 ** -13


def split_matmul(a, w, scale, split=split_activation):
    """(a @ w) * scale with the three-product scheme, float64 accumulation of the exact FP16 products."""
    hs, lo = split(a)
    b_hi, b_lo = split_weight(w, np.float32(scale))
    return hs @ b_hi + lo @ b_hi + hs @ b_lo


def test_coupling_mlp_as_linear_plus_relu_term():
    """O = leaky_relu(x W0 + b0) W2 + b2 evaluated as the kernel does -- H' = 32 [x | 1] W0', O' = 1024 ([x | 1] V +
    relu(H) U) with V = a W0' W2 (+ b2 on the constant row) formed in float64 and rounded to float32, every contraction by
    the three-product FP16 scheme -- against the float64 MLP: the error stays at the level of a float32 evaluation."""
    rng = np.random.default_rng(3)
    n, d, hidden, u = 4000, 32, 128, 32
    x = (rng.standard_normal((n, d)) * 1.5).astype(np.float32)
    W0 = (rng.standard_normal((d, hidden)) * 0.3 / np.sqrt(d) * 4).astype(np.float32)
    b0 = (rng.standard_normal(hidden) * 0.1).astype(np.float32)
    W2 = (rng.standard_normal((hidden, u)) * 0.15 / np.sqrt(hidden) * 8).astype(np.float32)
    b2 = (rng.standard_normal(u) * 0.1).astype(np.float32)
    a_lin, b_abs = 0.01, 1 - 0.01
    x1 = np.concatenate([x, np.ones((n, 1), np.float32)], axis=1)
    W0p = np.concatenate([W0, b0[None, :]], axis=0).astype(np.float64)
    V = a_lin * W0p @ W2.astype(np.float64)
    V[-1] += b2
    U = (b_abs * W2.astype(np.float64)).astype(np.float32)
    # kernel arithmetic
    H32 = split_matmul(x1, W0p.astype(np.float32), 32.0).astype(np.float32)  # what E1 reads from the accumulator
    Olin = split_matmul(x1, V.astype(np.float32), 1024.0)
    Oabs = split_matmul(H32, U, 32.0, split_relu_rz)
    got = (Olin + Oabs) / 1024.0
    # float64 truth and a float32 evaluation of the reference formula
    H = x1.astype(np.float64) @ W0p
    want = np.where(H > 0, H, 0.01 * H) @ W2.astype(np.float64) + b2
    H_f32 = (x.astype(np.float32) @ W0 + b0).astype(np.float32)
    ref32 = (np.where(H_f32 > 0, H_f32, np.float32(0.01) * H_f32).astype(np.float32) @ W2 + b2).astype(np.float32)
    err, err32 = np.abs(got - want), np.abs(ref32.astype(np.float64) - want)
    scale = np.abs(np.where(H > 0, H, 0.01 * H)) @ np.abs(W2.astype(np.float64)) + np.abs(b2)  # conditioning of the sum
    assert (err / scale).max() < 2.0 ** -20
    assert np.sqrt(np.mean(err ** 2)) < 3.0 * np.sqrt(np.mean(err32 ** 2)) + 1e-7


def test_relu_split_rounded_towards_zero_keeps_21_bits():
    """hs rounded towards zero leaves a remainder below one half-precision ulp of r (11 significant bits after the
    second conversion): r = hs + lo to 2^-21 relative in the normal range, exactly zero for h <= 0, NaN stays NaN."""
    rng = np.random.default_rng(5)
    h = (rng.standard_normal(200000) * np.exp(rng.uniform(np.log(0.25), np.log(6.0e4), 200000))).astype(np.float32)
    h = h[np.abs(h) < 65000]
    hs, lo = split_relu_rz(h)
    r = np.maximum(h.astype(np.float64), 0)
    assert (hs[h <= 0] == 0).all() and (lo[h <= 0] == 0).all()
    pos = h > 0.25
    assert (np.abs(hs + lo - r)[pos] <= 2.0 ** -21 * r[pos]).all()
    hs, lo = split_relu_rz(np.array([np.nan, -np.inf], np.float32))
    assert np.isnan(hs[0]) and hs[1] == 0 and lo[1] == 0


def test_register_kernel_linear_plus_relu_form():
    """realnvp_small_kernel evaluates a coupling MLP as y V + v0 + relu(y W0 + b0) (0.99 W2) with
    V = 0.01 W0 W2 and v0 = b2 + 0.01 b0 W2 formed in float64 and rounded to float32 once: in float32 arithmetic the
    form is as close to the float64 MLP as the float32 evaluation of leaky_relu(y W0 + b0) W2 + b2 itself."""
    rng = np.random.default_rng(11)
    n, d, hidden, u = 5000, 2, 128, 3
    y = (rng.standard_normal((n, d)) * 1.5).astype(np.float32)
    W0 = (rng.standard_normal((d, hidden)) * 0.6).astype(np.float32)
    b0 = (rng.standard_normal(hidden) * 0.3).astype(np.float32)
    W2 = (rng.standard_normal((hidden, u)) * 0.1).astype(np.float32)
    b2 = (rng.standard_normal(u) * 0.1).astype(np.float32)
    V = (0.01 * W0.astype(np.float64) @ W2.astype(np.float64)).astype(np.float32)
    v0 = (b2.astype(np.float64) + 0.01 * b0.astype(np.float64) @ W2.astype(np.float64)).astype(np.float32)
    Wc = (0.99 * W2.astype(np.float64)).astype(np.float32)
    H32 = (y @ W0 + b0).astype(np.float32)
    got = (y @ V + v0 + np.maximum(H32, np.float32(0)) @ Wc).astype(np.float64)
    ref32 = (np.where(H32 > 0, H32, np.float32(0.01) * H32).astype(np.float32) @ W2 + b2).astype(np.float64)
    H = y.astype(np.float64) @ W0.astype(np.float64) + b0
    want = np.where(H > 0, H, 0.01 * H) @ W2.astype(np.float64) + b2
    err, err32 = np.abs(got - want), np.abs(ref32 - want)
    assert err.max() < 2e-6 and np.sqrt(np.mean(err ** 2)) < 1.5 * np.sqrt(np.mean(err32 ** 2)) + 1e-8
