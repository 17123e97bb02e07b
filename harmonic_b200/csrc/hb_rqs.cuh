// Device helpers shared by the RQSpline kernels (hb_flow_simt.cu: CUDA-core paths, hb_rqspline_tc.cu: tcgen05 path):
// accurate softplus, the 7-instruction tanh, and distrax's rational-quadratic spline forward with its parameters in
// registers (harmonic/flows.py:315-332 -> distrax.RationalQuadraticSpline).
#pragma once
#include <math_constants.h>

namespace hb {

constexpr int RQF_NP = 28;  // 3*8+1 = 25 spline parameters padded to 28

__device__ __forceinline__ float softplus_acc(float a) {
  return fmaxf(a, 0.f) + log1pf(expf(-fabsf(a)));
}

__device__ __forceinline__ float tanh_fast(float x) {
  // tanh|x| = (1 - e) / (1 + e) with e = exp(-2|x|) in (0, 1] (MUFU ex2 / rcp): seven instructions.  The ABSOLUTE
  // error is <= ~1.5e-7 everywhere (e carries 2^-22 relative, i.e. <= 2.4e-7 absolute, halved by the quotient's
  // slope at 0); the relative error grows towards x = 0 (cancellation in 1 - e), which does not matter here: every
  // tanh output only enters FMAs against the next Dense's weights and bias, where its absolute error counts like the
  // rounding of any other O(1) float32 term (round 1 carried an odd series for |x| < 0.25 beside this: twice the
  // instructions for no difference in ln phi at the 1e-7 level, profiles/README.md).
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-2.8853900817779268f * fabsf(x)));
  return copysignf(__fdividef(1.0f - e, 1.0f + e), x);
}

// distrax RationalQuadraticSpline forward, K = 8, parameters in registers (static indexing only)
__device__ __forceinline__ void rqs_forward_reg(float x, const float (&p)[RQF_NP], float lo, float hi,
                                                float& yout, float& ldout) {
  constexpr int K = 8;
  const float mbs = 1e-4f, mks = 1e-4f;
  const float off = 0.5411666523385311f;  // log(exp(1 - 1e-4) - 1)
  const float span = (hi - lo) - (float)K * mbs;
  float mw = p[0], mh = p[K];
#pragma unroll
  for (int k = 1; k < K; ++k) {
    mw = fmaxf(mw, p[k]);
    mh = fmaxf(mh, p[K + k]);
  }
  float ew[K], eh[K], sw = 0.f, shh = 0.f;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    ew[k] = __expf(p[k] - mw);
    eh[k] = __expf(p[K + k] - mh);
    sw += ew[k];
    shh += eh[k];
  }
  const float iw = span / sw, ih = span / shh;
  float xl = lo, yl = lo, cw = 0.f, chh = 0.f;
  float bxl = lo, bxr = lo, byl = lo, byr = lo, sl_raw = p[2 * K], sr_raw = p[2 * K + 1];
  bool found = false;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    cw += fmaf(ew[k], iw, mbs);
    chh += fmaf(eh[k], ih, mbs);
    const float xr = (k == K - 1) ? hi : lo + cw;
    const float yr = (k == K - 1) ? hi : lo + chh;
    const bool in = (!found) && (x >= xl) && (x < xr);
    if (k == 0 || in) {  // bin 0 is the default (e.g. NaN input)
      bxl = xl; bxr = xr; byl = yl; byr = yr;
      sl_raw = p[2 * K + k];
      sr_raw = p[2 * K + k + 1];
    }
    found = found || in;
    xl = xr;
    yl = yr;
  }
  const bool below = x <= lo, above = x >= hi;
  if (below) { sl_raw = p[2 * K]; }
  if (above) { sl_raw = p[3 * K]; }
  const float sl = softplus_acc(sl_raw + off) + mks;
  const float sr = softplus_acc(sr_raw + off) + mks;
  const float bw = bxr - bxl, bh = byr - byl;
  const float delta = bh / bw;
  float z = (x - bxl) / bw;
  z = fminf(fmaxf(z, 0.f), 1.f);
  if (x != x) z = x;  // jnp.clip propagates NaN
  const float sq_z = z * z;
  const float z1mz = z - sq_z;
  const float omz = 1.f - z;
  const float slopes_term = sr + sl - 2.f * delta;
  const float numerator = bh * (delta * sq_z + sl * z1mz);
  const float denominator = delta + slopes_term * z1mz;
  float yv = byl + numerator / denominator;
  float ldv = 2.f * __logf(delta) + __logf(sr * sq_z + 2.f * delta * z1mz + sl * omz * omz) -
              2.f * __logf(denominator);
  if (below) { yv = (x - lo) * sl + lo; ldv = __logf(sl); }
  if (above) { yv = (x - hi) * sl + hi; ldv = __logf(sl); }
  yout = yv;
  ldout = ldv;
}

// ---- the same spline with MUFU-level arithmetic (rqspline_tc_kernel: once the h1 x 25 contraction runs on the tensor
// cores the spline is 40 % of the kernel's instructions).  Differences from rqs_forward_reg, each at the 1e-7 level:
//   * exponentials as ex2(fma(p, log2 e, -max log2 e)), the two softmax normalisations, the bin's 1 / width and the
//     spline's 1 / denominator by rcp.approx (2^-23 relative) instead of IEEE divisions;
//   * softplus(a) = max(a, 0) + log1p(exp(-|a|)) with log1p(t) = 2 atanh(t / (2 + t)) as a degree-4 polynomial in the
//     square (relative error 8e-9, hb_realnvp_tc.cu) instead of log1pf(expf());
//   * the log-det as ONE logarithm, ln(delta^2 q / den^2) (delta^2 in [2.5e-11, 4e10], q / den^2 of order one: inside
//     float32's range for any admissible bin), instead of three.
__device__ __forceinline__ float rqs_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rqs_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float softplus_mufu(float a) {
  const float t = rqs_ex2(-1.4426950408889634f * fabsf(a));
  const float u = t * rqs_rcp(2.0f + t);
  const float z = u * u;
  const float pl = fmaf(z, fmaf(z, fmaf(z, fmaf(z, 0.14087678f, 0.13980642f), 0.20012431f), 0.33333158f), 1.0f);
  return fmaf(u + u, pl, fmaxf(a, 0.f));  // NaN in -> NaN out (t, u are NaN)
}
__device__ __forceinline__ void rqs_forward_mufu(float x, const float (&p)[RQF_NP], float lo, float hi,
                                                 float& yout, float& ldout) {
  constexpr int K = 8;
  const float mbs = 1e-4f, mks = 1e-4f;
  const float off = 0.5411666523385311f;  // log(exp(1 - 1e-4) - 1)
  const float L2E = 1.4426950408889634f;
  const float span = (hi - lo) - (float)K * mbs;
  float mw = p[0], mh = p[K];
#pragma unroll
  for (int k = 1; k < K; ++k) {
    mw = fmaxf(mw, p[k]);
    mh = fmaxf(mh, p[K + k]);
  }
  const float cw0 = -mw * L2E, ch0 = -mh * L2E;
  float ew[K], eh[K], sw = 0.f, shh = 0.f;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    ew[k] = rqs_ex2(fmaf(p[k], L2E, cw0));
    eh[k] = rqs_ex2(fmaf(p[K + k], L2E, ch0));
    sw += ew[k];
    shh += eh[k];
  }
  const float iw = span * rqs_rcp(sw), ih = span * rqs_rcp(shh);
  float xl = lo, yl = lo, cw = 0.f, chh = 0.f;
  float bxl = lo, bxr = lo, byl = lo, byr = lo, sl_raw = p[2 * K], sr_raw = p[2 * K + 1];
  bool found = false;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    cw += fmaf(ew[k], iw, mbs);
    chh += fmaf(eh[k], ih, mbs);
    const float xr = (k == K - 1) ? hi : lo + cw;
    const float yr = (k == K - 1) ? hi : lo + chh;
    const bool in = (!found) && (x >= xl) && (x < xr);
    if (k == 0 || in) {  // bin 0 is the default (e.g. NaN input)
      bxl = xl; bxr = xr; byl = yl; byr = yr;
      sl_raw = p[2 * K + k];
      sr_raw = p[2 * K + k + 1];
    }
    found = found || in;
    xl = xr;
    yl = yr;
  }
  const bool below = x <= lo, above = x >= hi;
  if (below) { sl_raw = p[2 * K]; }
  if (above) { sl_raw = p[3 * K]; }
  const float sl = softplus_mufu(sl_raw + off) + mks;
  const float sr = softplus_mufu(sr_raw + off) + mks;
  const float bw = bxr - bxl, bh = byr - byl;
  const float rbw = rqs_rcp(bw);
  const float delta = bh * rbw;
  float z = (x - bxl) * rbw;
  z = fminf(fmaxf(z, 0.f), 1.f);
  if (x != x) z = x;  // jnp.clip propagates NaN
  const float sq_z = z * z;
  const float z1mz = z - sq_z;
  const float omz = 1.f - z;
  const float slopes_term = sr + sl - 2.f * delta;
  const float numerator = bh * (delta * sq_z + sl * z1mz);
  const float denominator = delta + slopes_term * z1mz;
  const float rden = rqs_rcp(denominator);
  float yv = fmaf(numerator, rden, byl);
  const float q = sr * sq_z + 2.f * delta * z1mz + sl * omz * omz;
  const float dr = delta * rden;
  float ldv = __logf(dr * dr * q);
  if (below) { yv = (x - lo) * sl + lo; ldv = __logf(sl); }
  if (above) { yv = (x - hi) * sl + hi; ldv = __logf(sl); }
  yout = yv;
  ldout = ldv;
}

}  // namespace hb
