// CUDA-core FP32 flow kernels (one sample per thread): the general-shape path for RealNVP and the
// RQSpline path, both fused with the per-chain fold.  Also the on-device cross-check for the
// tcgen05 RealNVP kernel.
//
// Reference arithmetic replaced: harmonic/flows.py:41-116,159-180 (RealNVP via tfp-jax) and
// :234-288,335-438 (RQSpline via distrax); harmonic/model.py:222-235 (standardisation);
// fused with harmonic/evidence.py:278-352.  Per-sample restatement: SURVEY.md Appendix A.
#include <math.h>
#include <string.h>

#include <vector>

#include "hb_internal.h"

#include "hb_rqs.cuh"

namespace hb {

constexpr int FT = 128;  // threads per CTA = rows per tile
constexpr int MAX_LAYERS = 64;

struct SimtLayer {
  int w0, b0, w2, b2;  // offsets (floats) into the packed buffer
  int n2, n2pad;       // outputs of the second dense (u or 2u) and padded width
  int scaled;
};
struct RealNVPSimtParams {
  const float* pack;
  const float* pre;  // pre_offset[D], pre_amp[D] or nullptr
  int D, d, u, nlayers;
  float T, sum_log_amp;
  SimtLayer layers[MAX_LAYERS];
};


// out[j] = act(b[j] + sum_i in[phys(i)] * W[i][j]); activations live in shared memory as
// [feature][thread] columns; weights are read through the read-only path (warp-uniform
// addresses -> one broadcast transaction per float4).
template <int ACT, class Phys, class Use>
__device__ __forceinline__ void dense_cols(const float* in_s, Phys phys, Use use, int n_in,
                                           const float* __restrict__ W, int ldw,
                                           const float* __restrict__ b, int n_out, float* out_s,
                                           int tid) {
  for (int j0 = 0; j0 < n_out; j0 += 8) {
    float acc[8];
    {
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(b + j0));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(b + j0 + 4));
      acc[0] = b0.x; acc[1] = b0.y; acc[2] = b0.z; acc[3] = b0.w;
      acc[4] = b1.x; acc[5] = b1.y; acc[6] = b1.z; acc[7] = b1.w;
    }
    for (int i = 0; i < n_in; ++i) {
      if (!use(i)) continue;
      const float a = in_s[phys(i) * FT + tid];
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(W + (size_t)i * ldw + j0));
      const float4 w1 = __ldg(reinterpret_cast<const float4*>(W + (size_t)i * ldw + j0 + 4));
      acc[0] = fmaf(a, w0.x, acc[0]); acc[1] = fmaf(a, w0.y, acc[1]);
      acc[2] = fmaf(a, w0.z, acc[2]); acc[3] = fmaf(a, w0.w, acc[3]);
      acc[4] = fmaf(a, w1.x, acc[4]); acc[5] = fmaf(a, w1.y, acc[5]);
      acc[6] = fmaf(a, w1.z, acc[6]); acc[7] = fmaf(a, w1.w, acc[7]);
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      if (j0 + q < n_out) {
        float v = acc[q];
        if (ACT == 1) v = tanhf(v);
        if (ACT == 2) v = (v >= 0.f) ? v : 0.01f * v;
        out_s[(j0 + q) * FT + tid] = v;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// RealNVP, any ndim >= 2.  Shared memory: y[D][128] | h[128][128] | o[n2max][128].
// ------------------------------------------------------------------------------------------
template <typename XT, typename LT, bool FOLD>
__global__ void __launch_bounds__(FT)
realnvp_simt_kernel(RealNVPSimtParams P, const XT* __restrict__ x, long long ld,
                    const LT* __restrict__ lnpost, float* __restrict__ lnpred_out, FoldParams fp,
                    long long nrows, long long rows_per_cta) {
  extern __shared__ float smem[];
  __shared__ double scratch[3 * (FT / 32)];
  const int tid = threadIdx.x;
  const int D = P.D, d = P.d, u = P.u;
  float* ys = smem;
  float* hs = ys + (size_t)D * FT;
  float* os = hs + (size_t)128 * FT;
  const long long row_lo = (long long)blockIdx.x * rows_per_cta;
  long long row_hi = row_lo + rows_per_cta;
  if (row_hi > nrows) row_hi = nrows;
  const bool had_rows = row_lo < nrows;
  Fold<FT> fold(fp, blockIdx.x, tid, scratch, (FOLD && had_rows) ? row_lo : 0);
  const float invT = 1.f / P.T;
  const float base_const = -(float)D * logf(P.T) - 0.5f * (float)D * 1.8378770664093453f;

  for (long long t0 = row_lo; t0 < row_hi; t0 += FT) {
    const long long r = t0 + tid;
    const bool live = r < row_hi;
    const long long rr = live ? r : t0;  // dead threads recompute row t0 (never stored/folded)
    for (int i = 0; i < D; ++i) {
      float v = (float)x[rr * ld + i];
      if (P.pre) v = (v - P.pre[i]) / P.pre[D + i];
      ys[i * FT + tid] = v;
    }
    float ldj = 0.f;
    for (int l = 0; l < P.nlayers; ++l) {
      const SimtLayer L = P.layers[l];
      const int rot = l % D;  // logical y[i] lives in physical row (i + l) mod D
      auto phys = [&](int i) { int k = i + rot; return k >= D ? k - D : k; };
      auto all = [](int) { return true; };
      dense_cols<2>(ys, phys, all, d, P.pack + L.w0, 128, P.pack + L.b0, 128, hs, tid);
      dense_cols<0>(hs, [](int i) { return i; }, all, 128, P.pack + L.w2, L.n2pad, P.pack + L.b2,
                    L.n2, os, tid);
      for (int k = 0; k < u; ++k) {
        const int pk = phys(d + k);
        float y1 = ys[pk * FT + tid] - os[k * FT + tid];
        if (L.scaled) {
          float sc = softplus_acc(os[(u + k) * FT + tid]);
          sc = fminf(fmaxf(sc, 1e-3f), 1e3f);
          y1 = y1 / sc;
          ldj -= logf(sc);
        }
        ys[pk * FT + tid] = y1;
      }
    }
    float ss = 0.f;
    for (int i = 0; i < D; ++i) {
      const float z = ys[i * FT + tid] * invT;
      ss = fmaf(z, z, ss);
    }
    const float lnphi = -0.5f * ss + base_const + ldj - P.sum_log_amp;
    if (live && lnpred_out) lnpred_out[r] = lnphi;
    if (FOLD) {
      float a[1] = {lnphi};
      float b[1] = {live ? (float)lnpost[r] : 0.f};
      const long long t1 = (t0 + FT < row_hi) ? t0 + FT : row_hi;
      fold.tile<1>(t0, t1, a, b);
    }
  }
  if (FOLD) fold.finish(had_rows);
}

// ------------------------------------------------------------------------------------------
// RealNVP, small ndim (2..8): everything in registers, SM_R samples per thread (4 up to ndim 5, else 2),
// all layers' weights resident in shared memory as per-hidden-unit records
//   scaled layer    rec[j] = { b0[j], W0[0..d-1][j], c Wt[j][0..u-1], c Ws[j][0..u-1] }   (padded to 4 floats)
//   unscaled layer  rec[j] = { b0[j], W0[0..d-1][j], c Wt[j][0..u-1] }
// so one warp-uniform LDS.128 stream feeds d + 1 + 2u (resp. d + 1 + u) arithmetic instructions per sample and
// hidden unit; the hidden activation is never stored.  leaky_relu(h) = s h + c relu(h) with s = 0.01, c = 1 - s
// (harmonic/flows.py:170): the linear part of a coupling MLP, s (y W0 + b0) W2 + b2 = y V + v0, is a d x 2u matrix
// formed once on the host in float64 (the tail of each layer's block) and seeds the output accumulators, the loop over
// the hidden units keeps only relu -- one FMNMX instead of FMUL + FMNMX per sample and hidden unit.
// ------------------------------------------------------------------------------------------
constexpr int SM_THREADS = 256;  // (128 -> 256: 32 instead of 16 resident warps per SM at cfg3, 5.04 -> 4.61 ms)
constexpr int small_r(int D) { return D <= 5 ? 4 : 2; }  // samples per thread
constexpr int small_d(int D) { return (D == 2) ? 1 : (D == 3) ? 2 : (D == 4) ? 2 : (D == 5) ? 2 : (D == 6) ? 3 : 4; }

struct SmallParams {
  const float* pack;   // scaled layers first: 128 records + tail each; tail = (1 + d) rows of 2 x (4-padded u) floats
  const float* pre;    // pre_offset[D], pre_amp[D] or nullptr
  int nlayers, nscaled, total_floats;
  int lf_scaled, lf_unscaled;  // floats per layer block
  float T, sum_log_amp;
  int x_f64, lp_f64;
};

// One coupling layer for the R samples of a thread.  L: the layer's block in shared memory.
template <int D, int R, bool SCALED>
__device__ __forceinline__ void small_layer(const float* __restrict__ L, float (&y)[R][D], float (&ldj)[R]) {
  constexpr int d = small_d(D), u = D - d, UP = (u + 3) & ~3;
  constexpr int NW = 1 + d + (SCALED ? 2 * u : u);
  constexpr int RP = (NW + 3) & ~3;
  const float* tail = L + 128 * RP;
  float sh[R][u], sc[R][SCALED ? u : 1];
#pragma unroll
  for (int k = 0; k < u; ++k) {
    float vt[1 + d], vs[1 + d];
#pragma unroll
    for (int i = 0; i <= d; ++i) {
      vt[i] = tail[i * 2 * UP + k];
      vs[i] = SCALED ? tail[i * 2 * UP + UP + k] : 0.f;
    }
#pragma unroll
    for (int q = 0; q < R; ++q) {
      float t = vt[0], s = vs[0];
#pragma unroll
      for (int i = 0; i < d; ++i) {
        t = fmaf(y[q][i], vt[1 + i], t);
        if (SCALED) s = fmaf(y[q][i], vs[1 + i], s);
      }
      sh[q][k] = t;
      if (SCALED) sc[q][k] = s;
    }
  }
#pragma unroll 4
  for (int j = 0; j < 128; ++j) {
    const float* rec = L + j * RP;
    float w[RP];
#pragma unroll
    for (int i = 0; i < RP; i += 4) {
      const float4 t4 = *reinterpret_cast<const float4*>(rec + i);
      w[i] = t4.x;
      w[i + 1] = t4.y;
      w[i + 2] = t4.z;
      w[i + 3] = t4.w;
    }
#pragma unroll
    for (int q = 0; q < R; ++q) {
      float h = w[0];
#pragma unroll
      for (int i = 0; i < d; ++i) h = fmaf(y[q][i], w[1 + i], h);
      h = fmaxf(h, 0.f);
#pragma unroll
      for (int k = 0; k < u; ++k) {
        sh[q][k] = fmaf(h, w[1 + d + k], sh[q][k]);
        if (SCALED) sc[q][k] = fmaf(h, w[1 + d + u + k], sc[q][k]);
      }
    }
  }
#pragma unroll
  for (int q = 0; q < R; ++q) {
#pragma unroll
    for (int k = 0; k < u; ++k) {
      float y1 = y[q][d + k] - sh[q][k];
      if (SCALED) {
        float s = softplus_acc(sc[q][k]);
        s = fminf(fmaxf(s, 1e-3f), 1e3f);
        s = (sc[q][k] != sc[q][k]) ? sc[q][k] : s;
        y1 = y1 / s;
        ldj[q] -= logf(s);
      }
      y[q][d + k] = y1;
    }
  }
}

#ifndef HB_SMALL_MINB
#define HB_SMALL_MINB 1
#endif
template <int D, bool FOLD>
__global__ void __launch_bounds__(SM_THREADS, HB_SMALL_MINB)
realnvp_small_kernel(SmallParams P, const void* __restrict__ xv, long long ld,
                     const void* __restrict__ lnpostv, float* __restrict__ lnpred_out, FoldParams fp,
                     long long nrows, long long rows_per_cta) {
  constexpr int SM_R = small_r(D);
  extern __shared__ float wsm[];
  __shared__ double scratch[3 * (SM_THREADS / 32)];
  const int tid = threadIdx.x;
  for (int i = tid; i < P.total_floats; i += SM_THREADS) wsm[i] = P.pack[i];
  __syncthreads();
  const long long row_lo = (long long)blockIdx.x * rows_per_cta;
  long long row_hi = row_lo + rows_per_cta;
  if (row_hi > nrows) row_hi = nrows;
  const bool had_rows = row_lo < nrows;
  Fold<SM_THREADS> fold(fp, blockIdx.x, tid, scratch, (FOLD && had_rows) ? row_lo : 0);
  const float invT = 1.f / P.T;
  const float base_const = -(float)D * logf(P.T) - 0.5f * (float)D * 1.8378770664093453f;

  for (long long t0 = row_lo; t0 < row_hi; t0 += SM_THREADS * SM_R) {
    float y[SM_R][D];
    float ldj[SM_R];
    long long rq[SM_R];
#pragma unroll
    for (int q = 0; q < SM_R; ++q) {
      const long long r = t0 + (long long)tid * SM_R + q;
      rq[q] = r;
      const long long rr = (r < row_hi) ? r : row_lo;
#pragma unroll
      for (int i = 0; i < D; ++i) {
        float v = P.x_f64 ? (float)static_cast<const double*>(xv)[rr * ld + i]
                          : static_cast<const float*>(xv)[rr * ld + i];
        if (P.pre) v = (v - P.pre[i]) / P.pre[D + i];
        y[q][i] = v;
      }
      ldj[q] = 0.f;
    }
    // relu swallows a NaN hidden unit (fmaxf(NaN, 0) = 0) where leaky_relu propagates it; the linear term y V + v0
    // keeps the propagation: it multiplies EVERY conditioning input (NaN * 0 = NaN, inf * v = +-inf or NaN), so a
    // non-finite input still ends in a non-finite ln phi (tests/test_gpu_flows.py::test_nan_propagates).
    for (int l = 0; l < P.nlayers; ++l) {
      if (l < P.nscaled)
        small_layer<D, SM_R, true>(wsm + (size_t)l * P.lf_scaled, y, ldj);
      else
        small_layer<D, SM_R, false>(wsm + (size_t)P.nscaled * P.lf_scaled + (size_t)(l - P.nscaled) * P.lf_unscaled,
                                    y, ldj);
      if (l + 1 < P.nlayers) {  // inverse permute: rotate left by one
#pragma unroll
        for (int q = 0; q < SM_R; ++q) {
          const float y0 = y[q][0];
#pragma unroll
          for (int i = 0; i < D - 1; ++i) y[q][i] = y[q][i + 1];
          y[q][D - 1] = y0;
        }
      }
    }
    float a[SM_R], b[SM_R];
#pragma unroll
    for (int q = 0; q < SM_R; ++q) {
      float ss = 0.f;
#pragma unroll
      for (int i = 0; i < D; ++i) {
        const float z = y[q][i] * invT;
        ss = fmaf(z, z, ss);
      }
      a[q] = -0.5f * ss + base_const + ldj[q] - P.sum_log_amp;
      const bool live = rq[q] < row_hi;
      if (live && lnpred_out) lnpred_out[rq[q]] = a[q];
      b[q] = 0.f;
      if (FOLD && live)
        b[q] = P.lp_f64 ? (float)static_cast<const double*>(lnpostv)[rq[q]]
                        : static_cast<const float*>(lnpostv)[rq[q]];
    }
    if (FOLD) {
      long long t1 = t0 + SM_THREADS * SM_R;
      if (t1 > row_hi) t1 = row_hi;
      fold.tile<SM_R>(t0, t1, a, b);
    }
  }
  if (FOLD) fold.finish(had_rows);
}

// ------------------------------------------------------------------------------------------
// RQSpline, any ndim / hidden sizes.  Packed weights per flow layer i:
//   MLP dense m = 0..nmlp-2 (each followed by tanh): W[in][outpad], b[outpad]
//   per transformed feature j: fused (last MLP dense) x (output dense rows of j):
//       M_j[in_last][npar_pad], c_j[npar_pad]          (no activation in between, flows.py:434-438)
//   scales[D], shifts[D]
// ------------------------------------------------------------------------------------------
constexpr int RQ_MAX_MLP = HB_MAX_HIDDEN + 1;
struct RQLayer {
  int w[RQ_MAX_MLP], b[RQ_MAX_MLP];  // hidden denses (with tanh)
  int fused;                         // offset of the first feature block; blocks are contiguous
  int scales, shifts;
};
struct RQParams {
  const float* pack;
  const float* pre;
  int D, K, npar, npar_pad, nlayers;
  int nact;                  // number of tanh denses = nmlp - 1
  int width[RQ_MAX_MLP + 1];  // width[0] = D (input), width[m+1] = out of dense m
  int wpad[RQ_MAX_MLP + 1];
  int in_last;                // input width of the fused dense
  int maxw;                   // max activation width
  float lo, hi, T, sum_log_amp;
  const RQLayer* layers;  // device
};

// distrax.RationalQuadraticSpline forward + log-det for one scalar; params in a smem column.
__device__ __forceinline__ void rqs_forward(float x, const float* ps, int tid, int K, float lo,
                                            float hi, float& yout, float& ldout) {
  const float mbs = 1e-4f, mks = 1e-4f;
  const float off = logf(expf(1.f - mks) - 1.f);
  const float R = hi - lo;
  const float span = R - (float)K * mbs;
  float mw = -CUDART_INF_F, mh = -CUDART_INF_F;
  for (int k = 0; k < K; ++k) {
    mw = fmaxf(mw, ps[k * FT + tid]);
    mh = fmaxf(mh, ps[(K + k) * FT + tid]);
  }
  float sw = 0.f, shh = 0.f;
  for (int k = 0; k < K; ++k) {
    sw += expf(ps[k * FT + tid] - mw);
    shh += expf(ps[(K + k) * FT + tid] - mh);
  }
  if (x <= lo) {
    const float s0 = softplus_acc(ps[(2 * K) * FT + tid] + off) + mks;
    yout = (x - lo) * s0 + lo;
    ldout = logf(s0);
    return;
  }
  if (x >= hi) {
    const float sK = softplus_acc(ps[(3 * K) * FT + tid] + off) + mks;
    yout = (x - hi) * sK + hi;
    ldout = logf(sK);
    return;
  }
  // knots: xk[0]=lo, xk[k]=lo+cumsum(w)[k-1], xk[K]=hi (same for yk)
  float xl = lo, yl = lo, cw = 0.f, chh = 0.f;
  float bxl = lo, bxr = lo, byl = lo, byr = lo;
  int kb = -1;
  for (int k = 0; k < K; ++k) {
    const float w = expf(ps[k * FT + tid] - mw) / sw * span + mbs;
    const float h = expf(ps[(K + k) * FT + tid] - mh) / shh * span + mbs;
    cw += w;
    chh += h;
    const float xr = (k == K - 1) ? hi : lo + cw;
    const float yr = (k == K - 1) ? hi : lo + chh;
    if (k == 0) { bxl = xl; bxr = xr; byl = yl; byr = yr; }  // default bin 0 (e.g. NaN input)
    if (kb < 0 && x >= xl && x < xr) {
      kb = k; bxl = xl; bxr = xr; byl = yl; byr = yr;
    }
    xl = xr;
    yl = yr;
  }
  if (kb < 0) kb = 0;
  const float sl = softplus_acc(ps[(2 * K + kb) * FT + tid] + off) + mks;
  const float sr = softplus_acc(ps[(2 * K + kb + 1) * FT + tid] + off) + mks;
  const float bw = bxr - bxl, bh = byr - byl;
  const float delta = bh / bw;
  float z = (x - bxl) / bw;
  z = fminf(fmaxf(z, 0.f), 1.f);
  if (isnan(x)) z = x;  // jnp.clip propagates NaN; fminf/fmaxf would drop it
  const float sq_z = z * z;
  const float z1mz = z - sq_z;
  const float omz = 1.f - z;
  const float sq_1mz = omz * omz;
  const float slopes_term = sr + sl - 2.f * delta;
  const float numerator = bh * (delta * sq_z + sl * z1mz);
  const float denominator = delta + slopes_term * z1mz;
  yout = byl + numerator / denominator;
  ldout = 2.f * logf(delta) + logf(sr * sq_z + 2.f * delta * z1mz + sl * sq_1mz) -
          2.f * logf(denominator);
}

// distrax.RationalQuadraticSpline inverse (quadratic root) for one scalar; params in a smem column.
// Used by the sampling direction only (harmonic/flows.py:290-313).
__device__ __forceinline__ float rqs_inverse(float y, const float* ps, int tid, int K, float lo, float hi) {
  const float mbs = 1e-4f, mks = 1e-4f;
  const float off = logf(expf(1.f - mks) - 1.f);
  const float R = hi - lo;
  const float span = R - (float)K * mbs;
  if (y <= lo) {
    const float s0 = softplus_acc(ps[(2 * K) * FT + tid] + off) + mks;
    return (y - lo) / s0 + lo;
  }
  if (y >= hi) {
    const float sK = softplus_acc(ps[(3 * K) * FT + tid] + off) + mks;
    return (y - hi) / sK + hi;
  }
  float mw = -CUDART_INF_F, mh = -CUDART_INF_F;
  for (int k = 0; k < K; ++k) {
    mw = fmaxf(mw, ps[k * FT + tid]);
    mh = fmaxf(mh, ps[(K + k) * FT + tid]);
  }
  float sw = 0.f, shh = 0.f;
  for (int k = 0; k < K; ++k) {
    sw += expf(ps[k * FT + tid] - mw);
    shh += expf(ps[(K + k) * FT + tid] - mh);
  }
  float xl = lo, yl = lo, cw = 0.f, chh = 0.f;
  float bxl = lo, bxr = lo, byl = lo, byr = lo;
  int kb = -1;
  for (int k = 0; k < K; ++k) {
    const float w = expf(ps[k * FT + tid] - mw) / sw * span + mbs;
    const float h = expf(ps[(K + k) * FT + tid] - mh) / shh * span + mbs;
    cw += w;
    chh += h;
    const float xr = (k == K - 1) ? hi : lo + cw;
    const float yr = (k == K - 1) ? hi : lo + chh;
    if (k == 0) { bxl = xl; bxr = xr; byl = yl; byr = yr; }  // default bin 0 (e.g. NaN input)
    if (kb < 0 && y >= yl && y < yr) {
      kb = k; bxl = xl; bxr = xr; byl = yl; byr = yr;
    }
    xl = xr;
    yl = yr;
  }
  if (kb < 0) kb = 0;
  const float sl = softplus_acc(ps[(2 * K + kb) * FT + tid] + off) + mks;
  const float sr = softplus_acc(ps[(2 * K + kb + 1) * FT + tid] + off) + mks;
  const float bw = bxr - bxl, bh = byr - byl;
  const float delta = bh / bw;
  float w = (y - byl) / bh;
  w = fminf(fmaxf(w, 0.f), 1.f);
  if (isnan(y)) w = y;
  const float slopes_term = sr + sl - 2.f * delta;
  const float c = -delta * w;
  const float b = sl - slopes_term * w;
  const float a = delta - b;
  float z = -2.f * c / (b + sqrtf(b * b - 4.f * a * c));
  z = fminf(fmaxf(z, 0.f), 1.f);
  if (isnan(y)) z = y;
  return bw * z + bxl;
}

template <typename XT, typename LT, bool FOLD>
__global__ void __launch_bounds__(FT)
rqspline_simt_kernel(RQParams P, const XT* __restrict__ x, long long ld,
                     const LT* __restrict__ lnpost, float* __restrict__ lnpred_out, FoldParams fp,
                     long long nrows, long long rows_per_cta) {
  extern __shared__ float smem[];
  __shared__ double scratch[3 * (FT / 32)];
  const int tid = threadIdx.x;
  const int D = P.D, K = P.K;
  float* ys = smem;                              // [D][FT]
  float* bufA = ys + (size_t)D * FT;             // [maxw][FT]
  float* bufB = bufA + (size_t)P.maxw * FT;      // [maxw][FT]
  float* ps = bufB + (size_t)P.maxw * FT;        // [npar_pad][FT]
  const long long row_lo = (long long)blockIdx.x * rows_per_cta;
  long long row_hi = row_lo + rows_per_cta;
  if (row_hi > nrows) row_hi = nrows;
  const bool had_rows = row_lo < nrows;
  Fold<FT> fold(fp, blockIdx.x, tid, scratch, (FOLD && had_rows) ? row_lo : 0);
  const float base_const = -0.5f * (float)D * 1.8378770664093453f - 0.5f * (float)D * logf(P.T);
  auto ident = [](int i) { return i; };
  auto all = [](int) { return true; };

  for (long long t0 = row_lo; t0 < row_hi; t0 += FT) {
    const long long r = t0 + tid;
    const bool live = r < row_hi;
    const long long rr = live ? r : t0;
    for (int i = 0; i < D; ++i) {
      float v = (float)x[rr * ld + i];
      if (P.pre) v = (v - P.pre[i]) / P.pre[D + i];
      ys[i * FT + tid] = v;
    }
    float ldet = 0.f;
    for (int i = P.nlayers - 1; i >= 0; --i) {  // Inverse(Chain).inverse: last layer first
      const RQLayer L = P.layers[i];
      const int par = i & 1;  // mask = (j%2==1) for even i, negated for odd i; True = pass-through
      auto visible = [&](int j) { return ((j & 1) == 1) != (par == 1); };
      // conditioner MLP on where(mask, y, 0)
      const float* cur = ys;
      float* nxt = bufA;
      for (int m = 0; m < P.nact; ++m) {
        if (m == 0)
          dense_cols<1>(cur, ident, visible, P.width[0], P.pack + L.w[0], P.wpad[1], P.pack + L.b[0],
                        P.width[1], nxt, tid);
        else
          dense_cols<1>(cur, ident, all, P.width[m], P.pack + L.w[m], P.wpad[m + 1], P.pack + L.b[m],
                        P.width[m + 1], nxt, tid);
        cur = nxt;
        nxt = (nxt == bufA) ? bufB : bufA;
      }
      // spline on every non-visible feature, params from the fused last dense
      int blk = 0;
      const size_t blk_floats = (size_t)P.in_last * P.npar_pad + P.npar_pad;
      for (int j = 0; j < D; ++j) {
        if (visible(j)) continue;
        const float* Mj = P.pack + L.fused + blk * blk_floats;
        const float* cj = Mj + (size_t)P.in_last * P.npar_pad;
        if (P.nact == 0)
          dense_cols<0>(ys, ident, visible, P.in_last, Mj, P.npar_pad, cj, P.npar, ps, tid);
        else
          dense_cols<0>(cur, ident, all, P.in_last, Mj, P.npar_pad, cj, P.npar, ps, tid);
        float yo, lo_;
        rqs_forward(ys[j * FT + tid], ps, tid, K, P.lo, P.hi, yo, lo_);
        ys[j * FT + tid] = yo;
        ldet += lo_;
        ++blk;
      }
      // scalar affine, all dims
      for (int j = 0; j < D; ++j) {
        const float sc = __ldg(P.pack + L.scales + j), sh = __ldg(P.pack + L.shifts + j);
        ys[j * FT + tid] = fmaf(ys[j * FT + tid], sc, sh);
        ldet += logf(fabsf(sc));
      }
    }
    float ss = 0.f;
    for (int j = 0; j < D; ++j) {
      const float v = ys[j * FT + tid];
      ss = fmaf(v, v, ss);
    }
    const float lnphi = -0.5f * ss / P.T + base_const + ldet - P.sum_log_amp;
    if (live && lnpred_out) lnpred_out[r] = lnphi;
    if (FOLD) {
      float a[1] = {lnphi};
      float b[1] = {live ? (float)lnpost[r] : 0.f};
      const long long t1 = (t0 + FT < row_hi) ? t0 + FT : row_hi;
      fold.tile<1>(t0, t1, a, b);
    }
  }
  if (FOLD) fold.finish(had_rows);
}

// ------------------------------------------------------------------------------------------
// RQSpline fast path: n_hidden == 2, ndim 2..8, 8 bins.  Everything in registers, two samples per
// thread; weights are read through the read-only path with warp-uniform float4 loads.  Per flow
// layer:  a1 = tanh(Dense_0(where(mask, y, 0)))            (ndim values)
//         for every transformed feature f:  P_f = c_f + sum_j tanh(Dense_1(a1))_j * M_f[j][:]
//         (M_f = last MLP Dense x output Dense rows of f, pre-multiplied on the host; the hidden
//         activation is recomputed per feature instead of being stored)
//         spline forward on feature f, then the scalar affine on all features.
// Packed per layer: W1[D][D4] b1[D4] | W2T[h1][D4 + 4] (row j: W2[0..D-1][j], b2[j]) |
//                   per transformed feature: M[h1][28], c[28] | scales[D4] shifts[D4]
// ------------------------------------------------------------------------------------------
constexpr int RQF_THREADS = 128;
constexpr int RQF_R = 2;

struct RQFastParams {
  const float* pack;
  const float* pre;
  int nlayers, h1, layer_floats_even, layer_floats_odd;  // layers alternate parity -> sizes differ
  float lo, hi, T, sum_log_amp;
  int x_f64, lp_f64;
};

template <int D, bool FOLD>
__global__ void __launch_bounds__(RQF_THREADS)
rqspline_fast_kernel(RQFastParams P, const void* __restrict__ xv, long long ld,
                     const void* __restrict__ lnpostv, float* __restrict__ lnpred_out, FoldParams fp,
                     long long nrows, long long rows_per_cta) {
  constexpr int D4 = (D + 3) & ~3;
  __shared__ double scratch[3 * (RQF_THREADS / 32)];
  const int tid = threadIdx.x;
  const long long row_lo = (long long)blockIdx.x * rows_per_cta;
  long long row_hi = row_lo + rows_per_cta;
  if (row_hi > nrows) row_hi = nrows;
  const bool had_rows = row_lo < nrows;
  Fold<RQF_THREADS> fold(fp, blockIdx.x, tid, scratch, (FOLD && had_rows) ? row_lo : 0);
  const float base_const = -0.5f * (float)D * 1.8378770664093453f - 0.5f * (float)D * logf(P.T);
  const float invT = 1.f / P.T;
  const int h1 = P.h1;

  for (long long t0 = row_lo; t0 < row_hi; t0 += RQF_THREADS * RQF_R) {
    float y[RQF_R][D], ldet[RQF_R];
    long long rq[RQF_R];
#pragma unroll
    for (int q = 0; q < RQF_R; ++q) {
      const long long r = t0 + (long long)tid * RQF_R + q;
      rq[q] = r;
      const long long rr = (r < row_hi) ? r : row_lo;
#pragma unroll
      for (int i = 0; i < D; ++i) {
        float v = P.x_f64 ? (float)static_cast<const double*>(xv)[rr * ld + i]
                          : static_cast<const float*>(xv)[rr * ld + i];
        if (P.pre) v = (v - P.pre[i]) / P.pre[D + i];
        y[q][i] = v;
      }
      ldet[q] = 0.f;
    }
    for (int li = P.nlayers - 1; li >= 0; --li) {  // Inverse(Chain).inverse: last layer first
      const int par = li & 1;  // mask = (j % 2 == 1) for even layers, negated for odd; True = pass-through
      // offset of layer li: even layers first have size fe, odd fo; layers are stored in order
      const size_t loff = (size_t)((li + 1) >> 1) * P.layer_floats_even + (size_t)(li >> 1) * P.layer_floats_odd;
      const float* L = P.pack + loff;
      const float* W1 = L;                  // [D][D4]
      const float* b1 = W1 + D * D4;        // [D4]
      const float* W2T = b1 + D4;           // [h1][D4 + 4]
      const float* feat = W2T + (size_t)h1 * (D4 + 4);
      // ---- a1 = tanh(Dense_0(where(mask, y, 0)))
      float a1[RQF_R][D];
#pragma unroll
      for (int q = 0; q < RQF_R; ++q)
#pragma unroll
        for (int o = 0; o < D; ++o) a1[q][o] = __ldg(b1 + o);
#pragma unroll
      for (int i = 0; i < D; ++i) {
        const bool vis = ((i & 1) == 1) != (par == 1);
        if (!vis) continue;  // compile-time per (i, par) after unrolling both parities below
#pragma unroll
        for (int o = 0; o < D; ++o) {
          const float w = __ldg(W1 + i * D4 + o);
#pragma unroll
          for (int q = 0; q < RQF_R; ++q) a1[q][o] = fmaf(y[q][i], w, a1[q][o]);
        }
      }
#pragma unroll
      for (int q = 0; q < RQF_R; ++q)
#pragma unroll
        for (int o = 0; o < D; ++o) a1[q][o] = tanh_fast(a1[q][o]);
      // ---- every transformed feature
      int blk = 0;
      const size_t blk_floats = (size_t)h1 * RQF_NP + RQF_NP;
#pragma unroll
      for (int f = 0; f < D; ++f) {
        const bool vis = ((f & 1) == 1) != (par == 1);
        if (vis) continue;
        const float* M = feat + (size_t)blk * blk_floats;
        const float* cb = M + (size_t)h1 * RQF_NP;
        float out[RQF_R][RQF_NP];
#pragma unroll
        for (int p4 = 0; p4 < RQF_NP; p4 += 4) {
          const float4 c4 = __ldg(reinterpret_cast<const float4*>(cb + p4));
#pragma unroll
          for (int q = 0; q < RQF_R; ++q) {
            out[q][p4] = c4.x; out[q][p4 + 1] = c4.y; out[q][p4 + 2] = c4.z; out[q][p4 + 3] = c4.w;
          }
        }
#pragma unroll 2
        for (int j = 0; j < h1; ++j) {
          float w2[D4 + 4];
#pragma unroll
          for (int i4 = 0; i4 < D4 + 4; i4 += 4) {
            const float4 t4 = __ldg(reinterpret_cast<const float4*>(W2T + (size_t)j * (D4 + 4) + i4));
            w2[i4] = t4.x; w2[i4 + 1] = t4.y; w2[i4 + 2] = t4.z; w2[i4 + 3] = t4.w;
          }
          float a2[RQF_R];
#pragma unroll
          for (int q = 0; q < RQF_R; ++q) {
            float h = w2[D4];  // b2[j]
#pragma unroll
            for (int i = 0; i < D; ++i) h = fmaf(a1[q][i], w2[i], h);
            a2[q] = tanh_fast(h);
          }
#pragma unroll
          for (int p4 = 0; p4 < RQF_NP; p4 += 4) {
            const float4 m4 = __ldg(reinterpret_cast<const float4*>(M + (size_t)j * RQF_NP + p4));
#pragma unroll
            for (int q = 0; q < RQF_R; ++q) {
              out[q][p4] = fmaf(a2[q], m4.x, out[q][p4]);
              out[q][p4 + 1] = fmaf(a2[q], m4.y, out[q][p4 + 1]);
              out[q][p4 + 2] = fmaf(a2[q], m4.z, out[q][p4 + 2]);
              out[q][p4 + 3] = fmaf(a2[q], m4.w, out[q][p4 + 3]);
            }
          }
        }
#pragma unroll
        for (int q = 0; q < RQF_R; ++q) {
          float yo, lo_;
          rqs_forward_reg(y[q][f], out[q], P.lo, P.hi, yo, lo_);
          y[q][f] = yo;
          ldet[q] += lo_;
        }
        ++blk;
      }
      // ---- scalar affine, all dims
      const float* sc = feat + (size_t)blk * blk_floats;
      const float* sh = sc + D4;
#pragma unroll
      for (int i = 0; i < D; ++i) {
        const float s = __ldg(sc + i), t = __ldg(sh + i);
        const float ls = __logf(fabsf(s));
#pragma unroll
        for (int q = 0; q < RQF_R; ++q) {
          y[q][i] = fmaf(y[q][i], s, t);
          ldet[q] += ls;
        }
      }
    }
    float a[RQF_R], b[RQF_R];
#pragma unroll
    for (int q = 0; q < RQF_R; ++q) {
      float ss = 0.f;
#pragma unroll
      for (int i = 0; i < D; ++i) ss = fmaf(y[q][i], y[q][i], ss);
      a[q] = -0.5f * ss * invT + base_const + ldet[q] - P.sum_log_amp;
      const bool live = rq[q] < row_hi;
      if (live && lnpred_out) lnpred_out[rq[q]] = a[q];
      b[q] = 0.f;
      if (FOLD && live)
        b[q] = P.lp_f64 ? (float)static_cast<const double*>(lnpostv)[rq[q]]
                        : static_cast<const float*>(lnpostv)[rq[q]];
    }
    if (FOLD) {
      long long t1 = t0 + RQF_THREADS * RQF_R;
      if (t1 > row_hi) t1 = row_hi;
      fold.tile<RQF_R>(t0, t1, a, b);
    }
  }
  if (FOLD) fold.finish(had_rows);
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static inline int pad8(int n) { return (n + 7) & ~7; }

// Packed SIMT weights are built lazily and cached on the Flow (rq_pack is reused for RealNVP).
static int realnvp_simt_prepare(Flow* f, RealNVPSimtParams* P) {
  if (f->simt_params) {
    *P = *static_cast<RealNVPSimtParams*>(f->simt_params);
    return HB_OK;
  }
  const int D = f->desc.ndim, d = f->d, u = f->u;
  const int nl = f->nlayers;
  if (nl > MAX_LAYERS) {
    set_error("RealNVP: more than 64 coupling layers are not supported");
    return HB_ERR_UNSUPPORTED;
  }
  std::vector<float> pack;
  auto align4 = [&]() { while (pack.size() % 4) pack.push_back(0.f); };
  const float* w = f->host_weights.data();
  for (int l = 0; l < nl; ++l) {
    const bool scaled = l < f->desc.n_scaled;
    const float* base = w + f->layer_off[l];
    const float* W0 = base;
    const float* b0 = W0 + (size_t)d * 128;
    const float* Wt = b0 + 128;
    const float* bt = Wt + (size_t)128 * u;
    const float* Ws = bt + u;
    const float* bs = Ws + (size_t)128 * u;
    SimtLayer& L = P->layers[l];
    L.scaled = scaled;
    L.n2 = scaled ? 2 * u : u;
    L.n2pad = pad8(L.n2);
    align4();
    L.w0 = (int)pack.size();
    pack.insert(pack.end(), W0, W0 + (size_t)d * 128);
    L.b0 = (int)pack.size();
    pack.insert(pack.end(), b0, b0 + 128);
    L.w2 = (int)pack.size();
    for (int j = 0; j < 128; ++j)
      for (int k = 0; k < L.n2pad; ++k) {
        float v = 0.f;
        if (k < u) v = Wt[(size_t)j * u + k];
        else if (scaled && k < 2 * u) v = Ws[(size_t)j * u + (k - u)];
        pack.push_back(v);
      }
    L.b2 = (int)pack.size();
    for (int k = 0; k < L.n2pad; ++k) {
      float v = 0.f;
      if (k < u) v = bt[k];
      else if (scaled && k < 2 * u) v = bs[k - u];
      pack.push_back(v);
    }
  }
  {
    HB_CUDA(cudaMalloc(&f->rq_pack, pack.size() * sizeof(float)));
    HB_CUDA(cudaMemcpy(f->rq_pack, pack.data(), pack.size() * sizeof(float), cudaMemcpyHostToDevice));
    f->rq_pack_floats = pack.size();
  }
  P->pack = f->rq_pack;
  P->pre = f->desc.standardize ? f->dev_weights : nullptr;
  P->D = D;
  P->d = d;
  P->u = u;
  P->nlayers = nl;
  P->sum_log_amp = f->sum_log_amp;
  f->simt_params = new RealNVPSimtParams(*P);
  return HB_OK;
}

void realnvp_simt_release(Flow* f) {
  delete static_cast<RealNVPSimtParams*>(f->simt_params);
  f->simt_params = nullptr;
}

#define HB_DISPATCH_XL(KERNEL, FOLDFLAG, ...)                                                    \
  do {                                                                                           \
    if (x_dtype == HB_F32 && lp_dtype == HB_F32)                                                 \
      KERNEL<float, float, FOLDFLAG><<<grid, FT, smem, st>>>(P, (const float*)x, ld,             \
                                                             (const float*)lnpost, __VA_ARGS__); \
    else if (x_dtype == HB_F32)                                                                  \
      KERNEL<float, double, FOLDFLAG><<<grid, FT, smem, st>>>(P, (const float*)x, ld,            \
                                                              (const double*)lnpost, __VA_ARGS__); \
    else if (lp_dtype == HB_F32)                                                                 \
      KERNEL<double, float, FOLDFLAG><<<grid, FT, smem, st>>>(P, (const double*)x, ld,           \
                                                              (const float*)lnpost, __VA_ARGS__); \
    else                                                                                         \
      KERNEL<double, double, FOLDFLAG><<<grid, FT, smem, st>>>(P, (const double*)x, ld,          \
                                                               (const double*)lnpost, __VA_ARGS__); \
  } while (0)

#define HB_SET_SMEM(KERNEL, FOLDFLAG)                                                              \
  do {                                                                                             \
    cudaFuncSetAttribute(KERNEL<float, float, FOLDFLAG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   \
    cudaFuncSetAttribute(KERNEL<float, double, FOLDFLAG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);  \
    cudaFuncSetAttribute(KERNEL<double, float, FOLDFLAG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);  \
    cudaFuncSetAttribute(KERNEL<double, double, FOLDFLAG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
  } while (0)

static void plan_grid(int device, long long n, size_t smem, long long* nfl, long long* rpc) {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  int per_sm = (int)((227 * 1024) / (smem + 2048));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 8) per_sm = 8;
  long long f = (long long)sms * per_sm;
  long long r = (n + f - 1) / f;
  r = ((r + FT - 1) / FT) * FT;
  *nfl = (n + r - 1) / r;
  *rpc = r;
}

// ---- small-ndim register path ------------------------------------------------------------------
struct SmallCache {
  float* dev = nullptr;
  SmallParams P;
};

static int realnvp_small_prepare(Flow* f, SmallParams* P) {
  if (f->small_cache) {
    *P = static_cast<SmallCache*>(f->small_cache)->P;
    return HB_OK;
  }
  const int d = f->d, u = f->u, nl = f->nlayers, ns = f->desc.n_scaled;
  const int up = (u + 3) & ~3;
  const int rp_s = (1 + d + 2 * u + 3) & ~3, rp_u = (1 + d + u + 3) & ~3;
  const int tail = (1 + d) * 2 * up;
  const int lf_s = 128 * rp_s + tail, lf_u = 128 * rp_u + tail;
  const double slope = 0.01, c = 1.0 - slope;  // leaky_relu(h) = slope h + c relu(h)
  std::vector<float> pack((size_t)ns * lf_s + (size_t)(nl - ns) * lf_u, 0.f);
  const float* w = f->host_weights.data();
  for (int l = 0; l < nl; ++l) {
    const bool scaled = l < ns;
    const int rp = scaled ? rp_s : rp_u;
    const float* W0 = w + f->layer_off[l];
    const float* b0 = W0 + (size_t)d * 128;
    const float* Wt = b0 + 128;
    const float* bt = Wt + (size_t)128 * u;
    const float* Ws = bt + u;
    const float* bs = Ws + (size_t)128 * u;
    float* L = pack.data() + (scaled ? (size_t)l * lf_s : (size_t)ns * lf_s + (size_t)(l - ns) * lf_u);
    for (int j = 0; j < 128; ++j) {
      float* rec = L + (size_t)j * rp;
      rec[0] = b0[j];
      for (int i = 0; i < d; ++i) rec[1 + i] = W0[(size_t)i * 128 + j];
      for (int k = 0; k < u; ++k) rec[1 + d + k] = (float)(c * (double)Wt[(size_t)j * u + k]);
      if (scaled)
        for (int k = 0; k < u; ++k) rec[1 + d + u + k] = (float)(c * (double)Ws[(size_t)j * u + k]);
    }
    // linear part: row 0 = b2 + slope b0 W2, rows 1..d = slope W0 W2 (float64 sums, rounded once)
    float* tl = L + (size_t)128 * rp;
    for (int half = 0; half < (scaled ? 2 : 1); ++half) {
      const float* W2 = half ? Ws : Wt;
      const float* b2 = half ? bs : bt;
      for (int k = 0; k < u; ++k) {
        double v0 = (double)b2[k];
        for (int j = 0; j < 128; ++j) v0 += slope * (double)b0[j] * (double)W2[(size_t)j * u + k];
        tl[half * up + k] = (float)v0;
        for (int i = 0; i < d; ++i) {
          double v = 0.0;
          for (int j = 0; j < 128; ++j) v += (double)W0[(size_t)i * 128 + j] * (double)W2[(size_t)j * u + k];
          tl[(size_t)(1 + i) * 2 * up + half * up + k] = (float)(slope * v);
        }
      }
    }
  }
  SmallCache* cch = new SmallCache();
  HB_CUDA(cudaMalloc(&cch->dev, pack.size() * sizeof(float)));
  HB_CUDA(cudaMemcpy(cch->dev, pack.data(), pack.size() * sizeof(float), cudaMemcpyHostToDevice));
  cch->P.pack = cch->dev;
  cch->P.pre = f->desc.standardize ? f->dev_weights : nullptr;
  cch->P.nlayers = nl;
  cch->P.nscaled = ns;
  cch->P.total_floats = (int)pack.size();
  cch->P.lf_scaled = lf_s;
  cch->P.lf_unscaled = lf_u;
  cch->P.sum_log_amp = f->sum_log_amp;
  f->small_cache = cch;
  *P = cch->P;
  return HB_OK;
}

void realnvp_small_release(Flow* f) {
  if (!f->small_cache) return;
  SmallCache* c = static_cast<SmallCache*>(f->small_cache);
  if (c->dev) cudaFree(c->dev);
  delete c;
  f->small_cache = nullptr;
}

template <int D>
static int small_launch(const SmallParams& P, bool fold, int grid, size_t smem, const void* x, long long ld,
                        const void* lnpost, float* lnpred_out, const FoldParams& fp, long long n,
                        long long rpc, cudaStream_t st) {
  if (fold) {
    auto k = realnvp_small_kernel<D, true>;
    HB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, SM_THREADS, smem, st>>>(P, x, ld, lnpost, lnpred_out, fp, n, rpc);
  } else {
    auto k = realnvp_small_kernel<D, false>;
    HB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, SM_THREADS, smem, st>>>(P, x, ld, lnpost, lnpred_out, fp, n, rpc);
  }
  return HB_OK;
}

// resident CTAs per SM of the instantiation a call will launch (registers and shared memory)
template <int D>
static int small_occupancy(bool fold, size_t smem) {
  int nb = 0;
  cudaError_t e;
  if (fold) {
    auto k = realnvp_small_kernel<D, true>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k, SM_THREADS, smem);
  } else {
    auto k = realnvp_small_kernel<D, false>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k, SM_THREADS, smem);
  }
  return (e == cudaSuccess && nb > 0) ? nb : 1;
}

// returns HB_ERR_UNSUPPORTED (without setting an error) when the shape does not fit this path
static int realnvp_small_run(Flow* f, float T, const void* x, int x_dtype, long long ld, long long n,
                             const void* lnpost, int lp_dtype, float* lnpred_out, Accum* acc,
                             const int64_t* start, long long chain_lo, long long chain_hi, long long a,
                             long long b, cudaStream_t st) {
  const int D = f->desc.ndim;
  if (D < 2 || D > 8) return HB_ERR_UNSUPPORTED;
  SmallParams P;
  int rc = realnvp_small_prepare(f, &P);
  if (rc) return rc;
  const size_t smem = (size_t)P.total_floats * sizeof(float);
  if (smem > 200 * 1024) return HB_ERR_UNSUPPORTED;
  P.T = T;
  P.x_f64 = x_dtype == HB_F64;
  P.lp_f64 = lp_dtype == HB_F64;
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, f->device);
  int per_sm = 1;
  switch (D) {
    case 2: per_sm = small_occupancy<2>(acc != nullptr, smem); break;
    case 3: per_sm = small_occupancy<3>(acc != nullptr, smem); break;
    case 4: per_sm = small_occupancy<4>(acc != nullptr, smem); break;
    case 5: per_sm = small_occupancy<5>(acc != nullptr, smem); break;
    case 6: per_sm = small_occupancy<6>(acc != nullptr, smem); break;
    case 7: per_sm = small_occupancy<7>(acc != nullptr, smem); break;
    default: per_sm = small_occupancy<8>(acc != nullptr, smem); break;
  }
  const long long tile = SM_THREADS * small_r(D);
  long long nfl = (long long)sms * per_sm;
  long long rpc = (n + nfl - 1) / nfl;
  rpc = ((rpc + tile - 1) / tile) * tile;
  nfl = (n + rpc - 1) / rpc;
  FoldParams fp{};
  if (acc) {
    rc = prepare_fold(acc, start, chain_lo, chain_hi, a, b, nfl, rpc, &fp, st);
    if (rc) return rc;
  }
  timing_begin(st);
  switch (D) {
    case 2: rc = small_launch<2>(P, acc != nullptr, (int)nfl, smem, x, ld, lnpost, lnpred_out, fp, n, rpc, st); break;
    case 3: rc = small_launch<3>(P, acc != nullptr, (int)nfl, smem, x, ld, lnpost, lnpred_out, fp, n, rpc, st); break;
    case 4: rc = small_launch<4>(P, acc != nullptr, (int)nfl, smem, x, ld, lnpost, lnpred_out, fp, n, rpc, st); break;
    case 5: rc = small_launch<5>(P, acc != nullptr, (int)nfl, smem, x, ld, lnpost, lnpred_out, fp, n, rpc, st); break;
    case 6: rc = small_launch<6>(P, acc != nullptr, (int)nfl, smem, x, ld, lnpost, lnpred_out, fp, n, rpc, st); break;
    case 7: rc = small_launch<7>(P, acc != nullptr, (int)nfl, smem, x, ld, lnpost, lnpred_out, fp, n, rpc, st); break;
    default: rc = small_launch<8>(P, acc != nullptr, (int)nfl, smem, x, ld, lnpost, lnpred_out, fp, n, rpc, st); break;
  }
  timing_end(st);
  if (rc) return rc;
  count_launch();
  HB_CUDA(cudaGetLastError());
  if (acc) return launch_merge(acc, fp, chain_lo, nfl, st);
  return HB_OK;
}

int realnvp_simt_run(Flow* f, float T, const void* x, int x_dtype, long long ld, long long n,
                     const void* lnpost, int lp_dtype, float* lnpred_out, Accum* acc,
                     const int64_t* start, long long chain_lo, long long chain_hi, long long a,
                     long long b, cudaStream_t st) {
  if (n <= 0) return HB_OK;
  if (!f->force_generic) {
    const int rs = realnvp_small_run(f, T, x, x_dtype, ld, n, lnpost, lp_dtype, lnpred_out, acc, start,
                                     chain_lo, chain_hi, a, b, st);
    if (rs != HB_ERR_UNSUPPORTED) return rs;
  }
  RealNVPSimtParams P;
  int rc = realnvp_simt_prepare(f, &P);
  if (rc) return rc;
  P.T = T;
  const int D = P.D;
  int n2max = 2 * P.u;
  const size_t smem = ((size_t)D + 128 + n2max) * FT * sizeof(float);
  if (smem > 227 * 1024) {
    set_error("RealNVP SIMT path: ndim too large for shared memory");
    return HB_ERR_UNSUPPORTED;
  }
  long long nfl, rpc;
  plan_grid(f->device, n, smem, &nfl, &rpc);
  FoldParams fp{};
  if (acc) {
    rc = prepare_fold(acc, start, chain_lo, chain_hi, a, b, nfl, rpc, &fp, st);
    if (rc) return rc;
  }
  const int grid = (int)nfl;
  timing_begin(st);
  if (acc) {
    HB_SET_SMEM(realnvp_simt_kernel, true);
    HB_DISPATCH_XL(realnvp_simt_kernel, true, lnpred_out, fp, n, rpc);
  } else {
    HB_SET_SMEM(realnvp_simt_kernel, false);
    lp_dtype = HB_F32;
    HB_DISPATCH_XL(realnvp_simt_kernel, false, lnpred_out, fp, n, rpc);
  }
  timing_end(st);
  count_launch();
  HB_CUDA(cudaGetLastError());
  if (acc) return launch_merge(acc, fp, chain_lo, nfl, st);
  return HB_OK;
}


// ---- RQSpline fast path packing ---------------------------------------------------------------
static bool rqspline_fast_eligible(const hb_flow_desc& ds) {
  return ds.n_hidden == 2 && ds.ndim >= 2 && ds.ndim <= 8 && ds.n_bins == 8 && ds.hidden[0] >= 1;
}

static int rqspline_fast_prepare(Flow* f) {
  const hb_flow_desc& ds = f->desc;
  if (!rqspline_fast_eligible(ds)) return HB_OK;
  const int D = ds.ndim, K = ds.n_bins, npar = 3 * K + 1, D4 = (D + 3) & ~3;
  const int h1 = ds.hidden[0], h2 = ds.hidden[1];
  std::vector<float> pack;
  const float* w = f->host_weights.data() + 2 * (size_t)D;
  size_t fe = 0, fo = 0;
  for (int i = 0; i < ds.n_layers; ++i) {
    const size_t begin = pack.size();
    const float* W1 = w; w += (size_t)D * D;
    const float* b1 = w; w += D;
    const float* W2 = w; w += (size_t)D * h1;
    const float* b2 = w; w += h1;
    const float* W3 = w; w += (size_t)h1 * h2;
    const float* b3 = w; w += h2;
    const float* Wout = w; w += (size_t)h2 * D * npar;
    const float* bout = w; w += (size_t)D * npar;
    const float* scales = w; w += D;
    const float* shifts = w; w += D;
    for (int r = 0; r < D; ++r)
      for (int c = 0; c < D4; ++c) pack.push_back(c < D ? W1[(size_t)r * D + c] : 0.f);
    for (int c = 0; c < D4; ++c) pack.push_back(c < D ? b1[c] : 0.f);
    for (int j = 0; j < h1; ++j) {
      for (int c = 0; c < D4; ++c) pack.push_back(c < D ? W2[(size_t)c * h1 + j] : 0.f);
      pack.push_back(b2[j]);
      pack.push_back(0.f); pack.push_back(0.f); pack.push_back(0.f);
    }
    const int par = i & 1;
    for (int j = 0; j < D; ++j) {
      const bool visible = ((j & 1) == 1) != (par == 1);
      if (visible) continue;
      for (int r = 0; r < h1; ++r)
        for (int c = 0; c < RQF_NP; ++c) {
          double acc = 0.0;
          if (c < npar)
            for (int q = 0; q < h2; ++q)
              acc += (double)W3[(size_t)r * h2 + q] * (double)Wout[(size_t)q * D * npar + (size_t)j * npar + c];
          pack.push_back((float)acc);
        }
      for (int c = 0; c < RQF_NP; ++c) {
        double acc = 0.0;
        if (c < npar) {
          acc = bout[(size_t)j * npar + c];
          for (int q = 0; q < h2; ++q) acc += (double)b3[q] * (double)Wout[(size_t)q * D * npar + (size_t)j * npar + c];
        }
        pack.push_back((float)acc);
      }
    }
    for (int c = 0; c < D4; ++c) pack.push_back(c < D ? scales[c] : 1.f);
    for (int c = 0; c < D4; ++c) pack.push_back(c < D ? shifts[c] : 0.f);
    if (par == 0) fe = pack.size() - begin; else fo = pack.size() - begin;
  }
  HB_CUDA(cudaMalloc(&f->rqf_pack, pack.size() * sizeof(float)));
  HB_CUDA(cudaMemcpy(f->rqf_pack, pack.data(), pack.size() * sizeof(float), cudaMemcpyHostToDevice));
  f->rqf_fe = (int)fe;
  f->rqf_fo = (int)(fo ? fo : fe);
  return HB_OK;
}

template <int D>
static int rqf_launch(const RQFastParams& P, bool fold, int grid, const void* x, long long ld,
                      const void* lnpost, float* lnpred_out, const FoldParams& fp, long long n,
                      long long rpc, cudaStream_t st) {
  if (fold) rqspline_fast_kernel<D, true><<<grid, RQF_THREADS, 0, st>>>(P, x, ld, lnpost, lnpred_out, fp, n, rpc);
  else rqspline_fast_kernel<D, false><<<grid, RQF_THREADS, 0, st>>>(P, x, ld, lnpost, lnpred_out, fp, n, rpc);
  return HB_OK;
}

static int rqspline_fast_run(Flow* f, float T, const void* x, int x_dtype, long long ld, long long n,
                             const void* lnpost, int lp_dtype, float* lnpred_out, Accum* acc,
                             const int64_t* start, long long chain_lo, long long chain_hi, long long a,
                             long long b, cudaStream_t st) {
  const hb_flow_desc& ds = f->desc;
  RQFastParams P;
  P.pack = f->rqf_pack;
  P.pre = ds.standardize ? f->dev_weights : nullptr;
  P.nlayers = ds.n_layers;
  P.h1 = ds.hidden[0];
  P.layer_floats_even = f->rqf_fe;
  P.layer_floats_odd = f->rqf_fo;
  P.lo = ds.range_min;
  P.hi = ds.range_max;
  P.T = T;
  P.sum_log_amp = f->sum_log_amp;
  P.x_f64 = x_dtype == HB_F64;
  P.lp_f64 = lp_dtype == HB_F64;
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, f->device);
  const long long tile = RQF_THREADS * RQF_R;
  long long nfl = (long long)sms * 8;
  long long rpc = (n + nfl - 1) / nfl;
  rpc = ((rpc + tile - 1) / tile) * tile;
  nfl = (n + rpc - 1) / rpc;
  FoldParams fp{};
  int rc;
  if (acc) {
    rc = prepare_fold(acc, start, chain_lo, chain_hi, a, b, nfl, rpc, &fp, st);
    if (rc) return rc;
  }
  timing_begin(st);
  switch (ds.ndim) {
    case 2: rqf_launch<2>(P, acc != nullptr, (int)nfl, x, ld, lnpost, lnpred_out, fp, n, rpc, st); break;
    case 3: rqf_launch<3>(P, acc != nullptr, (int)nfl, x, ld, lnpost, lnpred_out, fp, n, rpc, st); break;
    case 4: rqf_launch<4>(P, acc != nullptr, (int)nfl, x, ld, lnpost, lnpred_out, fp, n, rpc, st); break;
    case 5: rqf_launch<5>(P, acc != nullptr, (int)nfl, x, ld, lnpost, lnpred_out, fp, n, rpc, st); break;
    case 6: rqf_launch<6>(P, acc != nullptr, (int)nfl, x, ld, lnpost, lnpred_out, fp, n, rpc, st); break;
    case 7: rqf_launch<7>(P, acc != nullptr, (int)nfl, x, ld, lnpost, lnpred_out, fp, n, rpc, st); break;
    default: rqf_launch<8>(P, acc != nullptr, (int)nfl, x, ld, lnpost, lnpred_out, fp, n, rpc, st); break;
  }
  timing_end(st);
  count_launch();
  HB_CUDA(cudaGetLastError());
  if (acc) return launch_merge(acc, fp, chain_lo, nfl, st);
  return HB_OK;
}

// ---- RQSpline packing ---------------------------------------------------------------------
int rqspline_prepare(Flow* f) {
  const hb_flow_desc& ds = f->desc;
  const int D = ds.ndim, K = ds.n_bins, npar = 3 * K + 1, npar_pad = pad8(npar);
  const int nmlp = ds.n_hidden + 1;  // Dense(D), Dense(h1), ..., Dense(h_last)
  std::vector<int> width(nmlp + 1);
  width[0] = D;
  width[1] = D;
  for (int m = 0; m < ds.n_hidden; ++m) width[m + 2] = ds.hidden[m];
  const int nact = nmlp - 1;
  const int in_last = width[nmlp - 1];   // input width of the last MLP dense
  const int out_last = width[nmlp];      // its output width (= input of the output dense)
  std::vector<float> pack;
  std::vector<RQLayer> layers(ds.n_layers);
  auto align4 = [&]() { while (pack.size() % 4) pack.push_back(0.f); };
  const float* w = f->host_weights.data() + 2 * (size_t)D;
  for (int i = 0; i < ds.n_layers; ++i) {
    RQLayer& L = layers[i];
    std::vector<const float*> Wm(nmlp), bm(nmlp);
    for (int m = 0; m < nmlp; ++m) {
      Wm[m] = w;
      w += (size_t)width[m] * width[m + 1];
      bm[m] = w;
      w += width[m + 1];
    }
    const float* Wout = w;
    w += (size_t)out_last * D * npar;
    const float* bout = w;
    w += (size_t)D * npar;
    const float* scales = w;
    w += D;
    const float* shifts = w;
    w += D;
    for (int m = 0; m < nact; ++m) {
      const int wp = pad8(width[m + 1]);
      align4();
      L.w[m] = (int)pack.size();
      for (int r = 0; r < width[m]; ++r)
        for (int c = 0; c < wp; ++c) pack.push_back(c < width[m + 1] ? Wm[m][(size_t)r * width[m + 1] + c] : 0.f);
      L.b[m] = (int)pack.size();
      for (int c = 0; c < wp; ++c) pack.push_back(c < width[m + 1] ? bm[m][c] : 0.f);
    }
    // fused last MLP dense x output dense, per transformed feature (float64 product)
    align4();
    L.fused = (int)pack.size();
    const int par = i & 1;
    for (int j = 0; j < D; ++j) {
      const bool visible = ((j & 1) == 1) != (par == 1);
      if (visible) continue;
      const float* Wl = Wm[nmlp - 1];
      const float* bl = bm[nmlp - 1];
      for (int r = 0; r < in_last; ++r)
        for (int c = 0; c < npar_pad; ++c) {
          double acc = 0.0;
          if (c < npar)
            for (int q = 0; q < out_last; ++q)
              acc += (double)Wl[(size_t)r * out_last + q] * (double)Wout[(size_t)q * D * npar + (size_t)j * npar + c];
          pack.push_back((float)acc);
        }
      for (int c = 0; c < npar_pad; ++c) {
        double acc = 0.0;
        if (c < npar) {
          acc = bout[(size_t)j * npar + c];
          for (int q = 0; q < out_last; ++q)
            acc += (double)bl[q] * (double)Wout[(size_t)q * D * npar + (size_t)j * npar + c];
        }
        pack.push_back((float)acc);
      }
    }
    align4();
    L.scales = (int)pack.size();
    pack.insert(pack.end(), scales, scales + D);
    L.shifts = (int)pack.size();
    pack.insert(pack.end(), shifts, shifts + D);
  }
  const size_t lay_bytes = layers.size() * sizeof(RQLayer);
  const size_t lay_floats = (lay_bytes + 3) / 4;
  align4();
  const size_t lay_off = pack.size();
  pack.resize(pack.size() + lay_floats);
  memcpy(pack.data() + lay_off, layers.data(), lay_bytes);
  HB_CUDA(cudaMalloc(&f->rq_pack, pack.size() * sizeof(float)));
  HB_CUDA(cudaMemcpy(f->rq_pack, pack.data(), pack.size() * sizeof(float), cudaMemcpyHostToDevice));
  f->rq_pack_floats = pack.size();
  f->rq_off.assign(1, lay_off);
  const int rc = rqspline_fast_prepare(f);
  if (rc) return rc;
  return rqspline_tc_prepare(f);
}

// kernel parameters of the generic RQSpline kernels (log_prob and sampling direction)
static void rq_params(Flow* f, float T, RQParams* out) {
  const hb_flow_desc& ds = f->desc;
  RQParams& P = *out;
  P.pack = f->rq_pack;
  P.pre = ds.standardize ? f->dev_weights : nullptr;
  P.D = ds.ndim;
  P.K = ds.n_bins;
  P.npar = 3 * ds.n_bins + 1;
  P.npar_pad = pad8(P.npar);
  P.nlayers = ds.n_layers;
  const int nmlp = ds.n_hidden + 1;
  P.nact = nmlp - 1;
  P.width[0] = ds.ndim;
  P.width[1] = ds.ndim;
  for (int m = 0; m < ds.n_hidden; ++m) P.width[m + 2] = ds.hidden[m];
  P.maxw = 1;
  for (int m = 0; m <= nmlp; ++m) {
    P.wpad[m] = pad8(P.width[m]);
    if (m >= 1 && m <= P.nact && P.width[m] > P.maxw) P.maxw = P.width[m];
  }
  P.in_last = P.width[nmlp - 1];
  P.lo = ds.range_min;
  P.hi = ds.range_max;
  P.T = T;
  P.sum_log_amp = f->sum_log_amp;
  P.layers = reinterpret_cast<const RQLayer*>(f->rq_pack + f->rq_off[0]);
}

int rqspline_run(Flow* f, float T, const void* x, int x_dtype, long long ld, long long n,
                 const void* lnpost, int lp_dtype, float* lnpred_out, Accum* acc,
                 const int64_t* start, long long chain_lo, long long chain_hi, long long a,
                 long long b, cudaStream_t st) {
  if (n <= 0) return HB_OK;
  if (f->path == HB_PATH_TCGEN05 || (f->path == HB_PATH_AUTO && f->rqtc))
    return rqspline_tc_run(f, T, x, x_dtype, ld, n, lnpost, lp_dtype, lnpred_out, acc, start, chain_lo, chain_hi,
                           a, b, st);
  if (f->rqf_pack && !f->force_generic)
    return rqspline_fast_run(f, T, x, x_dtype, ld, n, lnpost, lp_dtype, lnpred_out, acc, start, chain_lo,
                             chain_hi, a, b, st);
  RQParams P;
  rq_params(f, T, &P);
  const size_t smem = ((size_t)P.D + 2 * (size_t)P.maxw + P.npar_pad) * FT * sizeof(float);
  if (smem > 227 * 1024) {
    set_error("RQSpline: ndim / hidden sizes too large for shared memory");
    return HB_ERR_UNSUPPORTED;
  }
  long long nfl, rpc;
  plan_grid(f->device, n, smem, &nfl, &rpc);
  FoldParams fp{};
  int rc;
  if (acc) {
    rc = prepare_fold(acc, start, chain_lo, chain_hi, a, b, nfl, rpc, &fp, st);
    if (rc) return rc;
  }
  const int grid = (int)nfl;
  timing_begin(st);
  if (acc) {
    HB_SET_SMEM(rqspline_simt_kernel, true);
    HB_DISPATCH_XL(rqspline_simt_kernel, true, lnpred_out, fp, n, rpc);
  } else {
    HB_SET_SMEM(rqspline_simt_kernel, false);
    lp_dtype = HB_F32;
    HB_DISPATCH_XL(rqspline_simt_kernel, false, lnpred_out, fp, n, rpc);
  }
  timing_end(st);
  count_launch();
  HB_CUDA(cudaGetLastError());
  if (acc) return launch_merge(acc, fp, chain_lo, nfl, st);
  return HB_OK;
}


// ------------------------------------------------------------------------------------------
// Sampling direction: FlowModel.sample (harmonic/model.py:239-275) with the standard-normal draws supplied
// by the caller.  RealNVP (flows.py:118-138): tfb.Chain.forward applies [S0, P, S1, ..] right to left, a
// coupling's forward is y1 = x1 * scale(x0) + shift(x0), base = T * eps.  RQSpline (flows.py:290-313):
// Inverse(Chain).forward = the inverses of [affine_0, spline_0, affine_1, ..] in list order,
// base = sqrt(T) * eps.  Same shared-memory layout and weight packs as the log_prob kernels above.
// ------------------------------------------------------------------------------------------
template <typename ET>
__global__ void __launch_bounds__(FT)
realnvp_sample_kernel(RealNVPSimtParams P, const ET* __restrict__ eps, long long ld,
                      float* __restrict__ out, long long ldo, long long nrows) {
  extern __shared__ float smem[];
  const int tid = threadIdx.x;
  const int D = P.D, d = P.d, u = P.u;
  float* ys = smem;
  float* hs = ys + (size_t)D * FT;
  float* os = hs + (size_t)128 * FT;
  for (long long t0 = (long long)blockIdx.x * FT; t0 < nrows; t0 += (long long)gridDim.x * FT) {
    const long long r = t0 + tid;
    const bool live = r < nrows;
    const long long rr = live ? r : t0;
    {
      const int rot = (P.nlayers - 1) % D;  // base space: logical i lives in physical row (i + L - 1) mod D
      for (int i = 0; i < D; ++i) {
        int k = i + rot;
        if (k >= D) k -= D;
        ys[k * FT + tid] = P.T * (float)eps[rr * ld + i];
      }
    }
    for (int l = P.nlayers - 1; l >= 0; --l) {
      const SimtLayer L = P.layers[l];
      const int rot = l % D;
      auto phys = [&](int i) { int k = i + rot; return k >= D ? k - D : k; };
      auto all = [](int) { return true; };
      dense_cols<2>(ys, phys, all, d, P.pack + L.w0, 128, P.pack + L.b0, 128, hs, tid);
      dense_cols<0>(hs, [](int i) { return i; }, all, 128, P.pack + L.w2, L.n2pad, P.pack + L.b2,
                    L.n2, os, tid);
      for (int k = 0; k < u; ++k) {
        const int pk = phys(d + k);
        float y1 = ys[pk * FT + tid];
        if (L.scaled) {
          float sc = softplus_acc(os[(u + k) * FT + tid]);
          sc = fminf(fmaxf(sc, 1e-3f), 1e3f);
          y1 = y1 * sc;
        }
        ys[pk * FT + tid] = y1 + os[k * FT + tid];
      }
    }
    if (live)
      for (int i = 0; i < D; ++i) {
        float v = ys[i * FT + tid];
        if (P.pre) v = fmaf(v, P.pre[D + i], P.pre[i]);  // samples * pre_amp + pre_offset (model.py:272-273)
        out[r * ldo + i] = v;
      }
  }
}

template <typename ET>
__global__ void __launch_bounds__(FT)
rqspline_sample_kernel(RQParams P, const ET* __restrict__ eps, long long ld, float* __restrict__ out,
                       long long ldo, long long nrows) {
  extern __shared__ float smem[];
  const int tid = threadIdx.x;
  const int D = P.D, K = P.K;
  float* ys = smem;
  float* bufA = ys + (size_t)D * FT;
  float* bufB = bufA + (size_t)P.maxw * FT;
  float* ps = bufB + (size_t)P.maxw * FT;
  auto ident = [](int i) { return i; };
  auto all = [](int) { return true; };
  const float sT = sqrtf(P.T);  // MultivariateNormalFullCovariance(cov = T I): scale sqrt(T) (flows.py:264-269)
  for (long long t0 = (long long)blockIdx.x * FT; t0 < nrows; t0 += (long long)gridDim.x * FT) {
    const long long r = t0 + tid;
    const bool live = r < nrows;
    const long long rr = live ? r : t0;
    for (int i = 0; i < D; ++i) ys[i * FT + tid] = sT * (float)eps[rr * ld + i];
    for (int i = 0; i < P.nlayers; ++i) {
      const RQLayer L = P.layers[i];
      const int par = i & 1;
      auto visible = [&](int j) { return ((j & 1) == 1) != (par == 1); };
      // scalar affine inverse, all dims
      for (int j = 0; j < D; ++j) {
        const float sc = __ldg(P.pack + L.scales + j), sh = __ldg(P.pack + L.shifts + j);
        ys[j * FT + tid] = (ys[j * FT + tid] - sh) / sc;
      }
      // conditioner on the pass-through features (unchanged by this layer in either direction)
      const float* cur = ys;
      float* nxt = bufA;
      for (int m = 0; m < P.nact; ++m) {
        if (m == 0)
          dense_cols<1>(cur, ident, visible, P.width[0], P.pack + L.w[0], P.wpad[1], P.pack + L.b[0],
                        P.width[1], nxt, tid);
        else
          dense_cols<1>(cur, ident, all, P.width[m], P.pack + L.w[m], P.wpad[m + 1], P.pack + L.b[m],
                        P.width[m + 1], nxt, tid);
        cur = nxt;
        nxt = (nxt == bufA) ? bufB : bufA;
      }
      int blk = 0;
      const size_t blk_floats = (size_t)P.in_last * P.npar_pad + P.npar_pad;
      for (int j = 0; j < D; ++j) {
        if (visible(j)) continue;
        const float* Mj = P.pack + L.fused + blk * blk_floats;
        const float* cj = Mj + (size_t)P.in_last * P.npar_pad;
        if (P.nact == 0)
          dense_cols<0>(ys, ident, visible, P.in_last, Mj, P.npar_pad, cj, P.npar, ps, tid);
        else
          dense_cols<0>(cur, ident, all, P.in_last, Mj, P.npar_pad, cj, P.npar, ps, tid);
        ys[j * FT + tid] = rqs_inverse(ys[j * FT + tid], ps, tid, K, P.lo, P.hi);
        ++blk;
      }
    }
    if (live)
      for (int i = 0; i < D; ++i) {
        float v = ys[i * FT + tid];
        if (P.pre) v = fmaf(v, P.pre[D + i], P.pre[i]);
        out[r * ldo + i] = v;
      }
  }
}

static int sample_grid(int device, long long n, size_t smem) {
  long long nfl, rpc;
  plan_grid(device, n, smem, &nfl, &rpc);
  const long long tiles = (n + FT - 1) / FT;
  return (int)(nfl < tiles ? nfl : tiles);
}

int flow_sample_run(Flow* f, float T, const void* eps, int eps_dtype, long long ld, long long n, float* out,
                    long long ldo, cudaStream_t st) {
  if (n <= 0) return HB_OK;
  const hb_flow_desc& ds = f->desc;
  if (ds.kind == HB_FLOW_REALNVP) {
    RealNVPSimtParams P;
    int rc = realnvp_simt_prepare(f, &P);
    if (rc) return rc;
    P.T = T;
    const size_t smem = ((size_t)P.D + 128 + 2 * (size_t)P.u) * FT * sizeof(float);
    if (smem > 227 * 1024) {
      set_error("RealNVP sample: ndim too large for shared memory");
      return HB_ERR_UNSUPPORTED;
    }
    const int grid = sample_grid(f->device, n, smem);
    cudaFuncSetAttribute(realnvp_sample_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(realnvp_sample_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    timing_begin(st);
    if (eps_dtype == HB_F32)
      realnvp_sample_kernel<float><<<grid, FT, smem, st>>>(P, (const float*)eps, ld, out, ldo, n);
    else
      realnvp_sample_kernel<double><<<grid, FT, smem, st>>>(P, (const double*)eps, ld, out, ldo, n);
    timing_end(st);
  } else {
    RQParams P;
    rq_params(f, T, &P);
    const size_t smem = ((size_t)P.D + 2 * (size_t)P.maxw + P.npar_pad) * FT * sizeof(float);
    if (smem > 227 * 1024) {
      set_error("RQSpline sample: ndim / hidden sizes too large for shared memory");
      return HB_ERR_UNSUPPORTED;
    }
    const int grid = sample_grid(f->device, n, smem);
    cudaFuncSetAttribute(rqspline_sample_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(rqspline_sample_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    timing_begin(st);
    if (eps_dtype == HB_F32)
      rqspline_sample_kernel<float><<<grid, FT, smem, st>>>(P, (const float*)eps, ld, out, ldo, n);
    else
      rqspline_sample_kernel<double><<<grid, FT, smem, st>>>(P, (const double*)eps, ld, out, ldo, n);
    timing_end(st);
  }
  count_launch();
  HB_CUDA(cudaGetLastError());
  return HB_OK;
}

}  // namespace hb
