// RealNVP log_prob (+ fused harmonic-mean fold) for small ndim (2..8) with the coupling MLP's second contraction on the
// tcgen05 tensor cores (harmonic/flows.py:41-116,159-180; the Pima-Indian example and BASELINE cfg3 are ndim 5).
//
// A coupling layer conditions on d = round(ndim / 2) features and has a fixed hidden width of 128 (flows.py:171):
//   H = y0 W0 + b0                         128 x d FMAs per sample: too thin for a GEMM, CUDA cores
//   [shift | scale pre-activation] = leaky_relu(H) [Wt | Ws] + [bt | bs]     128 x 2u: K = 128, N <= 8 -> TENSOR CORES
// With leaky_relu(h) = s h + c relu(h) (s = 0.01, c = 0.99) the linear part s (y0 W0 + b0) W2 + b2 = y0 V + v0 is a
// d x 2u matrix formed once on the host in float64 (as in realnvp_small_kernel) and the contraction that remains is
// relu(H) (c W2).  It runs as the three-product FP16 split of hb_realnvp_tc.cu: r = relu(h) as
// hs = fp16(r) rounded towards zero and lo = fp16(r - hs), both by conversions whose clamp is a modifier
// (cvt.rz.relu / cvt.rn.relu) and one mixed-precision FMA per element; c W2 = w_hi + w_lo travelling times 2^5;
// tcgen05.mma kind::f16, M = 128 samples, N = 16, A from TMEM, FP32 accumulation.  On the CUDA cores
// (realnvp_small_kernel) the contraction is 6 of the ~9.75 instructions per hidden unit and sample of a scaled ndim-5
// layer; here a hidden unit costs d FMAs, 2 instructions of the split and one LDS.
//
// Range: hs = fp16(relu(h)) needs h < 65504.  |h| <= max_j sum_i |W0[i][j]| max|y0| + max|b0|; every epilogue thread
// tracks max |conditioning input| of its row over the layers, rows beyond the limit derived at flow creation are
// counted into the flow's guard counter and HB_PATH_AUTO answers by evaluating the call again on the CUDA cores
// (hb_api.cu, as for the ndim-64 kernel).  NaN inputs do not trip the guard: they propagate (relu(NaN) stays NaN in
// both conversions).
//
// One persistent CTA per SM: four epilogue warpgroups, each owning one 128-sample tile (TMEM lane = sample), and one
// auxiliary warpgroup whose warp t issues tile t's 24 MMAs per layer.  All layers' weights are resident in shared memory
// (8 KiB of FP16 blocks + 2-4 KiB of float32 records per layer).  Seven mbarriers per tile: one per K = 32 quarter of A (the
// quarter's 6 MMAs run while the epilogue computes the next one), two "buffer free" commits and OFULL (tcgen05.commit after the
// last quarter).
// TMEM per tile (80 columns): A (two K = 32 quarter buffers of 32 columns: per 16 hidden units 8 columns of packed lo pairs + 8
// of packed hs pairs; quarters 2 / 3 reuse the buffers of quarters 0 / 1 once their MMAs have completed -- with all four
// quarters resident only three tiles fit TMEM, and three warps per scheduler left the issue slots 39 % idle) | O (16).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "hb_internal.h"
#include "hb_tc_ptx.cuh"

namespace hb {

namespace stc {

using namespace tcx;

constexpr int TILE = 128;
constexpr int NT = 4;
constexpr int THREADS = NT * 128 + 128;
constexpr int HID = 128;
constexpr int NO = 16;                  // GEMM N: 2u <= 8 outputs padded to 16
constexpr float SU = 32.0f, SU_INV = 1.0f / SU;
constexpr int Q_BYTES = NO * 64 * 2;    // one K = 32 quarter: [16 x 64 K'], K' < 32 w_hi, K' >= 32 w_lo
constexpr int L_BYTES = 4 * Q_BYTES;    // per layer
constexpr uint32_t TM_A = 0, TM_O = 64, TM_TILE = 64 + NO;  // A: two K = 32 quarter buffers (quarter q lives in buffer q & 1)
static_assert(NT * TM_TILE <= 512, "TMEM");
constexpr int B_AQ = 0 /* [quarter][tile]: one K = 32 quarter of A has been written */,
              B_QD = 4 * NT /* [buffer][tile]: the MMAs of quarter 0 / 1 have completed, the buffer may be rewritten */,
              B_OFULL = 6 * NT, B_COUNT = 7 * NT;
__host__ __device__ constexpr int small_d(int D) { return (D == 2) ? 1 : (D == 3) ? 2 : (D == 4) ? 2 : (D == 5) ? 2 : (D == 6) ? 3 : 4; }
__host__ __device__ constexpr int rec_floats(int D) { return 1 + small_d(D); }  // {b0[j], W0[0..d-1][j]}, unpadded: four
// records are (1 + d) LDS.128 (shared-memory wavefronts bound this loop, profiles/README.md)
// float32 block of a layer: 128 records, then the linear part: (1 + d) rows of 2 x (4-padded u) floats
__host__ __device__ constexpr int scal_floats(int D) {
  return HID * rec_floats(D) + (1 + small_d(D)) * 2 * (((D - small_d(D)) + 3) & ~3);
}

struct Params {
  const uint8_t* bimg;  // [layer] of L_BYTES
  const float* scal;    // [layer][scal_floats(D)]
  const float* pre;     // pre_offset[D], pre_amp[D] or nullptr
  int nlayers, nscaled;
  float T, sum_log_amp;
  unsigned long long* guard;
  float guard_limit;
  long long nrows, rows_per_wg;
  int nbatches;
  int x_f64, lp_f64;
};

// clip(softplus(a), 1e-3, 1e3) (harmonic/flows.py:177) as in hb_realnvp_tc.cu: t = exp(-|a|) by MUFU ex2,
// log1p(t) = 2 atanh(t / (2 + t)) as a degree-4 polynomial in the square (relative error 8e-9)
__device__ __forceinline__ float softplus_clip_fast(float a) {
  const float t = ex2_approx(-1.4426950408889634f * fabsf(a));
  const float uu = t * rcp_approx(2.0f + t);
  const float z = uu * uu;
  const float p = fmaf(z, fmaf(z, fmaf(z, fmaf(z, 0.14087678f, 0.13980642f), 0.20012431f), 0.33333158f), 1.0f);
  return fminf(fmaxf(fmaf(uu + uu, p, fmaxf(a, 0.f)), 1e-3f), 1e3f);
}

template <int D, bool FOLD>
__global__ void __launch_bounds__(THREADS, 1)
realnvp_small_tc_kernel(Params P, const void* __restrict__ xv, long long ld, const void* __restrict__ lnpostv,
                        float* __restrict__ lnpred_out, FoldParams fp) {
  constexpr int d = small_d(D), u = D - d, UP = (u + 3) & ~3, RF = rec_floats(D), SF = scal_floats(D);
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int nl = P.nlayers;
  const uint32_t img_bytes = (uint32_t)nl * L_BYTES;
  uint8_t* s_img = smem;
  float* s_scal = reinterpret_cast<float*>(smem + img_bytes);
  uint8_t* s_tail = smem + img_bytes + (((uint32_t)nl * SF * 4 + 15) & ~15u);
  const uint32_t bar0 = smem_u32(s_tail);
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(s_tail + 8 * B_COUNT);
  double* scratch_all = reinterpret_cast<double*>(s_tail + 8 * B_COUNT + 16);

  if (threadIdx.x == 0) {
    for (int t = 0; t < NT; ++t) {
      for (int q = 0; q < 4; ++q) mbar_init(BAR(B_AQ + NT * q + t), 128);
      for (int q = 0; q < 2; ++q) mbar_init(BAR(B_QD + NT * q + t), 1);
      mbar_init(BAR(B_OFULL + t), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  {
    const uint4* src = reinterpret_cast<const uint4*>(P.bimg);
    uint4* dst = reinterpret_cast<uint4*>(s_img);
    for (uint32_t i = threadIdx.x; i < img_bytes / 16; i += THREADS) dst[i] = src[i];
    for (int i = threadIdx.x; i < nl * SF; i += THREADS) s_scal[i] = P.scal[i];
    fence_proxy_async();  // the MMAs read the blocks through the async proxy
  }
  if (warp == NT * 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32((const void*)tmem_ptr)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp >= NT * 4) {
    // ===== MMA issuers: warp NT * 4 + t serves tile t =====
    const int t = warp - NT * 4;
    if (t < NT) {
      const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0) + t * TM_TILE;
      const uint32_t idesc = make_idesc_f16(NO);
      const uint32_t lbo = NO * 16;
      const uint64_t kstep = (uint64_t)((2 * lbo) >> 4);
      const uint32_t img0 = smem_u32(s_img);
      uint32_t it = 0;
      for (int b = 0; b < P.nbatches; ++b) {
        for (int l = 0; l < nl; ++l, ++it) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {  // quarter q as soon as the epilogue has written it: the MMAs run under the
            mbar_wait(BAR(B_AQ + NT * q + t), it & 1);  // epilogue's work on quarter q + 1
            tc_fence_after();
            if (elect_one()) {
              const uint64_t bd = make_desc(img0 + (uint32_t)l * L_BYTES + (uint32_t)q * Q_BYTES, lbo, 128);
              const uint32_t a0 = tb + TM_A + 32 * (q & 1);
#pragma unroll
              for (int half = 0; half < 2; ++half) {
                const uint32_t a_lo = a0 + 16 * half, a_hs = a_lo + 8;
                mma_ts_f16(tb + TM_O, a_lo, bd + (uint64_t)half * kstep, idesc, (q | half) ? 1u : 0u);
                mma_ts_f16(tb + TM_O, a_hs, bd + (uint64_t)(2 + half) * kstep, idesc, 1);
                mma_ts_f16(tb + TM_O, a_hs, bd + (uint64_t)half * kstep, idesc, 1);
              }
              if (q < 2) tc_commit(BAR(B_QD + NT * q + t));
              if (q == 3) tc_commit(BAR(B_OFULL + t));
            }
            __syncwarp();
          }
        }
      }
    }
  } else {
    // ===== epilogue warpgroups: one sample per thread =====
    const int wg = warp >> 2;
    const int tid = threadIdx.x & 127;
    const uint32_t lane_base = ((uint32_t)((warp & 3) * 32)) << 16;
    const uint32_t tm_a = tmem_base + lane_base + wg * TM_TILE + TM_A;
    const uint32_t tm_o = tmem_base + lane_base + wg * TM_TILE + TM_O;
    double* scratch = scratch_all + wg * 12;
    const int flusher = blockIdx.x * NT + wg;
    const long long row_lo = (long long)flusher * P.rows_per_wg;
    long long row_hi = row_lo + P.rows_per_wg;
    if (row_hi > P.nrows) row_hi = P.nrows;
    const bool had_rows = row_lo < P.nrows;
    Fold<128> fold(fp, flusher, tid, scratch, (FOLD && had_rows) ? row_lo : 0, 1 + wg);  // named barriers 1..NT
    const float invT = 1.f / P.T;
    const float base_const = -(float)D * logf(P.T) - 0.5f * (float)D * 1.8378770664093453f;
    uint32_t it = 0;
    unsigned int nguard = 0;
    for (int b = 0; b < P.nbatches; ++b) {
      const long long t0 = row_lo + (long long)b * TILE;
      const long long r = t0 + tid;
      const bool live = r < row_hi;
      const long long rr = live ? r : (had_rows ? row_lo : 0);
      float y[D];
#pragma unroll
      for (int i = 0; i < D; ++i) {
        float v = P.x_f64 ? (float)static_cast<const double*>(xv)[rr * ld + i]
                          : static_cast<const float*>(xv)[rr * ld + i];
        if (P.pre) v = (v - P.pre[i]) / P.pre[D + i];
        y[i] = v;
      }
      float lp = 0.f;
      if (FOLD && live)
        lp = P.lp_f64 ? (float)static_cast<const double*>(lnpostv)[r] : static_cast<const float*>(lnpostv)[r];
      float ldj = 0.f, gmax = 0.f;
      for (int l = 0; l < nl; ++l, ++it) {
        const bool scaled = l < P.nscaled;
        const float* S = s_scal + l * SF;
#pragma unroll
        for (int i = 0; i < d; ++i) gmax = fmaxf(gmax, fabsf(y[i]));  // fmaxf drops NaN: NaN rows do not trip the guard
        // ---- relu(H) and its FP16 split, 16 hidden units per TMEM chunk
#pragma unroll 1
        for (int c = 0; c < HID / 16; ++c) {
          if (c == 4 || c == 6) {  // quarters 2 / 3 reuse the buffers of quarters 0 / 1: their MMAs must have completed
            mbar_wait(BAR(B_QD + NT * ((c >> 1) & 1) + wg), it & 1);
            tc_fence_after();
          }
          const uint32_t ta = tm_a + 32 * ((c >> 1) & 1) + 16 * (c & 1);
          uint32_t wl[8], wh[8];
#pragma unroll
          for (int g = 0; g < 4; ++g) {  // four hidden units per group: their records are 4 (1 + d) contiguous floats
            const float4* rec = reinterpret_cast<const float4*>(S + (16 * c + 4 * g) * RF);
            float w[4 * RF];
#pragma unroll
            for (int i4 = 0; i4 < RF; ++i4) {
              const float4 t4 = rec[i4];
              w[4 * i4] = t4.x; w[4 * i4 + 1] = t4.y; w[4 * i4 + 2] = t4.z; w[4 * i4 + 3] = t4.w;
            }
            float h[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float acc = w[e * RF];
#pragma unroll
              for (int i = 0; i < d; ++i) acc = fmaf(y[i], w[e * RF + 1 + i], acc);
              h[e] = acc;
            }
#pragma unroll
            for (int pr = 0; pr < 2; ++pr) {
              const uint32_t hs = pack_relu_rz_f16x2(h[2 * pr + 1], h[2 * pr]);
              float l0, l1;
              split_rest<false>(hs, h[2 * pr], h[2 * pr + 1], l0, l1);
              wh[2 * g + pr] = hs;
              wl[2 * g + pr] = pack_relu_f16x2(l1, l0);
            }
          }
          tmem_st8(ta, wl);
          tmem_st8(ta + 8, wh);
          if (c & 1) {  // a K = 32 quarter is complete
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(BAR(B_AQ + NT * (c >> 1) + wg));
          }
        }
        // ---- the linear part y0 V + v0 while the MMAs run
        const float* tail = S + HID * RF;
        float sh[u], sc[u];
#pragma unroll
        for (int k = 0; k < u; ++k) {
          float t = tail[k], s = tail[UP + k];
#pragma unroll
          for (int i = 0; i < d; ++i) {
            t = fmaf(y[i], tail[(1 + i) * 2 * UP + k], t);
            s = fmaf(y[i], tail[(1 + i) * 2 * UP + UP + k], s);
          }
          sh[k] = t;
          sc[k] = s;
        }
        mbar_wait(BAR(B_OFULL + wg), it & 1);
        tc_fence_after();
        uint32_t o[16];
        tmem_ld16(tm_o, o);
        tmem_ld_wait();
        tc_fence_before();  // orders this load before the next layer's MMAs (issued after the next AFULL)
        float prod = 1.f;  // u <= 4 factors in [1e-3, 1e3] stay inside the float32 range
#pragma unroll
        for (int k = 0; k < u; ++k) {
          float y1 = y[d + k] - fmaf(__uint_as_float(o[k]), SU_INV, sh[k]);
          if (scaled) {
            const float a = fmaf(__uint_as_float(o[u + k]), SU_INV, sc[k]);
            float s = softplus_clip_fast(a);
            s = (a != a) ? a : s;  // keep NaN (fminf / fmaxf would drop it)
            y1 *= rcp_approx(s);
            prod *= s;
          }
          y[d + k] = y1;
        }
        if (scaled) ldj -= 0.6931471805599453f * lg2_approx(prod);
        if (l + 1 < nl) {  // inverse permute: rotate left by one
          const float y0 = y[0];
#pragma unroll
          for (int i = 0; i < D - 1; ++i) y[i] = y[i + 1];
          y[D - 1] = y0;
        }
      }
      float ss = 0.f;
#pragma unroll
      for (int i = 0; i < D; ++i) {
        const float z = y[i] * invT;
        ss = fmaf(z, z, ss);
      }
      const float lnphi = -0.5f * ss + base_const + ldj - P.sum_log_amp;
      nguard += (live && gmax > P.guard_limit) ? 1u : 0u;
      if (live && lnpred_out) lnpred_out[r] = lnphi;
      if (FOLD) {
        float a[1] = {lnphi};
        float bb[1] = {lp};
        long long t1 = t0 + TILE;
        if (t1 > row_hi) t1 = row_hi;
        if (t0 < row_hi) fold.tile<1>(t0, t1, a, bb);
      }
    }
    if (FOLD) fold.finish(had_rows);
    if (P.guard && nguard) atomicAdd(P.guard, (unsigned long long)nguard);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == NT * 4) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

}  // namespace stc

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
using namespace tcx;

struct SmallTcCache {
  uint8_t* bimg = nullptr;
  float* scal = nullptr;
  size_t smem = 0;
  float guard_limit = 0.f;
};

// Builds the resident weight image of the small-ndim tensor path; leaves f->stc null (the register kernel serves the
// flow) when a block leaves the half range or the image does not fit shared memory.
int realnvp_small_tc_prepare(Flow* f) {
  const hb_flow_desc& ds = f->desc;
  const int D = ds.ndim;
  if (ds.kind != HB_FLOW_REALNVP || D < 2 || D > 8) return HB_OK;
  const int d = f->d, u = f->u, nl = f->nlayers, ns = ds.n_scaled;
  if (d != stc::small_d(D) || nl < 1) return HB_OK;
  const int RF = stc::rec_floats(D), SF = stc::scal_floats(D), UP = (u + 3) & ~3;
  const size_t img_bytes = (size_t)nl * stc::L_BYTES;
  const size_t smem = img_bytes + (((size_t)nl * SF * 4 + 15) & ~(size_t)15) + 8 * stc::B_COUNT + 16 +
                      stc::NT * 12 * sizeof(double);
  if (smem > 200 * 1024) return HB_OK;
  const double slope = 0.01, c = 1.0 - slope;  // leaky_relu(h) = slope h + c relu(h)
  std::vector<uint8_t> img(img_bytes, 0);
  std::vector<float> scal((size_t)nl * SF, 0.f);
  const float* w = f->host_weights.data();
  bool overflow = false;
  double hsum_max = 0.0, hbias_max = 0.0;
  for (int l = 0; l < nl; ++l) {
    const bool scaled = l < ns;
    const float* W0 = w + f->layer_off[l];
    const float* b0 = W0 + (size_t)d * 128;
    const float* Wt = b0 + 128;
    const float* bt = Wt + (size_t)128 * u;
    const float* Ws = bt + u;
    const float* bs = Ws + (size_t)128 * u;
    float* S = scal.data() + (size_t)l * SF;
    for (int j = 0; j < 128; ++j) {
      float* rec = S + (size_t)j * RF;
      rec[0] = b0[j];
      double sum = 0.0;
      for (int i = 0; i < d; ++i) {
        rec[1 + i] = W0[(size_t)i * 128 + j];
        sum += fabs((double)W0[(size_t)i * 128 + j]);
      }
      if (sum > hsum_max) hsum_max = sum;
      if (fabs((double)b0[j]) > hbias_max) hbias_max = fabs((double)b0[j]);
    }
    float* tl = S + (size_t)128 * RF;
    for (int half = 0; half < (scaled ? 2 : 1); ++half) {
      const float* W2 = half ? Ws : Wt;
      const float* b2 = half ? bs : bt;
      for (int k = 0; k < u; ++k) {
        double v0 = (double)b2[k];
        for (int j = 0; j < 128; ++j) v0 += slope * (double)b0[j] * (double)W2[(size_t)j * u + k];
        tl[half * UP + k] = (float)v0;
        for (int i = 0; i < d; ++i) {
          double v = 0.0;
          for (int j = 0; j < 128; ++j) v += (double)W0[(size_t)i * 128 + j] * (double)W2[(size_t)j * u + k];
          tl[(size_t)(1 + i) * 2 * UP + half * UP + k] = (float)(slope * v);
        }
      }
    }
    // B blocks: rows 0..u-1 = c Wt, rows u..2u-1 = c Ws (scaled layers), times SU, hi | lo
    for (int q = 0; q < 4; ++q) {
      uint16_t* c16 = reinterpret_cast<uint16_t*>(img.data() + (size_t)l * stc::L_BYTES + (size_t)q * stc::Q_BYTES);
      for (int kk = 0; kk < 32; ++kk) {
        const int j = 32 * q + kk;
        for (int n = 0; n < (scaled ? 2 * u : u); ++n) {
          const float v = (float)(c * (double)((n < u) ? Wt[(size_t)j * u + n] : Ws[(size_t)j * u + (n - u)]));
          float hi, lo;
          split_hi_lo(v, &hi, &lo);
          c16[kmajor_off16(stc::NO, n, kk)] = f32_to_f16(hi * stc::SU, &overflow);
          c16[kmajor_off16(stc::NO, n, 32 + kk)] = f32_to_f16(lo * stc::SU, &overflow);
        }
      }
    }
  }
  if (overflow) return HB_OK;
  double limit = 1e30;
  if (hsum_max > 0.0) limit = (60000.0 - hbias_max) / hsum_max;
  if (!(limit > 1.0)) return HB_OK;
  SmallTcCache* cch = new SmallTcCache();
  HB_CUDA(cudaMalloc(&cch->bimg, img.size()));
  HB_CUDA(cudaMemcpy(cch->bimg, img.data(), img.size(), cudaMemcpyHostToDevice));
  HB_CUDA(cudaMalloc(&cch->scal, scal.size() * sizeof(float)));
  HB_CUDA(cudaMemcpy(cch->scal, scal.data(), scal.size() * sizeof(float), cudaMemcpyHostToDevice));
  if (!f->tc_guard) {
    HB_CUDA(cudaMalloc(&f->tc_guard, sizeof(unsigned long long)));
    HB_CUDA(cudaMemset(f->tc_guard, 0, sizeof(unsigned long long)));
  }
  cch->smem = smem;
  cch->guard_limit = (float)(limit < 1e30 ? limit : 1e30);
  f->stc = cch;
  f->tc_ok = true;
  return HB_OK;
}

void realnvp_small_tc_release(Flow* f) {
  if (!f->stc) return;
  SmallTcCache* c = static_cast<SmallTcCache*>(f->stc);
  if (c->bimg) cudaFree(c->bimg);
  if (c->scal) cudaFree(c->scal);
  delete c;
  f->stc = nullptr;
}

template <int D>
static int stc_launch(const stc::Params& P, bool fold, int grid, size_t smem, const void* x, long long ld,
                      const void* lnpost, float* lnpred_out, const FoldParams& fp, cudaStream_t st) {
  if (fold) {
    auto k = stc::realnvp_small_tc_kernel<D, true>;
    HB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, stc::THREADS, smem, st>>>(P, x, ld, lnpost, lnpred_out, fp);
  } else {
    auto k = stc::realnvp_small_tc_kernel<D, false>;
    HB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, stc::THREADS, smem, st>>>(P, x, ld, lnpost, lnpred_out, fp);
  }
  return HB_OK;
}

int realnvp_small_tc_run(Flow* f, float T, const void* x, int x_dtype, long long ld, long long n,
                         const void* lnpost, int lp_dtype, float* lnpred_out, Accum* acc,
                         const int64_t* start, long long chain_lo, long long chain_hi, long long a,
                         long long b, unsigned long long* guard, cudaStream_t st) {
  if (!f->stc) {
    set_error("tcgen05 path not available for this flow");
    return HB_ERR_UNSUPPORTED;
  }
  if (n <= 0) return HB_OK;
  const SmallTcCache* c = static_cast<const SmallTcCache*>(f->stc);
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, f->device);
  const int NWG = stc::NT;
  long long nwg = (long long)sms * NWG;
  long long rpw = (n + nwg - 1) / nwg;
  rpw = ((rpw + stc::TILE - 1) / stc::TILE) * stc::TILE;
  nwg = (n + rpw - 1) / rpw;
  const int grid = (int)((nwg + NWG - 1) / NWG);
  const long long nfl = (long long)grid * NWG;
  stc::Params P;
  P.bimg = c->bimg;
  P.scal = c->scal;
  P.pre = f->desc.standardize ? f->dev_weights : nullptr;
  P.nlayers = f->nlayers;
  P.nscaled = f->desc.n_scaled;
  P.T = T;
  P.sum_log_amp = f->sum_log_amp;
  P.guard = guard;
  P.guard_limit = c->guard_limit;
  P.nrows = n;
  P.rows_per_wg = rpw;
  P.nbatches = (int)(rpw / stc::TILE);
  P.x_f64 = x_dtype == HB_F64;
  P.lp_f64 = lp_dtype == HB_F64;
  FoldParams fp{};
  int rc;
  if (acc) {
    rc = prepare_fold(acc, start, chain_lo, chain_hi, a, b, nfl, rpw, &fp, st);
    if (rc) return rc;
  }
  timing_begin(st);
  const bool fold = acc != nullptr;
  switch (f->desc.ndim) {
    case 2: rc = stc_launch<2>(P, fold, grid, c->smem, x, ld, lnpost, lnpred_out, fp, st); break;
    case 3: rc = stc_launch<3>(P, fold, grid, c->smem, x, ld, lnpost, lnpred_out, fp, st); break;
    case 4: rc = stc_launch<4>(P, fold, grid, c->smem, x, ld, lnpost, lnpred_out, fp, st); break;
    case 5: rc = stc_launch<5>(P, fold, grid, c->smem, x, ld, lnpost, lnpred_out, fp, st); break;
    case 6: rc = stc_launch<6>(P, fold, grid, c->smem, x, ld, lnpost, lnpred_out, fp, st); break;
    case 7: rc = stc_launch<7>(P, fold, grid, c->smem, x, ld, lnpost, lnpred_out, fp, st); break;
    default: rc = stc_launch<8>(P, fold, grid, c->smem, x, ld, lnpost, lnpred_out, fp, st); break;
  }
  timing_end(st);
  if (rc) return rc;
  count_launch();
  HB_CUDA(cudaGetLastError());
  if (acc) return launch_merge(acc, fp, chain_lo, nfl, st);
  return HB_OK;
}

}  // namespace hb
