// RealNVP log_prob (+ fused harmonic-mean fold) for small ndim (2..8) with BOTH contractions of the coupling MLP on the
// tcgen05 tensor cores (harmonic/flows.py:41-116,159-180; the Pima-Indian example and BASELINE cfg3 are ndim 5) -- the
// structure of hb_realnvp_tc.cu in miniature, with all weights resident in shared memory.
//
// A coupling layer conditions on d = round(ndim / 2) <= 4 features and has a fixed hidden width of 128 (flows.py:171):
//   GEMM1  H = y0 W0 + b0                      ONE MMA: the three products of the FP16 split and the bias fit a single K = 16
//                                              operand,  A = [hs(y0) 0.. | 1 | lo(y0) 0.. | 0 | hs(y0) 0.. | 1 | 0]  (shared memory),
//                                              B rows  = [w_hi      | b_hi | w_hi      | 0 | w_lo      | b_lo | 0],  N = 128
//   E1     r = relu(H) -> FP16 pair (hs, r - hs) in place of H: conversions whose clamp is a modifier (cvt.rz.relu /
//          cvt.rn.relu) and one mixed-precision FMA per element, 16 columns per step
//   GEMM2  O[128 x 16] = relu(H) (c W2), c = 0.99: per K = 32 quarter 6 MMAs (A from TMEM), issued while E1 works on the
//          next quarter
//   E2     leaky_relu(h) = s h + c relu(h): the linear part s (y0 W0 + b0) W2 + b2 = y0 V + v0 (a d x 2u matrix formed on
//          the host in float64, as in realnvp_small_kernel) is evaluated in registers while GEMM1 runs; shift, softplus-clip
//          scale, log-det as in the ndim-64 kernel's E2
// Weights travel times 2^5 (so that the lo halves stay normal halves): H accumulates 2^5, O 2^10 times its value.  The
// first version of this kernel computed H on the CUDA cores from per-hidden-unit records in shared memory: one sample per
// thread (TMEM lane = sample) leaves nothing to amortise the records over, and the loop was bound by shared-memory
// wavefronts (ncu: 75 % of the cycles; profiles/README.md).  Here a hidden unit costs the CUDA cores two instructions.
//
// Range: hs = fp16(2^5 relu(h)) needs h < 2047 (|h| <= max_j sum_i |W0[i][j]| max|y0| + max|b0|), and far outside the trained
// range (conditioning inputs beyond ~128 units; the limit is 64) the truncating accumulator costs accuracy.  Every epilogue thread
// tracks max |conditioning input| of its row over the layers, rows beyond the limit derived at flow creation are
// counted into the flow's guard counter and HB_PATH_AUTO answers by evaluating the call again on the CUDA cores
// (hb_api.cu, as for the ndim-64 kernel).  NaN inputs do not trip the guard: they propagate.
//
// One persistent CTA per SM: three epilogue warpgroups, each owning one 128-sample tile (TMEM lane = sample), and one
// auxiliary warpgroup whose warp t issues tile t's 25 MMAs per layer.  Shared memory: 12 KiB of FP16 blocks per layer, 4 KiB
// of A operand per tile.  Seven mbarriers per tile: A1FULL / HFULL around GEMM1, one per quarter of E1, OFULL.
// TMEM per tile (144 columns): P (128: H, then per 16-column chunk 8 columns of packed lo pairs + 8 of packed hs pairs) | O (16).
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
constexpr int NT = 3;
constexpr int THREADS = NT * 128 + 128;
constexpr int HID = 128;
constexpr int NO = 16;                  // GEMM2 N: 2u <= 8 outputs padded to 16
constexpr float SH = 32.0f, SU = 32.0f, SO_INV = 1.0f / (SH * SU);
constexpr int G1_BYTES = HID * 16 * 2;  // GEMM1 block [128 x 16 K']
constexpr int Q_BYTES = NO * 64 * 2;    // one K = 32 quarter of GEMM2: [16 x 64 K'], K' < 32 w_hi, K' >= 32 w_lo
constexpr int L_BYTES = G1_BYTES + 4 * Q_BYTES;  // per layer
constexpr int A1_BYTES = TILE * 16 * 2;  // GEMM1 A operand of a tile
constexpr uint32_t TM_P = 0, TM_O = HID, TM_TILE = HID + NO;
static_assert(NT * TM_TILE <= 512, "TMEM");
constexpr int B_A1FULL = 0, B_HFULL = NT, B_AQ = 2 * NT /* [quarter][tile] */, B_OFULL = 6 * NT, B_COUNT = 7 * NT;
// K' slots of the GEMM1 operand
constexpr int K_HS = 0, K_ONE = 4, K_LO = 5, K_HS2 = 10, K_ONE2 = 14;
__host__ __device__ constexpr int small_d(int D) { return (D == 2) ? 1 : (D == 3) ? 2 : (D == 4) ? 2 : (D == 5) ? 2 : (D == 6) ? 3 : 4; }
// float32 block of a layer: the linear part, (1 + d) rows of 2 x (4-padded u) floats
__host__ __device__ constexpr int scal_floats(int D) { return (1 + small_d(D)) * 2 * (((D - small_d(D)) + 3) & ~3); }

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

// E1 for one 16-column chunk of H (hb_realnvp_tc.cu, e1_chunk): wh = 8 words of packed hs = fp16(relu(H)) pairs rounded
// towards zero, wl = 8 words of packed fp16(relu(H) - hs) pairs
__device__ __forceinline__ void e1_chunk(const uint32_t (&v)[16], uint32_t (&wl)[8], uint32_t (&wh)[8]) {
#pragma unroll
  for (int jj = 0; jj < 8; ++jj) {
    const float h0 = __uint_as_float(v[2 * jj]), h1 = __uint_as_float(v[2 * jj + 1]);
    const uint32_t hs = pack_relu_rz_f16x2(h1, h0);
    float l0, l1;
    split_rest<false>(hs, h0, h1, l0, l1);
    wh[jj] = hs;
    wl[jj] = pack_relu_f16x2(l1, l0);
  }
}

template <int D, bool FOLD>
__global__ void __launch_bounds__(THREADS, 1)
realnvp_small_tc_kernel(Params P, const void* __restrict__ xv, long long ld, const void* __restrict__ lnpostv,
                        float* __restrict__ lnpred_out, FoldParams fp) {
  constexpr int d = small_d(D), u = D - d, UP = (u + 3) & ~3, SF = scal_floats(D);
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int nl = P.nlayers;
  const uint32_t img_bytes = (uint32_t)nl * L_BYTES;
  uint8_t* s_img = smem;
  uint8_t* s_a1 = smem + img_bytes;
  float* s_scal = reinterpret_cast<float*>(s_a1 + NT * A1_BYTES);
  uint8_t* s_tail = reinterpret_cast<uint8_t*>(s_scal) + (((uint32_t)nl * SF * 4 + 15) & ~15u);
  const uint32_t bar0 = smem_u32(s_tail);
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(s_tail + 8 * B_COUNT);
  double* scratch_all = reinterpret_cast<double*>(s_tail + 8 * B_COUNT + 16);

  if (threadIdx.x == 0) {
    for (int t = 0; t < NT; ++t) {
      mbar_init(BAR(B_A1FULL + t), 128);
      mbar_init(BAR(B_HFULL + t), 1);
      for (int q = 0; q < 4; ++q) mbar_init(BAR(B_AQ + NT * q + t), 128);
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
    // ===== MMA issuers: warp NT * 4 + t serves tile t; the whole warp runs the loop, one elected lane issues =====
    const int t = warp - NT * 4;
    if (t < NT) {
      const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0) + t * TM_TILE;
      const uint32_t ih = make_idesc_f16(HID), io = make_idesc_f16(NO);
      const uint32_t lbo2 = NO * 16;
      const uint64_t kstep = (uint64_t)((2 * lbo2) >> 4);
      const uint32_t img0 = smem_u32(s_img);
      const uint64_t adesc = make_desc(smem_u32(s_a1) + (uint32_t)t * A1_BYTES, TILE * 16, 128);
      uint32_t it = 0;
      for (int b = 0; b < P.nbatches; ++b) {
        for (int l = 0; l < nl; ++l, ++it) {
          const uint32_t lay = img0 + (uint32_t)l * L_BYTES;
          // ---- GEMM1: one MMA
          mbar_wait(BAR(B_A1FULL + t), it & 1);
          tc_fence_after();
          if (elect_one()) {
            mma_ss_f16(tb + TM_P, adesc, make_desc(lay, HID * 16, 128), ih, 0);
            tc_commit(BAR(B_HFULL + t));
          }
          __syncwarp();
          // ---- GEMM2: quarter q as soon as E1 has written it
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            mbar_wait(BAR(B_AQ + NT * q + t), it & 1);
            tc_fence_after();
            if (elect_one()) {
              const uint64_t bd = make_desc(lay + G1_BYTES + (uint32_t)q * Q_BYTES, lbo2, 128);
              const uint32_t a0 = tb + TM_P + 32 * q;
#pragma unroll
              for (int half = 0; half < 2; ++half) {
                const uint32_t a_lo = a0 + 16 * half, a_hs = a_lo + 8;
                // small terms first (lo . w_hi, hs . w_lo), then the main term hs . w_hi
                mma_ts_f16(tb + TM_O, a_lo, bd + (uint64_t)half * kstep, io, (q | half) ? 1u : 0u);
                mma_ts_f16(tb + TM_O, a_hs, bd + (uint64_t)(2 + half) * kstep, io, 1);
                mma_ts_f16(tb + TM_O, a_hs, bd + (uint64_t)half * kstep, io, 1);
              }
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
    const uint32_t tm_p = tmem_base + lane_base + wg * TM_TILE + TM_P;
    const uint32_t tm_o = tmem_base + lane_base + wg * TM_TILE + TM_O;
    uint8_t* a1 = s_a1 + wg * A1_BYTES;
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
        const float* tail = s_scal + l * SF;
        // ---- GEMM1's A operand: [hs(y0) | 1 | lo(y0) | 0 | hs(y0) | 1 | 0] as 16 halves
        {
          uint16_t k16[16];
#pragma unroll
          for (int k = 0; k < 16; ++k) k16[k] = 0;
          k16[K_ONE] = 0x3C00;
          k16[K_ONE2] = 0x3C00;
#pragma unroll
          for (int i = 0; i < d; ++i) {
            gmax = fmaxf(gmax, fabsf(y[i]));  // fmaxf drops NaN: NaN rows do not trip the guard
            const uint16_t hs = to_f16(y[i]);
            k16[K_HS + i] = hs;
            k16[K_HS2 + i] = hs;
            k16[K_LO + i] = to_f16(y[i] - from_f16(hs));
          }
          uint4 g0, g1;
          g0.x = k16[0] | ((uint32_t)k16[1] << 16); g0.y = k16[2] | ((uint32_t)k16[3] << 16);
          g0.z = k16[4] | ((uint32_t)k16[5] << 16); g0.w = k16[6] | ((uint32_t)k16[7] << 16);
          g1.x = k16[8] | ((uint32_t)k16[9] << 16); g1.y = k16[10] | ((uint32_t)k16[11] << 16);
          g1.z = k16[12] | ((uint32_t)k16[13] << 16); g1.w = k16[14] | ((uint32_t)k16[15] << 16);
          *reinterpret_cast<uint4*>(a1 + tid * 16) = g0;
          *reinterpret_cast<uint4*>(a1 + TILE * 16 + tid * 16) = g1;
        }
        fence_proxy_async();
        tc_fence_before();  // orders the previous layer's TMEM loads (O) before the next MMAs
        mbar_arrive(BAR(B_A1FULL + wg));
        // ---- the linear part y0 V + v0 while GEMM1 runs
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
        // ---- E1: relu(H) -> (hs, lo) in place, the next chunk's load in flight under the current chunk's arithmetic
        mbar_wait(BAR(B_HFULL + wg), it & 1);
        tc_fence_after();
        {
          uint32_t va[16], vb[16];
          tmem_ld16(tm_p, va);
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            uint32_t(&cur)[16] = (c & 1) ? vb : va;
            uint32_t(&nxt)[16] = (c & 1) ? va : vb;
            tmem_ld_wait16(cur);
            if (c < 7) tmem_ld16(tm_p + (c + 1) * 16, nxt);
            uint32_t wl[8], wh[8];
            e1_chunk(cur, wl, wh);
            tmem_st8(tm_p + c * 16, wl);
            tmem_st8(tm_p + c * 16 + 8, wh);
            if (c & 1) {  // a K = 32 quarter is complete
              tmem_st_wait();
              tc_fence_before();
              mbar_arrive(BAR(B_AQ + NT * (c >> 1) + wg));
            }
          }
        }
        // ---- E2
        mbar_wait(BAR(B_OFULL + wg), it & 1);
        tc_fence_after();
        uint32_t o[16];
        tmem_ld16(tm_o, o);
        tmem_ld_wait();
        float prod = 1.f;  // u <= 4 factors in [1e-3, 1e3] stay inside the float32 range
#pragma unroll
        for (int k = 0; k < u; ++k) {
          float y1 = y[d + k] - fmaf(__uint_as_float(o[k]), SO_INV, sh[k]);
          if (scaled) {
            const float a = fmaf(__uint_as_float(o[u + k]), SO_INV, sc[k]);
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
  const int SF = stc::scal_floats(D), UP = (u + 3) & ~3;
  const size_t img_bytes = (size_t)nl * stc::L_BYTES;
  const size_t smem = img_bytes + (size_t)stc::NT * stc::A1_BYTES + (((size_t)nl * SF * 4 + 15) & ~(size_t)15) +
                      8 * stc::B_COUNT + 16 + stc::NT * 12 * sizeof(double);
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
    uint8_t* lay = img.data() + (size_t)l * stc::L_BYTES;
    // GEMM1 block [128 x 16 K'] times SH: hi rows meet hs and lo of the inputs, lo rows meet hs; the bias meets the ones
    uint16_t* g16 = reinterpret_cast<uint16_t*>(lay);
    for (int j = 0; j < 128; ++j) {
      double sum = 0.0;
      for (int i = 0; i < d; ++i) {
        float hi, lo;
        split_hi_lo(W0[(size_t)i * 128 + j], &hi, &lo);
        const uint16_t h16 = f32_to_f16(hi * stc::SH, &overflow), l16 = f32_to_f16(lo * stc::SH, &overflow);
        g16[kmajor_off16(128, j, stc::K_HS + i)] = h16;
        g16[kmajor_off16(128, j, stc::K_LO + i)] = h16;
        g16[kmajor_off16(128, j, stc::K_HS2 + i)] = l16;
        sum += fabs((double)W0[(size_t)i * 128 + j]);
      }
      float hi, lo;
      split_hi_lo(b0[j], &hi, &lo);
      g16[kmajor_off16(128, j, stc::K_ONE)] = f32_to_f16(hi * stc::SH, &overflow);
      g16[kmajor_off16(128, j, stc::K_ONE2)] = f32_to_f16(lo * stc::SH, &overflow);
      if (sum > hsum_max) hsum_max = sum;
      if (fabs((double)b0[j]) > hbias_max) hbias_max = fabs((double)b0[j]);
    }
    // the linear part: row 0 = b2 + slope b0 W2, rows 1..d = slope W0 W2 (float64 sums, rounded once)
    float* tl = scal.data() + (size_t)l * SF;
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
    // GEMM2 blocks: rows 0..u-1 = c Wt, rows u..2u-1 = c Ws (scaled layers), times SU, hi | lo
    for (int q = 0; q < 4; ++q) {
      uint16_t* c16 = reinterpret_cast<uint16_t*>(lay + stc::G1_BYTES + (size_t)q * stc::Q_BYTES);
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
  // hs = fp16(SH relu(h)) and fp16(y0) must stay finite; and far outside the trained range the truncating accumulator
  // of the K = 128 contraction shows against a float32 FMA chain: on the cfg3 flow with inputs 4 - 8 sigma wide every row
  // that leaves max(1e-5, 3 x float32's own error) has a conditioning input beyond 128, none below (tools/
  // stc_guard_calib.py, profiles/r2_stc_guard_calib.txt).  Rows beyond HALF of that are counted and AUTO evaluates the
  // call on the CUDA cores (the ndim-64 kernel's policy)
  double limit = 64.0;
  if (const char* e = getenv("HB_STC_GUARD_LIMIT")) limit = atof(e);  // development aid (tools/stc_guard_calib.py)
  if (hsum_max > 0.0) {
    const double hl = (60000.0 / stc::SH - hbias_max) / hsum_max;
    if (hl < limit) limit = hl;
  }
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
  cch->guard_limit = (float)limit;
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
