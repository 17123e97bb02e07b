// RQSpline log_prob (+ fused harmonic-mean fold) with the conditioner's wide contraction on the tcgen05 tensor cores.
//
// Shape served: ndim 2 (harmonic's Rosenbrock example and BASELINE cfg2), two hidden layers h1 in {32, 64}, 8 bins.
// Per flow layer (harmonic/flows.py:234-288,335-438; App. A.2 of SURVEY.md) and sample, with v the conditioning feature
// and t the transformed one (the mask alternates):
//   a1 = tanh(Dense_0(where(mask, y, 0)))                      2 values,            CUDA cores
//   a2 = tanh(Dense_1(a1))                                     h1 values,           CUDA cores (2 FMAs + 6 instructions of a paired tanh)
//   P  = c + a2 . M      M = last MLP Dense x output Dense     h1 x 25 contraction  TENSOR CORES
//   y_t <- RationalQuadraticSpline(P)(y_t), then the scalar affine on both features                CUDA cores
// M is formed on the host in float64 (no activation between the two Denses, flows.py:357,434-438).  The contraction is
// 93 % of the flow's FLOPs (2 h1 25 of 2 (2 + 2 h1 + 25 h1) per layer) and the only part with a GEMM shape: 128 samples
// x h1 x 32.  On the CUDA cores it costs 25 of the ~40 instructions per hidden unit and sample (rqspline_fast_kernel);
// here it costs the FP16 split of a2 (2 instructions per hidden unit) and 3 h1 / 16 MMAs per 128 samples.
//
// Arithmetic: the three-product FP16 split of hb_realnvp_tc.cu -- a2 = hs + lo (hs = fp16(a2), lo = the
// exact remainder rounded to fp16), M = m_hi + m_lo (11 leading bits + rest), a2 M ~= hs m_hi + lo m_hi + hs m_lo with
// FP32 accumulation in TMEM; M travels times 2^5 so that m_lo stays a normal half, the factor leaves in the epilogue's
// FMA that adds the bias c.  |a2| <= 1, so there is no range to guard; weight blocks that leave the half range select
// the CUDA-core kernel at flow creation.
//
// One persistent CTA per SM: four epilogue warpgroups, each owning one 128-sample tile (TMEM lane = sample), and one
// auxiliary warpgroup whose warp t issues tile t's MMAs.  All layers' weights are resident in shared memory (8 KiB of
// FP16 blocks + 1.2 KiB of float32 scalars per layer at h1 = 64).  Per tile and layer two mbarriers: AFULL (the
// warpgroup has written a2's split into TMEM) and PFULL (tcgen05.commit: the 3 h1 / 16 MMAs have completed).  The
// chain of one tile is serial (a2 -> MMAs -> spline -> next layer); the four tiles of a CTA fill each other's gaps on
// the CUDA cores, which are the bound of this kernel.
// TMEM per tile (96 columns): A (h1 columns: per 16 hidden units 8 columns of packed lo pairs + 8 of packed hs pairs) |
// P (32 columns, FP32).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "hb_internal.h"
#include "hb_rqs.cuh"
#include "hb_tc_ptx.cuh"

namespace hb {

namespace rqtc {

using namespace tcx;

constexpr int D = 2;
constexpr int TILE = 128;
constexpr int NT = 4;                  // tiles (= epilogue warpgroups) in flight per CTA
constexpr int THREADS = NT * 128 + 128;
constexpr int NP = 32;                 // GEMM N: 25 spline parameters padded to 32
constexpr int MAXH = 64;
constexpr float SU = 32.0f, SU_INV = 1.0f / SU;
constexpr int Q_BYTES = NP * 64 * 2;   // one K = 32 quarter: [32 x 64 K'], K' < 32 m_hi, K' >= 32 m_lo
constexpr uint32_t TM_A = 0, TM_P = MAXH, TM_TILE = MAXH + NP;
static_assert(NT * TM_TILE <= 512, "TMEM");
// float32 scalars of a layer: w1[2] b1[2] | scale[2] shift[2] | per hidden unit {W2[0][j], W2[1][j], b2[j], 0} | c[28]
// (w1, b1, W2, b2 times 2 log2(e): tanh_pair_scaled)
__host__ __device__ constexpr int scal_floats(int h1) { return 8 + 4 * h1 + RQF_NP; }
constexpr int B_AFULL = 0, B_PFULL = NT, B_COUNT = 2 * NT;

constexpr double TANH_ARG = 2.8853900817779268;  // 2 log2(e): both tanh Denses travel times this factor

// tanh(x) from s = 2 log2(e) x (the factor sits in the Dense's weights and bias, host side): with e = 2^-|s| in (0, 1],
// tanh|x| = (1 - e) / (1 + e) = 2 / (1 + e) - 1.  The kernel is bound by the MUFU pipe (ncu: 73 % busy at 2 MUFU per
// tanh), so two hidden units share ONE reciprocal: R = 1 / ((1 + ea)(1 + eb)) -- the product stays in (1, 4] --,
// 1 / (1 + ea) = (1 + eb) R.  Per pair 2 MUFU ex2 + 1 MUFU rcp + 9 other instructions.  Absolute error <= ~3e-7 (2^-22
// relative on e, the reciprocal's ulp and two more roundings, halved by the quotient's slope), which is what the
// consumers of a tanh output see (hb_rqs.cuh, tanh_fast).  s = +-inf gives +-1, NaN stays NaN.
__device__ __forceinline__ void tanh_pair_scaled(float sa, float sb, float& ta, float& tb) {
  const float da = ex2_approx(-fabsf(sa)) + 1.0f, db = ex2_approx(-fabsf(sb)) + 1.0f;
  const float R = rcp_approx(da * db);
  ta = copysignf(fmaf(2.0f, db * R, -1.0f), sa);
  tb = copysignf(fmaf(2.0f, da * R, -1.0f), sb);
}

struct Params {
  const uint8_t* bimg;   // FP16 blocks: [layer][quarter] of Q_BYTES
  const float* scal;     // [layer][scal_floats(h1)]
  const float* pre;      // pre_offset[2], pre_amp[2] or nullptr
  int nlayers, h1;
  float lo, hi, T, sum_log_amp;
  long long nrows;
  long long rows_per_wg;  // contiguous rows per epilogue warpgroup (multiple of TILE)
  int nbatches;           // tiles per warpgroup
};

template <typename XT, typename LT, bool FOLD>
__global__ void __launch_bounds__(THREADS, 1)
rqspline_tc_kernel(Params P, const XT* __restrict__ x, long long ld, const LT* __restrict__ lnpost,
                   float* __restrict__ lnpred_out, FoldParams fp) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int nl = P.nlayers, h1 = P.h1, nq = h1 >> 5;
  const int SF = scal_floats(h1);
  // shared memory carve-up
  const uint32_t img_bytes = (uint32_t)nl * nq * Q_BYTES;
  uint8_t* s_img = smem;
  float* s_scal = reinterpret_cast<float*>(smem + img_bytes);
  uint8_t* s_tail = smem + img_bytes + (((uint32_t)nl * SF * 4 + 15) & ~15u);
  const uint32_t bar0 = smem_u32(s_tail);
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(s_tail + 8 * B_COUNT);
  double* scratch_all = reinterpret_cast<double*>(s_tail + 8 * B_COUNT + 16);

  // ---- one-time setup: barriers, resident weights, TMEM --------------------------------------
  if (threadIdx.x == 0) {
    for (int t = 0; t < NT; ++t) {
      mbar_init(BAR(B_AFULL + t), 128);
      mbar_init(BAR(B_PFULL + t), 1);
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
      const uint32_t idesc = make_idesc_f16(NP);
      const uint32_t lbo = NP * 16;
      const uint64_t kstep = (uint64_t)((2 * lbo) >> 4);
      const uint32_t img0 = smem_u32(s_img);
      uint32_t it = 0;
      for (int b = 0; b < P.nbatches; ++b) {
        for (int li = nl - 1; li >= 0; --li, ++it) {
          mbar_wait(BAR(B_AFULL + t), it & 1);
          tc_fence_after();
          if (elect_one()) {
            for (int q = 0; q < nq; ++q) {
              const uint64_t bd = make_desc(img0 + (uint32_t)(li * nq + q) * Q_BYTES, lbo, 128);
              const uint32_t a0 = tb + TM_A + 32 * q;
#pragma unroll
              for (int half = 0; half < 2; ++half) {
                const uint32_t a_lo = a0 + 16 * half, a_hs = a_lo + 8;
                // small terms first (lo . m_hi, hs . m_lo), then the main term hs . m_hi
                mma_ts_f16(tb + TM_P, a_lo, bd + (uint64_t)half * kstep, idesc, (q | half) ? 1u : 0u);
                mma_ts_f16(tb + TM_P, a_hs, bd + (uint64_t)(2 + half) * kstep, idesc, 1);
                mma_ts_f16(tb + TM_P, a_hs, bd + (uint64_t)half * kstep, idesc, 1);
              }
            }
            tc_commit(BAR(B_PFULL + t));
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ===== epilogue warpgroups: one sample per thread =====
    const int wg = warp >> 2;
    const int tid = threadIdx.x & 127;
    const uint32_t lane_base = ((uint32_t)((warp & 3) * 32)) << 16;
    const uint32_t tm_a = tmem_base + lane_base + wg * TM_TILE + TM_A;
    const uint32_t tm_p = tmem_base + lane_base + wg * TM_TILE + TM_P;
    double* scratch = scratch_all + wg * 12;
    const int flusher = blockIdx.x * NT + wg;
    const long long row_lo = (long long)flusher * P.rows_per_wg;
    long long row_hi = row_lo + P.rows_per_wg;
    if (row_hi > P.nrows) row_hi = P.nrows;
    const bool had_rows = row_lo < P.nrows;
    Fold<128> fold(fp, flusher, tid, scratch, (FOLD && had_rows) ? row_lo : 0, 1 + wg);  // named barriers 1..NT
    const float invT = 1.f / P.T;
    const float base_const = -0.5f * (float)D * 1.8378770664093453f - 0.5f * (float)D * logf(P.T);
    float pre_s[2] = {1.f, 1.f}, pre_o[2] = {0.f, 0.f};
    if (P.pre) {
#pragma unroll
      for (int i = 0; i < D; ++i) {
        pre_o[i] = P.pre[i];
        pre_s[i] = P.pre[D + i];
      }
    }
    uint32_t it = 0;
    for (int b = 0; b < P.nbatches; ++b) {
      const long long t0 = row_lo + (long long)b * TILE;
      const long long r = t0 + tid;
      const bool live = r < row_hi;
      const long long rr = live ? r : (had_rows ? row_lo : 0);
      float y[D];
#pragma unroll
      for (int i = 0; i < D; ++i) {
        float v = (float)x[rr * ld + i];
        if (P.pre) v = (v - pre_o[i]) / pre_s[i];
        y[i] = v;
      }
      float lp = 0.f;
      if (FOLD) lp = live ? (float)lnpost[r] : 0.f;
      float ldet = 0.f;
      for (int li = nl - 1; li >= 0; --li, ++it) {  // Inverse(Chain).inverse: last layer first
        const float* S = s_scal + li * SF;
        const int par = li & 1;  // even layers: feature 1 conditions, feature 0 is transformed; odd layers the reverse
        const float yv = par ? y[0] : y[1];
        const float4 s0 = *reinterpret_cast<const float4*>(S);  // w1[v][0], w1[v][1], b1[0], b1[1]
        float a10, a11;
        tanh_pair_scaled(fmaf(yv, s0.x, s0.z), fmaf(yv, s0.y, s0.w), a10, a11);
        // ---- a2 = tanh(Dense_1(a1)) and its FP16 split, 16 hidden units per TMEM chunk
        const float4* W2 = reinterpret_cast<const float4*>(S + 8);
        for (int c = 0; c < (h1 >> 4); ++c) {
          uint32_t wl[8], wh[8];
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) {
            const float4 wa = W2[16 * c + 2 * jj], wb = W2[16 * c + 2 * jj + 1];
            float ta, tb2;
            tanh_pair_scaled(fmaf(a10, wa.x, fmaf(a11, wa.y, wa.z)), fmaf(a10, wb.x, fmaf(a11, wb.y, wb.z)), ta, tb2);
            split_x_pair(ta, tb2, wh[jj], wl[jj]);
          }
          tmem_st8(tm_a + 16 * c, wl);
          tmem_st8(tm_a + 16 * c + 8, wh);
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(BAR(B_AFULL + wg));
        // ---- P = c + a2 . M
        mbar_wait(BAR(B_PFULL + wg), it & 1);
        tc_fence_after();
        uint32_t v[32];
        tmem_ld32(tm_p, v);
        tmem_ld_wait();
        tc_fence_before();  // orders this load before the next layer's MMAs (issued after the next AFULL)
        float p[RQF_NP];
        const float* cb = S + 8 + 4 * h1;
#pragma unroll
        for (int k4 = 0; k4 < RQF_NP; k4 += 4) {
          const float4 c4 = *reinterpret_cast<const float4*>(cb + k4);
          p[k4] = fmaf(__uint_as_float(v[k4]), SU_INV, c4.x);
          p[k4 + 1] = fmaf(__uint_as_float(v[k4 + 1]), SU_INV, c4.y);
          p[k4 + 2] = fmaf(__uint_as_float(v[k4 + 2]), SU_INV, c4.z);
          p[k4 + 3] = fmaf(__uint_as_float(v[k4 + 3]), SU_INV, c4.w);
        }
        // ---- spline on the transformed feature, then the scalar affine on both
        float yo, lo_;
        rqs_forward_mufu(par ? y[1] : y[0], p, P.lo, P.hi, yo, lo_);
        if (par) y[1] = yo; else y[0] = yo;
        ldet += lo_;
        const float4 s1 = *reinterpret_cast<const float4*>(S + 4);  // scale[2], shift[2]
        y[0] = fmaf(y[0], s1.x, s1.z);
        y[1] = fmaf(y[1], s1.y, s1.w);
        ldet += __logf(fabsf(s1.x)) + __logf(fabsf(s1.y));
      }
      const float ss = fmaf(y[0], y[0], y[1] * y[1]);
      const float lnphi = -0.5f * ss * invT + base_const + ldet - P.sum_log_amp;
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
  }

  // ---- teardown ---------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (warp == NT * 4) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

}  // namespace rqtc

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
using namespace tcx;

struct RqTcCache {
  uint8_t* bimg = nullptr;
  float* scal = nullptr;
  size_t smem = 0;
};

bool rqspline_tc_eligible(const hb_flow_desc& ds) {
  return ds.kind == HB_FLOW_RQSPLINE && ds.ndim == rqtc::D && ds.n_hidden == 2 && ds.n_bins == 8 &&
         (ds.hidden[0] == 32 || ds.hidden[0] == 64) && ds.hidden[1] >= 1 && ds.n_layers >= 1;
}

// Builds the resident weight image; leaves f->rqtc null (CUDA-core kernel serves the flow) when the shape is not
// served, a block leaves the half range, or the image does not fit shared memory.
int rqspline_tc_prepare(Flow* f) {
  const hb_flow_desc& ds = f->desc;
  if (!rqspline_tc_eligible(ds)) return HB_OK;
  const int D = ds.ndim, K = ds.n_bins, npar = 3 * K + 1;
  const int h1 = ds.hidden[0], h2 = ds.hidden[1], nq = h1 / 32, SF = rqtc::scal_floats(h1);
  const size_t img_bytes = (size_t)ds.n_layers * nq * rqtc::Q_BYTES;
  const size_t smem = img_bytes + (((size_t)ds.n_layers * SF * 4 + 15) & ~(size_t)15) + 8 * rqtc::B_COUNT + 16 +
                      rqtc::NT * 12 * sizeof(double);
  if (smem > 200 * 1024) return HB_OK;
  std::vector<uint8_t> img(img_bytes, 0);
  std::vector<float> scal((size_t)ds.n_layers * SF, 0.f);
  const float* w = f->host_weights.data() + 2 * (size_t)D;
  bool overflow = false;
  for (int i = 0; i < ds.n_layers; ++i) {
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
    const int par = i & 1;
    const int vis = par ? 0 : 1, tr = par ? 1 : 0;  // mask = (j % 2 == 1) on even layers, negated on odd ones
    float* S = scal.data() + (size_t)i * SF;
    const double ta = rqtc::TANH_ARG;  // tanh_pair_scaled takes 2 log2(e) x
    S[0] = (float)(ta * (double)W1[(size_t)vis * D + 0]);
    S[1] = (float)(ta * (double)W1[(size_t)vis * D + 1]);
    S[2] = (float)(ta * (double)b1[0]);
    S[3] = (float)(ta * (double)b1[1]);
    S[4] = scales[0];
    S[5] = scales[1];
    S[6] = shifts[0];
    S[7] = shifts[1];
    for (int j = 0; j < h1; ++j) {
      S[8 + 4 * j + 0] = (float)(ta * (double)W2[(size_t)0 * h1 + j]);
      S[8 + 4 * j + 1] = (float)(ta * (double)W2[(size_t)1 * h1 + j]);
      S[8 + 4 * j + 2] = (float)(ta * (double)b2[j]);
    }
    float* cb = S + 8 + 4 * h1;
    for (int c = 0; c < npar; ++c) {
      double acc = bout[(size_t)tr * npar + c];
      for (int q = 0; q < h2; ++q) acc += (double)b3[q] * (double)Wout[(size_t)q * D * npar + (size_t)tr * npar + c];
      cb[c] = (float)acc;
    }
    // M[r][c] = sum_q W3[r][q] Wout[q][tr][c] in float64, rounded to float32, split, times SU
    for (int qd = 0; qd < nq; ++qd) {
      uint16_t* c16 = reinterpret_cast<uint16_t*>(img.data() + ((size_t)i * nq + qd) * rqtc::Q_BYTES);
      for (int kk = 0; kk < 32; ++kk) {
        const int r = 32 * qd + kk;
        for (int c = 0; c < npar; ++c) {
          double acc = 0.0;
          for (int q = 0; q < h2; ++q)
            acc += (double)W3[(size_t)r * h2 + q] * (double)Wout[(size_t)q * D * npar + (size_t)tr * npar + c];
          float hi, lo;
          split_hi_lo((float)acc, &hi, &lo);
          c16[kmajor_off16(rqtc::NP, c, kk)] = f32_to_f16(hi * rqtc::SU, &overflow);
          c16[kmajor_off16(rqtc::NP, c, 32 + kk)] = f32_to_f16(lo * rqtc::SU, &overflow);
        }
      }
    }
  }
  if (overflow) return HB_OK;
  RqTcCache* c = new RqTcCache();
  HB_CUDA(cudaMalloc(&c->bimg, img.size()));
  HB_CUDA(cudaMemcpy(c->bimg, img.data(), img.size(), cudaMemcpyHostToDevice));
  HB_CUDA(cudaMalloc(&c->scal, scal.size() * sizeof(float)));
  HB_CUDA(cudaMemcpy(c->scal, scal.data(), scal.size() * sizeof(float), cudaMemcpyHostToDevice));
  c->smem = smem;
  f->rqtc = c;
  f->tc_ok = true;
  return HB_OK;
}

void rqspline_tc_release(Flow* f) {
  if (!f->rqtc) return;
  RqTcCache* c = static_cast<RqTcCache*>(f->rqtc);
  if (c->bimg) cudaFree(c->bimg);
  if (c->scal) cudaFree(c->scal);
  delete c;
  f->rqtc = nullptr;
}

template <typename XT, typename LT, bool FOLD>
static int rqtc_launch(const rqtc::Params& P, int grid, size_t smem, const void* x, long long ld, const void* lnpost,
                       float* lnpred_out, const FoldParams& fp, cudaStream_t st) {
  auto kern = rqtc::rqspline_tc_kernel<XT, LT, FOLD>;
  HB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, rqtc::THREADS, smem, st>>>(P, (const XT*)x, ld, (const LT*)lnpost, lnpred_out, fp);
  return HB_OK;
}

int rqspline_tc_run(Flow* f, float T, const void* x, int x_dtype, long long ld, long long n,
                    const void* lnpost, int lp_dtype, float* lnpred_out, Accum* acc,
                    const int64_t* start, long long chain_lo, long long chain_hi, long long a,
                    long long b, cudaStream_t st) {
  if (!f->rqtc) {
    set_error("tcgen05 path not available for this flow");
    return HB_ERR_UNSUPPORTED;
  }
  if (n <= 0) return HB_OK;
  const RqTcCache* c = static_cast<const RqTcCache*>(f->rqtc);
  const hb_flow_desc& ds = f->desc;
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, f->device);
  const int NWG = rqtc::NT;
  long long nwg = (long long)sms * NWG;
  long long rpw = (n + nwg - 1) / nwg;
  rpw = ((rpw + rqtc::TILE - 1) / rqtc::TILE) * rqtc::TILE;
  nwg = (n + rpw - 1) / rpw;
  const int grid = (int)((nwg + NWG - 1) / NWG);
  const long long nfl = (long long)grid * NWG;
  rqtc::Params P;
  P.bimg = c->bimg;
  P.scal = c->scal;
  P.pre = ds.standardize ? f->dev_weights : nullptr;
  P.nlayers = ds.n_layers;
  P.h1 = ds.hidden[0];
  P.lo = ds.range_min;
  P.hi = ds.range_max;
  P.T = T;
  P.sum_log_amp = f->sum_log_amp;
  P.nrows = n;
  P.rows_per_wg = rpw;
  P.nbatches = (int)(rpw / rqtc::TILE);
  FoldParams fp{};
  int rc;
  if (acc) {
    rc = prepare_fold(acc, start, chain_lo, chain_hi, a, b, nfl, rpw, &fp, st);
    if (rc) return rc;
  }
  timing_begin(st);
  if (acc) {
    if (x_dtype == HB_F32 && lp_dtype == HB_F32) rc = rqtc_launch<float, float, true>(P, grid, c->smem, x, ld, lnpost, lnpred_out, fp, st);
    else if (x_dtype == HB_F32) rc = rqtc_launch<float, double, true>(P, grid, c->smem, x, ld, lnpost, lnpred_out, fp, st);
    else if (lp_dtype == HB_F32) rc = rqtc_launch<double, float, true>(P, grid, c->smem, x, ld, lnpost, lnpred_out, fp, st);
    else rc = rqtc_launch<double, double, true>(P, grid, c->smem, x, ld, lnpost, lnpred_out, fp, st);
  } else {
    if (x_dtype == HB_F32) rc = rqtc_launch<float, float, false>(P, grid, c->smem, x, ld, nullptr, lnpred_out, fp, st);
    else rc = rqtc_launch<double, float, false>(P, grid, c->smem, x, ld, nullptr, lnpred_out, fp, st);
  }
  timing_end(st);
  if (rc) return rc;
  count_launch();
  HB_CUDA(cudaGetLastError());
  if (acc) return launch_merge(acc, fp, chain_lo, nfl, st);
  return HB_OK;
}

}  // namespace hb
