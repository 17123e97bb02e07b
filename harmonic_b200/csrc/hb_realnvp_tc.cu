// tcgen05 RealNVP kernel (sm_100a): fused flow evaluation + per-chain fold.
//
// Replaces harmonic/flows.py:41-116,159-180 (tfp-jax RealNVP log_prob) + harmonic/model.py:222-235
// + harmonic/evidence.py:278-352 for the shapes the tensor path supports (ndim = 64 today).
//
// Mapping.  One persistent CTA per SM, 320 threads:
//   warps 0-3 / 4-7 : two "epilogue" warpgroups; each owns one 128-sample tile at a time, one sample
//                     per thread (TMEM lane = sample), y[ndim] in registers through all layers
//   warp 8          : weight producer: streams each layer's pre-split weight image (hi/lo TF32,
//                     UMMA canonical no-swizzle K-major layout, built once on the host) from L2 into
//                     a 4-slot shared-memory ring with cp.async.bulk + mbarrier
//   warp 9          : TMEM allocator + single-thread tcgen05.mma issuer
// Per coupling layer and tile:   GEMM1  H[128x128]  = y0[128xd]  . W0[dx128]      (A, B from smem)
//                                E1     h = leaky_relu(H + b0), split hi/lo, back to TMEM
//                                GEMM2  O[128xN2]   = h[128x128] . [Wt|Ws]        (A from TMEM)
//                                E2     shift/scale epilogue on y1, log-det
// Each GEMM is 3xTF32: a.w ~= a_lo.w_hi + a_hi.w_lo + a_hi.w_hi with FP32 accumulation in TMEM
// (hi = top 19 bits, lo = exact remainder), so per-sample ln phi keeps ~FP32 accuracy.
// TMEM (512 columns), per tile t at column 256 t:  P (128: H, then h_hi in place) | Q (64: h_lo of
// the hidden half in flight) | O (64: GEMM2 accumulators).  GEMM2 runs as two K-halves so that only
// half of h_lo has to live in TMEM; the epilogue of the second half overlaps the first half's MMAs.
// The MMA thread issues whichever tile's next GEMM is ready (layers stay in lock-step so that the
// weight ring is consumed in order), so one tile's epilogue overlaps the other tile's tensor work.
#include <math.h>
#include <string.h>

#include <vector>

#include "hb_internal.h"

namespace hb {

namespace tc {

constexpr int D = 64;
constexpr int DH = 32;        // d = conditioning width (round(D/2))
constexpr int U = 32;         // transformed width
constexpr int HID = 128;
constexpr int TILE = 128;
constexpr int NWG = 2;        // epilogue warpgroups = tiles in flight
constexpr int THREADS = NWG * 128 + 64;
constexpr int SLOT_BYTES = 32768;
constexpr int NSLOTS = 4;
constexpr int MAX_TC_LAYERS = 16;

// shared memory carve-up (bytes)
constexpr int SM_RING = 0;
constexpr int SM_A1 = SM_RING + NSLOTS * SLOT_BYTES;            // per WG: hi 16 KB | lo 16 KB
constexpr int SM_BIAS = SM_A1 + NWG * 32768;                    // MAX_TC_LAYERS * 192 floats
constexpr int SM_PRE = SM_BIAS + MAX_TC_LAYERS * 192 * 4;       // standardisation: inv_amp[D], -offset*inv_amp[D]
constexpr int SM_BAR = SM_PRE + 2 * D * 4;                      // mbarriers
constexpr int SM_MISC = SM_BAR + 256;                           // tmem ptr, fold scratch
constexpr int SM_TOTAL = SM_MISC + 256;

// barrier indices (8 bytes each)
enum {
  B_WFULL = 0, B_WEMPTY = 4, B_A1FULL = 8, B_HFULL = 10, B_A2A = 12, B_QFREE = 14, B_A2B = 16, B_OFULL = 18,
  B_COUNT = 20
};

// TMEM columns (per tile: + 256 * tile)
constexpr uint32_t TM_P = 0, TM_Q = 128, TM_O = 192, TM_TILE = 256;

struct Params {
  const uint8_t* image;   // weight image: chunks of SLOT_BYTES in consumption order
  const float* bias;      // per layer 192 floats: b0[128], b2[64]
  const float* pre;       // pre_offset[D], pre_amp[D] or nullptr
  int nlayers;
  int nscaled;
  float T, sum_log_amp;
  long long nrows;
  long long rows_per_wg;  // contiguous rows per epilogue warpgroup (multiple of TILE)
  int nbatches;           // tiles per warpgroup
};

// ---- PTX wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}\n" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem]
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
#define HB_R32(v, o)                                                                                  \
  "=r"(v[o + 0]), "=r"(v[o + 1]), "=r"(v[o + 2]), "=r"(v[o + 3]), "=r"(v[o + 4]), "=r"(v[o + 5]),     \
      "=r"(v[o + 6]), "=r"(v[o + 7]), "=r"(v[o + 8]), "=r"(v[o + 9]), "=r"(v[o + 10]), "=r"(v[o + 11]), \
      "=r"(v[o + 12]), "=r"(v[o + 13]), "=r"(v[o + 14]), "=r"(v[o + 15]), "=r"(v[o + 16]),            \
      "=r"(v[o + 17]), "=r"(v[o + 18]), "=r"(v[o + 19]), "=r"(v[o + 20]), "=r"(v[o + 21]),            \
      "=r"(v[o + 22]), "=r"(v[o + 23]), "=r"(v[o + 24]), "=r"(v[o + 25]), "=r"(v[o + 26]),            \
      "=r"(v[o + 27]), "=r"(v[o + 28]), "=r"(v[o + 29]), "=r"(v[o + 30]), "=r"(v[o + 31])
#define HB_W32(v, o)                                                                                \
  "r"(v[o + 0]), "r"(v[o + 1]), "r"(v[o + 2]), "r"(v[o + 3]), "r"(v[o + 4]), "r"(v[o + 5]),         \
      "r"(v[o + 6]), "r"(v[o + 7]), "r"(v[o + 8]), "r"(v[o + 9]), "r"(v[o + 10]), "r"(v[o + 11]),   \
      "r"(v[o + 12]), "r"(v[o + 13]), "r"(v[o + 14]), "r"(v[o + 15]), "r"(v[o + 16]), "r"(v[o + 17]), \
      "r"(v[o + 18]), "r"(v[o + 19]), "r"(v[o + 20]), "r"(v[o + 21]), "r"(v[o + 22]), "r"(v[o + 23]), \
      "r"(v[o + 24]), "r"(v[o + 25]), "r"(v[o + 26]), "r"(v[o + 27]), "r"(v[o + 28]), "r"(v[o + 29]), \
      "r"(v[o + 30]), "r"(v[o + 31])
#define HB_LIST32                                                                                     \
  "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, " \
  "%21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}"
#define HB_LIST32_FROM1                                                                              \
  "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, " \
  "%22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32}"

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 " HB_LIST32 ", [%32];"
               : HB_R32(v, 0)
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], " HB_LIST32_FROM1 ";" ::"r"(taddr), HB_W32(v, 0)
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// UMMA shared-memory descriptor, K-major, no swizzle: element (r, k) of a [rows x K] tf32 operand
// lives at (r/8)*SBO + (r%8)*16 + (k/4)*LBO + (k%4)*4 bytes.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) |
         (1ull << 46);
}
__device__ __forceinline__ uint32_t make_idesc(int n) {
  // c=F32 (1<<4), a=b=TF32 (2<<7, 2<<10), K-major both, N>>3 at bit 17, M>>4 at bit 24
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

__device__ __forceinline__ float softplus_clip(float a) {
  float sp = fmaxf(a, 0.f) + log1pf(expf(-fabsf(a)));
  return fminf(fmaxf(sp, 1e-3f), 1e3f);
}

template <typename XT>
__device__ __forceinline__ void load_row(const XT* __restrict__ x, long long ld, long long r, float (&y)[D]);

template <>
__device__ __forceinline__ void load_row<float>(const float* __restrict__ x, long long ld, long long r,
                                                float (&y)[D]) {
  const float* p = x + r * ld;
  if ((((uintptr_t)p) & 15) == 0) {
#pragma unroll
    for (int i = 0; i < D / 4; ++i) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(p) + i);
      y[4 * i] = v.x; y[4 * i + 1] = v.y; y[4 * i + 2] = v.z; y[4 * i + 3] = v.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < D; ++i) y[i] = __ldg(p + i);
  }
}
template <>
__device__ __forceinline__ void load_row<double>(const double* __restrict__ x, long long ld, long long r,
                                                 float (&y)[D]) {
  const double* p = x + r * ld;
  if ((((uintptr_t)p) & 15) == 0) {
#pragma unroll
    for (int i = 0; i < D / 2; ++i) {
      const double2 v = __ldg(reinterpret_cast<const double2*>(p) + i);
      y[2 * i] = (float)v.x; y[2 * i + 1] = (float)v.y;
    }
  } else {
#pragma unroll
    for (int i = 0; i < D; ++i) y[i] = (float)__ldg(p + i);
  }
}

__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// clip(softplus(a), 1e-3, 1e3) with MUFU ex2/lg2 (harmonic/flows.py:177).  t = exp(-|a|) in (0, 1];
// log1p(t) by a 4-term series below 0.06 (rel. err < 3e-6) and ln2 * lg2(1 + t) above.
__device__ __forceinline__ float softplus_clip_fast(float a) {
  const float t = ex2_approx(-1.4426950408889634f * fabsf(a));
  const float series = t * fmaf(t, fmaf(t, fmaf(t, -0.25f, 0.33333334f), -0.5f), 1.0f);
  const float via_log = 0.6931471805599453f * lg2_approx(1.0f + t);
  const float sp = fmaxf(a, 0.f) + ((t < 0.06f) ? series : via_log);
  return fminf(fmaxf(sp, 1e-3f), 1e3f);  // NaN in -> NaN out is handled by the caller's y1 path
}

// E1 for one 32-column chunk: h = leaky_relu(H + b0) -> hi (top 19 bits) / lo (exact remainder)
__device__ __forceinline__ void e1_chunk(uint32_t (&v)[32], uint32_t (&w)[32], const float* b0c) {
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    float h = __uint_as_float(v[j]) + b0c[j];
    h = fmaxf(h, 0.01f * h);
    const uint32_t hb = __float_as_uint(h) & 0xffffe000u;
    v[j] = hb;
    w[j] = __float_as_uint(h - __uint_as_float(hb));
  }
}

template <typename XT, typename LT, bool FOLD>
__global__ void __launch_bounds__(THREADS, 1)
realnvp_tc_kernel(Params P, const XT* __restrict__ x, long long ld, const LT* __restrict__ lnpost,
                  float* __restrict__ lnpred_out, FoldParams fp) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar0 = sbase + SM_BAR;
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(smem + SM_MISC);
  float* bias_s = reinterpret_cast<float*>(smem + SM_BIAS);
  float* pre_s = reinterpret_cast<float*>(smem + SM_PRE);
  const int nl = P.nlayers;

  // ---- one-time setup -------------------------------------------------------------------
  if (threadIdx.x == 0) {
    for (int i = 0; i < NSLOTS; ++i) {
      mbar_init(BAR(B_WFULL + i), 1);
      mbar_init(BAR(B_WEMPTY + i), 1);
    }
    for (int t = 0; t < NWG; ++t) {
      mbar_init(BAR(B_A1FULL + t), 128);
      mbar_init(BAR(B_HFULL + t), 1);
      mbar_init(BAR(B_A2A + t), 128);
      mbar_init(BAR(B_QFREE + t), 1);
      mbar_init(BAR(B_A2B + t), 128);
      mbar_init(BAR(B_OFULL + t), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < nl * 192; i += THREADS) bias_s[i] = P.bias[i];
  for (int i = threadIdx.x; i < D; i += THREADS) {
    const float inv = P.pre ? 1.f / P.pre[D + i] : 1.f;
    pre_s[i] = inv;
    pre_s[D + i] = P.pre ? -P.pre[i] * inv : 0.f;
  }
  if (warp == 9) {
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

  if (warp == 8) {
    // ===== weight producer =====
    if (lane == 0) {
      uint32_t g = 0;
      for (int b = 0; b < P.nbatches; ++b) {
        const uint8_t* src = P.image;
        for (int l = 0; l < nl; ++l) {
          const int nch = (l < P.nscaled) ? 3 : 2;
          for (int c = 0; c < nch; ++c, ++g) {
            const uint32_t slot = g & (NSLOTS - 1), use = g / NSLOTS;
            mbar_wait(BAR(B_WEMPTY + slot), (use & 1) ^ 1);
            mbar_expect_tx(BAR(B_WFULL + slot), SLOT_BYTES);
            bulk_g2s(sbase + SM_RING + slot * SLOT_BYTES, src, SLOT_BYTES, BAR(B_WFULL + slot));
            src += SLOT_BYTES;
          }
        }
      }
    }
  } else if (warp == 9) {
    // ===== MMA issuer: the whole warp runs the (warp-uniform) scheduling loop so that descriptors
    // stay in uniform registers; one elected lane issues.  Within a layer it serves whichever
    // tile is ready.
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t id128 = make_idesc(128), id64 = make_idesc(64), id32 = make_idesc(32);
    uint32_t g = 0, it = 0;
    for (int b = 0; b < P.nbatches; ++b) {
      for (int l = 0; l < nl; ++l, ++it) {
        const bool scaled = l < P.nscaled;
        const uint32_t par = it & 1;
        const uint32_t s0 = g & (NSLOTS - 1);
        const uint32_t s1 = (g + 1) & (NSLOTS - 1), s2 = (g + 2) & (NSLOTS - 1);
        const uint32_t w0 = sbase + SM_RING + s0 * SLOT_BYTES;  // hi [128x32] 16 KB | lo 16 KB
        const uint32_t lbo2 = (scaled ? 64u : 32u) * 16u;
        const uint32_t w_hi = sbase + SM_RING + s1 * SLOT_BYTES;
        const uint32_t w_lo = scaled ? sbase + SM_RING + s2 * SLOT_BYTES : w_hi + 16384;
        const uint32_t id2 = scaled ? id64 : id32;
        const uint64_t wd_hi = make_desc(w0, 2048, 128), wd_lo = make_desc(w0 + 16384, 2048, 128);
        const uint64_t bd_hi = make_desc(w_hi, lbo2, 128), bd_lo = make_desc(w_lo, lbo2, 128);
        const uint64_t kstep2 = (uint64_t)((2 * lbo2) >> 4);
        mbar_wait(BAR(B_WFULL + s0), (g / NSLOTS) & 1);
        bool w2_ready = false, w0_released = false;
        int stage0 = 0, stage1 = 0;  // 0: GEMM1, 1: GEMM2 first half, 2: GEMM2 second half, 3: done
        while (stage0 < 3 || stage1 < 3) {
#pragma unroll
          for (int t = 0; t < NWG; ++t) {
            const int stage = t ? stage1 : stage0;
            const uint32_t tb = tmem_u + t * TM_TILE;
            if (stage == 0) {
              if (!__all_sync(0xffffffffu, mbar_test(BAR(B_A1FULL + t), par))) continue;
              tc_fence_after();
              // GEMM1: H = y0 . W0 (K = 32 -> 4 k-steps, three split passes, small terms first)
              const uint32_t a_hi = sbase + SM_A1 + t * 32768;
              const uint64_t ad_hi = make_desc(a_hi, 2048, 128), ad_lo = make_desc(a_hi + 16384, 2048, 128);
              if (elect_one()) {
                uint32_t acc = 0;
#pragma unroll
                for (int pass = 0; pass < 3; ++pass) {
                  const uint64_t aa = (pass == 0) ? ad_lo : ad_hi;
                  const uint64_t bb = (pass == 1) ? wd_lo : wd_hi;
#pragma unroll
                  for (int k = 0; k < DH / 8; ++k) {
                    mma_ss(tb + TM_P, aa + (uint64_t)(k * 256), bb + (uint64_t)(k * 256), id128, acc);
                    acc = 1;
                  }
                }
                tc_commit(BAR(B_HFULL + t));
              }
              __syncwarp();
            } else if (stage <= 2) {
              const int half = stage - 1;
              if (!__all_sync(0xffffffffu, mbar_test(BAR((half ? B_A2B : B_A2A) + t), par))) continue;
              if (!w2_ready) {
                mbar_wait(BAR(B_WFULL + s1), ((g + 1) / NSLOTS) & 1);
                if (scaled) mbar_wait(BAR(B_WFULL + s2), ((g + 2) / NSLOTS) & 1);
                w2_ready = true;
              }
              tc_fence_after();
              // GEMM2 half: O (+)= h[:, 64 half ..] . W2[64 half .., :]  (A from TMEM, K = 64)
              if (elect_one()) {
                uint32_t acc = half ? 1u : 0u;
#pragma unroll
                for (int pass = 0; pass < 3; ++pass) {
                  const uint32_t aa = (pass == 0) ? tb + TM_Q : tb + TM_P + half * 64;
                  const uint64_t bb = ((pass == 1) ? bd_lo : bd_hi) + (uint64_t)(half * 8) * kstep2;
#pragma unroll
                  for (int k = 0; k < 8; ++k) {
                    mma_ts(tb + TM_O, aa + k * 8, bb + (uint64_t)k * kstep2, id2, acc);
                    acc = 1;
                  }
                }
                tc_commit(BAR((half ? B_OFULL : B_QFREE) + t));
              }
              __syncwarp();
            } else {
              continue;
            }
            if (t) ++stage1; else ++stage0;
          }
          if (!w0_released && stage0 >= 1 && stage1 >= 1) {  // both GEMM1s issued: free W0's slot
            if (elect_one()) tc_commit(BAR(B_WEMPTY + s0));
            __syncwarp();
            w0_released = true;
          }
        }
        if (elect_one()) {
          tc_commit(BAR(B_WEMPTY + s1));
          if (scaled) tc_commit(BAR(B_WEMPTY + s2));
        }
        __syncwarp();
        g += scaled ? 3 : 2;
      }
    }
  } else {
    // ===== epilogue warpgroups =====
    const int wg = warp >> 2;              // 0 / 1
    const int tid = threadIdx.x & 127;     // TMEM lane == sample row within the tile
    const uint32_t lane_base = ((uint32_t)((warp & 3) * 32)) << 16;
    const uint32_t tm_p = tmem_base + lane_base + wg * TM_TILE + TM_P;
    const uint32_t tm_q = tmem_base + lane_base + wg * TM_TILE + TM_Q;
    const uint32_t tm_o = tmem_base + lane_base + wg * TM_TILE + TM_O;
    uint8_t* a1 = smem + SM_A1 + wg * 32768;
    double* scratch = reinterpret_cast<double*>(smem + SM_MISC + 16) + wg * 12;
    const int flusher = blockIdx.x * NWG + wg;
    const long long row_lo = (long long)flusher * P.rows_per_wg;
    long long row_hi = row_lo + P.rows_per_wg;
    if (row_hi > P.nrows) row_hi = P.nrows;
    const bool had_rows = row_lo < P.nrows;
    // named barriers 1 and 2 are private to the two warpgroups
    Fold<128> fold(fp, flusher, tid, scratch, (FOLD && had_rows) ? row_lo : 0, 1 + wg);
    const float invT = 1.f / P.T;
    const float base_const = -(float)D * logf(P.T) - 0.5f * (float)D * 1.8378770664093453f;
    uint32_t it = 0;

    for (int b = 0; b < P.nbatches; ++b) {
      const long long t0 = row_lo + (long long)b * TILE;
      const long long r = t0 + tid;
      const bool live = r < row_hi;
      const long long rr = live ? r : 0;
      float y[D];
      load_row<XT>(x, ld, rr, y);
      if (r + TILE < row_hi) {  // pull the next tile's row into L2 while this tile computes
        const char* nx = reinterpret_cast<const char*>(x + (r + TILE) * ld);
#pragma unroll
        for (int q = 0; q < (int)(D * sizeof(XT)); q += 128) prefetch_l2(nx + q);
      }
#pragma unroll
      for (int i = 0; i < D; ++i) y[i] = fmaf(y[i], pre_s[i], pre_s[D + i]);
      float ldj = 0.f;
      for (int l = 0; l < nl; ++l, ++it) {
        const uint32_t par = it & 1;
        const bool scaled = l < P.nscaled;
        const float* b0 = bias_s + l * 192;
        const float* b2 = b0 + 128;
        // ---- A1: y0 split hi/lo into the canonical smem layout (row tid, 4 k per 16-byte chunk)
#pragma unroll
        for (int c = 0; c < DH / 4; ++c) {
          uint4 hi, lo;
          uint32_t* ph = &hi.x;
          uint32_t* pl = &lo.x;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float v = y[4 * c + q];
            const uint32_t hb = __float_as_uint(v) & 0xffffe000u;
            ph[q] = hb;
            pl[q] = __float_as_uint(v - __uint_as_float(hb));
          }
          *reinterpret_cast<uint4*>(a1 + c * 2048 + tid * 16) = hi;
          *reinterpret_cast<uint4*>(a1 + 16384 + c * 2048 + tid * 16) = lo;
        }
        fence_proxy_async();
        tc_fence_before();  // orders the previous layer's TMEM loads (O) before the next MMAs
        mbar_arrive(BAR(B_A1FULL + wg));
        // ---- E1: h = leaky_relu(H + b0) -> hi in place, lo into Q; two 64-column halves
        mbar_wait(BAR(B_HFULL + wg), par);
        tc_fence_after();
        {
          uint32_t v[32], w[32];
#pragma unroll 1
          for (int c = 0; c < 2; ++c) {  // first hidden half
            tmem_ld32(tm_p + c * 32, v);
            tmem_ld_wait();
            e1_chunk(v, w, b0 + c * 32);
            tmem_st32(tm_p + c * 32, v);
            tmem_st32(tm_q + c * 32, w);
          }
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(BAR(B_A2A + wg));
          // second half: compute chunk 2 while GEMM2's first half still reads Q
          tmem_ld32(tm_p + 64, v);
          tmem_ld_wait();
          e1_chunk(v, w, b0 + 64);
          tmem_st32(tm_p + 64, v);
          mbar_wait(BAR(B_QFREE + wg), par);
          tc_fence_after();
          tmem_st32(tm_q, w);
          tmem_ld32(tm_p + 96, v);
          tmem_ld_wait();
          e1_chunk(v, w, b0 + 96);
          tmem_st32(tm_p + 96, v);
          tmem_st32(tm_q + 32, w);
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(BAR(B_A2B + wg));
        }
        // ---- E2: y1 = (y1 - shift) / scale, ldj -= sum log scale
        mbar_wait(BAR(B_OFULL + wg), par);
        tc_fence_after();
        {
          uint32_t sh[32];
          tmem_ld32(tm_o, sh);
          if (scaled) {
            uint32_t sc[32];
            tmem_ld32(tm_o + 32, sc);
            tmem_ld_wait();
#pragma unroll
            for (int k0 = 0; k0 < U; k0 += 8) {
              float prod = 1.f;  // 8 factors in [1e-3, 1e3] stay inside the float32 range
#pragma unroll
              for (int k = k0; k < k0 + 8; ++k) {
                const float shift = __uint_as_float(sh[k]) + b2[k];
                const float a = __uint_as_float(sc[k]) + b2[U + k];
                float s = softplus_clip_fast(a);
                s = (a != a) ? a : s;  // keep NaN (fminf/fmaxf would drop it)
                y[DH + k] = (y[DH + k] - shift) * rcp_approx(s);
                prod *= s;
              }
              ldj -= 0.6931471805599453f * lg2_approx(prod);
            }
          } else {
            tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < U; ++k) y[DH + k] -= __uint_as_float(sh[k]) + b2[k];
          }
        }
        // ---- inverse of tfb.Permute([D-1, 0, .., D-2]): rotate left by one (flows.py:65-66)
        if (l + 1 < nl) {
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
      if (live && lnpred_out) lnpred_out[r] = lnphi;
      if (FOLD) {
        float a[1] = {lnphi};
        float bb[1] = {live ? (float)lnpost[r] : 0.f};
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
  if (warp == 9) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

}  // namespace tc

// ---------------------------------------------------------------------------------------------
// host side: weight image (hi/lo split, canonical K-major no-swizzle layout)
// ---------------------------------------------------------------------------------------------
static inline void split_tf32(float v, float* hi, float* lo) {
  uint32_t b;
  memcpy(&b, &v, 4);
  b &= 0xffffe000u;
  memcpy(hi, &b, 4);
  *lo = v - *hi;
}

// writes B[n][k] = W[k*ldw + n] (flax kernel is (in, out)) for n < N, k < K into hi/lo blocks
static void pack_b(float* hi, float* lo, int N, int K, const float* W, int ldw, int n0, int nvalid) {
  for (int n = 0; n < nvalid; ++n)
    for (int k = 0; k < K; ++k) {
      const size_t off = (size_t)((n0 + n) / 8) * 32 + (size_t)((n0 + n) % 8) * 4 + (size_t)(k / 4) * (N * 4) + (k % 4);
      split_tf32(W[(size_t)k * ldw + n], &hi[off], &lo[off]);
    }
}

int realnvp_tc_prepare(Flow* f) {
  f->tc_ok = false;
  const hb_flow_desc& ds = f->desc;
  if (ds.kind != HB_FLOW_REALNVP || ds.ndim != tc::D) return HB_OK;
  const int nl = f->nlayers;
  if (nl > tc::MAX_TC_LAYERS) return HB_OK;
  const int d = f->d, u = f->u;
  size_t nchunks = 0;
  for (int l = 0; l < nl; ++l) nchunks += (l < ds.n_scaled) ? 3 : 2;
  std::vector<float> img(nchunks * (tc::SLOT_BYTES / 4), 0.f);
  std::vector<float> bias((size_t)nl * 192, 0.f);
  const float* w = f->host_weights.data();
  size_t ch = 0;
  for (int l = 0; l < nl; ++l) {
    const bool scaled = l < ds.n_scaled;
    const float* W0 = w + f->layer_off[l];
    const float* b0 = W0 + (size_t)d * 128;
    const float* Wt = b0 + 128;
    const float* bt = Wt + (size_t)128 * u;
    const float* Ws = bt + u;
    const float* bs = Ws + (size_t)128 * u;
    float* c0 = img.data() + ch * (tc::SLOT_BYTES / 4);
    pack_b(c0, c0 + 4096, 128, d, W0, 128, 0, 128);  // hi 16 KB | lo 16 KB
    ++ch;
    if (scaled) {
      float* ca = img.data() + ch * (tc::SLOT_BYTES / 4);
      float* cb = ca + tc::SLOT_BYTES / 4;
      pack_b(ca, cb, 64, 128, Wt, u, 0, u);
      pack_b(ca, cb, 64, 128, Ws, u, u, u);
      ch += 2;
    } else {
      float* ca = img.data() + ch * (tc::SLOT_BYTES / 4);
      pack_b(ca, ca + 4096, 32, 128, Wt, u, 0, u);
      ++ch;
    }
    for (int j = 0; j < 128; ++j) bias[(size_t)l * 192 + j] = b0[j];
    for (int k = 0; k < u; ++k) bias[(size_t)l * 192 + 128 + k] = bt[k];
    if (scaled)
      for (int k = 0; k < u; ++k) bias[(size_t)l * 192 + 128 + u + k] = bs[k];
  }
  const size_t img_bytes = img.size() * 4, bias_bytes = bias.size() * 4;
  HB_CUDA(cudaMalloc(&f->tc_image, img_bytes + bias_bytes));
  HB_CUDA(cudaMemcpy(f->tc_image, img.data(), img_bytes, cudaMemcpyHostToDevice));
  HB_CUDA(cudaMemcpy((char*)f->tc_image + img_bytes, bias.data(), bias_bytes, cudaMemcpyHostToDevice));
  f->tc_image_bytes = img_bytes;
  f->tc_ok = true;
  return HB_OK;
}

void realnvp_tc_release(Flow* f) {
  if (f->tc_image) cudaFree(f->tc_image);
  f->tc_image = nullptr;
}

template <typename XT, typename LT, bool FOLD>
static int tc_launch(const tc::Params& P, int grid, const void* x, long long ld, const void* lnpost,
                     float* lnpred_out, const FoldParams& fp, cudaStream_t st) {
  auto kern = tc::realnvp_tc_kernel<XT, LT, FOLD>;
  HB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SM_TOTAL));
  kern<<<grid, tc::THREADS, tc::SM_TOTAL, st>>>(P, (const XT*)x, ld, (const LT*)lnpost, lnpred_out, fp);
  return HB_OK;
}

int realnvp_tc_run(Flow* f, float T, const void* x, int x_dtype, long long ld, long long n,
                   const void* lnpost, int lp_dtype, float* lnpred_out, Accum* acc,
                   const int64_t* start, long long chain_lo, long long chain_hi, long long a,
                   long long b, cudaStream_t st) {
  if (!f->tc_ok) {
    set_error("tcgen05 path not available for this flow");
    return HB_ERR_UNSUPPORTED;
  }
  if (n <= 0) return HB_OK;
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, f->device);
  long long nwg = (long long)sms * tc::NWG;
  long long rpw = (n + nwg - 1) / nwg;
  rpw = ((rpw + tc::TILE - 1) / tc::TILE) * tc::TILE;
  nwg = (n + rpw - 1) / rpw;
  const int grid = (int)((nwg + tc::NWG - 1) / tc::NWG);
  const long long nfl = (long long)grid * tc::NWG;
  tc::Params P;
  P.image = static_cast<const uint8_t*>(f->tc_image);
  P.bias = reinterpret_cast<const float*>(static_cast<const char*>(f->tc_image) + f->tc_image_bytes);
  P.pre = f->desc.standardize ? f->dev_weights : nullptr;
  P.nlayers = f->nlayers;
  P.nscaled = f->desc.n_scaled;
  P.T = T;
  P.sum_log_amp = f->sum_log_amp;
  P.nrows = n;
  P.rows_per_wg = rpw;
  P.nbatches = (int)(rpw / tc::TILE);
  FoldParams fp{};
  int rc;
  if (acc) {
    rc = prepare_fold(acc, start, chain_lo, chain_hi, a, b, nfl, rpw, &fp, st);
    if (rc) return rc;
  }
  timing_begin(st);
  if (acc) {
    if (x_dtype == HB_F32 && lp_dtype == HB_F32) rc = tc_launch<float, float, true>(P, grid, x, ld, lnpost, lnpred_out, fp, st);
    else if (x_dtype == HB_F32) rc = tc_launch<float, double, true>(P, grid, x, ld, lnpost, lnpred_out, fp, st);
    else if (lp_dtype == HB_F32) rc = tc_launch<double, float, true>(P, grid, x, ld, lnpost, lnpred_out, fp, st);
    else rc = tc_launch<double, double, true>(P, grid, x, ld, lnpost, lnpred_out, fp, st);
  } else {
    if (x_dtype == HB_F32) rc = tc_launch<float, float, false>(P, grid, x, ld, nullptr, lnpred_out, fp, st);
    else rc = tc_launch<double, float, false>(P, grid, x, ld, nullptr, lnpred_out, fp, st);
  }
  timing_end(st);
  if (rc) return rc;
  count_launch();
  HB_CUDA(cudaGetLastError());
  if (acc) return launch_merge(acc, fp, chain_lo, nfl, st);
  return HB_OK;
}

}  // namespace hb
