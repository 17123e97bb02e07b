// tcgen05 RealNVP kernel (sm_100a): fused flow evaluation + per-chain fold.
//
// Replaces harmonic/flows.py:41-116,159-180 (tfp-jax RealNVP log_prob) + harmonic/model.py:222-235
// + harmonic/evidence.py:278-352 for the shapes the tensor path supports (ndim = 64, <= 8 coupling
// layers; everything else runs on the CUDA-core kernels).
//
// Mapping.  One persistent CTA per SM, three 128-sample tiles in flight:
//   3 epilogue warpgroups : each owns one tile at a time, one sample per thread (TMEM lane = sample).  A thread keeps
//                     the 32 initially transformed features + the 7 that can enter, owned by ORIGINAL index
//   producer warp   : streams the pre-split weight image (FP16 blocks, UMMA canonical no-swizzle K-major layout,
//                     built once on the host) from L2 into an 8-slot ring of 20 KiB slots (3 - 4 chunks per layer)
//                     with cp.async.bulk + mbarrier
//   3 MMA warps     : tcgen05.mma issuers, warp W_MMA + t serves tile t (the first also allocates TMEM); warp-uniform
//                     straight-line program with blocking mbarrier waits, one elected lane issues.  One issuing warp
//                     sustains one tcgen05.mma per ~49 cycles whatever N <= 64 is, while the tensor pipe itself needs
//                     N / 2 cycles (profiles/r2_issue_probe.txt)
// The epilogue warpgroups run with 160 registers, the producer / MMA warpgroup with 32 (setmaxnreg).
//
// Arithmetic.  leaky_relu(h) = a h + b relu(h) with a = 0.01, b = 1 - 0.01 (harmonic/flows.py:170
// LeakyReLU slope 0.01), so with H = x' W0' (x' = conditioning inputs + constant one, W0' = [W0 ; b0]) a coupling MLP is
//     O = leaky_relu(H) W2 + b2 = x' V + relu(H) U,  V = a W0' W2 (+ b2 on the constant row),  U = b W2
// V (34 x 32) is formed once on the host in float64.  The linear term rides in GEMM1 as 32 more output columns
// (N = 160: [W0' | V]); the epilogue between the GEMMs (E1) is then only relu(H) and its FP16 split -- and the clamp
// is a modifier of the conversion instructions: no arithmetic for the activation,
// no bias, nothing but conversions -- which is what the kernel's issue slots go to (profiles/r2_alu_probe.txt).
//
// Per coupling layer and tile
//   GEMM1  [H | O] [128 x 160] = [32 input slots | newest column | 1] . [W0' | V]   A, B from shared memory, 7 MMAs
//   E1     r = relu(H) -> FP16 pair (hs, r - hs), in place of H              eight 16-column chunks, loads prefetched
//   GEMM2  O[128 x 32] += r[:, quarter] . U[quarter, :]                      A from TMEM, 6 FP16 MMAs per K = 32 quarter
//   E2     y -= O (shift) or softplus-clip scale epilogue on y, log-det
// Scaled layers run GEMM1 with V_scale, GEMM2 with U_scale, hand O to the epilogue, then a second pass O = x' V_shift
// (7 small MMAs) + relu(H) U_shift over the same operands, which runs under the softplus / reciprocal / log-det arithmetic.
//
// Split-precision contractions with FP32 accumulation in TMEM, float32-equivalent products (~2^-22), every MMA
// kind::f16.  A weight is w_hi (11 significant bits) + w_lo; an activation a is hs = fp16(a) (round to nearest) and
// the exact remainder lo = a - hs, rounded to fp16:
//     a w ~= hs w_hi + lo w_hi + hs w_lo                                       (the lo w_lo term, <= 2^-22, dropped)
// with s = 1 everywhere: hs = fp16(a) straight out of the register / accumulator and lo = a - hs by ONE
// mixed-precision FMA (FHFMA), both against the SAME hi block.  So that w_lo stays a normal half the weights travel
// scaled: H accumulates SH = 2^5 times its value (relu is positively homogeneous, E1 does not care), O accumulates
// SO = 2^10 times its value; E2 removes the factor inside its FMA.
// What remains against a float32 FMA chain is the tensor cores' truncating accumulator
// (profiles/r1_tensor_accum_rounding.txt), see DESIGN.md 4.1.
//
// Range guard.  hs = fp16(relu(H)) needs H < 65504 and the truncating accumulator costs accuracy where the flow is
// evaluated far outside its trained range.  Every epilogue thread tracks max |conditioning input| of its rows (the 32
// raw features and each layer's new column); rows beyond Params::guard_limit are counted into a device counter which
// the host reads with the results and answers by re-evaluating the call on the CUDA-core path (hb_api / evidence.py).
// NaN inputs do not trip it: they propagate through the MMAs to a NaN ln phi as in the reference.
//
// GEMM1 without re-staging activations.  The flow permutes by one position per layer
// (harmonic/flows.py:65-66), so layer l conditions on ORIGINAL features l .. l+31: the features below
// 32 are still the raw (standardised) inputs, and the j < l features 32+j were frozen by layer j's
// epilogue.  The A operand is therefore written ONCE per tile -- the raw x(0..31) in 32 input slots -- plus one
// new column per layer: layer l's epilogue writes feature 32 + l into the K' = 16 "correction" operand
// [lo, hs, hs, 1, 1, 0 ...] (met by ONE MMA against [hi, hi, lo of its weights; hi, lo of the bias]) and into input
// slot l, which no later layer reads as a raw input, where the layers after the next find it.  Each layer's
// [W0' | V] is stored accordingly ("main", K = 32: 6 MMAs).  The main part of layer l+1 is issued as soon as layer
// l's epilogue has read O, i.e. it runs under layer l's epilogue -- slot l is being overwritten at that time and
// carries zero weights in layer l+1 --; only the 1-MMA correction waits for the new column.
//
// GEMM2 is pipelined with E1 at quarter granularity: quarter q's operands replace its own 32 H columns, its
// GEMM2 step runs while the epilogue computes quarter q + 1 (one mbarrier per quarter: the epilogue may run a
// whole layer ahead of the MMA warp, never two phases of one barrier).
// TMEM: tile t at column 160 t: P (128: H, then per 16-column chunk 8 columns of packed lo pairs + 8 of packed hs
// pairs) | O (32).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "hb_internal.h"
#include "hb_tc_ptx.cuh"

namespace hb {

namespace tc {

constexpr int D = 64;
constexpr int DH = 32;        // d = conditioning width (round(D/2))
constexpr int U = 32;         // transformed width
constexpr int TILE = 128;
constexpr int NT = 3;         // tiles (= epilogue warpgroups) in flight per CTA
constexpr int NSLOTS = 8;
constexpr int SLOT_BYTES = 20480;
constexpr int MAX_TC_LAYERS = 8;  // the epilogue keeps 7 features that can enter the transformed half in registers
constexpr int KA = 34;            // rows of a layer's effective GEMM1 weight matrix: 32 input slots, the newest column, the constant
constexpr int A_TILE = 10 * 2048;     // FP16 block, 10 chunks of 8 columns: lo 0-3 | hs 4-7 | correction operand 8 | zero 9
constexpr int N1 = 128 + U;       // GEMM1 output columns: H | O
constexpr float SH = 32.0f;                        // H accumulates SH times its value (so that w_lo is a normal half)
constexpr float SU = 32.0f;                        // GEMM2 weights travel times SU (u_lo a normal half)
constexpr float SO = SH * SU, SO_INV = 1.0f / SO;  // hence O accumulates SO times its value
constexpr double LEAKY_SLOPE = 0.01;

// weight image, per layer (FP16 blocks, hi = the 11 leading bits of a weight, lo = the rest; ONE hi block serves
// both halves of an activation's split):
//   chunk 0  GA   [160 x 64 K']: K' < 32 hi (meets lo AND hs of the 32 input slots), K' >= 32 lo (meets hs)
//   chunk 1  GB   correction [160 x 16 K'] over [lo new, hs new, hs new, 1, 1, 0 ...]: hi, hi, lo of the newest
//                 column's weights, hi, lo of the bias;
//                 then unscaled layers: GEMM2 quarters 0, 1; scaled layers: VT = the GEMM1 blocks of V_shift (N = 32)
//   chunk 2, (3)  GEMM2 quarters 2, 3 (unscaled) / 0, 1 and 2, 3 (scaled)
//   a GEMM2 quarter (K = 32 hidden units) is [N2 x 64 K']: K' < 32 hi (meets hs - and lo), K' >= 32 lo (meets hs);
//   scaled layers N2 = 64: rows 0..31 shift, 32..63 scale
constexpr int GA_BYTES = N1 * 64 * 2;
constexpr int GB_CORR_BYTES = N1 * 16 * 2;     // one correction block
constexpr int Q32_BYTES = 32 * 64 * 2;
constexpr int Q64_BYTES = 64 * 64 * 2;
constexpr int VT_BYTES = U * (64 + 16) * 2;
constexpr int GB_U_BYTES = GB_CORR_BYTES + 2 * Q32_BYTES, GB_S_BYTES = GB_CORR_BYTES + VT_BYTES;
static_assert(GA_BYTES <= SLOT_BYTES && GB_U_BYTES <= SLOT_BYTES && GB_S_BYTES <= SLOT_BYTES && 2 * Q64_BYTES <= SLOT_BYTES, "slot");
__host__ __device__ __forceinline__ constexpr int layer_chunks(bool scaled) { return scaled ? 4 : 3; }
__host__ __device__ __forceinline__ constexpr uint32_t chunk_bytes(bool scaled, int j) {
  return j == 0 ? GA_BYTES : j == 1 ? (scaled ? GB_S_BYTES : GB_U_BYTES) : (scaled ? 2 * Q64_BYTES : 2 * Q32_BYTES);
}

constexpr int THREADS = NT * 128 + 128;  // + one auxiliary warpgroup: producer warp, one MMA warp per tile
constexpr int REGS_EPI = 160, REGS_AUX = 32;  // setmaxnreg: the CTA launches with 512 x 128 registers; the auxiliary warpgroup hands
                                              // back 128 x 96 = 12288, the three epilogue warpgroups take 3 x 128 x 32 = all of them
// shared memory carve-up (bytes)
constexpr int SM_RING = 0;
constexpr int SM_A1 = SM_RING + NSLOTS * SLOT_BYTES;  // per tile: the FP16 A operand (A_TILE)
constexpr int SM_PRE = SM_A1 + NT * A_TILE;           // standardisation: inv_amp[D], -offset*inv_amp[D]
constexpr int SM_BAR = SM_PRE + 2 * D * 4;            // mbarriers
constexpr int SM_MISC = SM_BAR + 512;                 // tmem ptr, fold scratch
constexpr int SM_TOTAL = SM_MISC + 512;
// barrier indices (8 bytes each); per-tile barriers are indexed + tile
constexpr int B_WFULL = 0, B_WEMPTY = 8, B_A1FULL = 16, B_HFULL = B_A1FULL + NT,
              B_A2 = B_HFULL + NT /* [quarter][tile], one phase per layer */, B_OFULL = B_A2 + 4 * NT /* one phase per pass */,
              B_OCONS = B_OFULL + NT /* O is in the epilogue's registers: one phase per pass */, B_COUNT = B_OCONS + NT;
static_assert(B_COUNT * 8 <= 512 && NSLOTS <= 8, "barrier block");
// TMEM columns (per tile: + TM_TILE * tile)
constexpr uint32_t TM_P = 0, TM_O = 128, TM_TILE = 160;

struct Params {
  const uint8_t* image;   // weight image: chunks of SLOT_BYTES in consumption order
  const float* pre;       // pre_offset[D], pre_amp[D] or nullptr
  int nlayers;
  int nscaled;
  float T, sum_log_amp;
  unsigned long long* guard;  // rows beyond guard_limit (device counter), nullptr = not counted
  float guard_raw;            // largest |standardised input| (accuracy of the truncating accumulator)
  float guard_all;            // largest |conditioning input|, new columns included (|H| must stay a finite half)
  long long nrows;
  long long rows_per_wg;  // contiguous rows per epilogue warpgroup (multiple of TILE)
  int nbatches;           // tiles per warpgroup
};

// ---- optional phase trace (-DHB_TC_TRACE, development builds only; see tools/tc_trace.py) ----
#ifdef HB_TC_TRACE
constexpr int TRACE_LEN = 4096;
__device__ unsigned long long g_trace[4][TRACE_LEN];
__device__ int g_trace_n[4];
#define TR(code)                                                                              \
  do {                                                                                        \
    if (tr_on && tr_n < TRACE_LEN)                                                            \
      g_trace[tr_s][tr_n++] = ((unsigned long long)clock64() << 8) | (unsigned long long)(code); \
  } while (0)
#else
#define TR(code)
#endif

using namespace tcx;  // PTX wrappers and weight-image helpers (hb_tc_ptx.cuh)

// clip(softplus(a), 1e-3, 1e3) (harmonic/flows.py:177).  t = exp(-|a|) in (0, 1] by MUFU ex2;
// log1p(t) = 2 atanh(u), u = t / (2 + t) in (0, 1/3], as u times a degree-4 minimax polynomial in u^2
// (relative error 8e-9): for a < 0 the scale IS log1p(t), so it has to be accurate in relative terms.
__device__ __forceinline__ float softplus_clip_fast(float a) {
  const float t = ex2_approx(-1.4426950408889634f * fabsf(a));
  const float u = t * rcp_approx(2.0f + t);
  const float z = u * u;
  const float p = fmaf(z, fmaf(z, fmaf(z, fmaf(z, 0.14087678f, 0.13980642f), 0.20012431f), 0.33333158f), 1.0f);
  const float sp = fmaf(u + u, p, fmaxf(a, 0.f));
  return fminf(fmaxf(sp, 1e-3f), 1e3f);  // NaN in -> NaN out is handled by the caller's y1 path
}

// DH consecutive values of one sample row -> float32 registers (16-byte loads when the row allows)
template <typename XT>
__device__ __forceinline__ void load_half(const XT* __restrict__ p, float (&y)[DH]);

template <>
__device__ __forceinline__ void load_half<float>(const float* __restrict__ p, float (&y)[DH]) {
  if ((((uintptr_t)p) & 15) == 0) {
#pragma unroll
    for (int i = 0; i < DH / 4; ++i) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(p) + i);
      y[4 * i] = v.x; y[4 * i + 1] = v.y; y[4 * i + 2] = v.z; y[4 * i + 3] = v.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < DH; ++i) y[i] = __ldg(p + i);
  }
}
template <>
__device__ __forceinline__ void load_half<double>(const double* __restrict__ p, float (&y)[DH]) {
  if ((((uintptr_t)p) & 15) == 0) {
#pragma unroll
    for (int i = 0; i < DH / 2; ++i) {
      const double2 v = __ldg(reinterpret_cast<const double2*>(p) + i);
      y[2 * i] = (float)v.x; y[2 * i + 1] = (float)v.y;
    }
  } else {
#pragma unroll
    for (int i = 0; i < DH; ++i) y[i] = (float)__ldg(p + i);
  }
}

// E1 for one 16-column chunk of H: r = relu(H) as wh = 8 words of packed hs = fp16(r) pairs (rounded towards zero) and
// wl = 8 words of packed fp16(r - hs) pairs.  Per pair: F2FP.RELU.RZ, 2 FHFMA, F2FP.RELU -- the clamp is part of the
// conversions: for H < 0, hs = 0 and the "remainder" H - 0 < 0 is clamped by the second one; for H > 0 the remainder
// is non-negative because hs was rounded towards zero.
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

template <typename XT, typename LT, bool FOLD>
__global__ void __launch_bounds__(THREADS, 1)
realnvp_tc_kernel(Params P, const XT* __restrict__ x, long long ld, const LT* __restrict__ lnpost,
                  float* __restrict__ lnpred_out, FoldParams fp) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int NWG = NT;
  constexpr int W_PROD = 4 * NT, W_MMA = W_PROD + 1;  // weight producer warp; the MMA warps follow
  // lane-0 broadcast: tells the compiler the warp index (and with it the role dispatch below) is
  // warp-uniform, so the MMA warps' addresses, descriptors and loop state live on the uniform datapath
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar0 = sbase + SM_BAR;
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(smem + SM_MISC);
  float* pre_s = reinterpret_cast<float*>(smem + SM_PRE);
  const int nl = P.nlayers;
#ifdef HB_TC_TRACE
  int tr_n = 0;
#ifdef HB_TC_TRACE_SKEW  // the four warps of warpgroup 0 instead (warp skew at the barriers)
  const bool tr_on = blockIdx.x == 0 && lane == 0 && warp < 4;
  const int tr_s = warp;
#else
  const bool tr_on = blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 4 || warp == W_MMA || warp == W_MMA + 1);
  const int tr_s = warp == 0 ? 0 : warp == 4 ? 1 : warp == W_MMA ? 2 : 3;
#endif
#endif

  // ---- one-time setup -------------------------------------------------------------------
  if (threadIdx.x == 0) {
    for (int i = 0; i < NSLOTS; ++i) {
      mbar_init(BAR(B_WFULL + i), 1);
      mbar_init(BAR(B_WEMPTY + i), NWG);  // every MMA warp releases a slot
    }
    for (int t = 0; t < NWG; ++t) {
      mbar_init(BAR(B_A1FULL + t), 128);
      mbar_init(BAR(B_HFULL + t), 1);
      for (int q = 0; q < 4; ++q) mbar_init(BAR(B_A2 + NT * q + t), 128);
      mbar_init(BAR(B_OFULL + t), 1);
      mbar_init(BAR(B_OCONS + t), 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < D; i += THREADS) {
    const float inv = P.pre ? 1.f / P.pre[D + i] : 1.f;
    pre_s[i] = inv;
    pre_s[D + i] = P.pre ? -P.pre[i] * inv : 0.f;
  }
  if (warp == W_MMA) {
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
  // (setmaxnreg, first thing in every role: the epilogue warpgroups take the registers the producer / MMA warpgroups
  // do not need)

  if (warp == W_PROD) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_AUX));
    // ===== weight producer: the image is one batch worth of chunks, replayed every batch =====
    if (lane == 0) {
      uint32_t g = 0;
      for (int b = 0; b < P.nbatches; ++b) {
        const uint8_t* src = P.image;
        for (int l = 0; l < nl; ++l) {
          const bool scaled = l < P.nscaled;
          const int nc = layer_chunks(scaled);
          for (int j = 0; j < nc; ++j, ++g) {
            const uint32_t bytes = chunk_bytes(scaled, j);  // only the occupied prefix of a slot travels
            const uint32_t slot = g % NSLOTS, use = g / NSLOTS;
            mbar_wait_relaxed(BAR(B_WEMPTY + slot), (use & 1) ^ 1);
            mbar_expect_tx(BAR(B_WFULL + slot), bytes);
            bulk_g2s(sbase + SM_RING + slot * SLOT_BYTES, src, bytes, BAR(B_WFULL + slot));
            src += SLOT_BYTES;
          }
        }
      }
    }
  } else if (warp >= W_MMA) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_AUX));
    // ===== MMA issuers: warp W_MMA + t serves tile t.  Each runs a straight-line program per layer with blocking
    // mbarrier waits; the whole warp executes it (warp-uniform, so the descriptors stay in uniform registers) and
    // one elected lane issues.  ONE warp issues everything that accumulates into a tile's O: MMAs of different warps
    // are not executed in issue order (measured with two alternating issuers per tile: 10 % faster, but 3 - 20 % of
    // the rows differed in the last bit from run to run -- the accumulator truncates, so the order of the partial
    // sums shows), and results must not depend on scheduling.  A weight slot is released when every MMA warp has
    // committed on it.
    const int t = warp - W_MMA;
    const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0) + t * TM_TILE;
    const uint32_t ih1 = make_idesc_f16(N1), ih32 = make_idesc_f16(32);
    const uint32_t a_h = sbase + SM_A1 + t * A_TILE;  // FP16 block: lo raw 0-3 | hs raw 4-7 | lo new 8 | hs new 9
    const uint64_t adx = make_desc(a_h, 2048, 128), adn = make_desc(a_h + 8 * 2048, 2048, 128);
    auto wwait = [&](uint32_t gidx) {
      mbar_wait(BAR(B_WFULL + (gidx % NSLOTS)), (gidx / NSLOTS) & 1);
    };
    auto chunk_addr = [&](uint32_t gidx) { return sbase + SM_RING + (gidx % NSLOTS) * SLOT_BYTES; };
    auto release = [&](uint32_t gidx) { tc_commit(BAR(B_WEMPTY + (gidx % NSLOTS))); };
    // x' . B for the K = 32 raw inputs ("main"): B = the [n x 64 K'] block at ga (hi | lo).  Small terms first
    // (lo . hi, hs . lo), then the main term hs . hi
    auto issue_main = [&](uint32_t d, uint32_t ga, uint32_t n, uint32_t idesc, uint32_t acc0) {
      const uint64_t wh = make_desc(ga, n * 16, 128);
      const uint64_t kstep = (uint64_t)((n * 32) >> 4);
#pragma unroll
      for (int k = 0; k < 2; ++k) mma_ss_f16(d, adx + (uint64_t)(k * 256), wh + (uint64_t)k * kstep, idesc, k ? 1u : acc0);
#pragma unroll
      for (int k = 0; k < 2; ++k) mma_ss_f16(d, adx + (uint64_t)((2 + k) * 256), wh + (uint64_t)(2 + k) * kstep, idesc, 1);
#pragma unroll
      for (int k = 0; k < 2; ++k) mma_ss_f16(d, adx + (uint64_t)((2 + k) * 256), wh + (uint64_t)k * kstep, idesc, 1);
    };
    // the newest column and the constant ("corr"): ONE K' = 16 MMA over A = [lo new, hs new, hs new, 1, 1, 0 ...]
    auto issue_corr = [&](uint32_t d, uint32_t gc, uint32_t n, uint32_t idesc) {
      mma_ss_f16(d, adn, make_desc(gc, n * 16, 128), idesc, 1);
    };
    // one GEMM2 quarter: 6 MMAs over chunk pair (2 q, 2 q + 1) of P against the quarter block described by bd
    auto issue_quarter = [&](int q, uint64_t bd, uint64_t kstep2) {
      const uint32_t a0 = tb + TM_P + 32 * q;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const uint32_t a_lo = a0 + 16 * half, a_hs = a_lo + 8;
        mma_ts_f16(tb + TM_O, a_lo, bd + (uint64_t)half * kstep2, ih32, 1);
        mma_ts_f16(tb + TM_O, a_hs, bd + (uint64_t)(2 + half) * kstep2, ih32, 1);
        mma_ts_f16(tb + TM_O, a_hs, bd + (uint64_t)half * kstep2, ih32, 1);
      }
    };
    // main part of a layer's GEMM1 (chunk g = GA, waited for by the caller), released as soon as its MMAs have completed
    auto gemm1_main = [&](uint32_t g) {
      tc_fence_after();
      if (elect_one()) {
        issue_main(tb + TM_P, chunk_addr(g), N1, ih1, 0);
        release(g);
      }
      __syncwarp();
    };
    uint32_t g = 0, it = 0, pit = 0;
    for (int b = 0; b < P.nbatches; ++b) {
      for (int l = 0; l < nl; ++l, ++it) {
        const bool scaled = l < P.nscaled;
        const uint32_t par = it & 1;
        // this layer's GEMM2 operand blocks: quarter q lives in chunk gq(q) at offset qoff(q); the descriptors are
        // formed here, off the per-quarter path.  Scaled layers take the SCALE rows (32..63 of each quarter block:
        // + 4 row groups of 128 bytes) in the first pass: their epilogue is the long one
        const uint32_t lbo2 = scaled ? 64u * 16u : 32u * 16u;
        const uint32_t qbytes = scaled ? (uint32_t)Q64_BYTES : (uint32_t)Q32_BYTES;
        const uint64_t kstep2 = (uint64_t)((2 * lbo2) >> 4);
        auto gq = [&](int q) { return g + (scaled ? 2u : 1u) + (uint32_t)(q >> 1); };
        auto qoff = [&](int q) { return ((!scaled && q < 2) ? (uint32_t)GB_CORR_BYTES : 0u) + (uint32_t)(q & 1) * qbytes; };
        uint64_t bd[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) bd[q] = make_desc(chunk_addr(gq(q)) + qoff(q), lbo2, 128);
        const uint64_t pass1 = scaled ? (uint64_t)(512 >> 4) : 0ull;
        // ---- GEMM1: (main, unless issued early) + correction
        if (l > 0) wwait(g + 1);  // (idle anyway: the chunk checks come ahead of the epilogue's signals)
        mbar_wait(BAR(B_A1FULL + t), par);
        TR(20);
        if (l == 0) {
          wwait(g);
          gemm1_main(g);
          wwait(g + 1);
        }
        tc_fence_after();
        if (elect_one()) {
          issue_corr(tb + TM_P, chunk_addr(g + 1), N1, ih1);
          tc_commit(BAR(B_HFULL + t));
        }
        __syncwarp();
        TR(21);
        // ---- GEMM2: four K = 32 quarters, each as soon as its part of E1 has landed
        if (scaled) wwait(g + 2);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (q == 2) wwait(gq(2));
          mbar_wait(BAR(B_A2 + NT * q + t), par);
          TR(22 + q);
          tc_fence_after();
          if (elect_one()) {
            issue_quarter(q, bd[q] + pass1, kstep2);
            if (q == 3) tc_commit(BAR(B_OFULL + t));
            if (!scaled && (q & 1)) release(gq(q));
          }
          __syncwarp();
          TR(26 + q);
        }
        if (scaled) {
          // ---- second pass: O = x' V_shift + relu(H) U_shift (rows 0..31 of each quarter block) once the epilogue has
          // read the scale pre-activations out of O
          mbar_wait(BAR(B_OCONS + t), pit & 1);
          ++pit;
          tc_fence_after();
          TR(32);
          if (elect_one()) {
            const uint32_t vt = chunk_addr(g + 1) + GB_CORR_BYTES;
            issue_main(tb + TM_O, vt, U, ih32, 0);
            issue_corr(tb + TM_O, vt + U * 64 * 2, U, ih32);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              issue_quarter(q, bd[q], kstep2);
              if (q & 1) release(gq(q));
            }
            tc_commit(BAR(B_OFULL + t));
            release(g + 1);
          }
          __syncwarp();
          TR(33);
        }
        g += layer_chunks(scaled);
        // ---- O is in the epilogue's registers (and with it every MMA that read P has completed): the main part of
        // the next layer overwrites P and O under this layer's E2
        if (l + 1 < nl) wwait(g);
        mbar_wait(BAR(B_OCONS + t), pit & 1);
        ++pit;
        TR(30);
        if (l + 1 < nl) {
          gemm1_main(g);
          TR(31);
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_EPI));
    // ===== epilogue warpgroups =====
    // Features are owned by ORIGINAL index and nothing rotates.  yt[i] = feature 32 + i,
    // x7[i] = feature i < 7 (feature i joins the transformed half from layer i + 1 on); the features that no layer
    // touches (7..31) only contribute their squares, summed at the tile start.  At layer l feature f is GEMM2 output
    // column k = (f - 32 - l) mod 64, so O is read at a column offset of -l: entry i of the 32-column load is
    // feature 32 + i for i >= l (below that the feature is frozen and the word -- whatever lies left of O -- is
    // discarded); features i < l sit at columns 32 - l + i.  Scaled layers: O receives the scale
    // pre-activations, then (second MMA pass, after OCONS, under the scale arithmetic) the shifts.
    const int wg = warp >> 2;              // 0..2 = tile
    const int tid = threadIdx.x & 127;     // TMEM lane == sample row within the tile
    const uint32_t lane_base = ((uint32_t)((warp & 3) * 32)) << 16;
    const uint32_t tm_p = tmem_base + lane_base + wg * TM_TILE + TM_P;
    const uint32_t tm_o = tmem_base + lane_base + wg * TM_TILE + TM_O;
    uint8_t* a1 = smem + SM_A1 + wg * A_TILE;
    double* scratch = reinterpret_cast<double*>(smem + SM_MISC + 16) + wg * 12;
    const int flusher = blockIdx.x * NWG + wg;
    const long long row_lo = (long long)flusher * P.rows_per_wg;
    long long row_hi = row_lo + P.rows_per_wg;
    if (row_hi > P.nrows) row_hi = P.nrows;
    const bool had_rows = row_lo < P.nrows;
    Fold<128> fold(fp, flusher, tid, scratch, (FOLD && had_rows) ? row_lo : 0, 1 + wg);  // named barriers 1..3
    const float invT = 1.f / P.T;
    const float base_const = -(float)D * logf(P.T) - 0.5f * (float)D * 1.8378770664093453f;
    constexpr int NX = MAX_TC_LAYERS - 1;  // features that can enter the transformed half
    uint32_t it = 0, pit = 0;
    unsigned int nguard = 0;

    // The conditioning half of a tile's rows -- all the A operand needs -- is fetched into registers under the last
    // layer of the tile before (xpre); the transformed half is loaded under GEMM1 of layer 0.
    float xpre[DH];
    {
      const long long r0 = row_lo + tid;
      load_half<XT>(x + (r0 < row_hi ? r0 : 0) * ld, xpre);
    }
    for (int b = 0; b < P.nbatches; ++b) {
      const long long t0 = row_lo + (long long)b * TILE;
      const long long r = t0 + tid;
      const bool live = r < row_hi;
      const long long rr = live ? r : 0;
      float yt[U], x7[NX], ss = 0.f, gmax = 0.f, gnew = 0.f;
      TR(1);
      const XT* xrow = x + rr * ld;
      float lp = 0.f;
      if (FOLD) lp = live ? (float)lnpost[r] : 0.f;  // in flight until the fold at the end of the tile
      {
        float y[DH];
#pragma unroll
        for (int i = 0; i < DH; ++i) y[i] = fmaf(xpre[i], pre_s[i], pre_s[D + i]);
        // A operand, written once per tile: chunk c = columns 8 c .. 8 c + 7 (lo), chunk 4 + c the same columns (hs)
#pragma unroll
        for (int c = 0; c < DH / 8; ++c) {
          uint4 plo, phi;
          uint32_t* ql = &plo.x;
          uint32_t* qh = &phi.x;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float va = y[8 * c + 2 * q], vb = y[8 * c + 2 * q + 1];
            split_x_pair(va, vb, qh[q], ql[q]);
            gmax = fmaxf(gmax, fmaxf(fabsf(va), fabsf(vb)));  // fmaxf drops NaN: NaN rows do not trip the guard
          }
          *reinterpret_cast<uint4*>(a1 + c * 2048 + tid * 16) = plo;
          *reinterpret_cast<uint4*>(a1 + (4 + c) * 2048 + tid * 16) = phi;
        }
        {
          // correction operand [lo new, hs new, hs new, 1, 1, 0, 0, 0 | 0 x 8]: no new column yet
          *reinterpret_cast<uint4*>(a1 + 8 * 2048 + tid * 16) = make_uint4(0u, 0x3C000000u, 0x00003C00u, 0u);
          *reinterpret_cast<uint4*>(a1 + 9 * 2048 + tid * 16) = make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int i = 0; i < NX; ++i) x7[i] = y[i];
#pragma unroll
        for (int i = NX; i < DH; ++i) {
          const float z = y[i] * invT;
          ss = fmaf(z, z, ss);
        }
      }
      float ldj = 0.f;
      for (int l = 0; l < nl; ++l, ++it) {
        const uint32_t par = it & 1;
        const bool scaled = l < P.nscaled;
        fence_proxy_async();
        tc_fence_before();  // orders the previous layer's TMEM loads (O) before the next MMAs
        mbar_arrive(BAR(B_A1FULL + wg));
        TR(2);
        if (l == 0) {
          load_half<XT>(xrow + DH, yt);  // raw; standardised below, once GEMM1 has been waited for
          if (r + TILE < row_hi) {  // pull the next tile's row into L2 while this tile computes
            const char* nx = reinterpret_cast<const char*>(x + (r + TILE) * ld);
#pragma unroll
            for (int q = 0; q < (int)(D * sizeof(XT)); q += 128) prefetch_l2(nx + q);
          }
        }
        // ---- E1
        mbar_wait(BAR(B_HFULL + wg), par);
        tc_fence_after();
        TR(3);
        {
          uint32_t va[16], vb[16];
          tmem_ld16(tm_p, va);
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            uint32_t(&cur)[16] = (c & 1) ? vb : va;
            uint32_t(&nxt)[16] = (c & 1) ? va : vb;
            const int q = c >> 1;
            tmem_ld_wait16(cur);
            if (c < 7) tmem_ld16(tm_p + (c + 1) * 16, nxt);
            uint32_t wl[8], wh[8];
            e1_chunk(cur, wl, wh);
            tmem_st8(tm_p + c * 16, wl);
            tmem_st8(tm_p + c * 16 + 8, wh);
            if (c & 1) {
              tmem_st_wait();
              tc_fence_before();
              mbar_arrive(BAR(B_A2 + NT * q + wg));
              TR(4 + q);
            }
          }
        }
        if (l == 0) {
#pragma unroll
          for (int i = 0; i < U; ++i) yt[i] = fmaf(yt[i], pre_s[DH + i], pre_s[D + DH + i]);
        }
        if (l == nl - 1 && b + 1 < P.nbatches) {  // next tile's conditioning half: in flight under this layer's E2 and the fold
          const long long rn = r + TILE;
          load_half<XT>(x + (rn < row_hi ? rn : 0) * ld, xpre);
        }
        // ---- E2
        mbar_wait(BAR(B_OFULL + wg), pit & 1);
        ++pit;
        tc_fence_after();
        TR(9);
#ifdef HB_TC_DELAY_NS  // experiment (profiles/README.md): an idle wait in every layer's serial chain -- is the kernel
        // bound by the chain or by the power cap?
        asm volatile("nanosleep.u32 %0;" ::"n"(HB_TC_DELAY_NS));
#endif
        if (!scaled) {
          uint32_t v0[U], v1[8];
          tmem_ld32(tm_o - l, v0);
          tmem_ld8(tm_o + U - l, v1);
          tmem_ld_wait();
          tc_fence_before();
          mbar_arrive(BAR(B_OCONS + wg));
#pragma unroll
          for (int i = 0; i < U; ++i) {
            const float yn = fmaf(__uint_as_float(v0[i]), -SO_INV, yt[i]);
            yt[i] = (i >= NX || i >= l) ? yn : yt[i];
          }
#pragma unroll
          for (int i = 0; i < NX; ++i) {
            const float xn = fmaf(__uint_as_float(v1[i]), -SO_INV, x7[i]);
            x7[i] = (i < l) ? xn : x7[i];
          }
        } else {
          // Scaled layer: O holds the SCALE pre-activations first.  Once they are in registers the MMA warp may
          // overwrite O with the shift pass (OCONS), which then runs under the softplus / reciprocal / log-det
          // arithmetic below -- the long part of E2; the shifts are applied last: y = (y - t) * (1 / s).
          uint32_t v0[U], v1[8];
          tmem_ld32(tm_o - l, v0);
          tmem_ld8(tm_o + U - l, v1);
          tmem_ld_wait();
          tc_fence_before();
          // ptxas schedules across asm statements: without a data dependency it sinks this arrive below the ~2 000
          // cycles of arithmetic that follow and the second pass starts that much later.  The scale factor of the
          // pre-activations therefore takes (runtime-zero) bits from the arrive's state word.
          const uint32_t tok = mbar_arrive_token(BAR(B_OCONS + wg));
          const float so_inv = __uint_as_float(__float_as_uint(SO_INV) | (tok & (uint32_t)(P.nlayers >> 16)));
          // v0 / v1 become 1 / clip(softplus(a)) in place; ldj -= sum log scale
#pragma unroll
          for (int k0 = 0; k0 < U; k0 += 8) {
            float prod = 1.f;  // 8 factors in [1e-3, 1e3] stay inside the float32 range
#pragma unroll
            for (int i = k0; i < k0 + 8; ++i) {
              const bool act = i >= NX || i >= l;
              const float a = __uint_as_float(v0[i]) * so_inv;
              float sc = softplus_clip_fast(a);
              sc = (a != a) ? a : sc;  // keep NaN (fminf/fmaxf would drop it)
              v0[i] = __float_as_uint(rcp_approx(sc));
              prod *= act ? sc : 1.f;
            }
            ldj -= 0.6931471805599453f * lg2_approx(prod);
          }
          if (l > 0) {
            float prod = 1.f;
#pragma unroll
            for (int i = 0; i < NX; ++i) {
              if (i < l) {
                const float a = __uint_as_float(v1[i]) * so_inv;
                float sc = softplus_clip_fast(a);
                sc = (a != a) ? a : sc;
                v1[i] = __float_as_uint(rcp_approx(sc));
                prod *= sc;
              }
            }
            ldj -= 0.6931471805599453f * lg2_approx(prod);
          }
          // shifts (second MMA pass), in two 16-column loads + the columns of the entered features
          TR(11);
          mbar_wait(BAR(B_OFULL + wg), pit & 1);
          ++pit;
          tc_fence_after();
          TR(12);
          uint32_t w0[16], w1[8];
          tmem_ld16(tm_o - l, w0);
          tmem_ld8(tm_o + U - l, w1);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float yn = fmaf(__uint_as_float(w0[i]), -SO_INV, yt[i]) * __uint_as_float(v0[i]);
            yt[i] = (i >= NX || i >= l) ? yn : yt[i];
          }
          tmem_ld16(tm_o - l + 16, w0);
#pragma unroll
          for (int i = 0; i < NX; ++i) {
            const float xn = fmaf(__uint_as_float(w1[i]), -SO_INV, x7[i]) * __uint_as_float(v1[i]);
            x7[i] = (i < l) ? xn : x7[i];
          }
          tmem_ld_wait16(w0);
          tc_fence_before();
          mbar_arrive(BAR(B_OCONS + wg));
#pragma unroll
          for (int i = 16; i < U; ++i)  // features 48.. are transformed by every layer
            yt[i] = fmaf(__uint_as_float(w0[i - 16]), -SO_INV, yt[i]) * __uint_as_float(v0[i]);
        }
        if (l + 1 < nl) {
          // feature 32 + l is final from here on: it is the new conditioning column l of the following layers
          float v = yt[0];
#pragma unroll
          for (int i = 1; i < NX; ++i) v = (l == i) ? yt[i] : v;
          gnew = fmaxf(gnew, fabsf(v));
          // the next layer meets it in the correction operand; the layers after that in input slot l (dead as a raw
          // input from layer l + 1 on; the next layer's main MMAs, already in flight, meet that slot with zero weights)
          const uint32_t hs16 = to_f16(v), lo16 = to_f16(v - from_f16((uint16_t)hs16));
          *reinterpret_cast<uint2*>(a1 + 8 * 2048 + tid * 16) = make_uint2(lo16 | (hs16 << 16), hs16 | 0x3C000000u);
          *reinterpret_cast<uint16_t*>(a1 + tid * 16 + l * 2) = (uint16_t)lo16;
          *reinterpret_cast<uint16_t*>(a1 + 4 * 2048 + tid * 16 + l * 2) = (uint16_t)hs16;
        }
        TR(10);
      }
#pragma unroll
      for (int i = 0; i < U; ++i) {
        const float z = yt[i] * invT;
        ss = fmaf(z, z, ss);
      }
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        const float z = x7[i] * invT;
        ss = fmaf(z, z, ss);
      }
      const float lnphi = -0.5f * ss + base_const + ldj - P.sum_log_amp;
      nguard += (live && (gmax > P.guard_raw || gnew > P.guard_all)) ? 1u : 0u;
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
#ifdef HB_TC_TRACE
  if (tr_on) g_trace_n[tr_s] = tr_n;
#endif

  // ---- teardown ---------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

}  // namespace tc

#ifdef HB_TC_TRACE
extern "C" __attribute__((visibility("default"))) int hb_debug_trace_read(unsigned long long* dst, int* n) {
  if (cudaMemcpyFromSymbol(dst, tc::g_trace, sizeof(unsigned long long) * 4 * tc::TRACE_LEN) != cudaSuccess) return -1;
  if (cudaMemcpyFromSymbol(n, tc::g_trace_n, sizeof(int) * 4) != cudaSuccess) return -1;
  return 0;
}
#endif

// ---------------------------------------------------------------------------------------------
// host side: weight image (hi/lo split, canonical K-major no-swizzle layout, SLOT_BYTES chunks)
// ---------------------------------------------------------------------------------------------
using namespace tcx;  // split_hi_lo, kmajor_off16, f32_to_f16
namespace {

// The GEMM1-shaped blocks of an effective [KA x n_out] weight matrix M (rows 0..31 multiply the 32 input slots of the
// A operand, row 32 the newest column, row 33 the constant one) for output rows [n0, n0 + n_out) of a block that is N
// rows wide, stored times `scale`:
//   ga: [N x 64 K']: K' = k hi (meets lo and hs of slot k), K' = 32 + k lo (meets hs)
//   gc: [N x 16 K']: K' = 0, 1 hi and 2 lo of the newest column's row (meet lo new, hs new, hs new), 3 hi and 4 lo of
//       the constant row (meet the two ones)
struct Gemm1Writer {
  uint16_t *ga, *gc;
  int N;
  bool overflow = false;
  void put(const std::vector<double>& M, int n_out, int n0, float scale) {
    for (int k = 0; k < tc::KA; ++k)
      for (int n = 0; n < n_out; ++n) {
        const float w = (float)M[(size_t)k * n_out + n];
        float hi, lo;
        split_hi_lo(w, &hi, &lo);
        const uint16_t h = f32_to_f16(hi * scale, &overflow), l = f32_to_f16(lo * scale, &overflow);
        const int row = n0 + n;
        if (k < 32) {
          ga[kmajor_off16(N, row, k)] = h;
          ga[kmajor_off16(N, row, 32 + k)] = l;
        } else if (k == 32) {
          gc[kmajor_off16(N, row, 0)] = h;
          gc[kmajor_off16(N, row, 1)] = h;
          gc[kmajor_off16(N, row, 2)] = l;
        } else {
          gc[kmajor_off16(N, row, 3)] = h;
          gc[kmajor_off16(N, row, 4)] = l;
        }
      }
  }
};

}  // namespace

int realnvp_tc_prepare(Flow* f) {
  f->tc_ok = false;
  const hb_flow_desc& ds = f->desc;
  if (ds.kind != HB_FLOW_REALNVP || ds.ndim != tc::D) return HB_OK;
  const int nl = f->nlayers;
  if (nl > tc::MAX_TC_LAYERS) return HB_OK;
  const int d = f->d, u = f->u;
  if (d != tc::DH || u != tc::U) return HB_OK;
  const double alpha = tc::LEAKY_SLOPE, beta = 1.0 - tc::LEAKY_SLOPE;  // leaky_relu(h) = alpha h + beta relu(h)
  size_t nchunks = 0;
  for (int l = 0; l < nl; ++l) nchunks += tc::layer_chunks(l < ds.n_scaled);
  std::vector<uint8_t> img(nchunks * tc::SLOT_BYTES, 0);
  const float* w = f->host_weights.data();
  size_t ch = 0;
  bool overflow = false;
  double hsum_max = 0.0, hbias_max = 0.0;  // |H| <= hsum_max * max|x'| + hbias_max
  for (int l = 0; l < nl; ++l) {
    const bool scaled = l < ds.n_scaled;
    const float* W0 = w + f->layer_off[l];       // [d][128]
    const float* b0 = W0 + (size_t)d * 128;
    const float* Wt = b0 + 128;                  // [128][u]
    const float* bt = Wt + (size_t)128 * u;
    const float* Ws = bt + u;
    const float* bs = Ws + (size_t)128 * u;
    // W0eff[k][j]: input slot k < 32 holds original feature k for k >= l (W0 row k - l) and the frozen feature 32 + k
    // for k < l - 1 (W0 row 32 - l + k); slot l - 1 is being overwritten while this layer's main MMAs run and gets
    // zero weights -- its feature, the newest column, is row 32 (W0 row 31); row 33 is b0
    std::vector<double> W0eff((size_t)tc::KA * 128, 0.0);
    for (int j = 0; j < 128; ++j) {
      for (int k = l; k < 32; ++k) W0eff[(size_t)k * 128 + j] = W0[(size_t)(k - l) * 128 + j];
      for (int k = 0; k + 1 < l; ++k) W0eff[(size_t)k * 128 + j] = W0[(size_t)(32 - l + k) * 128 + j];
      if (l > 0) W0eff[(size_t)32 * 128 + j] = W0[(size_t)31 * 128 + j];
      W0eff[(size_t)33 * 128 + j] = b0[j];
      double sum = 0.0;
      for (int k = 0; k < 33; ++k) sum += fabs(W0eff[(size_t)k * 128 + j]);
      if (sum > hsum_max) hsum_max = sum;
      if (fabs((double)b0[j]) > hbias_max) hbias_max = fabs((double)b0[j]);
    }
    // V[k][n] = alpha sum_j W0eff[k][j] W2[j][n]  (+ b2[n] on the constant row)
    auto make_v = [&](const float* W2, const float* b2) {
      std::vector<double> V((size_t)tc::KA * u, 0.0);
      for (int k = 0; k < tc::KA; ++k)
        for (int n = 0; n < u; ++n) {
          double acc = 0.0;
          for (int j = 0; j < 128; ++j) acc += W0eff[(size_t)k * 128 + j] * (double)W2[(size_t)j * u + n];
          V[(size_t)k * u + n] = alpha * acc + (k == 33 ? (double)b2[n] : 0.0);
        }
      return V;
    };
    const std::vector<double> Vt = make_v(Wt, bt);
    uint8_t* base = img.data() + ch * tc::SLOT_BYTES;
    uint8_t* gb = base + tc::SLOT_BYTES;
    {
      Gemm1Writer wr{reinterpret_cast<uint16_t*>(base), reinterpret_cast<uint16_t*>(gb), tc::N1};
      wr.put(W0eff, 128, 0, tc::SH);
      wr.put(scaled ? make_v(Ws, bs) : Vt, u, 128, tc::SO);
      overflow |= wr.overflow;
    }
    // GEMM2: one K = 32 quarter = FP16 block [N2 x 64 K'] of u = beta W2 times SU: K' < 32 u_hi (meets lo and hs),
    // 32..63 u_lo (meets hs).  Scaled layers: rows 0..31 shift, 32..63 scale.
    const int N2 = scaled ? 64 : 32;
    const size_t qbytes = scaled ? tc::Q64_BYTES : tc::Q32_BYTES;
    for (int q = 0; q < 4; ++q) {
      uint8_t* qb = scaled ? base + (size_t)(2 + (q >> 1)) * tc::SLOT_BYTES + (q & 1) * qbytes
                           : (q < 2 ? gb + tc::GB_CORR_BYTES + q * qbytes
                                    : base + (size_t)2 * tc::SLOT_BYTES + (q & 1) * qbytes);
      uint16_t* c16 = reinterpret_cast<uint16_t*>(qb);
      for (int kk = 0; kk < 32; ++kk)
        for (int n = 0; n < N2; ++n) {
          const int k = 32 * q + kk;
          const float v = (float)(beta * (double)((n < u) ? Wt[(size_t)k * u + n] : Ws[(size_t)k * u + (n - u)]));
          float hi, lo;
          split_hi_lo(v, &hi, &lo);
          c16[kmajor_off16(N2, n, kk)] = f32_to_f16(hi * tc::SU, &overflow);
          c16[kmajor_off16(N2, n, 32 + kk)] = f32_to_f16(lo * tc::SU, &overflow);
        }
    }
    if (scaled) {
      uint8_t* vt = gb + tc::GB_CORR_BYTES;
      Gemm1Writer wr{reinterpret_cast<uint16_t*>(vt), reinterpret_cast<uint16_t*>(vt + u * 64 * 2), u};
      wr.put(Vt, u, 0, tc::SO);
      overflow |= wr.overflow;
    }
    ch += tc::layer_chunks(scaled);
  }
  if (overflow) return HB_OK;  // a weight block leaves the half range: the CUDA-core path serves this flow
  // range guard: hs = fp16(SH |H|) must stay finite (|H| <= hsum_max |x'|_max + |b0|_max), and beyond ~32 standardised
  // units the truncating accumulator shows (tests/test_gpu_flows.py::test_realnvp_tcgen05_large_magnitude)
  double hard = 1e30;
  if (hsum_max > 0.0) hard = (60000.0 / tc::SH - hbias_max) / hsum_max;
  if (!(hard > 1.0)) return HB_OK;
  f->tc_guard_all = (float)hard;
  f->tc_guard_raw = (float)(hard < 32.0 ? hard : 32.0);
  HB_CUDA(cudaMalloc(&f->tc_image, img.size()));
  HB_CUDA(cudaMemcpy(f->tc_image, img.data(), img.size(), cudaMemcpyHostToDevice));
  HB_CUDA(cudaMalloc(&f->tc_guard, sizeof(unsigned long long)));
  HB_CUDA(cudaMemset(f->tc_guard, 0, sizeof(unsigned long long)));
  f->tc_image_bytes = img.size();
  f->tc_ok = true;
  return HB_OK;
}

void realnvp_tc_release(Flow* f) {
  if (f->tc_image) cudaFree(f->tc_image);
  if (f->tc_guard) cudaFree(f->tc_guard);
  f->tc_image = nullptr;
  f->tc_guard = nullptr;
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
                   long long b, unsigned long long* guard, cudaStream_t st) {
  if (!f->tc_ok) {
    set_error("tcgen05 path not available for this flow");
    return HB_ERR_UNSUPPORTED;
  }
  if (n <= 0) return HB_OK;
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, f->device);
  const int NWG = tc::NT;
  long long nwg = (long long)sms * NWG;
  long long rpw = (n + nwg - 1) / nwg;
  rpw = ((rpw + tc::TILE - 1) / tc::TILE) * tc::TILE;
  nwg = (n + rpw - 1) / rpw;
  const int grid = (int)((nwg + NWG - 1) / NWG);
  const long long nfl = (long long)grid * NWG;
  tc::Params P;
  P.image = static_cast<const uint8_t*>(f->tc_image);
  P.pre = f->desc.standardize ? f->dev_weights : nullptr;
  P.nlayers = f->nlayers;
  P.nscaled = f->desc.n_scaled;
  P.T = T;
  P.sum_log_amp = f->sum_log_amp;
  P.guard = guard;
  P.guard_raw = f->tc_guard_raw;
  P.guard_all = f->tc_guard_all;
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
