// tcgen05 RealNVP kernel (sm_100a): fused flow evaluation + per-chain fold.
//
// Replaces harmonic/flows.py:41-116,159-180 (tfp-jax RealNVP log_prob) + harmonic/model.py:222-235
// + harmonic/evidence.py:278-352 for the shapes the tensor path supports (ndim = 64, <= 8 coupling
// layers; everything else runs on the CUDA-core kernels).
//
// Mapping.  One persistent CTA per SM, NT 128-sample tiles in flight (template parameter; 3 by default,
// HB_TC_TILES=2 selects the two-tile layout for A/B measurements), 128 NT + 32 (NT + 1) threads:
//   NT epilogue warpgroups : each owns one tile at a time, one sample per thread (TMEM lane = sample).
//                     NT = 3 (128 registers per thread): the 32 initially transformed features + the 7 that
//                     can enter, owned by ORIGINAL index; NT = 2: y[64] with a register rotation per layer
//   producer warp   : streams the pre-split weight image (FP16 blocks, UMMA canonical no-swizzle K-major
//                     layout, built once on the host) from L2 into an 8-slot ring of 16 KiB chunks with
//                     cp.async.bulk + mbarrier
//   NT MMA warps    : tcgen05.mma issuers, one per tile (the first also allocates TMEM); warp-uniform
//                     straight-line program with blocking mbarrier waits, one elected lane issues
//
// Per coupling layer and tile
//   GEMM1  H[128x128] = [x(0..31) | new(0..6) | 1] . W0'        A, B from shared memory, K = 40
//   E1     h = leaky_relu(H), compensated FP16 split, in place   eight 16-column chunks, TMEM loads prefetched
//   GEMM2  O[128xN2] += h[:, quarter] . [Wt|Ws][quarter, :]      A from TMEM, 6 FP16 MMAs per K = 32 quarter
//   E2     shift / softplus-clip scale epilogue on y1, log-det
// Split-precision contractions with FP32 accumulation in TMEM, float32-equivalent products (~2^-22), every MMA
// kind::f16 (see the comment at H2_SCALE).  An activation a (GEMM1: a standardised input, GEMM2: a hidden unit) is
// split with compensation, hs = fp16(a * 2^-11), lo = fp16(a - 2^11 hs); a weight as w_hi (11 significant bits) +
// w_lo; then  a.w ~= hs.(w_hi 2^11) + hs.(w_lo 2^11) + lo.w_hi  -- the SAME hs block feeds two MMAs.  A tcgen05.mma
// costs max(~45, N/2) cycles whatever its kind (profiles/r1_mma_probe.txt), so the instruction count is the
// resource and an FP16 MMA covers K = 16 where a TF32 MMA covers K = 8: GEMM1 8 MMAs per layer (2 main + 4
// correction over x(0..31), 2 over the new columns + bias), GEMM2 6 per K = 32 quarter.
// What remains against a float32 FMA chain is the tensor cores' truncating accumulator
// (profiles/r1_tensor_accum_rounding.txt), see DESIGN.md 4.1.
//
// GEMM1 without re-staging activations.  The flow permutes by one position per layer
// (harmonic/flows.py:65-66), so layer l conditions on ORIGINAL features l .. l+31: the features below
// 32 are still the raw (standardised) inputs, and the j < l features 32+j were frozen by layer j's
// epilogue.  The A operand is therefore written ONCE per tile -- the raw x(0..31) -- plus one new
// column per layer; each layer's W0 is stored shifted by l rows ("main", K = 32) next to a K = 8
// "correction" block whose rows multiply the new columns and a constant-one column that carries the
// bias b0.  The main part of layer l+1 is issued as soon as layer l's last GEMM2 quarter has drained,
// i.e. it runs under layer l's epilogue; only the 2-MMA correction waits for the new column.
//
// GEMM2 is pipelined with E1 at quarter granularity: quarter q's operands replace its own 32 H columns, its
// GEMM2 step runs while the epilogue computes quarter q + 1 (one mbarrier per quarter: the epilogue may run a
// whole layer ahead of the MMA warp, never two phases of one barrier).
// TMEM, P = 128 columns per tile (H, then per 16-column chunk 8 columns of packed lo pairs + 8 of packed hs pairs):
//   NT = 3: tile t at column 160 t: P | O (32).  Scaled layers run GEMM2 twice over the same h operands: the scale
//           rows first, then -- once the epilogue has read the pre-activations out of O (OCONS) -- the shift rows
//           (OFULL2), which so run under the softplus / reciprocal / log-det arithmetic of the epilogue
//   NT = 2: tile t at column 256 t: P | 64 unused | O (64: one GEMM2 pass, N = 64 on scaled layers).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "hb_internal.h"

namespace hb {

namespace tc {

constexpr int D = 64;
constexpr int DH = 32;        // d = conditioning width (round(D/2))
constexpr int U = 32;         // transformed width
constexpr int TILE = 128;
constexpr int SLOT_BYTES = 16384;
constexpr int MAX_TC_LAYERS = 8;  // one K = 8 correction step holds 7 new columns + the bias column
constexpr int KA = 40;            // A operand columns: 32 raw + 7 new + 1 constant
constexpr int A_HALF = 0;         // (offset of the FP16 block inside a tile's A operand: it is the whole operand)
constexpr int A_TILE = KA / 4 * 2048;  // FP16 block, 10 chunks of 8 columns: lo raw 0-3 | hs raw 4-7 | lo new 8 | hs new 9

// Layout for NT tiles (= epilogue warpgroups) in flight per CTA.
//   NT = 2: TMEM per tile P 128 | 64 unused | O 64; y[64] per thread with a register rotation
//   NT = 3: TMEM per tile P 128 | O 32 (480 of 512 columns): scaled layers run GEMM2 in TWO N = 32 passes over the
//           same h operands (scale, then shift, see the kernel); 128 registers per thread, so the
//           epilogue keeps only the 32 transformed features + the 7 that can enter (owned by ORIGINAL index)
template <int NT>
struct L {
  static constexpr int NSLOTS = 8;
  static constexpr int THREADS = NT * 128 + 32 * (NT + 1);  // + producer warp + one MMA warp per tile
  // shared memory carve-up (bytes)
  static constexpr int SM_RING = 0;
  static constexpr int SM_A1 = SM_RING + NSLOTS * SLOT_BYTES;       // per tile: the FP16 A operand (A_TILE)
  static constexpr int SM_BIAS = SM_A1 + NT * A_TILE;               // b2: MAX_TC_LAYERS * 64 floats
  static constexpr int SM_PRE = SM_BIAS + MAX_TC_LAYERS * 64 * 4;   // standardisation: inv_amp[D], -offset*inv_amp[D]
  static constexpr int SM_BAR = SM_PRE + 2 * D * 4;                 // mbarriers
  static constexpr int SM_MISC = SM_BAR + 512;                      // tmem ptr, fold scratch
  static constexpr int SM_TOTAL = SM_MISC + 512;
  // barrier indices (8 bytes each); per-tile barriers are indexed + tile
  static constexpr int B_WFULL = 0, B_WEMPTY = 8, B_A1FULL = 16, B_HFULL = B_A1FULL + NT,
                       B_A2 = B_HFULL + NT /* [quarter][tile], one phase per layer */, B_OFULL = B_A2 + 4 * NT,
                       B_OCONS = B_OFULL + NT /* NT = 3: first pass read, O may be overwritten */,
                       B_OFULL2 = B_OCONS + NT /* NT = 3: second (shift) pass complete */, B_COUNT = B_OFULL2 + NT;
  static_assert(B_COUNT * 8 <= 512, "barrier block");
  // TMEM columns (per tile: + TM_TILE * tile)
  static constexpr uint32_t TM_P = 0, TM_O = NT == 3 ? 128 : 192, TM_TILE = NT == 3 ? 160 : 256;
};

struct Params {
  const uint8_t* image;   // weight image: chunks of SLOT_BYTES in consumption order
  const float* bias;      // per layer 64 floats: bt[32], bs[32]
  const float* pre;       // pre_offset[D], pre_amp[D] or nullptr
  int nlayers;
  int nscaled;
  int chunks_per_batch;
  float T, sum_log_amp;
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
__device__ __forceinline__ void mma_ss_f16(uint32_t d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void mma_ts_f16(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// {upper 16 bits: fp16(hi_elem), lower 16 bits: fp16(lo_elem)}, round to nearest, saturating
__device__ __forceinline__ uint32_t pack_f16x2(float hi_elem, float lo_elem) {
  uint32_t d;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi_elem), "f"(lo_elem));
  return d;
}
__device__ __forceinline__ uint16_t to_f16(float v) {
  uint16_t d;
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(d) : "f"(v));
  return d;
}
__device__ __forceinline__ float from_f16(uint16_t h) {
  float d;
  asm("cvt.f32.f16 %0, %1;" : "=f"(d) : "h"(h));
  return d;
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
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}
// wait::ld that also names the registers of the load it completes: a prefetched chunk's values cannot be
// consumed (or moved by the compiler) ahead of the wait
__device__ __forceinline__ void tmem_ld_wait16(uint32_t (&v)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                 "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15])
               :
               : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]),
               "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
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

__device__ __forceinline__ uint32_t make_idesc_f16(int n) {
  // c=F32 (1<<4), a=b=F16 (0), K-major both, N>>3 at bit 17, M>>4 at bit 24
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
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

// Both GEMMs run entirely as kind::f16: one FP16 MMA covers K = 16 at the same ~45-cycle cost as a TF32 MMA
// covering K = 8 (profiles/r1_mma_probe.txt), and an 11-bit significand is all the main term needs.
// Compensated split of an activation h (GEMM1: a standardised input x):  hs = fp16(h * 2^-11) (round to nearest),  lo = h - 2^11 * float(hs)  (exact in
// float32, whatever hs lost to rounding or to the half subnormal range), stored as fp16(lo).  Then
//   h . w  ~=  hs . (w_hi * 2^11)  +  hs . (w_lo * 2^11)  +  lo . w_hi       (the lo . w_lo term, <= 2^-22, dropped)
// with all products exact in the tensor core.  hs is a normal half for |h| in [0.125, 1.3e8]; below that lo
// carries the remainder (|lo| <= 6e-5, absolute error <= 3e-8).  Beyond 1.3e8 the halves saturate.
constexpr float H2_SCALE = 4.8828125e-4f;  // 2^-11
constexpr float H2_INV = 2048.0f;
__device__ __forceinline__ float2 unpack_f16x2(uint32_t v) {  // {low half, high half}
  float2 r;
  asm("{.reg .b16 a, b; mov.b32 {a, b}, %2; cvt.f32.f16 %0, a; cvt.f32.f16 %1, b;}" : "=f"(r.x), "=f"(r.y) : "r"(v));
  return r;
}

// E1 for one 16-column chunk: h = leaky_relu(H) -> wh: 8 words of packed hs pairs, wl: 8 words of packed lo pairs
// (the compensated split above).  The bias b0 is already inside H (constant-one column of the A operand).
__device__ __forceinline__ void e1_chunk(const uint32_t (&v)[16], uint32_t (&wl)[8], uint32_t (&wh)[8]) {
#pragma unroll
  for (int jj = 0; jj < 8; ++jj) {
    float h0 = __uint_as_float(v[2 * jj]), h1 = __uint_as_float(v[2 * jj + 1]);
    h0 = fmaxf(h0, 0.01f * h0);
    h1 = fmaxf(h1, 0.01f * h1);
    const uint32_t hs = pack_f16x2(h1 * H2_SCALE, h0 * H2_SCALE);
    const float2 u = unpack_f16x2(hs);
    wh[jj] = hs;
    wl[jj] = pack_f16x2(fmaf(u.y, -H2_INV, h1), fmaf(u.x, -H2_INV, h0));
  }
}

__device__ __forceinline__ int layer_chunks(bool scaled) { return scaled ? 7 : 5; }

template <typename XT, typename LT, bool FOLD, int NT>
__global__ void __launch_bounds__(L<NT>::THREADS, 1)
realnvp_tc_kernel(Params P, const XT* __restrict__ x, long long ld, const LT* __restrict__ lnpost,
                  float* __restrict__ lnpred_out, FoldParams fp) {
  extern __shared__ __align__(1024) uint8_t smem[];
  using C = L<NT>;
  constexpr int NWG = NT, THREADS = C::THREADS, NSLOTS = C::NSLOTS;
  constexpr int SM_RING = C::SM_RING, SM_A1 = C::SM_A1, SM_BIAS = C::SM_BIAS, SM_PRE = C::SM_PRE, SM_BAR = C::SM_BAR,
                SM_MISC = C::SM_MISC;
  constexpr int B_WFULL = C::B_WFULL, B_WEMPTY = C::B_WEMPTY, B_A1FULL = C::B_A1FULL, B_HFULL = C::B_HFULL,
                B_A2 = C::B_A2, B_OFULL = C::B_OFULL, B_OCONS = C::B_OCONS, B_OFULL2 = C::B_OFULL2;
  constexpr uint32_t TM_P = C::TM_P, TM_O = C::TM_O, TM_TILE = C::TM_TILE;
  constexpr int W_PROD = 4 * NT, W_MMA = W_PROD + 1;  // weight producer warp; the MMA warps follow
  (void)B_OCONS; (void)B_OFULL2;
  // lane-0 broadcast: tells the compiler the warp index (and with it the role dispatch below) is
  // warp-uniform, so the MMA warps' addresses, descriptors and loop state live on the uniform datapath
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar0 = sbase + SM_BAR;
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(smem + SM_MISC);
  float* bias_s = reinterpret_cast<float*>(smem + SM_BIAS);
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
      mbar_init(BAR(B_WEMPTY + i), NWG);  // both MMA warps release a slot
    }
    for (int t = 0; t < NWG; ++t) {
      mbar_init(BAR(B_A1FULL + t), 128);
      mbar_init(BAR(B_HFULL + t), 1);
      for (int q = 0; q < 4; ++q) mbar_init(BAR(B_A2 + NT * q + t), 128);
      mbar_init(BAR(B_OFULL + t), 1);
      mbar_init(BAR(B_OCONS + t), 128);
      mbar_init(BAR(B_OFULL2 + t), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < nl * 64; i += THREADS) bias_s[i] = P.bias[i];
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

  if (warp == W_PROD) {
    // ===== weight producer: the image is one batch worth of chunks, replayed every batch =====
    if (lane == 0) {
      uint32_t g = 0;
      for (int b = 0; b < P.nbatches; ++b) {
        const uint8_t* src = P.image;
        for (int l = 0; l < nl; ++l) {
          const int nc = layer_chunks(l < P.nscaled);
          for (int j = 0; j < nc; ++j, ++g) {
            // only the occupied prefix of a chunk travels: GEMM1 main hs block 8 KiB, [lo | hs] block 16 KiB, the
            // two new-column blocks 8 KiB, a GEMM2 chunk (one N2 = 64 quarter or two N2 = 32 quarters) 12 KiB
            const uint32_t bytes = j == 1 ? 16384u : (j < 3 ? 8192u : 12288u);
            const uint32_t slot = g % NSLOTS, use = g / NSLOTS;
            mbar_wait(BAR(B_WEMPTY + slot), (use & 1) ^ 1);
            mbar_expect_tx(BAR(B_WFULL + slot), bytes);
            bulk_g2s(sbase + SM_RING + slot * SLOT_BYTES, src, bytes, BAR(B_WFULL + slot));
            src += SLOT_BYTES;
          }
        }
      }
    }
  } else if (warp >= W_MMA) {
    // ===== MMA issuers: warp 9 serves tile 0, warp 10 tile 1.  Each runs a straight-line program
    // per layer with blocking mbarrier waits; the whole warp executes it (warp-uniform, so the
    // descriptors stay in uniform registers) and one elected lane issues.  A weight slot is
    // released when both warps have committed on it (WEMPTY expects two arrivals).
    const int t = warp - W_MMA;
    const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0) + t * TM_TILE;
    const uint32_t ih128 = make_idesc_f16(128), ih64 = make_idesc_f16(64), ih32 = make_idesc_f16(32);
    const uint32_t a_h = sbase + SM_A1 + t * A_TILE;  // FP16 block: lo raw 0-3 | hs raw 4-7 | lo new 8 | hs new 9
    const uint64_t adx_f16 = make_desc(a_h, 2048, 128), adn_f16 = make_desc(a_h + 8 * 2048, 2048, 128);
    auto wwait = [&](uint32_t gidx) {
      mbar_wait(BAR(B_WFULL + (gidx % NSLOTS)), (gidx / NSLOTS) & 1);
    };
    auto chunk_addr = [&](uint32_t gidx) { return sbase + SM_RING + (gidx % NSLOTS) * SLOT_BYTES; };
    auto release = [&](uint32_t gidx) {
      if (elect_one()) tc_commit(BAR(B_WEMPTY + (gidx % NSLOTS)));
      __syncwarp();
    };
    // main part of GEMM1: H = x(0..31) . W0main, three FP16 products on the compensated split of x (as GEMM2).
    // Chunk g+1 = FP16 [128 x 64 K']: w_hi (meets lo) | w_lo * 2^11 (meets hs); chunk g = FP16 [128 x 32 K']:
    // w_hi * 2^11 (meets the same hs block); corrections first (small terms), then the main term.
    auto issue_main = [&](uint32_t g) {
      wwait(g);
      wwait(g + 1);
      tc_fence_after();
      const uint64_t wt = make_desc(chunk_addr(g), 2048, 128), wh = make_desc(chunk_addr(g + 1), 2048, 128);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          mma_ss_f16(tb + TM_P, adx_f16 + (uint64_t)(k * 256), wh + (uint64_t)(k * 256), ih128, k ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 2; ++k)
          mma_ss_f16(tb + TM_P, adx_f16 + (uint64_t)((2 + k) * 256), wt + (uint64_t)(k * 256), ih128, 1);
      }
      __syncwarp();
    };
    uint32_t g = 0, it = 0, sc_it = 0;
    (void)sc_it;
    for (int b = 0; b < P.nbatches; ++b) {
      for (int l = 0; l < nl; ++l, ++it) {
        const bool scaled = l < P.nscaled;
        const uint32_t par = it & 1;
        // ---- GEMM1: (main, unless issued early) + correction
        mbar_wait(BAR(B_A1FULL + t), par);
        TR(20);
        if (l == 0) issue_main(g);
        wwait(g + 2);
        tc_fence_after();
        {
          const uint64_t wc = make_desc(chunk_addr(g + 2), 2048, 128);
          if (elect_one()) {
            // K' = 16 each over A = [lo new | hs new]: B = [w_hi ; w_lo * 2^11], then B = [0 ; w_hi * 2^11]
            mma_ss_f16(tb + TM_P, adn_f16, wc + (uint64_t)(4096 >> 4), ih128, 1);
            mma_ss_f16(tb + TM_P, adn_f16, wc, ih128, 1);
            tc_commit(BAR(B_HFULL + t));
            tc_commit(BAR(B_WEMPTY + (g % NSLOTS)));
            tc_commit(BAR(B_WEMPTY + ((g + 1) % NSLOTS)));
            tc_commit(BAR(B_WEMPTY + ((g + 2) % NSLOTS)));
          }
          __syncwarp();
          TR(21);
        }
        // ---- GEMM2: four K = 32 quarters, each as soon as its half of E1 has landed
        const uint32_t n2 = scaled ? 64u : 32u;
        const uint32_t lbo2 = n2 * 16u;
        const uint32_t qbytes = 12 * lbo2;  // one quarter's weight block: 12 K'-chunks of 8 halves
        const uint64_t kstep2 = (uint64_t)((2 * lbo2) >> 4);
        // NT = 3: O holds 32 columns, scaled layers take the SCALE rows (32..63 of each quarter block, + 4 row
        // groups of 128 bytes) in this pass: their epilogue is the long one, the shift pass runs under it
        const bool two_pass = NT == 3 && scaled;
        const uint32_t ih2 = (scaled && NT != 3) ? ih64 : ih32;
#pragma unroll 1
        for (int q = 0; q < 4; ++q) {
          const uint32_t gw = g + 3 + (scaled ? q : (q >> 1));
          // (one barrier per quarter: with two phases per layer on one barrier the epilogue, no longer held back by
          // a Q-buffer hand-back, could complete both before this warp has seen the first)
          mbar_wait(BAR(B_A2 + NT * q + t), par);
          TR(22 + q);
          wwait(gw);
          tc_fence_after();
          // quarter block in the chunk: FP16 [N2 x 96 K']: K' 0-31 w_hi (meets lo), 32-63 w_lo * 2^11 and
          // 64-95 w_hi * 2^11 (both meet the scaled hi block).  Chunk c = 2 q + half holds its operands in
          // place of its 16 H columns: 8 columns of packed lo pairs, 8 of packed scaled hi pairs.
          const uint32_t wbase = chunk_addr(gw) + ((!scaled && (q & 1)) ? qbytes : 0) + (two_pass ? 512u : 0u);
          const uint64_t bd = make_desc(wbase, lbo2, 128);
          const uint32_t a0 = tb + TM_P + 32 * q;
          if (elect_one()) {
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              const uint32_t a_lo = a0 + 16 * half, a_hs = a_lo + 8;
              mma_ts_f16(tb + TM_O, a_lo, bd + (uint64_t)half * kstep2, ih2, (q | half) ? 1u : 0u);
              mma_ts_f16(tb + TM_O, a_hs, bd + (uint64_t)(2 + half) * kstep2, ih2, 1);
              mma_ts_f16(tb + TM_O, a_hs, bd + (uint64_t)(4 + half) * kstep2, ih2, 1);
            }
            if (q == 3) tc_commit(BAR(B_OFULL + t));
            if (!two_pass && (scaled || (q & 1))) tc_commit(BAR(B_WEMPTY + (gw % NSLOTS)));
          }
          __syncwarp();
          TR(26 + q);
        }
        if constexpr (NT == 3) {
          if (scaled) {
            // ---- second pass: the shift rows (0..31 of each quarter block) over the same h operands, into the
            // same 32 O columns once the epilogue has read the scale pre-activations out of them
            mbar_wait(BAR(B_OCONS + t), sc_it & 1);
            tc_fence_after();
            TR(32);
            if (elect_one()) {
#pragma unroll 1
              for (int q = 0; q < 4; ++q) {
                const uint64_t bd = make_desc(chunk_addr(g + 3 + q), lbo2, 128);
                const uint32_t a0 = tb + TM_P + 32 * q;
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                  const uint32_t a_lo = a0 + 16 * half, a_hs = a_lo + 8;
                  mma_ts_f16(tb + TM_O, a_lo, bd + (uint64_t)half * kstep2, ih32, (q | half) ? 1u : 0u);
                  mma_ts_f16(tb + TM_O, a_hs, bd + (uint64_t)(2 + half) * kstep2, ih32, 1);
                  mma_ts_f16(tb + TM_O, a_hs, bd + (uint64_t)(4 + half) * kstep2, ih32, 1);
                }
                tc_commit(BAR(B_WEMPTY + ((g + 3 + q) % NSLOTS)));
              }
              tc_commit(BAR(B_OFULL2 + t));
            }
            __syncwarp();
            TR(33);
          }
        }
        g += layer_chunks(scaled);
        // ---- early main part of the next layer: runs under this layer's E2.  P is GEMM2's A operand
        // until the last quarter has drained.
        if (l + 1 < nl) {
          if (two_pass) mbar_wait(BAR(B_OFULL2 + t), sc_it & 1);
          else mbar_wait(BAR(B_OFULL + t), par);
          TR(30);
          issue_main(g);
          TR(31);
        }
        if (two_pass) ++sc_it;
      }
    }
  } else if constexpr (NT == 2) {
    // ===== epilogue warpgroups =====
    const int wg = warp >> 2;              // 0 / 1
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
      TR(1);
      load_row<XT>(x, ld, rr, y);
      if (r + TILE < row_hi) {  // pull the next tile's row into L2 while this tile computes
        const char* nx = reinterpret_cast<const char*>(x + (r + TILE) * ld);
#pragma unroll
        for (int q = 0; q < (int)(D * sizeof(XT)); q += 128) prefetch_l2(nx + q);
      }
#pragma unroll
      for (int i = 0; i < D; ++i) y[i] = fmaf(y[i], pre_s[i], pre_s[D + i]);
      // ---- A operand, written once per tile.  TF32 block: hi of raw features 0..31 (K-chunks 0-7),
      // zeroed new columns and a constant one in column 39 (K-chunks 8, 9).  FP16 block (8 values per
      // 16-byte chunk): lo of raw 0..31 (chunks 0-3), hi of raw (4-7), lo of new (8), hi of new (9).
      // (GEMM1 of the previous tile has completed: its HFULL was waited.)
#pragma unroll
      for (int c = 0; c < DH / 8; ++c) {
        uint4 plo, phi;
        uint32_t* ql = &plo.x;
        uint32_t* qh = &phi.x;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float va = y[8 * c + 2 * q], vb = y[8 * c + 2 * q + 1];
          const uint32_t hs = pack_f16x2(vb * H2_SCALE, va * H2_SCALE);
          const float2 bk = unpack_f16x2(hs);
          ql[q] = pack_f16x2(fmaf(bk.y, -H2_INV, vb), fmaf(bk.x, -H2_INV, va));
          qh[q] = hs;
        }
        *reinterpret_cast<uint4*>(a1 + A_HALF + c * 2048 + tid * 16) = plo;
        *reinterpret_cast<uint4*>(a1 + A_HALF + (4 + c) * 2048 + tid * 16) = phi;
      }
      {
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(a1 + A_HALF + 8 * 2048 + tid * 16) = z;
        *reinterpret_cast<uint4*>(a1 + A_HALF + 9 * 2048 + tid * 16) = make_uint4(0u, 0u, 0u, 0x10000000u);
      }
      float ldj = 0.f;
      for (int l = 0; l < nl; ++l, ++it) {
        const uint32_t par = it & 1;
        const bool scaled = l < P.nscaled;
        const float* b2 = bias_s + l * 64;
        fence_proxy_async();
        tc_fence_before();  // orders the previous layer's TMEM loads (O) before the next MMAs
        mbar_arrive(BAR(B_A1FULL + wg));
        TR(2);
        // ---- E1 in eight 16-column chunks, software pipelined: chunk c + 1 is in flight from TMEM while
        // chunk c is computed.  Chunks 2q, 2q + 1 form quarter q: its FP16 operands replace its H columns and
        // its GEMM2 step runs under the following chunks.
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
            tmem_st8(tm_p + c * 16, wl);      // in place of the chunk's own H columns (already in registers)
            tmem_st8(tm_p + c * 16 + 8, wh);
            if (c & 1) {
              tmem_st_wait();
              tc_fence_before();
              mbar_arrive(BAR(B_A2 + NT * q + wg));
              TR(4 + q);
            }
          }
        }
        // ---- E2: y1 = (y1 - shift) / scale, ldj -= sum log scale
        mbar_wait(BAR(B_OFULL + wg), par);
        tc_fence_after();
        TR(9);
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
        if (l + 1 < nl) {
          // feature (logical 32) is frozen from here on: it is the new conditioning column l of the
          // following layers
          const float v = y[DH];
          const uint16_t hs16 = to_f16(v * H2_SCALE);
          *reinterpret_cast<uint16_t*>(a1 + A_HALF + 8 * 2048 + tid * 16 + l * 2) = to_f16(fmaf(from_f16(hs16), -H2_INV, v));
          *reinterpret_cast<uint16_t*>(a1 + A_HALF + 9 * 2048 + tid * 16 + l * 2) = hs16;
          // inverse of tfb.Permute([D-1, 0, .., D-2]): rotate left by one (flows.py:65-66)
          const float y0 = y[0];
#pragma unroll
          for (int i = 0; i < D - 1; ++i) y[i] = y[i + 1];
          y[D - 1] = y0;
        }
        TR(10);
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
  } else {
    // ===== epilogue warpgroups, NT = 3 =====
    // 128 registers per thread: features are owned by ORIGINAL index and nothing rotates.  yt[i] = feature 32 + i,
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
    uint32_t it = 0, sc_it = 0;

    for (int b = 0; b < P.nbatches; ++b) {
      const long long t0 = row_lo + (long long)b * TILE;
      const long long r = t0 + tid;
      const bool live = r < row_hi;
      const long long rr = live ? r : 0;
      float yt[U], x7[NX], ss = 0.f;
      TR(1);
      {
        float y[D];
        load_row<XT>(x, ld, rr, y);
        if (r + TILE < row_hi) {  // pull the next tile's row into L2 while this tile computes
          const char* nx = reinterpret_cast<const char*>(x + (r + TILE) * ld);
#pragma unroll
          for (int q = 0; q < (int)(D * sizeof(XT)); q += 128) prefetch_l2(nx + q);
        }
#pragma unroll
        for (int i = 0; i < D; ++i) y[i] = fmaf(y[i], pre_s[i], pre_s[D + i]);
        // A operand, written once per tile (layout as in the NT = 2 branch)
#pragma unroll
        for (int c = 0; c < DH / 8; ++c) {
          uint4 plo, phi;
          uint32_t* ql = &plo.x;
          uint32_t* qh = &phi.x;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float va = y[8 * c + 2 * q], vb = y[8 * c + 2 * q + 1];
            const uint32_t hs = pack_f16x2(vb * H2_SCALE, va * H2_SCALE);
            const float2 bk = unpack_f16x2(hs);
            ql[q] = pack_f16x2(fmaf(bk.y, -H2_INV, vb), fmaf(bk.x, -H2_INV, va));
            qh[q] = hs;
          }
          *reinterpret_cast<uint4*>(a1 + A_HALF + c * 2048 + tid * 16) = plo;
          *reinterpret_cast<uint4*>(a1 + A_HALF + (4 + c) * 2048 + tid * 16) = phi;
        }
        {
          const uint4 z = make_uint4(0u, 0u, 0u, 0u);
          *reinterpret_cast<uint4*>(a1 + A_HALF + 8 * 2048 + tid * 16) = z;
          *reinterpret_cast<uint4*>(a1 + A_HALF + 9 * 2048 + tid * 16) = make_uint4(0u, 0u, 0u, 0x10000000u);
        }
#pragma unroll
        for (int i = 0; i < NX; ++i) x7[i] = y[i];
#pragma unroll
        for (int i = NX; i < DH; ++i) {
          const float z = y[i] * invT;
          ss = fmaf(z, z, ss);
        }
#pragma unroll
        for (int i = 0; i < U; ++i) yt[i] = y[DH + i];
      }
      float ldj = 0.f;
      for (int l = 0; l < nl; ++l, ++it) {
        const uint32_t par = it & 1;
        const bool scaled = l < P.nscaled;
        const float* bl = bias_s + l * 64 - l;  // bl[i] = bt[i - l], bl[32 + i] = bt[32 - l + i] = bs[i - l], bl[64 + i] = bs[32 - l + i]
        fence_proxy_async();
        tc_fence_before();  // orders the previous layer's TMEM loads (O) before the next MMAs
        mbar_arrive(BAR(B_A1FULL + wg));
        TR(2);
        // ---- E1 (as in the NT = 2 branch)
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
        // ---- E2
        mbar_wait(BAR(B_OFULL + wg), par);
        tc_fence_after();
        TR(9);
        if (!scaled) {
          uint32_t v0[U], v1[8];
          tmem_ld32(tm_o - l, v0);
          tmem_ld8(tm_o + U - l, v1);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < U; ++i) {
            const float shift = __uint_as_float(v0[i]) + bl[i];
            yt[i] -= (i >= NX || i >= l) ? shift : 0.f;
          }
#pragma unroll
          for (int i = 0; i < NX; ++i) {
            const float shift = __uint_as_float(v1[i]) + bl[U + i];
            x7[i] -= (i < l) ? shift : 0.f;
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
          mbar_arrive(BAR(B_OCONS + wg));
          // v0 / v1 become 1 / clip(softplus(a)) in place; ldj -= sum log scale
#pragma unroll
          for (int k0 = 0; k0 < U; k0 += 8) {
            float prod = 1.f;  // 8 factors in [1e-3, 1e3] stay inside the float32 range
#pragma unroll
            for (int i = k0; i < k0 + 8; ++i) {
              const bool act = i >= NX || i >= l;
              const float a = __uint_as_float(v0[i]) + bl[U + i];
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
                const float a = __uint_as_float(v1[i]) + bl[2 * U + i];
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
          mbar_wait(BAR(B_OFULL2 + wg), sc_it & 1);
          tc_fence_after();
          TR(12);
          ++sc_it;
          uint32_t w0[16], w1[8];
          tmem_ld16(tm_o - l, w0);
          tmem_ld8(tm_o + U - l, w1);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float yn = (yt[i] - (__uint_as_float(w0[i]) + bl[i])) * __uint_as_float(v0[i]);
            yt[i] = (i >= NX || i >= l) ? yn : yt[i];
          }
          tmem_ld16(tm_o - l + 16, w0);
#pragma unroll
          for (int i = 0; i < NX; ++i) {
            const float xn = (x7[i] - (__uint_as_float(w1[i]) + bl[U + i])) * __uint_as_float(v1[i]);
            x7[i] = (i < l) ? xn : x7[i];
          }
          tmem_ld_wait16(w0);
#pragma unroll
          for (int i = 16; i < U; ++i) {
            const float yn = (yt[i] - (__uint_as_float(w0[i - 16]) + bl[i])) * __uint_as_float(v0[i]);
            yt[i] = yn;  // features 48.. are transformed by every layer
          }
        }
        if (l + 1 < nl) {
          // feature 32 + l is final from here on: it is the new conditioning column l of the following layers
          float v = yt[0];
#pragma unroll
          for (int i = 1; i < NX; ++i) v = (l == i) ? yt[i] : v;
          const uint16_t hs16 = to_f16(v * H2_SCALE);
          *reinterpret_cast<uint16_t*>(a1 + A_HALF + 8 * 2048 + tid * 16 + l * 2) = to_f16(fmaf(from_f16(hs16), -H2_INV, v));
          *reinterpret_cast<uint16_t*>(a1 + A_HALF + 9 * 2048 + tid * 16 + l * 2) = hs16;
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
// host side: weight image (hi/lo split, canonical K-major no-swizzle layout, 16 KiB chunks)
// ---------------------------------------------------------------------------------------------
static inline void split_tf32(float v, float* hi, float* lo) {
  uint32_t b;
  memcpy(&b, &v, 4);
  b &= 0xffffe000u;
  memcpy(hi, &b, 4);
  *lo = v - *hi;
}

// element (n, k) of a K-major [N x K] 16-bit operand in the canonical no-swizzle layout (8 elements per 16-byte
// chunk), in halves
static inline size_t kmajor_off16(int N, int n, int k) {
  return (size_t)(n / 8) * 64 + (size_t)(n % 8) * 8 + (size_t)(k / 8) * (N * 8) + (k % 8);
}

// float -> IEEE half, round to nearest even, saturating to +-65504 (NaN stays NaN)
static inline uint16_t f32_to_f16(float f) {
  uint32_t x;
  memcpy(&x, &f, 4);
  const uint32_t sign = (x >> 16) & 0x8000u;
  const uint32_t absx = x & 0x7fffffffu;
  if (absx > 0x7f800000u) return (uint16_t)(sign | 0x7e00u);
  if (absx >= 0x477ff000u) return (uint16_t)(sign | 0x7bffu);  // >= 65520 rounds past the largest half
  if (absx < 0x33000001u) return (uint16_t)sign;                // < 2^-25 (rounds to zero)
  const int e = (int)(absx >> 23) - 127;
  uint32_t mant = (absx & 0x7fffffu) | 0x800000u;
  if (e < -14) {  // subnormal half: value = mant * 2^(e-23), unit = 2^-24
    const int shift = -e - 1;  // 13 + (-14 - e)
    const uint32_t q = mant >> shift, rem = mant & ((1u << shift) - 1), half = 1u << (shift - 1);
    uint32_t r = q;
    if (rem > half || (rem == half && (q & 1))) ++r;
    return (uint16_t)(sign | r);
  }
  uint32_t q = mant >> 13, rem = mant & 0x1fffu;
  if (rem > 0x1000u || (rem == 0x1000u && (q & 1))) ++q;
  uint32_t he = (uint32_t)(e + 15);
  if (q & 0x800u) {  // mantissa overflowed to 2.0
    q >>= 1;
    ++he;
  }
  return (uint16_t)(sign | (he << 10) | (q & 0x3ffu));
}

// one weight w = hi + lo (hi: 11 significant bits) into the three FP16 operands: m16[off_main] = fp16(hi * 2^11)
// (main term, meets hs), h16[off_hi] = fp16(hi) (meets lo), h16[off_lo] = fp16(lo * 2^11) (meets hs)
static inline void put_weight(float w, uint16_t* m16, size_t off_main, uint16_t* h16, size_t off_hi, size_t off_lo) {
  float hi, lo;
  split_tf32(w, &hi, &lo);
  m16[off_main] = f32_to_f16(hi * 2048.0f);
  h16[off_hi] = f32_to_f16(hi);
  h16[off_lo] = f32_to_f16(lo * 2048.0f);
}

int realnvp_tc_prepare(Flow* f) {
  f->tc_ok = false;
  const hb_flow_desc& ds = f->desc;
  if (ds.kind != HB_FLOW_REALNVP || ds.ndim != tc::D) return HB_OK;
  const int nl = f->nlayers;
  if (nl > tc::MAX_TC_LAYERS) return HB_OK;
  const int d = f->d, u = f->u;
  // weights (and b0, which rides in W0') are stored * 2^11 as halves: |w| must stay below 65504 / 2048 (else the
  // CUDA-core path runs)
  for (int l = 0; l < nl; ++l) {
    const float* W0 = f->host_weights.data() + f->layer_off[l];
    const size_t n2 = (size_t)d * 128 + 128 + (size_t)128 * u * ((l < ds.n_scaled) ? 2 : 1) + u;  // W0, b0, Wt, bt, (Ws)
    for (size_t i = 0; i < n2; ++i)
      if (!(fabsf(W0[i]) < 16.0f)) return HB_OK;
  }
  const size_t CF = tc::SLOT_BYTES / 4;  // floats per chunk
  size_t nchunks = 0;
  for (int l = 0; l < nl; ++l) nchunks += (l < ds.n_scaled) ? 7 : 5;
  std::vector<float> img(nchunks * CF, 0.f);
  std::vector<float> bias((size_t)nl * 64, 0.f);
  const float* w = f->host_weights.data();
  size_t ch = 0;
  for (int l = 0; l < nl; ++l) {
    const bool scaled = l < ds.n_scaled;
    const float* W0 = w + f->layer_off[l];       // [d][128]
    const float* b0 = W0 + (size_t)d * 128;
    const float* Wt = b0 + 128;                  // [128][u]
    const float* bt = Wt + (size_t)128 * u;
    const float* Ws = bt + u;
    const float* bs = Ws + (size_t)128 * u;
    // GEMM1 main: A column i (original feature i < 32) meets W0 row i - l (zero for i < l).
    // chunk 0: FP16 [128 x 32 K'] fp16(hi * 2^11); chunk 1: FP16 [128 x 64 K']: K' < 32 fp16(hi), K' >= 32
    // fp16(lo * 2^11)
    {
      uint16_t* c0 = reinterpret_cast<uint16_t*>(img.data() + ch * CF);
      uint16_t* c1 = reinterpret_cast<uint16_t*>(img.data() + (ch + 1) * CF);
      for (int i = 0; i < 32; ++i)
        for (int n = 0; n < 128; ++n) {
          const float v = (i >= l) ? W0[(size_t)(i - l) * 128 + n] : 0.f;
          put_weight(v, c0, kmajor_off16(128, n, i), c1, kmajor_off16(128, n, i), kmajor_off16(128, n, 32 + i));
        }
    }
    // GEMM1 correction (chunk 2): k = 0..6 <-> frozen feature 32 + j meets W0 row 32 - l + j (j < l),
    // k = 7 <-> constant-one column carrying the bias b0.  FP16 [128 x 16 K'] 4 KB: K' < 8 zero (meets lo new),
    // K' >= 8 fp16(hi * 2^11) (meets hs new) | FP16 [128 x 16 K'] 4 KB: fp16(hi) | fp16(lo * 2^11)
    {
      uint16_t* c = reinterpret_cast<uint16_t*>(img.data() + (ch + 2) * CF);
      uint16_t* ch16 = c + 2048;
      for (int k = 0; k < 8; ++k)
        for (int n = 0; n < 128; ++n) {
          float v = 0.f;
          if (k == 7) v = b0[n];
          else if (k < l) v = W0[(size_t)(32 - l + k) * 128 + n];
          put_weight(v, c, kmajor_off16(128, n, 8 + k), ch16, kmajor_off16(128, n, k), kmajor_off16(128, n, 8 + k));
        }
    }
    // GEMM2: B[n][k] = [Wt | Ws][k][n]; one K = 32 quarter = FP16 block [N2 x 96 K']: K' < 32 fp16(w_hi) (meets
    // lo), 32..63 fp16(w_lo * 2^11), 64..95 fp16(w_hi * 2^11) (both meet hs = fp16(h * 2^-11)).  12 K'-chunks of 8 halves:
    // 192 N2 bytes per quarter, one quarter per 16 KiB slot for scaled layers, two for unscaled ones.
    const int N2 = scaled ? 64 : 32;
    const size_t qh = (size_t)12 * N2 * 8;  // halves per quarter block
    for (int q = 0; q < 4; ++q) {
      uint16_t* c16 = reinterpret_cast<uint16_t*>(scaled ? img.data() + (ch + 3 + q) * CF
                                                         : img.data() + (ch + 3 + (q >> 1)) * CF) +
                      ((!scaled && (q & 1)) ? qh : 0);
      for (int kk = 0; kk < 32; ++kk)
        for (int n = 0; n < N2; ++n) {
          const int k = 32 * q + kk;
          const float v = (n < u) ? Wt[(size_t)k * u + n] : Ws[(size_t)k * u + (n - u)];
          float hi, lo;
          split_tf32(v, &hi, &lo);
          c16[kmajor_off16(N2, n, kk)] = f32_to_f16(hi);
          c16[kmajor_off16(N2, n, 32 + kk)] = f32_to_f16(lo * 2048.0f);
          c16[kmajor_off16(N2, n, 64 + kk)] = f32_to_f16(hi * 2048.0f);
        }
    }
    ch += scaled ? 7 : 5;
    for (int k = 0; k < u; ++k) bias[(size_t)l * 64 + k] = bt[k];
    if (scaled)
      for (int k = 0; k < u; ++k) bias[(size_t)l * 64 + u + k] = bs[k];
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

// tiles in flight per CTA: HB_TC_TILES = 2 | 3 (A/B measurements)
static int tc_tiles() {
  static const int v = [] {
    const char* e = getenv("HB_TC_TILES");
    return (e && e[0] == '2') ? 2 : 3;
  }();
  return v;
}

template <typename XT, typename LT, bool FOLD>
static int tc_launch(const tc::Params& P, int grid, const void* x, long long ld, const void* lnpost,
                     float* lnpred_out, const FoldParams& fp, cudaStream_t st) {
  if (tc_tiles() == 3) {
    auto kern = tc::realnvp_tc_kernel<XT, LT, FOLD, 3>;
    HB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::L<3>::SM_TOTAL));
    kern<<<grid, tc::L<3>::THREADS, tc::L<3>::SM_TOTAL, st>>>(P, (const XT*)x, ld, (const LT*)lnpost, lnpred_out, fp);
  } else {
    auto kern = tc::realnvp_tc_kernel<XT, LT, FOLD, 2>;
    HB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::L<2>::SM_TOTAL));
    kern<<<grid, tc::L<2>::THREADS, tc::L<2>::SM_TOTAL, st>>>(P, (const XT*)x, ld, (const LT*)lnpost, lnpred_out, fp);
  }
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
  const int NWG = tc_tiles();
  long long nwg = (long long)sms * NWG;
  long long rpw = (n + nwg - 1) / nwg;
  rpw = ((rpw + tc::TILE - 1) / tc::TILE) * tc::TILE;
  nwg = (n + rpw - 1) / rpw;
  const int grid = (int)((nwg + NWG - 1) / NWG);
  const long long nfl = (long long)grid * NWG;
  tc::Params P;
  P.image = static_cast<const uint8_t*>(f->tc_image);
  P.bias = reinterpret_cast<const float*>(static_cast<const char*>(f->tc_image) + f->tc_image_bytes);
  P.pre = f->desc.standardize ? f->dev_weights : nullptr;
  P.nlayers = f->nlayers;
  P.nscaled = f->desc.n_scaled;
  P.T = T;
  P.sum_log_amp = f->sum_log_amp;
  P.chunks_per_batch = (int)(f->tc_image_bytes / tc::SLOT_BYTES);
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
