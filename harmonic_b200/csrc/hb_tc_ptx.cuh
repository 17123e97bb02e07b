// PTX wrappers (mbarrier, cp.async.bulk, tcgen05 MMA / TMEM load / store, conversions) and the host-side helpers of the
// FP16 hi / lo weight images (canonical K-major no-swizzle UMMA layout) shared by the tcgen05 kernels
// (hb_realnvp_tc.cu, hb_rqspline_tc.cu).
#pragma once
#include <stdint.h>
#include <string.h>

namespace hb {
namespace tcx {

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
// arrive that hands back the barrier's state word: a value that data-depends on it cannot be computed before the arrive
__device__ __forceinline__ uint32_t mbar_arrive_token(uint32_t bar) {
  uint64_t st;
  asm volatile("mbarrier.arrive.shared::cta.b64 %0, [%1];" : "=l"(st) : "r"(bar) : "memory");
  return (uint32_t)st;
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
// the producer's wait: its lane would otherwise spend ~8 % of the SM's issue slots polling (profiles/README.md, r1l)
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "nanosleep.u32 500;\n\t"
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
__device__ __forceinline__ void mma_ss_f16(uint32_t d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem]
__device__ __forceinline__ void mma_ts_f16(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// {upper 16 bits: fp16(hi_elem), lower 16 bits: fp16(lo_elem)}, round to nearest (overflow -> inf: visible, and
// excluded by the range guard)
__device__ __forceinline__ uint32_t pack_f16x2(float hi_elem, float lo_elem) {
  uint32_t d;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi_elem), "f"(lo_elem));
  return d;
}
// the same with negative inputs clamped to zero (NaN stays NaN); RZ: round towards zero, so that the remainder of a
// positive value is never negative
__device__ __forceinline__ uint32_t pack_relu_rz_f16x2(float hi_elem, float lo_elem) {
  uint32_t d;
  asm("cvt.rz.relu.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi_elem), "f"(lo_elem));
  return d;
}
__device__ __forceinline__ uint32_t pack_relu_f16x2(float hi_elem, float lo_elem) {
  uint32_t d;
  asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi_elem), "f"(lo_elem));
  return d;
}
__device__ __forceinline__ uint16_t to_f16(float v) {
  uint16_t d;
  asm("cvt.rn.f16.f32 %0, %1;" : "=h"(d) : "f"(v));
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
#define HB_LIST32                                                                                     \
  "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, " \
  "%21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}"
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 " HB_LIST32 ", [%32];"
               : HB_R32(v, 0)
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
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
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]),
               "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// UMMA shared-memory descriptor, K-major, no swizzle: element (r, k) of a [rows x K] 16-bit operand
// lives at (r/8)*SBO + (r%8)*16 + (k/8)*LBO + (k%8)*2 bytes.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) |
         (1ull << 46);
}
__device__ __forceinline__ uint32_t make_idesc_f16(int n) {
  // c=F32 (1<<4), a=b=F16 (0), K-major both, N>>3 at bit 17, M>>4 at bit 24
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
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
// a - hs for the two halves of a packed FP16 pair hs and two float32 values (exact: hs is a rounded to 11 bits), by
// one mixed-precision FMA each (FHFMA: hs * -1 + a); with ABS the float32 operands enter as |a|
template <bool ABS>
__device__ __forceinline__ void split_rest(uint32_t hs, float a0, float a1, float& l0, float& l1) {
  const uint16_t m1 = 0xBC00;  // fp16(-1)
  if (ABS) {
    a0 = fabsf(a0);
    a1 = fabsf(a1);
  }
  asm("{.reg .b16 a, b; mov.b32 {a, b}, %2; fma.rn.f32.f16 %0, a, %5, %3; fma.rn.f32.f16 %1, b, %5, %4;}"
      : "=f"(l0), "=f"(l1)
      : "r"(hs), "f"(a0), "f"(a1), "h"(m1));
}
// Split of a pair of conditioning inputs: hs = fp16(v) packed, lo = fp16(v - hs) packed
__device__ __forceinline__ void split_x_pair(float va, float vb, uint32_t& hs, uint32_t& lo) {
  hs = pack_f16x2(vb, va);
  float la, lb;
  split_rest<false>(hs, va, vb, la, lb);
  lo = pack_f16x2(lb, la);
}

// ---------------------------------------------------------------------------------------------
// host side: hi / lo split and the operand layout
// ---------------------------------------------------------------------------------------------
// w = hi + lo, hi = the 11 leading significant bits (truncated), lo exact in float32
static inline void split_hi_lo(float v, float* hi, float* lo) {
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

// float -> IEEE half, round to nearest even; *overflow is set when the value does not fit (NaN stays NaN)
static inline uint16_t f32_to_f16(float f, bool* overflow) {
  uint32_t x;
  memcpy(&x, &f, 4);
  const uint32_t sign = (x >> 16) & 0x8000u;
  const uint32_t absx = x & 0x7fffffffu;
  if (absx > 0x7f800000u) return (uint16_t)(sign | 0x7e00u);
  if (absx >= 0x477ff000u) {  // >= 65520 rounds past the largest half
    *overflow = true;
    return (uint16_t)(sign | 0x7bffu);
  }
  if (absx < 0x33000001u) return (uint16_t)sign;  // < 2^-25 (rounds to zero)
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

}  // namespace tcx
}  // namespace hb
