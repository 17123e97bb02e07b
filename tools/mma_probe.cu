// tcgen05.mma issue / execution rates on one SM for the instruction shapes realnvp_tc_kernel uses
// (M = 128, cta_group::1, K-major no-swizzle operands).  Development probe, not part of the library:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/mma_probe.bin tools/mma_probe.cu
//   tools/mma_probe.bin             (on a B200; prints cycles per MMA for every variant)
// One CTA, one issuing thread: R back-to-back MMAs into one accumulator, then tcgen05.commit and a wait on
// the mbarrier; clock64() around the whole sequence.  Operand contents are zeros (timing does not depend on
// the data).  Results of round 1 are kept in profiles/r1_mma_probe.txt.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ uint32_t idesc_tf32(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ uint32_t idesc_f16(int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
template <int KIND, int TS>  // KIND 0 = tf32 (K = 8), 1 = f16 (K = 16); TS: A from TMEM
__device__ __forceinline__ void mma(uint32_t d, uint64_t adesc, uint32_t atmem, uint64_t bdesc, uint32_t idesc) {
  if (KIND == 0 && !TS)
    asm volatile("{.reg .pred p; setp.ne.b32 p, 1, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;}" ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc) : "memory");
  if (KIND == 0 && TS)
    asm volatile("{.reg .pred p; setp.ne.b32 p, 1, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;}" ::"r"(d), "r"(atmem), "l"(bdesc), "r"(idesc) : "memory");
  if (KIND == 1 && !TS)
    asm volatile("{.reg .pred p; setp.ne.b32 p, 1, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;}" ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc) : "memory");
  if (KIND == 1 && TS)
    asm volatile("{.reg .pred p; setp.ne.b32 p, 1, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;}" ::"r"(d), "r"(atmem), "l"(bdesc), "r"(idesc) : "memory");
}

// GROUP = 0: one kind only; GROUP = g > 0: g tf32 MMAs, then g f16 MMAs, then g tf32 ... (kind switches)
template <int KIND, int TS, int GROUP>
__global__ void __launch_bounds__(128, 1) probe(int n, int reps, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t bar;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tmem_base_s;
  if (threadIdx.x == 0) {
    const uint32_t sb = smem_u32(smem);
    // A: [128 x K] at offset 0 (LBO 2048, SBO 128); B: [N x K] at 32 KiB (LBO 16 N, SBO 128)
    const uint64_t ad = make_desc(sb, 2048, 128);
    const uint64_t bd = make_desc(sb + 32768, 16 * n, 128);
    const uint32_t i0 = idesc_tf32(n), i1 = idesc_f16(n);
    long long best = 1ll << 60;
    for (int trial = 0; trial < 5; ++trial) {
      const long long t0 = clock64();
      for (int r = 0; r < reps; r += 64) {
#pragma unroll
        for (int j = 0; j < 64; ++j) {  // straight-line issue: 64 MMAs per loop iteration
          const bool f16 = GROUP <= 0 ? (KIND == 1) : (((j / (GROUP > 0 ? GROUP : 1)) & 1) != 0);
          // GROUP < 0: rotate over -GROUP independent accumulators (64 columns apart) instead of one
          const uint32_t d = GROUP < 0 ? tb + 64u * (uint32_t)(j % (GROUP < 0 ? -GROUP : 1)) : tb;
          if (f16) mma<1, TS>(d, ad, tb + 320, bd, i1);
          else mma<0, TS>(d, ad, tb + 320, bd, i0);
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      asm volatile(
          "{.reg .pred p; WAIT_%=: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1; @!p bra WAIT_%=;}" ::"r"(smem_u32(&bar)),
          "r"((uint32_t)(trial & 1))
          : "memory");
      const long long t1 = clock64();
      if (t1 - t0 < best) best = t1 - t0;
    }
    out[0] = best;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512u) : "memory");
  }
}

template <int KIND, int TS, int GROUP>
static void run(const char* name, int n, long long* dout) {
  const int reps = 512;
  auto k = probe<KIND, TS, GROUP>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  k<<<1, 128, 96 * 1024>>>(n, reps, dout);
  long long cyc = 0;
  cudaError_t e = cudaMemcpy(&cyc, dout, sizeof(cyc), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) {
    printf("%-34s N=%3d  CUDA error %s\n", name, n, cudaGetErrorString(e));
    return;
  }
  const double per = (double)cyc / reps;
  const double kk = KIND ? 16.0 : 8.0;
  printf("%-34s N=%3d  %7.1f cycles per MMA   (M x N x K = 128 x %d x %d)\n", name, n, per, n, (int)kk);
}

int main() {
  long long* dout;
  cudaMalloc(&dout, 8);
  const int ns[4] = {32, 64, 128, 256};
  for (int i = 0; i < 4; ++i) run<0, 0, 0>("tf32 SS (A, B from smem)", ns[i], dout);
  for (int i = 0; i < 4; ++i) run<0, 1, 0>("tf32 TS (A from TMEM)", ns[i], dout);
  for (int i = 0; i < 4; ++i) run<1, 0, 0>("f16 SS", ns[i], dout);
  for (int i = 0; i < 4; ++i) run<1, 1, 0>("f16 TS", ns[i], dout);
  for (int i = 0; i < 3; ++i) run<0, 1, 1>("TS, tf32 / f16 in groups of 1", ns[i], dout);
  for (int i = 0; i < 3; ++i) run<0, 1, 2>("TS, tf32 / f16 in groups of 2", ns[i], dout);
  for (int i = 0; i < 3; ++i) run<0, 1, 4>("TS, tf32 / f16 in groups of 4", ns[i], dout);
  for (int i = 0; i < 3; ++i) run<0, 1, 8>("TS, tf32 / f16 in groups of 8", ns[i], dout);
  for (int i = 0; i < 3; ++i) run<0, 1, 16>("TS, tf32 / f16 in groups of 16", ns[i], dout);
  for (int i = 0; i < 3; ++i) run<0, 1, 32>("TS, tf32 / f16 in groups of 32", ns[i], dout);
  for (int i = 0; i < 3; ++i) run<0, 0, 4>("SS, tf32 / f16 in groups of 4", ns[i], dout);
  for (int i = 0; i < 2; ++i) run<0, 1, -2>("tf32 TS, 2 independent accumulators", ns[i], dout);
  for (int i = 0; i < 2; ++i) run<0, 1, -4>("tf32 TS, 4 independent accumulators", ns[i], dout);
  for (int i = 0; i < 2; ++i) run<1, 1, -2>("f16 TS, 2 independent accumulators", ns[i], dout);
  for (int i = 0; i < 2; ++i) run<0, 0, -2>("tf32 SS, 2 independent accumulators", ns[i], dout);
  for (int i = 0; i < 2; ++i) run<0, 0, -4>("tf32 SS, 4 independent accumulators", ns[i], dout);
  cudaFree(dout);
  return 0;
}
