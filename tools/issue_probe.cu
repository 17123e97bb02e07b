// Two rate probes behind round-2 kernel decisions (development aid, not part of the library):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/issue_probe.bin tools/issue_probe.cu
// (1) tcgen05.mma issued from SEVERAL warps at once.  tools/mma_probe.cu measured ~45 cycles per MMA for
//     N <= 64 from ONE issuing thread; this separates a per-warp issue limit from a floor of the (shared)
//     tensor pipe: NW warps issue R MMAs each into their own accumulators, the aggregate cycles per MMA is
//     printed.  If it falls with NW the floor is per issuing warp.
// (2) packed FP32 (fma.rn.f32x2 / mul.f32x2 / add.f32x2, sm_100) against scalar FMA chains: issue slots and
//     pipe cycles per lane-FMA with 1..8 warps per scheduler.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ uint32_t idesc_f16(int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

// TS = 1: A from TMEM (GEMM2 of the kernel), 0: A from shared memory (GEMM1)
template <int TS>
__global__ void __launch_bounds__(256, 1) mma_multi(int n, int nw, int reps, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t bar[8];
  __shared__ long long t_end[8], t_beg[8];
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 256) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tmem_base_s;
  const int w = threadIdx.x >> 5;
  for (int trial = 0; trial < 4; ++trial) {
    __syncthreads();
    if (w < nw && (threadIdx.x & 31) == 0) {
      const uint32_t sb = smem_u32(smem);
      const uint64_t ad = make_desc(sb, 2048, 128);
      const uint64_t bd = make_desc(sb + 32768, 16 * n, 128);
      const uint32_t id = idesc_f16(n);
      // accumulators: N <= 64 -> 64 columns per warp (up to 4 warps in 256 columns); operands (TS) at column 320+
      const uint32_t d = tb + (uint32_t)((n <= 64 ? 64 : 128) * (w % (n <= 64 ? 4 : 2)));
      const uint32_t a = tb + 320 + 16 * (w % 8);
      const long long t0 = clock64();
      for (int r = 0; r < reps; r += 32) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          if (TS)
            asm volatile("{.reg .pred p; setp.ne.b32 p, 1, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;}" ::"r"(d), "r"(a), "l"(bd), "r"(id) : "memory");
          else
            asm volatile("{.reg .pred p; setp.ne.b32 p, 1, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;}" ::"r"(d), "l"(ad), "l"(bd), "r"(id) : "memory");
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[w])) : "memory");
      asm volatile("{.reg .pred p; WAIT_%=: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1; @!p bra WAIT_%=;}" ::"r"(smem_u32(&bar[w])), "r"((uint32_t)(trial & 1)) : "memory");
      t_beg[w] = t0;
      t_end[w] = clock64();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      long long b = t_beg[0], e = t_end[0];
      for (int i = 1; i < nw; ++i) {
        if (t_beg[i] < b) b = t_beg[i];
        if (t_end[i] > e) e = t_end[i];
      }
      if (trial == 0 || e - b < out[0]) out[0] = e - b;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512u) : "memory");
  }
}

template <int TS>
static void run_mma(const char* name, int n, int nw, long long* dout) {
  const int reps = 256;
  auto k = mma_multi<TS>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  k<<<1, 256, 96 * 1024>>>(n, nw, reps, dout);
  long long cyc = 0;
  cudaError_t e = cudaMemcpy(&cyc, dout, sizeof(cyc), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) {
    printf("%-20s N=%3d warps=%d  CUDA error %s\n", name, n, nw, cudaGetErrorString(e));
    return;
  }
  printf("%-20s N=%3d issuing warps=%d  %7.1f cycles per MMA aggregate  (%7.1f per warp-MMA)\n", name, n, nw,
         (double)cyc / (reps * nw), (double)cyc / reps);
}


// (3) Are tcgen05.mma accumulations into ONE accumulator from SEVERAL issuing warps lossless?  A (TMEM) and B (shared
//     memory) are all ones, so every MMA adds exactly 16 to each of the 128 x N accumulator words; nw warps issue
//     `reps` MMAs each with the accumulate flag set into the same zeroed columns.  Any lost update shows as a word
//     that differs from 16 * reps * nw.
__global__ void __launch_bounds__(256, 1) mma_atomic(int n, int nw, int reps, int trials, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t bar[8];
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 256) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u;  // halves 1.0
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tmem_base_s;
  const int w = threadIdx.x >> 5;
  long long bad = 0;
  for (int trial = 0; trial < trials; ++trial) {
    if (w < 4) {  // warps 0-3 own TMEM lanes 32 w .. 32 w + 31: zero D (columns 0..127), ones in A (columns 320..327)
      const uint32_t lane_base = ((uint32_t)(w * 32)) << 16;
      for (int c = 0; c < 128; c += 8)
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(tb + lane_base + c), "r"(0u) : "memory");
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(tb + lane_base + 320), "r"(0x3C003C00u) : "memory");
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (w < nw && (threadIdx.x & 31) == 0) {
      const uint32_t sb = smem_u32(smem);
      const uint64_t bd = make_desc(sb + 32768, 16 * n, 128);
      const uint32_t id = idesc_f16(n);
      for (int r = 0; r < reps; r += 16) {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          asm volatile("{.reg .pred p; setp.ne.b32 p, 1, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;}" ::"r"(tb), "r"(tb + 320), "l"(bd), "r"(id) : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[w])) : "memory");
    }
    // every thread waits for every issuing warp's commit
    for (int i = 0; i < nw; ++i)
      asm volatile("{.reg .pred p; WAIT_%=: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1; @!p bra WAIT_%=;}" ::"r"(smem_u32(&bar[i])), "r"((uint32_t)(trial & 1)) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (w < 4) {
      const uint32_t lane_base = ((uint32_t)(w * 32)) << 16;
      const float want = 16.0f * (float)reps * (float)nw;
      for (int c = 0; c < n; c += 8) {
        uint32_t v[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                     : "r"(tb + lane_base + c) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int k = 0; k < 8; ++k) bad += (__uint_as_float(v[k]) != want) ? 1 : 0;
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
  }
  if (bad) atomicAdd(reinterpret_cast<unsigned long long*>(out), (unsigned long long)bad);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512u) : "memory");
  }
}

static void run_atomic(int n, int nw, long long* dout) {
  const int reps = 256, trials = 64;
  cudaMemset(dout, 0, 8);
  cudaFuncSetAttribute(mma_atomic, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  mma_atomic<<<1, 256, 96 * 1024>>>(n, nw, reps, trials, dout);
  long long bad = -1;
  cudaError_t e = cudaMemcpy(&bad, dout, sizeof(bad), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) {
    printf("shared accumulator N=%3d warps=%d  CUDA error %s\n", n, nw, cudaGetErrorString(e));
    return;
  }
  printf("shared accumulator   N=%3d issuing warps=%d  %lld wrong words of %d (x %d trials; %d MMAs per warp)\n", n, nw,
         bad, 128 * n, trials, reps);
}

// ---- packed FP32 ---------------------------------------------------------------------------
// MODE 0: scalar fma chains (8 independent accumulators per thread), 1: fma.rn.f32x2 (4 packed accumulators = the
// same 8 lane-FMAs per step), 2: scalar FMNMX+FMUL (leaky-relu shape), 3: mul.f32x2 + 2 FMNMX
template <int MODE>
__global__ void __launch_bounds__(1024, 1) fp32_probe(int iters, float seed, float* out, long long* cyc) {
  float a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = seed + (float)(threadIdx.x + i);
  const float m = 1.0000001f, c = 1e-7f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (MODE == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], m, c);
      } else if (MODE == 1) {
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
          asm volatile("{.reg .b64 x, y, z; mov.b64 x, {%0, %1}; mov.b64 y, {%2, %2}; mov.b64 z, {%3, %3};\n\t"
                       "fma.rn.f32x2 x, x, y, z; mov.b64 {%0, %1}, x;}"
                       : "+f"(a[i]), "+f"(a[i + 1])
                       : "f"(m), "f"(c));
        }
      } else if (MODE == 2) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fmaxf(a[i], 0.01f * a[i]) + c;
      } else {
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
          float p, q;
          asm volatile("{.reg .b64 x, y; mov.b64 x, {%2, %3}; mov.b64 y, {%4, %4};\n\t"
                       "mul.f32x2 x, x, y; mov.b64 {%0, %1}, x;}"
                       : "=f"(p), "=f"(q)
                       : "f"(a[i]), "f"(a[i + 1]), "f"(0.01f));
          a[i] = fmaxf(a[i], p) + c;
          a[i + 1] = fmaxf(a[i + 1], q) + c;
        }
      }
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

template <int MODE>
static void run_fp(const char* name, int threads, float* dout, long long* dcyc) {
  const int iters = 2000;
  fp32_probe<MODE><<<1, threads>>>(iters, 1.0f, dout, dcyc);
  long long cyc = 0;
  cudaError_t e = cudaMemcpy(&cyc, dcyc, sizeof(cyc), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) {
    printf("%-34s CUDA error %s\n", name, cudaGetErrorString(e));
    return;
  }
  const double lane_ops = (double)iters * 64.0;  // 8 x 8 element-steps per thread and iteration
  printf("%-34s warps/scheduler=%d  %6.3f cycles per element-step per warp-slot  (%.3f x 1 warp-inst/cycle/scheduler)\n",
         name, threads / 128, (double)cyc / lane_ops / (threads / 128.0) * 1.0, lane_ops * (threads / 128.0) / (double)cyc);
}

int main() {
  long long* dout;
  cudaMalloc(&dout, 64);
  const int ns[3] = {32, 64, 128};
  for (int i = 0; i < 3; ++i)
    for (int nw = 1; nw <= 4; ++nw) {
      if (ns[i] == 128 && nw > 2) continue;
      run_mma<1>("f16 TS (A in TMEM)", ns[i], nw, dout);
    }
  for (int nw = 1; nw <= 2; ++nw) run_mma<0>("f16 SS (A in smem)", 128, nw, dout);
  for (int nw = 1; nw <= 3; ++nw) run_mma<0>("f16 SS (A in smem)", 64, nw, dout);
  for (int nw = 1; nw <= 4; ++nw) run_atomic(32, nw, dout);
  for (int nw = 2; nw <= 4; ++nw) run_atomic(64, nw, dout);
  float* fout;
  cudaMalloc(&fout, 1024 * 4);
  const int th[4] = {128, 256, 512, 1024};
  for (int i = 0; i < 4; ++i) {
    run_fp<0>("scalar fma", th[i], fout, dout);
    run_fp<1>("fma.rn.f32x2", th[i], fout, dout);
    run_fp<2>("scalar mul + max + add", th[i], fout, dout);
    run_fp<3>("mul.f32x2 + 2 max + 2 add", th[i], fout, dout);
  }
  cudaFree(dout);
  cudaFree(fout);
  return 0;
}
