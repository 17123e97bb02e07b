// Accumulation-only path, merge and finalise kernels, and the hb_accum_* C ABI.
//
// Reference code replaced: harmonic/evidence.py:278-358 (batched add_chains after predict),
// :125-189 (process_run), :307-319 (shift selection).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <vector>

#include "hb_common.cuh"
#include "hb_internal.h"

namespace hb {

// ------------------------------------------------------------------------------------------
// accumulate_precomputed_kernel: HBM-bound streaming fold of (lnpred, lnpost) rows.
// One CTA = one flusher with a contiguous row range; 256 threads x 8 rows per tile.
// Algorithmic traffic: sizeof(PT) + sizeof(LT) bytes per sample, read once.
// ------------------------------------------------------------------------------------------
constexpr int ACC_THREADS = 256;
constexpr int ACC_R = 8;
constexpr int ACC_TILE = ACC_THREADS * ACC_R;

template <typename T>
__device__ __forceinline__ void load8(const T* __restrict__ p, long long r, float (&v)[8]);

// rows r .. r+3 and r + 4*ACC_THREADS .. +3: each 128-bit load is fully coalesced across the warp
template <>
__device__ __forceinline__ void load8<float>(const float* __restrict__ p, long long r, float (&v)[8]) {
  const float4* q = reinterpret_cast<const float4*>(p + r);
  float4 a = __ldcs(q), b = __ldcs(q + 256);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <>
__device__ __forceinline__ void load8<double>(const double* __restrict__ p, long long r, float (&v)[8]) {
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    const double2* q = reinterpret_cast<const double2*>(p + r + (long long)g * 4 * 256);
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      double2 a = __ldcs(q + k);
      v[4 * g + 2 * k] = (float)a.x;
      v[4 * g + 2 * k + 1] = (float)a.y;
    }
  }
}

template <typename PT, typename LT, bool VEC>
__global__ void __launch_bounds__(ACC_THREADS, 6)
accumulate_precomputed_kernel(const PT* __restrict__ lnpred, const LT* __restrict__ lnpost,
                              FoldParams p) {
  __shared__ double scratch[3 * (ACC_THREADS / 32)];
  const long long row_lo = (long long)blockIdx.x * p.rows_per_flusher;
  long long row_hi = row_lo + p.rows_per_flusher;
  if (row_hi > p.nrows) row_hi = p.nrows;
  const bool had_rows = row_lo < p.nrows;
  Fold<ACC_THREADS> fold(p, blockIdx.x, threadIdx.x, scratch, had_rows ? row_lo : 0);
  for (long long t0 = row_lo; t0 < row_hi; t0 += ACC_TILE) {
    long long t1 = t0 + ACC_TILE;
    float a[ACC_R], b[ACC_R];
    const long long r0 = t0 + (long long)threadIdx.x * 4;
    if (VEC && t1 <= row_hi) {
      load8<PT>(lnpred, r0, a);
      load8<LT>(lnpost, r0, b);
    } else {
      if (t1 > row_hi) t1 = row_hi;
#pragma unroll
      for (int q = 0; q < ACC_R; ++q) {
        const long long r = r0 + (long long)(q / 4) * 4 * ACC_THREADS + (q % 4);
        a[q] = (r < t1) ? (float)lnpred[r] : 0.f;
        b[q] = (r < t1) ? (float)lnpost[r] : 0.f;
      }
    }
    fold.tile<ACC_R, 4>(t0, t1, a, b);
  }
  fold.finish(had_rows);
}

// ------------------------------------------------------------------------------------------
// merge_kernel: fold the (chain, flusher) records of one launch into the persistent state in a
// fixed order; block 0 also reduces the per-call scalars.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void merge_pair(double& m0, double& s0, double m1, double s1) {
  if (!(s1 > 0.0)) return;
  if (!(s0 > 0.0)) {
    m0 = m1;
    s0 = s1;
    return;
  }
  const double M = fmax(m0, m1);
  s0 = s0 * exp(m0 - M) + s1 * exp(m1 - M);
  m0 = M;
}

__global__ void merge_kernel(ChainState* __restrict__ state, double* __restrict__ globals,
                             FoldParams p, long long chain_lo, long long nflushers) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < p.nch) {
    const long long a = p.start[c], b = p.start[c + 1];
    if (b > a) {
      ChainState st = state[chain_lo + c];
      const long long f0 = a / p.rows_per_flusher, f1 = (b - 1) / p.rows_per_flusher;
      for (long long f = f0; f <= f1; ++f) {
        const ChainRec r = p.recs[c + f];
        if (r.valid != 0.0) {
          merge_pair(st.m, st.s, r.m, r.s);
          st.nnan += r.nnan;
        }
      }
      st.n += (double)(b - a);
      state[chain_lo + c] = st;
    }
  }
  // per-call scalars: warp k of block 0 reduces column k of gpart (lane-strided, then a shuffle
  // tree: a fixed order, so the result is reproducible)
  if (blockIdx.x == 0 && (threadIdx.x >> 5) < G_N) {
    const int k = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double acc = (k < 2) ? 0.0 : ((k & 1) ? CUDART_INF : -CUDART_INF);
    for (long long f = lane; f < nflushers; f += 32) {
      const double o = p.gpart[f * G_N + k];
      if (k < 2) acc += o;
      else if (k & 1) acc = fmin(acc, o);
      else acc = fmax(acc, o);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      const double v = __shfl_xor_sync(0xffffffffu, acc, o);
      if (k < 2) acc += v;
      else if (k & 1) acc = fmin(acc, v);
      else acc = fmax(acc, v);
    }
    if (lane == 0) {
      const double g = globals[k];
      globals[k] = (k < 2) ? g + acc : ((k & 1) ? fmin(g, acc) : fmax(g, acc));
    }
  }
}

// fold nranks packed blocks ([4n partials | 8 globals], block r at packed + r*stride) in rank order
__global__ void merge_packed_kernel(ChainState* __restrict__ state, double* __restrict__ globals,
                                    const double* __restrict__ packed, long long stride, int nranks,
                                    long long n) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c < n) {
    ChainState st = state[c];
    for (int r = 0; r < nranks; ++r) {
      const double* part = packed + (long long)r * stride;
      merge_pair(st.m, st.s, part[4 * c + 0], part[4 * c + 1]);
      st.n += part[4 * c + 2];
      st.nnan += part[4 * c + 3];
    }
    state[c] = st;
  }
  if (blockIdx.x == 0 && threadIdx.x < G_N) {
    const int k = threadIdx.x;
    double acc = globals[k];
    for (int r = 0; r < nranks; ++r) {
      const double o = packed[(long long)r * stride + 4 * n + k];
      if (k < 2) acc += o;
      else if (k & 1) acc = fmin(acc, o);
      else acc = fmax(acc, o);
    }
    globals[k] = acc;
  }
}

__global__ void reset_globals_kernel(double* g) {
  const int k = threadIdx.x;
  if (k < G_N) g[k] = (k < 2) ? 0.0 : ((k & 1) ? CUDART_INF : -CUDART_INF);
}

__global__ void reset_state_kernel(ChainState* st, long long n) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c < n) {
    ChainState z;
    z.m = -CUDART_INF;
    z.s = 0.0;
    z.n = 0.0;
    z.nnan = 0.0;
    st[c] = z;
  }
}

// merge host-provided partials (other ranks / checkpoints) into the state
__global__ void merge_partials_kernel(ChainState* __restrict__ state, double* __restrict__ globals,
                                      const double* __restrict__ part, const double* __restrict__ g,
                                      long long n) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c < n) {
    ChainState st = state[c];
    merge_pair(st.m, st.s, part[4 * c + 0], part[4 * c + 1]);
    st.n += part[4 * c + 2];
    st.nnan += part[4 * c + 3];
    state[c] = st;
  }
  if (blockIdx.x == 0 && threadIdx.x < G_N && g != nullptr) {
    const int k = threadIdx.x;
    double acc = globals[k];
    const double o = g[k];
    if (k < 2) acc += o;
    else if (k & 1) acc = fmin(acc, o);
    else acc = fmax(acc, o);
    globals[k] = acc;
  }
}

// ------------------------------------------------------------------------------------------
// finalize_kernel: Evidence.process_run (harmonic/evidence.py:125-189) in float64, shift-free.
//   L_c = m_c + ln s_c = ln running_sum[c] - shift
//   ln rho = LSE_c(L_c) - ln N                                     (:136-139,178)
//   ln z_c = ln rho + ln|expm1(L_c - ln n_c - ln rho)|             (:156)
//   lnV  = LSE_c(2 ln z_c + ln n_c) - ln N                         (:157-161, minus 2*shift)
//   lnK  = LSE_c(4 ln z_c + ln n_c) - ln N - 2 lnV                 (:162-169)
//   n_eff = N^2 / sum n_c^2                                        (:172-174)
//   ln var = lnV - ln(n_eff-1);  ln varvar = 2 lnV - 3 ln n_eff + ln((K-1) + 2/(n_eff-1))  (:179-187)
// Single CTA; out layout: hb_result, then 4 arrays of nchains doubles.
// ------------------------------------------------------------------------------------------
constexpr int FIN_THREADS = 256;

__device__ double block_max(double v, double* sh) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = sh[0];
  for (int w = 1; w < FIN_THREADS / 32; ++w) r = fmax(r, sh[w]);
  return r;
}
__device__ double block_sum(double v, double* sh) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = 0.0;
  for (int w = 0; w < FIN_THREADS / 32; ++w) r += sh[w];
  return r;
}
// NaN-propagating logsumexp of f(c) over chains (scipy.special.logsumexp semantics)
template <typename F>
__device__ double block_lse(long long n, F f, double* sh) {
  double mx = -CUDART_INF;
  bool has_nan = false;
  for (long long c = threadIdx.x; c < n; c += FIN_THREADS) {
    const double v = f(c);
    if (isnan(v)) has_nan = true;
    mx = fmax(mx, v);
  }
  const double nanflag = block_sum(has_nan ? 1.0 : 0.0, sh);
  mx = block_max(mx, sh);
  if (nanflag > 0.0) return CUDART_NAN;
  if (isinf(mx)) return mx;  // all -inf -> -inf ; +inf -> +inf
  double sum = 0.0;
  for (long long c = threadIdx.x; c < n; c += FIN_THREADS) sum += exp(f(c) - mx);
  sum = block_sum(sum, sh);
  return mx + log(sum);
}

__global__ void __launch_bounds__(FIN_THREADS)
finalize_kernel(const ChainState* __restrict__ state, const double* __restrict__ globals,
                long long nch, int shift_mode, double shift_in, double* __restrict__ out) {
  __shared__ double sh[FIN_THREADS / 32];
  hb_result* res = reinterpret_cast<hb_result*>(out);
  double* ln_inv_pc = out + sizeof(hb_result) / sizeof(double);
  double* rsum = ln_inv_pc + nch;
  double* npc = rsum + nch;
  double* neff_pc = npc + nch;

  // shift selection (harmonic/evidence.py:307-319) -- only cosmetic in the shift-free form
  double shift = shift_in;
  if (isnan(shift_in)) {
    if (shift_mode == HB_MEAN_SHIFT) shift = -globals[G_SUM] / globals[G_CNT];
    else if (shift_mode == HB_MAX_SHIFT) shift = -globals[G_ARGMAX];
    else if (shift_mode == HB_MIN_SHIFT) shift = -globals[G_ARGMIN];
    else shift = (fabs(globals[G_ARGMAX]) >= fabs(globals[G_ARGMIN])) ? -globals[G_ARGMAX]
                                                                     : -globals[G_ARGMIN];
  }

  double nsum = 0.0, n2sum = 0.0;
  for (long long c = threadIdx.x; c < nch; c += FIN_THREADS) {
    const double n = state[c].n;
    nsum += n;
    n2sum += n * n;
  }
  const double N = block_sum(nsum, sh);
  const double N2 = block_sum(n2sum, sh);
  const double lnN = log(N);
  auto L = [&](long long c) {
    const ChainState st = state[c];
    return (st.s > 0.0) ? st.m + log(st.s) : -CUDART_INF;
  };
  const double ln_rho = block_lse(nch, L, sh) - lnN;
  auto lnz = [&](long long c) {
    const double lrc = L(c) - log(state[c].n);
    if (isinf(ln_rho) && ln_rho < 0) return -CUDART_INF;  // all sums zero
    return ln_rho + log(fabs(expm1(lrc - ln_rho)));
  };
  const double lnV =
      block_lse(nch, [&](long long c) { return 2.0 * lnz(c) + log(state[c].n); }, sh) - lnN;
  const double lnK =
      block_lse(nch, [&](long long c) { return 4.0 * lnz(c) + log(state[c].n); }, sh) - lnN -
      2.0 * lnV;
  const double n_eff = N * N / N2;
  for (long long c = threadIdx.x; c < nch; c += FIN_THREADS) {
    const ChainState st = state[c];
    const double Lc = L(c);
    ln_inv_pc[c] = Lc - log(st.n);
    rsum[c] = exp(Lc + shift);
    npc[c] = st.n;
    neff_pc[c] = st.n - st.nnan;
  }
  if (threadIdx.x == 0) {
    const double kur = exp(lnK);
    res->ln_evidence_inv = ln_rho;
    res->ln_evidence_inv_var = lnV - log(n_eff - 1.0);
    res->ln_evidence_inv_var_var =
        2.0 * lnV - 3.0 * log(n_eff) + log((kur - 1.0) + 2.0 / (n_eff - 1.0));
    res->ln_kurtosis = lnK;
    res->n_eff = n_eff;
    res->shift_value = shift;
    const bool any_arg = globals[G_CNT] > 0.0;
    res->lnargmax = any_arg ? globals[G_ARGMAX] + shift : CUDART_NAN;
    res->lnargmin = any_arg ? globals[G_ARGMIN] + shift : CUDART_NAN;
    res->lnprobmax = globals[G_PROBMAX];
    res->lnprobmin = globals[G_PROBMIN];
    res->lnpredictmax = globals[G_PREDMAX];
    res->lnpredictmin = globals[G_PREDMIN];
    res->nsamples_total = N;
    res->reserved[0] = res->reserved[1] = res->reserved[2] = 0.0;
  }
}

// process_run on real-space totals (reference formulas verbatim in float64; sums may be <= 0)
__global__ void __launch_bounds__(FIN_THREADS)
finalize_realspace_kernel(const double* __restrict__ S, const double* __restrict__ n, long long nch,
                          double shift, double* __restrict__ out) {
  __shared__ double sh[FIN_THREADS / 32];
  hb_result* res = reinterpret_cast<hb_result*>(out);
  double* ln_inv_pc = out + sizeof(hb_result) / sizeof(double);
  double ssum = 0.0, nsum = 0.0, n2sum = 0.0;
  for (long long c = threadIdx.x; c < nch; c += FIN_THREADS) {
    ssum += S[c];
    nsum += n[c];
    n2sum += n[c] * n[c];
  }
  const double Ssum = block_sum(ssum, sh);
  const double N = block_sum(nsum, sh);
  const double N2 = block_sum(n2sum, sh);
  const double rho = Ssum / N;
  const double lnN = log(N);
  auto lnz = [&](long long c) { return log(fabs(S[c] / n[c] - rho)); };
  const double lnV = block_lse(nch, [&](long long c) { return 2.0 * lnz(c) + log(n[c]); }, sh) - lnN;
  const double lnK =
      block_lse(nch, [&](long long c) { return 4.0 * lnz(c) + log(n[c]); }, sh) - lnN - 2.0 * lnV;
  const double n_eff = N * N / N2;
  for (long long c = threadIdx.x; c < nch; c += FIN_THREADS) ln_inv_pc[c] = log(S[c]) - shift - log(n[c]);
  if (threadIdx.x == 0) {
    const double kur = exp(lnK);
    res->ln_evidence_inv = log(rho) - shift;
    res->ln_evidence_inv_var = lnV - 2.0 * shift - log(n_eff - 1.0);
    res->ln_evidence_inv_var_var =
        2.0 * lnV - 3.0 * log(n_eff) - 4.0 * shift + log((kur - 1.0) + 2.0 / (n_eff - 1.0));
    res->ln_kurtosis = lnK;
    res->n_eff = n_eff;
    res->shift_value = shift;
    res->lnargmax = res->lnargmin = res->lnprobmax = res->lnprobmin = CUDART_NAN;
    res->lnpredictmax = res->lnpredictmin = CUDART_NAN;
    res->nsamples_total = N;
    res->reserved[0] = res->reserved[1] = res->reserved[2] = 0.0;
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
int ensure_scratch(Accum* acc, long long nch, long long nflushers) {
  const size_t need_recs = (size_t)(nch + nflushers);
  if (need_recs > acc->recs_cap) {
    if (acc->recs) cudaFree(acc->recs);
    acc->recs = nullptr;
    acc->recs_cap = 0;
    HB_CUDA(cudaMalloc(&acc->recs, need_recs * sizeof(ChainRec)));
    acc->recs_cap = need_recs;
  }
  const size_t need_g = (size_t)nflushers * G_N;
  if (need_g > acc->gpart_cap) {
    if (acc->gpart) cudaFree(acc->gpart);
    acc->gpart = nullptr;
    acc->gpart_cap = 0;
    HB_CUDA(cudaMalloc(&acc->gpart, need_g * sizeof(double)));
    acc->gpart_cap = need_g;
  }
  return HB_OK;
}

int launch_merge(Accum* acc, const FoldParams& p, long long chain_lo, long long nflushers,
                 cudaStream_t st) {
  const int threads = 256;
  const int blocks = (int)((p.nch + threads - 1) / threads);
  merge_kernel<<<blocks > 0 ? blocks : 1, threads, 0, st>>>(acc->state, acc->globals, p, chain_lo,
                                                             nflushers);
  count_launch();
  HB_CUDA(cudaGetLastError());
  return HB_OK;
}

// Build the local (clamped) start offsets of chains [chain_lo, chain_hi) for rows [a, b) of the
// shard, upload them, and zero the record buffer.
int prepare_fold(Accum* acc, const int64_t* start, long long chain_lo, long long chain_hi,
                 long long a, long long b, long long nflushers, long long rows_per_flusher,
                 FoldParams* p, cudaStream_t st) {
  const long long nch = chain_hi - chain_lo;
  int rc = ensure_scratch(acc, nch, nflushers);
  if (rc) return rc;
  static thread_local std::vector<long long> local;
  local.resize(nch + 1);
  const long long base = start[chain_lo];
  for (long long c = 0; c <= nch; ++c) {
    long long v = start[chain_lo + c] - base - a;
    if (v < 0) v = 0;
    if (v > b - a) v = b - a;
    local[c] = v;
  }
  // start_dev is reused across launches on the same stream: the copy is stream-ordered, and the
  // pageable source is consumed before cudaMemcpyAsync returns.
  HB_CUDA(cudaMemcpyAsync(acc->start_dev, local.data(), (nch + 1) * sizeof(long long),
                          cudaMemcpyHostToDevice, st));
  HB_CUDA(cudaMemsetAsync(acc->recs, 0, (size_t)(nch + nflushers) * sizeof(ChainRec), st));
  p->start = acc->start_dev;
  p->nch = (int)nch;
  p->nrows = b - a;
  p->rows_per_flusher = rows_per_flusher;
  p->recs = acc->recs;
  p->gpart = acc->gpart;
  return HB_OK;
}

int reset_call_scalars(Accum* acc, cudaStream_t st) {
  reset_globals_kernel<<<1, 32, 0, st>>>(acc->globals);
  count_launch();
  HB_CUDA(cudaGetLastError());
  return HB_OK;
}

template <typename PT, typename LT>
static int launch_acc(Accum* acc, const PT* lnpred, const LT* lnpost, const int64_t* start,
                      long long chain_lo, long long chain_hi, long long a, long long b,
                      cudaStream_t st) {
  const long long nrows = b - a;
  if (nrows <= 0) return HB_OK;
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, acc->device);
  // many more CTAs than resident slots (64 per SM): measured on B200 with 1e9 rows, kernel time falls
  // from 1.54 ms at 8 CTAs/SM to 1.36 ms at 64 and stays flat beyond (profiles/README.md);
  // HB_ACC_CTAS_PER_SM overrides for experiments
  static int per_sm = [] {
    const char* e = getenv("HB_ACC_CTAS_PER_SM");
    return e ? atoi(e) : 64;
  }();
  long long nfl = (long long)sms * (per_sm > 0 ? per_sm : 64);
  long long rpf = (nrows + nfl - 1) / nfl;
  rpf = ((rpf + ACC_TILE - 1) / ACC_TILE) * ACC_TILE;
  nfl = (nrows + rpf - 1) / rpf;
  FoldParams p;
  int rc = prepare_fold(acc, start, chain_lo, chain_hi, a, b, nfl, rpf, &p, st);
  if (rc) return rc;
  const bool vec = ((uintptr_t)lnpred % 16 == 0) && ((uintptr_t)lnpost % 16 == 0);
  timing_begin(st);
  if (vec)
    accumulate_precomputed_kernel<PT, LT, true><<<(int)nfl, ACC_THREADS, 0, st>>>(lnpred, lnpost, p);
  else
    accumulate_precomputed_kernel<PT, LT, false><<<(int)nfl, ACC_THREADS, 0, st>>>(lnpred, lnpost, p);
  timing_end(st);
  count_launch();
  HB_CUDA(cudaGetLastError());
  return launch_merge(acc, p, chain_lo, nfl, st);
}

int accumulate_precomputed_device(Accum* acc, const void* lnpred, int pd, const void* lnpost, int ld,
                                  const int64_t* start, long long chain_lo, long long chain_hi,
                                  long long a, long long b, cudaStream_t st) {
  if (pd == HB_F32 && ld == HB_F32)
    return launch_acc(acc, (const float*)lnpred, (const float*)lnpost, start, chain_lo, chain_hi, a, b, st);
  if (pd == HB_F32 && ld == HB_F64)
    return launch_acc(acc, (const float*)lnpred, (const double*)lnpost, start, chain_lo, chain_hi, a, b, st);
  if (pd == HB_F64 && ld == HB_F32)
    return launch_acc(acc, (const double*)lnpred, (const float*)lnpost, start, chain_lo, chain_hi, a, b, st);
  return launch_acc(acc, (const double*)lnpred, (const double*)lnpost, start, chain_lo, chain_hi, a, b, st);
}

}  // namespace hb

using namespace hb;

// Released accumulators are parked here and handed out again for the same (device, nchains):
// creating one costs several cudaMalloc / stream / event calls, which a per-call Evidence would
// otherwise pay on every add_chains.
static std::mutex g_pool_mu;
static std::vector<Accum*> g_pool;

extern "C" {

int hb_accum_create(int64_t nchains, int device, hb_accum** out) {
  HB_REQUIRE(out != nullptr, "hb_accum_create: out is NULL");
  HB_REQUIRE(nchains >= 1, "nchains must be greater than 0.");
  int rc = require_device(device);
  if (rc) return rc;
  {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    for (size_t i = 0; i < g_pool.size(); ++i) {
      if (g_pool[i]->device == device && g_pool[i]->nchains == nchains) {
        Accum* a = g_pool[i];
        g_pool.erase(g_pool.begin() + i);
        *out = reinterpret_cast<hb_accum*>(a);
        HB_CUDA(cudaSetDevice(device));
        return hb_accum_reset(*out);
      }
    }
  }
  Accum* acc = new Accum();
  acc->device = device;
  acc->nchains = nchains;
  HB_CUDA(cudaSetDevice(device));
  HB_CUDA(cudaMalloc(&acc->state, nchains * sizeof(ChainState)));
  HB_CUDA(cudaMalloc(&acc->globals, G_N * sizeof(double)));
  HB_CUDA(cudaMalloc(&acc->start_dev, (nchains + 1) * sizeof(long long)));
  HB_CUDA(cudaMalloc(&acc->fin, sizeof(hb_result) + 4 * nchains * sizeof(double)));
  HB_CUDA(cudaStreamCreateWithFlags(&acc->copy_stream, cudaStreamNonBlocking));
  for (int i = 0; i < 2; ++i) {
    HB_CUDA(cudaEventCreateWithFlags(&acc->ev_copy[i], cudaEventDisableTiming));
    HB_CUDA(cudaEventCreateWithFlags(&acc->ev_done[i], cudaEventDisableTiming));
  }
  *out = reinterpret_cast<hb_accum*>(acc);
  return hb_accum_reset(*out);
}

int hb_accum_destroy(hb_accum* h) {
  if (!h) return HB_OK;
  Accum* acc = reinterpret_cast<Accum*>(h);
  cudaSetDevice(acc->device);
  cudaDeviceSynchronize();
  {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    if (g_pool.size() < 8) {
      g_pool.push_back(acc);
      return HB_OK;
    }
  }
  cudaFree(acc->state);
  cudaFree(acc->globals);
  cudaFree(acc->start_dev);
  cudaFree(acc->recs);
  cudaFree(acc->gpart);
  cudaFree(acc->fin);
  for (int i = 0; i < 2; ++i) {
    if (acc->dev[i]) cudaFree(acc->dev[i]);
    if (acc->pin[i]) cudaFreeHost(acc->pin[i]);
    if (acc->ev_copy[i]) cudaEventDestroy(acc->ev_copy[i]);
    if (acc->ev_done[i]) cudaEventDestroy(acc->ev_done[i]);
  }
  if (acc->copy_stream) cudaStreamDestroy(acc->copy_stream);
  delete acc;
  return HB_OK;
}

int hb_accum_reset(hb_accum* h) {
  HB_REQUIRE(h != nullptr, "hb_accum_reset: NULL handle");
  Accum* acc = reinterpret_cast<Accum*>(h);
  HB_CUDA(cudaSetDevice(acc->device));
  const int threads = 256;
  reset_state_kernel<<<(int)((acc->nchains + threads - 1) / threads), threads>>>(acc->state,
                                                                                 acc->nchains);
  reset_globals_kernel<<<1, 32>>>(acc->globals);
  count_launch(2);
  HB_CUDA(cudaGetLastError());
  return HB_OK;
}

int hb_accumulate_precomputed(hb_accum* h, const void* lnpred, int lnpred_dtype,
                              const void* ln_posterior, int lp_dtype, int mem,
                              const int64_t* start_indices, int64_t chain_lo, int64_t chain_hi,
                              int flags, void* stream) {
  HB_REQUIRE(h != nullptr, "hb_accumulate_precomputed: NULL handle");
  Accum* acc = reinterpret_cast<Accum*>(h);
  HB_REQUIRE(start_indices != nullptr, "start_indices is NULL");
  HB_REQUIRE(0 <= chain_lo && chain_lo <= chain_hi && chain_hi <= acc->nchains,
             "nchains do not match");
  HB_REQUIRE((lnpred_dtype == HB_F32 || lnpred_dtype == HB_F64) &&
                 (lp_dtype == HB_F32 || lp_dtype == HB_F64),
             "dtype must be HB_F32 or HB_F64");
  HB_CUDA(cudaSetDevice(acc->device));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (flags & HB_ACC_NEW_CALL) {
    int rc = reset_call_scalars(acc, st);
    if (rc) return rc;
  }
  const long long nrows = start_indices[chain_hi] - start_indices[chain_lo];
  if (nrows == 0 || chain_lo == chain_hi) return HB_OK;
  HB_REQUIRE(lnpred != nullptr && ln_posterior != nullptr, "NULL input buffer");
  if (mem == HB_MEM_DEVICE)
    return accumulate_precomputed_device(acc, lnpred, lnpred_dtype, ln_posterior, lp_dtype,
                                         start_indices, chain_lo, chain_hi, 0, nrows, st);
  // HOST buffers: chunked, double-buffered staging overlapped with the fold kernel
  if (!is_pinned_host(lnpred) || !is_pinned_host(ln_posterior)) {
    // pageable memory: host threads convert to float32 into pinned staging (see hb_accumulate_flow)
    const long long chunkp = 1LL << 23;
    const size_t offl = (size_t)chunkp * 4;
    int rcp = ensure_stage(acc, 2 * offl);
    if (rcp) return rcp;
    rcp = ensure_pinned(acc, 2 * offl);
    if (rcp) return rcp;
    const int nth = host_threads();
    int ip = 0;
    for (long long a = 0; a < nrows; a += chunkp, ++ip) {
      const long long b = (a + chunkp < nrows) ? a + chunkp : nrows;
      const int sl = ip & 1;
      char* hp = static_cast<char*>(acc->pin[sl]);
      char* d = static_cast<char*>(acc->dev[sl]);
      HB_CUDA(cudaEventSynchronize(acc->ev_copy[sl]));
      if (lnpred_dtype == HB_F64) convert_rows((const double*)lnpred, 1, 1, a, b, (float*)hp, nth);
      else convert_rows((const float*)lnpred, 1, 1, a, b, (float*)hp, nth);
      if (lp_dtype == HB_F64) convert_rows((const double*)ln_posterior, 1, 1, a, b, (float*)(hp + offl), nth);
      else convert_rows((const float*)ln_posterior, 1, 1, a, b, (float*)(hp + offl), nth);
      HB_CUDA(cudaStreamWaitEvent(acc->copy_stream, acc->ev_done[sl], 0));
      HB_CUDA(cudaMemcpyAsync(d, hp, offl + (size_t)(b - a) * 4, cudaMemcpyHostToDevice, acc->copy_stream));
      HB_CUDA(cudaEventRecord(acc->ev_copy[sl], acc->copy_stream));
      HB_CUDA(cudaStreamWaitEvent(st, acc->ev_copy[sl], 0));
      rcp = accumulate_precomputed_device(acc, d, HB_F32, d + offl, HB_F32, start_indices, chain_lo, chain_hi,
                                          a, b, st);
      if (rcp) return rcp;
      HB_CUDA(cudaEventRecord(acc->ev_done[sl], st));
    }
    HB_CUDA(cudaStreamSynchronize(st));
    return HB_OK;
  }
  const size_t es_p = lnpred_dtype == HB_F32 ? 4 : 8, es_l = lp_dtype == HB_F32 ? 4 : 8;
  const long long chunk = 1LL << 24;  // rows per chunk
  int rc = ensure_stage(acc, (size_t)chunk * (es_p + es_l));
  if (rc) return rc;
  int i = 0;
  for (long long a = 0; a < nrows; a += chunk, ++i) {
    const long long b = (a + chunk < nrows) ? a + chunk : nrows;
    const int sl = i & 1;
    char* d = static_cast<char*>(acc->dev[sl]);
    char* dl = d + (size_t)chunk * es_p;
    HB_CUDA(cudaStreamWaitEvent(acc->copy_stream, acc->ev_done[sl], 0));
    HB_CUDA(cudaMemcpyAsync(d, (const char*)lnpred + a * es_p, (b - a) * es_p,
                            cudaMemcpyHostToDevice, acc->copy_stream));
    HB_CUDA(cudaMemcpyAsync(dl, (const char*)ln_posterior + a * es_l, (b - a) * es_l,
                            cudaMemcpyHostToDevice, acc->copy_stream));
    HB_CUDA(cudaEventRecord(acc->ev_copy[sl], acc->copy_stream));
    HB_CUDA(cudaStreamWaitEvent(st, acc->ev_copy[sl], 0));
    rc = accumulate_precomputed_device(acc, d, lnpred_dtype, dl, lp_dtype, start_indices, chain_lo,
                                       chain_hi, a, b, st);
    if (rc) return rc;
    HB_CUDA(cudaEventRecord(acc->ev_done[sl], st));
  }
  HB_CUDA(cudaStreamSynchronize(st));
  return HB_OK;
}

int hb_accum_export(hb_accum* h, double* partials, double* globals) {
  HB_REQUIRE(h != nullptr, "hb_accum_export: NULL handle");
  Accum* acc = reinterpret_cast<Accum*>(h);
  HB_CUDA(cudaSetDevice(acc->device));
  HB_CUDA(cudaDeviceSynchronize());
  if (partials)
    HB_CUDA(cudaMemcpy(partials, acc->state, acc->nchains * sizeof(ChainState),
                       cudaMemcpyDeviceToHost));
  if (globals) HB_CUDA(cudaMemcpy(globals, acc->globals, G_N * sizeof(double), cudaMemcpyDeviceToHost));
  return HB_OK;
}

int hb_accum_merge(hb_accum* h, const double* partials, const double* globals) {
  HB_REQUIRE(h != nullptr && partials != nullptr, "hb_accum_merge: NULL argument");
  Accum* acc = reinterpret_cast<Accum*>(h);
  HB_CUDA(cudaSetDevice(acc->device));
  double *dpart = nullptr, *dg = nullptr;
  HB_CUDA(cudaMalloc(&dpart, acc->nchains * 4 * sizeof(double)));
  HB_CUDA(cudaMemcpy(dpart, partials, acc->nchains * 4 * sizeof(double), cudaMemcpyHostToDevice));
  if (globals) {
    HB_CUDA(cudaMalloc(&dg, G_N * sizeof(double)));
    HB_CUDA(cudaMemcpy(dg, globals, G_N * sizeof(double), cudaMemcpyHostToDevice));
  }
  const int threads = 256;
  merge_partials_kernel<<<(int)((acc->nchains + threads - 1) / threads), threads>>>(
      acc->state, acc->globals, dpart, dg, acc->nchains);
  count_launch();
  cudaError_t e = cudaDeviceSynchronize();
  cudaFree(dpart);
  if (dg) cudaFree(dg);
  HB_CUDA(e);
  return HB_OK;
}

int hb_accum_pack(hb_accum* h, double* dst, int mem, void* stream) {
  HB_REQUIRE(h != nullptr && dst != nullptr, "hb_accum_pack: NULL argument");
  Accum* acc = reinterpret_cast<Accum*>(h);
  HB_CUDA(cudaSetDevice(acc->device));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const cudaMemcpyKind kind = mem == HB_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
  HB_CUDA(cudaMemcpyAsync(dst, acc->state, acc->nchains * sizeof(ChainState), kind, st));
  HB_CUDA(cudaMemcpyAsync(dst + 4 * acc->nchains, acc->globals, G_N * sizeof(double), kind, st));
  if (mem != HB_MEM_DEVICE) HB_CUDA(cudaStreamSynchronize(st));
  return HB_OK;
}

int hb_accum_merge_packed(hb_accum* h, int nranks, const double* packed, int64_t stride, int mem,
                          void* stream) {
  HB_REQUIRE(h != nullptr && packed != nullptr && nranks >= 1, "hb_accum_merge_packed: bad argument");
  Accum* acc = reinterpret_cast<Accum*>(h);
  HB_REQUIRE(stride >= 4 * acc->nchains + G_N, "hb_accum_merge_packed: stride too small");
  HB_CUDA(cudaSetDevice(acc->device));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const double* src = packed;
  double* tmp = nullptr;
  if (mem != HB_MEM_DEVICE) {
    const size_t bytes = (size_t)nranks * stride * sizeof(double);
    HB_CUDA(cudaMalloc(&tmp, bytes));
    cudaError_t e = cudaMemcpyAsync(tmp, packed, bytes, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) {
      cudaFree(tmp);
      return cuda_fail(e, "H2D", __FILE__, __LINE__);
    }
    src = tmp;
  }
  const int threads = 256;
  merge_packed_kernel<<<(int)((acc->nchains + threads - 1) / threads), threads, 0, st>>>(
      acc->state, acc->globals, src, stride, nranks, acc->nchains);
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (tmp) {
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(tmp);
  }
  HB_CUDA(e);
  return HB_OK;
}

int hb_finalize(hb_accum* h, int shift_mode, double shift_value_in, hb_result* out,
                double* ln_inv_per_chain, double* running_sum, double* nsamples_per_chain,
                double* nsamples_eff_per_chain) {
  HB_REQUIRE(h != nullptr && out != nullptr, "hb_finalize: NULL argument");
  HB_REQUIRE(shift_mode >= HB_MEAN_SHIFT && shift_mode <= HB_ABS_MAX_SHIFT, "unknown shift mode");
  Accum* acc = reinterpret_cast<Accum*>(h);
  HB_CUDA(cudaSetDevice(acc->device));
  HB_CUDA(cudaDeviceSynchronize());  // accumulate launches may sit on caller streams
  finalize_kernel<<<1, FIN_THREADS>>>(acc->state, acc->globals, acc->nchains, shift_mode,
                                      shift_value_in, acc->fin);
  count_launch();
  HB_CUDA(cudaGetLastError());
  const size_t off = sizeof(hb_result) / sizeof(double);
  const size_t n = (size_t)acc->nchains;
  const bool want_arrays = ln_inv_per_chain || running_sum || nsamples_per_chain || nsamples_eff_per_chain;
  static thread_local std::vector<double> host;
  host.resize(off + (want_arrays ? 4 * n : 0));
  HB_CUDA(cudaMemcpy(host.data(), acc->fin, host.size() * sizeof(double), cudaMemcpyDeviceToHost));  // one D2H
  memcpy(out, host.data(), sizeof(hb_result));
  if (ln_inv_per_chain) memcpy(ln_inv_per_chain, host.data() + off, n * sizeof(double));
  if (running_sum) memcpy(running_sum, host.data() + off + n, n * sizeof(double));
  if (nsamples_per_chain) memcpy(nsamples_per_chain, host.data() + off + 2 * n, n * sizeof(double));
  if (nsamples_eff_per_chain) memcpy(nsamples_eff_per_chain, host.data() + off + 3 * n, n * sizeof(double));
  return HB_OK;
}

int hb_finalize_realspace(int device, int64_t nchains, const double* running_sum,
                          const double* nsamples_per_chain, double shift_value, hb_result* out,
                          double* ln_inv_per_chain) {
  HB_REQUIRE(running_sum && nsamples_per_chain && out, "hb_finalize_realspace: NULL argument");
  HB_REQUIRE(nchains >= 1, "nchains must be greater than 0.");
  int rc = require_device(device);
  if (rc) return rc;
  HB_CUDA(cudaSetDevice(device));
  double* buf = nullptr;
  const size_t head = sizeof(hb_result) / sizeof(double);
  HB_CUDA(cudaMalloc(&buf, (head + 3 * (size_t)nchains) * sizeof(double)));
  double* dS = buf + head + nchains;
  double* dn = dS + nchains;
  cudaError_t e = cudaMemcpy(dS, running_sum, nchains * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(dn, nsamples_per_chain, nchains * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    finalize_realspace_kernel<<<1, FIN_THREADS>>>(dS, dn, nchains, shift_value, buf);
    count_launch();
    e = cudaDeviceSynchronize();
  }
  if (e == cudaSuccess) e = cudaMemcpy(out, buf, sizeof(hb_result), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && ln_inv_per_chain)
    e = cudaMemcpy(ln_inv_per_chain, buf + head, nchains * sizeof(double), cudaMemcpyDeviceToHost);
  cudaFree(buf);
  HB_CUDA(e);
  return HB_OK;
}

}  // extern "C"
