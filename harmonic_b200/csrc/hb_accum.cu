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

// rows r .. r+3 and r + 4*ACC_THREADS .. +3 in the element type of the buffer: each 128-bit load is fully
// coalesced across the warp
__device__ __forceinline__ void load8(const float* __restrict__ p, long long r, float (&v)[8]) {
  const float4* q = reinterpret_cast<const float4*>(p + r);
  float4 a = __ldcs(q), b = __ldcs(q + 256);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8(const double* __restrict__ p, long long r, double (&v)[8]) {
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    const double2* q = reinterpret_cast<const double2*>(p + r + (long long)g * 4 * 256);
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      double2 a = __ldcs(q + k);
      v[4 * g + 2 * k] = a.x;
      v[4 * g + 2 * k + 1] = a.y;
    }
  }
}

// lnarg = lnpred - ln_posterior.  With a float64 operand the difference is formed in double and rounded to
// float32 once (|ln_posterior| ~ 1e4 would otherwise cost ~1e-3 of cancellation error per lnarg): what numpy
// does in the reference's per-sample branch (harmonic/evidence.py:285-303).
template <typename PT, typename LT, bool VEC>
__global__ void __launch_bounds__(ACC_THREADS, 6)
accumulate_precomputed_kernel(const PT* __restrict__ lnpred, const LT* __restrict__ lnpost,
                              FoldParams p) {
  __shared__ double scratch[3 * (ACC_THREADS / 32)];
  constexpr bool WIDE = sizeof(PT) == 8 || sizeof(LT) == 8;
  const long long row_lo = (long long)blockIdx.x * p.rows_per_flusher;
  long long row_hi = row_lo + p.rows_per_flusher;
  if (row_hi > p.nrows) row_hi = p.nrows;
  const bool had_rows = row_lo < p.nrows;
  Fold<ACC_THREADS> fold(p, blockIdx.x, threadIdx.x, scratch, had_rows ? row_lo : 0);
  for (long long t0 = row_lo; t0 < row_hi; t0 += ACC_TILE) {
    long long t1 = t0 + ACC_TILE;
    PT pa[ACC_R];
    LT pb[ACC_R];
    const long long r0 = t0 + (long long)threadIdx.x * 4;
    if (VEC && t1 <= row_hi) {
      load8(lnpred, r0, pa);
      load8(lnpost, r0, pb);
    } else {
      if (t1 > row_hi) t1 = row_hi;
#pragma unroll
      for (int q = 0; q < ACC_R; ++q) {
        const long long r = r0 + (long long)(q / 4) * 4 * ACC_THREADS + (q % 4);
        pa[q] = (r < t1) ? lnpred[r] : (PT)0;
        pb[q] = (r < t1) ? lnpost[r] : (LT)0;
      }
    }
    float a[ACC_R], b[ACC_R], d[ACC_R];
#pragma unroll
    for (int q = 0; q < ACC_R; ++q) {
      a[q] = (float)pa[q];
      b[q] = (float)pb[q];
      if (WIDE) d[q] = (float)((double)pa[q] - (double)pb[q]);
    }
    fold.tile<ACC_R, 4>(t0, t1, a, b, WIDE ? d : nullptr);
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

__global__ void __launch_bounds__(256)
merge_kernel(ChainState* __restrict__ state, double* __restrict__ globals, FoldParams p, long long chain_lo,
             long long nflushers, int overwrite_globals) {
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
  // per-call scalars: block 0 reduces the [nflushers][G_N] partials.  Thread t takes rows t, t + 256, ...
  // (independent 64-byte loads), then a shuffle tree and one pass over the 8 warps: a fixed order, so the
  // result is reproducible.
  if (blockIdx.x == 0) {
    __shared__ double sh[8][G_N];
    double acc[G_N];
#pragma unroll
    for (int k = 0; k < G_N; ++k) acc[k] = (k < 2) ? 0.0 : ((k & 1) ? CUDART_INF : -CUDART_INF);
    for (long long f = threadIdx.x; f < nflushers; f += 256) {
      const double2* row = reinterpret_cast<const double2*>(p.gpart + f * G_N);
#pragma unroll
      for (int k = 0; k < G_N; k += 2) {
        const double2 o = row[k >> 1];
        if (k < 2) {
          acc[k] += o.x;
          acc[k + 1] += o.y;
        } else {
          acc[k] = fmax(acc[k], o.x);          // even k = maxima
          acc[k + 1] = fmin(acc[k + 1], o.y);  // odd k = minima
        }
      }
    }
#pragma unroll
    for (int k = 0; k < G_N; ++k) {
#pragma unroll
      for (int o = 16; o; o >>= 1) {
        const double v = __shfl_xor_sync(0xffffffffu, acc[k], o);
        if (k < 2) acc[k] += v;
        else if (k & 1) acc[k] = fmin(acc[k], v);
        else acc[k] = fmax(acc[k], v);
      }
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
      for (int k = 0; k < G_N; ++k) sh[threadIdx.x >> 5][k] = acc[k];
    }
    __syncthreads();
    if (threadIdx.x < G_N) {
      const int k = threadIdx.x;
      double r = sh[0][k];
      for (int w = 1; w < 8; ++w) {
        const double o = sh[w][k];
        if (k < 2) r += o;
        else if (k & 1) r = fmin(r, o);
        else r = fmax(r, o);
      }
      if (!overwrite_globals) {
        const double g = globals[k];
        r = (k < 2) ? g + r : ((k & 1) ? fmin(g, r) : fmax(g, r));
      }
      globals[k] = r;
    }
  }
}

// fold nranks packed blocks ([4n partials | 8 globals], block r at packed + r*stride) in rank order;
// overwrite != 0: the result REPLACES the state (the all-gather target needs no separate reset)
__global__ void merge_packed_kernel(ChainState* __restrict__ state, double* __restrict__ globals,
                                    const double* __restrict__ packed, long long stride, int nranks,
                                    long long n, int overwrite) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c < n) {
    ChainState st;
    if (overwrite) {
      st.m = -CUDART_INF;
      st.s = st.n = st.nnan = 0.0;
    } else {
      st = state[c];
    }
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
    double acc = overwrite ? ((k < 2) ? 0.0 : ((k & 1) ? CUDART_INF : -CUDART_INF)) : globals[k];
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

// the whole packed block [n x (m, s, n, n_nan) | G_N scalars] in one launch
__global__ void reset_packed_kernel(double* packed, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 4 * n) packed[i] = ((i & 3) == 0) ? -CUDART_INF : 0.0;
  else if (i < 4 * n + G_N) {
    const int k = (int)(i - 4 * n);
    packed[i] = (k < 2) ? 0.0 : ((k & 1) ? CUDART_INF : -CUDART_INF);
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
constexpr int FIN_THREADS = 1024;
constexpr int FIN_WARPS = FIN_THREADS / 32;

__device__ double block_max(double v, double* sh) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = sh[0];
  for (int w = 1; w < FIN_WARPS; ++w) r = fmax(r, sh[w]);
  return r;
}
__device__ double block_sum(double v, double* sh) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = 0.0;
  for (int w = 0; w < FIN_WARPS; ++w) r += sh[w];
  return r;
}
// NaN-propagating logsumexp of f(c) over chains (scipy.special.logsumexp semantics); f must be cheap (it is
// evaluated twice per chain: the callers read values stored by an earlier pass)
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

// One CTA of 1024 threads.  The four per-chain output arrays double as scratch so that every logarithm /
// exponential of a chain is evaluated once (float64 transcendentals are what this kernel costs: 1e4 chains took
// 236 us with 256 threads and per-use recomputation, see profiles/README.md).
__global__ void __launch_bounds__(FIN_THREADS)
finalize_kernel(const ChainState* __restrict__ state, const double* __restrict__ globals,
                long long nch, int shift_mode, double shift_in, double* __restrict__ out) {
  __shared__ double sh[FIN_WARPS];
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

  // pass 1: L_c -> rsum[] (scratch), ln n_c -> neff_pc[] (scratch); totals
  double nsum = 0.0, n2sum = 0.0, nansum = 0.0;
  for (long long c = threadIdx.x; c < nch; c += FIN_THREADS) {
    const ChainState st = state[c];
    nsum += st.n;
    n2sum += st.n * st.n;
    nansum += st.nnan;
    rsum[c] = (st.s > 0.0) ? st.m + log(st.s) : -CUDART_INF;
    neff_pc[c] = log(st.n);
  }
  const double N = block_sum(nsum, sh);
  const double N2 = block_sum(n2sum, sh);
  const double NNAN = block_sum(nansum, sh);
  const double lnN = log(N);
  const double ln_rho = block_lse(nch, [&](long long c) { return rsum[c]; }, sh) - lnN;
  // pass 2: ln z_c -> ln_inv_pc[] (scratch)
  const bool all_zero = isinf(ln_rho) && ln_rho < 0;  // all sums zero
  for (long long c = threadIdx.x; c < nch; c += FIN_THREADS)
    ln_inv_pc[c] = all_zero ? -CUDART_INF : ln_rho + log(fabs(expm1(rsum[c] - neff_pc[c] - ln_rho)));
  __syncthreads();
  const double lnV = block_lse(nch, [&](long long c) { return 2.0 * ln_inv_pc[c] + neff_pc[c]; }, sh) - lnN;
  const double lnK =
      block_lse(nch, [&](long long c) { return 4.0 * ln_inv_pc[c] + neff_pc[c]; }, sh) - lnN - 2.0 * lnV;
  const double n_eff = N * N / N2;
  __syncthreads();
  // pass 3: the public per-chain arrays replace the scratch values
  for (long long c = threadIdx.x; c < nch; c += FIN_THREADS) {
    const ChainState st = state[c];
    const double Lc = rsum[c];
    ln_inv_pc[c] = Lc - neff_pc[c];
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
    res->mean_nsamples_eff = (N - NNAN) / (double)nch;
    res->guard_trips = 0.0;  // filled in by the host (Accum::guard_trips)
    res->reserved = 0.0;
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
    res->mean_nsamples_eff = N / (double)nch;
    res->guard_trips = 0.0;
    res->reserved = 0.0;
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

// every entry point that touches the accumulator on `st` orders itself behind a pending reset
// (hb_accum_create / hb_accum_reset without a stream run it on the legacy default stream)
int order_after_reset(Accum* acc, cudaStream_t st) {
  if (acc->reset_pending) {
    HB_CUDA(cudaStreamWaitEvent(st, acc->ev_reset, 0));
    acc->reset_pending = false;
  }
  return HB_OK;
}

int launch_merge(Accum* acc, const FoldParams& p, long long chain_lo, long long nflushers,
                 cudaStream_t st) {
  const int threads = 256;
  const int blocks = (int)((p.nch + threads - 1) / threads);
  merge_kernel<<<blocks > 0 ? blocks : 1, threads, 0, st>>>(acc->state, acc->globals, p, chain_lo,
                                                             nflushers, 0);
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

// First launch of an add_chains call: the per-call scalars start afresh.
int begin_call(Accum* acc, bool has_rows, cudaStream_t st) {
  (void)has_rows;
  int rc = order_after_reset(acc, st);
  if (rc) return rc;
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

// ---- staging buffers: a per-process free list, borrowed for the duration of one host-input call ----
static std::mutex g_stage_mu;
static std::vector<Stage*> g_stage_free;

static void stage_free_buffers(Stage* s) {
  cudaSetDevice(s->device);
  for (int i = 0; i < 2; ++i) {
    if (s->dev[i]) cudaFree(s->dev[i]);
    if (s->pin[i]) cudaFreeHost(s->pin[i]);
    if (s->ev_copy[i]) cudaEventDestroy(s->ev_copy[i]);
    if (s->ev_done[i]) cudaEventDestroy(s->ev_done[i]);
  }
  if (s->copy_stream) cudaStreamDestroy(s->copy_stream);
  delete s;
}

Stage* stage_acquire(int device) {
  {
    std::lock_guard<std::mutex> lk(g_stage_mu);
    for (size_t i = 0; i < g_stage_free.size(); ++i) {
      if (g_stage_free[i]->device == device) {
        Stage* s = g_stage_free[i];
        g_stage_free.erase(g_stage_free.begin() + i);
        return s;
      }
    }
  }
  Stage* s = new Stage();
  s->device = device;
  bool ok = cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking) == cudaSuccess;
  for (int i = 0; i < 2 && ok; ++i) {
    ok = cudaEventCreateWithFlags(&s->ev_copy[i], cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&s->ev_done[i], cudaEventDisableTiming) == cudaSuccess;
  }
  if (!ok) {
    cuda_fail(cudaGetLastError(), "staging stream / events", __FILE__, __LINE__);
    stage_free_buffers(s);
    return nullptr;
  }
  return s;
}

void stage_release(Stage* s, cudaStream_t st) {
  if (!s) return;
  // nothing of this call may still be in flight when the buffers change hands (error paths included)
  cudaStreamSynchronize(s->copy_stream);
  cudaStreamSynchronize(st);
  std::lock_guard<std::mutex> lk(g_stage_mu);
  // at most two parked stages per process keep their buffers (double-buffered pinned + device chunks,
  // <= ~0.7 GB each); hb_pool_trim() frees them
  if (g_stage_free.size() < 2) g_stage_free.push_back(s);
  else stage_free_buffers(s);
}

}  // namespace hb

using namespace hb;

// Released accumulators are parked here and handed out again for the same (device, nchains):
// creating one costs several cudaMalloc / event calls, which a per-call Evidence would otherwise pay
// on every add_chains.  They hold O(nchains + flushers) bytes only (staging lives in the Stage list);
// on a miss with a full pool the oldest entry is evicted.
static std::mutex g_pool_mu;
static std::vector<Accum*> g_pool;
constexpr size_t POOL_MAX = 8;

static void accum_free(Accum* acc) {
  cudaSetDevice(acc->device);
  cudaFree(acc->packed);
  cudaFree(acc->shadow);
  cudaFree(acc->start_dev);
  cudaFree(acc->recs);
  cudaFree(acc->gpart);
  cudaFree(acc->fin);
  if (acc->fin_host) cudaFreeHost(acc->fin_host);
  if (acc->ev_reset) cudaEventDestroy(acc->ev_reset);
  delete acc;
}

static int accum_reset_on(Accum* acc, cudaStream_t st) {
  const int threads = 256;
  const long long n = 4 * acc->nchains + G_N;
  reset_packed_kernel<<<(int)((n + threads - 1) / threads), threads, 0, st>>>(acc->packed, acc->nchains);
  count_launch();
  HB_CUDA(cudaGetLastError());
  acc->guard_trips = 0;
  return HB_OK;
}

extern "C" {

int hb_accum_create(int64_t nchains, int device, hb_accum** out) {
  HB_REQUIRE(out != nullptr, "hb_accum_create: out is NULL");
  HB_REQUIRE(nchains >= 1, "nchains must be greater than 0.");
  int rc = require_device(device);
  if (rc) return rc;
  HB_CUDA(cudaSetDevice(device));
  Accum* acc = nullptr;
  {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    for (size_t i = 0; i < g_pool.size(); ++i) {
      if (g_pool[i]->device == device && g_pool[i]->nchains == nchains) {
        acc = g_pool[i];
        g_pool.erase(g_pool.begin() + i);
        break;
      }
    }
  }
  if (!acc) {
    acc = new Accum();
    acc->device = device;
    acc->nchains = nchains;
    cudaError_t e = cudaMalloc(&acc->packed, (4 * nchains + G_N) * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&acc->start_dev, (nchains + 1) * sizeof(long long));
    if (e == cudaSuccess) e = cudaMalloc(&acc->fin, sizeof(hb_result) + 4 * nchains * sizeof(double));
    if (e == cudaSuccess)
      e = cudaHostAlloc(reinterpret_cast<void**>(&acc->fin_host), sizeof(hb_result) + 4 * nchains * sizeof(double),
                        cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&acc->ev_reset, cudaEventDisableTiming);
    if (e != cudaSuccess) {
      accum_free(acc);
      return cuda_fail(e, "accumulator allocation", __FILE__, __LINE__);
    }
    acc->state = reinterpret_cast<ChainState*>(acc->packed);
    acc->globals = acc->packed + 4 * nchains;
  }
  *out = reinterpret_cast<hb_accum*>(acc);
  return hb_accum_reset(*out, nullptr);
}

int hb_accum_destroy(hb_accum* h) {
  if (!h) return HB_OK;
  Accum* acc = reinterpret_cast<Accum*>(h);
  cudaSetDevice(acc->device);
  cudaDeviceSynchronize();
  Accum* evict = nullptr;
  {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    if (g_pool.size() >= POOL_MAX) {
      evict = g_pool.front();
      g_pool.erase(g_pool.begin());
    }
    g_pool.push_back(acc);
  }
  if (evict) accum_free(evict);
  return HB_OK;
}

int hb_pool_trim(void) {
  std::vector<Accum*> accs;
  std::vector<Stage*> stages;
  {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    accs.swap(g_pool);
  }
  {
    std::lock_guard<std::mutex> lk(g_stage_mu);
    stages.swap(g_stage_free);
  }
  for (Accum* a : accs) accum_free(a);
  for (Stage* s : stages) stage_free_buffers(s);
  return HB_OK;
}

int hb_accum_reset(hb_accum* h, void* stream) {
  HB_REQUIRE(h != nullptr, "hb_accum_reset: NULL handle");
  Accum* acc = reinterpret_cast<Accum*>(h);
  HB_CUDA(cudaSetDevice(acc->device));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int rc = accum_reset_on(acc, st);
  if (rc) return rc;
  // later calls may arrive on any stream: they wait on this event (order_after_reset)
  HB_CUDA(cudaEventRecord(acc->ev_reset, st));
  acc->reset_pending = true;
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
  const long long nrows = start_indices[chain_hi] - start_indices[chain_lo];
  const bool has_rows = nrows > 0 && chain_lo < chain_hi;
  int rc = (flags & HB_ACC_NEW_CALL) ? begin_call(acc, has_rows, st) : order_after_reset(acc, st);
  if (rc) return rc;
  if (!has_rows) return HB_OK;
  HB_REQUIRE(lnpred != nullptr && ln_posterior != nullptr, "NULL input buffer");
  if (mem == HB_MEM_DEVICE)
    return accumulate_precomputed_device(acc, lnpred, lnpred_dtype, ln_posterior, lp_dtype,
                                         start_indices, chain_lo, chain_hi, 0, nrows, st);
  // HOST buffers: chunked, double-buffered staging overlapped with the fold kernel.  Pinned buffers are
  // copied straight from the caller's memory; pageable ones go through pinned staging filled by host threads.
  // The element types travel unchanged (float64 stays float64: the kernel forms lnpred - ln_posterior in
  // double before rounding to float32, as numpy does in the reference's per-sample branch,
  // harmonic/evidence.py:285-303).
  StageLease lease(acc->device, st);
  Stage* sg = lease.get();
  if (!sg) return HB_ERR_CUDA;
  const bool pinned = is_pinned_host(lnpred) && is_pinned_host(ln_posterior);
  const size_t es_p = lnpred_dtype == HB_F32 ? 4 : 8, es_l = lp_dtype == HB_F32 ? 4 : 8;
  const long long chunk = pinned ? (1LL << 24) : (1LL << 22);  // rows per chunk
  const size_t off_l = (size_t)chunk * es_p;
  rc = ensure_stage(sg, (size_t)chunk * (es_p + es_l));
  if (rc) return rc;
  if (!pinned) {
    rc = ensure_pinned(sg, (size_t)chunk * (es_p + es_l));
    if (rc) return rc;
  }
  const int nth = host_threads();
  int i = 0;
  for (long long a = 0; a < nrows; a += chunk, ++i) {
    const long long b = (a + chunk < nrows) ? a + chunk : nrows;
    const int sl = i & 1;
    char* d = static_cast<char*>(sg->dev[sl]);
    char* dl = d + off_l;
    if (pinned) {
      HB_CUDA(cudaStreamWaitEvent(sg->copy_stream, sg->ev_done[sl], 0));
      HB_CUDA(cudaMemcpyAsync(d, (const char*)lnpred + a * es_p, (b - a) * es_p, cudaMemcpyHostToDevice,
                              sg->copy_stream));
      HB_CUDA(cudaMemcpyAsync(dl, (const char*)ln_posterior + a * es_l, (b - a) * es_l, cudaMemcpyHostToDevice,
                              sg->copy_stream));
    } else {
      char* hp = static_cast<char*>(sg->pin[sl]);
      HB_CUDA(cudaEventSynchronize(sg->ev_copy[sl]));  // the pinned slot's previous copy has left
      copy_bytes((const char*)lnpred + a * es_p, hp, (size_t)(b - a) * es_p, nth);
      copy_bytes((const char*)ln_posterior + a * es_l, hp + off_l, (size_t)(b - a) * es_l, nth);
      HB_CUDA(cudaStreamWaitEvent(sg->copy_stream, sg->ev_done[sl], 0));
      HB_CUDA(cudaMemcpyAsync(d, hp, off_l + (size_t)(b - a) * es_l, cudaMemcpyHostToDevice, sg->copy_stream));
    }
    HB_CUDA(cudaEventRecord(sg->ev_copy[sl], sg->copy_stream));
    HB_CUDA(cudaStreamWaitEvent(st, sg->ev_copy[sl], 0));
    rc = accumulate_precomputed_device(acc, d, lnpred_dtype, dl, lp_dtype, start_indices, chain_lo,
                                       chain_hi, a, b, st);
    if (rc) return rc;
    HB_CUDA(cudaEventRecord(sg->ev_done[sl], st));
  }
  HB_CUDA(cudaStreamSynchronize(st));
  return HB_OK;
}

int hb_accum_export(hb_accum* h, double* partials, double* globals, void* stream) {
  HB_REQUIRE(h != nullptr, "hb_accum_export: NULL handle");
  Accum* acc = reinterpret_cast<Accum*>(h);
  HB_CUDA(cudaSetDevice(acc->device));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int rc = order_after_reset(acc, st);
  if (rc) return rc;
  if (partials)
    HB_CUDA(cudaMemcpyAsync(partials, acc->state, acc->nchains * sizeof(ChainState), cudaMemcpyDeviceToHost, st));
  if (globals)
    HB_CUDA(cudaMemcpyAsync(globals, acc->globals, G_N * sizeof(double), cudaMemcpyDeviceToHost, st));
  HB_CUDA(cudaStreamSynchronize(st));
  return HB_OK;
}

int hb_accum_merge(hb_accum* h, const double* partials, const double* globals, void* stream) {
  HB_REQUIRE(h != nullptr && partials != nullptr, "hb_accum_merge: NULL argument");
  Accum* acc = reinterpret_cast<Accum*>(h);
  HB_CUDA(cudaSetDevice(acc->device));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int rc = order_after_reset(acc, st);
  if (rc) return rc;
  double *dpart = nullptr, *dg = nullptr;
  HB_CUDA(cudaMalloc(&dpart, (acc->nchains * 4 + G_N) * sizeof(double)));
  cudaError_t e = cudaMemcpyAsync(dpart, partials, acc->nchains * 4 * sizeof(double), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess && globals) {
    dg = dpart + acc->nchains * 4;
    e = cudaMemcpyAsync(dg, globals, G_N * sizeof(double), cudaMemcpyHostToDevice, st);
  }
  if (e == cudaSuccess) {
    const int threads = 256;
    merge_partials_kernel<<<(int)((acc->nchains + threads - 1) / threads), threads, 0, st>>>(
        acc->state, acc->globals, dpart, dg, acc->nchains);
    count_launch();
    e = cudaGetLastError();
  }
  const cudaError_t e2 = cudaStreamSynchronize(st);
  cudaFree(dpart);
  HB_CUDA(e);
  HB_CUDA(e2);
  return HB_OK;
}

int hb_accum_packed_ptr(hb_accum* h, double** out, int64_t* n_doubles) {
  HB_REQUIRE(h != nullptr && out != nullptr, "hb_accum_packed_ptr: NULL argument");
  Accum* acc = reinterpret_cast<Accum*>(h);
  *out = acc->packed;
  if (n_doubles) *n_doubles = 4 * acc->nchains + G_N;
  return HB_OK;
}

int hb_accum_pack(hb_accum* h, double* dst, int mem, void* stream) {
  HB_REQUIRE(h != nullptr && dst != nullptr, "hb_accum_pack: NULL argument");
  Accum* acc = reinterpret_cast<Accum*>(h);
  HB_CUDA(cudaSetDevice(acc->device));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int rc = order_after_reset(acc, st);
  if (rc) return rc;
  const cudaMemcpyKind kind = mem == HB_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
  HB_CUDA(cudaMemcpyAsync(dst, acc->packed, (4 * acc->nchains + G_N) * sizeof(double), kind, st));
  if (mem != HB_MEM_DEVICE) HB_CUDA(cudaStreamSynchronize(st));
  return HB_OK;
}

int hb_accum_merge_packed(hb_accum* h, int nranks, const double* packed, int64_t stride, int mem,
                          int flags, void* stream) {
  HB_REQUIRE(h != nullptr && packed != nullptr && nranks >= 1, "hb_accum_merge_packed: bad argument");
  Accum* acc = reinterpret_cast<Accum*>(h);
  HB_REQUIRE(stride >= 4 * acc->nchains + G_N, "hb_accum_merge_packed: stride too small");
  HB_CUDA(cudaSetDevice(acc->device));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int rc = order_after_reset(acc, st);
  if (rc) return rc;
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
      acc->state, acc->globals, src, stride, nranks, acc->nchains, (flags & HB_MERGE_OVERWRITE) ? 1 : 0);
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
                double* nsamples_eff_per_chain, void* stream) {
  HB_REQUIRE(h != nullptr && out != nullptr, "hb_finalize: NULL argument");
  HB_REQUIRE(shift_mode >= HB_MEAN_SHIFT && shift_mode <= HB_ABS_MAX_SHIFT, "unknown shift mode");
  Accum* acc = reinterpret_cast<Accum*>(h);
  HB_CUDA(cudaSetDevice(acc->device));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int rc = order_after_reset(acc, st);
  if (rc) return rc;
  // stream-ordered behind the accumulate / merge launches of the same stream: no device-wide sync
  finalize_kernel<<<1, FIN_THREADS, 0, st>>>(acc->state, acc->globals, acc->nchains, shift_mode,
                                             shift_value_in, acc->fin);
  count_launch();
  HB_CUDA(cudaGetLastError());
  const size_t off = sizeof(hb_result) / sizeof(double);
  const size_t n = (size_t)acc->nchains;
  const bool want_arrays = ln_inv_per_chain || running_sum || nsamples_per_chain || nsamples_eff_per_chain;
  // one D2H into the pinned mirror: the 128-byte result alone unless the per-chain arrays are asked for
  // (hb_finalize_arrays fetches them later)
  HB_CUDA(cudaMemcpyAsync(acc->fin_host, acc->fin, (off + (want_arrays ? 4 * n : 0)) * sizeof(double),
                          cudaMemcpyDeviceToHost, st));
  HB_CUDA(cudaStreamSynchronize(st));
  memcpy(out, acc->fin_host, sizeof(hb_result));
  out->guard_trips = (double)acc->guard_trips;
  const double* host = acc->fin_host;
  if (ln_inv_per_chain) memcpy(ln_inv_per_chain, host + off, n * sizeof(double));
  if (running_sum) memcpy(running_sum, host + off + n, n * sizeof(double));
  if (nsamples_per_chain) memcpy(nsamples_per_chain, host + off + 2 * n, n * sizeof(double));
  if (nsamples_eff_per_chain) memcpy(nsamples_eff_per_chain, host + off + 3 * n, n * sizeof(double));
  return HB_OK;
}

int hb_finalize_arrays(hb_accum* h, double* ln_inv_per_chain, double* running_sum, double* nsamples_per_chain,
                       double* nsamples_eff_per_chain, void* stream) {
  HB_REQUIRE(h != nullptr, "hb_finalize_arrays: NULL handle");
  Accum* acc = reinterpret_cast<Accum*>(h);
  HB_CUDA(cudaSetDevice(acc->device));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t off = sizeof(hb_result) / sizeof(double);
  const size_t n = (size_t)acc->nchains;
  HB_CUDA(cudaMemcpyAsync(acc->fin_host + off, acc->fin + off, 4 * n * sizeof(double), cudaMemcpyDeviceToHost, st));
  HB_CUDA(cudaStreamSynchronize(st));
  const double* host = acc->fin_host;
  if (ln_inv_per_chain) memcpy(ln_inv_per_chain, host + off, n * sizeof(double));
  if (running_sum) memcpy(running_sum, host + off + n, n * sizeof(double));
  if (nsamples_per_chain) memcpy(nsamples_per_chain, host + off + 2 * n, n * sizeof(double));
  if (nsamples_eff_per_chain) memcpy(nsamples_eff_per_chain, host + off + 3 * n, n * sizeof(double));
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
