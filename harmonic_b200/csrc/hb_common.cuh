// Shared device/host helpers for libharmonic_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include <string>

#include "../../include/harmonic_b200.h"

namespace hb {

// ------------------------------------------------------------------------------------------
// host-side error plumbing
// ------------------------------------------------------------------------------------------
void set_error(const std::string& msg);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);
void count_launch(int n = 1);

#define HB_CUDA(call)                                                         \
  do {                                                                        \
    cudaError_t _e = (call);                                                  \
    if (_e != cudaSuccess) return hb::cuda_fail(_e, #call, __FILE__, __LINE__); \
  } while (0)

#define HB_REQUIRE(cond, msg)   \
  do {                          \
    if (!(cond)) {              \
      hb::set_error(msg);       \
      return HB_ERR_INVALID;    \
    }                           \
  } while (0)

// ------------------------------------------------------------------------------------------
// per-chain streaming state and launch partials
// ------------------------------------------------------------------------------------------
// state[c] = {m, s, n, n_nan}: ln sum_i exp(lnarg_i) = m + ln s over the finite lnargs of chain c
// (the shift-free form of Evidence.running_sum, harmonic/evidence.py:321-334), n = all samples
// (NaNs included, :337-338), n_nan (:342-345).
struct ChainState {
  double m, s, n, nnan;
};
// one record per (chain, flusher) pair written by the accumulate kernels; merged in a fixed
// order by merge_kernel so the result does not depend on scheduling.
struct ChainRec {
  double m, s, nnan, valid;
};
enum { G_SUM = 0, G_CNT, G_ARGMAX, G_ARGMIN, G_PROBMAX, G_PROBMIN, G_PREDMAX, G_PREDMIN, G_N };

struct FoldParams {
  const long long* start;  // device, nch+1 local row offsets of the chains of this launch
  int nch;
  long long nrows;
  long long rows_per_flusher;  // contiguous rows owned by one flusher (multiple of its tile)
  ChainRec* recs;              // [nch + nflushers], zeroed before the launch
  double* gpart;               // [nflushers][G_N]
};

#ifdef __CUDACC__
// ------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__device__ __forceinline__ float nanmaxf(float a, float b) { return fmaxf(a, b); }  // fmaxf drops NaN
__device__ __forceinline__ float nanminf(float a, float b) { return fminf(a, b); }

// Fold: every thread of a "flusher" group (NT threads synchronised on a named barrier) feeds
// the rows it owns; rows of one flusher are a contiguous range, walked tile by tile in order.
// Implements lnarg = lnpred - lnpost, +-inf -> NaN (harmonic/evidence.py:278-283), NaN counting
// (:342-345), the per-chain sums (:321-334, as a max-rescaled online logsumexp) and the six
// nan-aware extrema + mean-shift sums (:307-319,347-352).
template <int NT>
struct Fold {
  // uniform across the group
  const FoldParams& p;
  int flusher;
  int cur_chain;
  long long cur_end;  // local row where cur_chain ends
  int tid;            // 0..NT-1 within the group
  int bar;            // named barrier id private to the group (0 = whole CTA when NT == blockDim)
  double* scratch;    // shared, >= 3*(NT/32) doubles, private to the group
  // per thread
  float m, s;
  int nnan;
  float gsum;
  int gcnt;
  float amax, amin, pmax, pmin, qmax, qmin;

  __device__ Fold(const FoldParams& p_, int flusher_, int tid_, double* scratch_, long long row_lo,
                  int bar_ = 0)
      : p(p_), flusher(flusher_), tid(tid_), bar(bar_), scratch(scratch_) {
    m = -CUDART_INF_F;
    s = 0.f;
    nnan = 0;
    gsum = 0.f;
    gcnt = 0;
    amax = pmax = qmax = -CUDART_INF_F;
    amin = pmin = qmin = CUDART_INF_F;
    // chain containing row_lo: largest c with start[c] <= row_lo (then skip empty chains)
    int lo = 0, hi = p.nch;  // invariant: start[lo] <= row_lo
    while (hi - lo > 1) {
      int mid = (lo + hi) >> 1;
      if (p.start[mid] <= row_lo) lo = mid; else hi = mid;
    }
    cur_chain = lo;
    cur_end = (cur_chain < p.nch) ? p.start[cur_chain + 1] : (long long)0x7fffffffffffffffLL;
    while (cur_chain < p.nch && cur_end <= row_lo) {
      ++cur_chain;
      cur_end = (cur_chain < p.nch) ? p.start[cur_chain + 1] : (long long)0x7fffffffffffffffLL;
    }
  }

  // `diff` (optional): lnpred - lnpost already formed at higher precision by the caller (float64 inputs)
  __device__ __forceinline__ void sample(float lnpred, float lnpost, const float* diff = nullptr) {
    float a = diff ? *diff : lnpred - lnpost;
    qmax = nanmaxf(qmax, lnpred);
    qmin = nanminf(qmin, lnpred);
    pmax = nanmaxf(pmax, lnpost);
    pmin = nanminf(pmin, lnpost);
    if (isnan(a) || isinf(a)) {
      ++nnan;
    } else {
      if (a > m) {
        s = s * __expf(m - a) + 1.f;
        m = a;
      } else {
        s += __expf(a - m);
      }
      gsum += a;
      ++gcnt;
      amax = fmaxf(amax, a);
      amin = fminf(amin, a);
    }
  }

  // group-wide reduction of (m, s, nnan); result valid in tid 0. All NT threads must call.
  __device__ void flush() {
    constexpr int NW = NT / 32;
    float mw = m;
#pragma unroll
    for (int o = 16; o; o >>= 1) mw = fmaxf(mw, __shfl_xor_sync(0xffffffffu, mw, o));
    double sw = (s > 0.f) ? (double)s * exp((double)m - (double)mw) : 0.0;
    int nw = nnan;
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      sw += __shfl_xor_sync(0xffffffffu, sw, o);
      nw += __shfl_xor_sync(0xffffffffu, nw, o);
    }
    if ((tid & 31) == 0) {
      scratch[(tid >> 5) * 3 + 0] = (double)mw;
      scratch[(tid >> 5) * 3 + 1] = sw;
      scratch[(tid >> 5) * 3 + 2] = (double)nw;
    }
    named_bar_sync(bar, NT);
    if (tid == 0) {
      double M = -CUDART_INF, S = 0.0, NN = 0.0;
      for (int w = 0; w < NW; ++w) M = fmax(M, scratch[w * 3]);
      for (int w = 0; w < NW; ++w) {
        double sv = scratch[w * 3 + 1];
        if (sv > 0.0) S += sv * exp(scratch[w * 3] - M);
        NN += scratch[w * 3 + 2];
      }
      ChainRec r;
      r.m = M;
      r.s = S;
      r.nnan = NN;
      r.valid = 1.0;
      p.recs[cur_chain + flusher] = r;
    }
    named_bar_sync(bar, NT);
    m = -CUDART_INF_F;
    s = 0.f;
    nnan = 0;
  }

  // Whole tile inside the open chain (the common case): no per-row range tests, one rescale of
  // the running sum per thread and tile, predicated (branch-free) accumulation.
  template <int R>
  __device__ __forceinline__ void fold_full(const float (&lnpred)[R], const float (&lnpost)[R],
                                            const float* diff = nullptr) {
    float a[R];
    float lmax = -CUDART_INF_F;
#pragma unroll
    for (int q = 0; q < R; ++q) {
      a[q] = diff ? diff[q] : lnpred[q] - lnpost[q];
      qmax = fmaxf(qmax, lnpred[q]);
      qmin = fminf(qmin, lnpred[q]);
      pmax = fmaxf(pmax, lnpost[q]);
      pmin = fminf(pmin, lnpost[q]);
      const bool fin = fabsf(a[q]) < CUDART_INF_F;  // false for NaN and +-inf (evidence.py:282-283)
      const float af = fin ? a[q] : -CUDART_INF_F;
      lmax = fmaxf(lmax, af);
      amin = fminf(amin, fin ? a[q] : CUDART_INF_F);
      gsum += fin ? a[q] : 0.f;
      gcnt += fin ? 1 : 0;
      nnan += fin ? 0 : 1;
      a[q] = af;
    }
    amax = fmaxf(amax, lmax);
    if (lmax > m) {
      s *= __expf(m - lmax);  // m = -inf, s = 0 at the start: 0 * 0
      m = lmax;
    }
    if (m > -CUDART_INF_F) {
#pragma unroll
      for (int q = 0; q < R; ++q) s += __expf(a[q] - m);  // exp(-inf) = 0 for the excluded rows
    }
  }

  // Feed one tile [t0, t1) of the flusher's rows.  Thread owns R rows in groups of V consecutive
  // rows: row(q) = t0 + (q / V) * NT * V + tid * V + q % V  (V = R: R consecutive rows per thread;
  // V = 4, R = 8: two fully coalesced float4 loads per thread).
  template <int R, int V = R>
  __device__ __forceinline__ void tile(long long t0, long long t1, const float (&lnpred)[R],
                                       const float (&lnpost)[R], const float* diff = nullptr) {
    if (t1 - t0 == (long long)NT * R && t1 <= cur_end) {
      fold_full<R>(lnpred, lnpost, diff);
      if (t1 < cur_end) return;
      // the chain ends exactly at the tile boundary
      flush();
      const long long seg_begin = cur_end;
      do {
        ++cur_chain;
        cur_end = (cur_chain < p.nch) ? p.start[cur_chain + 1] : (long long)0x7fffffffffffffffLL;
      } while (cur_chain < p.nch && cur_end <= seg_begin);
      return;
    }
    long long seg_begin = t0;
    const long long r0 = t0 + (long long)tid * V;
    while (true) {
      const long long seg_end = (t1 < cur_end) ? t1 : cur_end;
#pragma unroll
      for (int q = 0; q < R; ++q) {
        const long long r = r0 + (long long)(q / V) * NT * V + (q % V);
        if (r >= seg_begin && r < seg_end) sample(lnpred[q], lnpost[q], diff ? diff + q : nullptr);
      }
      if (cur_chain < p.nch && cur_end <= t1) {
        flush();
        seg_begin = cur_end;
        do {
          ++cur_chain;
          cur_end = (cur_chain < p.nch) ? p.start[cur_chain + 1] : (long long)0x7fffffffffffffffLL;
        } while (cur_chain < p.nch && cur_end <= seg_begin);
        if (seg_begin >= t1 || cur_chain >= p.nch) break;
      } else {
        break;
      }
    }
  }

  // End of the flusher's range: flush the open chain and the per-call scalars.
  __device__ void finish(bool had_rows) {
    if (had_rows && cur_chain < p.nch) flush();
    constexpr int NW = NT / 32;
    double gs = gsum, gc = gcnt;
    float v[6] = {amax, amin, pmax, pmin, qmax, qmin};
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      gs += __shfl_xor_sync(0xffffffffu, gs, o);
      gc += __shfl_xor_sync(0xffffffffu, gc, o);
#pragma unroll
      for (int k = 0; k < 6; k += 2) {
        v[k] = fmaxf(v[k], __shfl_xor_sync(0xffffffffu, v[k], o));
        v[k + 1] = fminf(v[k + 1], __shfl_xor_sync(0xffffffffu, v[k + 1], o));
      }
    }
    // scratch holds 3*NW doubles; reuse it in three rounds of (sum,cnt | extrema)
    named_bar_sync(bar, NT);
    double out[G_N];
    for (int k = 0; k < G_N; ++k) {
      double mine = (k == 0) ? gs : (k == 1) ? gc : (double)v[k - 2];
      if ((tid & 31) == 0) scratch[tid >> 5] = mine;
      named_bar_sync(bar, NT);
      if (tid == 0) {
        double acc = scratch[0];
        for (int w = 1; w < NW; ++w) {
          double o = scratch[w];
          if (k < 2) acc += o;
          else if (k & 1) acc = fmin(acc, o);  // odd k = minima (ARGMIN=3, PROBMIN=5, PREDMIN=7)
          else acc = fmax(acc, o);
        }
        out[k] = acc;
      }
      named_bar_sync(bar, NT);
    }
    if (tid == 0) {
      for (int k = 0; k < G_N; ++k) p.gpart[(long long)flusher * G_N + k] = out[k];
    }
  }
};

template <typename T>
__device__ __forceinline__ float ldf(const T* p, long long i) {
  return (float)p[i];
}
#endif  // __CUDACC__

// ------------------------------------------------------------------------------------------
// host-side handles
// ------------------------------------------------------------------------------------------
struct Accum {
  int device = 0;
  long long nchains = 0;
  // ONE allocation [nchains x ChainState | G_N per-call scalars] = the packed block that crosses
  // NVLink (hb_accum_packed_ptr): state and globals point into it
  double* packed = nullptr;
  ChainState* state = nullptr;  // = packed
  double* globals = nullptr;    // = packed + 4 * nchains
  long long* start_dev = nullptr;  // device [nchains+1] scratch for the local offsets
  ChainRec* recs = nullptr;        // device scratch
  size_t recs_cap = 0;
  double* gpart = nullptr;  // device scratch
  size_t gpart_cap = 0;
  // reset runs on the legacy stream (hb_accum_create has no stream argument): the first use on another
  // stream waits on this event
  cudaEvent_t ev_reset = nullptr;
  bool reset_pending = false;
  // tensor-path range guard: hb_accumulate_flow calls of this accumulator re-run on the CUDA cores since the last
  // reset; shadow = copy of the packed block taken at the start of such a call (allocated on first use)
  long long guard_trips = 0;
  double* shadow = nullptr;
  // finalize outputs
  double* fin = nullptr;       // device: hb_result + 4*nchains doubles
  double* fin_host = nullptr;  // pinned mirror (the 128-byte result travels alone unless arrays are asked for)
};

// Staging for HOST inputs (pinned + device, double-buffered).  Borrowed per call from a per-device free
// list (stage_acquire / StageLease), so parked accumulators hold no large buffers.
struct Stage {
  int device = 0;
  void* pin[2] = {nullptr, nullptr};
  void* dev[2] = {nullptr, nullptr};
  size_t stage_cap = 0;
  size_t pin_cap = 0;
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_copy[2] = {nullptr, nullptr};
  cudaEvent_t ev_done[2] = {nullptr, nullptr};
};

}  // namespace hb
