/*
 * harmonic_b200.h -- C ABI of libharmonic_b200.so (sm_100a only, no CPU fallback).
 *
 * The reference (astro-informatics/harmonic) has NO FFI for this path: its boundary is the
 * duck-typed Python Model contract (harmonic/model_abstract.py:6-66) consumed by
 * Evidence.add_chains (harmonic/evidence.py:215-358).  Each entry point below cites the
 * reference code it replaces; harmonic_b200/{model,evidence}.py mirror the reference classes
 * on top of these calls through ctypes.
 *
 * Conventions
 *  - every function returns 0 on success, a negative hb_status otherwise; the message for the
 *    calling thread is available from hb_last_error().
 *  - buffers are caller-owned and borrowed for the duration of the call.  `mem` says where a
 *    buffer lives: HB_MEM_HOST (pageable or pinned host memory; the library stages it with
 *    chunked cudaMemcpyAsync overlapped with compute) or HB_MEM_DEVICE (zero-copy).
 *  - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream).  Calls
 *    taking HOST buffers synchronise before returning; calls on DEVICE buffers are
 *    asynchronous on `stream` unless stated otherwise.
 *  - samples are row-major (n, ndim) with leading dimension `ld` >= ndim (in elements), as in
 *    harmonic.chains.Chains.samples (harmonic/chains.py:28-33,65-69).
 */
#ifndef HARMONIC_B200_H
#define HARMONIC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

#define HB_VERSION 200 /* 0.2.0 */

typedef enum {
  HB_OK = 0,
  HB_ERR_INVALID = -1,     /* bad argument (the Python layer raises ValueError first) */
  HB_ERR_CUDA = -2,        /* CUDA runtime error; text in hb_last_error() */
  HB_ERR_UNSUPPORTED = -3, /* shape outside what the kernels support */
  HB_ERR_NO_DEVICE = -4    /* no sm_100 device: there is no CPU fallback */
} hb_status;

enum { HB_FLOW_REALNVP = 1, HB_FLOW_RQSPLINE = 2 };
enum { HB_F32 = 0, HB_F64 = 1 };
enum { HB_MEM_HOST = 0, HB_MEM_DEVICE = 1 };
/* harmonic.evidence.Shifting (harmonic/evidence.py:11-20) */
enum { HB_MEAN_SHIFT = 1, HB_MAX_SHIFT = 2, HB_MIN_SHIFT = 3, HB_ABS_MAX_SHIFT = 4 };
/* kernel selection for RealNVP: AUTO picks tcgen05 when the flow is supported there (ndim 64, hidden 128, <= 8
 * coupling layers, every conditioner weight and bias |w| < 16: the tensor path stores them * 2^11 as halves) */
enum { HB_PATH_AUTO = 0, HB_PATH_SIMT = 1, HB_PATH_TCGEN05 = 2, HB_PATH_SIMT_GENERIC = 3 /* no small-ndim register path */ };
/* flags for hb_accumulate_* */
enum { HB_ACC_NEW_CALL = 1 /* first launch of an add_chains call: reset the per-call scalars */ };
/* flags for hb_accum_merge_packed */
enum { HB_MERGE_OVERWRITE = 1 /* the merged blocks REPLACE the accumulator's state (no separate reset) */ };

#define HB_MAX_HIDDEN 8

/* Hyper-parameters of a fitted flow: the constructor arguments of RealNVPModel
 * (harmonic/model.py:286-337) / RQSplineModel (harmonic/model.py:348-404) that shape the
 * network, plus `standardize` (harmonic/model.py:118,174-183). */
typedef struct {
  int32_t kind;                  /* HB_FLOW_REALNVP | HB_FLOW_RQSPLINE */
  int32_t ndim;                  /* >= 2 for RealNVP (one masked + one transformed feature) */
  int32_t n_scaled;              /* RealNVP n_scaled_layers   (> 0) */
  int32_t n_unscaled;            /* RealNVP n_unscaled_layers (>= 0) */
  int32_t n_layers;              /* RQSpline n_layers */
  int32_t n_bins;                /* RQSpline n_bins (K) */
  int32_t n_hidden;              /* RQSpline len(hidden_size) (<= HB_MAX_HIDDEN) */
  int32_t hidden[HB_MAX_HIDDEN]; /* RQSpline hidden_size */
  float range_min, range_max;    /* RQSpline spline_range */
  int32_t standardize;           /* 0/1: apply (x - pre_offset) / pre_amp, subtract sum(log pre_amp) */
} hb_flow_desc;

/* Flat float32 weight buffer, in this order (all Dense kernels row-major (in, out) as flax stores
 * them; names as in model.state.params):
 *   pre_offset[ndim], pre_amp[ndim]                      (ignored unless standardize)
 *   RealNVP, for layer in scaled_layers_0..,unscaled_layers_0..:
 *       Dense_0.kernel[d*128], Dense_0.bias[128], Dense_1.kernel[128*u], Dense_1.bias[u],
 *       (scaled only) Dense_2.kernel[128*u], Dense_2.bias[u]          d = round(ndim/2), u = ndim-d
 *   RQSpline, for i in 0..n_layers-1:
 *       conditioner_i MLP Dense m = 0..n_hidden: kernel[in*out], bias[out]  (widths ndim, hidden...)
 *       conditioner_i output Dense: kernel[h_last * ndim*(3K+1)], bias[ndim*(3K+1)]
 *       scalar_i.scales[ndim], scalar_i.shifts[ndim]
 * hb_flow_weight_count() returns the expected length. */
typedef struct hb_flow hb_flow;
typedef struct hb_accum hb_accum;

/* Scalars produced by the finalise kernel: Evidence.process_run (harmonic/evidence.py:125-189). */
typedef struct {
  double ln_evidence_inv;
  double ln_evidence_inv_var;
  double ln_evidence_inv_var_var;
  double ln_kurtosis;
  double n_eff;
  double shift_value;  /* the shift in effect (harmonic/evidence.py:307-319) */
  double lnargmax, lnargmin;         /* of the last call, shifted (harmonic/evidence.py:347-348) */
  double lnprobmax, lnprobmin;       /* harmonic/evidence.py:349-350 */
  double lnpredictmax, lnpredictmin; /* harmonic/evidence.py:351-352 */
  double nsamples_total;
  double mean_nsamples_eff; /* mean over chains of nsamples_eff_per_chain (harmonic/evidence.py:372-378) */
  double guard_trips;       /* tensor-path range guard: hb_accumulate_flow calls re-run on the CUDA cores since the last reset */
  double reserved;
} hb_result;

int hb_version(void);
const char* hb_last_error(void);
/* number of visible sm_100 devices (0 => every compute call fails with HB_ERR_NO_DEVICE) */
int hb_device_count(void);

size_t hb_flow_weight_count(const hb_flow_desc* desc);

/* Replaces: reading model.state.params / pre_offset / pre_amp inside FlowModel.predict
 * (harmonic/model.py:222-235).  Copies and re-packs the weights on `device`.  RealNVP: 2 <= ndim <= 163 (the generic
 * kernel keeps ndim + 128 + 2u activation columns in shared memory; wider flows return HB_ERR_UNSUPPORTED here). */
int hb_flow_create(const hb_flow_desc* desc, const float* weights, size_t n_weights, int device,
                   hb_flow** out);
int hb_flow_destroy(hb_flow* flow);
/* HB_PATH_*.  HB_PATH_TCGEN05 exists for RealNVP at ndim 64 (<= 8 coupling layers) and at ndim 2..8, and for RQSpline
 * at ndim 2 (two hidden layers, first one 32 or 64 wide, 8 bins); HB_PATH_AUTO takes it where it exists (RealNVP at
 * ndim 2..8: from 2^20 rows per call on, below that the register kernel).  Returns
 * HB_ERR_UNSUPPORTED if the path cannot run this flow. */
int hb_flow_set_path(hb_flow* flow, int path);
/* 1 if the tcgen05 path will be used for this flow */
int hb_flow_uses_tcgen05(const hb_flow* flow);
/* Range guard of the tcgen05 path under HB_PATH_AUTO: sample rows found outside the range its FP16-split arithmetic
 * serves (|standardised input| > 32, or a hidden pre-activation that would leave the half range) since the flow was
 * created, and the hb_flow_predict / hb_accumulate_flow calls that were therefore evaluated again on the CUDA-core
 * path (float32 FMA arithmetic, as harmonic/flows.py:159-180 runs it on XLA's float32).  With HB_PATH_TCGEN05 forced
 * nothing is re-run. */
int hb_flow_guard_stats(const hb_flow* flow, int64_t* rows, int64_t* calls);

/* Replaces FlowModel.predict (harmonic/model.py:196-237): out[i] = ln phi(x_i) at `temperature`
 * (0 < T <= 1 is validated by the caller, as the reference does on every call). */
int hb_flow_predict(hb_flow* flow, float temperature, const void* x, int x_dtype, int x_mem,
                    int64_t n, int64_t ld, float* out, int out_mem, void* stream);

/* Replaces FlowModel.sample (harmonic/model.py:239-275; flows.py:118-138, 290-313): out[i][0..ndim) = the flow
 * in the sampling direction applied to the base draw built from the standard-normal row eps[i] (RealNVP:
 * T * eps, RQSpline: sqrt(T) * eps), de-standardised (x * pre_amp + pre_offset) when the flow standardises.
 * The random draws are the caller's (the reference takes a JAX PRNG key), so the call is deterministic.
 * eps: (n, ld_eps >= ndim) HB_F32 / HB_F64; out: (n, ndim) float32, contiguous; both in `mem`. */
int hb_flow_sample(hb_flow* flow, float temperature, const void* eps, int eps_dtype, int64_t n,
                   int64_t ld_eps, float* out, int mem, void* stream);

/* Streaming per-chain state of Evidence (harmonic/evidence.py:65-67,83-85): for every chain the
 * max-rescaled pair (m, s) with ln sum_i exp(lnarg_i) = m + ln s, the sample count n and the NaN
 * count; plus the per-call scalars. */
/* hb_accum_create resets on the legacy default stream; later calls on any stream order themselves behind
 * that reset.  hb_accum_destroy parks the accumulator (O(nchains) bytes) for reuse by the next create with
 * the same (device, nchains); hb_pool_trim frees everything parked, staging buffers included. */
int hb_accum_create(int64_t nchains, int device, hb_accum** out);
int hb_accum_destroy(hb_accum* acc);
int hb_accum_reset(hb_accum* acc, void* stream);
int hb_pool_trim(void);

/* Replaces the batched branch of Evidence.add_chains (harmonic/evidence.py:260-283,321-352) fused
 * with FlowModel.predict: for chains [chain_lo, chain_hi) evaluate ln phi, form
 * lnarg = ln phi - ln_posterior (+-inf -> NaN), and fold into the per-chain state.
 * `start_indices` (HOST, nchains+1 entries, Chains.start_indices, harmonic/chains.py:29) is the
 * global layout; `x` / `ln_posterior` point at row start_indices[chain_lo] (a rank passes just its
 * shard).  `lnpred_out` (optional, same mem space as x, float32) receives ln phi per row. */
int hb_accumulate_flow(hb_accum* acc, hb_flow* flow, float temperature, const void* x, int x_dtype,
                       int64_t ld, const void* ln_posterior, int lp_dtype, int mem,
                       const int64_t* start_indices, int64_t chain_lo, int64_t chain_hi,
                       float* lnpred_out, int flags, void* stream);

/* Same with a precomputed ln phi (accumulation-only path; also what a user-supplied Model without
 * a flow would feed, harmonic/evidence.py:285-303). */
int hb_accumulate_precomputed(hb_accum* acc, const void* lnpred, int lnpred_dtype,
                              const void* ln_posterior, int lp_dtype, int mem,
                              const int64_t* start_indices, int64_t chain_lo, int64_t chain_hi,
                              int flags, void* stream);

/* partials: nchains x 4 doubles (m, s, n, n_nan); globals: 8 doubles
 * (sum lnarg, #finite, lnarg max, lnarg min, lnprob max, lnprob min, lnpred max, lnpred min).
 * Both run on `stream` and synchronise it.  Merge folds another rank's / an earlier checkpoint's partials
 * in; this is what crosses NVLink (one all-gather of O(nchains) doubles) and what gets pickled. */
int hb_accum_export(hb_accum* acc, double* partials, double* globals, void* stream);
int hb_accum_merge(hb_accum* acc, const double* partials, const double* globals, void* stream);

/* Packed form for the multi-GPU exchange: [nchains x 4 partials | 8 globals] = 4*nchains + 8 doubles.
 * hb_accum_pack writes this rank's block to `dst` (HOST or DEVICE; device copies are asynchronous on
 * `stream`).  The accumulator keeps its state in exactly this layout in ONE device allocation, so
 * hb_accum_packed_ptr hands out the send buffer of the path's one all-gather without a copy (valid until
 * the accumulator is destroyed; contents are stream-ordered behind the accumulate calls).
 * hb_accum_merge_packed folds `nranks` such blocks (block r at packed + r*stride doubles, rank order =
 * merge order) in one kernel launch; HB_MERGE_OVERWRITE makes them replace the previous state. */
int hb_accum_pack(hb_accum* acc, double* dst, int mem, void* stream);
int hb_accum_packed_ptr(hb_accum* acc, double** device_ptr, int64_t* n_doubles);
int hb_accum_merge_packed(hb_accum* acc, int nranks, const double* packed, int64_t stride, int mem,
                          int flags, void* stream);

/* Replaces Evidence.process_run (harmonic/evidence.py:125-189) and the shift selection
 * (harmonic/evidence.py:307-319): a single-CTA float64 kernel.  `shift_value_in` is the frozen
 * shift of an earlier call, or NaN to derive it from this call's scalars with `shift_mode`.
 * Optional outputs (HOST, nchains doubles each, may be NULL): ln_evidence_inv_per_chain,
 * running_sum (= exp(m + ln s + shift)), nsamples_per_chain, nsamples_eff_per_chain.
 * Runs on `stream` behind the accumulate / merge launches issued there (no device-wide synchronisation),
 * copies the 128-byte result (plus the arrays when asked for) and synchronises `stream`.
 * hb_finalize_arrays fetches the per-chain arrays of the most recent hb_finalize later. */
int hb_finalize(hb_accum* acc, int shift_mode, double shift_value_in, hb_result* out,
                double* ln_inv_per_chain, double* running_sum, double* nsamples_per_chain,
                double* nsamples_eff_per_chain, void* stream);
int hb_finalize_arrays(hb_accum* acc, double* ln_inv_per_chain, double* running_sum,
                       double* nsamples_per_chain, double* nsamples_eff_per_chain, void* stream);

/* Evidence.process_run (harmonic/evidence.py:125-189) on caller-supplied real-space running totals
 * (HOST arrays of nchains doubles): what the reference computes when process_run() is invoked on
 * hand-set `running_sum` / `nsamples_per_chain` / `shift_value` (tests/test_evidence.py:95-144);
 * sums may be negative there, so this variant works in real space.  Synchronises. */
int hb_finalize_realspace(int device, int64_t nchains, const double* running_sum,
                          const double* nsamples_per_chain, double shift_value, hb_result* out,
                          double* ln_inv_per_chain);

/* pinned host buffers for callers that want full-rate host->device staging */
int hb_host_alloc(size_t bytes, void** out);
int hb_host_free(void* p);
/* plain device buffers + copies so that Python callers need no other CUDA binding */
int hb_device_alloc(int device, size_t bytes, void** out);
int hb_device_free(void* p);
int hb_memcpy_h2d(void* dst, const void* src, size_t bytes, void* stream);
int hb_memcpy_d2h(void* dst, const void* src, size_t bytes, void* stream);
int hb_stream_sync(void* stream);

/* timing aid: number of kernels launched by this library in the calling process so far */
int64_t hb_launch_count(void);
/* duration (ms, CUDA events on the call's stream) of the flow/accumulate kernel of the most
 * recent hb_accumulate_* / hb_flow_predict call on `device`; for bench.py's roofline block.
 * Enabled by hb_set_kernel_timing(1); the state is per device and mutex-guarded. */
int hb_set_kernel_timing(int on);
double hb_last_kernel_ms(int device);

/* Host-side helpers for the staged (HOST buffer) path.  Pageable inputs are converted to float32 by up to 16
 * std::threads per chunk (HB_HOST_THREADS overrides; they inherit the caller's CPU affinity).
 * hb_bind_host_to_device_node pins the calling thread to the CPUs of the GPU's NUMA node (returns the CPU
 * count, 0 if the topology is not exposed): call it once per rank before allocating host buffers.
 * hb_host_copy_gbs: STREAM-style copy rate of the host with `threads` workers (GB/s, read + write). */
int hb_bind_host_to_device_node(int device);
double hb_host_copy_gbs(size_t bytes, int threads);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* HARMONIC_B200_H */
