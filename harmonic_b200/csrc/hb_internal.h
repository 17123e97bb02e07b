// Host-side internal declarations shared by the translation units of libharmonic_b200.
#pragma once
#include <string.h>

#include <thread>
#include <vector>

#include "hb_common.cuh"

namespace hb {

int require_device(int device);
int ensure_stage(Stage* sg, size_t bytes_per_slot);
int ensure_pinned(Stage* sg, size_t bytes_per_slot);
void timing_begin(cudaStream_t st);
void timing_end(cudaStream_t st);

int ensure_scratch(Accum* acc, long long nch, long long nflushers);
int order_after_reset(Accum* acc, cudaStream_t st);
int begin_call(Accum* acc, bool has_rows, cudaStream_t st);
int launch_merge(Accum* acc, const FoldParams& p, long long chain_lo, long long nflushers,
                 cudaStream_t st);
int prepare_fold(Accum* acc, const int64_t* start, long long chain_lo, long long chain_hi,
                 long long a, long long b, long long nflushers, long long rows_per_flusher,
                 FoldParams* p, cudaStream_t st);
int accumulate_precomputed_device(Accum* acc, const void* lnpred, int pd, const void* lnpost, int ld,
                                  const int64_t* start, long long chain_lo, long long chain_hi,
                                  long long a, long long b, cudaStream_t st);

// ---- host staging helpers --------------------------------------------------------------------
Stage* stage_acquire(int device);
void stage_release(Stage* s, cudaStream_t st);
// borrows a Stage for one host-input call; on every exit path (errors included) the copy stream and the
// compute stream are drained before the buffers return to the free list
struct StageLease {
  Stage* s;
  cudaStream_t st;
  StageLease(int device, cudaStream_t st_) : s(stage_acquire(device)), st(st_) {}
  ~StageLease() { stage_release(s, st); }
  StageLease(const StageLease&) = delete;
  StageLease& operator=(const StageLease&) = delete;
  Stage* get() const { return s; }
};

inline bool is_pinned_host(const void* p) {
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return attr.type == cudaMemoryTypeHost;
}

// Host worker threads for the pageable-input staging (std::thread, spawned per chunk: the library keeps
// no thread pool).  They inherit the caller's CPU affinity, so a rank bound to its GPU's NUMA node
// (hb_bind_host_to_device_node) converts node-locally.  HB_HOST_THREADS overrides the count.
int host_threads();

template <typename F>
inline void parallel_rows(long long r0, long long r1, int nthreads, long long min_rows, F work) {
  const long long n = r1 - r0;
  if (nthreads <= 1 || n < min_rows) {
    work(r0, r1);
    return;
  }
  std::vector<std::thread> th;
  const long long per = (n + nthreads - 1) / nthreads;
  for (int t = 0; t < nthreads; ++t) {
    const long long a = r0 + t * per, b = (a + per < r1) ? a + per : r1;
    if (a >= b) break;
    th.emplace_back(work, a, b);
  }
  for (auto& t : th) t.join();
}

// rows [r0, r1) of a (n, ld) host matrix -> contiguous float32 (n, ncol), split across host threads
template <typename T>
inline void convert_rows(const T* src, long long ld, int ncol, long long r0, long long r1, float* dst,
                         int nthreads) {
  parallel_rows(r0, r1, nthreads, 4096, [=](long long a, long long b) {
    for (long long r = a; r < b; ++r) {
      const T* s = src + r * ld;
      float* d = dst + (r - r0) * ncol;
      for (int c = 0; c < ncol; ++c) d[c] = (float)s[c];
    }
  });
}

inline void copy_bytes(const char* src, char* dst, size_t bytes, int nthreads) {
  parallel_rows(0, (long long)bytes, nthreads, 1 << 20, [=](long long a, long long b) {
    memcpy(dst + a, src + a, (size_t)(b - a));
  });
}

// ---- flows -------------------------------------------------------------------------------
struct RealNVPLayerDev {
  // SIMT layout (row-major, float32):
  const float* W0;  // [d][128]
  const float* b0;  // [128]
  const float* Wt;  // [128][u]
  const float* bt;  // [u]
  const float* Ws;  // [128][u] or nullptr
  const float* bs;  // [u] or nullptr
};

struct Flow {
  hb_flow_desc desc;
  int device = 0;
  int path = HB_PATH_AUTO;
  std::vector<float> host_weights;  // canonical flat buffer (header order)
  float* dev_weights = nullptr;     // same, on device
  float sum_log_amp = 0.f;          // sum(log(pre_amp)) (model.py:234-235)
  // RealNVP
  int d = 0, u = 0, nlayers = 0;
  std::vector<size_t> layer_off;  // offset of each layer's block in the flat buffer
  // RQSpline (re-packed, see hb_rqspline.cu)
  float* rq_pack = nullptr;
  std::vector<size_t> rq_off;
  size_t rq_pack_floats = 0;
  void* simt_params = nullptr;  // cached RealNVPSimtParams (hb_flow_simt.cu)
  float* rqf_pack = nullptr;    // RQSpline fast-path weights (hb_flow_simt.cu)
  int rqf_fe = 0, rqf_fo = 0;   // floats per even / odd flow layer in rqf_pack
  void* small_cache = nullptr;  // cached small-ndim register-path weights (hb_flow_simt.cu)
  bool force_generic = false;   // HB_PATH_SIMT_GENERIC: skip the small-ndim register path
  // tcgen05 image (hi/lo split, UMMA canonical layout, see hb_realnvp_tc.cu)
  void* tc_image = nullptr;
  size_t tc_image_bytes = 0;
  bool tc_ok = false;
  void* stc = nullptr;   // small-ndim RealNVP tcgen05 path: resident weight image (hb_realnvp_small_tc.cu); sets tc_ok
  void* rqtc = nullptr;  // RQSpline tcgen05 path: resident weight image (hb_rqspline_tc.cu); sets tc_ok
  // range guard of the tensor path (hb_realnvp_tc.cu): device counter of rows beyond the limits since the last
  // hb_flow_guard_take, and the limits on |standardised input| / |any conditioning input|
  unsigned long long* tc_guard = nullptr;
  float tc_guard_raw = 0.f, tc_guard_all = 0.f;
  bool force_simt = false;                   // set while a guarded call is being re-run on the CUDA cores
  long long guard_rows = 0, guard_calls = 0; // rows counted / calls re-run since creation (hb_flow_guard_stats)
};

// kernels: evaluate lnphi for rows [0, n) of x; if acc != nullptr also fold with lnpost into acc.
// lnpred_out may be nullptr.
int realnvp_simt_run(Flow* f, float T, const void* x, int x_dtype, long long ld, long long n,
                     const void* lnpost, int lp_dtype, float* lnpred_out, Accum* acc,
                     const int64_t* start, long long chain_lo, long long chain_hi, long long a,
                     long long b, cudaStream_t st);
int rqspline_run(Flow* f, float T, const void* x, int x_dtype, long long ld, long long n,
                 const void* lnpost, int lp_dtype, float* lnpred_out, Accum* acc,
                 const int64_t* start, long long chain_lo, long long chain_hi, long long a,
                 long long b, cudaStream_t st);
int realnvp_tc_run(Flow* f, float T, const void* x, int x_dtype, long long ld, long long n,
                   const void* lnpost, int lp_dtype, float* lnpred_out, Accum* acc,
                   const int64_t* start, long long chain_lo, long long chain_hi, long long a,
                   long long b, unsigned long long* guard, cudaStream_t st);
// sampling direction: out[r][0..ndim) = flow(eps[r]) for standard-normal draws eps (f32 / f64), generic SIMT kernels
int flow_sample_run(Flow* f, float T, const void* eps, int eps_dtype, long long ld, long long n, float* out,
                    long long ldo, cudaStream_t st);
int realnvp_tc_prepare(Flow* f);  // builds tc_image; sets tc_ok
int realnvp_small_tc_prepare(Flow* f);  // builds stc for ndim 2..8; sets tc_ok, allocates tc_guard
void realnvp_small_tc_release(Flow* f);
int realnvp_small_tc_run(Flow* f, float T, const void* x, int x_dtype, long long ld, long long n,
                         const void* lnpost, int lp_dtype, float* lnpred_out, Accum* acc,
                         const int64_t* start, long long chain_lo, long long chain_hi, long long a,
                         long long b, unsigned long long* guard, cudaStream_t st);
int rqspline_tc_prepare(Flow* f);  // builds rqtc for the shapes the RQSpline tensor path serves; sets tc_ok
void rqspline_tc_release(Flow* f);
int rqspline_tc_run(Flow* f, float T, const void* x, int x_dtype, long long ld, long long n,
                    const void* lnpost, int lp_dtype, float* lnpred_out, Accum* acc,
                    const int64_t* start, long long chain_lo, long long chain_hi, long long a,
                    long long b, cudaStream_t st);
int rqspline_prepare(Flow* f);
void realnvp_tc_release(Flow* f);
void realnvp_simt_release(Flow* f);
void realnvp_small_release(Flow* f);

}  // namespace hb
