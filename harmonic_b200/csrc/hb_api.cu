// C ABI glue: error state, device checks, flow handles, predict / accumulate dispatch with
// host-buffer staging.  See include/harmonic_b200.h for the contract.
#include <math.h>
#include <string.h>

#include <sched.h>
#include <stdio.h>
#include <stdlib.h>

#include <atomic>
#include <chrono>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "hb_internal.h"

namespace hb {

static thread_local std::string g_err;
static std::atomic<long long> g_launches{0};

// kernel timing (bench.py's roofline block): one event pair per device, guarded by a mutex -- several host
// threads may drive handles on the same device
constexpr int MAX_DEV = 64;
struct DevTiming {
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool pending = false;
};
static std::mutex g_timing_mu;
static bool g_timing = false;
static DevTiming g_dev_timing[MAX_DEV];

void set_error(const std::string& msg) { g_err = msg; }
void count_launch(int n) { g_launches += n; }

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  g_err = std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what + " (" + file + ":" +
          std::to_string(line) + ")";
  return HB_ERR_CUDA;
}

static DevTiming* timing_slot() {  // caller holds g_timing_mu
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEV) return nullptr;
  DevTiming* t = &g_dev_timing[dev];
  if (!t->ev0) {
    if (cudaEventCreate(&t->ev0) != cudaSuccess || cudaEventCreate(&t->ev1) != cudaSuccess) return nullptr;
  }
  return t;
}
void timing_begin(cudaStream_t st) {
  std::lock_guard<std::mutex> lk(g_timing_mu);
  if (!g_timing) return;
  if (DevTiming* t = timing_slot()) cudaEventRecord(t->ev0, st);
}
void timing_end(cudaStream_t st) {
  std::lock_guard<std::mutex> lk(g_timing_mu);
  if (!g_timing) return;
  if (DevTiming* t = timing_slot()) {
    cudaEventRecord(t->ev1, st);
    t->pending = true;
  }
}

int host_threads() {
  static const int forced = [] {
    const char* e = getenv("HB_HOST_THREADS");
    return e ? atoi(e) : 0;
  }();
  if (forced > 0) return forced;
  int n = 0;
  cpu_set_t set;
  if (sched_getaffinity(0, sizeof(set), &set) == 0) n = CPU_COUNT(&set);
  if (n <= 0) n = (int)std::thread::hardware_concurrency();
  return n > 16 ? 16 : (n < 1 ? 1 : n);
}

int require_device(int device) {
  // cudaGetDeviceProperties costs milliseconds per call; remember devices already validated
  static std::atomic<unsigned long long> ok_mask{0};
  if (device >= 0 && device < 64 && ((ok_mask.load() >> device) & 1ull)) return HB_OK;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    cudaGetLastError();
    set_error("no CUDA device visible: libharmonic_b200 has no CPU fallback");
    return HB_ERR_NO_DEVICE;
  }
  if (device < 0 || device >= n) {
    set_error("device index out of range");
    return HB_ERR_INVALID;
  }
  cudaDeviceProp prop;
  HB_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10 || prop.minor != 0) {  // the binary holds sm_100a code only
    set_error(std::string("device is sm_") + std::to_string(prop.major * 10 + prop.minor) +
              "; libharmonic_b200 is built for sm_100a only");
    return HB_ERR_NO_DEVICE;
  }
  if (device < 64) ok_mask.fetch_or(1ull << device);
  return HB_OK;
}

int ensure_pinned(Stage* acc, size_t bytes_per_slot) {
  if (bytes_per_slot <= acc->pin_cap) return HB_OK;
  for (int i = 0; i < 2; ++i) {
    if (acc->pin[i]) cudaFreeHost(acc->pin[i]);
    acc->pin[i] = nullptr;
  }
  acc->pin_cap = 0;
  for (int i = 0; i < 2; ++i) HB_CUDA(cudaHostAlloc(&acc->pin[i], bytes_per_slot, cudaHostAllocDefault));
  acc->pin_cap = bytes_per_slot;
  return HB_OK;
}

int ensure_stage(Stage* acc, size_t bytes_per_slot) {
  if (bytes_per_slot <= acc->stage_cap) return HB_OK;
  for (int i = 0; i < 2; ++i) {
    if (acc->dev[i]) cudaFree(acc->dev[i]);
    acc->dev[i] = nullptr;
  }
  acc->stage_cap = 0;
  for (int i = 0; i < 2; ++i) HB_CUDA(cudaMalloc(&acc->dev[i], bytes_per_slot));
  acc->stage_cap = bytes_per_slot;
  return HB_OK;
}

static int round_half_even_half(int ndim) { return (int)nearbyint(ndim * 0.5); }

// Small-ndim RealNVP under HB_PATH_AUTO: the tensor path's range guard ends every call with a device-to-host read of its
// counter (~30 us); below this many rows that costs more than realnvp_small_tc_kernel saves over the register kernel
// (cfg1, 5e5 rows: kernel 0.16 -> 0.10 ms, whole step 0.31 -> 0.34 ms).  HB_STC_MIN_ROWS overrides (tests: 0).
static long long stc_min_rows() {
  static const long long v = [] {
    const char* e = getenv("HB_STC_MIN_ROWS");
    return e ? atoll(e) : (long long)1 << 20;
  }();
  return v;
}

static bool use_tc(const Flow* f, long long n) {
  if (f->path == HB_PATH_TCGEN05) return true;
  if (!(f->path == HB_PATH_AUTO && f->tc_ok && !f->force_simt)) return false;
  return !(f->stc && n < stc_min_rows());
}

static int flow_run(Flow* f, float T, const void* x, int x_dtype, long long ld, long long n,
                    const void* lnpost, int lp_dtype, float* lnpred_out, Accum* acc,
                    const int64_t* start, long long chain_lo, long long chain_hi, long long a,
                    long long b, cudaStream_t st) {
  if (f->desc.kind == HB_FLOW_RQSPLINE)
    return rqspline_run(f, T, x, x_dtype, ld, n, lnpost, lp_dtype, lnpred_out, acc, start, chain_lo,
                        chain_hi, a, b, st);
  const bool tc = use_tc(f, n);
  if (tc && f->stc)
    return realnvp_small_tc_run(f, T, x, x_dtype, ld, n, lnpost, lp_dtype, lnpred_out, acc, start,
                                chain_lo, chain_hi, a, b, f->tc_guard, st);
  if (tc)
    return realnvp_tc_run(f, T, x, x_dtype, ld, n, lnpost, lp_dtype, lnpred_out, acc, start,
                          chain_lo, chain_hi, a, b, f->tc_guard, st);
  return realnvp_simt_run(f, T, x, x_dtype, ld, n, lnpost, lp_dtype, lnpred_out, acc, start,
                          chain_lo, chain_hi, a, b, st);
}

// Range guard of the tensor path (hb_realnvp_tc.cu): under HB_PATH_AUTO a call whose rows left the range the
// FP16-split arithmetic serves is evaluated again on the CUDA cores (float32 FMA arithmetic, any magnitude).
static bool guarded(const Flow* f, long long n) {
  return f->desc.kind == HB_FLOW_REALNVP && f->path == HB_PATH_AUTO && f->tc_guard && use_tc(f, n);
}
// rows counted since the last take; stream-ordered read + reset, then waits for the stream
static int guard_take(Flow* f, cudaStream_t st, long long* count) {
  unsigned long long host = 0;
  HB_CUDA(cudaMemcpyAsync(&host, f->tc_guard, sizeof(host), cudaMemcpyDeviceToHost, st));
  HB_CUDA(cudaStreamSynchronize(st));
  if (host) HB_CUDA(cudaMemsetAsync(f->tc_guard, 0, sizeof(host), st));
  *count = (long long)host;
  return HB_OK;
}
struct ForceSimt {
  Flow* f;
  explicit ForceSimt(Flow* f_) : f(f_) { f->force_simt = true; }
  ~ForceSimt() { f->force_simt = false; }
};

int accumulate_flow_once(Accum* acc, Flow* f, float temperature, const void* x, int x_dtype, long long ld,
                         const void* ln_posterior, int lp_dtype, int mem, const int64_t* start_indices,
                         long long chain_lo, long long chain_hi, long long nrows, float* lnpred_out,
                         cudaStream_t st);

static int accumulate_flow_pageable(Accum* acc, Flow* f, float temperature, const void* x, int x_dtype,
                                    long long ld, const void* lnpost, int lp_dtype,
                                    const int64_t* start_indices, long long chain_lo, long long chain_hi,
                                    long long nrows, float* lnpred_out, cudaStream_t st) {
  const int D = f->desc.ndim;
  const int nthreads = host_threads();
  long long chunk = (long long)((64ull << 20) / ((size_t)D * 4 + 4));
  chunk = (chunk / 4096) * 4096;
  if (chunk < 4096) chunk = 4096;
  if (chunk > nrows) chunk = ((nrows + 255) / 256) * 256;
  const size_t off_l = (((size_t)chunk * D * 4) + 255) & ~(size_t)255;
  const size_t off_o = (off_l + (size_t)chunk * 4 + 255) & ~(size_t)255;
  const size_t slot = off_o + (lnpred_out ? (size_t)chunk * 4 : 0) + 256;
  StageLease lease(acc->device, st);
  Stage* sg = lease.get();
  if (!sg) return HB_ERR_CUDA;
  int rc = ensure_stage(sg, slot);
  if (rc) return rc;
  rc = ensure_pinned(sg, off_o + 256);
  if (rc) return rc;
  int i = 0;
  for (long long a = 0; a < nrows; a += chunk, ++i) {
    const long long b = (a + chunk < nrows) ? a + chunk : nrows;
    const int sl = i & 1;
    char* hp = static_cast<char*>(sg->pin[sl]);
    char* dxp = static_cast<char*>(sg->dev[sl]);
    // the pinned slot is free once the H2D copy of chunk i-2 has been consumed
    HB_CUDA(cudaEventSynchronize(sg->ev_copy[sl]));
    if (x_dtype == HB_F64) convert_rows((const double*)x, ld, D, a, b, reinterpret_cast<float*>(hp), nthreads);
    else convert_rows((const float*)x, ld, D, a, b, reinterpret_cast<float*>(hp), nthreads);
    if (lp_dtype == HB_F64) convert_rows((const double*)lnpost, 1, 1, a, b, reinterpret_cast<float*>(hp + off_l), nthreads);
    else convert_rows((const float*)lnpost, 1, 1, a, b, reinterpret_cast<float*>(hp + off_l), nthreads);
    HB_CUDA(cudaStreamWaitEvent(sg->copy_stream, sg->ev_done[sl], 0));  // device slot free
    HB_CUDA(cudaMemcpyAsync(dxp, hp, off_l + (size_t)(b - a) * 4, cudaMemcpyHostToDevice, sg->copy_stream));
    HB_CUDA(cudaEventRecord(sg->ev_copy[sl], sg->copy_stream));
    HB_CUDA(cudaStreamWaitEvent(st, sg->ev_copy[sl], 0));
    float* dop = lnpred_out ? reinterpret_cast<float*>(dxp + off_o) : nullptr;
    rc = flow_run(f, temperature, dxp, HB_F32, D, b - a, dxp + off_l, HB_F32, dop, acc, start_indices,
                  chain_lo, chain_hi, a, b, st);
    if (rc) return rc;
    if (dop)
      HB_CUDA(cudaMemcpyAsync(lnpred_out + a, dop, (size_t)(b - a) * 4, cudaMemcpyDeviceToHost, st));
    HB_CUDA(cudaEventRecord(sg->ev_done[sl], st));
  }
  HB_CUDA(cudaStreamSynchronize(st));
  return HB_OK;
}

}  // namespace hb

using namespace hb;

extern "C" {

int hb_version(void) { return HB_VERSION; }
const char* hb_last_error(void) { return g_err.c_str(); }

int hb_device_count(void) {
  static int cached = -1;
  if (cached >= 0) return cached;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  int ok = 0;
  for (int i = 0; i < n; ++i) {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, i) == cudaSuccess && prop.major == 10 && prop.minor == 0) ++ok;
  }
  if (ok > 0) cached = ok;
  return ok;
}

int64_t hb_launch_count(void) { return g_launches.load(); }

int hb_set_kernel_timing(int on) {
  std::lock_guard<std::mutex> lk(g_timing_mu);
  g_timing = on != 0;
  return HB_OK;
}
double hb_last_kernel_ms(int device) {
  cudaEvent_t e0, e1;
  {
    std::lock_guard<std::mutex> lk(g_timing_mu);
    if (device < 0 || device >= MAX_DEV) return -1.0;
    const DevTiming& t = g_dev_timing[device];
    if (!t.pending || !t.ev0) return -1.0;
    e0 = t.ev0;
    e1 = t.ev1;
  }
  float ms = 0.f;
  if (cudaEventSynchronize(e1) != cudaSuccess) return -1.0;
  if (cudaEventElapsedTime(&ms, e0, e1) != cudaSuccess) return -1.0;
  return (double)ms;
}

// Pin the calling thread (and the staging threads it spawns later) to the CPUs of the NUMA node the GPU
// hangs off, read from sysfs; pinned staging buffers allocated afterwards land on that node (first touch).
// Returns the number of CPUs bound, 0 when the topology is not exposed (nothing changed).
int hb_bind_host_to_device_node(int device) {
  char bus[32];
  if (cudaDeviceGetPCIBusId(bus, sizeof(bus), device) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  for (char* c = bus; *c; ++c)
    if (*c >= 'A' && *c <= 'Z') *c = (char)(*c - 'A' + 'a');
  char path[128];
  snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/numa_node", bus);
  FILE* fh = fopen(path, "r");
  if (!fh) return 0;
  int node = -1;
  const int got = fscanf(fh, "%d", &node);
  fclose(fh);
  if (got != 1 || node < 0) return 0;
  snprintf(path, sizeof(path), "/sys/devices/system/node/node%d/cpulist", node);
  fh = fopen(path, "r");
  if (!fh) return 0;
  char list[4096];
  const bool ok = fgets(list, sizeof(list), fh) != nullptr;
  fclose(fh);
  if (!ok) return 0;
  cpu_set_t set;
  CPU_ZERO(&set);
  int ncpu = 0;
  for (char* p = list; *p;) {  // "0-15,32-47"
    char* end;
    const long a = strtol(p, &end, 10);
    if (end == p) break;
    long b = a;
    p = end;
    if (*p == '-') {
      b = strtol(p + 1, &end, 10);
      p = end;
    }
    for (long c = a; c <= b && c < CPU_SETSIZE; ++c) {
      CPU_SET((int)c, &set);
      ++ncpu;
    }
    if (*p == ',') ++p;
    else break;
  }
  if (ncpu == 0) return 0;
  cpu_set_t cur, both;
  if (sched_getaffinity(0, sizeof(cur), &cur) == 0) {  // never widen what the launcher allowed
    CPU_AND(&both, &cur, &set);
    if (CPU_COUNT(&both) == 0) return 0;
    set = both;
    ncpu = CPU_COUNT(&set);
  }
  if (sched_setaffinity(0, sizeof(set), &set) != 0) return 0;
  return ncpu;
}

// STREAM-style host copy rate (GB/s, read + write bytes) with `threads` workers over `bytes` per pass: the
// ceiling of the host-staged end-to-end path (bench.py reports it next to e2e).
double hb_host_copy_gbs(size_t bytes, int threads) {
  if (bytes < (1u << 20) || threads < 1) return 0.0;
  std::vector<char> src(bytes, 1), dst(bytes, 0);
  copy_bytes(src.data(), dst.data(), bytes, threads);  // warm / fault in
  double best = 0.0;
  for (int rep = 0; rep < 3; ++rep) {
    const auto t0 = std::chrono::steady_clock::now();
    copy_bytes(src.data(), dst.data(), bytes, threads);
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (dt > 0) best = std::max(best, 2.0 * (double)bytes / dt / 1e9);
  }
  return best;
}

size_t hb_flow_weight_count(const hb_flow_desc* ds) {
  if (!ds || ds->ndim < 1) return 0;
  const size_t D = ds->ndim;
  size_t n = 2 * D;
  if (ds->kind == HB_FLOW_REALNVP) {
    const size_t d = round_half_even_half(ds->ndim), u = D - d;
    n += (size_t)ds->n_scaled * (d * 128 + 128 + 2 * (128 * u + u));
    n += (size_t)ds->n_unscaled * (d * 128 + 128 + (128 * u + u));
  } else if (ds->kind == HB_FLOW_RQSPLINE) {
    if (ds->n_hidden < 0 || ds->n_hidden > HB_MAX_HIDDEN) return 0;
    const size_t npar = 3 * (size_t)ds->n_bins + 1;
    size_t per = D * D + D;
    size_t prev = D;
    for (int m = 0; m < ds->n_hidden; ++m) {
      per += prev * ds->hidden[m] + ds->hidden[m];
      prev = ds->hidden[m];
    }
    per += prev * D * npar + D * npar + 2 * D;
    n += (size_t)ds->n_layers * per;
  } else {
    return 0;
  }
  return n;
}

int hb_flow_create(const hb_flow_desc* ds, const float* weights, size_t n_weights, int device,
                   hb_flow** out) {
  HB_REQUIRE(ds && weights && out, "hb_flow_create: NULL argument");
  HB_REQUIRE(ds->ndim >= 1, "Dimension must be greater than 0.");
  HB_REQUIRE(ds->kind == HB_FLOW_REALNVP || ds->kind == HB_FLOW_RQSPLINE, "unknown flow kind");
  if (ds->kind == HB_FLOW_REALNVP) {
    HB_REQUIRE(ds->n_scaled > 0, "Number of scaled layers must be greater than 0.");
    HB_REQUIRE(ds->n_unscaled >= 0, "n_unscaled_layers must be >= 0");
    if (ds->ndim < 2 || round_half_even_half(ds->ndim) < 1) {
      set_error("RealNVP needs ndim >= 2 (one conditioning and one transformed feature)");
      return HB_ERR_UNSUPPORTED;
    }
  } else {
    HB_REQUIRE(ds->n_layers >= 1 && ds->n_bins >= 1, "RQSpline needs n_layers >= 1 and n_bins >= 1");
    HB_REQUIRE(ds->n_hidden >= 0 && ds->n_hidden <= HB_MAX_HIDDEN, "too many hidden layers");
    HB_REQUIRE(ds->range_max > ds->range_min, "spline_range must be increasing");
    HB_REQUIRE((double)ds->n_bins * 1e-4 <= (double)ds->range_max - ds->range_min,
               "spline range too small for n_bins * min_bin_size");
  }
  const size_t expect = hb_flow_weight_count(ds);
  HB_REQUIRE(expect == n_weights, "weight buffer length does not match the flow description");
  int rc = require_device(device);
  if (rc) return rc;
  HB_CUDA(cudaSetDevice(device));
  Flow* f = new Flow();
  f->desc = *ds;
  f->device = device;
  f->host_weights.assign(weights, weights + n_weights);
  const int D = ds->ndim;
  double sla = 0.0;
  if (ds->standardize)
    for (int i = 0; i < D; ++i) sla += log((double)weights[D + i]);
  f->sum_log_amp = (float)sla;
  cudaError_t e = cudaMalloc(&f->dev_weights, n_weights * sizeof(float));
  if (e == cudaSuccess)
    e = cudaMemcpy(f->dev_weights, weights, n_weights * sizeof(float), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    delete f;
    return cuda_fail(e, "weight upload", __FILE__, __LINE__);
  }
  if (ds->kind == HB_FLOW_REALNVP) {
    f->d = round_half_even_half(D);
    f->u = D - f->d;
    f->nlayers = ds->n_scaled + ds->n_unscaled;
    size_t off = 2 * (size_t)D;
    for (int l = 0; l < f->nlayers; ++l) {
      f->layer_off.push_back(off);
      const bool scaled = l < ds->n_scaled;
      off += (size_t)f->d * 128 + 128 + (scaled ? 2 : 1) * ((size_t)128 * f->u + f->u);
    }
    // the generic kernel keeps (ndim + 128 + 2u) activation columns of 128 threads in shared memory: say at
    // creation, not at the first predict, that a flow is too wide for it (ndim <= 163)
    if (((size_t)D + 128 + 2 * (size_t)f->u) * 128 * sizeof(float) > 227 * 1024) {
      hb_flow_destroy(reinterpret_cast<hb_flow*>(f));
      set_error("RealNVP: ndim too large for the shared-memory activation layout (supported: ndim <= 163)");
      return HB_ERR_UNSUPPORTED;
    }
    rc = realnvp_tc_prepare(f);  // sets tc_ok when the shape is supported (ndim 64)
    if (!rc) rc = realnvp_small_tc_prepare(f);  // ndim 2..8
  } else {
    rc = rqspline_prepare(f);
  }
  if (rc) {
    hb_flow_destroy(reinterpret_cast<hb_flow*>(f));
    return rc;
  }
  *out = reinterpret_cast<hb_flow*>(f);
  return HB_OK;
}

int hb_flow_destroy(hb_flow* h) {
  if (!h) return HB_OK;
  Flow* f = reinterpret_cast<Flow*>(h);
  cudaSetDevice(f->device);
  cudaDeviceSynchronize();
  if (f->dev_weights) cudaFree(f->dev_weights);
  if (f->rq_pack) cudaFree(f->rq_pack);
  if (f->rqf_pack) cudaFree(f->rqf_pack);
  realnvp_tc_release(f);
  realnvp_small_tc_release(f);
  rqspline_tc_release(f);
  realnvp_simt_release(f);
  realnvp_small_release(f);
  delete f;
  return HB_OK;
}

int hb_flow_set_path(hb_flow* h, int path) {
  HB_REQUIRE(h != nullptr, "hb_flow_set_path: NULL handle");
  Flow* f = reinterpret_cast<Flow*>(h);
  HB_REQUIRE(path == HB_PATH_AUTO || path == HB_PATH_SIMT || path == HB_PATH_TCGEN05 ||
                 path == HB_PATH_SIMT_GENERIC,
             "unknown path");
  f->force_generic = path == HB_PATH_SIMT_GENERIC;
  if (path == HB_PATH_TCGEN05 && !f->tc_ok) {
    set_error("tcgen05 path does not support this flow shape");
    return HB_ERR_UNSUPPORTED;
  }
  f->path = path;
  return HB_OK;
}

int hb_flow_uses_tcgen05(const hb_flow* h) {
  if (!h) return 0;
  const Flow* f = reinterpret_cast<const Flow*>(h);
  return (f->path == HB_PATH_TCGEN05) || (f->path == HB_PATH_AUTO && f->tc_ok);
}

static int flow_predict_once(Flow* f, float temperature, const void* x, int x_dtype, int x_mem, int64_t n,
                             int64_t ld, float* out, int out_mem, cudaStream_t st);

int hb_flow_predict(hb_flow* h, float temperature, const void* x, int x_dtype, int x_mem, int64_t n,
                    int64_t ld, float* out, int out_mem, void* stream) {
  HB_REQUIRE(h != nullptr, "hb_flow_predict: NULL handle");
  Flow* f = reinterpret_cast<Flow*>(h);
  HB_REQUIRE(temperature <= 1.f, "Scaling must not be greater than 1.");
  HB_REQUIRE(temperature > 0.f, "Scaling must be positive.");
  HB_REQUIRE(n >= 0 && ld >= f->desc.ndim, "bad shape");
  HB_REQUIRE(x_dtype == HB_F32 || x_dtype == HB_F64, "dtype must be HB_F32 or HB_F64");
  if (n == 0) return HB_OK;
  HB_REQUIRE(x != nullptr && out != nullptr, "NULL buffer");
  HB_CUDA(cudaSetDevice(f->device));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool guard = guarded(f, n);
  int rc = flow_predict_once(f, temperature, x, x_dtype, x_mem, n, ld, out, out_mem, st);
  if (rc || !guard) return rc;
  long long cnt = 0;
  rc = guard_take(f, st, &cnt);
  if (rc || cnt == 0) return rc;
  f->guard_rows += cnt;
  ++f->guard_calls;
  ForceSimt scope(f);
  return flow_predict_once(f, temperature, x, x_dtype, x_mem, n, ld, out, out_mem, st);
}

int hb_flow_guard_stats(const hb_flow* h, int64_t* rows, int64_t* calls) {
  HB_REQUIRE(h != nullptr, "hb_flow_guard_stats: NULL handle");
  const Flow* f = reinterpret_cast<const Flow*>(h);
  if (rows) *rows = f->guard_rows;
  if (calls) *calls = f->guard_calls;
  return HB_OK;
}

static int flow_predict_once(Flow* f, float temperature, const void* x, int x_dtype, int x_mem, int64_t n,
                             int64_t ld, float* out, int out_mem, cudaStream_t st) {
  if (x_mem == HB_MEM_DEVICE && out_mem == HB_MEM_DEVICE)
    return flow_run(f, temperature, x, x_dtype, ld, n, nullptr, HB_F32, out, nullptr, nullptr, 0, 0,
                    0, n, st);
  // host staging (synchronous per chunk: predict-only is not the hot path).  Large pageable inputs are
  // converted to float32 by host threads into a pinned buffer first (see hb_accumulate_flow).
  const size_t es = x_dtype == HB_F32 ? 4 : 8;
  const int D = f->desc.ndim;
  const bool threaded = x_mem == HB_MEM_HOST && n >= 65536 && !is_pinned_host(x);
  const long long chunk = threaded ? (long long)((64ull << 20) / ((size_t)D * 4)) : (1LL << 20);
  void* dx = nullptr;
  void* hpin = nullptr;
  float* dout = nullptr;
  const long long cap = n < chunk ? n : chunk;
  int rc = HB_OK;
  cudaError_t e = cudaSuccess;
  if (x_mem == HB_MEM_HOST) e = cudaMalloc(&dx, (size_t)cap * (threaded ? (size_t)D * 4 : (size_t)ld * es));
  if (e == cudaSuccess && threaded) e = cudaHostAlloc(&hpin, (size_t)cap * D * 4, cudaHostAllocDefault);
  if (e == cudaSuccess && out_mem == HB_MEM_HOST) e = cudaMalloc(&dout, (size_t)cap * sizeof(float));
  if (e != cudaSuccess) rc = cuda_fail(e, "staging allocation", __FILE__, __LINE__);
  const int nthreads = host_threads();
  for (long long a = 0; a < n && rc == HB_OK; a += chunk) {
    const long long b = (a + chunk < n) ? a + chunk : n;
    const void* xin = (const char*)x + (size_t)a * ld * es;
    int xdt = x_dtype;
    long long xld = ld;
    if (threaded) {
      if (x_dtype == HB_F64) convert_rows((const double*)x, ld, D, a, b, (float*)hpin, nthreads);
      else convert_rows((const float*)x, ld, D, a, b, (float*)hpin, nthreads);
      e = cudaMemcpyAsync(dx, hpin, (size_t)(b - a) * D * 4, cudaMemcpyHostToDevice, st);
      if (e != cudaSuccess) { rc = cuda_fail(e, "H2D", __FILE__, __LINE__); break; }
      xin = dx;
      xdt = HB_F32;
      xld = D;
    } else if (x_mem == HB_MEM_HOST) {
      e = cudaMemcpyAsync(dx, xin, (size_t)(b - a) * ld * es, cudaMemcpyHostToDevice, st);
      if (e != cudaSuccess) { rc = cuda_fail(e, "H2D", __FILE__, __LINE__); break; }
      xin = dx;
    }
    float* o = (out_mem == HB_MEM_HOST) ? dout : out + a;
    rc = flow_run(f, temperature, xin, xdt, xld, b - a, nullptr, HB_F32, o, nullptr, nullptr, 0, 0, 0, b - a, st);
    if (rc) break;
    if (out_mem == HB_MEM_HOST) {
      e = cudaMemcpyAsync(out + a, dout, (size_t)(b - a) * sizeof(float), cudaMemcpyDeviceToHost, st);
      if (e != cudaSuccess) { rc = cuda_fail(e, "D2H", __FILE__, __LINE__); break; }
    }
    e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { rc = cuda_fail(e, "sync", __FILE__, __LINE__); break; }
  }
  if (dx) cudaFree(dx);
  if (hpin) cudaFreeHost(hpin);
  if (dout) cudaFree(dout);
  return rc;
}

int hb_flow_sample(hb_flow* h, float temperature, const void* eps, int eps_dtype, int64_t n, int64_t ld,
                   float* out, int mem, void* stream) {
  HB_REQUIRE(h != nullptr, "hb_flow_sample: NULL handle");
  Flow* f = reinterpret_cast<Flow*>(h);
  HB_REQUIRE(temperature <= 1.f, "Scaling must not be greater than 1.");
  HB_REQUIRE(temperature > 0.f, "Scaling must be positive.");
  HB_REQUIRE(n >= 0 && ld >= f->desc.ndim, "bad shape");
  HB_REQUIRE(eps_dtype == HB_F32 || eps_dtype == HB_F64, "dtype must be HB_F32 or HB_F64");
  if (n == 0) return HB_OK;
  HB_REQUIRE(eps != nullptr && out != nullptr, "NULL buffer");
  HB_CUDA(cudaSetDevice(f->device));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int D = f->desc.ndim;
  if (mem == HB_MEM_DEVICE) return flow_sample_run(f, temperature, eps, eps_dtype, ld, n, out, D, st);
  // host buffers: staged in chunks (sampling is a diagnostic path, not the hot one)
  const size_t es = eps_dtype == HB_F32 ? 4 : 8;
  const long long chunk = 1LL << 20;
  const long long cap = n < chunk ? n : chunk;
  void* de = nullptr;
  float* dout = nullptr;
  int rc = HB_OK;
  cudaError_t e = cudaMalloc(&de, (size_t)cap * ld * es);
  if (e == cudaSuccess) e = cudaMalloc(&dout, (size_t)cap * D * sizeof(float));
  if (e != cudaSuccess) rc = cuda_fail(e, "staging allocation", __FILE__, __LINE__);
  for (long long a = 0; a < n && rc == HB_OK; a += chunk) {
    const long long b = (a + chunk < n) ? a + chunk : n;
    e = cudaMemcpyAsync(de, (const char*)eps + (size_t)a * ld * es, (size_t)(b - a) * ld * es,
                        cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) { rc = cuda_fail(e, "H2D", __FILE__, __LINE__); break; }
    rc = flow_sample_run(f, temperature, de, eps_dtype, ld, b - a, dout, D, st);
    if (rc) break;
    e = cudaMemcpyAsync(out + (size_t)a * D, dout, (size_t)(b - a) * D * sizeof(float), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { rc = cuda_fail(e, "D2H", __FILE__, __LINE__); break; }
  }
  if (de) cudaFree(de);
  if (dout) cudaFree(dout);
  return rc;
}

int hb_accumulate_flow(hb_accum* ha, hb_flow* hf, float temperature, const void* x, int x_dtype,
                       int64_t ld, const void* ln_posterior, int lp_dtype, int mem,
                       const int64_t* start_indices, int64_t chain_lo, int64_t chain_hi,
                       float* lnpred_out, int flags, void* stream) {
  HB_REQUIRE(ha != nullptr && hf != nullptr, "hb_accumulate_flow: NULL handle");
  Accum* acc = reinterpret_cast<Accum*>(ha);
  Flow* f = reinterpret_cast<Flow*>(hf);
  HB_REQUIRE(acc->device == f->device, "flow and accumulator live on different devices");
  HB_REQUIRE(temperature <= 1.f, "Scaling must not be greater than 1.");
  HB_REQUIRE(temperature > 0.f, "Scaling must be positive.");
  HB_REQUIRE(start_indices != nullptr, "start_indices is NULL");
  HB_REQUIRE(0 <= chain_lo && chain_lo <= chain_hi && chain_hi <= acc->nchains,
             "nchains do not match");
  HB_REQUIRE(ld >= f->desc.ndim, "Chains ndim inconsistent");
  HB_REQUIRE((x_dtype == HB_F32 || x_dtype == HB_F64) && (lp_dtype == HB_F32 || lp_dtype == HB_F64),
             "dtype must be HB_F32 or HB_F64");
  HB_CUDA(cudaSetDevice(acc->device));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long nrows = start_indices[chain_hi] - start_indices[chain_lo];
  const bool has_rows = nrows > 0 && chain_lo < chain_hi;
  int rc = (flags & HB_ACC_NEW_CALL) ? begin_call(acc, has_rows, st) : order_after_reset(acc, st);
  if (rc) return rc;
  if (!has_rows) return HB_OK;
  HB_REQUIRE(x != nullptr && ln_posterior != nullptr, "NULL input buffer");
  if (!guarded(f, nrows))
    return accumulate_flow_once(acc, f, temperature, x, x_dtype, ld, ln_posterior, lp_dtype, mem, start_indices,
                                chain_lo, chain_hi, nrows, lnpred_out, st);
  // guarded call: keep a copy of the (small) state block so that the call can be folded again on the CUDA cores
  const size_t state_bytes = (size_t)(4 * acc->nchains + G_N) * sizeof(double);
  if (!acc->shadow) HB_CUDA(cudaMalloc(&acc->shadow, state_bytes));
  HB_CUDA(cudaMemcpyAsync(acc->shadow, acc->packed, state_bytes, cudaMemcpyDeviceToDevice, st));
  rc = accumulate_flow_once(acc, f, temperature, x, x_dtype, ld, ln_posterior, lp_dtype, mem, start_indices,
                            chain_lo, chain_hi, nrows, lnpred_out, st);
  if (rc) return rc;
  long long cnt = 0;
  rc = guard_take(f, st, &cnt);
  if (rc || cnt == 0) return rc;
  f->guard_rows += cnt;
  ++f->guard_calls;
  ++acc->guard_trips;
  HB_CUDA(cudaMemcpyAsync(acc->packed, acc->shadow, state_bytes, cudaMemcpyDeviceToDevice, st));
  ForceSimt scope(f);
  return accumulate_flow_once(acc, f, temperature, x, x_dtype, ld, ln_posterior, lp_dtype, mem, start_indices,
                              chain_lo, chain_hi, nrows, lnpred_out, st);
}

}  // extern "C"

namespace hb {
int accumulate_flow_once(Accum* acc, Flow* f, float temperature, const void* x, int x_dtype, long long ld,
                         const void* ln_posterior, int lp_dtype, int mem, const int64_t* start_indices,
                         long long chain_lo, long long chain_hi, long long nrows, float* lnpred_out,
                         cudaStream_t st) {
  int rc = HB_OK;
  if (mem == HB_MEM_DEVICE)
    return flow_run(f, temperature, x, x_dtype, ld, nrows, ln_posterior, lp_dtype, lnpred_out, acc,
                    start_indices, chain_lo, chain_hi, 0, nrows, st);
  // HOST buffers.  Pageable memory (what numpy / harmonic.chains.Chains hand over, float64) would move
  // at ~11 GB/s through cudaMemcpyAsync; instead host threads convert it to float32 into pinned staging
  // buffers while the previous chunk is copied and computed (half the PCIe bytes, full PCIe rate).
  if (!is_pinned_host(x) || !is_pinned_host(ln_posterior))
    return accumulate_flow_pageable(acc, f, temperature, x, x_dtype, ld, ln_posterior, lp_dtype,
                                    start_indices, chain_lo, chain_hi, nrows, lnpred_out, st);
  // Pinned buffers: chunked double-buffered async copies straight from the caller's memory; the copy of
  // chunk i+1 overlaps the compute of chunk i
  const size_t es_x = x_dtype == HB_F32 ? 4 : 8, es_l = lp_dtype == HB_F32 ? 4 : 8;
  const size_t row_bytes = (size_t)ld * es_x + es_l + (lnpred_out ? 4 : 0);
  long long chunk = (long long)((256ull << 20) / row_bytes);
  chunk = (chunk / 4096) * 4096;
  if (chunk < 4096) chunk = 4096;
  if (chunk > nrows) chunk = ((nrows + 255) / 256) * 256;
  const size_t off_l = (((size_t)chunk * ld * es_x) + 255) & ~(size_t)255;
  const size_t off_o = (off_l + (size_t)chunk * es_l + 255) & ~(size_t)255;
  StageLease lease(acc->device, st);
  Stage* sg = lease.get();
  if (!sg) return HB_ERR_CUDA;
  rc = ensure_stage(sg, off_o + (lnpred_out ? (size_t)chunk * 4 : 0) + 256);
  if (rc) return rc;
  int i = 0;
  for (long long a = 0; a < nrows; a += chunk, ++i) {
    const long long b = (a + chunk < nrows) ? a + chunk : nrows;
    const int sl = i & 1;
    char* dxp = static_cast<char*>(sg->dev[sl]);
    char* dlp = dxp + off_l;
    float* dop = lnpred_out ? reinterpret_cast<float*>(dxp + off_o) : nullptr;
    HB_CUDA(cudaStreamWaitEvent(sg->copy_stream, sg->ev_done[sl], 0));
    HB_CUDA(cudaMemcpyAsync(dxp, (const char*)x + (size_t)a * ld * es_x, (size_t)(b - a) * ld * es_x,
                            cudaMemcpyHostToDevice, sg->copy_stream));
    HB_CUDA(cudaMemcpyAsync(dlp, (const char*)ln_posterior + (size_t)a * es_l, (size_t)(b - a) * es_l,
                            cudaMemcpyHostToDevice, sg->copy_stream));
    HB_CUDA(cudaEventRecord(sg->ev_copy[sl], sg->copy_stream));
    HB_CUDA(cudaStreamWaitEvent(st, sg->ev_copy[sl], 0));
    rc = flow_run(f, temperature, dxp, x_dtype, ld, b - a, dlp, lp_dtype, dop, acc, start_indices,
                  chain_lo, chain_hi, a, b, st);
    if (rc) return rc;
    if (dop)
      HB_CUDA(cudaMemcpyAsync(lnpred_out + a, dop, (size_t)(b - a) * 4, cudaMemcpyDeviceToHost, st));
    HB_CUDA(cudaEventRecord(sg->ev_done[sl], st));
  }
  HB_CUDA(cudaStreamSynchronize(st));
  return HB_OK;
}
}  // namespace hb

extern "C" {

int hb_host_alloc(size_t bytes, void** out) {
  HB_REQUIRE(out != nullptr, "hb_host_alloc: out is NULL");
  HB_CUDA(cudaHostAlloc(out, bytes, cudaHostAllocDefault));
  return HB_OK;
}
int hb_host_free(void* p) {
  if (p) HB_CUDA(cudaFreeHost(p));
  return HB_OK;
}
int hb_device_alloc(int device, size_t bytes, void** out) {
  HB_REQUIRE(out != nullptr, "hb_device_alloc: out is NULL");
  int rc = require_device(device);
  if (rc) return rc;
  HB_CUDA(cudaSetDevice(device));
  HB_CUDA(cudaMalloc(out, bytes));
  return HB_OK;
}
int hb_device_free(void* p) {
  if (p) HB_CUDA(cudaFree(p));
  return HB_OK;
}
int hb_memcpy_h2d(void* dst, const void* src, size_t bytes, void* stream) {
  HB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, reinterpret_cast<cudaStream_t>(stream)));
  return HB_OK;
}
int hb_memcpy_d2h(void* dst, const void* src, size_t bytes, void* stream) {
  HB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, reinterpret_cast<cudaStream_t>(stream)));
  HB_CUDA(cudaStreamSynchronize(reinterpret_cast<cudaStream_t>(stream)));
  return HB_OK;
}
int hb_stream_sync(void* stream) {
  HB_CUDA(cudaStreamSynchronize(reinterpret_cast<cudaStream_t>(stream)));
  return HB_OK;
}

}  // extern "C"
