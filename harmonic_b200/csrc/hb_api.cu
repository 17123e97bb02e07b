// C ABI glue: error state, device checks, flow handles, predict / accumulate dispatch with
// host-buffer staging.  See include/harmonic_b200.h for the contract.
#include <math.h>
#include <string.h>

#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "hb_internal.h"

namespace hb {

static thread_local std::string g_err;
static std::atomic<long long> g_launches{0};
static bool g_timing = false;
static cudaEvent_t g_ev0 = nullptr, g_ev1 = nullptr;
static bool g_timing_pending = false;

void set_error(const std::string& msg) { g_err = msg; }
void count_launch(int n) { g_launches += n; }

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  g_err = std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what + " (" + file + ":" +
          std::to_string(line) + ")";
  return HB_ERR_CUDA;
}

void timing_begin(cudaStream_t st) {
  if (!g_timing) return;
  if (!g_ev0) {
    cudaEventCreate(&g_ev0);
    cudaEventCreate(&g_ev1);
  }
  cudaEventRecord(g_ev0, st);
}
void timing_end(cudaStream_t st) {
  if (!g_timing) return;
  cudaEventRecord(g_ev1, st);
  g_timing_pending = true;
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
  if (prop.major != 10) {
    set_error(std::string("device is sm_") + std::to_string(prop.major * 10 + prop.minor) +
              "; libharmonic_b200 is built for sm_100a only");
    return HB_ERR_NO_DEVICE;
  }
  if (device < 64) ok_mask.fetch_or(1ull << device);
  return HB_OK;
}

int ensure_pinned(Accum* acc, size_t bytes_per_slot) {
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

int ensure_stage(Accum* acc, size_t bytes_per_slot) {
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

static int flow_run(Flow* f, float T, const void* x, int x_dtype, long long ld, long long n,
                    const void* lnpost, int lp_dtype, float* lnpred_out, Accum* acc,
                    const int64_t* start, long long chain_lo, long long chain_hi, long long a,
                    long long b, cudaStream_t st) {
  if (f->desc.kind == HB_FLOW_RQSPLINE)
    return rqspline_run(f, T, x, x_dtype, ld, n, lnpost, lp_dtype, lnpred_out, acc, start, chain_lo,
                        chain_hi, a, b, st);
  const bool tc = (f->path == HB_PATH_TCGEN05) || (f->path == HB_PATH_AUTO && f->tc_ok);
  if (tc)
    return realnvp_tc_run(f, T, x, x_dtype, ld, n, lnpost, lp_dtype, lnpred_out, acc, start,
                          chain_lo, chain_hi, a, b, st);
  return realnvp_simt_run(f, T, x, x_dtype, ld, n, lnpost, lp_dtype, lnpred_out, acc, start,
                          chain_lo, chain_hi, a, b, st);
}


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
  int rc = ensure_stage(acc, slot);
  if (rc) return rc;
  rc = ensure_pinned(acc, off_o + 256);
  if (rc) return rc;
  int i = 0;
  for (long long a = 0; a < nrows; a += chunk, ++i) {
    const long long b = (a + chunk < nrows) ? a + chunk : nrows;
    const int sl = i & 1;
    char* hp = static_cast<char*>(acc->pin[sl]);
    char* dxp = static_cast<char*>(acc->dev[sl]);
    // the pinned slot is free once the H2D copy of chunk i-2 has been consumed
    HB_CUDA(cudaEventSynchronize(acc->ev_copy[sl]));
    if (x_dtype == HB_F64) convert_rows((const double*)x, ld, D, a, b, reinterpret_cast<float*>(hp), nthreads);
    else convert_rows((const float*)x, ld, D, a, b, reinterpret_cast<float*>(hp), nthreads);
    if (lp_dtype == HB_F64) convert_rows((const double*)lnpost, 1, 1, a, b, reinterpret_cast<float*>(hp + off_l), nthreads);
    else convert_rows((const float*)lnpost, 1, 1, a, b, reinterpret_cast<float*>(hp + off_l), nthreads);
    HB_CUDA(cudaStreamWaitEvent(acc->copy_stream, acc->ev_done[sl], 0));  // device slot free
    HB_CUDA(cudaMemcpyAsync(dxp, hp, off_l + (size_t)(b - a) * 4, cudaMemcpyHostToDevice, acc->copy_stream));
    HB_CUDA(cudaEventRecord(acc->ev_copy[sl], acc->copy_stream));
    HB_CUDA(cudaStreamWaitEvent(st, acc->ev_copy[sl], 0));
    float* dop = lnpred_out ? reinterpret_cast<float*>(dxp + off_o) : nullptr;
    rc = flow_run(f, temperature, dxp, HB_F32, D, b - a, dxp + off_l, HB_F32, dop, acc, start_indices,
                  chain_lo, chain_hi, a, b, st);
    if (rc) return rc;
    if (dop)
      HB_CUDA(cudaMemcpyAsync(lnpred_out + a, dop, (size_t)(b - a) * 4, cudaMemcpyDeviceToHost, st));
    HB_CUDA(cudaEventRecord(acc->ev_done[sl], st));
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
    if (cudaGetDeviceProperties(&prop, i) == cudaSuccess && prop.major == 10) ++ok;
  }
  if (ok > 0) cached = ok;
  return ok;
}

int64_t hb_launch_count(void) { return g_launches.load(); }

int hb_set_kernel_timing(int on) {
  g_timing = on != 0;
  return HB_OK;
}
double hb_last_kernel_ms(void) {
  if (!g_timing_pending || !g_ev0) return -1.0;
  float ms = 0.f;
  if (cudaEventSynchronize(g_ev1) != cudaSuccess) return -1.0;
  if (cudaEventElapsedTime(&ms, g_ev0, g_ev1) != cudaSuccess) return -1.0;
  return (double)ms;
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
    rc = realnvp_tc_prepare(f);  // sets tc_ok when the shape is supported
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
  if (f->desc.kind != HB_FLOW_REALNVP) return 0;
  return (f->path == HB_PATH_TCGEN05) || (f->path == HB_PATH_AUTO && f->tc_ok);
}

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
  if (flags & HB_ACC_NEW_CALL) {
    int rc = reset_call_scalars(acc, st);
    if (rc) return rc;
  }
  const long long nrows = start_indices[chain_hi] - start_indices[chain_lo];
  if (nrows == 0 || chain_lo == chain_hi) return HB_OK;
  HB_REQUIRE(x != nullptr && ln_posterior != nullptr, "NULL input buffer");
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
  int rc = ensure_stage(acc, off_o + (lnpred_out ? (size_t)chunk * 4 : 0) + 256);
  if (rc) return rc;
  int i = 0;
  for (long long a = 0; a < nrows; a += chunk, ++i) {
    const long long b = (a + chunk < nrows) ? a + chunk : nrows;
    const int sl = i & 1;
    char* dxp = static_cast<char*>(acc->dev[sl]);
    char* dlp = dxp + off_l;
    float* dop = lnpred_out ? reinterpret_cast<float*>(dxp + off_o) : nullptr;
    HB_CUDA(cudaStreamWaitEvent(acc->copy_stream, acc->ev_done[sl], 0));
    HB_CUDA(cudaMemcpyAsync(dxp, (const char*)x + (size_t)a * ld * es_x, (size_t)(b - a) * ld * es_x,
                            cudaMemcpyHostToDevice, acc->copy_stream));
    HB_CUDA(cudaMemcpyAsync(dlp, (const char*)ln_posterior + (size_t)a * es_l, (size_t)(b - a) * es_l,
                            cudaMemcpyHostToDevice, acc->copy_stream));
    HB_CUDA(cudaEventRecord(acc->ev_copy[sl], acc->copy_stream));
    HB_CUDA(cudaStreamWaitEvent(st, acc->ev_copy[sl], 0));
    rc = flow_run(f, temperature, dxp, x_dtype, ld, b - a, dlp, lp_dtype, dop, acc, start_indices,
                  chain_lo, chain_hi, a, b, st);
    if (rc) return rc;
    if (dop)
      HB_CUDA(cudaMemcpyAsync(lnpred_out + a, dop, (size_t)(b - a) * 4, cudaMemcpyDeviceToHost, st));
    HB_CUDA(cudaEventRecord(acc->ev_done[sl], st));
  }
  HB_CUDA(cudaStreamSynchronize(st));
  return HB_OK;
}

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
