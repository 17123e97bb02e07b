// tcgen05 RealNVP path (placeholder until the kernel lands).
#include "hb_internal.h"
namespace hb {
int realnvp_tc_prepare(Flow* f) { f->tc_ok = false; return HB_OK; }
void realnvp_tc_release(Flow* f) { if (f->tc_image) cudaFree(f->tc_image); f->tc_image = nullptr; }
int realnvp_tc_run(Flow*, float, const void*, int, long long, long long, const void*, int, float*, Accum*,
                   const int64_t*, long long, long long, long long, long long, cudaStream_t) {
  set_error("tcgen05 path not available for this flow");
  return HB_ERR_UNSUPPORTED;
}
}  // namespace hb
