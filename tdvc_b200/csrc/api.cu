// C-ABI plumbing: error text, version, conv dispatch.
#include <stdarg.h>
#include <string.h>
#include "common.cuh"

namespace tdvc {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int conv2d_validate(const TdvcConvParams* p);
int conv2d_simt(const TdvcConvParams& p, cudaStream_t st);
int conv2d_tc_supported(const TdvcConvParams& p);
int conv2d_tc(const TdvcConvParams& p, cudaStream_t st, int* rows_only = nullptr);
int conv2d_small_supported(const TdvcConvParams& p);
int conv2d_small(const TdvcConvParams& p, cudaStream_t st);
}  // namespace tdvc
extern "C" int tdvc_conv2d_f16_is_split(const TdvcConvParams* p);

extern "C" int tdvc_version(void) { return 100; }
extern "C" const char* tdvc_last_error(void) { return tdvc::g_err; }

extern "C" int tdvc_conv2d(const TdvcConvParams* p, void* stream) {
  int rc = tdvc::conv2d_validate(p);
  if (rc != TDVC_OK) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (p->impl == 2) {
    if (!tdvc::conv2d_tc_supported(*p)) {
      tdvc::set_error("conv2d: impl=2 (tcgen05) does not support this shape");
      return TDVC_EINVAL;
    }
    return tdvc::conv2d_tc(*p, st);
  }
  if ((p->out_absmax != nullptr || p->chan_sum != nullptr) &&
      (p->impl == 3 || (p->impl == 0 && p->products == 0 && tdvc::conv2d_small_supported(*p)))) {
    tdvc::set_error("conv2d: out_absmax / chan_sum are not available on the <= 4-channel kernel");
    return TDVC_EINVAL;
  }
  if (p->impl == 3) return tdvc::conv2d_small(*p, st);
  // (a one-product layer stays on the tensor cores even with <= 4 output channels: the exact fp32 SIMT kernel is slower there)
  if (p->impl == 0 && p->products == 0 && tdvc::conv2d_small_supported(*p)) return tdvc::conv2d_small(*p, st);
  if (p->impl == 0 && tdvc::conv2d_tc_supported(*p)) return tdvc::conv2d_tc(*p, st);
  if (p->chan_sum != nullptr) {
    tdvc::set_error("conv2d: chan_sum is produced by the tcgen05 kernel only (ask tdvc_conv2d_chan_sum_rows first)");
    return TDVC_EINVAL;
  }
  return tdvc::conv2d_simt(*p, st);
}

// which kernel tdvc_conv2d would run for *p and what it costs the tensor pipe: 0 = an exact fp32 SIMT kernel (no MMA),
// otherwise the fp16 MMA products issued per algorithmic MAC (4 = hi/lo rows, 3 = split scheme, 1 = one product)
extern "C" int tdvc_conv2d_products(const TdvcConvParams* p) {
  if (p == nullptr || tdvc::conv2d_validate(p) != TDVC_OK) return -1;
  if (p->impl == 1 || p->impl == 3) return 0;
  if (p->impl == 0 && p->products == 0 && tdvc::conv2d_small_supported(*p)) return 0;
  if (!tdvc::conv2d_tc_supported(*p)) return p->impl == 2 ? -1 : 0;
  if (p->products == 1) return 1;
  return tdvc_conv2d_f16_is_split(p) ? 3 : 4;
}

// rows of the `chan_sum` buffer ([rows][N][cout] floats) tdvc_conv2d would fill for *p; 0 = this launch cannot produce channel
// sums (an exact fp32 SIMT kernel would run, several output-channel tiles, planar or pixel-shuffled output)
extern "C" int tdvc_conv2d_chan_sum_rows(const TdvcConvParams* p) {
  if (tdvc_conv2d_products(p) <= 0) return 0;
  int rows = 0;
  if (tdvc::conv2d_tc(*p, nullptr, &rows) != TDVC_OK) return 0;
  return rows;
}
