// Shared helpers for the tdvc_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/tdvc_b200.h"

namespace tdvc {

void set_error(const char* fmt, ...);

#define TDVC_REQUIRE(cond, ...)                 \
  do {                                          \
    if (!(cond)) {                              \
      ::tdvc::set_error(__VA_ARGS__);           \
      return TDVC_EINVAL;                       \
    }                                           \
  } while (0)

#define TDVC_CHECK_LAUNCH(name)                                                    \
  do {                                                                             \
    cudaError_t e_ = cudaGetLastError();                                           \
    if (e_ != cudaSuccess) {                                                       \
      ::tdvc::set_error("%s: CUDA launch failed: %s", name, cudaGetErrorString(e_)); \
      return TDVC_ECUDA;                                                           \
    }                                                                              \
  } while (0)

static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

__device__ __forceinline__ float apply_act(float v, int act, float slope) {
  if (act == TDVC_ACT_RELU) return fmaxf(v, 0.f);
  if (act == TDVC_ACT_LRELU) return v > 0.f ? v : v * slope;
  if (act == TDVC_ACT_CLAMP01) return fminf(fmaxf(v, 0.f), 1.f);
  return v;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a per-device setting: `done` (one static array per launch site)
// remembers the largest size already set on each device, so that a process driving several GPUs (nn.DataParallel
// replicas) sets it on every one of them.  A racing second call sets the same value again, which is harmless.
constexpr int kMaxDevices = 64;
template <class K>
static inline int ensure_dynamic_smem(K kernel, size_t bytes, int (&done)[kMaxDevices], const char* name) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) dev = kMaxDevices - 1;
  if ((size_t)done[dev] >= bytes) return TDVC_OK;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) {
    set_error("%s: cudaFuncSetAttribute(%d bytes) failed: %s", name, (int)bytes, cudaGetErrorString(e));
    return TDVC_ECUDA;
  }
  done[dev] = (int)bytes;
  return TDVC_OK;
}

// 148 SMs on B200; persistent / grid-stride launches size their grids from this.
constexpr int kNumSMs = 148;

}  // namespace tdvc
