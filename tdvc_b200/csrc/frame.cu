// Per-frame plumbing kernels: accumulator reset, content hash of the reference frames (cache keys), bpp finish.
#include "common.cuh"

namespace tdvc {

__device__ __forceinline__ uint64_t mix64(uint64_t z) {   // splitmix64 finaliser
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

// out[2*s], out[2*s+1] += sum_i mix(word_i, i) with two different mixes.  Integer sums: order-free, hence deterministic.
__global__ void slices_hash_kernel(const uint32_t* __restrict__ base, int64_t words, int64_t stride_words, uint64_t* __restrict__ out) {
  const int s = blockIdx.y;
  const uint32_t* p = base + (int64_t)s * stride_words;
  uint64_t h0 = 0, h1 = 0;
  const int64_t n4 = words >> 2;
  const bool vec = (reinterpret_cast<uintptr_t>(p) & 15) == 0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (vec) {
    for (int64_t i = t; i < n4; i += stride) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(p) + i);
      const uint64_t a = ((uint64_t)v.y << 32) | v.x, b = ((uint64_t)v.w << 32) | v.z;
      const uint64_t k = (uint64_t)i * 0x9E3779B97F4A7C15ull;
      h0 += mix64(a + k) + mix64(b ^ (k + 0x632BE59BD9B4E019ull));
      h1 += mix64((a ^ 0xD6E8FEB86659FD93ull) * 3 + (k >> 1)) ^ mix64(b + ~k);
    }
    for (int64_t i = (n4 << 2) + t; i < words; i += stride) {
      const uint64_t a = p[i], k = (uint64_t)i * 0xC2B2AE3D27D4EB4Full;
      h0 += mix64(a + k);
      h1 += mix64(~a ^ k);
    }
  } else {
    for (int64_t i = t; i < words; i += stride) {
      const uint64_t a = p[i], k = (uint64_t)i * 0xC2B2AE3D27D4EB4Full;
      h0 += mix64(a + k);
      h1 += mix64(~a ^ k);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    h0 += __shfl_xor_sync(0xffffffffu, h0, o);
    h1 += __shfl_xor_sync(0xffffffffu, h1, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(reinterpret_cast<unsigned long long*>(out + 2 * s), (unsigned long long)h0);
    atomicAdd(reinterpret_cast<unsigned long long*>(out + 2 * s + 1), (unsigned long long)h1);
  }
}

__global__ void bpp_finish_kernel(const double* __restrict__ acc, float* __restrict__ bpp, double scale) {
  if (threadIdx.x < 2) bpp[threadIdx.x] = (float)((acc[2 * threadIdx.x] + acc[2 * threadIdx.x + 1]) * scale);
}

}  // namespace tdvc

using namespace tdvc;

extern "C" int tdvc_zero_bytes(void* p, size_t n, void* stream) {
  TDVC_REQUIRE(p != nullptr || n == 0, "zero_bytes: null pointer");
  if (n == 0) return TDVC_OK;
  const cudaError_t e = cudaMemsetAsync(p, 0, n, (cudaStream_t)stream);
  if (e != cudaSuccess) {
    set_error("zero_bytes: cudaMemsetAsync failed: %s", cudaGetErrorString(e));
    return TDVC_ECUDA;
  }
  return TDVC_OK;
}

extern "C" int tdvc_slices_hash(const void* base, int64_t slice_words, int64_t stride_words, int n_slices, uint64_t* out, void* stream) {
  TDVC_REQUIRE(base && out && slice_words > 0 && n_slices > 0 && n_slices <= 65535, "slices_hash: bad args");
  TDVC_REQUIRE((reinterpret_cast<uintptr_t>(base) & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 7) == 0, "slices_hash: alignment");
  if (int rc = tdvc_zero_bytes(out, (size_t)n_slices * 16, stream)) return rc;
  int gx = cdiv(slice_words / 4 + 1, 256 * 4);
  const int cap = cdiv(kNumSMs * 8, n_slices);
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  slices_hash_kernel<<<dim3(gx, n_slices), 256, 0, (cudaStream_t)stream>>>(static_cast<const uint32_t*>(base), slice_words, stride_words, out);
  TDVC_CHECK_LAUNCH("slices_hash");
  return TDVC_OK;
}

extern "C" int tdvc_bpp_finish(const double* acc4, float* bpp2, double scale, void* stream) {
  TDVC_REQUIRE(acc4 && bpp2, "bpp_finish: null pointer");
  bpp_finish_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(acc4, bpp2, scale);
  TDVC_CHECK_LAUNCH("bpp_finish");
  return TDVC_OK;
}
