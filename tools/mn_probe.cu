// Developer probe: which shared-memory words does a tcgen05 kind::tf32 MMA read for an MN-major, 128-byte-swizzled A operand?
// A's region holds its own word indices; B (K-major, no swizzle: known-good in conv_tc.cu) is an 8x8 identity, so D[m][n] = A[k=n][m].
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tools/mn_probe tools/mn_probe.cu && tools/mn_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../tdvc_b200/csrc/tc_common.cuh"
namespace tdvc { void set_error(const char*, ...) {} }
using namespace tdvc::tc;

__global__ void probe(uint32_t start_off, uint32_t lbo, uint32_t sbo, uint32_t base_off, uint32_t swz, int div, float* out) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* gen = raw + (base - smem_u32(raw));
  float* A = reinterpret_cast<float*>(gen);                // 32 KB region of word indices
  float* B = reinterpret_cast<float*>(gen + 32768);        // K-major no-swizzle: [k/4][n][k%4], 16 rows
  uint32_t* slot = reinterpret_cast<uint32_t*>(gen + 32768 + 1024);
  const uint32_t bar = base + 32768 + 1024 + 16;
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) A[i] = (float)(div ? i / 1024 : i % 1024);
  for (int i = threadIdx.x; i < 128; i += blockDim.x) {    // core matrix cm = i / 64 (k 0-3 | 4-7), row n = (i / 4) % 16, kk = i % 4
    const int cm = i / 64, n = (i / 4) % 16, kk = i % 4;
    B[i] = (n == cm * 4 + kk) ? 1.f : 0.f;
  }
  if (threadIdx.x == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  fence_async_smem();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(32u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(slot);
  if (threadIdx.x == 0) {
    const uint32_t a_addr = base + start_off;
    const uint64_t adesc = (uint64_t)((a_addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) |
                           ((uint64_t)(base_off & 7) << 49) | ((uint64_t)swz << 61);
    const uint64_t bdesc = smem_desc(base + 32768, 256, 128);   // LBO: k 0-3 -> k 4-7 core matrix (16 rows x 16 B = 256 B), SBO: 8 rows
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (0u << 16) | ((16u >> 3) << 17) | ((128u >> 4) << 24);
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(0u) : "memory");
    tc_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  uint32_t v[16];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16), v);
  tmem_ld_wait();
  for (int j = 0; j < 16; ++j) out[(warp * 32 + lane) * 16 + j] = __uint_as_float(v[j]);
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32u) : "memory");
}

int main() {
  float* out;
  cudaMalloc(&out, 128 * 16 * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40960);
  static float lo[2048], hi[2048];
  struct { uint32_t off, lbo, sbo, bo, swz; } cases[] = {
      {0, 1024, 512, 0, 1}, {0, 5120, 512, 0, 1}, {128, 5120, 512, 0, 1}, {128, 5120, 512, 1, 1}, {256, 5120, 512, 2, 1},
      {384, 5120, 512, 0, 1}, {512, 5120, 512, 0, 1}, {640, 5120, 512, 0, 1}, {0, 5120, 1024, 0, 1}};
  for (auto& c : cases) {
    for (int div = 0; div < 2; ++div) {
      probe<<<1, 128, 40960>>>(c.off, c.lbo, c.sbo, c.bo, c.swz, div, out);
      cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess) { printf("launch error %s\n", cudaGetErrorString(e)); return 1; }
      e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
      cudaMemcpy(div ? hi : lo, out, sizeof(lo), cudaMemcpyDeviceToHost);
    }
    printf("== start +%u LBO %u SBO %u base_off %u swizzle %u: byte offset read for A[k][m]\n", c.off, c.lbo, c.sbo, c.bo, c.swz);
    for (int k = 0; k < 8; ++k) {
      printf(" k=%d:", k);
      for (int m : {0, 1, 3, 4, 8, 31, 32, 33, 64, 96, 127}) printf(" m%d@%d", m, 4 * ((int)hi[m * 16 + k] * 1024 + (int)lo[m * 16 + k]));
      printf("\n");
    }
    printf(" raw lo row m=5:"); for (int n = 0; n < 16; ++n) printf(" %g", lo[5 * 16 + n]); printf("\n");
    int extra = 0;
    for (int m = 0; m < 128; ++m) for (int n = 8; n < 16; ++n) extra += lo[m * 16 + n] != 0.f;
    printf(" nonzero in columns 8-15: %d\n", extra);
  }
  return 0;
}
