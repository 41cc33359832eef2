// Developer micro-benchmark (GPU box): issue rate of tcgen05.mma kind::f16 for the operand arrangements the
// convolution kernels can use, with both operands in shared memory (no-swizzle K-major layouts exactly as
// csrc/conv_tc.cu builds them) or with the M-side operand in TMEM.  Prints SM cycles per k-step (K = 16).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/mma_probe tools/mma_probe.cu -lcuda
#include "../tdvc_b200/csrc/tc_common.cuh"
#include <cstdlib>
#include <vector>

using namespace tdvc::tc;

__device__ __forceinline__ void tc_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}

constexpr int SMEM = 200 * 1024;

// mode 0: pixels = M (two 8x16 tiles): per tile  N=128 (A_hi x [W_hi|W_lo]) + N=64 (A_lo x W_hi)      [current conv_tc]
// mode 1: weights = M ([W_hi;W_lo], 128 rows), pixels = N=256 (8x32 tile): W x X_hi + W x X_lo          [swapped, SS]
// mode 2: as 1 with the weights read from TMEM                                                           [swapped, TS]
// mode 3: two N=128 MMAs   mode 4: two N=64 MMAs   mode 5: one N=256 MMA   (pixels = M)
// mode 6: pixels = M, cout tile 128: N=256 (A_hi) + N=128 (A_lo), one tile
// mode 7: as 1 but two N=128 pixel tiles (8x16 each) per operand instead of one N=256
// mode 8 / 9: does an M = 64 MMA cost less than an M = 128 one?
__global__ void __launch_bounds__(128, 1) probe(int mode, int iters, unsigned long long* cycles) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar_s;
  __shared__ uint32_t tmem_slot;
  // small pseudo-random fp16 values
  for (int i = threadIdx.x; i < SMEM / 2; i += blockDim.x) {
    uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u;
    h ^= h >> 15;
    reinterpret_cast<__half*>(smem)[i] = __float2half_rn(((int)(h & 1023) - 512) * (1.f / 2048.f));
  }
  const uint32_t bar = smem_u32(&bar_s);
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  fence_async_smem();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t s0 = smem_u32(smem);
    // pixel operand: [ch/8][NPIXP][16 B] hi plane then lo plane (64 channels);  weights: blocks of [128 rows][64] fp16
    const bool wm = mode == 1 || mode == 2 || mode == 8 || mode == 9;
    const int IW = wm ? 10 : 18;
    const int NPIXP = (wm ? 10 * 34 : 18 * 18) | 1;
    const uint32_t LBO_X = NPIXP * 16, SBO_X = IW * 16, X_HALF = 8 * NPIXP * 16;
    const uint32_t x_hi = s0, x_lo = s0 + X_HALF, w0 = s0 + 2 * X_HALF;   // 3 weight blocks of 16 KB follow
    const uint32_t LBO_W = 128, SBO_W = 8 * 128, W_BLOCK = 128 * 64 * 2;
    const uint32_t I256 = instr_desc(256), I128 = instr_desc(128), I64 = instr_desc(64);
    const uint32_t I256_M64 = (1u << 4) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(64 >> 4) << 24);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      for (int tap = 0; tap < 9; ++tap) {
        const uint32_t wblk = w0 + (tap % 3) * W_BLOCK;
        const uint32_t xoff = ((tap / 3) * IW + tap % 3) * 16;
        for (int s = 0; s < 4; ++s) {
          const uint64_t wd = smem_desc(wblk + s * 2 * LBO_W, LBO_W, SBO_W);
          const uint64_t xh = smem_desc(x_hi + xoff + s * 2 * LBO_X, LBO_X, SBO_X);
          const uint64_t xl = smem_desc(x_lo + xoff + s * 2 * LBO_X, LBO_X, SBO_X);
          const uint64_t xh2 = smem_desc(x_hi + xoff + 128 + s * 2 * LBO_X, LBO_X, SBO_X);
          const uint64_t xl2 = smem_desc(x_lo + xoff + 128 + s * 2 * LBO_X, LBO_X, SBO_X);
          const uint32_t acc = (it | tap | s) != 0;
          switch (mode) {
            case 0:
              tc_mma(tmem, xh, wd, I128, acc);       tc_mma(tmem, xl, wd, I64, 1);
              tc_mma(tmem + 128, xh2, wd, I128, acc); tc_mma(tmem + 128, xl2, wd, I64, 1);
              break;
            case 1:
              tc_mma(tmem, wd, xh, I256, acc); tc_mma(tmem, wd, xl, I256, 1);
              break;
            case 2:
              tc_mma_ts(tmem, tmem + 256 + (tap * 4 + s) % 24 * 8, xh, I256, acc);
              tc_mma_ts(tmem, tmem + 256 + (tap * 4 + s) % 24 * 8, xl, I256, 1);
              break;
            case 3:
              tc_mma(tmem, xh, wd, I128, acc); tc_mma(tmem + 128, xh2, wd, I128, acc);
              break;
            case 4:
              tc_mma(tmem, xl, wd, I64, acc); tc_mma(tmem + 128, xl2, wd, I64, acc);
              break;
            case 5:
              tc_mma(tmem, xh, wd, I256, acc);
              break;
            case 6:
              tc_mma(tmem, xh, wd, I256, acc); tc_mma(tmem, xl, wd, I128, 1);
              break;
            case 7:
              tc_mma(tmem, wd, xh, I128, acc);        tc_mma(tmem, wd, xl, I128, 1);
              tc_mma(tmem + 128, wd, xh2, I128, acc); tc_mma(tmem + 128, wd, xl2, I128, 1);
              break;
            case 8:   // M = 64 rows of weights, N = 256 pixels, twice
              tc_mma(tmem, wd, xh, I256_M64, acc); tc_mma(tmem, wd, xl, I256_M64, 1);
              break;
            case 9:   // M = 128 (x_hi) + M = 64 (x_lo: only the W_hi rows are needed)
              tc_mma(tmem, wd, xh, I256, acc); tc_mma(tmem, wd, xl, I256_M64, 1);
              break;
          }
        }
      }
    }
    tc_commit(bar);
    mbar_wait(bar, 0);
    const long long t1 = clock64();
    cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

int main(int argc, char** argv) {
  const int iters = argc > 1 ? atoi(argv[1]) : 100;
  const char* names[] = {"pixels=M 2 tiles: N128 + N64        (floor 192)", "weights=M SS: 2 x N256 px           (floor 256)",
                         "weights=M TS: 2 x N256 px           (floor 256)", "pixels=M: 2 x N128                  (floor 128)",
                         "pixels=M: 2 x N64                   (floor  64)", "pixels=M: 1 x N256                  (floor 128)",
                         "pixels=M cout128: N256 + N128       (floor 192)", "weights=M SS: 4 x N128 px           (floor 256)",
                         "weights=M (M=64) SS: 2 x N256 px    (floor 256?)", "weights=M: M128 N256 + M64 N256     (floor 256?)"};
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
  unsigned long long* d;
  cudaMalloc(&d, 148 * sizeof(unsigned long long));
  for (int mode = 0; mode < 10; ++mode) {
    std::vector<unsigned long long> h(148);
    for (int rep = 0; rep < 2; ++rep) {
      probe<<<148, 128, SMEM>>>(mode, iters, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
    }
    cudaMemcpy(h.data(), d, 148 * 8, cudaMemcpyDeviceToHost);
    double sum = 0, mx = 0, mn = 1e30;
    for (auto c : h) { sum += c; mx = mx > c ? mx : c; mn = mn < c ? mn : (double)c; }
    const double ks = (double)iters * 36;
    printf("mode %d %s: cycles/k-step avg %.1f min %.1f max %.1f\n", mode, names[mode], sum / 148 / ks, mn / ks, mx / ks);
  }
  return 0;
}
