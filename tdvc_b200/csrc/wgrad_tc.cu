// Weight gradient of the 3x3 stride-1 convolutions on the tensor cores (the training step; the reference
// takes this gradient from cuDNN through autograd, tools/train.py:125-159).  One TF32 product per MAC: the arithmetic of
// `enabled_amp=True` (the reference's shipped cfg/train.yaml; conv_bwd.cu keeps the fp32-class three-product kernel).
//
//   d W[tap][ci][co] = sum over pixels p of x[p + tap][ci] * g[p][co]
//
// is a GEMM whose reduction index is the PIXEL, and both operands live in channels-last tensors: a row of the operand (one
// pixel) holds the M (ci) or N (co) index contiguously.  That IS tcgen05's MN-major operand layout, so nothing is transposed
// and no thread touches an operand: tensor-map bulk copies (5-D maps over (32 channels, x, 32-channel group, y, image), 128-byte
// swizzle with 32-byte atoms) drop a patch of x with its halo and the matching patch of g into shared memory, and every MMA reads 8 pixels x
// (128 | 64) channels of them through MN-major descriptors (the first warp-level kernel staged 32-pixel chunks and issued
// m16n8k8 MMAs from scalar fragment loads: 0.85 ms for 64->64 at 8x256x256; a K-major tcgen05 version had to transpose with
// 4-byte cp.async and was slower still - conv_bwd.cu).
//
// Work item = R = 4 output rows x BW = 32 output columns of one image: x patch 6 rows x 36 columns (34 needed; 36 keeps every
// (row, channel-group) block a whole number of 512-byte swizzle atoms), g patch 4 x 32.  Shared-memory order [row][group of 32
// channels][column][32 channels]: consecutive (row, group) blocks are equally spaced, so ONE descriptor with that spacing as its
// leading-dimension offset spans the 64 input channels of kernel row ky AND of kernel row ky + 1 - an M = 128 operand for a
// 64-channel tile (M = 64 would run the tensor pipe at half rate).  3x3: per kernel column two accumulators [ky 0 | ky 1] and
// [ky 1 (dropped) | ky 2]: 6 x 64 TMEM columns.  A tap's shift is a shift of the operand's start address by whole pixels
// (128-byte rows inside the swizzle pattern, whose phase follows the absolute address).
// CTA = (64 ci, 64 co) tile x slab s: items s, s + S, ...; accumulators stay in TMEM over the whole slab; partial sums go to the
// workspace [S][tap][cin][cout] and conv_bwd.cu's fixed-order reduction adds them (deterministic).  Warp 0: bulk copies,
// warp 1: MMA issue, warps 2-5: column sums of g for the bias gradient from the staged patches, then the TMEM read-out.
#include <cuda.h>
#include <cstdlib>

#include "tc_common.cuh"

namespace tdvc {
namespace wgtc {

using namespace tc;

constexpr int R = 4, BW = 32, XW = 36, XROWS = R + 2;
constexpr int HALF_X = XW * 128, ROW_X = 2 * HALF_X, X_BYTES = XROWS * ROW_X;   // 4608, 9216, 55296
constexpr int HALF_G = BW * 128, ROW_G = 2 * HALF_G, G_BYTES = R * ROW_G;       // 4096, 8192, 32768
constexpr int STAGE = X_BYTES + G_BYTES;                                        // 88064 = 86 x 1024
constexpr int NSTAGE = 2;
constexpr int SMEM = NSTAGE * STAGE + 1024 /* alignment slack */ + 128 /* barriers, TMEM slot, bias exchange */ + 512;
constexpr int THREADS = 192;
constexpr int TMEM_COLS = 512;

struct Maps {
  CUtensorMap x, g;
};

struct Args {
  int N, H, W, cin, cout;
  int S, co_tiles, n_items, tiles_x, tiles_y;
  float* part;       // [S][9][cin][cout]
  float* part_bias;  // [S][cout]
};

// MN-major 32-bit operands have ONE legal shared-memory layout: "128-byte swizzle with 32-byte atoms" (layout type 1; the tensor
// map's SWIZZLE_128B_ATOM_32B): an atom is 4 rows (K = 4 pixels) x 128 bytes (32 channels), the 32-byte chunk index of a row is
// XORed with (row & 3).  LBO = distance between 32-channel atoms, SBO = distance between the two 4-pixel groups of a K = 8 MMA.
// Measured with tools/mn_probe.cu: the XOR phase comes from the ABSOLUTE address bits 7-8, so a start address moved by whole
// 128-byte rows reads the rows behind it with the phases the bulk copy wrote them with (base-offset field 0); the plain
// 128-byte swizzle (type 2) and the unswizzled layout make a kind::tf32 MMA with MN-major operands write zeros.
__device__ __forceinline__ uint64_t mn_desc(uint32_t addr, uint32_t lbo) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (1ull << 61);
}
// kind::tf32: D = f32, A = B = tf32 (format 2), both MN-major (bits 15, 16), M = 128, N = 64
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((64u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(kIdesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tma5(uint32_t dst, const CUtensorMap* m, int c0, int c1, int c2, int c3, int c4, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(bar)
      : "memory");
}

__global__ void __launch_bounds__(THREADS, 1) wgrad_tc_kernel(const Args a, const __grid_constant__ Maps maps) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* const gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bars = base + NSTAGE * STAGE;            // full[2], empty[2], acc
  uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(gen + NSTAGE * STAGE + 64);
  float* const bias_x = reinterpret_cast<float*>(gen + NSTAGE * STAGE + 128);   // [128]
  auto full = [&](int s) { return bars + 8u * s; };
  auto empty = [&](int s) { return bars + 16u + 8u * s; };
  const uint32_t acc_bar = bars + 32u;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x, s = blockIdx.y;
  const int co_t = tile % a.co_tiles, ci_t = tile / a.co_tiles;

  if (threadIdx.x == 0) {
    for (int i = 0; i < NSTAGE; ++i) { mbar_init(full(i), 1); mbar_init(empty(i), 5); }
    mbar_init(acc_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  if (warp == 0) {
    // ------------------------------------------------------------------ bulk copies: one x patch + one g patch per item
    if (elect_one()) {
      int st = 0, ph = 1;
      for (int item = s; item < a.n_items; item += a.S) {
        int r = item;
        const int tx = r % a.tiles_x; r /= a.tiles_x;
        const int ty = r % a.tiles_y;
        const int n = r / a.tiles_y;
        mbar_wait(empty(st), ph);
        mbar_expect_tx(full(st), STAGE);
        const uint32_t dst = base + st * STAGE;
        tma5(dst, &maps.x, 0, tx * BW - 1, 2 * ci_t, ty * R - 1, n, full(st));
        tma5(dst + X_BYTES, &maps.g, 0, tx * BW, 2 * co_t, ty * R, n, full(st));
        if (++st == NSTAGE) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issue: 16 K-steps x 3 kernel columns x 2 row pairs
    int st = 0, ph = 0;
    uint32_t first = 1;
    for (int item = s; item < a.n_items; item += a.S) {
      mbar_wait(full(st), ph);
      tc_fence_after();
      if (elect_one()) {
        uint32_t acc = first ^ 1u;
        const uint32_t xs = base + st * STAGE, gs = xs + X_BYTES;
#pragma unroll 1
        for (int r = 0; r < R; ++r) {
#pragma unroll
          for (int c0 = 0; c0 < BW; c0 += 8) {
            const uint64_t bdesc = mn_desc(gs + r * ROW_G + c0 * 128, HALF_G);
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
#pragma unroll
              for (int p = 0; p < 2; ++p) {
                const uint64_t adesc = mn_desc(xs + (r + p) * ROW_X + (c0 + kx) * 128, HALF_X);
                mma_tf32(tmem_base + (kx * 2 + p) * 64, adesc, bdesc, acc);
              }
            }
            acc = 1;
          }
        }
        tc_commit(empty(st));
      }
      first = 0;
      __syncwarp();
      if (++st == NSTAGE) { st = 0; ph ^= 1; }
    }
    if (elect_one()) tc_commit(acc_bar);
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ bias sums from the staged g patches, then the read-out
    const int et = threadIdx.x - 64;                 // 0..127
    const int co = et & 63, part = et >> 6;          // rows 2 part, 2 part + 1 of every patch
    const bool do_bias = a.part_bias != nullptr && ci_t == 0;
    float bsum = 0.f;
    int st = 0, ph = 0;
    for (int item = s; item < a.n_items; item += a.S) {
      mbar_wait(full(st), ph);
      if (do_bias) {
        const uint8_t* gs = gen + st * STAGE + X_BYTES + (co >> 5) * HALF_G + (co & 7) * 4;
        const int ch = (co & 31) >> 3;
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
          const uint8_t* row = gs + (part * 2 + rr) * ROW_G;
#pragma unroll 8
          for (int px = 0; px < BW; ++px) bsum += *reinterpret_cast<const float*>(row + px * 128 + ((ch ^ (px & 3)) << 5));
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty(st));
      if (++st == NSTAGE) { st = 0; ph ^= 1; }
    }
    if (do_bias) {
      bias_x[et] = bsum;
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (part == 0) a.part_bias[(int64_t)s * a.cout + co_t * 64 + co] = bias_x[co] + bias_x[64 + co];
    }
    mbar_wait(acc_bar, 0);
    tc_fence_after();
    const int q = warp & 3;                          // TMEM lane quarter this warp may read
    const int L = q * 32 + lane, ci = L & 63, upper = L >> 6;
#pragma unroll 1
    for (int acc_i = 0; acc_i < 6; ++acc_i) {
      const int kx = acc_i >> 1, p = acc_i & 1;
      if (p == 1 && upper == 0) continue;            // the duplicate of kernel row 1 (warp-uniform: `upper` is per warp)
      const int tap = (p + upper) * 3 + kx;
      float* dst = a.part + (((int64_t)s * 9 + tap) * a.cin + ci_t * 64 + ci) * a.cout + co_t * 64;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc_i * 64 + h * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(dst + h * 32 + j) =
              make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
  }
}

using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeFn encoder() {
  static EncodeFn encode = nullptr;   // idempotent; a benign race resolves it twice
  if (encode == nullptr) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || fn == nullptr ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    encode = reinterpret_cast<EncodeFn>(fn);
  }
  return encode;
}

static int slabs(int n_items, int tiles) {
  int S = kNumSMs / tiles;
  if (S < 1) S = 1;
  if (S > n_items) S = n_items;
  return S;
}

}  // namespace wgtc

// shapes the tensor-core weight gradient takes: 3x3, stride 1, pad 1, 64-channel tiles on both sides, whole 4 x 32 items
bool wgrad_tc_eligible(const float* x, int x_ld, const float* g, int g_ld, int H, int W, int cin, int cout, int k, int stride, int pad,
                       int in_square) {
  static const bool off = getenv("TDVC_B200_WGRAD_SIMT") != nullptr;   // developer A/B switch: warp-level MMA kernel everywhere
  if (off || k != 3 || stride != 1 || pad != 1 || in_square) return false;
  if (cin % 64 != 0 || cout % 64 != 0 || H % wgtc::R != 0 || W % wgtc::BW != 0) return false;
  if ((reinterpret_cast<uintptr_t>(x) & 15) != 0 || (reinterpret_cast<uintptr_t>(g) & 15) != 0 || x_ld % 4 != 0 || g_ld % 4 != 0) return false;
  return true;
}

size_t wgrad_tc_workspace_bytes(int N, int H, int W, int cin, int cout) {
  const int tiles = (cin / 64) * (cout / 64);
  const int n_items = N * (H / wgtc::R) * (W / wgtc::BW);
  const int S = wgtc::slabs(n_items, tiles);
  return (size_t)S * ((size_t)9 * cin * cout + cout) * sizeof(float);
}

// -> partial sums in the layout of conv_bwd.cu's reduction; *S_out slabs
int wgrad_tc_launch(const float* x, int x_ld, const float* g, int g_ld, int N, int H, int W, int cin, int cout, float* workspace,
                    float** part_bias_out, int* S_out, cudaStream_t st) {
  using namespace wgtc;
  static int smem_done[kMaxDevices] = {0};
  if (int rc = ensure_dynamic_smem(wgrad_tc_kernel, SMEM, smem_done, "wgrad_tc")) return rc;
  EncodeFn encode = encoder();
  if (encode == nullptr) {
    set_error("wgrad_tc: cuTensorMapEncodeTiled is not available from this driver");
    return TDVC_ECUDA;
  }
  Maps maps;
  const struct { const float* p; int ld, c, box_w, box_h; CUtensorMap* m; } t[2] = {{x, x_ld, cin, XW, XROWS, &maps.x}, {g, g_ld, cout, BW, R, &maps.g}};
  for (int i = 0; i < 2; ++i) {
    // (32 channels, x, channel group of 32, y, image); coordinates outside (the image border = the zero padding) are zero-filled
    const cuuint64_t ld = (cuuint64_t)t[i].ld;
    const cuuint64_t dims[5] = {32, (cuuint64_t)W, (cuuint64_t)(t[i].c / 32), (cuuint64_t)H, (cuuint64_t)N};
    const cuuint64_t strides[4] = {ld * 4, 128, (cuuint64_t)W * ld * 4, (cuuint64_t)H * W * ld * 4};
    const cuuint32_t box[5] = {32, (cuuint32_t)t[i].box_w, 2, (cuuint32_t)t[i].box_h, 1}, estr[5] = {1, 1, 1, 1, 1};
    const CUresult r = encode(t[i].m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(t[i].p), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("wgrad_tc: cuTensorMapEncodeTiled failed for operand %d (%d)", i, (int)r);
      return TDVC_ECUDA;
    }
  }
  Args a;
  a.N = N; a.H = H; a.W = W; a.cin = cin; a.cout = cout;
  a.co_tiles = cout / 64;
  a.tiles_x = W / BW; a.tiles_y = H / R;
  a.n_items = N * a.tiles_x * a.tiles_y;
  const int tiles = (cin / 64) * a.co_tiles;
  a.S = slabs(a.n_items, tiles);
  a.part = workspace;
  a.part_bias = workspace + (size_t)a.S * 9 * cin * cout;
  *part_bias_out = a.part_bias;
  *S_out = a.S;
  wgrad_tc_kernel<<<dim3(tiles, a.S), THREADS, SMEM, st>>>(a, maps);
  TDVC_CHECK_LAUNCH("wgrad_tc");
  return TDVC_OK;
}

}  // namespace tdvc
