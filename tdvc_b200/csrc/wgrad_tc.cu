// Weight gradient of the stride-1 convolutions on the tensor cores (the training step; the reference
// takes this gradient from cuDNN through autograd, tools/train.py:125-159).  One TF32 product per MAC - the arithmetic of
// `enabled_amp=True` (the reference's shipped cfg/train.yaml) - or three (fp32-class: operands split into hi + lo, the lo parts
// written next to the staged patches by the otherwise idle read-out warps; conv_bwd.cu keeps the shapes this file does not take).
//
//   d W[tap][ci][co] = sum over pixels p of x[p + tap][ci] * g[p][co]
//
// is a GEMM whose reduction index is the PIXEL, and both operands live in channels-last tensors: a row of the operand (one
// pixel) holds the M (ci) or N (co) index contiguously.  That IS tcgen05's MN-major operand layout, so nothing is transposed
// and no thread touches an operand: tensor-map bulk copies (4-D maps over (channel, x, y, image), boxes of 32 channels x one
// row segment, 128-byte swizzle with 32-byte atoms; channels past the tensor and pixels past the border are zero-filled by the
// hardware, which is the convolution's padding) drop a patch of x with its halo and the matching patch of g into shared
// memory, and every MMA reads 8 pixels x (128 | NB) channels of them through MN-major descriptors.  (The first warp-level
// kernel staged 32-pixel chunks and issued m16n8k8 MMAs from scalar fragment loads: 0.85 ms for 64->64 3x3 at 8x256x256,
// 4.5 ms for every 7x7 SPyNet layer whatever its width; a K-major tcgen05 version had to transpose with 4-byte cp.async and
// was slower still - conv_bwd.cu.  This one: 0.11 ms for the 3x3 layer.)
//
// Work item = R output rows x BW = 32 output columns of one image.  Shared-memory order of the x patch: [row][group of 32
// channels][column][32 channels] - consecutive (row, group) blocks are equally spaced, so ONE descriptor with that spacing as
// its leading-dimension offset spans the G channel groups of kernel row ky AND of the kernel rows behind it: an M = 128 operand
// from a 64-channel tile (2 kernel rows) or a <= 32-channel tile (4 kernel rows); M = 64 would run the tensor pipe at half
// rate.  Per kernel column the kernel rows are covered by NG such operands (3x3, 64 channels: [ky 0 | ky 1] and [ky 1 (dropped)
// | ky 2]), each with its own NB TMEM columns; kernels with more than 512 columns of accumulators split their kernel columns
// over CTAs.  A tap's shift is a shift of the operand's start address by whole pixels (128-byte rows inside the swizzle
// pattern, whose phase follows the absolute address).
// CTA = (ci tile, co tile, kernel-column group) x slab s: items s, s + S, ...; accumulators stay in TMEM over the whole slab;
// partial sums go to the workspace [S][tap][cin][cout] and conv_bwd.cu's fixed-order reduction adds them (deterministic).
// Warp 0: bulk copies, warp 1: MMA issue, warps 2-5: column sums of g for the bias gradient from the staged patches, then the
// TMEM read-out.
#include <cuda.h>
#include <cstdlib>

#include "tc_common.cuh"

namespace tdvc {
namespace wgtc {

using namespace tc;

constexpr int BW = 32;
constexpr int THREADS = 192;
constexpr int TMEM_COLS = 512;

template <int KS, int G, int NB, int R, int PR>
struct Cfg {
  static constexpr int PAD = KS / 2;
  static constexpr int RPM = 4 / G;                                   // kernel rows one M = 128 operand spans
  static constexpr int NG = (KS + RPM - 1) / RPM;                     // operands (accumulators) per kernel column
  static constexpr int XROWS = R + (KS > RPM ? KS : RPM) - 1;
  static constexpr int XW = (BW + KS - 1 + 3) / 4 * 4;                // whole 512-byte swizzle atoms per (row, group) block
  static constexpr int NBG = NB / 32;
  static constexpr int BLK_X = XW * 128, BLK_G = BW * 128;
  static constexpr int X_BYTES = XROWS * G * BLK_X, G_BYTES = R * NBG * BLK_G;
  static constexpr int TILE = X_BYTES + G_BYTES;                      // what the bulk copies bring
  static constexpr int STAGE = (PR == 3 ? 2 : 1) * TILE;              // PR == 3: + the lo parts of both patches, same layout
  static constexpr int NSTAGE = (PR == 3 && KS == 7 && G == 2) ? 1 : 2;   // (two such stages of the 64-channel 7x7 patch do not fit)
  static constexpr int KXC = KS < TMEM_COLS / (NG * NB) ? KS : TMEM_COLS / (NG * NB);   // kernel columns per CTA
  static constexpr int NKX = (KS + KXC - 1) / KXC;
  static constexpr int SMEM = NSTAGE * STAGE + 1024 /* alignment slack */ + 128 /* barriers, TMEM slot */ + 512 /* bias exchange */;
  static_assert(STAGE % 512 == 0 && KXC >= 1 && SMEM <= 227 * 1024, "wgrad_tc configuration");
  // first kernel row of operand gi: 0, RPM, 2 RPM, ..., the last one pulled back so that it ends on the last kernel row
  __host__ __device__ static constexpr int start(int gi) {
    const int last = KS > RPM ? KS - RPM : 0;
    return gi * RPM < last ? gi * RPM : last;
  }
};

struct Maps {
  CUtensorMap x, g;
};

struct Args {
  int N, H, W, cin, cout;
  int S, co_tiles, n_items, tiles_x, tiles_y;
  float* part;       // [S][k*k][cin][cout]
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
// kind::tf32: D = f32, A = B = tf32 (format 2), both MN-major (bits 15, 16), M = 128, N = n
__host__ __device__ constexpr uint32_t idesc_tf32(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tma4(uint32_t dst, const CUtensorMap* m, int c0, int c1, int c2, int c3, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
      : "memory");
}

template <int KS, int G, int NB, int R, int PR>
__global__ void __launch_bounds__(THREADS, 1) wgrad_tc_kernel(const Args a, const __grid_constant__ Maps maps) {
  using C = Cfg<KS, G, NB, R, PR>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* const gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bars = base + C::NSTAGE * C::STAGE;            // full[2], empty[2], acc, lo[2]
  uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(gen + C::NSTAGE * C::STAGE + 64);
  float* const bias_x = reinterpret_cast<float*>(gen + C::NSTAGE * C::STAGE + 128);   // [128]
  auto full = [&](int s) { return bars + 8u * s; };
  auto empty = [&](int s) { return bars + 16u + 8u * s; };
  const uint32_t acc_bar = bars + 32u;
  auto lo_bar = [&](int s) { return bars + 40u + 8u * s; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int s = blockIdx.y;
  const int kxs = blockIdx.x % C::NKX, tile = blockIdx.x / C::NKX;
  const int co_t = tile % a.co_tiles, ci_t = tile / a.co_tiles;
  const int kx0 = kxs * C::KXC;

  if (threadIdx.x == 0) {
    for (int i = 0; i < C::NSTAGE; ++i) { mbar_init(full(i), 1); mbar_init(empty(i), 5); mbar_init(lo_bar(i), 4); }
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
    // ------------------------------------------------------------------ bulk copies: the (row, group) blocks of one x patch and one g patch
    if (elect_one()) {
      int st = 0, ph = 1;
      for (int item = s; item < a.n_items; item += a.S) {
        int r = item;
        const int tx = r % a.tiles_x; r /= a.tiles_x;
        const int ty = r % a.tiles_y;
        const int n = r / a.tiles_y;
        mbar_wait(empty(st), ph);
        mbar_expect_tx(full(st), C::TILE);
        const uint32_t dst = base + st * C::STAGE;
#pragma unroll 1
        for (int row = 0; row < C::XROWS; ++row)
#pragma unroll
          for (int grp = 0; grp < G; ++grp)
            tma4(dst + (row * G + grp) * C::BLK_X, &maps.x, (ci_t * G + grp) * 32, tx * BW - C::PAD, ty * R - C::PAD + row, n, full(st));
#pragma unroll 1
        for (int row = 0; row < R; ++row)
#pragma unroll
          for (int grp = 0; grp < C::NBG; ++grp)
            tma4(dst + C::X_BYTES + (row * C::NBG + grp) * C::BLK_G, &maps.g, co_t * NB + grp * 32, tx * BW, ty * R + row, n, full(st));
        if (++st == C::NSTAGE) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issue: R x 4 K-steps x kernel columns x NG operands
    constexpr uint32_t idesc = idesc_tf32(NB);
    int st = 0, ph = 0;
    uint32_t first = 1;
    for (int item = s; item < a.n_items; item += a.S) {
      mbar_wait(full(st), ph);
      tc_fence_after();
      if (elect_one()) {
        uint32_t acc = first ^ 1u;
        // pass 0: x * g as staged (the tensor core reads the TF32 truncation of an fp32 word: the hi parts);
        // PR == 3, passes 1, 2: x_lo * g and x * g_lo from the converted copies - hi hi + lo hi + hi lo carries ~21 bits
#pragma unroll 1
        for (int pass = 0; pass < PR; ++pass) {
          if (pass == 1) {
            mbar_wait(lo_bar(st), ph);
            tc_fence_after();
          }
          const uint32_t xs = base + st * C::STAGE + (pass == 1 ? C::TILE : 0);
          const uint32_t gs = base + st * C::STAGE + C::X_BYTES + (pass == 2 ? C::TILE : 0);
#pragma unroll 1
          for (int r = 0; r < R; ++r) {
#pragma unroll
            for (int c0 = 0; c0 < BW; c0 += 8) {
              const uint64_t bdesc = mn_desc(gs + r * C::NBG * C::BLK_G + c0 * 128, C::BLK_G);
#pragma unroll
              for (int kxi = 0; kxi < C::KXC; ++kxi) {
                if (C::NKX > 1 && kx0 + kxi >= KS) break;
#pragma unroll
                for (int gi = 0; gi < C::NG; ++gi) {
                  const uint64_t adesc = mn_desc(xs + (r + C::start(gi)) * G * C::BLK_X + (c0 + kx0 + kxi) * 128, C::BLK_X);
                  mma_tf32(tmem_base + (kxi * C::NG + gi) * NB, adesc, bdesc, idesc, acc);
                }
              }
              acc = 1;
            }
          }
        }
        tc_commit(empty(st));
      }
      first = 0;
      __syncwarp();
      if (++st == C::NSTAGE) { st = 0; ph ^= 1; }
    }
    if (elect_one()) tc_commit(acc_bar);
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ bias sums from the staged g patches, then the read-out
    constexpr int NPART = 128 / NB;
    const int et = threadIdx.x - 64;                 // 0..127
    const int co = et % NB, part = et / NB;          // rows part, part + NPART, ... of every patch
    const bool do_bias = a.part_bias != nullptr && ci_t == 0 && kxs == 0;
    float bsum = 0.f;
    int st = 0, ph = 0;
    for (int item = s; item < a.n_items; item += a.S) {
      mbar_wait(full(st), ph);
      if (PR == 3) {
        // lo = v - tf32_truncation(v), exact in fp32, for both patches; the generic-proxy stores are fenced for the tensor core
        const float4* src = reinterpret_cast<const float4*>(gen + st * C::STAGE);
        float4* dst = reinterpret_cast<float4*>(gen + st * C::STAGE + C::TILE);
#pragma unroll 4
        for (int i = et; i < C::TILE / 16; i += 128) {
          const float4 v = src[i];
          float4 l;
          l.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
          l.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
          l.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
          l.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
          dst[i] = l;
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(lo_bar(st));
      }
      if (do_bias) {
        const uint8_t* gs = gen + st * C::STAGE + C::X_BYTES + (co >> 5) * C::BLK_G + (co & 7) * 4;
        const int ch = (co & 31) >> 3;
        for (int rr = part; rr < R; rr += NPART) {
          const uint8_t* row = gs + rr * C::NBG * C::BLK_G;
#pragma unroll 8
          for (int px = 0; px < BW; ++px) bsum += *reinterpret_cast<const float*>(row + px * 128 + ((ch ^ (px & 3)) << 5));
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty(st));
      if (++st == C::NSTAGE) { st = 0; ph ^= 1; }
    }
    if (do_bias) {
      bias_x[et] = bsum;
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (part == 0 && co_t * NB + co < a.cout) {
        float v = 0.f;
#pragma unroll
        for (int q = 0; q < NPART; ++q) v += bias_x[q * NB + co];
        a.part_bias[(int64_t)s * a.cout + co_t * NB + co] = v;
      }
    }
    mbar_wait(acc_bar, 0);
    tc_fence_after();
    const int q = warp & 3;                          // TMEM lane quarter this warp may read
    const int L = q * 32 + lane, cil = L % (32 * G), sub = L / (32 * G);   // `sub`: kernel row inside the operand (per warp)
    const int ci = ci_t * 32 * G + cil;
    const bool vec = (a.cout & 3) == 0 && (co_t + 1) * NB <= a.cout;
#pragma unroll 1
    for (int kxi = 0; kxi < C::KXC; ++kxi) {
      const int kx = kx0 + kxi;
      if (kx >= KS) break;
#pragma unroll 1
      for (int gi = 0; gi < C::NG; ++gi) {
        const int ky = C::start(gi) + sub;
        const int lo = gi > 0 ? C::start(gi - 1) + C::RPM : 0;      // kernel rows below `lo` belong to the operand before
        if (ky < lo || ky >= KS) continue;                          // warp-uniform
        float* dst = a.part + (((int64_t)s * KS * KS + ky * KS + kx) * a.cin + ci) * a.cout + co_t * NB;
#pragma unroll
        for (int h = 0; h < C::NBG; ++h) {
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (kxi * C::NG + gi) * NB + h * 32, v);
          tmem_ld_wait();
          if (ci < a.cin) {
            if (vec) {
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(dst + h * 32 + j) =
                    make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (co_t * NB + h * 32 + j < a.cout) dst[h * 32 + j] = __uint_as_float(v[j]);
            }
          }
        }
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

// the launch geometry of a shape: which instantiation (0 = none), its tiles and slabs
struct Plan {
  int id, G, NB, R, nkx, ci_tiles, co_tiles, tiles_x, tiles_y, n_items, S;
};
static Plan plan(int N, int H, int W, int cin, int cout, int k, int products) {
  Plan p = {};
  p.G = cin <= 32 ? 1 : 2;
  p.NB = cout <= 32 ? 32 : 64;
  // rows per item: the three-product mode keeps a second copy of both patches (their lo parts) in shared memory
  p.R = products == 3 ? 2 : ((k == 7 && p.G == 2) ? 2 : 4);
  if (k == 3) { p.id = 1; p.nkx = 1; }
  else if (k == 7 && !(p.G == 2 && p.NB == 64)) { p.id = 2; p.nkx = (p.G == 1 && p.NB == 32) ? 1 : 2; }
  else if (k == 1 && p.G == 2 && p.NB == 64) { p.id = 3; p.nkx = 1; }
  else return p;
  p.ci_tiles = cdiv(cin, 32 * p.G);
  p.co_tiles = cdiv(cout, p.NB);
  p.tiles_x = cdiv(W, BW);
  p.tiles_y = cdiv(H, p.R);
  p.n_items = N * p.tiles_x * p.tiles_y;
  const int ctas = p.ci_tiles * p.co_tiles * p.nkx;
  p.S = kNumSMs / ctas;
  if (p.S < 1) p.S = 1;
  if (p.S > p.n_items) p.S = p.n_items;
  return p;
}

template <int KS, int G, int NB, int R, int PR>
static int launch(const Plan& p, const float* x, int x_ld, const float* g, int g_ld, int N, int H, int W, int cin, int cout,
                  float* workspace, float** part_bias_out, cudaStream_t st) {
  using C = Cfg<KS, G, NB, R, PR>;
  TDVC_REQUIRE(p.nkx == C::NKX && p.R == R, "wgrad_tc: plan and kernel configuration disagree (k=%d G=%d NB=%d)", KS, G, NB);
  static int smem_done[kMaxDevices] = {0};
  if (int rc = ensure_dynamic_smem(wgrad_tc_kernel<KS, G, NB, R, PR>, C::SMEM, smem_done, "wgrad_tc")) return rc;
  EncodeFn encode = encoder();
  if (encode == nullptr) {
    set_error("wgrad_tc: cuTensorMapEncodeTiled is not available from this driver");
    return TDVC_ECUDA;
  }
  Maps maps;
  const struct { const float* p; int ld, c, box_w; CUtensorMap* m; } t[2] = {{x, x_ld, cin, C::XW, &maps.x}, {g, g_ld, cout, BW, &maps.g}};
  for (int i = 0; i < 2; ++i) {
    // (channel, x, y, image); box = 32 channels x one row segment.  Coordinates outside are zero-filled
    const cuuint64_t ld = (cuuint64_t)t[i].ld;
    const cuuint64_t dims[4] = {(cuuint64_t)t[i].c, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    const cuuint64_t strides[3] = {ld * 4, (cuuint64_t)W * ld * 4, (cuuint64_t)H * W * ld * 4};
    const cuuint32_t box[4] = {32, (cuuint32_t)t[i].box_w, 1, 1}, estr[4] = {1, 1, 1, 1};
    const CUresult r = encode(t[i].m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(t[i].p), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("wgrad_tc: cuTensorMapEncodeTiled failed for operand %d (%d)", i, (int)r);
      return TDVC_ECUDA;
    }
  }
  Args a;
  a.N = N; a.H = H; a.W = W; a.cin = cin; a.cout = cout;
  a.co_tiles = p.co_tiles;
  a.tiles_x = p.tiles_x; a.tiles_y = p.tiles_y;
  a.n_items = p.n_items;
  a.S = p.S;
  a.part = workspace;
  a.part_bias = workspace + (size_t)a.S * KS * KS * cin * cout;
  *part_bias_out = a.part_bias;
  wgrad_tc_kernel<KS, G, NB, R, PR><<<dim3(p.ci_tiles * p.co_tiles * C::NKX, a.S), THREADS, C::SMEM, st>>>(a, maps);
  TDVC_CHECK_LAUNCH("wgrad_tc");
  return TDVC_OK;
}

}  // namespace wgtc

// shapes the tensor-core weight gradient takes: stride 1, "same" padding, 1x1 / 3x3 / 7x7 (see wgtc::plan); products 1 | 3
bool wgrad_tc_eligible(const float* x, int x_ld, const float* g, int g_ld, int N, int H, int W, int cin, int cout, int k, int stride,
                       int pad, int in_square, int products) {
  static const bool off = getenv("TDVC_B200_WGRAD_SIMT") != nullptr;   // developer A/B switch: warp-level MMA kernel everywhere
  if (off || stride != 1 || pad != k / 2 || in_square || W < 16 || H < 4) return false;
  if ((reinterpret_cast<uintptr_t>(x) & 15) != 0 || (reinterpret_cast<uintptr_t>(g) & 15) != 0 || x_ld % 4 != 0 || g_ld % 4 != 0) return false;
  return wgtc::plan(N, H, W, cin, cout, k, products).id != 0;
}

// the larger of the two modes' needs (the query does not know the mode)
size_t wgrad_tc_workspace_bytes(int N, int H, int W, int cin, int cout, int k) {
  size_t need = 0;
  for (int products = 1; products <= 3; products += 2) {
    const wgtc::Plan p = wgtc::plan(N, H, W, cin, cout, k, products);
    if (p.id == 0) continue;
    const size_t b = (size_t)p.S * ((size_t)k * k * cin * cout + cout) * sizeof(float);
    if (b > need) need = b;
  }
  return need;
}

// -> partial sums in the layout of conv_bwd.cu's reduction; *S_out slabs
int wgrad_tc_launch(const float* x, int x_ld, const float* g, int g_ld, int N, int H, int W, int cin, int cout, int k, int products,
                    float* workspace, float** part_bias_out, int* S_out, cudaStream_t st) {
  using namespace wgtc;
  const Plan p = plan(N, H, W, cin, cout, k, products);
  *S_out = p.S;
#define TDVC_WG(KS, G_, NB_, R_, PR_)                                   \
  if (k == KS && p.G == G_ && p.NB == NB_ && products == PR_)          \
    return launch<KS, G_, NB_, R_, PR_>(p, x, x_ld, g, g_ld, N, H, W, cin, cout, workspace, part_bias_out, st);
  TDVC_WG(3, 2, 64, 4, 1)
  TDVC_WG(3, 1, 64, 4, 1)
  TDVC_WG(3, 2, 32, 4, 1)
  TDVC_WG(3, 1, 32, 4, 1)
  TDVC_WG(7, 1, 32, 4, 1)
  TDVC_WG(7, 1, 64, 4, 1)
  TDVC_WG(7, 2, 32, 2, 1)
  TDVC_WG(1, 2, 64, 4, 1)
  TDVC_WG(3, 2, 64, 2, 3)
  TDVC_WG(3, 1, 64, 2, 3)
  TDVC_WG(3, 2, 32, 2, 3)
  TDVC_WG(3, 1, 32, 2, 3)
  TDVC_WG(7, 1, 32, 2, 3)
  TDVC_WG(7, 1, 64, 2, 3)
  TDVC_WG(7, 2, 32, 2, 3)
  TDVC_WG(1, 2, 64, 2, 3)
#undef TDVC_WG
  set_error("wgrad_tc: no kernel for k=%d cin=%d cout=%d products=%d", k, cin, cout, products);
  return TDVC_EINVAL;
}

}  // namespace tdvc
