// DCNv2 forward on the tensor cores: bilinear gather producers + tcgen05 contraction, for the configuration the
// P-frame graph uses (3x3, stride 1, pad 1, dilation 1, 8 channels per deformable group, O <= 64).
//
// Reference: main/utils/dcnv2/src/cuda/dcn_v2_im2col_cuda.cu:25-54,125-195 (bilinear rule, one thread per
// (c, h, w), 9 taps each, result written to a (C*9) x (H*W) `columns` tensor) followed by
// main/utils/dcnv2/src/cuda/dcn_v2_cuda.cu:90-92 (at::matmul(weight_flat, columns) + bias); the fp16 rounding of the
// result is main/utils/dcnv2/dcn_v2_amp.py:67-69 and the LeakyReLU on it is main/model/pnet.py:180.
//
// Here the modulated columns of a 16x16-pixel tile never leave the SM: producers write them straight into the
// tcgen05 K-major shared-memory operand layout and the 576 -> O contraction runs as 3xFP16-split MMAs with fp32
// accumulators in TMEM (same numerics as conv_tc.cu).
//
// Data layouts chosen so that a warp-wide access touches few 128-byte lines (the gather is L1-wavefront bound):
//   input   "group planar": [(n*dg + g)][H][W][8 channels]  - a lane's bilinear corner is ONE 32-byte sector and the
//           32 lanes of a warp (consecutive x) read neighbouring sectors (tdvc_nhwc_to_group_planar makes it);
//   offsets / mask planar: plane (g*18 + 2*tap [+1]) resp. (g*9 + tap) of H*W floats - exactly the reference's NCHW
//           offset / mask tensors (dcn_v2.h:9-46), and what the offset/mask convolution writes with out_planar = 1.
//
// Work decomposition (persistent, one CTA per SM, 768 threads):
//   item = 16x8 output pixels = one MMA tile (M = 128: 8 rows x 16 px, row m = (y%8)*16 + x); kTiles tiles per item;
//   K is ordered k' = (g*9 + tap)*8 + c (one 16-byte core-matrix row per (pixel, g, tap)), padded per group from 72
//   to 80 (5 MMA k-steps); one pipeline stage = one deformable group.
//   warps 4-21  producers (18 warps, 80 registers: the gather is latency-bound, 12 warps at 94 registers were 8 % slower): lane = pixel, task = (32 pixels, tap): 3 coalesced parameter loads, 4 x 256-bit corner
//               loads, bilinear blend * mask for the group's 8 channels, fp16 hi/lo split, two 16-byte smem stores.
//   warp 23     streams the group's weight block [128 rows = w_hi | w_lo*2^12][80] (20 KB) by 1-D bulk TMA, 3-deep ring.
//   warp 22     one elected thread issues per (stage, tile, k-step): D[:,0:128] += A_hi*[W_hi|W_lo]^T, D[:,0:64] += A_lo*W_hi^T.
//   warps 0-3   epilogue: tcgen05.ld, hi+lo columns, bias, optional fp16 rounding + LeakyReLU as torch does on a
//               Half tensor, 256-bit stores (NHWC fp32); overlaps the next item (2 accumulator stages in TMEM).
#include "tc_common.cuh"

namespace tdvc {
namespace dcn {

using namespace tc;

constexpr int kEpiWarps = 4, kProdWarps = 18;
constexpr int kMmaWarp = kEpiWarps + kProdWarps, kLoadWarp = kMmaWarp + 1;
constexpr int kThreads = (kEpiWarps + kProdWarps + 2) * 32;  // 768
constexpr int kProdThreads = kProdWarps * 32;
constexpr int kTile = 16;                     // item width (pixels)
constexpr int kTiles = 1;                    // MMA tiles (8 rows x 16 px, M = 128) per item.  One tile keeps the CTA at 140 KB
                                             // of shared memory, so ~85 KB stay L1: the bilinear gather of a group (a ~26x18-pixel
                                             // window x 32 B, re-read by 9 taps x 4 corners) then hits L1 instead of L2
constexpr int kTileH = 8 * kTiles;
constexpr int NT = 64;                       // output channels (padded)
constexpr int KCH = 10;                      // 16-byte K chunks per group (9 taps + 1 zero pad)
constexpr int CHUNK_BYTES = 128 * 16;        // one K chunk of one MMA tile: 128 rows x 16 B
constexpr int A_PLANE = KCH * CHUNK_BYTES;   // hi (or lo) plane of one tile: 20 KB
constexpr int A_STAGE = kTiles * 2 * A_PLANE; // [tile][hi|lo]
constexpr int B_BLOCK = 2 * NT * KCH * 8 * 2;  // [128 rows][80] fp16: 20 KB
constexpr int NA = 2, NB = 3;
constexpr int SMEM = NA * A_STAGE + NB * B_BLOCK + 256;
constexpr int TMEM_COLS = 2 * kTiles * 128;   // 2 stages x tiles x 128 columns
constexpr int KSTEPS = KCH / 2;
static_assert(SMEM <= 227 * 1024, "shared memory budget");

__device__ __forceinline__ float sigmoid_exact(float x) { return 1.f / (1.f + expf(-x)); }

__global__ void __launch_bounds__(kThreads, 1) dcn_tc_kernel(const TdvcDcnParams p, int tiles_x, int tiles_y, int n_items) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* a_buf = smem;
  uint8_t* b_buf = smem + NA * A_STAGE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_buf + NB * B_BLOCK);
  constexpr int A_FULL = 0, A_EMPTY = A_FULL + NA, B_FULL = A_EMPTY + NA, B_EMPTY = B_FULL + NB, ACC_FULL = B_EMPTY + NB,
                ACC_EMPTY = ACC_FULL + 2, NBARS = ACC_EMPTY + 2;
  static_assert(NBARS * 8 + 8 <= 256, "barrier area");
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NBARS);
  const uint32_t bar0 = smem_u32(bars);
  auto bar = [&](int i) { return bar0 + 8u * i; };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = p.H, W = p.W, dg = p.dg;
  const int64_t HW = (int64_t)H * W;

  if (threadIdx.x == 0) {
    for (int i = 0; i < NA; ++i) { mbar_init(bar(A_FULL + i), kProdThreads); mbar_init(bar(A_EMPTY + i), 1); }
    for (int i = 0; i < NB; ++i) { mbar_init(bar(B_FULL + i), 1); mbar_init(bar(B_EMPTY + i), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(bar(ACC_FULL + i), 1); mbar_init(bar(ACC_EMPTY + i), kEpiWarps * 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // the pad chunk (k = 72..79 of every group) of every A plane is zero for the whole kernel
  for (int i = threadIdx.x; i < NA * kTiles * 2 * (CHUNK_BYTES / 16); i += kThreads) {
    const int plane = i / (CHUNK_BYTES / 16), r = i % (CHUNK_BYTES / 16);
    *reinterpret_cast<uint4*>(a_buf + plane * A_PLANE + (KCH - 1) * CHUNK_BYTES + r * 16) = make_uint4(0, 0, 0, 0);
  }
  fence_async_smem();
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  auto decode = [&](int item, int& n, int& y0, int& x0) {
    x0 = (item % tiles_x) * kTile;
    item /= tiles_x;
    y0 = (item % tiles_y) * kTileH;
    n = item / tiles_y;
  };

  if (warp < kEpiWarps) {
    // ===================================================================== epilogue
    const int quad = warp;
    const int m = quad * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
    const bool out_vec = (p.out_ld & 7) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 31) == 0;
    int acc_it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++acc_it) {
      int n, y0, x0;
      decode(item, n, y0, x0);
      const int sa = acc_it & 1;
      mbar_wait(bar(ACC_FULL + sa), (acc_it >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int t = 0; t < kTiles; ++t) {
        const int y = y0 + 8 * t + (m >> 4), x = x0 + (m & 15);
        const bool valid = y < H && x < W;
        const uint32_t tcol = lane_addr + (uint32_t)((sa * kTiles + t) * 2 * NT);
#pragma unroll 1
        for (int c0 = 0; c0 < NT; c0 += 32) {
          uint32_t ra[32], rb[32];
          tmem_ld32(tcol + c0, ra);
          tmem_ld32(tcol + NT + c0, rb);
          tmem_ld_wait();
          if (!valid || c0 >= p.O) continue;
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float o = fmaf(__uint_as_float(rb[j]), kLoUnscale, __uint_as_float(ra[j]));
            if (p.bias && c0 + j < p.O) o += __ldg(p.bias + c0 + j);
            if (p.round_fp16) {
              o = __half2float(__float2half_rn(o));
              if (p.act == TDVC_ACT_LRELU) {
                if (o < 0.f) o = __half2float(__float2half_rn(o * p.slope));  // torch Half leaky_relu
              } else {
                o = apply_act(o, p.act, p.slope);
              }
            } else {
              o = apply_act(o, p.act, p.slope);
            }
            v[j] = o;
          }
          float* op = p.out + (((int64_t)n * H + y) * W + x) * p.out_ld + c0;
          if (c0 + 32 <= p.O && out_vec) {
#pragma unroll
            for (int j = 0; j < 4; ++j) stg256(op + 8 * j, v + 8 * j);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (c0 + j < p.O) op[j] = v[j];
          }
        }
      }
      tc_fence_before();
      mbar_arrive(bar(ACC_EMPTY + sa));
    }
  } else if (warp < kEpiWarps + kProdWarps) {
    // ===================================================================== gather producers
    const int pw = warp - kEpiWarps;
    int a_it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      int n, y0, x0;
      decode(item, n, y0, x0);
      const float* off_n = p.offset + (int64_t)n * p.off_ld * HW;
      const float* msk_n = p.mask + (int64_t)n * p.mask_ld * HW;
      for (int g = 0; g < dg; ++g, ++a_it) {
        const int st = a_it % NA;
        mbar_wait(bar(A_EMPTY + st), ((a_it / NA) & 1) ^ 1);
        uint8_t* stage = a_buf + st * A_STAGE;
        const float* in_g = p.input_gp + ((int64_t)n * dg + g) * HW * 8;
        // RP row pairs x 9 taps warp tasks per stage, TPW per producer warp.  The three parameters (dy, dx, mask) of all
        // its tasks are loaded first (coalesced loads in flight), so each task then waits for ONE memory latency
        // (its four corner sectors) instead of two.
        constexpr int RP = 4 * kTiles;                 // row pairs of the item
        static_assert((RP * 9) % kProdWarps == 0, "tasks per producer warp");
        constexpr int TPW = RP * 9 / kProdWarps;
        float pdy[TPW], pdx[TPW], pmk[TPW];
#pragma unroll
        for (int j = 0; j < TPW; ++j) {
          const int q = pw + j * kProdWarps;
          const int rp = q % RP, tap = q / RP;
          const int y = y0 + 2 * rp + (lane >> 4), x = x0 + (lane & 15);
          pdy[j] = pdx[j] = pmk[j] = 0.f;
          if (y < H && x < W) {
            const int64_t pix = (int64_t)y * W + x;
            pdy[j] = __ldg(off_n + (int64_t)(g * 18 + 2 * tap) * HW + pix);
            pdx[j] = __ldg(off_n + (int64_t)(g * 18 + 2 * tap + 1) * HW + pix);
            pmk[j] = __ldg(msk_n + (int64_t)(g * 9 + tap) * HW + pix);
          }
        }
#pragma unroll
        for (int j = 0; j < TPW; ++j) {
          const int q = pw + j * kProdWarps;
          const int rp = q % RP, tap = q / RP;
          const int ly = 2 * rp + (lane >> 4), lx = lane & 15;
          const int y = y0 + ly, x = x0 + lx;
          const int t = ly >> 3, m = (ly & 7) * 16 + lx;
          float val[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) val[c] = 0.f;
          if (y < H && x < W) {
            const float dy = pdy[j], dx = pdx[j];
            float mk = pmk[j];
            if (p.mask_is_logit) mk = sigmoid_exact(mk);
            const int ki = tap / 3, kj = tap - ki * 3;
            const float h_im = (float)(y - 1 + ki) + dy;
            const float w_im = (float)(x - 1 + kj) + dx;
            if (h_im > -1.f && w_im > -1.f && h_im < (float)H && w_im < (float)W) {
              const float hf = floorf(h_im), wf = floorf(w_im);
              const int h_low = (int)hf, w_low = (int)wf;
              const int h_high = h_low + 1, w_high = w_low + 1;
              const float lh = h_im - hf, lw = w_im - wf;
              const float hh = 1.f - lh, hw = 1.f - lw;
              const float w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
              float c1[8], c2[8], c3[8], c4[8];
#pragma unroll
              for (int c = 0; c < 8; ++c) c1[c] = c2[c] = c3[c] = c4[c] = 0.f;
              const bool t_ok = h_low >= 0, b_ok = h_high <= H - 1, l_ok = w_low >= 0, r_ok = w_high <= W - 1;
              const float* base = in_g + ((int64_t)h_low * W + w_low) * 8;
              if (t_ok && l_ok) ldg256(base, c1);
              if (t_ok && r_ok) ldg256(base + 8, c2);
              if (b_ok && l_ok) ldg256(base + (int64_t)W * 8, c3);
              if (b_ok && r_ok) ldg256(base + (int64_t)W * 8 + 8, c4);
              // same evaluation order as the reference: w1*v1 + w2*v2 + w3*v3 + w4*v4, then * mask
#pragma unroll
              for (int c = 0; c < 8; ++c) val[c] = (w1 * c1[c] + w2 * c2[c] + w3 * c3[c] + w4 * c4[c]) * mk;
            }
          }
          uint2 h0, l0, h1, l1;
          split4(make_float4(val[0], val[1], val[2], val[3]), h0, l0);
          split4(make_float4(val[4], val[5], val[6], val[7]), h1, l1);
          uint8_t* dst = stage + (t * 2) * A_PLANE + tap * CHUNK_BYTES + m * 16;
          *reinterpret_cast<uint4*>(dst) = make_uint4(h0.x, h0.y, h1.x, h1.y);
          *reinterpret_cast<uint4*>(dst + A_PLANE) = make_uint4(l0.x, l0.y, l1.x, l1.y);
        }
        fence_async_smem();
        mbar_arrive(bar(A_FULL + st));
      }
    }
  } else if (warp == kMmaWarp) {
    // ===================================================================== MMA issuer
    if (elect_one()) {
      constexpr uint32_t IDESC_2N = instr_desc(2 * NT), IDESC_N = instr_desc(NT);
      const uint32_t a0 = smem_u32(a_buf), b0 = smem_u32(b_buf);
      int a_it = 0, acc_it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++acc_it) {
        const int sa = acc_it & 1;
        mbar_wait(bar(ACC_EMPTY + sa), ((acc_it >> 1) & 1) ^ 1);
        tc_fence_after();
        for (int g = 0; g < dg; ++g, ++a_it) {
          const int sA = a_it % NA, sB = a_it % NB;
          mbar_wait(bar(B_FULL + sB), (a_it / NB) & 1);
          mbar_wait(bar(A_FULL + sA), (a_it / NA) & 1);
          tc_fence_after();
          const uint32_t bblk = b0 + sB * B_BLOCK;
#pragma unroll
          for (int t = 0; t < kTiles; ++t) {
            const uint32_t d = tmem_base + (uint32_t)((sa * kTiles + t) * 2 * NT);
            const uint32_t a_hi = a0 + sA * A_STAGE + (t * 2) * A_PLANE, a_lo = a_hi + A_PLANE;
#pragma unroll
            for (int s = 0; s < KSTEPS; ++s) {
              const uint64_t bd = smem_desc(bblk + s * 2 * 128, 128, KCH * 128);
              const uint64_t adh = smem_desc(a_hi + s * 2 * CHUNK_BYTES, CHUNK_BYTES, 128);
              const uint64_t adl = smem_desc(a_lo + s * 2 * CHUNK_BYTES, CHUNK_BYTES, 128);
              tc_mma(d, adh, bd, IDESC_2N, (g | s) != 0);
              tc_mma(d, adl, bd, IDESC_N, 1u);
            }
          }
          tc_commit(bar(B_EMPTY + sB));
          tc_commit(bar(A_EMPTY + sA));
        }
        tc_commit(bar(ACC_FULL + sa));
      }
    }
  } else {
    // ===================================================================== weight loader
    if (elect_one()) {
      const uint8_t* wb = static_cast<const uint8_t*>(p.weight_f16);
      const uint32_t b0 = smem_u32(b_buf);
      int b_it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        for (int g = 0; g < dg; ++g, ++b_it) {
          const int sB = b_it % NB;
          mbar_wait(bar(B_EMPTY + sB), ((b_it / NB) & 1) ^ 1);
          mbar_expect_tx(bar(B_FULL + sB), B_BLOCK);
          bulk_g2s(b0 + sB * B_BLOCK, wb + (int64_t)g * B_BLOCK, B_BLOCK, bar(B_FULL + sB));
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
  }
}

// weight_packed [C*9][O_pad] fp32 (row = c*9 + tap) -> per group g the fp16 block [128 rows][80] in the canonical
// K-major interleaved layout [(n/8)][(k/8)][n%8][k%8]; rows 0..63 = hi, 64..127 = lo * 2^12; k = tap*8 + c%8.
__global__ void dcn_pack_f16_kernel(const float* __restrict__ w, __half* __restrict__ out, int O, int O_pad, int dg) {
  const int total = dg * 2 * NT * KCH * 8;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    int r = i;
    const int k8 = r % 8; r /= 8;
    const int n8 = r % 8; r /= 8;
    const int kc = r % KCH; r /= KCH;
    const int ng = r % (2 * NT / 8);
    const int g = r / (2 * NT / 8);
    const int n2 = ng * 8 + n8, o = n2 % NT;
    float v = 0.f;
    if (kc < 9 && o < O) v = w[(int64_t)((g * 8 + k8) * 9 + kc) * O_pad + o];
    v = fminf(fmaxf(v, -65504.f), 65504.f);
    const __half hi = __float2half_rn(v);
    out[i] = (n2 < NT) ? hi : __float2half_rn((v - __half2float(hi)) * kLoScale);
  }
}

// NHWC (ld floats per pixel) -> group planar [(n*G + g)][H][W][8]
__global__ void nhwc_to_gp_kernel(const float* __restrict__ src, int ld, float* __restrict__ dst, int N, int64_t HW, int G) {
  const int64_t total = (int64_t)N * HW * G;  // one 32-byte sector per thread
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int g = (int)(i % G);
    const int64_t pix = i / G;  // n*HW + hw
    const int64_t n = pix / HW, hw = pix - n * HW;
    float v[8];
    ldg256(src + pix * ld + g * 8, v);
    stg256(dst + ((n * G + g) * HW + hw) * 8, v);
  }
}

}  // namespace dcn

int dcn_tc_supported(const TdvcDcnParams& p) {
  return p.weight_f16 != nullptr && p.input_gp != nullptr && p.params_planar == 1 && p.C == 8 * p.dg && p.O <= dcn::NT &&
         (reinterpret_cast<uintptr_t>(p.input_gp) & 31) == 0 && (reinterpret_cast<uintptr_t>(p.weight_f16) & 15) == 0;
}

int dcn_tc(const TdvcDcnParams& p, cudaStream_t st) {
  static int smem_done[kMaxDevices] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  const bool first = dev >= 0 && dev < kMaxDevices && smem_done[dev] == 0;
  if (int rc = ensure_dynamic_smem(dcn::dcn_tc_kernel, dcn::SMEM, smem_done, "dcn_tc")) return rc;
  // ask for the smallest shared-memory carve-out that holds the CTA: the rest of the 228 KB stays L1 for the gather
  if (first) cudaFuncSetAttribute(dcn::dcn_tc_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (dcn::SMEM + 1024) * 100 / (228 * 1024) + 1);
  const int tiles_x = cdiv(p.W, dcn::kTile), tiles_y = cdiv(p.H, dcn::kTileH);
  const int64_t items = (int64_t)p.N * tiles_x * tiles_y;
  TDVC_REQUIRE(items < (1ll << 31), "dcn_tc: too many work items");
  const int grid = (int)(items < kNumSMs ? items : kNumSMs);
  dcn::dcn_tc_kernel<<<grid, dcn::kThreads, dcn::SMEM, st>>>(p, tiles_x, tiles_y, (int)items);
  TDVC_CHECK_LAUNCH("dcn_tc");
  return TDVC_OK;
}

}  // namespace tdvc

using namespace tdvc;

extern "C" size_t tdvc_dcn_f16_bytes(int dg) { return dg > 0 ? (size_t)dg * dcn::B_BLOCK : 0; }

extern "C" int tdvc_dcn_pack_f16(const float* weight_packed, int O, int O_pad, int dg, void* out, void* stream) {
  TDVC_REQUIRE(weight_packed && out && dg > 0 && O > 0 && O <= dcn::NT && O_pad >= O, "dcn_pack_f16: bad arguments (O <= 64)");
  dcn::dcn_pack_f16_kernel<<<cdiv((int64_t)dg * dcn::B_BLOCK / 2, 256), 256, 0, (cudaStream_t)stream>>>(
      weight_packed, static_cast<__half*>(out), O, O_pad, dg);
  TDVC_CHECK_LAUNCH("dcn_pack_f16");
  return TDVC_OK;
}

extern "C" int tdvc_nhwc_to_group_planar(const float* src, int src_ld, float* dst, int N, int H, int W, int C, void* stream) {
  TDVC_REQUIRE(src && dst && N > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "nhwc_to_group_planar: bad shape (C %% 8 == 0)");
  TDVC_REQUIRE(src_ld % 8 == 0 && (reinterpret_cast<uintptr_t>(src) & 31) == 0 && (reinterpret_cast<uintptr_t>(dst) & 31) == 0,
               "nhwc_to_group_planar: 32-byte alignment");
  const int64_t total = (int64_t)N * H * W * (C / 8);
  int grid = cdiv(total, 256);
  if (grid > kNumSMs * 16) grid = kNumSMs * 16;
  dcn::nhwc_to_gp_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, src_ld, dst, N, (int64_t)H * W, C / 8);
  TDVC_CHECK_LAUNCH("nhwc_to_group_planar");
  return TDVC_OK;
}
