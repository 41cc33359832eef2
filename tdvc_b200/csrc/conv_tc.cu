// Implicit-GEMM convolution on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM) for the
// dense KxK convolutions of the P-frame graph (reference main/model/pnet.py passim, main/utils/utils.py:43-56,
// main/model/flownet.py:187-227, compressai blocks of SURVEY.md App. A).
//
// Numerics: activations and weights stay fp32 in HBM.  Each operand is split into two fp16 terms
// (x = x_hi + x_lo, x_hi = fp16(x), x_lo = fp16(x - x_hi): 22 significand bits) and the product is evaluated as
// (w_hi + w_lo) * (x_hi + x_lo) with fp32 accumulation in TMEM: relative error per product ~2^-22, i.e. the
// fp32-class accuracy the parity bar needs (>= 99.9 % identical quantised symbols).  w_lo is stored scaled by 2^12
// (kept out of the fp16 subnormal range; its accumulator rows are rescaled by 2^-12 in the epilogue - both exact).
// fp16 range: |x| is saturated to 65504 on conversion (cvt.satfinite); activations of an image codec with [0,1]
// inputs sit orders of magnitude below that.
//
// Operand roles (tools/mma_probe.cu, profiles/r01_mma_probe.txt): with the pixels on the M side every tcgen05.mma
// of N <= 128 costs >= 88 cycles (the 128-row operand fetch is not hidden), 352 cycles per 256 px * 16 k.  Here the
// WEIGHTS are the M-side operand - 128 rows = 64 output channels x (hi, lo) - and 256 PIXELS are the N side: two
// MMAs (x_hi, x_lo) per k-step run at the math floor (259 cycles measured, floor 256).
//
// Work decomposition (persistent, one CTA per SM, 640 threads = 20 warps):
//   work item  = 8 x 32 output pixels (N = 256; pixel p = ty*8 + tx) x 64 output channels;
//                K loop = cin chunks of CK channels ("units") x KS*KS taps x CK/16 MMA k-steps.
//   warps 8-17 producers: read the fp32 halo of one unit from global (float4 per lane, 16*CK/4 B contiguous per
//              pixel, all loads of a batch in flight), split into fp16 hi / lo and store it to shared memory in the
//              tcgen05 K-major no-swizzle canonical layout [channel/8][halo pixel][8 channels]: 8 x-adjacent pixels
//              form one 8x16 B core matrix and the 32 tile rows are 32 row groups (SBO = halo pitch), so every tap of
//              the convolution is the SAME buffer read through a descriptor whose start address is shifted by
//              (ky*pitch + kx)*16 B.  No im2col copy.  Stride 2: the halo is de-interleaved into 4 parity planes.
//   warp 19    streams pre-packed fp16 weight blocks [128 rows][CK] (one per unit x tap) with cp.async.bulk
//              (1-D TMA) into a ring, completion on mbarriers.
//   warp 18    one elected thread issues per k-step  D[128][256] += W * X_hi^T ; D += W * X_lo^T
//              and tcgen05.commit's the ring slots back to the producers / loader.
//   warps 0-7  epilogue, straight from TMEM to global memory (no shared-memory transposition, no barrier; the first
//              version went through a slab and was bound by that latency chain: image-input 3x3 0.28 -> 0.20 ms).
//              TMEM lane = weight row.  64-channel items: lane quadrant q = channels 16q..16q+15, hi rows in lanes 0-15,
//              lo rows in lanes 16-31; a warp reads 16 pixels per chunk (tcgen05.ld), merges hi + lo * 2^-12 with one
//              lane-xor-16 shuffle per pixel, and each half-warp stores one tile row: 16 lanes = 64 contiguous bytes of
//              an NHWC pixel.  128-channel items (split scheme): lane = channel, 32 lanes = 128 contiguous bytes.
//              Bias / GDN / activation / residuals are applied in registers; GDN multipliers and the first residual
//              are loaded one chunk ahead.  Overlaps the next item's MMAs (two accumulator stages = all 512 TMEM columns).
#include <cuda.h>
#include <stdlib.h>
#include <type_traits>
#include "tc_common.cuh"

#ifndef TDVC_CONV_TC_NX_MAX
#define TDVC_CONV_TC_NX_MAX 3   // activation stages in shared memory; a 4th fits since the epilogue slab went away but measured slower (64->64 3x3: 0.424 vs 0.409 ms)
#endif

namespace tdvc {

namespace tc {

constexpr int kEpiWarps = 8, kProdWarps = 10;                // 20 warps = 5 per SM sub-partition: still 96 registers per thread
constexpr int kMmaWarp = kEpiWarps + kProdWarps, kLoadWarp = kMmaWarp + 1;
constexpr int kThreads = (kEpiWarps + kProdWarps + 2) * 32;  // 640
constexpr int kProdThreads = kProdWarps * 32;
constexpr int TW = 8, TH = 32, NPX = TW * TH;  // output tile = N of one MMA
constexpr int NT = 64;                         // output channels per item (M = 2*NT rows: hi, lo)

// SPLIT = 0: item = 64 output channels, weight rows [hi | lo*2^12] of those channels (4 products, 2 MMAs per k-step);
// SPLIT = 1: item = 128 output channels, separate W_hi / W_lo blocks (pre-scaled by 2^w_shift so that W_lo stays in
//            the fp16 normal range) accumulating into ONE row per channel: x_hi*w_hi + x_lo*w_hi + x_hi*w_lo,
//            3 MMAs per k-step for twice the channels (-25 % tensor work) and no hi/lo merge in the epilogue.
// NPH > 1 ("row phases", for layers with <= 64 / NPH output channels, 7x7 only): an item is 8 x 32*NPH output pixels; the
//            MMA's 256 N columns are the pixels of every NPH-th row (row-group pitch = NPH halo rows) and the 64
//            (hi, lo) row pairs of the M side are 64 / NPH channels x NPH row phases: the rows of phase f hold the
//            kernel shifted down by f rows (tap (ky', kx) -> W[ky' - f][kx], zero outside), so one pass over KS + NPH - 1
//            kernel rows produces NPH output rows per N column.  A 32-channel 7x7 layer then needs 8 x 7 MMA taps per
//            512 pixels instead of 7 x 7 per 256 with half of the M rows idle.
// PR = 1 ("one product", TdvcConvParams::products): operands are plain fp16(w) and fp16(x), ONE MMA per k-step, no lo planes
//            and no lo rows: the 128 M rows are 128 output channels (weights pre-scaled by 2^w_shift like the split scheme) or,
//            for 3x3 layers with <= 64 output channels, 64 channels x 2 row phases.  This is the arithmetic of the reference
//            under autocast; it is used for the stages behind the last quantiser only (precision budget).
// TMA = 1 ("TMA-fed activations", stride-1 KxK layers): the fp32 halo of a unit is not loaded by the producer warps' LDGs but by
//            tensor-map bulk copies (cp.async.bulk.tensor.4d over (channel, x, y, image) of every source; the hardware clips the
//            box at the image border and zero-fills, so the border / channel-tail predicates of the LDG path disappear).  One
//            elected lane issues a box of [IH][IW][SUB channels] fp32 per sub-unit into a small staging ring; the converter
//            warps read the staged box linearly (conflict-free 16-byte reads), split it to fp16 hi / lo and write the operand
//            planes.  Loads in flight no longer cost registers, and their depth is the staging ring, not a register batch.
template <int KS, int CK, int S, int SPLIT = 0, int NPH = 1, int PLN = 0, int PR = 0, int TMA = 0>
struct Cfg {
  static_assert(S == 1 || (S == 2 && (KS == 1 || KS == 3)), "stride");
  static_assert(TMA == 0 || (S == 1 && KS != 1), "TMA-fed activations: stride-1 KxK layers (1x1 layers stage with cp.async)");
  static constexpr bool TMAIN = (TMA != 0);
  static_assert(PR == 0 || (PR == 1 && SPLIT == 0), "one product: SPLIT is meaningless, pass 0");
  static_assert(NPH == 1 || (S == 1 && SPLIT == 0 && ((KS == 7 && PR == 0 && (NPH == 2 || NPH == 4)) || (KS == 3 && PR == 1 && NPH == 2))),
                "row phases");
  // DIRECT: one accumulator row per (channel, phase) - the epilogue stores TMEM lanes as they are; otherwise the item's rows are
  // (hi, lo) pairs that the epilogue merges
  static constexpr bool DIRECT = (SPLIT == 1 || PR == 1);
  static constexpr bool ONE = (PR == 1);
  static constexpr int THO = TH * NPH;                       // output rows of an item
  // warp roles: 8 epilogue + 10 producer warps; the 1x1 stride-1 split-scheme layers (GDN / IGDN, 1x1 with >= 96 output
  // channels) are bound by the epilogue's instruction stream while their cp.async-staged producers have little to do, so
  // there the same 640 threads are 12 epilogue + 6 producer warps (3 epilogue warps per TMEM lane quadrant: 5, 5 and 6 of the
  // sixteen 16-column chunks; 16 + 2 made the two producer warps the bottleneck: GDN 0.147 -> 0.194 ms)
#ifndef TDVC_CONV_TC_WIDE16
#define TDVC_CONV_TC_WIDE16 0   // A/B switch: 12 + 6 warp roles for the image-input 3x3 layers as well (see below)
#endif
  static constexpr bool WIDE_EPI = (KS == 1 && S == 1 && DIRECT) || (TDVC_CONV_TC_WIDE16 && KS == 3 && S == 1 && CK == 16 && NPH == 1 && !DIRECT);
  static constexpr int EPIW = WIDE_EPI ? 12 : kEpiWarps, PRODW = WIDE_EPI ? 6 : kProdWarps;
  static constexpr int CONVW = TMAIN ? PRODW - 1 : PRODW;   // warps that write operand planes (TMA: one producer warp issues the copies)
  static_assert(EPIW + PRODW + 2 == kThreads / 32 && EPIW % 4 == 0, "warp roles");
  static constexpr int NGRP = EPIW / 4;                      // epilogue warps per TMEM lane quadrant: they share the 16 chunks
  static constexpr int KY = KS + NPH - 1;                    // kernel rows walked by the MMA loop
  static_assert(CK == 16 || CK == 32, "cin chunk");
  static constexpr int PAD = KS / 2;
  // input halo of one 8x32 output tile: IH x IW input pixels, input pixel = origin + h*STEP
  static constexpr int STEP = (KS == 1) ? S : 1;            // 1x1: only every S-th input pixel is touched
  static constexpr int IH = (KS == 1) ? TH : (THO - 1) * S + KS;
  static constexpr int IW = (KS == 1) ? TW : (TW - 1) * S + KS;
  // stride-2 3x3: the halo is stored de-interleaved into 4 parity planes (row parity, column parity) so that every
  // tap again reads 8 x-adjacent plane pixels per core matrix: tap (ky,kx) -> plane (ky&1, kx&1), shift (ky>>1, kx>>1)
  static constexpr bool PLANES = (S == 2 && KS == 3);
  static constexpr int PW = PLANES ? TW + 1 : IW;           // plane (or halo) row pitch in pixels
  static constexpr int PH = PLANES ? TH + 1 : IH;
  static constexpr int NHALO = IH * IW;                     // pixels the producers visit
  static constexpr int NPIX = PLANES ? 4 * PH * PW : IH * IW;
  // channel-group pitch in pixels, chosen so that a warp's 8-byte stores spread evenly over the banks:
  // CK = 32 (4 channel groups x 4 pixels per store): pitch = 4 (mod 8);  CK = 16 (2 x 8): any odd pitch
  static constexpr int NPIXP = (CK == 32) ? ((NPIX + 3) / 8 * 8 + 4) : (NPIX | 1);
  static constexpr int NCH8 = CK / 8;
  static constexpr int X_HALF = NCH8 * NPIXP * 16;   // bytes of the hi (or lo) plane of one unit
  static constexpr int X_STAGE = (PR == 1 ? 1 : 2) * X_HALF;
  static constexpr int LBO_X = NPIXP * 16, SBO_X = NPH * PW * 16;
  static constexpr int NTT = (DIRECT ? 128 : 64) / NPH;  // output channels per item
  static constexpr int W_HALF = 128 * CK * 2;        // one [128 rows][CK] fp16 block
  static constexpr int W_BLOCK = SPLIT ? 2 * W_HALF : W_HALF;   // SPLIT: [W_hi block | W_lo block]
  static constexpr int LBO_W = 128, SBO_W = NCH8 * 128;
  static constexpr int KSTEPS = CK / 16;
  static constexpr int TAPS = KY * KS;               // weight blocks per unit
  // weight ring depth: a one-product block is consumed in KSTEPS * 128 cycles, well below the latency of its bulk copy
  static constexpr int NW = PR == 1 ? 8 : ((CK == 32) ? 4 : (SPLIT ? 4 : 6));
  // 1x1 stride-1 layers are bound by the depth of the producers' load pipeline (a few float4 per thread in registers), not by
  // the tensor pipe: their producers stage the fp32 unit in shared memory with cp.async, NSTG units ahead (each thread reads
  // back only the 16-byte slots it copied itself, so no barrier is involved), and convert from there.
  static constexpr bool STAGED = (KS == 1 && S == 1);
#ifndef TDVC_CONV_TC_NSTG
#define TDVC_CONV_TC_NSTG 2
#endif
  static constexpr int NSTG = TDVC_CONV_TC_NSTG;
  static constexpr int LOADS_PER_THREAD = (NHALO * (CK / 4) / 32 + PRODW - 1) / PRODW;   // = ProdCfg::PER_WARP
  static constexpr int STG_UNIT = LOADS_PER_THREAD * PRODW * 32 * 16;
  static constexpr int STG_BYTES = STAGED ? NSTG * STG_UNIT : 0;
  // PLN: planar (NCHW) output through tensor-map TMA stores: per epilogue warp two staging boxes of [32 channels][2 rows][8 px]
  static_assert(PLN == 0 || (SPLIT == 1 && NPH == 1), "TMA planar stores: split-scheme items only");
  static constexpr int PLN_BYTES = PLN ? EPIW * 2 * 2048 : 0;
  static constexpr int kNxMax = TDVC_CONV_TC_NX_MAX;
  static constexpr int kBarBytes = 384;
  // TMA staging ring: sub-units of SUB channels ([IH][IW][SUB] fp32 = NHALO * SUB * 4 bytes, a multiple of 128), NTS stages.
  // Preference: three operand stages with >= 2 staging stages of 16 channels, else 8-channel sub-units, else two operand stages
  static constexpr int fixed_bytes = NW * W_BLOCK + STG_BYTES + PLN_BYTES + kBarBytes;
  static constexpr bool tfits(int nx, int sub, int nts) { return nx * X_STAGE + fixed_bytes + nts * NHALO * sub * 4 <= 227 * 1024; }
  static constexpr int pick(int what) {   // what: 0 = NX, 1 = SUB, 2 = NTS
    if (!TMAIN) return what == 0 ? ((kNxMax >= 4 && tfits(4, 0, 0)) ? 4 : (tfits(3, 0, 0) ? 3 : 2)) : (what == 1 ? 16 : 0);
    const int cand[8][3] = {{3, 16, 3}, {3, 16, 2}, {3, 8, 3}, {2, 16, 3}, {2, 16, 2}, {2, 8, 3}, {2, 8, 2}, {2, 8, 1}};
    for (int i = 0; i < 8; ++i)
      if (tfits(cand[i][0], cand[i][1], cand[i][2])) return cand[i][what];
    return what == 0 ? 2 : (what == 1 ? 8 : 1);
  }
  static constexpr int NX = pick(0);
  static constexpr int SUB = pick(1), NSUB = CK / SUB, NTS = pick(2);
  static constexpr int TSTG = NHALO * SUB * 4;        // bytes of one staged sub-unit
  static constexpr int TS_BYTES = NTS * TSTG;
  static constexpr int SMEM = NX * X_STAGE + fixed_bytes + TS_BYTES;
  static_assert(!TMAIN || (TSTG % 128 == 0 && IW <= 256 && IH <= 256), "TMA box");
  static_assert(NPIXP >= NPIX, "pitch");
  static_assert(SMEM <= 227 * 1024, "shared memory budget");
  // smem pixel slot of halo pixel (hy, hx)
  __host__ __device__ static constexpr int slot(int hy, int hx) {
    return PLANES ? (((hy & 1) * 2 + (hx & 1)) * PH + (hy >> 1)) * PW + (hx >> 1) : hy * IW + hx;
  }
  // smem pixel slot read by output pixel (0,0) for tap (ky, kx)
  __host__ __device__ static constexpr int tap_slot(int ky, int kx) {
    return PLANES ? (((ky & 1) * 2 + (kx & 1)) * PH + (ky >> 1)) * PW + (kx >> 1) : ky * IW + kx;
  }
};

constexpr int TMEM_COLS = 2 * NPX;  // two accumulator stages

struct Item {
  int n, y0, x0, jt;
};

// `flip` >= 0: items are walked in descending order (item -> flip - item), see TdvcConvParams::order
__device__ __forceinline__ Item decode_item(int item, int n_jt, int tiles_x, int tiles_y, int flip, int th = TH) {
  Item it;
  if (flip >= 0) item = flip - item;
  it.jt = item % n_jt;
  int st = item / n_jt;
  it.x0 = (st % tiles_x) * TW;
  st /= tiles_x;
  it.y0 = (st % tiles_y) * th;
  it.n = st / tiles_y;
  return it;
}

// GDN / IGDN (in_square): x*x is pre-scaled by 2^-s so that max x*x < 2^15 stays inside the fp16 operand range (fp16 saturates
// at 65504, i.e. |x| > 255 would clamp silently); s comes from the device scalar the producing layer's epilogue maintains
// (TdvcConvParams::out_absmax -> in_absmax) and is undone exactly in the epilogue.  |x| <= 181: s = 0, nothing changes.
__device__ __forceinline__ int square_shift(const TdvcConvParams& p) {
  if (!p.in_square || p.in_absmax == nullptr) return 0;
  const float m = __ldg(p.in_absmax);
  if (!(m * m > 32768.f)) return 0;
  const int e = ((__float_as_int(m) >> 23) & 0xff) - 127;   // m = f * 2^e, 1 <= f < 2  =>  m*m < 2^(2e+2)
  const int s = 2 * e + 2 - 15;
  return s < 0 ? 0 : (s > 100 ? 100 : s);
}
__device__ __forceinline__ float pow2f(int e) { return __int_as_float((127 + e) << 23); }   // -126 <= e <= 127
// per-warp running max of |v| over the values an epilogue warp stored -> one atomic per warp (non-negative floats order like ints)
__device__ __forceinline__ void absmax_commit(float* dst, float amax) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  if ((threadIdx.x & 31) == 0 && dst != nullptr) atomicMax(reinterpret_cast<int*>(dst), __float_as_int(amax));
}

// ---- producer helpers -------------------------------------------------------------------------------------------
template <class C, int CK>
struct ProdCfg {
  static constexpr int LPP = CK / 4;                       // lanes (float4) per pixel
  static constexpr int PPI = 32 / LPP;                     // pixels per warp-wide load
  static constexpr int NLD = (C::NHALO + PPI - 1) / PPI;   // warp-wide loads per unit
  static constexpr int PER_WARP = (NLD + C::PRODW - 1) / C::PRODW;
  static constexpr int NBATCH = PER_WARP <= 12 ? 2 : 4;    // even: the two register buffers alternate
  static constexpr int BATCH = (PER_WARP + NBATCH - 1) / NBATCH;
};
struct ProdThread {
  int pw, psub;
  uint32_t vmask, lane_smem;
  float sq;   // in_square: exact power-of-two scale of x*x (square_shift), else unused
};
struct ProdUnit {
  const float* org;
  uint8_t* hi;
  int sld, iy0, ix0, stage;
  bool fast;
};

// issue the loads of batch B of a unit into v (zeros where the halo leaves the image / the channel range)
template <class C, class P, int B>
__device__ __forceinline__ void prod_issue(const TdvcConvParams& p, const ProdThread& th, const int (&tab)[P::PER_WARP],
                                           const ProdUnit& c, float4 (&v)[P::BATCH]) {
  if (c.fast) {
#pragma unroll
    for (int k = 0; k < P::BATCH; ++k) {
      const int kk = B * P::BATCH + k;
      v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (kk < P::PER_WARP) {
        if ((th.vmask >> kk) & 1) {
          int po = tab[kk < P::PER_WARP ? kk : 0];
          if (C::PLANES) po = (po >> 8) * p.W + (po & 255);
          v[k] = __ldg(reinterpret_cast<const float4*>(c.org + (int64_t)po * c.sld));
        }
      }
    }
  } else {
#pragma unroll
    for (int k = 0; k < P::BATCH; ++k) {
      const int kk = B * P::BATCH + k;
      const int sidx = (kk * C::PRODW + th.pw) * P::PPI + th.psub;
      const int hy = sidx / C::IW, hx = sidx - hy * C::IW;
      const int iy = c.iy0 + hy * C::STEP, ix = c.ix0 + hx * C::STEP;
      v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (kk < P::PER_WARP && sidx < C::NHALO && c.org != nullptr && iy >= 0 && iy < p.H && ix >= 0 && ix < p.W)
        v[k] = __ldg(reinterpret_cast<const float4*>(c.org + ((int64_t)(hy * C::STEP) * p.W + hx * C::STEP) * c.sld));
    }
  }
}

// split batch B into fp16 hi / lo and store it at its halo slots of the unit's stage
template <class C, class P, int B>
__device__ __forceinline__ void prod_convert(const TdvcConvParams& p, const ProdThread& th, const int (&tab)[P::PER_WARP],
                                             const ProdUnit& c, float4 (&v)[P::BATCH]) {
  if (p.in_square) {
#pragma unroll
    for (int k = 0; k < P::BATCH; ++k) {
      v[k].x = v[k].x * v[k].x * th.sq; v[k].y = v[k].y * v[k].y * th.sq;
      v[k].z = v[k].z * v[k].z * th.sq; v[k].w = v[k].w * v[k].w * th.sq;
    }
  }
#pragma unroll
  for (int k = 0; k < P::BATCH; ++k) {
    const int kk = B * P::BATCH + k;
    if (kk < P::PER_WARP) {
      if ((th.vmask >> kk) & 1) {
        uint2 hv, lv;
        if constexpr (C::ONE) {
          hv.x = pack_h2_sat(v[k].x, v[k].y);
          hv.y = pack_h2_sat(v[k].z, v[k].w);
        } else {
          split4(v[k], hv, lv);
        }
        uint8_t* dst;
        if (C::PLANES) {
          const int po = tab[kk < P::PER_WARP ? kk : 0];
          dst = c.hi - (th.pw * P::PPI + th.psub) * 16 + C::slot(po >> 8, po & 255) * 16;
        } else {
          dst = c.hi + kk * (C::PRODW * P::PPI * 16);   // flat slot = pixel index
        }
        *reinterpret_cast<uint2*>(dst) = hv;
        if constexpr (!C::ONE) *reinterpret_cast<uint2*>(dst + C::X_HALF) = lv;
      }
    }
  }
}

// batch index known only after unrolling the caller's loop (it is a compile-time constant there): dispatch 0..5
template <class C, class P>
__device__ __forceinline__ void prod_issue_rt(const TdvcConvParams& p, const ProdThread& th, const int (&tab)[P::PER_WARP],
                                              const ProdUnit& c, float4 (&v)[P::BATCH], int b) {
  if (b == 0) prod_issue<C, P, 0>(p, th, tab, c, v);
  else if (b == 1) prod_issue<C, P, 1>(p, th, tab, c, v);
  else if (b == 2) prod_issue<C, P, 2>(p, th, tab, c, v);
  else if (b == 3) prod_issue<C, P, 3>(p, th, tab, c, v);
  else if (b == 4) prod_issue<C, P, 4>(p, th, tab, c, v);
  else prod_issue<C, P, 5>(p, th, tab, c, v);
}
template <class C, class P>
__device__ __forceinline__ void prod_convert_rt(const TdvcConvParams& p, const ProdThread& th, const int (&tab)[P::PER_WARP],
                                                const ProdUnit& c, float4 (&v)[P::BATCH], int b) {
  if (b == 0) prod_convert<C, P, 0>(p, th, tab, c, v);
  else if (b == 1) prod_convert<C, P, 1>(p, th, tab, c, v);
  else if (b == 2) prod_convert<C, P, 2>(p, th, tab, c, v);
  else if (b == 3) prod_convert<C, P, 3>(p, th, tab, c, v);
  else if (b == 4) prod_convert<C, P, 4>(p, th, tab, c, v);
  else prod_convert<C, P, 5>(p, th, tab, c, v);
}

// PlanarMap: the 4-D tensor map (x, y, channel, image) of a planar output for the PLN variant, an empty tag otherwise
struct NoMap {};
// InMaps: one 4-D tensor map (channel, x, y, image) per concatenated source for the TMA variant, an empty tag otherwise
struct InMaps {
  CUtensorMap m[4];
};
template <int KS, int CK, int S, int SPLIT, int NPH, int PLN, int PR, int TMA>
__global__ void __launch_bounds__(kThreads, 1) conv_tc_kernel(const TdvcConvParams p, int tiles_x, int tiles_y, int n_jt,
                                                              int n_units, int n_items,
                                                              const __grid_constant__ std::conditional_t<PLN != 0, CUtensorMap, NoMap> omap,
                                                              const __grid_constant__ std::conditional_t<TMA != 0, InMaps, NoMap> imaps) {
  using C = Cfg<KS, CK, S, SPLIT, NPH, PLN, PR, TMA>;
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* ts_buf = smem;                                 // TMA staging ring of the fp32 sub-units (TS_BYTES, may be 0; 128-byte aligned)
  uint8_t* x_buf = smem + C::TS_BYTES;                    // NX stages of [hi plane | lo plane]
  uint8_t* w_buf = x_buf + C::NX * C::X_STAGE;            // NW weight blocks
  uint8_t* stg_buf = w_buf + C::NW * C::W_BLOCK;           // cp.async staging of the 1x1 producers (STG_BYTES, may be 0)
  uint8_t* pln_buf = stg_buf + C::STG_BYTES;              // TMA planar-store staging of the epilogue warps (PLN_BYTES, may be 0)
  uint64_t* bars = reinterpret_cast<uint64_t*>(pln_buf + C::PLN_BYTES);
  // barrier indices
  constexpr int X_FULL = 0, X_EMPTY = X_FULL + C::NX, W_FULL = X_EMPTY + C::NX, W_EMPTY = W_FULL + C::NW,
                ACC_FULL = W_EMPTY + C::NW, ACC_EMPTY = ACC_FULL + 2, TS_FULL = ACC_EMPTY + 2, TS_EMPTY = TS_FULL + C::NTS,
                NBARS = TS_EMPTY + C::NTS;
  static_assert(NBARS * 8 + 8 <= C::kBarBytes, "barrier area");
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NBARS);
  const uint32_t bar0 = smem_u32(bars);
  auto bar = [&](int i) { return bar0 + 8u * i; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int flip = p.order ? n_items - 1 : -1;

  if (threadIdx.x == 0) {
    for (int i = 0; i < C::NX; ++i) { mbar_init(bar(X_FULL + i), C::CONVW * 32); mbar_init(bar(X_EMPTY + i), 1); }
    for (int i = 0; i < C::NTS; ++i) { mbar_init(bar(TS_FULL + i), 1); mbar_init(bar(TS_EMPTY + i), C::CONVW * 32); }
    for (int i = 0; i < C::NW; ++i) { mbar_init(bar(W_FULL + i), 1); mbar_init(bar(W_EMPTY + i), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(bar(ACC_FULL + i), 1); mbar_init(bar(ACC_EMPTY + i), C::EPIW * 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == C::EPIW + C::PRODW) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  if (warp < C::EPIW) {
    // ===================================================================== epilogue (8 warps: lane quadrant x pixel half)
    // Warp = (TMEM lane quadrant, pixel half): it reads its 32 accumulator rows 16 pixels (two tile rows) at a time with
    // tcgen05.ld (the next chunk's load in flight while the current one is processed) and stores them itself.
    const int sq_shift = square_shift(p);   // GDN / IGDN: the accumulator holds the norm * 2^-sq_shift
    float amax = 0.f;                       // max |v| over the values this lane stored (TdvcConvParams::out_absmax)
    const bool track = p.out_absmax != nullptr;
    // TdvcConvParams::chan_sum (squeeze-excitation means fused into the producing layer): running sum of the values this lane
    // stored for its channel in image csum_n; flushed to row (CTA, epilogue column group, row phase) when the image changes.
    // Every (row, image, channel) cell has one owner and the items of a CTA are walked in a fixed order: deterministic.
    const bool csum_on = p.chan_sum != nullptr;
    float csum = 0.f;
    int csum_n = -1;
    if constexpr (C::DIRECT) {
    // ---- 128 output channels per item, one accumulator row per channel (TMEM lane = channel): every lane stores its own
    //      channel straight from the registers tcgen05.ld filled - per pixel the 32 lanes of a warp write 128 contiguous
    //      bytes of the NHWC pixel and the 4 quadrant warps cover its 512 bytes.  No shared-memory transposition.
    const int quad = warp & 3, grp = warp >> 2;   // TMEM lane quadrant, column group
    // this warp's share of the sixteen 16-pixel chunks of an item: chunks chunk0 .. chunk0 + nchunk - 1
    const int chunk0 = C::NGRP == 2 ? 8 * grp : (16 * grp) / C::NGRP;
    const int nchunk = C::NGRP == 2 ? 8 : (16 * (grp + 1)) / C::NGRP - chunk0;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
    const int t = quad * 32 + lane;
    // row phases (one-product 3x3 layers with <= 64 channels): row t = phase phi, channel cch of the item
    constexpr int NPHS = NPH;
    const int phi = t / C::NTT, cch = t - phi * C::NTT;
    const int Ho = p.Ho, Wo = p.Wo, cout = p.cout, act = p.act, post = p.post;
    const int sh = p.shuffle == 2 ? 2 : 1;
    const int cr = cout >> 2;
    const int oW = Wo * sh, oH = Ho * sh;
    const bool planar = p.out_planar != 0;
    const bool planar_vec = planar && (Wo & 7) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 31) == 0;
    const float a_neg = act == TDVC_ACT_NONE ? 1.f : (act == TDVC_ACT_LRELU ? p.slope : 0.f);
    const float a_hi = act == TDVC_ACT_CLAMP01 ? 1.f : __int_as_float(0x7f800000);
    const bool has_act = act != TDVC_ACT_NONE;
    const float unscale = pow2f(-p.w_shift) * pow2f(sq_shift);   // 2^(sq_shift - w_shift), exact
    const int o_rs = NPHS * sh * oW * p.out_ld, m_rs = NPHS * sh * oW * p.mul_ld, r1_rs = NPHS * sh * oW * p.res1_ld,
              r2_rs = NPHS * sh * oW * p.res2_ld;
    const int o_xs = sh * p.out_ld, m_xs = sh * p.mul_ld, r1_xs = sh * p.res1_ld, r2_xs = sh * p.res2_ld;
    int acc_it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++acc_it) {
      const Item it = decode_item(item, n_jt, tiles_x, tiles_y, flip, C::THO);
      const int sa = acc_it & 1;
      mbar_wait(bar(ACC_FULL + sa), (acc_it >> 1) & 1);
      tc_fence_after();
      const int co = it.jt * C::NTT + cch;
      const bool ch_ok = co < cout;
      const float wbias = (p.bias && ch_ok) ? __ldg(p.bias + co) : 0.f;
      if (csum_on && it.n != csum_n) {
        if (csum_n >= 0 && ch_ok)
          p.chan_sum[((int64_t)((blockIdx.x * C::NGRP + grp) * NPHS + phi) * p.N + csum_n) * cout + co] = csum;
        csum_n = it.n;
        csum = 0.f;
      }
      int oc = co, qy = 0, qx = 0;
      if (sh == 2) {
        const int q = ch_ok ? co / cr : 0;
        oc = co - q * cr;
        qy = q >> 1;
        qx = q & 1;
      }
      const int yb = it.y0 + phi;                   // first output row of this lane's phase
      const int ny = Ho > yb ? (Ho - yb + NPHS - 1) / NPHS : 0, nx = Wo - it.x0;   // valid tile rows / columns
      const int64_t pix0 = ((int64_t)it.n * oH + (yb * sh + qy)) * oW + (it.x0 * sh + qx);
      float* const o0 = planar ? p.out + (((int64_t)it.n * cout + co) * Ho + yb) * Wo + it.x0 : p.out + pix0 * p.out_ld + oc;
      const float* const m0 = post != TDVC_POST_NONE ? p.mul + pix0 * p.mul_ld + oc : nullptr;
      const float* const r10 = p.res1 ? p.res1 + pix0 * p.res1_ld + oc : nullptr;
      const float* const r20 = p.res2 ? p.res2 + pix0 * p.res2_ld + oc : nullptr;
      const bool side = m0 || r10 || r20;
      uint32_t r[16];
      tmem_ld16(lane_addr + (uint32_t)(sa * NPX + chunk0 * 16), r);
      // residual / GDN-multiplier rows are pulled into L2 two chunks (4 tile rows) ahead: lane l < 16 covers pixel (row l / 8,
      // column l % 8) of the chunk - the 128 bytes of this quadrant's 32 channels are one line
      auto l2_prefetch_rows = [&](int tyb) {
        const int ty = tyb + (lane >> 3), k = lane & 7;
        if (lane < 16 && ty < ny && k < nx && ch_ok) {
          if (m0) prefetch_l2(m0 + ty * m_rs + k * m_xs);
          if (r10) prefetch_l2(r10 + ty * r1_rs + k * r1_xs);
          if (r20) prefetch_l2(r20 + ty * r2_rs + k * r2_xs);
        }
      };
      // GDN multiplier (or, without one, the first residual) of the two rows a lane stores next, held in registers one
      // chunk ahead
      const float* const pf0 = m0 ? m0 : r10;
      const int pf_rs = m0 ? m_rs : r1_rs, pf_xs = m0 ? m_xs : r1_xs;
      float ra[16];
      auto res_load = [&](int tyb) {
        if (pf0 && ch_ok && nx >= 8) {
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {
            if (tyb + h2 < ny) {
              const float* rp = pf0 + (tyb + h2) * pf_rs;
#pragma unroll
              for (int k = 0; k < 8; ++k, rp += pf_xs) ra[h2 * 8 + k] = __ldg(rp);
            }
          }
        }
      };
      if (side) {
        l2_prefetch_rows(2 * chunk0);
        l2_prefetch_rows(2 * chunk0 + 2);
        res_load(2 * chunk0);
      }
#pragma unroll 1
      for (int c = 0; c < nchunk; ++c) {
        const int ty0 = 2 * chunk0 + c * 2;
        tmem_ld_wait();
        if (c == nchunk - 1) {
          tc_fence_before();
          mbar_arrive(bar(ACC_EMPTY + sa));
        }
        float o[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) o[j] = fmaf(__uint_as_float(r[j]), unscale, wbias);
        if (c < nchunk - 1) tmem_ld16(lane_addr + (uint32_t)(sa * NPX + chunk0 * 16 + (c + 1) * 16), r);
        if constexpr (PLN != 0) {
          // Planar (NCHW) output.  A lane's tile row is one 32-byte sector of its own plane, so a warp-wide st.global.v8
          // scatters over 32 lines - four times the LSU wavefronts of an NHWC store, and the producers' operand stores
          // queue behind them (64->216 3x3 @1024x1920: 1.71 ms planar vs 1.13 ms NHWC).  Instead the warp stages its
          // [32 channels][2 rows][8 px] box in shared memory and one lane hands it to the TMA engine (4-D tensor map over
          // (x, y, channel, image)), which also clips the box at the image and channel-count edges.
          if (planar && ty0 < ny) {
            if (has_act)
#pragma unroll
              for (int j = 0; j < 16; ++j) o[j] = fminf(fmaxf(o[j], a_neg * o[j]), a_hi);
            uint8_t* box = pln_buf + (warp * 2 + (c & 1)) * 2048;
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // the store issued two chunks ago has read this box
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 4; ++j)
              reinterpret_cast<float4*>(box + lane * 64)[j] = make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
            fence_async_smem();
            __syncwarp();
            if (lane == 0) {
              asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%1, %2, %3, %4}], [%5];" ::"l"(
                               reinterpret_cast<uint64_t>(&omap)),
                           "r"(it.x0), "r"(it.y0 + ty0), "r"(it.jt * C::NTT + quad * 32), "r"(it.n), "r"(smem_u32(box))
                           : "memory");
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            continue;
          }
        }
        if (!(ch_ok && ty0 < ny)) continue;
        if (side && c < nchunk - 2) l2_prefetch_rows(ty0 + 4);
        if (planar) {  // NCHW planes: the 8 x-adjacent pixels of a tile row are contiguous in this channel's plane
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {
            if (ty0 + h2 >= ny) continue;
            if (has_act)
#pragma unroll
              for (int k = 0; k < 8; ++k) o[h2 * 8 + k] = fminf(fmaxf(o[h2 * 8 + k], a_neg * o[h2 * 8 + k]), a_hi);
            float* op = o0 + (int64_t)(ty0 + h2) * NPHS * Wo;
            if (planar_vec && nx >= 8) {
              stg256(op, o + h2 * 8);   // one full 32-byte sector
            } else {
#pragma unroll
              for (int k = 0; k < 8; ++k)
                if (k < nx) op[k] = o[h2 * 8 + k];
            }
          }
          continue;
        }
        if (nx >= 8) {   // interior tile columns
          if (m0) {   // GDN / IGDN: v = mul * rsqrt(v) | mul * sqrt(v)   (rows past ny are never stored)
            if (post == TDVC_POST_IGDN) {
#pragma unroll
              for (int j = 0; j < 16; ++j) o[j] = ra[j] * sqrt_fast(o[j]);
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) o[j] = ra[j] * rsqrt_fast(o[j]);
            }
          }
          if (has_act)
#pragma unroll
            for (int j = 0; j < 16; ++j) o[j] = fminf(fmaxf(o[j], a_neg * o[j]), a_hi);
          if (r10) {
            if (m0) {   // multiplier and residual together: the residual is loaded in place
#pragma unroll
              for (int h2 = 0; h2 < 2; ++h2) {
                if (ty0 + h2 >= ny) continue;
                const float* rp = r10 + (ty0 + h2) * r1_rs;
                float rc[8];
#pragma unroll
                for (int k = 0; k < 8; ++k, rp += r1_xs) rc[k] = __ldg(rp);
#pragma unroll
                for (int k = 0; k < 8; ++k) o[h2 * 8 + k] += rc[k];
              }
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) o[j] += ra[j];
            }
          }
          if (pf0 && c < nchunk - 1) res_load(ty0 + 2);
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {
            if (ty0 + h2 >= ny) continue;
            if (r20) {
              const float* rp = r20 + (ty0 + h2) * r2_rs;
              float rc[8];
#pragma unroll
              for (int k = 0; k < 8; ++k, rp += r2_xs) rc[k] = __ldg(rp);
#pragma unroll
              for (int k = 0; k < 8; ++k) o[h2 * 8 + k] += rc[k];
            }
            float* op = o0 + (ty0 + h2) * o_rs;
#pragma unroll
            for (int k = 0; k < 8; ++k, op += o_xs) *op = o[h2 * 8 + k];
            if (track)
#pragma unroll
              for (int k = 0; k < 8; ++k) amax = fmaxf(amax, fabsf(o[h2 * 8 + k]));
            if (csum_on)
#pragma unroll
              for (int k = 0; k < 8; ++k) csum += o[h2 * 8 + k];
          }
        } else {         // ragged right edge (Wo % 8 != 0): one pixel at a time
#pragma unroll 1
          for (int j = 0; j < 16; ++j) {
            const int ty = ty0 + (j >> 3), k = j & 7;
            if (ty >= ny || k >= nx) continue;
            float v = o[0];
#pragma unroll
            for (int q = 1; q < 16; ++q) v = (j == q) ? o[q] : v;
            if (m0) {
              const float mvk = __ldg(m0 + ty * m_rs + k * m_xs);
              v = mvk * (post == TDVC_POST_IGDN ? sqrtf(v) : rsqrtf(v));
            }
            v = fminf(fmaxf(v, a_neg * v), a_hi);
            if (r10) v += __ldg(r10 + ty * r1_rs + k * r1_xs);
            if (r20) v += __ldg(r20 + ty * r2_rs + k * r2_xs);
            o0[ty * o_rs + k * o_xs] = v;
            amax = fmaxf(amax, fabsf(v));
            csum += v;
          }
        }
      }
    }
    if (csum_on && csum_n >= 0) {
      const int co = cch;   // chan_sum needs a single output-channel tile (launch() checks it)
      if (co < cout) p.chan_sum[((int64_t)((blockIdx.x * C::NGRP + grp) * NPHS + phi) * p.N + csum_n) * cout + co] = csum;
    }
    if constexpr (PLN != 0) {
      if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // every staged box has been written out
    }
    } else {
    // ---- 64 output channels per item, rows [hi | lo * 2^12].  Row order (pack_f16_kernel): TMEM lane quadrant q holds channels
    //      16q..16q+15, hi rows in lanes 0-15 and the matching lo rows in lanes 16-31, so the two terms of a channel sit in
    //      ONE warp and are merged with a lane-xor-16 shuffle - no shared-memory transposition, no barrier.  After the merge
    //      lanes 0-15 own the chunk's first tile row (8 pixels) and lanes 16-31 the second: per pixel the 16 lanes write 64
    //      contiguous bytes (two full sectors) of the NHWC pixel, and the 4 quadrant warps cover its 256 bytes.
    const int quad = warp & 3, grp = warp >> 2;   // TMEM lane quadrant, column group
    // this warp's share of the sixteen 16-pixel chunks of an item: chunks chunk0 .. chunk0 + nchunk - 1
    const int chunk0 = C::NGRP == 2 ? 8 * grp : (16 * grp) / C::NGRP;
    const int nchunk = C::NGRP == 2 ? 8 : (16 * (grp + 1)) / C::NGRP - chunk0;
    const int part = lane >> 4, cl = lane & 15;
    // row phases (Cfg::NPH): this lane's (hi, lo) row pair belongs to phase phi and channel cch of the item
    constexpr int NPHS = NPH;
    const int phi = (quad * 16 + cl) / C::NTT, cch = (quad * 16 + cl) - phi * C::NTT;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
    const int Ho = p.Ho, Wo = p.Wo, cout = p.cout, act = p.act, post = p.post;
    const int sh = p.shuffle == 2 ? 2 : 1;
    const int cr = cout >> 2;
    const int oW = Wo * sh, oH = Ho * sh;
    const bool planar = p.out_planar != 0;
    const bool planar_vec = planar && (Wo & 7) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 31) == 0;
    // a plain store may also fill the pad lanes of a channel-padded view (weight rows >= cout are zero, no bias there)
    const int cout_st = (!planar && sh == 1 && post == TDVC_POST_NONE && !p.res1 && !p.res2 && p.out_ld >= ((cout + 3) & ~3)) ? ((cout + 3) & ~3) : cout;
    // branch-free activation: a(v) = min(max(v, a_neg * v), a_hi)   (0 <= a_neg <= 1)
    const float a_neg = act == TDVC_ACT_NONE ? 1.f : (act == TDVC_ACT_LRELU ? p.slope : 0.f);
    const float a_hi = act == TDVC_ACT_CLAMP01 ? 1.f : __int_as_float(0x7f800000);
    const bool has_act = act != TDVC_ACT_NONE;
    const float rscale = (part ? kLoUnscale : 1.f) * pow2f(sq_shift);   // lo rows carry w_lo * 2^12
    // element strides of one tile row / one tile column in out / mul / res1 / res2 (all address the same logical pixel)
    // (a tile row is NPHS image rows apart; the phase offset is folded into the base pointers)
    const int o_rs = NPHS * sh * oW * p.out_ld, m_rs = NPHS * sh * oW * p.mul_ld, r1_rs = NPHS * sh * oW * p.res1_ld,
              r2_rs = NPHS * sh * oW * p.res2_ld;
    const int o_xs = sh * p.out_ld, m_xs = sh * p.mul_ld, r1_xs = sh * p.res1_ld, r2_xs = sh * p.res2_ld;
    int acc_it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++acc_it) {
      const Item it = decode_item(item, n_jt, tiles_x, tiles_y, flip, C::THO);
      const int sa = acc_it & 1;
      mbar_wait(bar(ACC_FULL + sa), (acc_it >> 1) & 1);
      tc_fence_after();
      const int co = it.jt * C::NTT + cch;   // this lane's output channel (same for its hi and lo lane)
      if (csum_on && it.n != csum_n) {   // (uniform over the warp: all lanes see the same item)
        const float both = csum + __shfl_xor_sync(0xffffffffu, csum, 16);   // the two tile rows of a chunk sit in lanes l, l ^ 16
        if (csum_n >= 0 && part == 0 && co < cout)
          p.chan_sum[((int64_t)((blockIdx.x * C::NGRP + grp) * NPHS + phi) * p.N + csum_n) * cout + co] = both;
        csum_n = it.n;
        csum = 0.f;
      }
      float wbias = 0.f;                           // added once, on the lo row
      if (part && p.bias && co < cout) wbias = __ldg(p.bias + co);
      // where the channel lands (PixelShuffle(2) folded into the store: co' = q*(cout/4) + c)
      int oc = co, qy = 0, qx = 0;
      if (sh == 2) {
        const int q = co < cout ? co / cr : 0;
        oc = co - q * cr;
        qy = q >> 1;
        qx = q & 1;
      }
      const bool ch_ok = co < cout_st;
      const int yb = it.y0 + phi;                   // first output row of this lane's phase
      const int ny = Ho > yb ? (Ho - yb + NPHS - 1) / NPHS : 0, nx = Wo - it.x0;   // valid tile rows / columns
      const int64_t pix0 = ((int64_t)it.n * oH + (yb * sh + qy)) * oW + (it.x0 * sh + qx);
      float* const o0 = planar ? p.out + (((int64_t)it.n * cout + co) * Ho + yb) * Wo + it.x0 : p.out + pix0 * p.out_ld + oc;
      const float* const m0 = post != TDVC_POST_NONE ? p.mul + pix0 * p.mul_ld + oc : nullptr;
      const float* const r10 = p.res1 ? p.res1 + pix0 * p.res1_ld + oc : nullptr;
      const float* const r20 = p.res2 ? p.res2 + pix0 * p.res2_ld + oc : nullptr;
      const bool side = m0 || r10 || r20;
      uint32_t r[16];
      tmem_ld16(lane_addr + (uint32_t)(sa * NPX + chunk0 * 16), r);
      // residual and GDN-multiplier rows: their DRAM latency (~1.5 us under load) is longer than a chunk, so they are pulled
      // into L2 two chunks (4 tile rows) ahead with prefetch.global.L2 - no registers held.  Lane cl < 8 covers tile column
      // cl of its part's row: the 64 bytes of this quadrant's 16 channels lie in one 128-byte line.
      auto l2_prefetch_row = [&](int ty) {
        if (cl < 8 && cl < nx && ty < ny && ch_ok) {
          if (m0) prefetch_l2(m0 + ty * m_rs + cl * m_xs);
          if (r10) prefetch_l2(r10 + ty * r1_rs + cl * r1_xs);
          if (r20) prefetch_l2(r20 + ty * r2_rs + cl * r2_xs);
        }
      };
      // GDN multiplier and first residual of the row a lane stores next, held in registers one chunk ahead
      float mv[8], ra[8];
      auto side_load = [&](int ty) {
        if (ch_ok && ty < ny && nx >= 8) {
          if (m0) {
            const float* mp = m0 + ty * m_rs;
#pragma unroll
            for (int k = 0; k < 8; ++k, mp += m_xs) mv[k] = __ldg(mp);
          }
          if (r10) {
            const float* rp = r10 + ty * r1_rs;
#pragma unroll
            for (int k = 0; k < 8; ++k, rp += r1_xs) ra[k] = __ldg(rp);
          }
        }
      };
      if (side) {
        l2_prefetch_row(2 * chunk0 + part);
        l2_prefetch_row(2 * chunk0 + 2 + part);
        side_load(2 * chunk0 + part);
      }
#pragma unroll 1
      for (int c = 0; c < nchunk; ++c) {
        const int ty = 2 * chunk0 + c * 2 + part;
        tmem_ld_wait();
        if (c == nchunk - 1) {  // all TMEM reads of this warp are done: the accumulator stage may be overwritten
          tc_fence_before();
          mbar_arrive(bar(ACC_EMPTY + sa));
        }
        // merge: lane (part 0) keeps columns 0-7 and receives their lo terms; lane (part 1) keeps columns 8-15 and receives
        // their hi terms.  a + b is commutative, so both orders give the same bits as hi + lo.
        float o[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float va = fmaf(__uint_as_float(r[k]), rscale, wbias), vb = fmaf(__uint_as_float(r[8 + k]), rscale, wbias);
          const float got = __shfl_xor_sync(0xffffffffu, part ? va : vb, 16);
          o[k] = (part ? vb : va) + got;
        }
        // next chunk's accumulator columns: the TMEM read overlaps the loads / stores below
        if (c < nchunk - 1) tmem_ld16(lane_addr + (uint32_t)(sa * NPX + chunk0 * 16 + (c + 1) * 16), r);
        if (!(ch_ok && ty < ny)) continue;
        if (side && c < nchunk - 2) l2_prefetch_row(ty + 4);
        if (planar) {  // NCHW planes: the 8 x-adjacent pixels of the tile row are contiguous in this channel's plane
          if (co < cout) {
            if (has_act)
#pragma unroll
              for (int k = 0; k < 8; ++k) o[k] = fminf(fmaxf(o[k], a_neg * o[k]), a_hi);
            float* op = o0 + (int64_t)ty * NPHS * Wo;
            if (planar_vec && nx >= 8) {
              stg256(op, o);   // one full 32-byte sector
            } else {
#pragma unroll
              for (int k = 0; k < 8; ++k)
                if (k < nx) op[k] = o[k];
            }
          }
          continue;
        }
        float* op = o0 + ty * o_rs;
        if (nx >= 8) {   // interior tile columns: no per-pixel predicates, running pointers
          if (side) {
            if (m0) {  // GDN / IGDN: v = mul * rsqrt(v) | mul * sqrt(v)
              if (post == TDVC_POST_IGDN) {
#pragma unroll
                for (int k = 0; k < 8; ++k) o[k] = mv[k] * sqrt_fast(o[k]);
              } else {
#pragma unroll
                for (int k = 0; k < 8; ++k) o[k] = mv[k] * rsqrt_fast(o[k]);
              }
            }
            if (has_act)
#pragma unroll
              for (int k = 0; k < 8; ++k) o[k] = fminf(fmaxf(o[k], a_neg * o[k]), a_hi);
            if (r10) {
#pragma unroll
              for (int k = 0; k < 8; ++k) o[k] += ra[k];
            }
            if (c < nchunk - 1) side_load(ty + 2);   // next chunk's multiplier / residual values: in flight across the stores,
                                            // the TMEM wait and the merge of the next iteration
            if (r20) {
              const float* rp = r20 + ty * r2_rs;
              float rc[8];
#pragma unroll
              for (int k = 0; k < 8; ++k, rp += r2_xs) rc[k] = __ldg(rp);
#pragma unroll
              for (int k = 0; k < 8; ++k) o[k] += rc[k];
            }
          } else {
            // activation a(v) = min(max(v, a_neg * v), a_hi): identity (a_neg 1), ReLU (0), LeakyReLU (slope), clamp01
            if (has_act)
#pragma unroll
              for (int k = 0; k < 8; ++k) o[k] = fminf(fmaxf(o[k], a_neg * o[k]), a_hi);
          }
#pragma unroll
          for (int k = 0; k < 8; ++k, op += o_xs) *op = o[k];
          if (track && co < cout)
#pragma unroll
            for (int k = 0; k < 8; ++k) amax = fmaxf(amax, fabsf(o[k]));
          if (csum_on)
#pragma unroll
            for (int k = 0; k < 8; ++k) csum += o[k];
        } else {         // ragged right edge (Wo % 8 != 0): one pixel at a time
#pragma unroll 1
          for (int k = 0; k < nx; ++k) {
            float v = o[0];
#pragma unroll
            for (int q = 1; q < 8; ++q) v = (k == q) ? o[q] : v;
            if (m0) {
              const float mvk = __ldg(m0 + ty * m_rs + k * m_xs);
              v = mvk * (post == TDVC_POST_IGDN ? sqrtf(v) : rsqrtf(v));
            }
            v = fminf(fmaxf(v, a_neg * v), a_hi);
            if (r10) v += __ldg(r10 + ty * r1_rs + k * r1_xs);
            if (r20) v += __ldg(r20 + ty * r2_rs + k * r2_xs);
            op[k * o_xs] = v;
            if (co < cout) amax = fmaxf(amax, fabsf(v));
            csum += v;
          }
        }
      }
    }
    if (csum_on) {
      const float both = csum + __shfl_xor_sync(0xffffffffu, csum, 16);
      if (csum_n >= 0 && part == 0 && cch < cout)
        p.chan_sum[((int64_t)((blockIdx.x * C::NGRP + grp) * NPHS + phi) * p.N + csum_n) * cout + cch] = both;
    }
    }
    if (track) absmax_commit(p.out_absmax, amax);
  } else if (warp < C::EPIW + C::PRODW) {
    // ===================================================================== producers: fp32 halo -> fp16 hi/lo planes
    // LPP lanes cover the CK channels of a pixel (one float4 each), 32/LPP pixels per warp-wide load; the halo is
    // walked as a flat pixel list s = (k*8 + warp)*PPI + lane/LPP.  Each thread keeps, for its PER_WARP loads, the
    // source-pixel offset in a register table (the same for every item and unit), so an interior tile costs one
    // address instruction per load.  The loads of a unit are split into NBATCH batches and software-pipelined across
    // batches AND units: the loads of batch b+1 (or of the next unit's batch 0) are issued before batch b is converted,
    // so the memory latency is covered by the fp32 -> fp16 hi/lo conversion of the previous batch.
    if constexpr (C::TMAIN) {
    const int pw = warp - C::EPIW;
    if (pw == C::PRODW - 1) {
      // ------------------------------------------------------------- activation loader: tensor-map bulk copies of the halo
      if (elect_one()) {
        int sT = 0, phT = 1;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
          const Item it = decode_item(item, n_jt, tiles_x, tiles_y, flip, C::THO);
          const int iy0 = it.y0 - C::PAD, ix0 = it.x0 - C::PAD;
          for (int u = 0; u < n_units; ++u) {
#pragma unroll 1
            for (int sub = 0; sub < C::NSUB; ++sub) {
              int cc = u * CK + sub * C::SUB, q = 0;   // first channel of the sub-unit in the concatenated input -> source q
#pragma unroll
              for (int k = 0; k < 3; ++k)
                if (q == k && k + 1 < p.n_src && cc >= p.src_c[k]) { cc -= p.src_c[k]; q = k + 1; }
              mbar_wait(bar(TS_EMPTY + sT), phT);
              mbar_expect_tx(bar(TS_FULL + sT), C::TSTG);
              asm volatile(
                  "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                  ::"r"(smem_u32(ts_buf + sT * C::TSTG)), "l"(reinterpret_cast<uint64_t>(&imaps.m[q])), "r"(cc), "r"(ix0), "r"(iy0),
                  "r"(it.n), "r"(bar(TS_FULL + sT))
                  : "memory");
              if (++sT == C::NTS) { sT = 0; phT ^= 1; }
            }
          }
        }
      }
    } else {
      // ------------------------------------------------------------- converters: staged fp32 box -> fp16 hi / lo operand planes
      constexpr int F4PP = C::SUB / 4;                         // float4 per pixel of a sub-unit
      constexpr int NF4 = C::NHALO * F4PP;                     // float4 per sub-unit
      constexpr int CT = C::CONVW * 32;
      constexpr int PER = (NF4 + CT - 1) / CT;
      const int ctid = threadIdx.x - C::EPIW * 32;
      const float sq = pow2f(-square_shift(p));
      int sX = 0, phX = 1, sT = 0, phT = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        for (int u = 0; u < n_units; ++u) {
          mbar_wait(bar(X_EMPTY + sX), phX);
          uint8_t* const hi = x_buf + sX * C::X_STAGE;
#pragma unroll 1
          for (int sub = 0; sub < C::NSUB; ++sub) {
            mbar_wait(bar(TS_FULL + sT), phT);
            const uint8_t* const stg = ts_buf + sT * C::TSTG;
            float4 v[PER];
#pragma unroll
            for (int k = 0; k < PER; ++k) {
              const int idx = k * CT + ctid;
              v[k] = idx < NF4 ? *reinterpret_cast<const float4*>(stg + idx * 16) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            // the box is in registers: the loader may refill the stage.  The refill is an async-proxy write of memory this
            // thread read through the generic proxy: the proxy fence orders the two (without it the first pixels of a box
            // were occasionally overwritten before the reads had been performed)
            fence_async_smem();
            mbar_arrive(bar(TS_EMPTY + sT));
            if (++sT == C::NTS) { sT = 0; phT ^= 1; }
#pragma unroll
            for (int k = 0; k < PER; ++k) {
              const int idx = k * CT + ctid;
              if (idx < NF4) {
                const int px = idx / F4PP, f = idx - px * F4PP;
                const int g = sub * (C::SUB / 8) + (f >> 1);
                float4 w = v[k];
                if (p.in_square) { w.x = w.x * w.x * sq; w.y = w.y * w.y * sq; w.z = w.z * w.z * sq; w.w = w.w * w.w * sq; }
                uint8_t* dst = hi + (g * C::NPIXP + px) * 16 + (f & 1) * 8;
                uint2 hv, lv;
                if constexpr (C::ONE) {
                  hv.x = pack_h2_sat(w.x, w.y);
                  hv.y = pack_h2_sat(w.z, w.w);
                  *reinterpret_cast<uint2*>(dst) = hv;
                } else {
                  split4(w, hv, lv);
                  *reinterpret_cast<uint2*>(dst) = hv;
                  *reinterpret_cast<uint2*>(dst + C::X_HALF) = lv;
                }
              }
            }
          }
          fence_async_smem();  // generic-proxy stores -> visible to the tensor core (async proxy)
          mbar_arrive(bar(X_FULL + sX));
          if (++sX == C::NX) { sX = 0; phX ^= 1; }
        }
      }
    }
    } else {
    using P = ProdCfg<C, CK>;
    const int pw = warp - C::EPIW;
    const int fi = lane % P::LPP, psub = lane / P::LPP;
    ProdThread th;
    th.pw = pw;
    th.psub = psub;
    th.vmask = 0;
    int tab[P::PER_WARP];   // non-PLANES: source pixel offset hy*STEP*W + hx*STEP;  PLANES: hy << 8 | hx
#pragma unroll
    for (int k = 0; k < P::PER_WARP; ++k) {
      const int sidx = (k * C::PRODW + pw) * P::PPI + psub;
      const int hy = sidx / C::IW, hx = sidx - hy * C::IW;
      tab[k] = C::PLANES ? ((hy << 8) | hx) : (hy * C::STEP * p.W + hx * C::STEP);
      if (sidx < C::NHALO) th.vmask |= 1u << k;
    }
    th.lane_smem = (uint32_t)(((fi >> 1) * C::NPIXP) * 16 + (fi & 1) * 8 + (pw * P::PPI + psub) * 16);
    th.sq = pow2f(-square_shift(p));

    // unit context: where the lane's 4 channels of unit (item, u) come from and which stage they go to
    int item = blockIdx.x, u = 0, sX = 0, phX = 1;
    auto setup = [&](ProdUnit& c) {
      const Item it = decode_item(item, n_jt, tiles_x, tiles_y, flip, C::THO);
      c.iy0 = it.y0 * S - C::PAD;
      c.ix0 = it.x0 * S - C::PAD;
      const float* sp = nullptr;
      int sld = 0;
      int cc = u * CK + fi * 4;   // index of this lane's first channel in the concatenated input
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (q < p.n_src && sp == nullptr) {
          if (cc < p.src_c[q]) { sp = p.src[q] + cc; sld = p.src_ld[q]; }
          else cc -= p.src_c[q];
        }
      }
      c.sld = sld;
      // pointer to this lane's channels of the halo origin pixel (may lie outside the image: only offsets are added)
      c.org = sp ? sp + (((int64_t)it.n * p.H + c.iy0) * p.W + c.ix0) * sld : nullptr;
      c.fast = sp != nullptr && c.iy0 >= 0 && c.ix0 >= 0 && c.iy0 + (C::IH - 1) * C::STEP < p.H &&
               c.ix0 + (C::IW - 1) * C::STEP < p.W;
      mbar_wait(bar(X_EMPTY + sX), phX);
      c.stage = sX;
      c.hi = x_buf + sX * C::X_STAGE + th.lane_smem;
      if (++sX == C::NX) { sX = 0; phX ^= 1; }
    };
    auto advance = [&]() {  // next (item, u) of this CTA; false when there is none
      if (++u == n_units) { u = 0; item += gridDim.x; }
      return item < n_items;
    };

    if constexpr (C::STAGED) {
      // ---- 1x1: global -> (cp.async) -> private staging slots -> fp16 hi/lo operand planes, NSTG units ahead
      static_assert(P::PER_WARP == C::LOADS_PER_THREAD, "staging size");
      const int ptid = threadIdx.x - C::EPIW * 32;
      uint8_t* const my_stg = stg_buf + ptid * 16;   // slot (d, k) at + (d * PER_WARP + k) * kProdThreads * 16
      int is_item = blockIdx.x, is_u = 0;            // next unit to issue
      auto issue = [&](int slot) {
        if (is_item < n_items) {
          const Item it = decode_item(is_item, n_jt, tiles_x, tiles_y, flip, C::THO);
          const float* sp = nullptr;
          int sld = 0;
          int cc = is_u * CK + fi * 4;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (q < p.n_src && sp == nullptr) {
              if (cc < p.src_c[q]) { sp = p.src[q] + cc; sld = p.src_ld[q]; }
              else cc -= p.src_c[q];
            }
          }
          const float* org = sp ? sp + (((int64_t)it.n * p.H + it.y0) * p.W + it.x0) * sld : p.src[0];
          const int ny = p.H - it.y0, nx = p.W - it.x0;
#pragma unroll
          for (int k = 0; k < P::PER_WARP; ++k) {
            const int sidx = (k * C::PRODW + pw) * P::PPI + psub;   // tile pixel: row sidx / 8, column sidx % 8
            if (sidx < C::NHALO) {
              const int hy = sidx >> 3, hx = sidx & 7;
              const bool ok = sp != nullptr && hy < ny && hx < nx;
              const float* g = ok ? org + ((int64_t)hy * p.W + hx) * sld : p.src[0];
              asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(my_stg + (slot * P::PER_WARP + k) * (C::PRODW * 32 * 16))),
                           "l"(g), "r"(ok ? 16 : 0)
                           : "memory");
            }
          }
          if (++is_u == n_units) { is_u = 0; is_item += gridDim.x; }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");   // (possibly empty: keeps the group count uniform)
      };
#pragma unroll
      for (int d = 0; d < C::NSTG; ++d) issue(d);
      int cv_item = blockIdx.x, cv_u = 0, slot = 0;
      while (cv_item < n_items) {
        asm volatile("cp.async.wait_group %0;" ::"n"(C::NSTG - 1) : "memory");
        mbar_wait(bar(X_EMPTY + sX), phX);
        uint8_t* const hi = x_buf + sX * C::X_STAGE + th.lane_smem;
#pragma unroll
        for (int k = 0; k < P::PER_WARP; ++k) {
          if ((th.vmask >> k) & 1) {
            float4 v = *reinterpret_cast<const float4*>(my_stg + (slot * P::PER_WARP + k) * (C::PRODW * 32 * 16));
            if (p.in_square) { v.x = v.x * v.x * th.sq; v.y = v.y * v.y * th.sq; v.z = v.z * v.z * th.sq; v.w = v.w * v.w * th.sq; }
            uint2 hv, lv;
            uint8_t* dst = hi + k * (C::PRODW * P::PPI * 16);
            if constexpr (C::ONE) {
              hv.x = pack_h2_sat(v.x, v.y);
              hv.y = pack_h2_sat(v.z, v.w);
              *reinterpret_cast<uint2*>(dst) = hv;
            } else {
              split4(v, hv, lv);
              *reinterpret_cast<uint2*>(dst) = hv;
              *reinterpret_cast<uint2*>(dst + C::X_HALF) = lv;
            }
          }
        }
        fence_async_smem();  // generic-proxy stores -> visible to the tensor core (async proxy)
        mbar_arrive(bar(X_FULL + sX));
        if (++sX == C::NX) { sX = 0; phX ^= 1; }
        issue(slot);         // refill the slot just consumed with the unit NSTG ahead
        if (++slot == C::NSTG) slot = 0;
        if (++cv_u == n_units) { cv_u = 0; cv_item += gridDim.x; }
      }
    } else {
    float4 va[P::BATCH], vb[P::BATCH];
    ProdUnit cur, nxt;
    bool have = item < n_items;
    if (have) {
      setup(cur);
      prod_issue<C, P, 0>(p, th, tab, cur, va);
    }
    while (have) {
      bool have_next = false;
#pragma unroll
      for (int b = 0; b < P::NBATCH; b += 2) {
        prod_issue_rt<C, P>(p, th, tab, cur, vb, b + 1);
        prod_convert_rt<C, P>(p, th, tab, cur, va, b);
        if (b + 2 < P::NBATCH) {
          prod_issue_rt<C, P>(p, th, tab, cur, va, b + 2);
        } else {
          have_next = advance();
          if (have_next) {
            setup(nxt);
            prod_issue<C, P, 0>(p, th, tab, nxt, va);
          }
        }
        prod_convert_rt<C, P>(p, th, tab, cur, vb, b + 1);
      }
      fence_async_smem();  // generic-proxy stores -> visible to the tensor core (async proxy)
      mbar_arrive(bar(X_FULL + cur.stage));
      cur = nxt;
      have = have_next;
    }
    }
    }
  } else if (warp == C::EPIW + C::PRODW) {
    // ===================================================================== MMA issuer (one elected thread)
    // Descriptors are built once; every MMA only adds a compile-time offset (16-byte units) to their low word.
    if (elect_one()) {
      constexpr uint32_t IDESC = instr_desc(NPX);
      constexpr uint32_t KX = 2 * C::LBO_X / 16, KW = 2 * C::LBO_W / 16;   // k-step advance of the two operands
      const uint64_t xdesc0 = smem_desc(smem_u32(x_buf), C::LBO_X, C::SBO_X);
      const uint64_t wdesc0 = smem_desc(smem_u32(w_buf), C::LBO_W, C::SBO_W);
      int sX = 0, phX = 0, sW = 0, phW = 0, acc_it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++acc_it) {
        const int sa = acc_it & 1;
        mbar_wait(bar(ACC_EMPTY + sa), ((acc_it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d = tmem_base + (uint32_t)(sa * NPX);
        for (int u = 0; u < n_units; ++u) {
          mbar_wait(bar(X_FULL + sX), phX);
          tc_fence_after();
          const uint64_t x_hi = desc_add(xdesc0, (uint32_t)(sX * (C::X_STAGE / 16)));
#pragma unroll 1
          for (int ky = 0; ky < C::KY; ++ky) {
            // first tap of the kernel row: tap_slot(ky, 0); the other taps of the row are compile-time offsets from it
            const uint32_t row0 = C::PLANES ? (uint32_t)(((ky & 1) * 2 * C::PH + (ky >> 1)) * C::PW) : (uint32_t)(ky * C::IW);
            const uint64_t x_row = desc_add(x_hi, row0);
#pragma unroll
            for (int kx = 0; kx < KS; ++kx) {
              mbar_wait(bar(W_FULL + sW), phW);
              tc_fence_after();
              const uint64_t wd = desc_add(wdesc0, (uint32_t)(sW * (C::W_BLOCK / 16)));
              const uint32_t xo = (uint32_t)(C::tap_slot(0, kx) - C::tap_slot(0, 0));
#pragma unroll
              for (int s = 0; s < C::KSTEPS; ++s) {
                const uint64_t wk = desc_add(wd, s * KW);
                const uint64_t xh = desc_add(x_row, xo + s * KX);
                const uint64_t xl = desc_add(x_row, xo + s * KX + C::X_HALF / 16);
                if (kx == 0 && s == 0) tc_mma(d, wk, xh, IDESC, (uint32_t)((u | ky) != 0));
                else tc_mma(d, wk, xh, IDESC, 1u);
                if constexpr (PR == 0) tc_mma(d, wk, xl, IDESC, 1u);
                if constexpr (SPLIT == 1) tc_mma(d, desc_add(wk, C::W_HALF / 16), xh, IDESC, 1u);   // W_lo * x_hi
              }
              tc_commit(bar(W_EMPTY + sW));
              if (++sW == C::NW) { sW = 0; phW ^= 1; }
            }
          }
          tc_commit(bar(X_EMPTY + sX));
          if (++sX == C::NX) { sX = 0; phX ^= 1; }
        }
        tc_commit(bar(ACC_FULL + sa));
      }
    }
  } else {
    // ===================================================================== weight loader (1-D bulk TMA)
    if (elect_one()) {
      const uint8_t* wb = static_cast<const uint8_t*>(p.weight_f16);
      const uint32_t w0 = smem_u32(w_buf);
      int sW = 0, phW = 1;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int jt = (flip >= 0 ? flip - item : item) % n_jt;
        const uint8_t* src = wb + ((int64_t)jt * n_units * C::TAPS) * C::W_BLOCK;
        for (int ut = 0; ut < n_units * C::TAPS; ++ut, src += C::W_BLOCK) {
          mbar_wait(bar(W_EMPTY + sW), phW);
          mbar_expect_tx(bar(W_FULL + sW), C::W_BLOCK);
          bulk_g2s(w0 + sW * C::W_BLOCK, src, C::W_BLOCK, bar(W_FULL + sW));
          if (++sW == C::NW) { sW = 0; phW ^= 1; }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == C::EPIW + C::PRODW) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS)
                 : "memory");
  }
}

// fp32 packed weights [T][cin_pad][cout_pad] -> per (cout tile of 64, unit, tap) fp16 block [128 rows][CK] in the
// canonical K-major no-swizzle layout [(row/8)][(k/8)][row%8][k%8].  Row order = TMEM lane order of the accumulator:
// lane quadrant q (rows 32q..32q+31) = channels 16q..16q+15: rows 32q+0..15 their hi terms, rows 32q+16..31 their lo * 2^12.
// split = 1: per (cout tile of 128, unit, tap) two such blocks, W_hi then W_lo, row = channel, both scaled by 2^w_shift.
// nph > 1 (row phases, see Cfg): T = (ks + nph - 1) * ks packed taps; the item's 64 row pairs are 64 / nph channels x nph phases
// and the rows of phase f take tap (ky' - f, kx) of the ks x ks kernel (zero where that leaves the kernel).
// pr = 1 (one product): per (cout tile, unit, tap) ONE block fp16(w * 2^w_shift); row = phase * (128 / nph) + channel.
__global__ void pack_f16_kernel(const float* __restrict__ w, __half* __restrict__ out, int T, int cin, int cin_pad,
                                 int cout, int cout_pad, int CK, int n_units, int n_jt, int split, float scale, int ks, int nph,
                                 int pr) {
  const int64_t per_block = (int64_t)(split ? 2 : 1) * 128 * CK;
  const int64_t total = (int64_t)n_jt * n_units * T * per_block;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i;
    const int k8 = (int)(r % 8); r /= 8;
    const int n8 = (int)(r % 8); r /= 8;
    const int kc = (int)(r % (CK / 8)); r /= (CK / 8);
    const int ng = (int)(r % 16); r /= 16;
    int which = 0;
    if (split) { which = (int)(r % 2); r /= 2; }
    const int tap = (int)(r % T); r /= T;
    const int u = (int)(r % n_units);
    const int jt = (int)(r / n_units);
    const int row = ng * 8 + n8, k = kc * 8 + k8;
    const int quad = row >> 5;
    const int ci = u * CK + k;
    const bool direct = split || pr;
    const int ntt = (direct ? 128 : NT) / nph;      // channels per item
    const int vc = direct ? row : quad * 16 + (row & 15);          // (phase, channel) pair of this row
    const int phase = vc / ntt;
    const int co = jt * ntt + (vc - phase * ntt);
    const bool lo = direct ? which == 1 : (row & 16) != 0;
    int src_tap = tap;
    if (nph > 1) {
      const int ky = tap / ks - phase, kx = tap % ks;
      src_tap = (ky >= 0 && ky < ks) ? ky * ks + kx : -1;
    }
    float v = 0.f;
    if (src_tap >= 0 && ci < cin_pad && co < cout_pad && ci < cin && co < cout)
      v = w[((int64_t)src_tap * cin_pad + ci) * cout_pad + co] * scale;
    v = fminf(fmaxf(v, -65504.f), 65504.f);
    const __half hi = __float2half_rn(v);
    out[i] = lo ? __float2half_rn((v - __half2float(hi)) * (direct ? 1.f : kLoScale)) : hi;
  }
}

struct Choice {
  int ks, ck, s, split, nph, pr;
};

// Every KxK (K in 1,3,5,7; pad K/2) stride-1 convolution and the stride-2 3x3 / 1x1 ones have a tensor-core path;
// input channels are processed in chunks of ck; output channels in tiles of 64 (4-product scheme) or, for layers
// with >= 96 output channels (a multiple of 4), in tiles of 128 with the 3-product split scheme.
static bool choose(const TdvcConvParams& p, Choice* c) {
  if (p.kh != p.kw || p.pad != p.kh / 2 || p.cin < 4 || p.cout < 1) return false;
  if (p.products == 1) {
    // one product: stride-1 3x3 (<= 64 output channels: 64 channels x 2 row phases per item) and 1x1 layers with >= 32 inputs
    if (p.stride != 1 || p.cin < 32) return false;
    if (p.kh == 3) { *c = {3, 32, 1, 0, p.cout <= 64 ? 2 : 1, 1}; return true; }
    if (p.kh == 1) { *c = {1, 32, 1, 0, 1, 1}; return true; }
    return false;
  }
  if (p.products != 0) return false;
  // (1x1 layers included: with the slab epilogue they were faster on 64-channel tiles, with the direct epilogue the
  // 128-channel tiles win - GDN 128->128 @512x960 0.230 -> 0.179 ms, 1x1 + residual 0.219 -> 0.135 ms)
  const int split = (p.cout >= 96 && (p.cout & 3) == 0 && p.kh != 7) ? 1 : 0;
  if (p.stride == 2) {
    if (p.kh == 3) { *c = {3, 16, 2, split, 1, 0}; return true; }
    if (p.kh == 1) { *c = {1, 32, 2, split, 1, 0}; return p.cin >= 32; }
    return false;
  }
  if (p.stride != 1) return false;
  const int ck = p.cin > 16 ? 32 : 16;
  if (p.kh == 3) { *c = {3, ck, 1, ck == 32 ? split : 0, 1, 0}; return true; }
  if (p.kh == 7) {
    // <= 32 output channels (SPyNet 8->32, 64->32, 32->16): two row phases fill the 64 row pairs of the M side
    if (p.cout <= 32) { *c = {7, 16, 1, 0, 2, 0}; return true; }
    *c = {7, ck, 1, 0, 1, 0};
    return true;
  }
  if (p.kh == 1 || p.kh == 5) { *c = {p.kh, 32, 1, split, 1, 0}; return p.cin >= 32; }
  return false;
}

// cuTensorMapEncodeTiled is looked up through the runtime at first use: linking libcuda.so directly would make the library
// unloadable on a host without a driver (the CPU-only build / symbol checks load it there)
using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeFn tensor_map_encoder() {
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

// can the sources of *p be walked by tensor maps in sub-units of `sub` channels?  (every source but the last must end on a
// sub-unit boundary; TMA needs 16-byte aligned bases and strides)
static bool tma_sources_ok(const TdvcConvParams& p, int ck) {
  if (getenv("TDVC_B200_CONV_LDG") != nullptr) return false;   // developer A/B switch: producers load with LDG
  for (int q = 0; q < p.n_src; ++q) {
    if ((reinterpret_cast<uintptr_t>(p.src[q]) & 15) != 0 || (p.src_ld[q] & 3) != 0) return false;
    if (q + 1 < p.n_src && p.src_c[q] % ck != 0) return false;
  }
  return true;
}

template <int KS, int CK, int S, int SPLIT, int NPH = 1, int PLN = 0, int PR = 0, int TMA = 0>
static int launch(const TdvcConvParams& p, cudaStream_t st, int* rows_only) {
  using C = Cfg<KS, CK, S, SPLIT, NPH, PLN, PR, TMA>;
  if (rows_only != nullptr) {   // query: rows of the chan_sum buffer this launch would write (CTAs x epilogue column groups x phases)
    const int64_t it = (int64_t)p.N * cdiv(p.Wo, TW) * cdiv(p.Ho, C::THO) * cdiv(p.cout, C::NTT);
    *rows_only = (p.cout <= C::NTT && !p.out_planar && p.shuffle == 0) ? (int)(it < kNumSMs ? it : kNumSMs) * C::NGRP * NPH : 0;
    return TDVC_OK;
  }
  static int smem_done[kMaxDevices] = {0};
  if (int rc = ensure_dynamic_smem(conv_tc_kernel<KS, CK, S, SPLIT, NPH, PLN, PR, TMA>, C::SMEM, smem_done, "conv_tc")) return rc;
  if (C::DIRECT) TDVC_REQUIRE(p.w_shift >= -100 && p.w_shift <= 100, "conv_tc: w_shift %d out of range", p.w_shift);
  TDVC_REQUIRE(NPH == 1 || !p.out_planar, "conv_tc: planar output is not available for row-phase items");
  TDVC_REQUIRE(p.chan_sum == nullptr || (p.cout <= C::NTT && !p.out_planar && p.shuffle == 0),
               "conv_tc: chan_sum needs a single output-channel tile (cout %d <= %d), NHWC output, no pixel shuffle", p.cout, C::NTT);
  const int tiles_x = cdiv(p.Wo, TW), tiles_y = cdiv(p.Ho, C::THO);
  const int n_jt = cdiv(p.cout, C::NTT), n_units = cdiv(p.cin, CK);
  const int64_t items = (int64_t)p.N * tiles_x * tiles_y * n_jt;
  TDVC_REQUIRE(items < (1ll << 31), "conv_tc: too many work items");
  const int grid = (int)(items < kNumSMs ? items : kNumSMs);
  std::conditional_t<TMA != 0, InMaps, NoMap> imaps;
  if constexpr (TMA != 0) {
    // one tensor map per source over (channel, x, y, image) fp32; box = SUB channels x IW x IH x 1 image.  Coordinates outside
    // the tensor (image border, channels past the source) are zero-filled by the hardware
    EncodeFn encode = tensor_map_encoder();
    if (encode == nullptr) {
      set_error("conv_tc: cuTensorMapEncodeTiled is not available from this driver");
      return TDVC_ECUDA;
    }
    for (int q = 0; q < p.n_src; ++q) {
      const cuuint64_t ld = (cuuint64_t)p.src_ld[q];
      const cuuint64_t dims[4] = {(cuuint64_t)p.src_c[q], (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.N};
      const cuuint64_t strides[3] = {ld * 4, (cuuint64_t)p.W * ld * 4, (cuuint64_t)p.H * p.W * ld * 4};
      const cuuint32_t box[4] = {(cuuint32_t)C::SUB, (cuuint32_t)C::IW, (cuuint32_t)C::IH, 1}, estr[4] = {1, 1, 1, 1};
      const CUresult r = encode(&imaps.m[q], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(p.src[q]), dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) {
        set_error("conv_tc: cuTensorMapEncodeTiled failed for source %d (%d)", q, (int)r);
        return TDVC_ECUDA;
      }
    }
    for (int q = p.n_src; q < 4; ++q) imaps.m[q] = imaps.m[0];
  }
  if constexpr (PLN != 0) {
    // tensor map of the planar output (N, cout, Ho, Wo) fp32: box = 8 px x 2 rows x 32 channels x 1 image
    CUtensorMap tm;
    const cuuint64_t dims[4] = {(cuuint64_t)p.Wo, (cuuint64_t)p.Ho, (cuuint64_t)p.cout, (cuuint64_t)p.N};
    const cuuint64_t strides[3] = {(cuuint64_t)p.Wo * 4, (cuuint64_t)p.Ho * p.Wo * 4, (cuuint64_t)p.cout * p.Ho * p.Wo * 4};
    const cuuint32_t box[4] = {8, 2, 32, 1}, estr[4] = {1, 1, 1, 1};
    EncodeFn encode = tensor_map_encoder();
    if (encode == nullptr) {
      set_error("conv_tc: cuTensorMapEncodeTiled is not available from this driver");
      return TDVC_ECUDA;
    }
    const CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, p.out, dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("conv_tc: cuTensorMapEncodeTiled failed (%d)", (int)r);
      return TDVC_ECUDA;
    }
    conv_tc_kernel<KS, CK, S, SPLIT, NPH, PLN, PR, TMA><<<grid, kThreads, C::SMEM, st>>>(p, tiles_x, tiles_y, n_jt, n_units, (int)items, tm, imaps);
  } else {
    conv_tc_kernel<KS, CK, S, SPLIT, NPH, PLN, PR, TMA><<<grid, kThreads, C::SMEM, st>>>(p, tiles_x, tiles_y, n_jt, n_units, (int)items, NoMap{}, imaps);
  }
  TDVC_CHECK_LAUNCH("conv_tc");
  return TDVC_OK;
}

}  // namespace tc

int conv2d_tc_supported(const TdvcConvParams& p) {
  tc::Choice c;
  if (p.weight_f16 == nullptr) return 0;
  if (!tc::choose(p, &c)) return 0;
  for (int s = 0; s < p.n_src; ++s)
    if (p.src_c[s] % 4 != 0 || p.src_ld[s] % 4 != 0 || (reinterpret_cast<uintptr_t>(p.src[s]) & 15) != 0) return 0;
  if ((reinterpret_cast<uintptr_t>(p.weight_f16) & 15) != 0) return 0;
  return 1;
}

int conv2d_tc(const TdvcConvParams& p, cudaStream_t st, int* rows_only) {
  tc::Choice c;
  if (!tc::choose(p, &c)) {
    set_error("conv_tc: unsupported shape");
    return TDVC_EINVAL;
  }
  // TMA-fed activations (stride-1 KxK layers).  Measured per shape (tools/conv_bench.py, tools/ab_frame.py): the
  // tensor-map path wins where the producers were the limit - the one-product layers and the 4-channel image layers - and costs
  // 1-2 % on the MMA-bound fp32-class layers (the staged box crosses shared memory twice), so it is the default for the former
  // only.  TDVC_B200_CONV_TMA=1 / TDVC_B200_CONV_LDG=1 force it on / off for every eligible layer (developer A/B).
  const bool tma_ok = c.s == 1 && c.ks != 1 && tc::tma_sources_ok(p, c.ck);
  const bool tma = tma_ok && (getenv("TDVC_B200_CONV_TMA") != nullptr || c.pr == 1 || (c.ks == 3 && c.ck == 16));
  if (c.pr == 1) {
    if (c.ks == 3 && c.nph == 2) return tma ? tc::launch<3, 32, 1, 0, 2, 0, 1, 1>(p, st, rows_only) : tc::launch<3, 32, 1, 0, 2, 0, 1>(p, st, rows_only);
    if (c.ks == 3) return tma ? tc::launch<3, 32, 1, 0, 1, 0, 1, 1>(p, st, rows_only) : tc::launch<3, 32, 1, 0, 1, 0, 1>(p, st, rows_only);
    if (c.ks == 1) return tc::launch<1, 32, 1, 0, 1, 0, 1>(p, st, rows_only);
  } else if (c.split) {
    if (c.s == 2 && c.ks == 3) return tc::launch<3, 16, 2, 1>(p, st, rows_only);
    if (c.s == 2 && c.ks == 1) return tc::launch<1, 32, 2, 1>(p, st, rows_only);
    // DCN offset / mask head: planar output through TMA tensor stores (needs 16-byte aligned planes and row pitch)
    if (c.ks == 3 && p.out_planar && (p.Wo & 3) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0)
      return tma ? tc::launch<3, 32, 1, 1, 1, 1, 0, 1>(p, st, rows_only) : tc::launch<3, 32, 1, 1, 1, 1>(p, st, rows_only);
    if (c.ks == 3) return tma ? tc::launch<3, 32, 1, 1, 1, 0, 0, 1>(p, st, rows_only) : tc::launch<3, 32, 1, 1>(p, st, rows_only);
    if (c.ks == 1) return tc::launch<1, 32, 1, 1>(p, st, rows_only);
    if (c.ks == 5) return tma ? tc::launch<5, 32, 1, 1, 1, 0, 0, 1>(p, st, rows_only) : tc::launch<5, 32, 1, 1>(p, st, rows_only);
  } else {
    if (c.s == 2 && c.ks == 3) return tc::launch<3, 16, 2, 0>(p, st, rows_only);
    if (c.s == 2 && c.ks == 1) return tc::launch<1, 32, 2, 0>(p, st, rows_only);
    if (c.ks == 3 && c.ck == 32) return tma ? tc::launch<3, 32, 1, 0, 1, 0, 0, 1>(p, st, rows_only) : tc::launch<3, 32, 1, 0>(p, st, rows_only);
    if (c.ks == 3 && c.ck == 16) return tma ? tc::launch<3, 16, 1, 0, 1, 0, 0, 1>(p, st, rows_only) : tc::launch<3, 16, 1, 0>(p, st, rows_only);
    if (c.ks == 1) return tc::launch<1, 32, 1, 0>(p, st, rows_only);
    if (c.ks == 5) return tma ? tc::launch<5, 32, 1, 0, 1, 0, 0, 1>(p, st, rows_only) : tc::launch<5, 32, 1, 0>(p, st, rows_only);
    if (c.ks == 7 && c.nph == 2) return tma ? tc::launch<7, 16, 1, 0, 2, 0, 0, 1>(p, st, rows_only) : tc::launch<7, 16, 1, 0, 2>(p, st, rows_only);
    if (c.ks == 7 && c.ck == 32) return tma ? tc::launch<7, 32, 1, 0, 1, 0, 0, 1>(p, st, rows_only) : tc::launch<7, 32, 1, 0>(p, st, rows_only);
    if (c.ks == 7 && c.ck == 16) return tma ? tc::launch<7, 16, 1, 0, 1, 0, 0, 1>(p, st, rows_only) : tc::launch<7, 16, 1, 0>(p, st, rows_only);
  }
  set_error("conv_tc: no instantiation for ks=%d ck=%d stride=%d split=%d", c.ks, c.ck, c.s, c.split);
  return TDVC_EINVAL;
}

}  // namespace tdvc

using namespace tdvc;

static size_t f16_elems(const TdvcConvParams* p, const tc::Choice& c) {
  const int ntt = ((c.split || c.pr) ? 128 : tc::NT) / c.nph;
  const int n_jt = cdiv(p->cout, ntt), n_units = cdiv(p->cin, c.ck);
  return (size_t)n_jt * n_units * (c.ks + c.nph - 1) * c.ks * (c.split ? 2 : 1) * 128 * c.ck;
}

extern "C" size_t tdvc_conv2d_f16_bytes(const TdvcConvParams* p) {
  tc::Choice c;
  if (p == nullptr || !tc::choose(*p, &c)) return 0;
  return f16_elems(p, c) * sizeof(__half);
}

extern "C" int tdvc_conv2d_f16_is_split(const TdvcConvParams* p) {
  tc::Choice c;
  if (p == nullptr || !tc::choose(*p, &c)) return 0;
  return c.split || c.pr;
}

extern "C" int tdvc_conv2d_pack_f16(const TdvcConvParams* p, void* out, void* stream) {
  tc::Choice c;
  TDVC_REQUIRE(p && out && p->weight, "conv2d_pack_f16: null pointer");
  TDVC_REQUIRE(tc::choose(*p, &c), "conv2d_pack_f16: shape has no tcgen05 path");
  const bool direct = c.split || c.pr;
  TDVC_REQUIRE(!direct || (p->w_shift >= -100 && p->w_shift <= 100), "conv2d_pack_f16: w_shift %d out of range", p->w_shift);
  const int ntt = (direct ? 128 : tc::NT) / c.nph;
  const int n_jt = cdiv(p->cout, ntt), n_units = cdiv(p->cin, c.ck);
  const int64_t total = (int64_t)f16_elems(p, c);
  int grid = cdiv(total, 256);
  if (grid > kNumSMs * 8) grid = kNumSMs * 8;
  const float scale = direct ? ldexpf(1.f, p->w_shift) : 1.f;
  tc::pack_f16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p->weight, static_cast<__half*>(out), (c.ks + c.nph - 1) * c.ks, p->cin,
                                                             p->cin_pad, p->cout, p->cout_pad, c.ck, n_units, n_jt, c.split, scale,
                                                             c.ks, c.nph, c.pr);
  TDVC_CHECK_LAUNCH("conv2d_pack_f16");
  return TDVC_OK;
}
