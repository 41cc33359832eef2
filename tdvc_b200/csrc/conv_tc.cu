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
// Work decomposition (persistent, one CTA per SM, 576 threads = 18 warps):
//   work item  = 8 x 32 output pixels (N = 256; pixel p = ty*8 + tx) x 64 output channels;
//                K loop = cin chunks of CK channels ("units") x KS*KS taps x CK/16 MMA k-steps.
//   warps 8-15 producers: read the fp32 halo of one unit from global (float4 per lane, 16*CK/4 B contiguous per
//              pixel, all loads of a batch in flight), split into fp16 hi / lo and store it to shared memory in the
//              tcgen05 K-major no-swizzle canonical layout [channel/8][halo pixel][8 channels]: 8 x-adjacent pixels
//              form one 8x16 B core matrix and the 32 tile rows are 32 row groups (SBO = halo pitch), so every tap of
//              the convolution is the SAME buffer read through a descriptor whose start address is shifted by
//              (ky*pitch + kx)*16 B.  No im2col copy.  Stride 2: the halo is de-interleaved into 4 parity planes.
//   warp 17    streams pre-packed fp16 weight blocks [128 rows][CK] (one per unit x tap) with cp.async.bulk
//              (1-D TMA) into a ring, completion on mbarriers.
//   warp 16    one elected thread issues per k-step  D[128][256] += W * X_hi^T ; D += W * X_lo^T
//              and tcgen05.commit's the ring slots back to the producers / loader.
//   warps 0-7  epilogue.  TMEM lane = weight row: lane quadrants 0 / 2 hold the hi rows of output channels 0-31 /
//              32-63, quadrants 1 / 3 the matching lo rows.  A lo warp reads its rows (tcgen05.ld, 32 pixels per
//              chunk), rescales and hands them to its hi partner through a double-buffered shared-memory slab
//              (named barrier per pair); the hi warp adds, applies bias / GDN / activation / residuals and stores:
//              lane = channel, so each store instruction writes 128 contiguous bytes of one NHWC pixel.
//              Overlaps the next item's MMAs (two accumulator stages = all 512 TMEM columns).
#include "tc_common.cuh"

namespace tdvc {

namespace tc {

constexpr int kEpiWarps = 8, kProdWarps = 10;                // 20 warps = 5 per SM sub-partition: still 96 registers per thread
constexpr int kMmaWarp = kEpiWarps + kProdWarps, kLoadWarp = kMmaWarp + 1;
constexpr int kThreads = (kEpiWarps + kProdWarps + 2) * 32;  // 640
constexpr int kProdThreads = kProdWarps * 32;
constexpr int TW = 8, TH = 32, NPX = TW * TH;  // output tile = N of one MMA
constexpr int NT = 64;                         // output channels per item (M = 2*NT rows: hi, lo)
constexpr int kStageBytes = 2 * 2 * 16 * 128 * 4;  // epilogue transposition: 2 pixel halves x 2 buffers x [16 px][128 rows] fp32

// SPLIT = 0: item = 64 output channels, weight rows [hi | lo*2^12] of those channels (4 products, 2 MMAs per k-step);
// SPLIT = 1: item = 128 output channels, separate W_hi / W_lo blocks (pre-scaled by 2^w_shift so that W_lo stays in
//            the fp16 normal range) accumulating into ONE row per channel: x_hi*w_hi + x_lo*w_hi + x_hi*w_lo,
//            3 MMAs per k-step for twice the channels (-25 % tensor work) and no hi/lo merge in the epilogue.
template <int KS, int CK, int S, int SPLIT = 0>
struct Cfg {
  static_assert(S == 1 || (S == 2 && (KS == 1 || KS == 3)), "stride");
  static_assert(CK == 16 || CK == 32, "cin chunk");
  static constexpr int PAD = KS / 2;
  // input halo of one 8x32 output tile: IH x IW input pixels, input pixel = origin + h*STEP
  static constexpr int STEP = (KS == 1) ? S : 1;            // 1x1: only every S-th input pixel is touched
  static constexpr int IH = (KS == 1) ? TH : (TH - 1) * S + KS;
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
  static constexpr int X_STAGE = 2 * X_HALF;
  static constexpr int LBO_X = NPIXP * 16, SBO_X = PW * 16;
  static constexpr int NTT = SPLIT ? 128 : 64;       // output channels per item
  static constexpr int W_HALF = 128 * CK * 2;        // one [128 rows][CK] fp16 block
  static constexpr int W_BLOCK = SPLIT ? 2 * W_HALF : W_HALF;   // SPLIT: [W_hi block | W_lo block]
  static constexpr int LBO_W = 128, SBO_W = NCH8 * 128;
  static constexpr int KSTEPS = CK / 16;
  static constexpr int TAPS = KS * KS;
  static constexpr int NW = (CK == 32) ? 4 : (SPLIT ? 4 : 6);
  static constexpr int NX = (3 * X_STAGE + NW * W_BLOCK + kStageBytes + 256 <= 227 * 1024) ? 3 : 2;
  static constexpr int SMEM = NX * X_STAGE + NW * W_BLOCK + kStageBytes + 256;
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
__device__ __forceinline__ Item decode_item(int item, int n_jt, int tiles_x, int tiles_y, int flip) {
  Item it;
  if (flip >= 0) item = flip - item;
  it.jt = item % n_jt;
  int st = item / n_jt;
  it.x0 = (st % tiles_x) * TW;
  st /= tiles_x;
  it.y0 = (st % tiles_y) * TH;
  it.n = st / tiles_y;
  return it;
}

// Rare epilogue path, kept out of line so that the hot loop stays small: ragged channel counts (cout % 4 != 0) or
// unaligned / odd-pitch views, one element at a time.
struct RaggedArgs {   // passed by value (registers): taking the kernel parameter block by reference would move it to local memory
  float* out; const float* mul; const float* res1; const float* res2;
  int out_ld, mul_ld, res1_ld, res2_ld, post, act, shuffle, cout, Ho, Wo;
  float slope;
};
__device__ __noinline__ void epilogue_ragged(RaggedArgs a, int n, int y0, int x0, const float* sb, int tx, int rd_off, int co, int ty0,
                                              int lo_off) {
  const int sh = a.shuffle == 2 ? 2 : 1;
  const int cr = a.cout >> 2;
  for (int h2 = 0; h2 < 2; ++h2) {
    const int y = y0 + ty0 + h2, x = x0 + tx;
    if (y >= a.Ho) continue;
    const float* rb = sb + (h2 * 8 + tx) * 128 + rd_off;
    for (int e = 0; e < 4; ++e) {
      const int ce = co + e;
      if (ce >= a.cout) break;
      int oce = ce;
      int64_t pe = ((int64_t)n * a.Ho + y) * a.Wo + x;
      if (sh == 2) {
        const int q = ce / cr;
        oce = ce - q * cr;
        pe = ((int64_t)n * (2 * a.Ho) + (2 * y + (q >> 1))) * (2 * a.Wo) + (2 * x + (q & 1));
      }
      float o = lo_off ? rb[e] + rb[e + lo_off] : rb[e];
      if (a.post != TDVC_POST_NONE) {
        const float mv = __ldg(a.mul + pe * a.mul_ld + oce);
        o = mv * (a.post == TDVC_POST_IGDN ? sqrtf(o) : rsqrtf(o));
      }
      o = apply_act(o, a.act, a.slope);
      if (a.res1) o += __ldg(a.res1 + pe * a.res1_ld + oce);
      if (a.res2) o += __ldg(a.res2 + pe * a.res2_ld + oce);
      a.out[pe * a.out_ld + oce] = o;
    }
  }
}

// ---- producer helpers -------------------------------------------------------------------------------------------
template <class C, int CK>
struct ProdCfg {
  static constexpr int LPP = CK / 4;                       // lanes (float4) per pixel
  static constexpr int PPI = 32 / LPP;                     // pixels per warp-wide load
  static constexpr int NLD = (C::NHALO + PPI - 1) / PPI;   // warp-wide loads per unit
  static constexpr int PER_WARP = (NLD + kProdWarps - 1) / kProdWarps;
  static constexpr int NBATCH = PER_WARP <= 12 ? 2 : 4;    // even: the two register buffers alternate
  static constexpr int BATCH = (PER_WARP + NBATCH - 1) / NBATCH;
};
struct ProdThread {
  int pw, psub;
  uint32_t vmask, lane_smem;
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
      const int sidx = (kk * kProdWarps + th.pw) * P::PPI + th.psub;
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
    for (int k = 0; k < P::BATCH; ++k) { v[k].x *= v[k].x; v[k].y *= v[k].y; v[k].z *= v[k].z; v[k].w *= v[k].w; }
  }
#pragma unroll
  for (int k = 0; k < P::BATCH; ++k) {
    const int kk = B * P::BATCH + k;
    if (kk < P::PER_WARP) {
      if ((th.vmask >> kk) & 1) {
        uint2 hv, lv;
        split4(v[k], hv, lv);
        uint8_t* dst;
        if (C::PLANES) {
          const int po = tab[kk < P::PER_WARP ? kk : 0];
          dst = c.hi - (th.pw * P::PPI + th.psub) * 16 + C::slot(po >> 8, po & 255) * 16;
        } else {
          dst = c.hi + kk * (kProdWarps * P::PPI * 16);   // flat slot = pixel index
        }
        *reinterpret_cast<uint2*>(dst) = hv;
        *reinterpret_cast<uint2*>(dst + C::X_HALF) = lv;
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

template <int KS, int CK, int S, int SPLIT>
__global__ void __launch_bounds__(kThreads, 1) conv_tc_kernel(const TdvcConvParams p, int tiles_x, int tiles_y, int n_jt,
                                                              int n_units, int n_items) {
  using C = Cfg<KS, CK, S, SPLIT>;
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* x_buf = smem;                                  // NX stages of [hi plane | lo plane]
  uint8_t* w_buf = smem + C::NX * C::X_STAGE;             // NW weight blocks
  float* stage_buf = reinterpret_cast<float*>(w_buf + C::NW * C::W_BLOCK);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(stage_buf) + kStageBytes);
  // barrier indices
  constexpr int X_FULL = 0, X_EMPTY = X_FULL + C::NX, W_FULL = X_EMPTY + C::NX, W_EMPTY = W_FULL + C::NW,
                ACC_FULL = W_EMPTY + C::NW, ACC_EMPTY = ACC_FULL + 2, NBARS = ACC_EMPTY + 2;
  static_assert(NBARS * 8 + 8 <= 256, "barrier area");
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NBARS);
  const uint32_t bar0 = smem_u32(bars);
  auto bar = [&](int i) { return bar0 + 8u * i; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int flip = p.order ? n_items - 1 : -1;

  if (threadIdx.x == 0) {
    for (int i = 0; i < C::NX; ++i) { mbar_init(bar(X_FULL + i), kProdThreads); mbar_init(bar(X_EMPTY + i), 1); }
    for (int i = 0; i < C::NW; ++i) { mbar_init(bar(W_FULL + i), 1); mbar_init(bar(W_EMPTY + i), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(bar(ACC_FULL + i), 1); mbar_init(bar(ACC_EMPTY + i), kEpiWarps * 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
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

  if (warp < kEpiWarps) {
    // ===================================================================== epilogue (8 warps: lane quadrant x pixel half)
    // The 4 warps of a pixel half (128 threads, named barrier 1 + half) move 16 pixels (two tile rows) at a time:
    //   write: warp = TMEM lane quadrant, lane = weight row -> slab[px][row] (conflict-free 128 B per pixel and warp);
    //          the lo rows are rescaled and carry the bias;
    //   read : thread = (tile column tx, 4 consecutive output channels) -> hi + lo as two float4, GDN / activation /
    //          residuals, one 16-byte store per tile row: 16 lanes cover the 256 contiguous bytes of an NHWC pixel.
    if constexpr (SPLIT == 1) {
    // ---- 128 output channels per item, one accumulator row per channel: write = TMEM lane quadrant q -> channels
    //      32q..32q+31 (scaled back by 2^-w_shift, bias added); read = thread (4 channels, tile columns txb and txb+4).
    const int quad = warp & 3, half = warp >> 2;
    float* slab = stage_buf + half * (2 * 16 * 128);
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
    const int t = quad * 32 + lane;
    const int c4 = t & 31, txb = t >> 5;
    const int rd_off = c4 * 4;
    const int Ho = p.Ho, Wo = p.Wo, cout = p.cout, act = p.act, post = p.post;
    const int sh = p.shuffle == 2 ? 2 : 1;
    const int cr = cout >> 2;
    const int oW = Wo * sh, oH = Ho * sh;
    const bool planar = p.out_planar != 0;
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    const bool vec_ok = !planar && (p.out_ld & 3) == 0 && al16(p.out) && (sh == 1 || (cr & 3) == 0) &&
                        (post == TDVC_POST_NONE || ((p.mul_ld & 3) == 0 && al16(p.mul))) &&
                        (!p.res1 || ((p.res1_ld & 3) == 0 && al16(p.res1))) && (!p.res2 || ((p.res2_ld & 3) == 0 && al16(p.res2)));
    const bool planar_vec = planar && (Wo & 7) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 31) == 0;
    const float a_neg = act == TDVC_ACT_NONE ? 1.f : (act == TDVC_ACT_LRELU ? p.slope : 0.f);
    const float a_hi = act == TDVC_ACT_CLAMP01 ? 1.f : __int_as_float(0x7f800000);
    const float unscale = __int_as_float((127 - p.w_shift) << 23);   // 2^-w_shift, exact
    const int o_rs = sh * oW * p.out_ld, m_rs = sh * oW * p.mul_ld, r1_rs = sh * oW * p.res1_ld, r2_rs = sh * oW * p.res2_ld;
    const int o_p4 = 4 * sh * p.out_ld, m_p4 = 4 * sh * p.mul_ld, r1_p4 = 4 * sh * p.res1_ld, r2_p4 = 4 * sh * p.res2_ld;  // 4 tile columns
    int acc_it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++acc_it) {
      const Item it = decode_item(item, n_jt, tiles_x, tiles_y, flip);
      const int sa = acc_it & 1;
      mbar_wait(bar(ACC_FULL + sa), (acc_it >> 1) & 1);
      tc_fence_after();
      float wbias = 0.f;
      {
        const int cw = it.jt * C::NTT + t;
        if (p.bias && cw < cout) wbias = __ldg(p.bias + cw);
      }
      const int co = it.jt * C::NTT + 4 * c4;
      int oc = co, qy = 0, qx = 0;
      if (sh == 2) {
        const int q = co < cout ? co / cr : 0;
        oc = co - q * cr;
        qy = q >> 1;
        qx = q & 1;
      }
      const bool c_ok = co < cout;
      const bool th_vec = vec_ok && co + 4 <= cout;
      const int ny = Ho - it.y0;
      const int nxv = Wo - it.x0 - txb;   // column txb valid iff nxv > 0, column txb + 4 iff nxv > 4
      const int64_t pix0 = ((int64_t)it.n * oH + (it.y0 * sh + qy)) * oW + ((it.x0 + txb) * sh + qx);
      float* const o0 = p.out + pix0 * p.out_ld + oc;
      const float* const m0 = post != TDVC_POST_NONE ? p.mul + pix0 * p.mul_ld + oc : nullptr;
      const float* const r10 = p.res1 ? p.res1 + pix0 * p.res1_ld + oc : nullptr;
      const float* const r20 = p.res2 ? p.res2 + pix0 * p.res2_ld + oc : nullptr;
      uint32_t r[16];
      tmem_ld16(lane_addr + (uint32_t)(sa * NPX + half * 128), r);
      // residual / GDN-multiplier rows are pulled into L2 two chunks ahead (see the 64-channel epilogue below)
      auto l2_prefetch_rows = [&](int tyb) {
        if (!(th_vec && c_ok)) return;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int h2 = k >> 1, g = k & 1;
          if (tyb + h2 < ny && nxv > 4 * g) {
            if (m0) prefetch_l2(m0 + (tyb + h2) * m_rs + g * m_p4);
            if (r10) prefetch_l2(r10 + (tyb + h2) * r1_rs + g * r1_p4);
            if (r20) prefetch_l2(r20 + (tyb + h2) * r2_rs + g * r2_p4);
          }
        }
      };
      if (m0 || r10 || r20) {
        l2_prefetch_rows(half * 16);
        l2_prefetch_rows(half * 16 + 2);
      }
#pragma unroll 1
      for (int c = 0; c < 8; ++c) {
        const int ty0 = half * 16 + c * 2;
        tmem_ld_wait();
        if (c == 7) {
          tc_fence_before();
          mbar_arrive(bar(ACC_EMPTY + sa));
        }
        float* sb = slab + (c & 1) * (16 * 128);
#pragma unroll
        for (int j = 0; j < 16; ++j) sb[j * 128 + t] = fmaf(__uint_as_float(r[j]), unscale, wbias);
        if (c < 7) tmem_ld16(lane_addr + (uint32_t)(sa * NPX + half * 128 + (c + 1) * 16), r);
        asm volatile("bar.sync %0, 128;" ::"r"(1 + half) : "memory");
        if (th_vec) {
          if (!c_ok) continue;
          if (c < 6 && (m0 || r10 || r20)) l2_prefetch_rows(ty0 + 4);
#pragma unroll
          for (int k = 0; k < 4; ++k) {   // (tile row, column group)
            const int h2 = k >> 1, g = k & 1;
            const int ty = ty0 + h2;
            if (ty >= ny || nxv <= 4 * g) continue;
            const float4 sv = *reinterpret_cast<const float4*>(sb + (h2 * 8 + txb + 4 * g) * 128 + rd_off);
            float v[4] = {sv.x, sv.y, sv.z, sv.w};
            if (m0) {
              const float4 m = __ldg(reinterpret_cast<const float4*>(m0 + ty * m_rs + g * m_p4));
              if (post == TDVC_POST_IGDN) { v[0] = m.x * sqrtf(v[0]); v[1] = m.y * sqrtf(v[1]); v[2] = m.z * sqrtf(v[2]); v[3] = m.w * sqrtf(v[3]); }
              else { v[0] = m.x * rsqrtf(v[0]); v[1] = m.y * rsqrtf(v[1]); v[2] = m.z * rsqrtf(v[2]); v[3] = m.w * rsqrtf(v[3]); }
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = fminf(fmaxf(v[e], a_neg * v[e]), a_hi);
            if (r10) {
              const float4 m = __ldg(reinterpret_cast<const float4*>(r10 + ty * r1_rs + g * r1_p4));
              v[0] += m.x; v[1] += m.y; v[2] += m.z; v[3] += m.w;
            }
            if (r20) {
              const float4 m = __ldg(reinterpret_cast<const float4*>(r20 + ty * r2_rs + g * r2_p4));
              v[0] += m.x; v[1] += m.y; v[2] += m.z; v[3] += m.w;
            }
            *reinterpret_cast<float4*>(o0 + ty * o_rs + g * o_p4) = make_float4(v[0], v[1], v[2], v[3]);
          }
        } else if (planar) {  // NCHW planes: thread = channel; 8 x-adjacent pixels of a tile row are contiguous in the plane
          const int cp = it.jt * C::NTT + t;
          if (cp < cout) {
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2) {
              const int ty = ty0 + h2;
              if (ty >= ny) continue;
              const float* rb = sb + h2 * (8 * 128) + t;
              float v[8];
#pragma unroll
              for (int x = 0; x < 8; ++x) {
                const float o = rb[x * 128];
                v[x] = fminf(fmaxf(o, a_neg * o), a_hi);
              }
              float* op = p.out + (((int64_t)it.n * cout + cp) * Ho + (it.y0 + ty)) * Wo + it.x0;
              if (planar_vec && it.x0 + 8 <= Wo) {
                stg256(op, v);   // one full 32-byte sector: two 16-byte stores made L2 read the sector back (partial writes)
              } else {
#pragma unroll
                for (int x = 0; x < 8; ++x)
                  if (it.x0 + x < Wo) op[x] = v[x];
              }
            }
          }
        } else if (c_ok) {
          const RaggedArgs ra{p.out, p.mul, p.res1, p.res2, p.out_ld, p.mul_ld, p.res1_ld, p.res2_ld, post, act, p.shuffle,
                              cout, Ho, Wo, p.slope};
          if (nxv > 0) epilogue_ragged(ra, it.n, it.y0, it.x0, sb, txb, rd_off, co, ty0, 0);
          if (nxv > 4) epilogue_ragged(ra, it.n, it.y0, it.x0, sb, txb + 4, rd_off, co, ty0, 0);
        }
      }
    }
    } else {
    const int quad = warp & 3, half = warp >> 2;
    const bool is_lo = (quad & 1) != 0;
    float* slab = stage_buf + half * (2 * 16 * 128);
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
    const int t = quad * 32 + lane;             // thread index within the half (= its TMEM lane = slab row)
    const int c4 = t & 15, tx = t >> 4;         // read phase: channels 4*c4..+3 of the item's 64, tile column tx
    const int rd_off = (c4 >> 3) * 64 + (c4 & 7) * 4;   // hi float4 of those channels inside a slab pixel; lo at +32
    const int Ho = p.Ho, Wo = p.Wo, cout = p.cout, act = p.act, post = p.post;
    const int sh = p.shuffle == 2 ? 2 : 1;
    const int cr = cout >> 2;
    const int oW = Wo * sh, oH = Ho * sh;
    const bool planar = p.out_planar != 0;
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    const bool vec_ok = !planar && (p.out_ld & 3) == 0 && al16(p.out) && (sh == 1 || (cr & 3) == 0) &&
                        (post == TDVC_POST_NONE || ((p.mul_ld & 3) == 0 && al16(p.mul))) &&
                        (!p.res1 || ((p.res1_ld & 3) == 0 && al16(p.res1))) && (!p.res2 || ((p.res2_ld & 3) == 0 && al16(p.res2)));
    const bool planar_vec = planar && (Wo & 7) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 31) == 0;
    // a plain store may also fill the pad lanes of a channel-padded view (weight rows >= cout are zero, no bias there)
    const int cout_st = (sh == 1 && post == TDVC_POST_NONE && !p.res1 && !p.res2 && p.out_ld >= ((cout + 3) & ~3)) ? ((cout + 3) & ~3) : cout;
    // branch-free activation: a(v) = min(max(v, a_neg * v), a_hi)   (0 <= a_neg <= 1)
    const float a_neg = act == TDVC_ACT_NONE ? 1.f : (act == TDVC_ACT_LRELU ? p.slope : 0.f);
    const float a_hi = act == TDVC_ACT_CLAMP01 ? 1.f : __int_as_float(0x7f800000);
    // element strides of one tile row in out / mul / res1 / res2 (all address the same logical pixel)
    const int o_rs = sh * oW * p.out_ld, m_rs = sh * oW * p.mul_ld, r1_rs = sh * oW * p.res1_ld, r2_rs = sh * oW * p.res2_ld;
    int acc_it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++acc_it) {
      const Item it = decode_item(item, n_jt, tiles_x, tiles_y, flip);
      const int sa = acc_it & 1;
      mbar_wait(bar(ACC_FULL + sa), (acc_it >> 1) & 1);
      tc_fence_after();
      // write phase: bias of this lane's weight row (lo rows only)
      float wbias = 0.f;
      if (is_lo && p.bias) {
        const int cw = it.jt * NT + (quad >> 1) * 32 + lane;
        if (cw < cout) wbias = __ldg(p.bias + cw);
      }
      // read phase: where this thread's 4 channels land (PixelShuffle(2) folded into the store: co' = q*(cout/4) + c)
      const int co = it.jt * NT + 4 * c4;
      int oc = co, qy = 0, qx = 0;
      if (sh == 2) {
        const int q = co < cout ? co / cr : 0;
        oc = co - q * cr;
        qy = q >> 1;
        qx = q & 1;
      }
      const bool th_ok = it.x0 + tx < Wo && co < cout;
      const bool th_vec = vec_ok && co + 4 <= cout_st;
      const int ny = Ho - it.y0;  // valid tile rows
      const int64_t pix0 = ((int64_t)it.n * oH + (it.y0 * sh + qy)) * oW + ((it.x0 + tx) * sh + qx);
      float* const o0 = p.out + pix0 * p.out_ld + oc;
      const float* const m0 = post != TDVC_POST_NONE ? p.mul + pix0 * p.mul_ld + oc : nullptr;
      const float* const r10 = p.res1 ? p.res1 + pix0 * p.res1_ld + oc : nullptr;
      const float* const r20 = p.res2 ? p.res2 + pix0 * p.res2_ld + oc : nullptr;
      uint32_t r[16];
      tmem_ld16(lane_addr + (uint32_t)(sa * NPX + half * 128), r);
      // residual (res1) and GDN-multiplier rows: their DRAM latency (~1.5 us under load) is longer than the slab exchange of
      // a chunk, so they are pulled into L2 two chunks (4 tile rows) ahead with prefetch.global.L2 - no registers held
      // (holding them in registers one chunk ahead spilled at the 96-register cap and slowed every variant)
      auto l2_prefetch_rows = [&](int tyb) {
        if (!(th_vec && th_ok)) return;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          if (tyb + q < ny) {
            if (m0) prefetch_l2(m0 + (tyb + q) * m_rs);
            if (r10) prefetch_l2(r10 + (tyb + q) * r1_rs);
            if (r20) prefetch_l2(r20 + (tyb + q) * r2_rs);
          }
        }
      };
      const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m0 || r10 || r20) {
        l2_prefetch_rows(half * 16);
        l2_prefetch_rows(half * 16 + 2);
      }
#pragma unroll 1
      for (int c = 0; c < 8; ++c) {
        const int ty0 = half * 16 + c * 2;
        float4 ma = z4, mb = z4, ra = z4, rbv = z4, rc = z4, rd = z4;
        const bool row0 = th_vec && th_ok && ty0 < ny, row1 = row0 && ty0 + 1 < ny;
        tmem_ld_wait();
        if (c == 7) {  // all TMEM reads of this warp are done: the accumulator stage may be overwritten
          tc_fence_before();
          mbar_arrive(bar(ACC_EMPTY + sa));
        }
        float* sb = slab + (c & 1) * (16 * 128);
        if (is_lo) {
#pragma unroll
          for (int j = 0; j < 16; ++j) sb[j * 128 + t] = fmaf(__uint_as_float(r[j]), kLoUnscale, wbias);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) sb[j * 128 + t] = __uint_as_float(r[j]);
        }
        // next chunk's accumulator columns: the TMEM read overlaps the barrier and the read phase below
        if (c < 7) tmem_ld16(lane_addr + (uint32_t)(sa * NPX + half * 128 + (c + 1) * 16), r);
        asm volatile("bar.sync %0, 128;" ::"r"(1 + half) : "memory");
        if (th_vec) {
          if (!row0) continue;
          const bool two = row1;
          if (c < 6 && (m0 || r10 || r20)) l2_prefetch_rows(ty0 + 4);
          if (m0) { ma = __ldg(reinterpret_cast<const float4*>(m0 + ty0 * m_rs)); if (two) mb = __ldg(reinterpret_cast<const float4*>(m0 + (ty0 + 1) * m_rs)); }
          if (r10) { ra = __ldg(reinterpret_cast<const float4*>(r10 + ty0 * r1_rs)); if (two) rbv = __ldg(reinterpret_cast<const float4*>(r10 + (ty0 + 1) * r1_rs)); }
          if (r20) { rc = __ldg(reinterpret_cast<const float4*>(r20 + ty0 * r2_rs)); if (two) rd = __ldg(reinterpret_cast<const float4*>(r20 + (ty0 + 1) * r2_rs)); }
          const float* rb = sb + tx * 128 + rd_off;
          const float4 h0 = *reinterpret_cast<const float4*>(rb), l0 = *reinterpret_cast<const float4*>(rb + 32);
          const float4 h1 = *reinterpret_cast<const float4*>(rb + 8 * 128), l1 = *reinterpret_cast<const float4*>(rb + 8 * 128 + 32);
          float4 v0 = make_float4(h0.x + l0.x, h0.y + l0.y, h0.z + l0.z, h0.w + l0.w);
          float4 v1 = make_float4(h1.x + l1.x, h1.y + l1.y, h1.z + l1.z, h1.w + l1.w);
          if (m0) {  // GDN / IGDN: v = mul * rsqrt(v) | mul * sqrt(v)
            if (post == TDVC_POST_IGDN) {
              v0 = make_float4(ma.x * sqrtf(v0.x), ma.y * sqrtf(v0.y), ma.z * sqrtf(v0.z), ma.w * sqrtf(v0.w));
              v1 = make_float4(mb.x * sqrtf(v1.x), mb.y * sqrtf(v1.y), mb.z * sqrtf(v1.z), mb.w * sqrtf(v1.w));
            } else {
              v0 = make_float4(ma.x * rsqrtf(v0.x), ma.y * rsqrtf(v0.y), ma.z * rsqrtf(v0.z), ma.w * rsqrtf(v0.w));
              v1 = make_float4(mb.x * rsqrtf(v1.x), mb.y * rsqrtf(v1.y), mb.z * rsqrtf(v1.z), mb.w * rsqrtf(v1.w));
            }
          }
          // activation a(v) = min(max(v, a_neg * v), a_hi): identity (a_neg 1), ReLU (0), LeakyReLU (slope), clamp01
          v0.x = fminf(fmaxf(v0.x, a_neg * v0.x), a_hi) + (ra.x + rc.x);
          v0.y = fminf(fmaxf(v0.y, a_neg * v0.y), a_hi) + (ra.y + rc.y);
          v0.z = fminf(fmaxf(v0.z, a_neg * v0.z), a_hi) + (ra.z + rc.z);
          v0.w = fminf(fmaxf(v0.w, a_neg * v0.w), a_hi) + (ra.w + rc.w);
          *reinterpret_cast<float4*>(o0 + ty0 * o_rs) = v0;
          if (two) {
            v1.x = fminf(fmaxf(v1.x, a_neg * v1.x), a_hi) + (rbv.x + rd.x);
            v1.y = fminf(fmaxf(v1.y, a_neg * v1.y), a_hi) + (rbv.y + rd.y);
            v1.z = fminf(fmaxf(v1.z, a_neg * v1.z), a_hi) + (rbv.z + rd.z);
            v1.w = fminf(fmaxf(v1.w, a_neg * v1.w), a_hi) + (rbv.w + rd.w);
            *reinterpret_cast<float4*>(o0 + (ty0 + 1) * o_rs) = v1;
          }
        } else if (planar) {  // NCHW planes: thread = (channel, tile row); 8 x-adjacent pixels are contiguous in the plane
          const int ch = t & 63, h2 = t >> 6;
          const int ty = ty0 + h2;
          const int cp = it.jt * NT + ch;
          if (cp < cout && ty < ny) {
            const float* rb = sb + h2 * (8 * 128) + (ch >> 5) * 64 + (ch & 31);
            float v[8];
#pragma unroll
            for (int x = 0; x < 8; ++x) {
              const float o = rb[x * 128] + rb[x * 128 + 32];
              v[x] = fminf(fmaxf(o, a_neg * o), a_hi);
            }
            float* op = p.out + (((int64_t)it.n * cout + cp) * Ho + (it.y0 + ty)) * Wo + it.x0;
            if (planar_vec && it.x0 + 8 <= Wo) {
              stg256(op, v);   // one full 32-byte sector: two 16-byte stores made L2 read the sector back (partial writes)
            } else {
#pragma unroll
              for (int x = 0; x < 8; ++x)
                if (it.x0 + x < Wo) op[x] = v[x];
            }
          }
        } else if (th_ok) {
          epilogue_ragged(RaggedArgs{p.out, p.mul, p.res1, p.res2, p.out_ld, p.mul_ld, p.res1_ld, p.res2_ld, post, act, p.shuffle,
                                     cout, Ho, Wo, p.slope},
                          it.n, it.y0, it.x0, sb, tx, rd_off, co, ty0, 32);
        }
      }
    }
    }
  } else if (warp < kEpiWarps + kProdWarps) {
    // ===================================================================== producers: fp32 halo -> fp16 hi/lo planes
    // LPP lanes cover the CK channels of a pixel (one float4 each), 32/LPP pixels per warp-wide load; the halo is
    // walked as a flat pixel list s = (k*8 + warp)*PPI + lane/LPP.  Each thread keeps, for its PER_WARP loads, the
    // source-pixel offset in a register table (the same for every item and unit), so an interior tile costs one
    // address instruction per load.  The loads of a unit are split into NBATCH batches and software-pipelined across
    // batches AND units: the loads of batch b+1 (or of the next unit's batch 0) are issued before batch b is converted,
    // so the memory latency is covered by the fp32 -> fp16 hi/lo conversion of the previous batch.
    using P = ProdCfg<C, CK>;
    const int pw = warp - kEpiWarps;
    const int fi = lane % P::LPP, psub = lane / P::LPP;
    ProdThread th;
    th.pw = pw;
    th.psub = psub;
    th.vmask = 0;
    int tab[P::PER_WARP];   // non-PLANES: source pixel offset hy*STEP*W + hx*STEP;  PLANES: hy << 8 | hx
#pragma unroll
    for (int k = 0; k < P::PER_WARP; ++k) {
      const int sidx = (k * kProdWarps + pw) * P::PPI + psub;
      const int hy = sidx / C::IW, hx = sidx - hy * C::IW;
      tab[k] = C::PLANES ? ((hy << 8) | hx) : (hy * C::STEP * p.W + hx * C::STEP);
      if (sidx < C::NHALO) th.vmask |= 1u << k;
    }
    th.lane_smem = (uint32_t)(((fi >> 1) * C::NPIXP) * 16 + (fi & 1) * 8 + (pw * P::PPI + psub) * 16);

    // unit context: where the lane's 4 channels of unit (item, u) come from and which stage they go to
    int item = blockIdx.x, u = 0, sX = 0, phX = 1;
    auto setup = [&](ProdUnit& c) {
      const Item it = decode_item(item, n_jt, tiles_x, tiles_y, flip);
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
  } else if (warp == kMmaWarp) {
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
          for (int ky = 0; ky < KS; ++ky) {
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
                tc_mma(d, wk, xl, IDESC, 1u);
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
  if (warp == kMmaWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS)
                 : "memory");
  }
}

// fp32 packed weights [T][cin_pad][cout_pad] -> per (cout tile of 64, unit, tap) fp16 block [128 rows][CK] in the
// canonical K-major no-swizzle layout [(row/8)][(k/8)][row%8][k%8].  Row order = TMEM lane order of the accumulator:
// rows 0-31 hi of channels 0-31, rows 32-63 lo * 2^12 of channels 0-31, rows 64-95 hi of 32-63, rows 96-127 lo of 32-63.
// split = 1: per (cout tile of 128, unit, tap) two such blocks, W_hi then W_lo, row = channel, both scaled by 2^w_shift.
__global__ void pack_f16_kernel(const float* __restrict__ w, __half* __restrict__ out, int T, int cin, int cin_pad,
                                 int cout, int cout_pad, int CK, int n_units, int n_jt, int split, float scale) {
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
    const int co = split ? jt * 128 + row : jt * NT + (quad >> 1) * 32 + (row & 31);
    const bool lo = split ? which == 1 : (quad & 1) != 0;
    float v = 0.f;
    if (ci < cin_pad && co < cout_pad && ci < cin && co < cout) v = w[((int64_t)tap * cin_pad + ci) * cout_pad + co] * scale;
    v = fminf(fmaxf(v, -65504.f), 65504.f);
    const __half hi = __float2half_rn(v);
    out[i] = lo ? __float2half_rn((v - __half2float(hi)) * (split ? 1.f : kLoScale)) : hi;
  }
}

struct Choice {
  int ks, ck, s, split;
};

// Every KxK (K in 1,3,5,7; pad K/2) stride-1 convolution and the stride-2 3x3 / 1x1 ones have a tensor-core path;
// input channels are processed in chunks of ck; output channels in tiles of 64 (4-product scheme) or, for layers
// with >= 96 output channels (a multiple of 4), in tiles of 128 with the 3-product split scheme.
static bool choose(const TdvcConvParams& p, Choice* c) {
  if (p.kh != p.kw || p.pad != p.kh / 2 || p.cin < 4 || p.cout < 1) return false;
  // 1x1 layers stay on 64-channel tiles: 128-channel split tiles were measured slower for them (GDN 128->128 @512x960:
  // 0.35 -> 0.48 ms) - they are bound by the epilogue's latency chain, and the smaller tiles give twice as many items
  const int split = (p.cout >= 96 && (p.cout & 3) == 0 && p.kh != 7 && p.kh != 1) ? 1 : 0;
  if (p.stride == 2) {
    if (p.kh == 3) { *c = {3, 16, 2, split}; return true; }
    if (p.kh == 1) { *c = {1, 32, 2, split}; return p.cin >= 32; }
    return false;
  }
  if (p.stride != 1) return false;
  const int ck = p.cin > 16 ? 32 : 16;
  if (p.kh == 3) { *c = {3, ck, 1, ck == 32 ? split : 0}; return true; }
  if (p.kh == 7) { *c = {7, ck, 1, 0}; return true; }
  if (p.kh == 1 || p.kh == 5) { *c = {p.kh, 32, 1, split}; return p.cin >= 32; }
  return false;
}

template <int KS, int CK, int S, int SPLIT>
static int launch(const TdvcConvParams& p, cudaStream_t st) {
  using C = Cfg<KS, CK, S, SPLIT>;
  static bool attr_set = false;  // idempotent; a benign race sets it twice
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<KS, CK, S, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
    if (e != cudaSuccess) {
      set_error("conv_tc: cudaFuncSetAttribute(%d bytes) failed: %s", C::SMEM, cudaGetErrorString(e));
      return TDVC_ECUDA;
    }
    attr_set = true;
  }
  if (SPLIT) TDVC_REQUIRE(p.w_shift >= -100 && p.w_shift <= 100, "conv_tc: w_shift %d out of range", p.w_shift);
  const int tiles_x = cdiv(p.Wo, TW), tiles_y = cdiv(p.Ho, TH);
  const int n_jt = cdiv(p.cout, C::NTT), n_units = cdiv(p.cin, CK);
  const int64_t items = (int64_t)p.N * tiles_x * tiles_y * n_jt;
  TDVC_REQUIRE(items < (1ll << 31), "conv_tc: too many work items");
  const int grid = (int)(items < kNumSMs ? items : kNumSMs);
  conv_tc_kernel<KS, CK, S, SPLIT><<<grid, kThreads, C::SMEM, st>>>(p, tiles_x, tiles_y, n_jt, n_units, (int)items);
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

int conv2d_tc(const TdvcConvParams& p, cudaStream_t st) {
  tc::Choice c;
  if (!tc::choose(p, &c)) {
    set_error("conv_tc: unsupported shape");
    return TDVC_EINVAL;
  }
  if (c.split) {
    if (c.s == 2 && c.ks == 3) return tc::launch<3, 16, 2, 1>(p, st);
    if (c.s == 2 && c.ks == 1) return tc::launch<1, 32, 2, 1>(p, st);
    if (c.ks == 3) return tc::launch<3, 32, 1, 1>(p, st);
    if (c.ks == 1) return tc::launch<1, 32, 1, 1>(p, st);
    if (c.ks == 5) return tc::launch<5, 32, 1, 1>(p, st);
  } else {
    if (c.s == 2 && c.ks == 3) return tc::launch<3, 16, 2, 0>(p, st);
    if (c.s == 2 && c.ks == 1) return tc::launch<1, 32, 2, 0>(p, st);
    if (c.ks == 3 && c.ck == 32) return tc::launch<3, 32, 1, 0>(p, st);
    if (c.ks == 3 && c.ck == 16) return tc::launch<3, 16, 1, 0>(p, st);
    if (c.ks == 1) return tc::launch<1, 32, 1, 0>(p, st);
    if (c.ks == 5) return tc::launch<5, 32, 1, 0>(p, st);
    if (c.ks == 7 && c.ck == 32) return tc::launch<7, 32, 1, 0>(p, st);
    if (c.ks == 7 && c.ck == 16) return tc::launch<7, 16, 1, 0>(p, st);
  }
  set_error("conv_tc: no instantiation for ks=%d ck=%d stride=%d split=%d", c.ks, c.ck, c.s, c.split);
  return TDVC_EINVAL;
}

}  // namespace tdvc

using namespace tdvc;

static size_t f16_elems(const TdvcConvParams* p, const tc::Choice& c) {
  const int ntt = c.split ? 128 : tc::NT;
  const int n_jt = cdiv(p->cout, ntt), n_units = cdiv(p->cin, c.ck);
  return (size_t)n_jt * n_units * c.ks * c.ks * (c.split ? 2 : 1) * 128 * c.ck;
}

extern "C" size_t tdvc_conv2d_f16_bytes(const TdvcConvParams* p) {
  tc::Choice c;
  if (p == nullptr || !tc::choose(*p, &c)) return 0;
  return f16_elems(p, c) * sizeof(__half);
}

extern "C" int tdvc_conv2d_f16_is_split(const TdvcConvParams* p) {
  tc::Choice c;
  if (p == nullptr || !tc::choose(*p, &c)) return 0;
  return c.split;
}

extern "C" int tdvc_conv2d_pack_f16(const TdvcConvParams* p, void* out, void* stream) {
  tc::Choice c;
  TDVC_REQUIRE(p && out && p->weight, "conv2d_pack_f16: null pointer");
  TDVC_REQUIRE(tc::choose(*p, &c), "conv2d_pack_f16: shape has no tcgen05 path");
  TDVC_REQUIRE(!c.split || (p->w_shift >= -100 && p->w_shift <= 100), "conv2d_pack_f16: w_shift %d out of range", p->w_shift);
  const int ntt = c.split ? 128 : tc::NT;
  const int n_jt = cdiv(p->cout, ntt), n_units = cdiv(p->cin, c.ck);
  const int64_t total = (int64_t)f16_elems(p, c);
  int grid = cdiv(total, 256);
  if (grid > kNumSMs * 8) grid = kNumSMs * 8;
  const float scale = c.split ? ldexpf(1.f, p->w_shift) : 1.f;
  tc::pack_f16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p->weight, static_cast<__half*>(out), c.ks * c.ks, p->cin,
                                                             p->cin_pad, p->cout, p->cout_pad, c.ck, n_units, n_jt, c.split, scale);
  TDVC_CHECK_LAUNCH("conv2d_pack_f16");
  return TDVC_OK;
}
