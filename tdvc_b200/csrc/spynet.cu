// SPyNet / OffsetGen memory-bound pieces (reference main/model/flownet.py:8-48,82-140; pnet.py:117,159,163).
// Images are NHWC with ld 4, so a pixel is one float4 access.
#include "common.cuh"

namespace tdvc {

// F.avg_pool2d(kernel 2, stride 2): raster-order sum then divide (flownet.py:102-114)
__global__ void avgpool2x2_kernel(const float* __restrict__ src, float* __restrict__ dst, int N, int H, int W, int C4) {
  const int Ho = H >> 1, Wo = W >> 1;
  const int64_t total = (int64_t)N * Ho * Wo * C4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int c = (int)(i % C4);
    int64_t r = i / C4;
    const int x = (int)(r % Wo); r /= Wo;
    const int y = (int)(r % Ho);
    const int n = (int)(r / Ho);
    const float4* s = reinterpret_cast<const float4*>(src) + (((int64_t)n * H + 2 * y) * W + 2 * x) * C4 + c;
    const float4 a = __ldg(s), b = __ldg(s + C4), d = __ldg(s + (int64_t)W * C4), e = __ldg(s + (int64_t)W * C4 + C4);
    float4 o;
    o.x = (((a.x + b.x) + d.x) + e.x) / 4.f; o.y = (((a.y + b.y) + d.y) + e.y) / 4.f;
    o.z = (((a.z + b.z) + d.z) + e.z) / 4.f; o.w = (((a.w + b.w) + d.w) + e.w) / 4.f;
    reinterpret_cast<float4*>(dst)[i] = o;
  }
}

// One SPyNet level's input assembly: x2 flow upsample (* 2), border-clamped bilinear backward warp of the support image along
// that flow, concat [ref(3), warped(3), flow(2)] (reference flownet.py:8-48, 116-138).  flow_prev: (N, h/2, w/2, 2) or NULL
// (level 0 => zero flow).
// A CTA produces a TW x TH tile of the level.  The support image is a float4 (ld 4) per pixel; its (TW + 2M) x (TH + 2M)
// neighbourhood of the tile is staged in shared memory with coalesced float4 row loads, and the four corners of a pixel's
// sample come from that staged tile (the flow of one level is a few pixels; a sample that leaves the margin M reads global
// memory instead - same values, same arithmetic).  Image reads thus are whole rows instead of per-lane scattered sectors.
// Measured at 1024x1920 (CUDA events inside the frame): 64x8 tiles 49 us, 128x16 tiles (72 KB, fewer resident CTAs) 69 us; the
// same kernel without staging, every corner an __ldg, 42-44 us - the 16-byte pixels of a 31 MB image that the previous kernel
// just wrote sit in L2 / L1 either way, so the staging buys coalescing the caches already provided.
constexpr int SP_TW = 64, SP_TH = 8, SP_M = 8;
constexpr int SP_SW = SP_TW + 2 * SP_M, SP_SH = SP_TH + 2 * SP_M;
__global__ void __launch_bounds__(256) spynet_prep_kernel(const float* __restrict__ ref4, const float* __restrict__ supp4,
                                                          const float* __restrict__ flow_prev, float* __restrict__ out8, int N,
                                                          int h, int w, int tiles_x, int tiles_y) {
  __shared__ float4 tile[SP_SH][SP_SW];
  const int hp = h >> 1, wp = w >> 1;
  // align_corners=True source scale, as ATen computes it: (in-1)/(out-1) in fp32
  const float rh = h > 1 ? (float)(hp - 1) / (float)(h - 1) : 0.f;
  const float rw = w > 1 ? (float)(wp - 1) / (float)(w - 1) : 0.f;
  const float wm1 = (float)(w - 1 > 1 ? w - 1 : 1), hm1 = (float)(h - 1 > 1 ? h - 1 : 1);
  int b = blockIdx.x;
  const int tx = b % tiles_x; b /= tiles_x;
  const int ty = b % tiles_y;
  const int n = b / tiles_y;
  const int bx0 = tx * SP_TW - SP_M, by0 = ty * SP_TH - SP_M;   // image coordinates of tile[0][0]
  const float4* sp = reinterpret_cast<const float4*>(supp4) + (int64_t)n * h * w;
  for (int i = threadIdx.x; i < SP_SH * SP_SW; i += blockDim.x) {
    const int sy = i / SP_SW, sx = i - sy * SP_SW;
    const int gy = by0 + sy, gx = bx0 + sx;
    tile[sy][sx] = (gy >= 0 && gy < h && gx >= 0 && gx < w) ? __ldg(sp + (int64_t)gy * w + gx) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();
  for (int t = threadIdx.x; t < SP_TW * SP_TH; t += blockDim.x) {
    const int y = ty * SP_TH + t / SP_TW, x = tx * SP_TW + t % SP_TW;
    if (y >= h || x >= w) continue;
    const int64_t i = ((int64_t)n * h + y) * w + x;
    float fx = 0.f, fy = 0.f;
    if (flow_prev != nullptr) {
      const float h1r = rh * (float)y, w1r = rw * (float)x;
      const int h1 = (int)h1r, w1 = (int)w1r;
      const int h1p = h1 < hp - 1 ? 1 : 0, w1p = w1 < wp - 1 ? 1 : 0;
      const float h1l = h1r - (float)h1, w1l = w1r - (float)w1;
      const float h0l = 1.f - h1l, w0l = 1.f - w1l;
      const float2* fp = reinterpret_cast<const float2*>(flow_prev) + ((int64_t)n * hp + h1) * wp + w1;
      const float2 v00 = __ldg(fp), v01 = __ldg(fp + w1p), v10 = __ldg(fp + (int64_t)h1p * wp), v11 = __ldg(fp + (int64_t)h1p * wp + w1p);
      fx = (h0l * (w0l * v00.x + w1l * v01.x) + h1l * (w0l * v10.x + w1l * v11.x)) * 2.0f;
      fy = (h0l * (w0l * v00.y + w1l * v01.y) + h1l * (w0l * v10.y + w1l * v11.y)) * 2.0f;
    }
    // flow_warp: normalise to [-1,1] then grid_sample's un-normalise (align_corners=True), border clamp
    const float gx = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, __fadd_rn((float)x, fx)), wm1), 1.0f);
    const float gy = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, __fadd_rn((float)y, fy)), hm1), 1.0f);
    float ix = __fmul_rn(__fdiv_rn(__fadd_rn(gx, 1.f), 2.f), (float)(w - 1));
    float iy = __fmul_rn(__fdiv_rn(__fadd_rn(gy, 1.f), 2.f), (float)(h - 1));
    ix = fminf((float)(w - 1), fmaxf(ix, 0.f));
    iy = fminf((float)(h - 1), fmaxf(iy, 0.f));
    const float ixf = floorf(ix), iyf = floorf(iy);
    const int x0 = (int)ixf, y0 = (int)iyf;
    const float txw = ix - ixf, tyw = iy - iyf;  // (ix - ix_nw), (iy - iy_nw)
    const float ux = (ixf + 1.f) - ix, uy = (iyf + 1.f) - iy;  // (ix_se - ix), (iy_se - iy)
    const float wnw = ux * uy, wne = txw * uy, wsw = ux * tyw, wse = txw * tyw;
    // corner fetch: staged tile when the 2x2 footprint lies inside it, global memory otherwise
    const int lx = x0 - bx0, ly = y0 - by0;
    const bool staged = lx >= 0 && ly >= 0 && lx + 1 < SP_SW && ly + 1 < SP_SH;
    auto corner = [&](int dy, int dx) {
      return staged ? tile[ly + dy][lx + dx] : __ldg(sp + (int64_t)(y0 + dy) * w + x0 + dx);
    };
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    {  // x0,y0 are always in range after the clamp
      const float4 v = corner(0, 0);
      acc.x += v.x * wnw; acc.y += v.y * wnw; acc.z += v.z * wnw;
    }
    if (x0 + 1 < w) {
      const float4 v = corner(0, 1);
      acc.x += v.x * wne; acc.y += v.y * wne; acc.z += v.z * wne;
    }
    if (y0 + 1 < h) {
      const float4 v = corner(1, 0);
      acc.x += v.x * wsw; acc.y += v.y * wsw; acc.z += v.z * wsw;
    }
    if (x0 + 1 < w && y0 + 1 < h) {
      const float4 v = corner(1, 1);
      acc.x += v.x * wse; acc.y += v.y * wse; acc.z += v.z * wse;
    }
    const float4 rf = __ldg(reinterpret_cast<const float4*>(ref4) + i);
    float4* o = reinterpret_cast<float4*>(out8) + i * 2;
    o[0] = make_float4(rf.x, rf.y, rf.z, acc.x);
    o[1] = make_float4(acc.y, acc.z, fx, fy);
  }
}

// ---- backward of spynet_prep with respect to the coarse flow (the images are inputs of the network: no gradient).
// out8 = [ref(3), warp(supp, f)(3), f(2)] with f = 2 * upsample_x2(flow_prev) (align_corners=True):
// d loss / d f = grad8[6..7] + sum_c grad8[3 + c] * d warp_c / d f   (grid_sample backward: the bilinear weights' derivative,
// corners outside the image contribute zero, a clamped coordinate has zero derivative - ATen's border-padding rule),
// then the adjoint of the x2 upsample, as a GATHER over the fine pixels that read a coarse one (deterministic).
__global__ void spynet_prep_bwd_fine_kernel(const float* __restrict__ supp4, const float* __restrict__ flow_prev,
                                            const float* __restrict__ grad8, float* __restrict__ gf, int N, int h, int w) {
  const int hp = h >> 1, wp = w >> 1;
  const float rh = h > 1 ? (float)(hp - 1) / (float)(h - 1) : 0.f;
  const float rw = w > 1 ? (float)(wp - 1) / (float)(w - 1) : 0.f;
  const float wm1 = (float)(w - 1 > 1 ? w - 1 : 1), hm1 = (float)(h - 1 > 1 ? h - 1 : 1);
  const int64_t total = (int64_t)N * h * w;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % w);
    const int64_t q = i / w;
    const int y = (int)(q % h), n = (int)(q / h);
    float fx = 0.f, fy = 0.f;
    {
      const float h1r = rh * (float)y, w1r = rw * (float)x;
      const int h1 = (int)h1r, w1 = (int)w1r;
      const int h1p = h1 < hp - 1 ? 1 : 0, w1p = w1 < wp - 1 ? 1 : 0;
      const float h1l = h1r - (float)h1, w1l = w1r - (float)w1;
      const float h0l = 1.f - h1l, w0l = 1.f - w1l;
      const float2* fp = reinterpret_cast<const float2*>(flow_prev) + ((int64_t)n * hp + h1) * wp + w1;
      const float2 v00 = __ldg(fp), v01 = __ldg(fp + w1p), v10 = __ldg(fp + (int64_t)h1p * wp), v11 = __ldg(fp + (int64_t)h1p * wp + w1p);
      fx = (h0l * (w0l * v00.x + w1l * v01.x) + h1l * (w0l * v10.x + w1l * v11.x)) * 2.0f;
      fy = (h0l * (w0l * v00.y + w1l * v01.y) + h1l * (w0l * v10.y + w1l * v11.y)) * 2.0f;
    }
    const float gxn = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, __fadd_rn((float)x, fx)), wm1), 1.0f);
    const float gyn = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, __fadd_rn((float)y, fy)), hm1), 1.0f);
    float ix = __fmul_rn(__fdiv_rn(__fadd_rn(gxn, 1.f), 2.f), (float)(w - 1));
    float iy = __fmul_rn(__fdiv_rn(__fadd_rn(gyn, 1.f), 2.f), (float)(h - 1));
    // d ix / d fx = (w - 1) / wm1 inside the image, 0 where the coordinate is clamped (ATen clip_coordinates_set_grad)
    const float mx = (ix <= 0.f || ix >= (float)(w - 1)) ? 0.f : (float)(w - 1) / wm1;
    const float my = (iy <= 0.f || iy >= (float)(h - 1)) ? 0.f : (float)(h - 1) / hm1;
    ix = fminf((float)(w - 1), fmaxf(ix, 0.f));
    iy = fminf((float)(h - 1), fmaxf(iy, 0.f));
    const float ixf = floorf(ix), iyf = floorf(iy);
    const int x0 = (int)ixf, y0 = (int)iyf;
    const float txw = ix - ixf, tyw = iy - iyf, ux = (ixf + 1.f) - ix, uy = (iyf + 1.f) - iy;
    const float4* sp = reinterpret_cast<const float4*>(supp4) + (int64_t)n * h * w;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 nw = __ldg(sp + (int64_t)y0 * w + x0);
    const float4 ne = x0 + 1 < w ? __ldg(sp + (int64_t)y0 * w + x0 + 1) : z;
    const float4 sw = y0 + 1 < h ? __ldg(sp + (int64_t)(y0 + 1) * w + x0) : z;
    const float4 se = (x0 + 1 < w && y0 + 1 < h) ? __ldg(sp + (int64_t)(y0 + 1) * w + x0 + 1) : z;
    const float* g = grad8 + i * 8;
    const float g0 = g[3], g1 = g[4], g2 = g[5];
    // d warp_c / d ix = (ne - nw) * uy + (se - sw) * tyw ; d warp_c / d iy = (sw - nw) * ux + (se - ne) * txw
    const float dix = g0 * ((ne.x - nw.x) * uy + (se.x - sw.x) * tyw) + g1 * ((ne.y - nw.y) * uy + (se.y - sw.y) * tyw) +
                      g2 * ((ne.z - nw.z) * uy + (se.z - sw.z) * tyw);
    const float diy = g0 * ((sw.x - nw.x) * ux + (se.x - ne.x) * txw) + g1 * ((sw.y - nw.y) * ux + (se.y - ne.y) * txw) +
                      g2 * ((sw.z - nw.z) * ux + (se.z - ne.z) * txw);
    gf[i * 2] = g[6] + dix * mx;
    gf[i * 2 + 1] = g[7] + diy * my;
  }
}

// adjoint of f = 2 * upsample_x2(flow_prev), align_corners=True: every coarse pixel gathers from the fine pixels whose
// interpolation footprint contains it; the source index / weights are recomputed exactly as the forward computes them.
__global__ void spynet_prep_bwd_coarse_kernel(const float* __restrict__ gf, float* __restrict__ gflow, int N, int h, int w) {
  const int hp = h >> 1, wp = w >> 1;
  const float rh = h > 1 ? (float)(hp - 1) / (float)(h - 1) : 0.f;
  const float rw = w > 1 ? (float)(wp - 1) / (float)(w - 1) : 0.f;
  const int64_t total = (int64_t)N * hp * wp;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int X = (int)(i % wp);
    const int64_t q = i / wp;
    const int Y = (int)(q % hp), n = (int)(q / hp);
    // fine rows whose h1 is Y - 1 or Y lie in [(Y - 1) / rh, (Y + 1) / rh); widened by 2 against rounding
    int ylo = 0, yhi = h - 1, xlo = 0, xhi = w - 1;
    if (rh > 0.f) { ylo = (int)floorf((float)(Y - 1) / rh) - 2; yhi = (int)ceilf((float)(Y + 1) / rh) + 2; }
    if (rw > 0.f) { xlo = (int)floorf((float)(X - 1) / rw) - 2; xhi = (int)ceilf((float)(X + 1) / rw) + 2; }
    ylo = ylo < 0 ? 0 : ylo; yhi = yhi > h - 1 ? h - 1 : yhi;
    xlo = xlo < 0 ? 0 : xlo; xhi = xhi > w - 1 ? w - 1 : xhi;
    float ax = 0.f, ay = 0.f;
    for (int y = ylo; y <= yhi; ++y) {
      const float h1r = rh * (float)y;
      const int h1 = (int)h1r;
      const int h1p = h1 < hp - 1 ? 1 : 0;
      const float h1l = h1r - (float)h1;
      float wy = 0.f;
      if (h1 == Y) wy += 1.f - h1l;
      if (h1 + h1p == Y) wy += h1l;
      if (wy == 0.f) continue;
      for (int x = xlo; x <= xhi; ++x) {
        const float w1r = rw * (float)x;
        const int w1 = (int)w1r;
        const int w1p = w1 < wp - 1 ? 1 : 0;
        const float w1l = w1r - (float)w1;
        float wx = 0.f;
        if (w1 == X) wx += 1.f - w1l;
        if (w1 + w1p == X) wx += w1l;
        if (wx == 0.f) continue;
        const float2 gv = *reinterpret_cast<const float2*>(gf + (((int64_t)n * h + y) * w + x) * 2);
        ax += wy * wx * gv.x;
        ay += wy * wx * gv.y;
      }
    }
    gflow[i * 2] = 2.f * ax;
    gflow[i * 2 + 1] = 2.f * ay;
  }
}

// nn.Upsample(scale_factor=2, bilinear, align_corners=False): src = 0.5*(dst+0.5)-0.5 clamped at 0.
// One block row per output row (blockIdx.y = n*Ho + y): the row's source rows and vertical weights are block constants and
// the index arithmetic is 32-bit; consecutive threads write consecutive float4 of the output row.
__global__ void upsample2x_kernel(const float* __restrict__ src, float* __restrict__ dst, int N, int H, int W, int C4) {
  const int Ho = 2 * H, Wo = 2 * W;
  const int y = blockIdx.y % Ho, n = blockIdx.y / Ho;
  float sy = 0.5f * ((float)y + 0.5f) - 0.5f;
  sy = sy < 0.f ? 0.f : sy;
  const int y1 = (int)sy;
  const int yp = y1 < H - 1 ? 1 : 0;
  const float ly = sy - (float)y1, hy = 1.f - ly;
  const float4* s0 = reinterpret_cast<const float4*>(src) + ((int64_t)n * H + y1) * W * C4;
  const float4* s1 = s0 + (int64_t)yp * W * C4;
  float4* drow = reinterpret_cast<float4*>(dst) + ((int64_t)n * Ho + y) * Wo * C4;
  const int row = Wo * C4;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < row; i += gridDim.x * blockDim.x) {
    const int x = i / C4, c = i - x * C4;
    float sx = 0.5f * ((float)x + 0.5f) - 0.5f;
    sx = sx < 0.f ? 0.f : sx;
    const int x1 = (int)sx;
    const int xp = x1 < W - 1 ? C4 : 0;
    const float lx = sx - (float)x1, hx = 1.f - lx;
    const int o0 = x1 * C4 + c;
    const float4 a = __ldg(s0 + o0), b = __ldg(s0 + o0 + xp), d = __ldg(s1 + o0), e = __ldg(s1 + o0 + xp);
    float4 o;
    o.x = hy * (hx * a.x + lx * b.x) + ly * (hx * d.x + lx * e.x);
    o.y = hy * (hx * a.y + lx * b.y) + ly * (hx * d.y + lx * e.y);
    o.z = hy * (hx * a.z + lx * b.z) + ly * (hx * d.z + lx * e.z);
    o.w = hy * (hx * a.w + lx * b.w) + ly * (hx * d.w + lx * e.w);
    drow[i] = o;
  }
}

__global__ void add_flow_tiled_kernel(const float* __restrict__ offset, const float* __restrict__ flow2,
                                      float* __restrict__ out, int64_t npix, int C4) {
  const int64_t total = npix * C4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t p = i / C4;
    const float2 f = __ldg(reinterpret_cast<const float2*>(flow2) + p);
    float4 v = __ldg(reinterpret_cast<const float4*>(offset) + i);
    v.x += f.x; v.y += f.y; v.z += f.x; v.w += f.y;
    reinterpret_cast<float4*>(out)[i] = v;
  }
}

static int grid_for(int64_t items) {
  int64_t b = (items + 255) / 256;
  const int64_t cap = (int64_t)kNumSMs * 8;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace tdvc

using namespace tdvc;

extern "C" int tdvc_avgpool2x2(const float* src, float* dst, int N, int H, int W, int C, void* stream) {
  TDVC_REQUIRE(src && dst && N > 0 && H >= 2 && W >= 2 && H % 2 == 0 && W % 2 == 0 && C % 4 == 0, "avgpool2x2: bad args");
  avgpool2x2_kernel<<<grid_for((int64_t)N * (H / 2) * (W / 2) * (C / 4)), 256, 0, (cudaStream_t)stream>>>(src, dst, N, H, W, C / 4);
  TDVC_CHECK_LAUNCH("avgpool2x2");
  return TDVC_OK;
}

extern "C" int tdvc_spynet_prep(const float* ref4, const float* supp4, const float* flow_prev, float* out8,
                                int N, int h, int w, void* stream) {
  TDVC_REQUIRE(ref4 && supp4 && out8 && N > 0 && h >= 2 && w >= 2, "spynet_prep: bad args");
  TDVC_REQUIRE(flow_prev == nullptr || (h % 2 == 0 && w % 2 == 0), "spynet_prep: odd level size");
  const int tiles_x = cdiv(w, SP_TW), tiles_y = cdiv(h, SP_TH);
  const int64_t blocks = (int64_t)N * tiles_x * tiles_y;
  TDVC_REQUIRE(blocks < (1ll << 31), "spynet_prep: too many tiles");
  spynet_prep_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(ref4, supp4, flow_prev, out8, N, h, w, tiles_x, tiles_y);
  TDVC_CHECK_LAUNCH("spynet_prep");
  return TDVC_OK;
}

extern "C" int tdvc_spynet_prep_backward(const float* supp4, const float* flow_prev, const float* grad_out8, float* grad_fine_ws,
                                         float* grad_flow_prev, int N, int h, int w, void* stream) {
  TDVC_REQUIRE(supp4 && flow_prev && grad_out8 && grad_fine_ws && grad_flow_prev && N > 0 && h >= 2 && w >= 2 && h % 2 == 0 && w % 2 == 0,
               "spynet_prep_backward: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  int grid = cdiv((int64_t)N * h * w, 256);
  if (grid > kNumSMs * 8) grid = kNumSMs * 8;
  spynet_prep_bwd_fine_kernel<<<grid, 256, 0, st>>>(supp4, flow_prev, grad_out8, grad_fine_ws, N, h, w);
  TDVC_CHECK_LAUNCH("spynet_prep_backward (fine)");
  int grid2 = cdiv((int64_t)N * (h / 2) * (w / 2), 128);
  if (grid2 > kNumSMs * 16) grid2 = kNumSMs * 16;
  spynet_prep_bwd_coarse_kernel<<<grid2, 128, 0, st>>>(grad_fine_ws, grad_flow_prev, N, h, w);
  TDVC_CHECK_LAUNCH("spynet_prep_backward (coarse)");
  return TDVC_OK;
}

extern "C" int tdvc_upsample2x(const float* src, float* dst, int N, int H, int W, int C, void* stream) {
  TDVC_REQUIRE(src && dst && N > 0 && H > 0 && W > 0 && C % 4 == 0, "upsample2x: bad args");
  TDVC_REQUIRE((int64_t)N * 2 * H <= 65535 && (int64_t)2 * W * (C / 4) < (1ll << 30), "upsample2x: image too large");
  const int row = 2 * W * (C / 4);
  dim3 grid((row + 255) / 256 > 8 ? 8 : (row + 255) / 256, N * 2 * H);
  upsample2x_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, dst, N, H, W, C / 4);
  TDVC_CHECK_LAUNCH("upsample2x");
  return TDVC_OK;
}

extern "C" int tdvc_add_flow_tiled(const float* offset, const float* flow2, float* out, int N, int H, int W, int C, void* stream) {
  TDVC_REQUIRE(offset && flow2 && out && N > 0 && H > 0 && W > 0 && C % 4 == 0, "add_flow_tiled: bad args");
  add_flow_tiled_kernel<<<grid_for((int64_t)N * H * W * (C / 4)), 256, 0, (cudaStream_t)stream>>>(offset, flow2, out, (int64_t)N * H * W, C / 4);
  TDVC_CHECK_LAUNCH("add_flow_tiled");
  return TDVC_OK;
}
