// Backward pieces of the convolutions (the training step's dgrad / wgrad; reference tools/train.py:125-159
// takes them from cuDNN through autograd).
//
// * dgrad needs no kernel of its own: d x = conv(d y, W^T flipped) is a stride-1 convolution of the output gradient with
//   the transposed, spatially flipped weight (pad' = k - 1 - pad), so it runs on the forward kernels (tcgen05 hi/lo split
//   included) from a re-packed weight.  A stride-s layer first spreads its output gradient over the input grid
//   (`zero_insert`: g_up[oy * s][ox * s] = g[oy][ox], zeros elsewhere).
// * `act_backward`: g * f'(y) from the layer's OUTPUT (ReLU / LeakyReLU / clamp keep the sign information there).
// * `wgrad`: d W[co][ci][ky][kx] = sum over pixels of x[p + tap][ci] * g[p][co]: a GEMM whose reduction runs over the pixels.
//   The pixel range is cut into S slabs, a CTA owns (slab, tap, 64 ci, 64 co) and contracts staged 32-pixel chunks with
//   warp-level TF32 MMAs on (hi, lo) operand pairs (three products: fp32-class); partial sums go to a workspace and a second
//   kernel adds the S partials in a fixed order (deterministic; the bias gradient rides along).  (A first exact-FFMA version
//   with a 4x4 register tile per thread spent 250 of the 436 ms of a training step here.  A tcgen05 / TMEM version - kind::tf32,
//   M = 128 rows of (tap, ci), D resident in TMEM over the slab, operands transposed into K-major rows by 4-byte cp.async -
//   was written, gave the right numbers and was dropped: 1.16 ms against 0.85 ms for 64->64 3x3 at 8x256x256, the transposing
//   producers issue one copy per element and the MMA pipe sat at 4 %.  What works is feeding the channels-last rows to the MMA
//   as MN-major operands, with no transposition at all: wgrad_tc.cu, which takes the one-product (enabled_amp) calls of every
//   stride-1 1x1 / 3x3 / 7x7 layer; this file keeps the three-product mode and the remaining shapes.  DESIGN.md section 7.)
#include <cstdlib>
#include "common.cuh"

namespace tdvc {

__global__ void act_backward_kernel(const float* __restrict__ y, const float* __restrict__ g, float* __restrict__ out, int64_t n,
                                    int act, float slope) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float v = y[i], gv = g[i];
    float d = 1.f;
    if (act == TDVC_ACT_RELU) d = v > 0.f ? 1.f : 0.f;
    else if (act == TDVC_ACT_LRELU) d = v > 0.f ? 1.f : slope;
    else if (act == TDVC_ACT_CLAMP01) d = (v > 0.f && v < 1.f) ? 1.f : 0.f;
    out[i] = gv * d;
  }
}

// out (N, H, W, C) <- g (N, Ho, Wo, C): out[n][oy*s][ox*s] = g[n][oy][ox], zero elsewhere.  One float4 per thread.
__global__ void zero_insert_kernel(const float* __restrict__ g, int g_ld, float* __restrict__ out, int out_ld, int N, int H, int W,
                                   int Ho, int Wo, int C4, int s) {
  const int64_t total = (int64_t)N * H * W * C4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int c4 = (int)(i % C4);
    int64_t p = i / C4;
    const int x = (int)(p % W);
    p /= W;
    const int y = (int)(p % H);
    const int n = (int)(p / H);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (y % s == 0 && x % s == 0 && y / s < Ho && x / s < Wo)
      v = *reinterpret_cast<const float4*>(g + (((int64_t)n * Ho + y / s) * Wo + x / s) * g_ld + 4 * c4);
    *reinterpret_cast<float4*>(out + (((int64_t)n * H + y) * W + x) * out_ld + 4 * c4) = v;
  }
}

// GDN / IGDN backward, element-wise parts (compressai GDN: out = x * norm^(-1/2), inverse: x * norm^(+1/2), norm = beta +
// gamma . x^2).  pre: the direct term d_x = g * norm^(-+1/2) and d_norm = -+1/2 * g * x * norm^(-3/2 | -1/2); the caller then
// sends d_norm through the 1x1 convolution's dgrad (-> d(x^2)) and wgrad (in_square -> d gamma, d beta); post: d_x += 2 x d(x^2).
__global__ void gdn_bwd_pre_kernel(const float* __restrict__ x, const float* __restrict__ norm, const float* __restrict__ g,
                                   float* __restrict__ dx_direct, float* __restrict__ dnorm, int64_t n4, int inverse) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 xv = __ldg(reinterpret_cast<const float4*>(x) + i), nv = __ldg(reinterpret_cast<const float4*>(norm) + i),
                 gv = __ldg(reinterpret_cast<const float4*>(g) + i);
    const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, ns[4] = {nv.x, nv.y, nv.z, nv.w}, gs[4] = {gv.x, gv.y, gv.z, gv.w};
    float a[4], b[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float r = rsqrtf(ns[j]);                      // norm^(-1/2)
      if (inverse) {
        a[j] = gs[j] * (ns[j] * r);                       // g * sqrt(norm)
        b[j] = 0.5f * gs[j] * xs[j] * r;                  // +1/2 g x norm^(-1/2)
      } else {
        a[j] = gs[j] * r;
        b[j] = -0.5f * gs[j] * xs[j] * (r * r * r);       // -1/2 g x norm^(-3/2)
      }
    }
    reinterpret_cast<float4*>(dx_direct)[i] = make_float4(a[0], a[1], a[2], a[3]);
    reinterpret_cast<float4*>(dnorm)[i] = make_float4(b[0], b[1], b[2], b[3]);
  }
}
__global__ void gdn_bwd_post_kernel(const float* __restrict__ dx_direct, const float* __restrict__ x, const float* __restrict__ dxsq,
                                    float* __restrict__ dx, int64_t n4) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(dx_direct) + i), xv = __ldg(reinterpret_cast<const float4*>(x) + i),
                 q = __ldg(reinterpret_cast<const float4*>(dxsq) + i);
    reinterpret_cast<float4*>(dx)[i] = make_float4(fmaf(2.f * xv.x, q.x, a.x), fmaf(2.f * xv.y, q.y, a.y),
                                                   fmaf(2.f * xv.z, q.z, a.z), fmaf(2.f * xv.w, q.w, a.w));
  }
}

// ---- per-(image, channel) pieces of the squeeze-excitation layer in training (reference inflate.py:159-208):
// chan_affine: out[n][p][c] = a[n][p][c] * s[n][c] + t[n][c]   (a == nullptr: out = s broadcast; t may be nullptr).
//   forward of x * gate, its grad_input (g * gate), and the backward of the spatial mean (a broadcast).
// chan_dot: out[n][c] = sum_p a[n][p][c] * b[n][p][c]  (b == nullptr: plain sums): the mean and the gate's gradient; two stages,
//   fixed order: deterministic.  NHWC, dense rows of C channels (C % 4 == 0).
__global__ void chan_affine_kernel(const float* __restrict__ a, const float* __restrict__ s, const float* __restrict__ t,
                                   float* __restrict__ out, int64_t HW, int C4) {
  const int n = blockIdx.y;
  const int64_t total = HW * C4;
  const float4* a4 = a != nullptr ? reinterpret_cast<const float4*>(a) + (int64_t)n * total : nullptr;
  const float4* s4 = reinterpret_cast<const float4*>(s) + (int64_t)n * C4;
  const float4* t4 = t != nullptr ? reinterpret_cast<const float4*>(t) + (int64_t)n * C4 : nullptr;
  float4* o4 = reinterpret_cast<float4*>(out) + (int64_t)n * total;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4);
    float4 v = __ldg(s4 + c);
    if (a4 != nullptr) {
      const float4 x = __ldg(a4 + i);
      v = make_float4(x.x * v.x, x.y * v.y, x.z * v.z, x.w * v.w);
    }
    if (t4 != nullptr) {
      const float4 y = __ldg(t4 + c);
      v = make_float4(v.x + y.x, v.y + y.y, v.z + y.z, v.w + y.w);
    }
    o4[i] = v;
  }
}

__global__ void __launch_bounds__(256) chan_dot_partial_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t HW,
                                                               int C4, float* __restrict__ partial) {
  extern __shared__ float4 cd_sh[];
  const int lanes_p = 256 / C4;
  const int c = threadIdx.x % C4, pl = threadIdx.x / C4;
  const int n = blockIdx.y, blk = blockIdx.x, nblk = gridDim.x;
  const int64_t chunk = (HW + nblk - 1) / nblk;
  const int64_t p0 = (int64_t)blk * chunk, p1 = p0 + chunk < HW ? p0 + chunk : HW;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (pl < lanes_p) {
    const float4* a4 = reinterpret_cast<const float4*>(a) + (int64_t)n * HW * C4 + c;
    const float4* b4 = b != nullptr ? reinterpret_cast<const float4*>(b) + (int64_t)n * HW * C4 + c : nullptr;
    for (int64_t p = p0 + pl; p < p1; p += lanes_p) {
      float4 x = __ldg(a4 + p * C4);
      if (b4 != nullptr) {
        const float4 y = __ldg(b4 + p * C4);
        x = make_float4(x.x * y.x, x.y * y.y, x.z * y.z, x.w * y.w);
      }
      acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w;
    }
  }
  cd_sh[threadIdx.x] = acc;
  __syncthreads();
  if (pl == 0) {
    for (int j = 1; j < lanes_p; ++j) {
      const float4 v = cd_sh[j * C4 + c];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    reinterpret_cast<float4*>(partial)[((int64_t)blk * gridDim.y + n) * C4 + c] = acc;
  }
}
__global__ void chan_dot_finish_kernel(const float* __restrict__ partial, int nblk, int NC, float scale, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= NC) return;
  float v = 0.f;
  for (int b = 0; b < nblk; ++b) v += partial[(int64_t)b * NC + i];
  out[i] = v * scale;
}

constexpr int WG_T = 64;      // ci x co tile of a CTA
constexpr int WG_P = 32;      // output pixels per staged chunk
constexpr int WG_LD = WG_T + 8;   // row pitch = 8 banks: the (k = lane & 3, m = lane >> 2) fragment reads are conflict-free

struct WgradArgs {
  const float* x; int x_ld;
  const float* g; int g_ld;
  int N, H, W, Ho, Wo, cin, cout, k, stride, pad;
  int S, ci_tiles, co_tiles;
  int in_square;      // x is squared on load: the weight gradient of a GDN's 1x1 convolution of x^2
  float* part;        // [S][k*k][cin][cout]
  float* part_bias;   // [S][cout] or nullptr
};

// fp32 -> (hi, lo) TF32 pair: hi = rna(v), lo = rna(v - hi); hi*hi' + lo*hi' + hi*lo' carries ~21 significand bits
__device__ __forceinline__ void tf32_split(float v, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(v));
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(v - __uint_as_float(hi)));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// One CTA = (tap, 64 ci, 64 co, pixel slab).  The reduction dimension of this GEMM is the PIXEL index, so both operands are
// read "transposed" from their channels-last rows: chunks of 32 pixels x 64 channels of x (shifted by the tap) and of grad_y
// are staged in shared memory and contracted with warp-level m16n8k8 TF32 MMAs, each fp32 operand split into a TF32 (hi, lo)
// pair and three products issued (fp32-class accuracy: the weight gradient sums ~10^5..10^6 terms).  8 warps: warp w owns ci
// rows 16 (w & 3) .. +16 and co columns 32 (w >> 2) .. +32 (four n-tiles).
template <int PRODUCTS>
__global__ void __launch_bounds__(256) wgrad_kernel(const WgradArgs a) {
  __shared__ __align__(16) float xs[WG_P][WG_LD];
  __shared__ __align__(16) float gs[WG_P][WG_LD];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int gq = lane >> 2, tq = lane & 3;
  const int kk = a.k * a.k;
  int r = blockIdx.x;            // the taps of one tile are neighbours in launch order: they read the same slab through L2
  const int tap = r % kk; r /= kk;
  const int co_t = r % a.co_tiles;
  const int ci_t = r / a.co_tiles;
  const int ky = tap / a.k, kx = tap - ky * a.k;
  const int s = blockIdx.y;
  const int64_t npix = (int64_t)a.N * a.Ho * a.Wo;
  const int64_t p_begin = npix * s / a.S, p_end = npix * (s + 1) / a.S;
  const int ci0 = ci_t * WG_T, co0 = co_t * WG_T;
  const bool do_bias = a.part_bias != nullptr && tap == 0 && ci_t == 0 && warp == 0;
  const int m0 = 16 * (warp & 3), n0 = 32 * (warp >> 2);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float bsum0 = 0.f, bsum1 = 0.f;
  // global -> registers for one chunk (two float4 of x and of g per thread); the loads of chunk i + 1 are in flight while the
  // MMAs of chunk i run
  // (the pixel coordinates of a thread's two staging slots advance by WG_P per chunk: tracked incrementally, no divisions)
  int sn[2], sy[2], sx[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int64_t p = p_begin + ((tid + 256 * h) >> 4);
    sx[h] = (int)(p % a.Wo);
    const int64_t q = p / a.Wo;
    sy[h] = (int)(q % a.Ho);
    sn[h] = (int)(q / a.Ho);
  }
  auto fetch = [&](int64_t pc, float4 (&xr)[2], float4 (&gr)[2]) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int i = tid + 256 * h;
      const int pl = i >> 4, c4 = (i & 15) * 4;
      const int64_t p = pc + pl;
      float4 xv = make_float4(0.f, 0.f, 0.f, 0.f), gv = xv;
      if (p < p_end) {
        const int ox = sx[h], oy = sy[h], n = sn[h];
        const int iy = oy * a.stride + ky - a.pad, ix = ox * a.stride + kx - a.pad;
        if (iy >= 0 && iy < a.H && ix >= 0 && ix < a.W) {
          const float* src = a.x + (((int64_t)n * a.H + iy) * a.W + ix) * a.x_ld;
          const int c = ci0 + c4;
          if (c + 3 < a.cin) xv = __ldg(reinterpret_cast<const float4*>(src + c));
          else {
            if (c < a.cin) xv.x = src[c];
            if (c + 1 < a.cin) xv.y = src[c + 1];
            if (c + 2 < a.cin) xv.z = src[c + 2];
          }
        }
        const float* gsrc = a.g + p * a.g_ld;
        const int c = co0 + c4;
        if (c + 3 < a.cout) gv = __ldg(reinterpret_cast<const float4*>(gsrc + c));
        else {
          if (c < a.cout) gv.x = gsrc[c];
          if (c + 1 < a.cout) gv.y = gsrc[c + 1];
          if (c + 2 < a.cout) gv.z = gsrc[c + 2];
        }
      }
      xr[h] = xv;
      gr[h] = gv;
      sx[h] += WG_P;
      while (sx[h] >= a.Wo) {
        sx[h] -= a.Wo;
        if (++sy[h] == a.Ho) { sy[h] = 0; ++sn[h]; }
      }
    }
  };
  float4 xr[2], gr[2];
  fetch(p_begin, xr, gr);
  for (int64_t pc = p_begin; pc < p_end; pc += WG_P) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int i = tid + 256 * h;
      float4 xv = xr[h];   // (squared here, not at the load: the prefetched registers must not be touched before they are needed)
      if (a.in_square) { xv.x *= xv.x; xv.y *= xv.y; xv.z *= xv.z; xv.w *= xv.w; }
      *reinterpret_cast<float4*>(&xs[i >> 4][(i & 15) * 4]) = xv;
      *reinterpret_cast<float4*>(&gs[i >> 4][(i & 15) * 4]) = gr[h];
    }
    __syncthreads();
    if (pc + WG_P < p_end) fetch(pc + WG_P, xr, gr);
#pragma unroll
    for (int k0 = 0; k0 < WG_P; k0 += 8) {
      // A fragment (16 ci x 8 pixels): element (row m, col k) = xs[k][m]; B fragments (8 pixels x 8 co): (row k, col n) = gs[k][n]
      if (PRODUCTS == 1) {   // one TF32 product (what cuDNN runs by default for fp32 training convolutions)
        uint32_t af[4], lo_;
        tf32_split(xs[k0 + tq][m0 + gq], af[0], lo_);
        tf32_split(xs[k0 + tq][m0 + gq + 8], af[1], lo_);
        tf32_split(xs[k0 + tq + 4][m0 + gq], af[2], lo_);
        tf32_split(xs[k0 + tq + 4][m0 + gq + 8], af[3], lo_);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          uint32_t b0, b1;
          tf32_split(gs[k0 + tq][n0 + 8 * nt + gq], b0, lo_);
          tf32_split(gs[k0 + tq + 4][n0 + 8 * nt + gq], b1, lo_);
          mma_tf32(acc[nt], af, b0, b1);
        }
      } else {
        uint32_t ah[4], al[4];
        tf32_split(xs[k0 + tq][m0 + gq], ah[0], al[0]);
        tf32_split(xs[k0 + tq][m0 + gq + 8], ah[1], al[1]);
        tf32_split(xs[k0 + tq + 4][m0 + gq], ah[2], al[2]);
        tf32_split(xs[k0 + tq + 4][m0 + gq + 8], ah[3], al[3]);
        uint32_t bh0[4], bl0[4], bh1[4], bl1[4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          tf32_split(gs[k0 + tq][n0 + 8 * nt + gq], bh0[nt], bl0[nt]);
          tf32_split(gs[k0 + tq + 4][n0 + 8 * nt + gq], bh1[nt], bl1[nt]);
        }
        // consecutive MMAs go to different accumulators (an accumulator's three products are 4 issues apart)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) mma_tf32(acc[nt], al, bh0[nt], bh1[nt]);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) mma_tf32(acc[nt], ah, bl0[nt], bl1[nt]);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) mma_tf32(acc[nt], ah, bh0[nt], bh1[nt]);
      }
    }
    if (do_bias) {
#pragma unroll 8
      for (int pl = 0; pl < WG_P; ++pl) {
        bsum0 += gs[pl][lane];
        bsum1 += gs[pl][lane + 32];
      }
    }
    __syncthreads();
  }
  float* dst = a.part + ((int64_t)s * kk + tap) * a.cin * a.cout;
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int ci = ci0 + m0 + gq + ((e & 2) ? 8 : 0);
      const int co = co0 + n0 + 8 * nt + 2 * tq + (e & 1);
      if (ci < a.cin && co < a.cout) dst[(int64_t)ci * a.cout + co] = acc[nt][e];
    }
  }
  if (do_bias) {
    if (co0 + lane < a.cout) a.part_bias[(int64_t)s * a.cout + co0 + lane] = bsum0;
    if (co0 + lane + 32 < a.cout) a.part_bias[(int64_t)s * a.cout + co0 + lane + 32] = bsum1;
  }
}

// grad_w[co][ci][ky][kx] (the layout of nn.Conv2d.weight) = sum_s part[s][tap][ci][co], s ascending; grad_b likewise
__global__ void wgrad_reduce_kernel(const float* __restrict__ part, const float* __restrict__ part_bias, int S, int kk, int cin,
                                    int cout, float* __restrict__ grad_w, float* __restrict__ grad_b) {
  const int64_t nw = (int64_t)kk * cin * cout;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nw) {
    const int co = (int)(i % cout);
    const int64_t r = i / cout;
    const int ci = (int)(r % cin), tap = (int)(r / cin);
    float v = 0.f;
    for (int s = 0; s < S; ++s) v += part[(int64_t)s * nw + i];
    grad_w[((int64_t)co * cin + ci) * kk + tap] = v;
  } else if (grad_b != nullptr && i < nw + cout) {
    const int co = (int)(i - nw);
    float v = 0.f;
    for (int s = 0; s < S; ++s) v += part_bias[(int64_t)s * cout + co];
    grad_b[co] = v;
  }
}

// nn.Conv2d.weight (O, I, k, k) -> [k*k][cin_pad][cout_pad], zero padded; transposed: the dgrad operator (flipped, channels swapped)
__global__ void pack_weight_kernel(const float* __restrict__ w, int O, int I, int kk, int transposed, float* __restrict__ out,
                                   int cin_pad, int cout_pad, const float* __restrict__ bias, float* __restrict__ bias_out) {
  const int64_t nw = (int64_t)kk * cin_pad * cout_pad;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nw) {
    const int co = (int)(i % cout_pad);
    const int64_t r = i / cout_pad;
    const int ci = (int)(r % cin_pad), tap = (int)(r / cin_pad);
    float v = 0.f;
    if (!transposed) {
      if (co < O && ci < I) v = w[((int64_t)co * I + ci) * kk + tap];
    } else {
      if (ci < O && co < I) v = w[((int64_t)ci * I + co) * kk + (kk - 1 - tap)];
    }
    out[i] = v;
  } else if (bias_out != nullptr && i < nw + cout_pad) {
    const int co = (int)(i - nw);
    bias_out[co] = (bias != nullptr && co < O) ? bias[co] : 0.f;
  }
}

static int wgrad_slabs(int64_t npix, int tiles) {
  int S = (4 * kNumSMs + tiles - 1) / tiles;           // ~4 CTAs per SM in total
  const int64_t max_s = (npix + 4 * WG_P - 1) / (4 * WG_P);   // at least 4 chunks per slab
  if (S > max_s) S = (int)max_s;
  return S < 1 ? 1 : S;
}

// wgrad_tc.cu: the tcgen05 kernels for stride-1 layers (one TF32 product per MAC, or three: fp32-class)
bool wgrad_tc_eligible(const float* x, int x_ld, const float* g, int g_ld, int N, int H, int W, int cin, int cout, int k, int stride,
                       int pad, int in_square, int products);
size_t wgrad_tc_workspace_bytes(int N, int H, int W, int cin, int cout, int k);
int wgrad_tc_launch(const float* x, int x_ld, const float* g, int g_ld, int N, int H, int W, int cin, int cout, int k, int products,
                    float* workspace, float** part_bias_out, int* S_out, cudaStream_t st);

}  // namespace tdvc

using namespace tdvc;

extern "C" int tdvc_act_backward(const float* y, const float* grad_y, float* grad_pre, int64_t n, int act, float slope, void* stream) {
  TDVC_REQUIRE(y && grad_y && grad_pre && n >= 0, "act_backward: bad args");
  TDVC_REQUIRE(act >= TDVC_ACT_NONE && act <= TDVC_ACT_CLAMP01, "act_backward: activation %d", act);
  if (n == 0) return TDVC_OK;
  int grid = cdiv(n, 256);
  if (grid > kNumSMs * 8) grid = kNumSMs * 8;
  act_backward_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(y, grad_y, grad_pre, n, act, slope);
  TDVC_CHECK_LAUNCH("act_backward");
  return TDVC_OK;
}

extern "C" int tdvc_zero_insert(const float* g, int g_ld, float* out, int out_ld, int N, int H, int W, int Ho, int Wo, int C,
                                int stride, void* stream) {
  TDVC_REQUIRE(g && out && N > 0 && H > 0 && W > 0 && Ho > 0 && Wo > 0 && C > 0 && stride >= 1, "zero_insert: bad args");
  TDVC_REQUIRE(C % 4 == 0 && g_ld % 4 == 0 && out_ld % 4 == 0 && g_ld >= C && out_ld >= C, "zero_insert: channels / ld must be multiples of 4");
  TDVC_REQUIRE((Ho - 1) * stride < H && (Wo - 1) * stride < W, "zero_insert: the gradient does not fit the input grid");
  int grid = cdiv((int64_t)N * H * W * (C / 4), 256);
  if (grid > kNumSMs * 16) grid = kNumSMs * 16;
  zero_insert_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(g, g_ld, out, out_ld, N, H, W, Ho, Wo, C / 4, stride);
  TDVC_CHECK_LAUNCH("zero_insert");
  return TDVC_OK;
}

extern "C" int tdvc_conv2d_pack_weight(const float* w, int O, int I, int k, int transposed, float* out, int cin_pad, int cout_pad,
                                       const float* bias_or_null, float* bias_out_or_null, void* stream) {
  TDVC_REQUIRE(w && out && O > 0 && I > 0 && k >= 1 && k <= 7, "conv2d_pack_weight: bad args");
  TDVC_REQUIRE(transposed ? (cin_pad >= O && cout_pad >= I) : (cin_pad >= I && cout_pad >= O), "conv2d_pack_weight: padded sizes");
  TDVC_REQUIRE(!(transposed && bias_out_or_null), "conv2d_pack_weight: the dgrad operator has no bias");
  const int64_t n = (int64_t)k * k * cin_pad * cout_pad + (bias_out_or_null ? cout_pad : 0);
  pack_weight_kernel<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(w, O, I, k * k, transposed ? 1 : 0, out, cin_pad, cout_pad,
                                                                     bias_or_null, bias_out_or_null);
  TDVC_CHECK_LAUNCH("conv2d_pack_weight");
  return TDVC_OK;
}

extern "C" size_t tdvc_conv2d_wgrad_workspace_bytes(int N, int Ho, int Wo, int cin, int cout, int k) {
  if (N <= 0 || Ho <= 0 || Wo <= 0 || cin <= 0 || cout <= 0 || k <= 0) return 0;
  const int tiles = k * k * cdiv(cin, WG_T) * cdiv(cout, WG_T);
  const int S = wgrad_slabs((int64_t)N * Ho * Wo, tiles);
  size_t need = (size_t)S * ((size_t)k * k * cin * cout + cout) * sizeof(float);
  const size_t tc = wgrad_tc_workspace_bytes(N, Ho, Wo, cin, cout, k);   // the tensor-core kernels cut their slabs differently; cover both
  if (tc > need) need = tc;
  return need;
}

extern "C" int tdvc_conv2d_wgrad(const float* x, int x_ld, const float* grad_y, int g_ld, int N, int H, int W, int cin, int cout,
                                 int k, int stride, int pad, int in_square, int products, float* grad_w, float* grad_b_or_null,
                                 void* workspace, size_t workspace_bytes, void* stream) {
  TDVC_REQUIRE(products == 1 || products == 3, "conv2d_wgrad: products must be 1 (TF32) or 3 (fp32-class TF32 split)");
  TDVC_REQUIRE(x && grad_y && grad_w && N > 0 && H > 0 && W > 0 && cin > 0 && cout > 0, "conv2d_wgrad: bad args");
  TDVC_REQUIRE(k >= 1 && k <= 7 && stride >= 1 && pad >= 0, "conv2d_wgrad: k=%d stride=%d pad=%d", k, stride, pad);
  TDVC_REQUIRE(x_ld >= cin && g_ld >= cout && x_ld % 4 == 0 && g_ld % 4 == 0, "conv2d_wgrad: leading dimensions");
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  TDVC_REQUIRE(Ho > 0 && Wo > 0, "conv2d_wgrad: empty output");
  const size_t need = tdvc_conv2d_wgrad_workspace_bytes(N, Ho, Wo, cin, cout, k);
  TDVC_REQUIRE(workspace != nullptr && workspace_bytes >= need, "conv2d_wgrad: workspace of %zu bytes, %zu needed", workspace_bytes, need);
  WgradArgs a;
  a.x = x; a.x_ld = x_ld; a.g = grad_y; a.g_ld = g_ld;
  a.N = N; a.H = H; a.W = W; a.Ho = Ho; a.Wo = Wo; a.cin = cin; a.cout = cout; a.k = k; a.stride = stride; a.pad = pad;
  a.in_square = in_square ? 1 : 0;
  a.ci_tiles = cdiv(cin, WG_T);
  a.co_tiles = cdiv(cout, WG_T);
  const int tiles = k * k * a.ci_tiles * a.co_tiles;
  a.S = wgrad_slabs((int64_t)N * Ho * Wo, tiles);
  a.part = (float*)workspace;
  a.part_bias = a.part + (size_t)a.S * k * k * cin * cout;
  cudaStream_t st = (cudaStream_t)stream;
  static const bool log_shapes = getenv("TDVC_B200_WGRAD_LOG") != nullptr;   // developer: one line per call (tools/train_profile.py)
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  if (log_shapes) {
    cudaEventCreate(&ev0);
    cudaEventCreate(&ev1);
    cudaEventRecord(ev0, st);
  }
  static const bool tc3 = getenv("TDVC_B200_WGRAD_TC3_OFF") == nullptr;   // developer A/B switch for the three-product tcgen05 mode
  if ((products == 1 || tc3) && wgrad_tc_eligible(x, x_ld, grad_y, g_ld, N, H, W, cin, cout, k, stride, pad, in_square, products)) {
    if (int rc = wgrad_tc_launch(x, x_ld, grad_y, g_ld, N, H, W, cin, cout, k, products, a.part, &a.part_bias, &a.S, st)) return rc;
  } else if (products == 1) wgrad_kernel<1><<<dim3(tiles, a.S), 256, 0, st>>>(a);
  else wgrad_kernel<3><<<dim3(tiles, a.S), 256, 0, st>>>(a);
  TDVC_CHECK_LAUNCH("conv2d_wgrad");
  const int64_t n_out = (int64_t)k * k * cin * cout + cout;
  wgrad_reduce_kernel<<<cdiv(n_out, 256), 256, 0, st>>>(a.part, a.part_bias, a.S, k * k, cin, cout, grad_w, grad_b_or_null);
  TDVC_CHECK_LAUNCH("conv2d_wgrad_reduce");
  if (log_shapes) {
    float ms = 0.f;
    cudaEventRecord(ev1, st);
    cudaEventSynchronize(ev1);
    cudaEventElapsedTime(&ms, ev0, ev1);
    fprintf(stderr, "wgrad N%d %dx%d cin%d cout%d k%d s%d p%d sq%d products%d ld%d/%d ms %.4f\n", N, H, W, cin, cout, k, stride, pad,
            in_square, products, x_ld, g_ld, ms);
    cudaEventDestroy(ev0);
    cudaEventDestroy(ev1);
  }
  return TDVC_OK;
}

extern "C" int tdvc_gdn_backward_pre(const float* x, const float* norm, const float* grad_out, float* dx_direct, float* dnorm,
                                     int64_t n, int inverse, void* stream) {
  TDVC_REQUIRE(x && norm && grad_out && dx_direct && dnorm && n > 0 && n % 4 == 0, "gdn_backward_pre: bad args");
  int grid = cdiv(n / 4, 256);
  if (grid > kNumSMs * 8) grid = kNumSMs * 8;
  gdn_bwd_pre_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, norm, grad_out, dx_direct, dnorm, n / 4, inverse);
  TDVC_CHECK_LAUNCH("gdn_backward_pre");
  return TDVC_OK;
}

extern "C" int tdvc_gdn_backward_post(const float* dx_direct, const float* x, const float* dxsq, float* dx, int64_t n, void* stream) {
  TDVC_REQUIRE(dx_direct && x && dxsq && dx && n > 0 && n % 4 == 0, "gdn_backward_post: bad args");
  int grid = cdiv(n / 4, 256);
  if (grid > kNumSMs * 8) grid = kNumSMs * 8;
  gdn_bwd_post_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(dx_direct, x, dxsq, dx, n / 4);
  TDVC_CHECK_LAUNCH("gdn_backward_post");
  return TDVC_OK;
}

extern "C" int tdvc_chan_affine(const float* a_or_null, const float* s, const float* t_or_null, float* out, int N, int64_t HW, int C,
                                void* stream) {
  TDVC_REQUIRE(s && out && N > 0 && HW > 0 && C > 0 && C % 4 == 0, "chan_affine: bad args");
  int gx = cdiv(HW * (C / 4), 256);
  if (gx > kNumSMs * 8) gx = kNumSMs * 8;
  chan_affine_kernel<<<dim3(gx, N), 256, 0, (cudaStream_t)stream>>>(a_or_null, s, t_or_null, out, HW, C / 4);
  TDVC_CHECK_LAUNCH("chan_affine");
  return TDVC_OK;
}

extern "C" size_t tdvc_chan_dot_workspace_bytes(int N, int64_t HW, int C) {
  if (N <= 0 || HW <= 0 || C <= 0) return 0;
  int64_t nblk = HW / 64;
  nblk = nblk < 1 ? 1 : (nblk > 296 ? 296 : nblk);
  return (size_t)nblk * N * C * sizeof(float);
}

extern "C" int tdvc_chan_dot(const float* a, const float* b_or_null, float* out, int N, int64_t HW, int C, float scale,
                             void* workspace, size_t workspace_bytes, void* stream) {
  TDVC_REQUIRE(a && out && N > 0 && HW > 0 && C > 0 && C % 4 == 0 && C <= 1024, "chan_dot: bad args");
  const size_t need = tdvc_chan_dot_workspace_bytes(N, HW, C);
  TDVC_REQUIRE(workspace && workspace_bytes >= need, "chan_dot: workspace of %zu bytes, %zu needed", workspace_bytes, need);
  const int nblk = (int)(need / ((size_t)N * C * sizeof(float)));
  cudaStream_t st = (cudaStream_t)stream;
  chan_dot_partial_kernel<<<dim3(nblk, N), 256, 256 * sizeof(float4), st>>>(a, b_or_null, HW, C / 4, (float*)workspace);
  TDVC_CHECK_LAUNCH("chan_dot");
  chan_dot_finish_kernel<<<cdiv((int64_t)N * C, 256), 256, 0, st>>>((const float*)workspace, nblk, N * C, scale, out);
  TDVC_CHECK_LAUNCH("chan_dot_finish");
  return TDVC_OK;
}
