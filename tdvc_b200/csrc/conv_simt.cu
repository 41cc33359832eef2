// Generic NHWC fp32 convolution as an implicit GEMM on the CUDA cores (FFMA), with the fused
// prologue/epilogue the P-frame graph needs.  This is the exact-fp32 workhorse for every shape; the
// tcgen05 tensor-core kernel (conv_tc.cu) takes over the heavy 3x3 shapes.
//
// Tiling: one CTA (128 threads) computes an 8x16-pixel output tile x TN output channels.  Input channels
// are consumed in chunks of CK=8: the (8-1)*S+K by (16-1)*S+K input halo of the chunk is staged in shared
// memory channel-major ([c][pixel], conflict-free for the broadcast reads), one kernel row of weights
// ([K][CK][TN]) is staged next to it, and each thread accumulates an 8-pixel x (TN/8)-channel register
// tile.  Global reads are float4 (two lanes cover one 32-byte sector of a pixel's 8 channels); stores
// are float4 along channels (coalesced: 8 neighbouring threads write 256 contiguous bytes of a pixel).
//
// Replaces: nn.Conv2d / Conv3d(1,3,3) / 1x1 / MaskedConv2d / GDN mix call sites of
// reference main/model/pnet.py, main/utils/utils.py:43-56, main/model/flownet.py:187-227 and the
// compressai layers of SURVEY.md App. A; torch.cat inputs become multi-source reads.
#include "common.cuh"

namespace tdvc {

constexpr int CK = 8;
constexpr int TY = 8, TX = 16;
constexpr int CONV_THREADS = 128;

struct OutLoc {
  int64_t pix;  // output pixel index (already shuffled)
  int c;        // output channel
};

__device__ __forceinline__ OutLoc out_loc(const TdvcConvParams& p, int n, int oy, int ox, int co) {
  OutLoc l;
  if (p.shuffle == 2) {
    int cr = p.cout >> 2;
    int q = co / cr;
    l.c = co - q * cr;
    l.pix = ((int64_t)n * (2 * p.Ho) + (2 * oy + (q >> 1))) * (2 * p.Wo) + (2 * ox + (q & 1));
  } else {
    l.c = co;
    l.pix = ((int64_t)n * p.Ho + oy) * p.Wo + ox;
  }
  return l;
}

__device__ __forceinline__ float epilogue1(const TdvcConvParams& p, float v, const OutLoc& l) {
  if (p.post != TDVC_POST_NONE) {
    float m = __ldg(p.mul + l.pix * p.mul_ld + l.c);
    v = m * (p.post == TDVC_POST_GDN ? rsqrtf(v) : sqrtf(v));
  }
  v = apply_act(v, p.act, p.slope);
  if (p.res1) v += __ldg(p.res1 + l.pix * p.res1_ld + l.c);
  if (p.res2) v += __ldg(p.res2 + l.pix * p.res2_ld + l.c);
  return v;
}

template <int TN, int K, int S>
__global__ void __launch_bounds__(CONV_THREADS) conv_simt_kernel(const TdvcConvParams p, int tiles_x, int vec_ok) {
  constexpr int CPT = TN / 8;  // output channels per thread
  constexpr int IH = (TY - 1) * S + K;
  constexpr int IW = (TX - 1) * S + K;
  constexpr int IHW = IH * IW;
  constexpr int IHWP = (IHW + 3) & ~3;
  extern __shared__ __align__(16) float smem[];
  float* halo = smem;              // [CK][IHWP]
  float* wsm = smem + CK * IHWP;   // [K][CK][TN]

  const int tid = threadIdx.x;
  const int cg = tid & 7;
  const int pg = tid >> 3;
  const int py = pg >> 1;
  const int px0 = (pg & 1) * 8;
  const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
  const int oy0 = ty * TY, ox0 = tx * TX;
  const int ct = blockIdx.y;
  const int n = blockIdx.z;
  const int iy0 = oy0 * S - p.pad, ix0 = ox0 * S - p.pad;
  const int u = tid & 1;  // which float4 half of the 8-channel chunk this thread stages

  float acc[8][CPT];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < CPT; ++j) acc[i][j] = 0.f;

  const int nchunks = p.cin_pad / CK;
  for (int cc = 0; cc < nchunks; ++cc) {
    // resolve this thread's float4 unit to a source tensor
    const float* sp = nullptr;
    int sld = 0;
    {
      int c0 = cc * CK + u * 4;
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        if (s < p.n_src && sp == nullptr) {
          if (c0 < p.src_c[s]) {
            sp = p.src[s] + c0;
            sld = p.src_ld[s];
          } else {
            c0 -= p.src_c[s];
          }
        }
      }
    }
    __syncthreads();  // previous chunk fully consumed
    for (int idx = tid; idx < IHW * 2; idx += CONV_THREADS) {
      const int pix = idx >> 1;
      const int hy = pix / IW, hx = pix - hy * IW;
      const int iy = iy0 + hy, ix = ix0 + hx;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (sp != nullptr && iy >= 0 && iy < p.H && ix >= 0 && ix < p.W)
        v = __ldg(reinterpret_cast<const float4*>(sp + (((int64_t)n * p.H + iy) * p.W + ix) * sld));
      if (p.in_square) {
        v.x *= v.x; v.y *= v.y; v.z *= v.z; v.w *= v.w;
      }
      float* h = halo + (u * 4) * IHWP + pix;
      h[0] = v.x; h[IHWP] = v.y; h[2 * IHWP] = v.z; h[3 * IHWP] = v.w;
    }
#pragma unroll 1
    for (int ky = 0; ky < K; ++ky) {
      if (ky > 0) __syncthreads();
      const float* wg = p.weight + ((int64_t)(ky * K) * p.cin_pad + cc * CK) * p.cout_pad + ct * TN;
      for (int idx = tid; idx < K * CK * TN / 4; idx += CONV_THREADS) {
        const int e = idx * 4;
        const int co = e % TN;
        const int r = e / TN;
        const int c = r % CK, kx = r / CK;
        const float4 w = __ldg(reinterpret_cast<const float4*>(wg + ((int64_t)kx * p.cin_pad + c) * p.cout_pad + co));
        *reinterpret_cast<float4*>(wsm + e) = w;
      }
      __syncthreads();
#pragma unroll
      for (int kx = 0; kx < K; ++kx) {
#pragma unroll
        for (int c = 0; c < CK; ++c) {
          const float* hrow = halo + c * IHWP + (py * S + ky) * IW + px0 * S + kx;
          float a[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) a[i] = hrow[i * S];
          const float* wrow = wsm + (kx * CK + c) * TN + cg * CPT;
          float b[CPT];
          if constexpr (CPT % 4 == 0) {
#pragma unroll
            for (int j = 0; j < CPT; j += 4) {
              const float4 w4 = *reinterpret_cast<const float4*>(wrow + j);
              b[j] = w4.x; b[j + 1] = w4.y; b[j + 2] = w4.z; b[j + 3] = w4.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < CPT; ++j) b[j] = wrow[j];
          }
#pragma unroll
          for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < CPT; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
      }
    }
  }

  // ---- epilogue
  const int oy = oy0 + py;
  if (oy >= p.Ho) return;
  const int cobase = ct * TN + cg * CPT;
  float amax = 0.f;   // TdvcConvParams::out_absmax: max |v| over the values this thread stored
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int ox = ox0 + px0 + i;
    if (ox >= p.Wo) continue;
    if constexpr (CPT % 4 == 0) {
      if (vec_ok) {
#pragma unroll
        for (int j = 0; j < CPT; j += 4) {
          const int co = cobase + j;
          if (co >= p.cout) continue;
          const OutLoc l = out_loc(p, n, oy, ox, co);
          float4 v = make_float4(acc[i][j], acc[i][j + 1], acc[i][j + 2], acc[i][j + 3]);
          if (p.bias) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + co));
            v.x += b4.x; v.y += b4.y; v.z += b4.z; v.w += b4.w;
          }
          if (p.post != TDVC_POST_NONE) {
            const float4 m = __ldg(reinterpret_cast<const float4*>(p.mul + l.pix * p.mul_ld + l.c));
            if (p.post == TDVC_POST_GDN) {
              v.x = m.x * rsqrtf(v.x); v.y = m.y * rsqrtf(v.y); v.z = m.z * rsqrtf(v.z); v.w = m.w * rsqrtf(v.w);
            } else {
              v.x = m.x * sqrtf(v.x); v.y = m.y * sqrtf(v.y); v.z = m.z * sqrtf(v.z); v.w = m.w * sqrtf(v.w);
            }
          }
          v.x = apply_act(v.x, p.act, p.slope); v.y = apply_act(v.y, p.act, p.slope);
          v.z = apply_act(v.z, p.act, p.slope); v.w = apply_act(v.w, p.act, p.slope);
          if (p.res1) {
            const float4 r = __ldg(reinterpret_cast<const float4*>(p.res1 + l.pix * p.res1_ld + l.c));
            v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
          }
          if (p.res2) {
            const float4 r = __ldg(reinterpret_cast<const float4*>(p.res2 + l.pix * p.res2_ld + l.c));
            v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
          }
          *reinterpret_cast<float4*>(p.out + l.pix * p.out_ld + l.c) = v;
          amax = fmaxf(fmaxf(amax, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
        }
        continue;
      }
    }
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
      const int co = cobase + j;
      if (co >= p.cout) continue;
      const OutLoc l = out_loc(p, n, oy, ox, co);
      float v = acc[i][j] + (p.bias ? __ldg(p.bias + co) : 0.f);
      if (p.out_planar) {
        v = apply_act(v, p.act, p.slope);
        p.out[(((int64_t)n * p.cout + co) * p.Ho + oy) * p.Wo + ox] = v;
      } else {
        v = epilogue1(p, v, l);
        p.out[l.pix * p.out_ld + l.c] = v;
      }
      amax = fmaxf(amax, fabsf(v));
    }
  }
  if (p.out_absmax != nullptr && amax > 0.f) atomicMax(reinterpret_cast<int*>(p.out_absmax), __float_as_int(amax));
}

template <int TN, int K, int S>
static int launch_simt(const TdvcConvParams& p, cudaStream_t st) {
  constexpr int IH = (TY - 1) * S + K, IW = (TX - 1) * S + K;
  constexpr int IHWP = (IH * IW + 3) & ~3;
  constexpr size_t smem = (size_t)(CK * IHWP + K * CK * TN) * sizeof(float);
  static_assert(smem <= 48 * 1024, "static smem budget");
  const int tiles_x = cdiv(p.Wo, TX), tiles_y = cdiv(p.Ho, TY);
  dim3 grid(tiles_x * tiles_y, p.cout_pad / TN, p.N);
  int vec_ok = (p.out_ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0);
  if (p.out_planar) vec_ok = 0;
  if (p.shuffle == 2) vec_ok = vec_ok && ((p.cout >> 2) % 4 == 0);
  else vec_ok = vec_ok && (p.cout % 4 == 0);
  if (p.post != TDVC_POST_NONE) vec_ok = vec_ok && (p.mul_ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.mul) & 15) == 0);
  if (p.res1) vec_ok = vec_ok && (p.res1_ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.res1) & 15) == 0);
  if (p.res2) vec_ok = vec_ok && (p.res2_ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.res2) & 15) == 0);
  conv_simt_kernel<TN, K, S><<<grid, CONV_THREADS, smem, st>>>(p, tiles_x, vec_ok);
  TDVC_CHECK_LAUNCH("conv_simt");
  return TDVC_OK;
}

template <int K, int S>
static int launch_simt_tn(const TdvcConvParams& p, cudaStream_t st) {
  if (p.cout_pad % 64 == 0) return launch_simt<64, K, S>(p, st);
  if (p.cout_pad % 32 == 0) return launch_simt<32, K, S>(p, st);
  return launch_simt<16, K, S>(p, st);
}

int conv2d_validate(const TdvcConvParams* p) {
  TDVC_REQUIRE(p != nullptr, "conv2d: null params");
  TDVC_REQUIRE(p->n_src >= 1 && p->n_src <= 4, "conv2d: n_src %d", p->n_src);
  int cin = 0;
  for (int s = 0; s < p->n_src; ++s) {
    TDVC_REQUIRE(p->src[s] != nullptr, "conv2d: src[%d] null", s);
    TDVC_REQUIRE(p->src_c[s] > 0 && p->src_c[s] % 4 == 0 && p->src_ld[s] % 4 == 0 && p->src_ld[s] >= p->src_c[s],
                 "conv2d: src[%d] c=%d ld=%d must be multiples of 4", s, p->src_c[s], p->src_ld[s]);
    TDVC_REQUIRE((reinterpret_cast<uintptr_t>(p->src[s]) & 15) == 0, "conv2d: src[%d] not 16-byte aligned", s);
    cin += p->src_c[s];
  }
  TDVC_REQUIRE(cin == p->cin, "conv2d: cin %d != sum of sources %d", p->cin, cin);
  TDVC_REQUIRE(p->cin_pad % CK == 0 && p->cin_pad >= p->cin, "conv2d: cin_pad %d", p->cin_pad);
  TDVC_REQUIRE(p->cout_pad % 16 == 0 && p->cout_pad >= p->cout && p->cout > 0, "conv2d: cout %d cout_pad %d", p->cout, p->cout_pad);
  TDVC_REQUIRE(p->kh == p->kw, "conv2d: only square kernels");
  TDVC_REQUIRE(p->weight != nullptr && p->out != nullptr, "conv2d: null weight/out");
  TDVC_REQUIRE((reinterpret_cast<uintptr_t>(p->weight) & 15) == 0, "conv2d: weight not 16-byte aligned");
  TDVC_REQUIRE(p->bias == nullptr || (reinterpret_cast<uintptr_t>(p->bias) & 15) == 0, "conv2d: bias not aligned");
  TDVC_REQUIRE(p->N > 0 && p->H > 0 && p->W > 0, "conv2d: empty input");
  TDVC_REQUIRE(p->Ho == (p->H + 2 * p->pad - p->kh) / p->stride + 1 && p->Wo == (p->W + 2 * p->pad - p->kw) / p->stride + 1,
               "conv2d: Ho/Wo (%d,%d) inconsistent", p->Ho, p->Wo);
  TDVC_REQUIRE(p->shuffle == 0 || (p->shuffle == 2 && p->cout % 4 == 0), "conv2d: shuffle %d", p->shuffle);
  TDVC_REQUIRE(p->post == TDVC_POST_NONE || p->mul != nullptr, "conv2d: post needs mul");
  TDVC_REQUIRE(p->out_planar == 0 || (p->shuffle == 0 && p->post == TDVC_POST_NONE && !p->res1 && !p->res2),
               "conv2d: out_planar excludes shuffle / post / residual");
  TDVC_REQUIRE(p->N <= 65535 && p->cout_pad / 16 <= 65535, "conv2d: grid too large");
  TDVC_REQUIRE(p->products == 0 || p->products == 1, "conv2d: products %d (0 = fp32-class split scheme, 1 = one fp16 product)", p->products);
  TDVC_REQUIRE((reinterpret_cast<uintptr_t>(p->out_absmax) & 3) == 0 && (reinterpret_cast<uintptr_t>(p->in_absmax) & 3) == 0,
               "conv2d: out_absmax / in_absmax not 4-byte aligned");
  return TDVC_OK;
}

int conv2d_simt(const TdvcConvParams& p, cudaStream_t st) {
  const int k = p.kh, s = p.stride;
  if (k == 3 && s == 1) return launch_simt_tn<3, 1>(p, st);
  if (k == 3 && s == 2) return launch_simt_tn<3, 2>(p, st);
  if (k == 1 && s == 1) return launch_simt_tn<1, 1>(p, st);
  if (k == 1 && s == 2) return launch_simt_tn<1, 2>(p, st);
  if (k == 5 && s == 1) return launch_simt_tn<5, 1>(p, st);
  if (k == 7 && s == 1) return launch_simt_tn<7, 1>(p, st);
  set_error("conv2d: unsupported kernel %d stride %d", k, s);
  return TDVC_EINVAL;
}

}  // namespace tdvc
