// KxK stride-1 convolutions with at most 4 output channels, exact fp32 on the CUDA cores.
//
// The SPyNet flow head (16 -> 2, 7x7; reference main/model/flownet.py:218-227) and the image heads (64 -> 3, 3x3;
// reference main/model/pnet.py:259-262) have 2-3 output channels: on the tensor-core kernel (conv_tc.cu) they occupy
// 4-6 of the 128 accumulator rows of every MMA and pay for all of them (16 -> 2 at 1024x1920: 1.22 ms).  As plain FFMA
// work they are 1.6-1.7 kMAC per pixel, a quarter of a millisecond.
//
// One CTA (128 threads) computes a 32 x 32 output tile: lane = tile column, each warp owns 8 consecutive tile rows,
// a thread accumulates its 8 rows x CO channels.  Input channels are consumed in chunks of 8; the chunk's
// (32+K-1)^2 halo sits in shared memory as [channel/4][halo pixel] float4 planes, so that the 32 lanes of a warp read 32
// consecutive float4 (conflict-free); the next chunk's halo is in flight (cp.async, double buffer) while one is consumed.  For one (channel quad, kx) the K x CO weight float4 stay in registers while the
// thread walks the 8+K-1 halo rows of its column: one LDS.128 feeds up to K x CO x 4 FFMA.
#include "common.cuh"

namespace tdvc {
namespace small {

constexpr int TS = 32;        // tile side (pixels)
constexpr int RPT = 8;        // rows per thread
constexpr int CC = 8;         // input channels per chunk (two float4 quads)
constexpr int THREADS = 128;

template <int K, int CO>
struct Cfg {
  static constexpr int IW = TS + K - 1;
  static constexpr int NPX = IW * IW;
  static constexpr int HALO_F4 = (CC / 4) * NPX;   // one chunk buffer; two of them are in flight
  static size_t smem(int cin) { return (size_t)(2 * HALO_F4 + K * K * (cin / 4) * CO) * sizeof(float4); }
};

// 16-byte async copy global -> shared; bytes = 0 zero-fills (halo pixels outside the image)
__device__ __forceinline__ void cp_async16(void* dst, const void* src, int bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src),
               "r"(bytes)
               : "memory");
}

template <int K, int CO>
__global__ void __launch_bounds__(THREADS) conv_small_kernel(const TdvcConvParams p, int tiles_x, int tiles_y) {
  using C = Cfg<K, CO>;
  extern __shared__ __align__(16) float4 sm4[];
  float4* halo = sm4;                     // [2 buffers][CC/4][NPX]
  float4* wsm = sm4 + 2 * C::HALO_F4;     // [K*K][cin/4][CO]: the 4 input channels of a quad for one output channel
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int b = blockIdx.x;
  const int tx = b % tiles_x; b /= tiles_x;
  const int ty = b % tiles_y;
  const int n = b / tiles_y;
  const int oy0 = ty * TS, ox0 = tx * TS;
  const int iy0 = oy0 - K / 2, ix0 = ox0 - K / 2;
  const float* src = p.src[0] + (int64_t)n * p.H * p.W * p.src_ld[0];
  const int sld = p.src_ld[0];
  const int nquads = p.cin / 4;
  const int nchunks = p.cin / CC;

  // halo of chunk `ch` -> buffer ch & 1: two lanes cover the 32 bytes of a pixel's 8 channels; every copy is in flight
  // at once (cp.async), zero-filled outside the image
  auto fill = [&](int ch) {
    float4* hb = halo + (ch & 1) * C::HALO_F4;
    for (int idx = tid; idx < C::NPX * 2; idx += THREADS) {
      const int px = idx >> 1, q = idx & 1;
      const int hy = px / C::IW, hx = px - hy * C::IW;
      const int iy = iy0 + hy, ix = ix0 + hx;
      const bool ok = iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
      const float* g = ok ? src + ((int64_t)iy * p.W + ix) * sld + ch * CC + q * 4 : src;
      cp_async16(hb + q * C::NPX + px, g, ok ? 16 : 0);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  fill(0);
  // all weights of the layer from the packed [tap][cin_pad][cout_pad] fp32 array
  for (int idx = tid; idx < K * K * nquads * CO; idx += THREADS) {
    const int co = idx % CO;
    const int q = (idx / CO) % nquads;
    const int tap = idx / (CO * nquads);
    float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
    if (co < p.cout) {
      const float* wp = p.weight + ((int64_t)tap * p.cin_pad + q * 4) * p.cout_pad + co;
      w = make_float4(__ldg(wp), __ldg(wp + p.cout_pad), __ldg(wp + 2 * p.cout_pad), __ldg(wp + 3 * p.cout_pad));
    }
    wsm[idx] = w;
  }

  float acc[RPT][CO];
#pragma unroll
  for (int i = 0; i < RPT; ++i)
#pragma unroll
    for (int j = 0; j < CO; ++j) acc[i][j] = 0.f;

  for (int ch = 0; ch < nchunks; ++ch) {
    if (ch + 1 < nchunks) {
      fill(ch + 1);   // its buffer was released by the barrier at the end of iteration ch - 1
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const float4* hb = halo + (ch & 1) * C::HALO_F4;
#pragma unroll 1
    for (int q = 0; q < CC / 4; ++q) {
#pragma unroll 1
      for (int kx = 0; kx < K; ++kx) {
        float4 w[K][CO];
#pragma unroll
        for (int ky = 0; ky < K; ++ky)
#pragma unroll
          for (int j = 0; j < CO; ++j) w[ky][j] = wsm[((ky * K + kx) * nquads + ch * (CC / 4) + q) * CO + j];
        const float4* col = hb + q * C::NPX + (warp * RPT) * C::IW + lane + kx;
#pragma unroll
        for (int r = 0; r < RPT + K - 1; ++r) {
          const float4 v = col[r * C::IW];
#pragma unroll
          for (int ky = 0; ky < K; ++ky) {
            const int i = r - ky;
            if (i >= 0 && i < RPT) {
#pragma unroll
              for (int j = 0; j < CO; ++j) {
                float a = acc[i][j];
                a = fmaf(v.x, w[ky][j].x, a);
                a = fmaf(v.y, w[ky][j].y, a);
                a = fmaf(v.z, w[ky][j].z, a);
                a = fmaf(v.w, w[ky][j].w, a);
                acc[i][j] = a;
              }
            }
          }
        }
      }
    }
    __syncthreads();   // the chunk's buffer may be refilled
  }

  // ---- epilogue: bias, activation, residuals, store (lane = x: consecutive pixels)
  const int ox = ox0 + lane;
  if (ox >= p.Wo) return;
#pragma unroll
  for (int i = 0; i < RPT; ++i) {
    const int oy = oy0 + warp * RPT + i;
    if (oy >= p.Ho) break;
    const int64_t pix = ((int64_t)n * p.Ho + oy) * p.Wo + ox;
#pragma unroll
    for (int j = 0; j < CO; ++j) {
      if (j >= p.cout) break;
      float v = acc[i][j] + (p.bias ? __ldg(p.bias + j) : 0.f);
      v = apply_act(v, p.act, p.slope);
      if (p.res1) v += __ldg(p.res1 + pix * p.res1_ld + j);
      if (p.res2) v += __ldg(p.res2 + pix * p.res2_ld + j);
      p.out[pix * p.out_ld + j] = v;
    }
  }
}

template <int K, int CO>
static int launch(const TdvcConvParams& p, cudaStream_t st) {
  using C = Cfg<K, CO>;
  const size_t smem = C::smem(p.cin);
  TDVC_REQUIRE(smem <= 200 * 1024, "conv_small: cin %d needs %zu bytes of shared memory", p.cin, smem);
  static int smem_done[kMaxDevices] = {0};
  if (int rc = ensure_dynamic_smem(conv_small_kernel<K, CO>, smem, smem_done, "conv_small")) return rc;
  const int tiles_x = cdiv(p.Wo, TS), tiles_y = cdiv(p.Ho, TS);
  const int64_t blocks = (int64_t)tiles_x * tiles_y * p.N;
  TDVC_REQUIRE(blocks < (1ll << 31), "conv_small: too many tiles");
  conv_small_kernel<K, CO><<<(unsigned)blocks, THREADS, smem, st>>>(p, tiles_x, tiles_y);
  TDVC_CHECK_LAUNCH("conv_small");
  return TDVC_OK;
}

}  // namespace small

int conv2d_small_supported(const TdvcConvParams& p) {
  return p.n_src == 1 && p.stride == 1 && p.kh == p.kw && (p.kh == 3 || p.kh == 7) && p.pad == p.kh / 2 && p.cout >= 1 &&
         p.cout <= 4 && p.cin % small::CC == 0 && p.src_c[0] == p.cin && p.shuffle == 0 && p.post == TDVC_POST_NONE &&
         p.out_planar == 0 && p.in_square == 0;
}

int conv2d_small(const TdvcConvParams& p, cudaStream_t st) {
  if (!conv2d_small_supported(p)) {
    set_error("conv_small: unsupported shape");
    return TDVC_EINVAL;
  }
  if (p.kh == 7) {
    if (p.cout <= 2) return small::launch<7, 2>(p, st);
    if (p.cout == 3) return small::launch<7, 3>(p, st);
    return small::launch<7, 4>(p, st);
  }
  if (p.cout <= 2) return small::launch<3, 2>(p, st);
  if (p.cout == 3) return small::launch<3, 3>(p, st);
  return small::launch<3, 4>(p, st);
}

}  // namespace tdvc
