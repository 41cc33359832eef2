// MS-SSIM on the device (reference main/model/ms_ssim_torch.py:21-87 `_ssim`, :138-200 `ms_ssim`; used by the evaluation
// loop tools/predict.py:93-96).  One level = separable 11-tap Gaussian filter (horizontal pass first, then vertical, "valid"
// padding) of x, y, x*x, y*y, x*y, the SSIM / contrast-structure maps and their sums - fused in one kernel: a CTA stages
// a (32+10) x (32+10) tile of both images in shared memory, filters it twice and adds the two map sums of its 32x32
// outputs to fp64 accumulators.  Nothing but the two sums per image leaves the kernel (memory-bound: 8 B/px read).
#include "common.cuh"

namespace tdvc {

constexpr int MS_T = 32;          // output tile edge
constexpr int MS_WIN = 11;        // the kernel is specialised for the reference's win_size = 11
constexpr int MS_IN = MS_T + MS_WIN - 1;

struct MsWin {
  float g[MS_WIN];
};

__global__ void __launch_bounds__(256) ssim_level_kernel(const float* __restrict__ x, const float* __restrict__ y, int H, int W,
                                                         int planes_per_image, MsWin win, float C1, float C2,
                                                         double* __restrict__ acc /* [N][2] : ssim sum, cs sum */) {
  __shared__ float sx[MS_IN][MS_IN + 1], sy[MS_IN][MS_IN + 1];
  __shared__ float hz[5][MS_IN][MS_T + 1];   // horizontally filtered x, y, xx, yy, xy
  __shared__ double red[2][8];
  const int plane = blockIdx.z;
  const int Ho = H - (MS_WIN - 1), Wo = W - (MS_WIN - 1);
  const int oy0 = blockIdx.y * MS_T, ox0 = blockIdx.x * MS_T;
  const float* xp = x + (int64_t)plane * H * W;
  const float* yp = y + (int64_t)plane * H * W;
  for (int i = threadIdx.x; i < MS_IN * MS_IN; i += blockDim.x) {
    const int r = i / MS_IN, c = i - r * MS_IN;
    const int gy = oy0 + r, gx = ox0 + c;
    const bool ok = gy < H && gx < W;
    sx[r][c] = ok ? __ldg(xp + (int64_t)gy * W + gx) : 0.f;
    sy[r][c] = ok ? __ldg(yp + (int64_t)gy * W + gx) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < MS_IN * MS_T; i += blockDim.x) {
    const int r = i / MS_T, c = i - r * MS_T;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f;
#pragma unroll
    for (int k = 0; k < MS_WIN; ++k) {
      const float xv = sx[r][c + k], yv = sy[r][c + k], g = win.g[k];
      a0 = fmaf(g, xv, a0);
      a1 = fmaf(g, yv, a1);
      a2 = fmaf(g, xv * xv, a2);
      a3 = fmaf(g, yv * yv, a3);
      a4 = fmaf(g, xv * yv, a4);
    }
    hz[0][r][c] = a0; hz[1][r][c] = a1; hz[2][r][c] = a2; hz[3][r][c] = a3; hz[4][r][c] = a4;
  }
  __syncthreads();
  double s_ssim = 0.0, s_cs = 0.0;
  for (int i = threadIdx.x; i < MS_T * MS_T; i += blockDim.x) {
    const int r = i / MS_T, c = i - r * MS_T;
    if (oy0 + r >= Ho || ox0 + c >= Wo) continue;
    float m1 = 0.f, m2 = 0.f, xx = 0.f, yy = 0.f, xy = 0.f;
#pragma unroll
    for (int k = 0; k < MS_WIN; ++k) {
      const float g = win.g[k];
      m1 = fmaf(g, hz[0][r + k][c], m1);
      m2 = fmaf(g, hz[1][r + k][c], m2);
      xx = fmaf(g, hz[2][r + k][c], xx);
      yy = fmaf(g, hz[3][r + k][c], yy);
      xy = fmaf(g, hz[4][r + k][c], xy);
    }
    const float m11 = m1 * m1, m22 = m2 * m2, m12 = m1 * m2;
    const float s1 = xx - m11, s2 = yy - m22, s12 = xy - m12;
    const float cs = (2.f * s12 + C2) / (s1 + s2 + C2);
    const float ss = ((2.f * m12 + C1) / (m11 + m22 + C1)) * cs;
    s_ssim += (double)ss;
    s_cs += (double)cs;
  }
  s_ssim = warp_sum_d(s_ssim);
  s_cs = warp_sum_d(s_cs);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[0][warp] = s_ssim; red[1][warp] = s_cs; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a += red[0][w]; b += red[1][w]; }
    const int n = plane / planes_per_image;
    atomicAdd(acc + 2 * n, a);
    atomicAdd(acc + 2 * n + 1, b);
  }
}

// F.avg_pool2d(x, kernel_size=2, padding=(H%2, W%2)) with the default count_include_pad=True (ms_ssim_torch.py:188-190)
__global__ void avgpool2_pad_kernel(const float* __restrict__ src, float* __restrict__ dst, int H, int W, int Ho, int Wo, int py,
                                    int px, int64_t total) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ox = (int)(i % Wo);
    int64_t r = i / Wo;
    const int oy = (int)(r % Ho);
    const int64_t plane = r / Ho;
    const float* s = src + plane * H * W;
    float a = 0.f;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const int iy = 2 * oy - py + dy, ix = 2 * ox - px + dx;
        if (iy >= 0 && iy < H && ix >= 0 && ix < W) a += __ldg(s + (int64_t)iy * W + ix);
      }
    dst[i] = a * 0.25f;
  }
}

}  // namespace tdvc

using namespace tdvc;

extern "C" int tdvc_ssim_level(const float* x, const float* y, int N, int C, int H, int W, const float* win11_host, float C1,
                               float C2, double* acc, void* stream) {
  TDVC_REQUIRE(x && y && acc && win11_host && N > 0 && C > 0, "ssim_level: bad args");
  TDVC_REQUIRE(H >= MS_WIN && W >= MS_WIN, "ssim_level: image %dx%d smaller than the 11-tap window", H, W);
  MsWin win;
  for (int k = 0; k < MS_WIN; ++k) win.g[k] = win11_host[k];
  const int Ho = H - (MS_WIN - 1), Wo = W - (MS_WIN - 1);
  dim3 grid(cdiv(Wo, MS_T), cdiv(Ho, MS_T), N * C);
  TDVC_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "ssim_level: grid too large");
  ssim_level_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, y, H, W, C, win, C1, C2, acc);
  TDVC_CHECK_LAUNCH("ssim_level");
  return TDVC_OK;
}

extern "C" int tdvc_avgpool2_pad(const float* src, float* dst, int planes, int H, int W, void* stream) {
  TDVC_REQUIRE(src && dst && planes > 0 && H > 0 && W > 0, "avgpool2_pad: bad args");
  const int py = H & 1, px = W & 1;
  const int Ho = (H + 2 * py - 2) / 2 + 1, Wo = (W + 2 * px - 2) / 2 + 1;
  const int64_t total = (int64_t)planes * Ho * Wo;
  int g = cdiv(total, 256);
  if (g > kNumSMs * 16) g = kNumSMs * 16;
  avgpool2_pad_kernel<<<g, 256, 0, (cudaStream_t)stream>>>(src, dst, H, W, Ho, Wo, py, px, total);
  TDVC_CHECK_LAUNCH("avgpool2_pad");
  return TDVC_OK;
}
