// Reference-based in-loop filter: patch match + block gather (reference main/model/pnet.py:213-257).
// The reference materialises F.unfold of the full-resolution reference features (1.06 GB at 1920x1024)
// and a gathered copy of the same size; here the match runs on pooled descriptors and the "gather" is a
// single streaming pass that reads each source block where it lies (512 B/pixel algorithmic traffic).
#include "common.cuh"

namespace tdvc {

// nn.AvgPool2d(scale): one CTA of 1024 threads per pooled cell; C/4 float4 lanes x (1024/(C/4)) pixel lanes, four
// independent partial sums per thread (loads in flight), fixed-order tree over the pixel lanes: deterministic.
__global__ void __launch_bounds__(1024) avgpool_scale_kernel(const float* __restrict__ x, int ld, float* __restrict__ out, int H,
                                                             int W, int C, int scale, int ph, int pw) {
  extern __shared__ float4 sh4[];
  const int lanes_c = C >> 2;
  const int lanes_p = blockDim.x / lanes_c;
  const int c4 = threadIdx.x % lanes_c, pl = threadIdx.x / lanes_c;
  const int cell = blockIdx.x, n = blockIdx.y;
  const int cy = cell / pw, cx = cell - cy * pw;
  const float* base = x + (((int64_t)n * H + (int64_t)cy * scale) * W + (int64_t)cx * scale) * ld + c4 * 4;
  float4 s[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) s[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (pl < lanes_p) {
    const int total = scale * scale;
    for (int q0 = pl; q0 < total; q0 += 4 * lanes_p) {
      float4 v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int q = q0 + j * lanes_p;
        v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (q < total) {
          const int qy = q / scale, qx = q - qy * scale;
          v[j] = __ldg(reinterpret_cast<const float4*>(base + ((int64_t)qy * W + qx) * ld));
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) { s[j].x += v[j].x; s[j].y += v[j].y; s[j].z += v[j].z; s[j].w += v[j].w; }
    }
  }
  float4 t = make_float4((s[0].x + s[1].x) + (s[2].x + s[3].x), (s[0].y + s[1].y) + (s[2].y + s[3].y),
                         (s[0].z + s[1].z) + (s[2].z + s[3].z), (s[0].w + s[1].w) + (s[2].w + s[3].w));
  sh4[threadIdx.x] = t;
  __syncthreads();
  for (int stride = 1; stride < lanes_p; stride <<= 1) {  // pairwise tree over the pixel lanes (lanes_p is a power of two)
    if (pl < lanes_p && (pl % (2 * stride)) == 0 && pl + stride < lanes_p) {
      float4 a = sh4[threadIdx.x];
      const float4 b = sh4[threadIdx.x + stride * lanes_c];
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
      sh4[threadIdx.x] = a;
    }
    __syncthreads();
  }
  if (threadIdx.x < lanes_c) {
    float4 r = sh4[threadIdx.x];
    const float d = (float)(scale * scale);
    r.x /= d; r.y /= d; r.z /= d; r.w /= d;
    reinterpret_cast<float4*>(out + (((int64_t)n * ph + cy) * pw + cx) * C)[threadIdx.x] = r;
  }
}

// F.unfold(k=3, pad=3, stride=3) + F.normalize(eps=1e-12).  One CTA per patch.
__global__ void ff_desc_kernel(const float* __restrict__ pooled, float* __restrict__ desc, int ph, int pw, int C, int PW, int P) {
  const int patch = blockIdx.x, n = blockIdx.y;
  const int pi = patch / PW, pj = patch - pi * PW;
  const int D = C * 9;
  float* d = desc + ((int64_t)n * P + patch) * D;
  float ss = 0.f;
  for (int f = threadIdx.x; f < D; f += blockDim.x) {
    const int c = f / 9, k = f - c * 9;
    const int y = 3 * pi - 3 + k / 3, x = 3 * pj - 3 + k % 3;
    float v = 0.f;
    if (y >= 0 && y < ph && x >= 0 && x < pw) v = pooled[(((int64_t)n * ph + y) * pw + x) * C + c];
    d[f] = v;
    ss += v * v;
  }
  ss = warp_sum(ss);
  __shared__ float sh[32];
  __shared__ float inv;
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = ss;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];
    inv = fmaxf(sqrtf(t), 1e-12f);
  }
  __syncthreads();
  const float nrm = inv;
  for (int f = threadIdx.x; f < D; f += blockDim.x) d[f] = d[f] / nrm;
}

// ind[n][q] = first argmax_r <desc_q[q], desc_r[r]>.  One warp per query patch.
__global__ void ff_match_kernel(const float* __restrict__ dq, const float* __restrict__ dr, int32_t* __restrict__ ind,
                                float* __restrict__ sim, int P, int D) {
  const int warps = blockDim.x >> 5;
  const int q = blockIdx.x * warps + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int n = blockIdx.y;
  if (q >= P) return;
  const float* a = dq + ((int64_t)n * P + q) * D;
  float best = -INFINITY;
  int besti = 0;
  for (int r = 0; r < P; ++r) {
    const float* b = dr + ((int64_t)n * P + r) * D;
    float s = 0.f;
    for (int f = lane; f < D; f += 32) s = fmaf(a[f], b[f], s);
    s = warp_sum(s);
    if (sim != nullptr && lane == 0) sim[((int64_t)n * P + q) * P + r] = s;
    if (s > best) { best = s; besti = r; }
  }
  if (lane == 0) ind[(int64_t)n * P + q] = besti;
}

// Block placement + cosine gate.  16 threads (float4 each) per pixel, C == 64.
__global__ void ff_gather_kernel(const float* __restrict__ f_in, const float* __restrict__ f_ref, const int32_t* __restrict__ ind,
                                 float* __restrict__ out_a, float* __restrict__ out_b, float* __restrict__ gathered,
                                 float* __restrict__ cor_out, int N, int H, int W, int bs, int PW, int P) {
  const int64_t total = (int64_t)N * H * W * 16;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int l = (int)(i & 15);
    const int64_t pix = i >> 4;
    const int x = (int)(pix % W);
    const int64_t r0 = pix / W;
    const int y = (int)(r0 % H);
    const int n = (int)(r0 / H);
    const int by = y / bs + 1, bx = x / bs + 1;
    const int r = __ldg(ind + (int64_t)n * P + by * PW + bx);
    const int ry = r / PW, rx = r - ry * PW;
    const int sy = (ry - 1) * bs + (y % bs), sx = (rx - 1) * bs + (x % bs);
    const float4 a = __ldg(reinterpret_cast<const float4*>(f_in) + pix * 16 + l);
    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
    if (sy >= 0 && sy < H && sx >= 0 && sx < W)
      b = __ldg(reinterpret_cast<const float4*>(f_ref) + (((int64_t)n * H + sy) * W + sx) * 16 + l);
    float sa = a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
    float sb = b.x * b.x + b.y * b.y + b.z * b.z + b.w * b.w;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      sa += __shfl_xor_sync(0xffffffffu, sa, o);
      sb += __shfl_xor_sync(0xffffffffu, sb, o);
    }
    const float na = fmaxf(sqrtf(sa), 1e-8f), nb = fmaxf(sqrtf(sb), 1e-8f);
    float dot = (a.x / na) * (b.x / nb) + (a.y / na) * (b.y / nb) + (a.z / na) * (b.z / nb) + (a.w / na) * (b.w / nb);
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    reinterpret_cast<float4*>(out_a)[pix * 16 + l] = make_float4(a.x * dot, a.y * dot, a.z * dot, a.w * dot);
    reinterpret_cast<float4*>(out_b)[pix * 16 + l] = make_float4(b.x * dot, b.y * dot, b.z * dot, b.w * dot);
    if (gathered != nullptr) reinterpret_cast<float4*>(gathered)[pix * 16 + l] = b;
    if (cor_out != nullptr && l == 0) cor_out[pix] = dot;
  }
}

}  // namespace tdvc

using namespace tdvc;

extern "C" int tdvc_avgpool_scale(const float* x, int ld, float* out, int N, int H, int W, int C, int scale, void* stream) {
  TDVC_REQUIRE(x && out && N > 0 && scale > 0 && H >= scale && W >= scale, "avgpool_scale: bad args");
  TDVC_REQUIRE(C % 4 == 0 && C <= 1024 && ld % 4 == 0 && ((C >> 2) & ((C >> 2) - 1)) == 0, "avgpool_scale: C/4 must be a power of two <= 256, ld %% 4 == 0");
  const int ph = H / scale, pw = W / scale;
  dim3 grid(ph * pw, N);
  avgpool_scale_kernel<<<grid, 1024, 1024 * sizeof(float4), (cudaStream_t)stream>>>(x, ld, out, H, W, C, scale, ph, pw);
  TDVC_CHECK_LAUNCH("avgpool_scale");
  return TDVC_OK;
}

extern "C" int tdvc_ff_descriptors(const float* pooled, float* desc, int N, int ph, int pw, int C, void* stream) {
  TDVC_REQUIRE(pooled && desc && N > 0 && ph > 0 && pw > 0 && C > 0, "ff_descriptors: bad args");
  const int PH = (ph + 3) / 3 + 1, PW = (pw + 3) / 3 + 1;
  dim3 grid(PH * PW, N);
  ff_desc_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(pooled, desc, ph, pw, C, PW, PH * PW);
  TDVC_CHECK_LAUNCH("ff_descriptors");
  return TDVC_OK;
}

extern "C" int tdvc_ff_match(const float* desc_q, const float* desc_r, int32_t* ind, float* sim_or_null, int N, int P, int D, void* stream) {
  TDVC_REQUIRE(desc_q && desc_r && ind && N > 0 && P > 0 && D > 0, "ff_match: bad args");
  dim3 grid(cdiv(P, 4), N);
  ff_match_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(desc_q, desc_r, ind, sim_or_null, P, D);
  TDVC_CHECK_LAUNCH("ff_match");
  return TDVC_OK;
}

extern "C" int tdvc_ff_gather(const float* f_in, const float* f_ref, const int32_t* ind, float* out_a, float* out_b,
                              float* gathered_or_null, float* cor_or_null, int N, int H, int W, int C, int scale, void* stream) {
  TDVC_REQUIRE(f_in && f_ref && ind && out_a && out_b && N > 0 && H > 0 && W > 0 && scale > 0, "ff_gather: bad args");
  TDVC_REQUIRE(C == 64, "ff_gather: C must be 64 (got %d)", C);
  const int bs = 3 * scale;
  const int PH = (H + bs) / bs + 1, PW = (W + bs) / bs + 1;
  const int64_t total = (int64_t)N * H * W * 16;
  int64_t b = (total + 255) / 256;
  if (b > kNumSMs * 16) b = kNumSMs * 16;
  ff_gather_kernel<<<(int)b, 256, 0, (cudaStream_t)stream>>>(f_in, f_ref, ind, out_a, out_b, gathered_or_null, cor_or_null,
                                                         N, H, W, bs, PW, PH * PW);
  TDVC_CHECK_LAUNCH("ff_gather");
  return TDVC_OK;
}
