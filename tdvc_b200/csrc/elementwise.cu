// Memory-bound glue kernels: layout changes, fused element-wise steps, squeeze-excitation, metrics.
// All are grid-stride, float4-vectorised where alignment allows, sized as multiples of the SM count.
#include "common.cuh"

namespace tdvc {

// ------------------------------------------------------------------ layout
// src [N][C][HW]  ->  dst [N][HW][ld]   (32x32 smem tile transpose, +1 padding against bank conflicts)
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, float* __restrict__ dst, int C, int64_t HW, int ld) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int64_t p0 = (int64_t)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const float* s = src + (int64_t)n * C * HW;
  float* d = dst + (int64_t)n * HW * ld;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i;
    const int64_t p = p0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && p < HW) ? s[(int64_t)c * HW + p] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int64_t p = p0 + i;
    const int c = c0 + threadIdx.x;
    if (p < HW && c < ld) d[p * ld + c] = tile[threadIdx.x][i];  // c in [C, ld) gets the zero fill
  }
}

// images (C <= 4, ld == 4): one pixel per thread, coalesced plane reads, one 16-byte store
__global__ void nchw_to_nhwc4_kernel(const float* __restrict__ src, float* __restrict__ dst, int C, int64_t HW, int64_t total) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = i / HW, px = i - n * HW;
    const float* s = src + n * C * HW + px;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    v.x = __ldg(s);
    if (C > 1) v.y = __ldg(s + HW);
    if (C > 2) v.z = __ldg(s + 2 * HW);
    if (C > 3) v.w = __ldg(s + 3 * HW);
    reinterpret_cast<float4*>(dst)[i] = v;
  }
}

// images (C <= 4, ld == 4): one pixel per thread, one 16-byte load, coalesced plane stores
__global__ void nhwc4_to_nchw_kernel(const float* __restrict__ src, float* __restrict__ dst, int C, int64_t HW, int64_t total) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = i / HW, px = i - n * HW;
    const float4 v = __ldg(reinterpret_cast<const float4*>(src) + i);
    float* d = dst + n * C * HW + px;
    d[0] = v.x;
    if (C > 1) d[HW] = v.y;
    if (C > 2) d[2 * HW] = v.z;
    if (C > 3) d[3 * HW] = v.w;
  }
}

__global__ void nhwc_to_nchw_kernel(const float* __restrict__ src, int ld, float* __restrict__ dst, int C, int64_t HW) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int64_t p0 = (int64_t)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const float* s = src + (int64_t)n * HW * ld;
  float* d = dst + (int64_t)n * C * HW;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int64_t p = p0 + i;
    const int c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (p < HW && c < C) ? s[p * ld + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i;
    const int64_t p = p0 + threadIdx.x;
    if (c < C && p < HW) d[(int64_t)c * HW + p] = tile[threadIdx.x][i];
  }
}

// ------------------------------------------------------------------ element-wise
__global__ void axpby_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out,
                             int64_t n, float alpha, float beta, int vec) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (vec) {
    const int64_t n4 = n >> 2;
    for (int64_t i = t; i < n4; i += stride) {
      const float4 x = __ldg(reinterpret_cast<const float4*>(a) + i);
      const float4 y = __ldg(reinterpret_cast<const float4*>(b) + i);
      float4 r;
      r.x = alpha * x.x + beta * y.x; r.y = alpha * x.y + beta * y.y;
      r.z = alpha * x.z + beta * y.z; r.w = alpha * x.w + beta * y.w;
      reinterpret_cast<float4*>(out)[i] = r;
    }
    for (int64_t i = (n4 << 2) + t; i < n; i += stride) out[i] = alpha * a[i] + beta * b[i];
  } else {
    for (int64_t i = t; i < n; i += stride) out[i] = alpha * a[i] + beta * b[i];
  }
}

__global__ void bcast_add_lrelu_kernel(const float* __restrict__ x, const float* __restrict__ tmp, float* __restrict__ out,
                                       int T, int64_t n4, float slope) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 b = __ldg(reinterpret_cast<const float4*>(tmp) + i);
    for (int t = 0; t < T; ++t) {
      float4 v = __ldg(reinterpret_cast<const float4*>(x) + (int64_t)t * n4 + i);
      v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
      v.x = v.x > 0.f ? v.x : v.x * slope; v.y = v.y > 0.f ? v.y : v.y * slope;
      v.z = v.z > 0.f ? v.z : v.z * slope; v.w = v.w > 0.f ? v.w : v.w * slope;
      reinterpret_cast<float4*>(out)[(int64_t)t * n4 + i] = v;
    }
  }
}

// out[j] = lrelu(x[j] + tmp) for up to four independent (x, out) pairs: tmp is read once for all of them
struct Ptr4 {
  const float* x[4];
  float* out[4];
};
__global__ void bcast_add_lrelu_multi_kernel(Ptr4 q, const float* __restrict__ tmp, int T, int64_t n4, float slope) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 b = __ldg(reinterpret_cast<const float4*>(tmp) + i);
    float4 v[4];
#pragma unroll
    for (int t = 0; t < 4; ++t)
      if (t < T) v[t] = __ldg(reinterpret_cast<const float4*>(q.x[t]) + i);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      if (t < T) {
        float4 o = v[t];
        o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
        o.x = o.x > 0.f ? o.x : o.x * slope; o.y = o.y > 0.f ? o.y : o.y * slope;
        o.z = o.z > 0.f ? o.z : o.z * slope; o.w = o.w > 0.f ? o.w : o.w * slope;
        reinterpret_cast<float4*>(q.out[t])[i] = o;
      }
    }
  }
}

__global__ void round_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = rintf(x[i]);
}

__global__ void sq_err_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n, double* acc) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  double s = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float d = a[i] - b[i];
    s += (double)d * d;
  }
  s = warp_sum_d(s);
  __shared__ double sh[32];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0;
    s = warp_sum_d(s);
    if (threadIdx.x == 0) atomicAdd(acc, s);
  }
}

// ------------------------------------------------------------------ squeeze-excitation
// partial[b][n][c] = sum over pixel slab b of x[n][p][c].  256 threads: C/4 float4 lanes x (1024/C) pixel lanes.
__global__ void se_partial_kernel(const float* __restrict__ x, int ld, int64_t HW, int C, float* __restrict__ partial) {
  extern __shared__ float4 sh4[];
  const int lanes_c = C >> 2;
  const int lanes_p = blockDim.x / lanes_c;
  const int c4 = threadIdx.x % lanes_c;
  const int pl = threadIdx.x / lanes_c;
  const int n = blockIdx.y, b = blockIdx.x, nblk = gridDim.x;
  const int64_t chunk = (HW + nblk - 1) / nblk;
  const int64_t p0 = (int64_t)b * chunk;
  const int64_t p1 = p0 + chunk < HW ? p0 + chunk : HW;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (pl < lanes_p) {
    const float* base = x + (int64_t)n * HW * ld + c4 * 4;
    // four independent loads in flight per thread (one per iteration left the kernel latency-bound at ~3 TB/s);
    // the four partial sums are combined in a fixed order: deterministic
    float4 a[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) a[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t p = p0 + pl; p < p1; p += 4 * lanes_p) {
      float4 v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int64_t q = p + (int64_t)j * lanes_p;
        v[j] = q < p1 ? __ldg(reinterpret_cast<const float4*>(base + q * ld)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) { a[j].x += v[j].x; a[j].y += v[j].y; a[j].z += v[j].z; a[j].w += v[j].w; }
    }
    s = make_float4((a[0].x + a[1].x) + (a[2].x + a[3].x), (a[0].y + a[1].y) + (a[2].y + a[3].y),
                    (a[0].z + a[1].z) + (a[2].z + a[3].z), (a[0].w + a[1].w) + (a[2].w + a[3].w));
  }
  sh4[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x < lanes_c) {
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int q = 0; q < lanes_p; ++q) {  // fixed order: deterministic
      const float4 v = sh4[q * lanes_c + threadIdx.x];
      t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
    }
    reinterpret_cast<float4*>(partial + ((int64_t)b * gridDim.y + n) * C)[threadIdx.x] = t;
  }
}

// Squeeze-excite gate of image n (reference inflate.py:189-207): channel means from the partial sums, C -> C/16 ReLU ->
// C sigmoid.  One block per image; the C gate values are written back over the image's first row of partial sums
// (partial[n*C + c], read by se_apply_kernel), so the streaming kernel's blocks do not each repeat the reduction.
__global__ void se_scale_kernel(float* __restrict__ partial, int nblk, const float* __restrict__ w1,
                                const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2,
                                int N, int64_t HW, int C, int Cr) {
  extern __shared__ float sh[];
  float* mean = sh;            // [C]
  float* hid = sh + C;         // [Cr]
  float* red = sh + C + Cr;    // [blockDim.x]
  const int n = blockIdx.x;
  const int nsl = blockDim.x / C;               // slices of the partial-sum rows summed in parallel (blockDim.x % C == 0)
  const int c = threadIdx.x % C, sl = threadIdx.x / C;
  float s = 0.f;
  if (sl < nsl) {   // 8 independent loads in flight per thread: the reduction is latency-bound, not bandwidth-bound
    const float* pp = partial + (int64_t)n * C + c;
    const int64_t rs = (int64_t)N * C;
    int b = sl;
    for (; b + 7 * nsl < nblk; b += 8 * nsl) {
      float v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = pp[(int64_t)(b + k * nsl) * rs];
#pragma unroll
      for (int k = 0; k < 8; ++k) s += v[k];
    }
    for (; b < nblk; b += nsl) s += pp[(int64_t)b * rs];
  }
  red[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x < C) {
    float t = 0.f;
    for (int k = 0; k < nsl; ++k) t += red[k * C + threadIdx.x];
    mean[threadIdx.x] = t / (float)HW;
  }
  __syncthreads();
  for (int r = threadIdx.x; r < Cr; r += blockDim.x) {
    float t = b1[r];
    for (int k = 0; k < C; ++k) t = fmaf(w1[r * C + k], mean[k], t);
    hid[r] = fmaxf(t, 0.f);
  }
  __syncthreads();
  if (threadIdx.x < C) {
    float t = b2[threadIdx.x];
    for (int r = 0; r < Cr; ++r) t = fmaf(w2[threadIdx.x * Cr + r], hid[r], t);
    partial[(int64_t)n * C + threadIdx.x] = 1.f / (1.f + expf(-t));
  }
}

__global__ void se_apply_kernel(const float* __restrict__ x, int ld, const float* __restrict__ gate,
                                int N, int64_t HW, int C, int act, float slope,
                                const float* __restrict__ res, int res_ld, float* __restrict__ out, int out_ld,
                                const float* __restrict__ sub_from, int sub_ld, float* __restrict__ out2, int out2_ld) {
  extern __shared__ float sh[];
  float* sc = sh;              // [C]
  const int n = blockIdx.y;
  for (int c = threadIdx.x; c < C; c += blockDim.x) sc[c] = gate[(int64_t)n * C + c];
  __syncthreads();
  const int lanes_c = C >> 2;
  const int64_t total = HW * lanes_c;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const float* xb = x + (int64_t)n * HW * ld;
  float* ob = out + (int64_t)n * HW * out_ld;
  const float* rb = res ? res + (int64_t)n * HW * res_ld : nullptr;
  const float* sb2 = sub_from ? sub_from + (int64_t)n * HW * sub_ld : nullptr;   // second output: out2 = sub_from - out
  float* ob2 = out2 ? out2 + (int64_t)n * HW * out2_ld : nullptr;
  // four float4 per thread and iteration, all loads issued before the first use (memory-level parallelism)
  const bool pow2 = (lanes_c & (lanes_c - 1)) == 0;
  const int lshift = __ffs(lanes_c) - 1;
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += 4 * stride) {
    float4 v[4], r[4], sf[4];
    int64_t pp[4];
    int cc[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t i = i0 + j * stride;
      pp[j] = pow2 ? (i >> lshift) : i / lanes_c;   // C/4 is a power of two for every SE block of the model
      cc[j] = (int)(i - pp[j] * lanes_c) * 4;
      v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      r[j] = v[j];
      if (i < total) {
        v[j] = __ldg(reinterpret_cast<const float4*>(xb + pp[j] * ld + cc[j]));
        if (rb) r[j] = __ldg(reinterpret_cast<const float4*>(rb + pp[j] * res_ld + cc[j]));
        if (sb2) sf[j] = __ldg(reinterpret_cast<const float4*>(sb2 + pp[j] * sub_ld + cc[j]));
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (i0 + j * stride >= total) continue;
      const int c = cc[j];
      float4 o;
      o.x = apply_act(v[j].x * sc[c], act, slope) + r[j].x;
      o.y = apply_act(v[j].y * sc[c + 1], act, slope) + r[j].y;
      o.z = apply_act(v[j].z * sc[c + 2], act, slope) + r[j].z;
      o.w = apply_act(v[j].w * sc[c + 3], act, slope) + r[j].w;
      *reinterpret_cast<float4*>(ob + pp[j] * out_ld + c) = o;
      if (sb2) *reinterpret_cast<float4*>(ob2 + pp[j] * out2_ld + c) = make_float4(sf[j].x - o.x, sf[j].y - o.y, sf[j].z - o.z, sf[j].w - o.w);
    }
  }
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace tdvc

using namespace tdvc;

extern "C" int tdvc_nchw_to_nhwc(const float* src, float* dst, int N, int C, int H, int W, int dst_ld, void* stream) {
  TDVC_REQUIRE(src && dst && N > 0 && C > 0 && H > 0 && W > 0 && dst_ld >= C, "nchw_to_nhwc: bad args");
  const int64_t HW = (int64_t)H * W;
  if (C <= 4 && dst_ld == 4 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
    const int64_t total = (int64_t)N * HW;
    int g = cdiv(total, 256);
    if (g > kNumSMs * 16) g = kNumSMs * 16;
    nchw_to_nhwc4_kernel<<<g, 256, 0, (cudaStream_t)stream>>>(src, dst, C, HW, total);
    TDVC_CHECK_LAUNCH("nchw_to_nhwc");
    return TDVC_OK;
  }
  dim3 grid(cdiv(HW, 32), cdiv(dst_ld, 32), N), block(32, 8);
  nchw_to_nhwc_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(src, dst, C, HW, dst_ld);
  TDVC_CHECK_LAUNCH("nchw_to_nhwc");
  return TDVC_OK;
}

extern "C" int tdvc_nhwc_to_nchw(const float* src, int src_ld, float* dst, int N, int C, int H, int W, void* stream) {
  TDVC_REQUIRE(src && dst && N > 0 && C > 0 && H > 0 && W > 0 && src_ld >= C, "nhwc_to_nchw: bad args");
  const int64_t HW = (int64_t)H * W;
  if (C <= 4 && src_ld == 4 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
    const int64_t total = (int64_t)N * HW;
    int g = cdiv(total, 256);
    if (g > kNumSMs * 16) g = kNumSMs * 16;
    nhwc4_to_nchw_kernel<<<g, 256, 0, (cudaStream_t)stream>>>(src, dst, C, HW, total);
    TDVC_CHECK_LAUNCH("nhwc_to_nchw");
    return TDVC_OK;
  }
  dim3 grid(cdiv(HW, 32), cdiv(C, 32), N), block(32, 8);
  nhwc_to_nchw_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(src, src_ld, dst, C, HW);
  TDVC_CHECK_LAUNCH("nhwc_to_nchw");
  return TDVC_OK;
}

static int ew_grid(int64_t work_items) {
  int64_t b = (work_items + 255) / 256;
  const int64_t cap = (int64_t)kNumSMs * 8;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

extern "C" int tdvc_axpby(const float* a, const float* b, float* out, int64_t n, float alpha, float beta, void* stream) {
  TDVC_REQUIRE(a && b && out && n >= 0, "axpby: bad args");
  if (n == 0) return TDVC_OK;
  const int vec = aligned16(a) && aligned16(b) && aligned16(out);
  axpby_kernel<<<ew_grid(n / 4 + 1), 256, 0, (cudaStream_t)stream>>>(a, b, out, n, alpha, beta, vec);
  TDVC_CHECK_LAUNCH("axpby");
  return TDVC_OK;
}

extern "C" int tdvc_bcast_add_lrelu(const float* x, const float* tmp, float* out, int T, int64_t n_per_t, float slope, void* stream) {
  TDVC_REQUIRE(x && tmp && out && T > 0 && n_per_t > 0 && n_per_t % 4 == 0, "bcast_add_lrelu: bad args");
  TDVC_REQUIRE(aligned16(x) && aligned16(tmp) && aligned16(out), "bcast_add_lrelu: alignment");
  bcast_add_lrelu_kernel<<<ew_grid(n_per_t / 4), 256, 0, (cudaStream_t)stream>>>(x, tmp, out, T, n_per_t / 4, slope);
  TDVC_CHECK_LAUNCH("bcast_add_lrelu");
  return TDVC_OK;
}

extern "C" int tdvc_bcast_add_lrelu_multi(const float* const* x, const float* tmp, float* const* out, int T, int64_t n,
                                          float slope, void* stream) {
  TDVC_REQUIRE(x && tmp && out && T > 0 && T <= 4 && n > 0 && n % 4 == 0 && aligned16(tmp), "bcast_add_lrelu_multi: bad args");
  Ptr4 q;
  for (int t = 0; t < 4; ++t) {
    q.x[t] = t < T ? x[t] : nullptr;
    q.out[t] = t < T ? out[t] : nullptr;
    TDVC_REQUIRE(t >= T || (q.x[t] && q.out[t] && aligned16(q.x[t]) && aligned16(q.out[t])), "bcast_add_lrelu_multi: pointer %d", t);
  }
  bcast_add_lrelu_multi_kernel<<<ew_grid(n / 4), 256, 0, (cudaStream_t)stream>>>(q, tmp, T, n / 4, slope);
  TDVC_CHECK_LAUNCH("bcast_add_lrelu_multi");
  return TDVC_OK;
}

extern "C" int tdvc_round_half_even(const float* x, float* out, int64_t n, void* stream) {
  TDVC_REQUIRE(x && out && n >= 0, "round: bad args");
  if (n == 0) return TDVC_OK;
  round_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(x, out, n);
  TDVC_CHECK_LAUNCH("round");
  return TDVC_OK;
}

extern "C" int tdvc_sq_err_sum(const float* a, const float* b, int64_t n, double* acc, void* stream) {
  TDVC_REQUIRE(a && b && acc && n > 0, "sq_err_sum: bad args");
  sq_err_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(a, b, n, acc);
  TDVC_CHECK_LAUNCH("sq_err_sum");
  return TDVC_OK;
}

extern "C" int tdvc_se_partial_sums(const float* x, int ld, int N, int64_t HW, int C, float* partial, int nblk, void* stream) {
  TDVC_REQUIRE(x && partial && N > 0 && HW > 0 && nblk > 0 && nblk <= 1024, "se_partial_sums: bad args");
  TDVC_REQUIRE(C % 4 == 0 && C <= 1024 && ld % 4 == 0 && aligned16(x) && aligned16(partial), "se_partial_sums: C/ld/alignment");
  dim3 grid(nblk, N);
  se_partial_kernel<<<grid, 256, 256 * sizeof(float4), (cudaStream_t)stream>>>(x, ld, HW, C, partial);
  TDVC_CHECK_LAUNCH("se_partial_sums");
  return TDVC_OK;
}

extern "C" int tdvc_se_apply(const float* x, int ld, float* partial, int nblk, const float* w1, const float* b1,
                             const float* w2, const float* b2, int N, int64_t HW, int C, int Cr, int act, float slope,
                             const float* res, int res_ld, float* out, int out_ld, const float* sub_from, int sub_ld,
                             float* out2, int out2_ld, void* stream) {
  TDVC_REQUIRE(x && partial && w1 && b1 && w2 && b2 && out && N > 0 && HW > 0 && Cr > 0, "se_apply: bad args");
  TDVC_REQUIRE((sub_from == nullptr) == (out2 == nullptr), "se_apply: sub_from and out2 go together");
  TDVC_REQUIRE(sub_from == nullptr || (sub_ld % 4 == 0 && out2_ld % 4 == 0 && aligned16(sub_from) && aligned16(out2)),
               "se_apply: sub_from / out2 alignment");
  TDVC_REQUIRE(C % 4 == 0 && ld % 4 == 0 && out_ld % 4 == 0 && aligned16(x) && aligned16(out), "se_apply: C/ld/alignment");
  TDVC_REQUIRE(res == nullptr || (res_ld % 4 == 0 && aligned16(res)), "se_apply: res alignment");
  int gx = ew_grid(HW * (C / 4));
  if (gx > kNumSMs * 8) gx = kNumSMs * 8;
  dim3 grid(gx, N);
  const int sthreads = (1024 / C) * C;   // a multiple of C
  TDVC_REQUIRE(C <= 1024, "se_apply: C %d > 1024", C);
  se_scale_kernel<<<N, sthreads, (C + Cr + sthreads) * sizeof(float), (cudaStream_t)stream>>>(
      partial, nblk, w1, b1, w2, b2, N, HW, C, Cr);
  TDVC_CHECK_LAUNCH("se_scale");
  se_apply_kernel<<<grid, 256, C * sizeof(float), (cudaStream_t)stream>>>(
      x, ld, partial, N, HW, C, act, slope, res, res_ld, out, out_ld, sub_from, sub_ld, out2, out2_ld);
  TDVC_CHECK_LAUNCH("se_apply");
  return TDVC_OK;
}
