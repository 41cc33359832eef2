// Quantise + likelihood + bits in one pass per latent (compressai EntropyBottleneck / GaussianConditional
// in eval mode; the log/sum reductions of reference main/model/pnet.py:38-43,62-67).
// Memory-bound: y-path reads 12 B/element (y, scale, mean), z-path 4 B/element (+4 B z_hat write).
// Reduction: fp64 per thread -> warp shuffle -> one atomicAdd(double) per CTA.
#include "common.cuh"

namespace tdvc {

__device__ __forceinline__ void block_accumulate(double s, double* acc) {
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

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// per-channel cumulative logits: widths 1,3,3,3,3,1.  Parameters in smem, laid out [param][C].
struct EbSmem {
  const float* m;  // 33 rows
  const float* b;  // 13 rows
  const float* f;  // 12 rows
  int C;
};

__device__ __forceinline__ float eb_logits(const EbSmem& s, int c, float v) {
  const int C = s.C;
  float h[3], g[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    float t = s.m[j * C + c] * v + s.b[j * C + c];
    h[j] = t + s.f[j * C + c] * tanhf(t);
  }
#pragma unroll
  for (int l = 0; l < 3; ++l) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const float* mr = s.m + (3 + l * 9 + i * 3) * C + c;
      float t = mr[0] * h[0] + mr[C] * h[1] + mr[2 * C] * h[2] + s.b[(3 + l * 3 + i) * C + c];
      g[i] = t + s.f[(3 + l * 3 + i) * C + c] * tanhf(t);
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) h[i] = g[i];
  }
  const float* mr = s.m + 30 * C + c;
  return mr[0] * h[0] + mr[C] * h[1] + mr[2 * C] * h[2] + s.b[12 * C + c];
}

// noise != nullptr: training-mode quantisation, z_hat = z + noise (compressai quantize(.., "noise"): the medians are not used)
__global__ void eb_bits_kernel(const float* __restrict__ z, const float* __restrict__ noise, float* __restrict__ z_hat,
                               const float* __restrict__ mats, const float* __restrict__ biases, const float* __restrict__ factors,
                               const float* __restrict__ medians, int64_t total, int C, double* acc) {
  extern __shared__ float sh[];
  float* sm = sh;                // [33][C]
  float* sb = sh + 33 * C;       // [13][C]
  float* sf = sb + 13 * C;       // [12][C]
  float* smed = sf + 12 * C;     // [C]
  for (int i = threadIdx.x; i < 33 * C; i += blockDim.x) sm[(i % 33) * C + i / 33] = mats[i];
  for (int i = threadIdx.x; i < 13 * C; i += blockDim.x) sb[(i % 13) * C + i / 13] = biases[i];
  for (int i = threadIdx.x; i < 12 * C; i += blockDim.x) sf[(i % 12) * C + i / 12] = factors[i];
  for (int i = threadIdx.x; i < C; i += blockDim.x) smed[i] = medians[i];
  __syncthreads();
  EbSmem s{sm, sb, sf, C};
  double sum = 0.0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int c = (int)(i % C);
    const float med = smed[c];
    const float q = noise != nullptr ? __fadd_rn(z[i], noise[i]) : __fadd_rn(rintf(__fsub_rn(z[i], med)), med);
    z_hat[i] = q;
    const float lower = eb_logits(s, c, q - 0.5f);
    const float upper = eb_logits(s, c, q + 0.5f);
    const float t = lower + upper;
    const float sign = t > 0.f ? -1.f : (t < 0.f ? 1.f : 0.f);
    float p = fabsf(sigmoidf_(sign * upper) - sigmoidf_(sign * lower));
    p = fmaxf(p, 1e-9f);
    sum += (double)logf(p);
  }
  block_accumulate(sum, acc);
}

// noise != nullptr: training mode, the likelihood is evaluated at y + noise instead of the dequantised value
__global__ void gc_bits_kernel(const float* __restrict__ y, const float* __restrict__ noise, const float* __restrict__ params,
                               int params_ld, int64_t total, int C, double* acc) {
  double sum = 0.0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const float kc = -0.70710678118654752440f;  // -(2 ** -0.5)
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t p = i / C;
    const int c = (int)(i - p * C);
    const float yv = __ldg(y + i);
    const float scale = __ldg(params + p * params_ld + c);
    const float mean = __ldg(params + p * params_ld + C + c);
    const float out = noise != nullptr ? __fadd_rn(yv, __ldg(noise + i))
                                       : __fadd_rn(rintf(__fsub_rn(yv, mean)), mean);  // dequantised value
    const float v = fabsf(__fsub_rn(out, mean));
    const float sc = fmaxf(scale, 0.11f);
    const float upper = 0.5f * erfcf(kc * __fdiv_rn(__fsub_rn(0.5f, v), sc));
    const float lower = 0.5f * erfcf(kc * __fdiv_rn(__fsub_rn(-0.5f, v), sc));
    float lik = fmaxf(upper - lower, 1e-9f);
    sum += (double)logf(lik);
  }
  block_accumulate(sum, acc);
}

// EntropyBottleneck.loss(): sum_c sum_k |logits_cumulative(quantiles[c][k]) - target[k]|  (compressai entropy_models.py;
// reference pnet.py:35,59 `aux_loss()`).  One block; fixed-order tree: deterministic.
__global__ void eb_aux_loss_kernel(const float* __restrict__ mats, const float* __restrict__ biases,
                                   const float* __restrict__ factors, const float* __restrict__ quantiles,
                                   const float* __restrict__ target, int C, float* __restrict__ out) {
  extern __shared__ float sh[];
  float* sm = sh;
  float* sb = sh + 33 * C;
  float* sf = sb + 13 * C;
  float* red = sf + 12 * C;   // [blockDim.x]
  for (int i = threadIdx.x; i < 33 * C; i += blockDim.x) sm[(i % 33) * C + i / 33] = mats[i];
  for (int i = threadIdx.x; i < 13 * C; i += blockDim.x) sb[(i % 13) * C + i / 13] = biases[i];
  for (int i = threadIdx.x; i < 12 * C; i += blockDim.x) sf[(i % 12) * C + i / 12] = factors[i];
  __syncthreads();
  EbSmem s{sm, sb, sf, C};
  float t = 0.f;
  for (int i = threadIdx.x; i < 3 * C; i += blockDim.x) {
    const int c = i / 3, k = i - 3 * c;
    t += fabsf(eb_logits(s, c, quantiles[i]) - target[k]);
  }
  red[threadIdx.x] = t;
  __syncthreads();
  for (int st = blockDim.x >> 1; st > 0; st >>= 1) {
    if (threadIdx.x < st) red[threadIdx.x] += red[threadIdx.x + st];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = red[0];
}

// d loss / d quantiles of eb_aux_loss (the only gradient the reference takes of it: `aux_loss.backward()` steps the aux
// optimiser on `.quantiles`, reference tools/train.py:101-113,147-159; compressai evaluates the logits with the matrices,
// biases and factors detached).  grad_q[c][k] = grad_out * sign(logits - target) * d logits / d q, one thread per (c, k).
__global__ void eb_aux_loss_grad_kernel(const float* __restrict__ mats, const float* __restrict__ biases,
                                        const float* __restrict__ factors, const float* __restrict__ quantiles,
                                        const float* __restrict__ target, const float* __restrict__ grad_out, int C,
                                        float* __restrict__ grad_q) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 3 * C) return;
  const int c = i / 3, k = i - 3 * c;
  const float* m = mats + c * 33;
  const float* b = biases + c * 13;
  const float* f = factors + c * 12;
  // forward value h and derivative d of every layer, widths 1,3,3,3,3,1
  float h[3], d[3], g[3], e[3];
  const float v = quantiles[i];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const float t = m[j] * v + b[j];
    const float th = tanhf(t);
    h[j] = t + f[j] * th;
    d[j] = m[j] * (1.f + f[j] * (1.f - th * th));
  }
#pragma unroll
  for (int l = 0; l < 3; ++l) {
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const float* mr = m + 3 + l * 9 + r * 3;
      const float t = mr[0] * h[0] + mr[1] * h[1] + mr[2] * h[2] + b[3 + l * 3 + r];
      const float dt = mr[0] * d[0] + mr[1] * d[1] + mr[2] * d[2];
      const float th = tanhf(t);
      const float fr = f[3 + l * 3 + r];
      g[r] = t + fr * th;
      e[r] = dt * (1.f + fr * (1.f - th * th));
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) { h[r] = g[r]; d[r] = e[r]; }
  }
  const float* mr = m + 30;
  const float logit = mr[0] * h[0] + mr[1] * h[1] + mr[2] * h[2] + b[12];
  const float dlogit = mr[0] * d[0] + mr[1] * d[1] + mr[2] * d[2];
  const float diff = logit - target[k];
  const float sgn = diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f);
  grad_q[i] = grad_out[0] * sgn * dlogit;
}

// Philox4x32-10 (Salmon et al. 2011), counter = element index / 4, key = seed: uniform noise in [-0.5, 0.5) for the
// training-mode quantisers (compressai quantize(.., "noise") draws torch.empty_like(x).uniform_(-0.5, 0.5)).
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u;
    key.y += 0xBB67AE85u;
  }
  return ctr;
}
__global__ void uniform_noise_kernel(float* __restrict__ out, int64_t n, uint64_t seed, uint64_t stream_id) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q * 4 < n; q += stride) {
    const uint4 r = philox4x32_10(make_uint4((uint32_t)q, (uint32_t)((uint64_t)q >> 32), (uint32_t)stream_id, (uint32_t)(stream_id >> 32)),
                                  make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    const uint32_t v[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (q * 4 + j < n) out[q * 4 + j] = (float)(v[j] >> 8) * (1.0f / 16777216.0f) - 0.5f;   // 24 random bits: exact in fp32
  }
}

// ---- real entropy coding: the "symbols" quantiser of z (compressai EntropyBottleneck.compress; reached from reference
// pnet.py:45-49,69-73 `compress()`).
// eb_symbols: z NHWC -> int32 symbols round(z - median) in NCHW order (the order compressai's EntropyBottleneck.compress
// flattens them in), indexes[i] = channel.
__global__ void eb_symbols_kernel(const float* __restrict__ z, int ld, const float* __restrict__ medians, int N, int HW, int C,
                                  int32_t* __restrict__ symbols, int32_t* __restrict__ indexes) {
  const int64_t total = (int64_t)N * C * HW;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int p = (int)(i % HW);
    const int c = (int)((i / HW) % C);
    const int n = (int)(i / ((int64_t)HW * C));
    const float v = __ldg(z + ((int64_t)n * HW + p) * ld + c);
    symbols[i] = __float2int_rn(__fsub_rn(v, medians[c]));
    indexes[i] = c;
  }
}

}  // namespace tdvc

using namespace tdvc;

extern "C" int tdvc_eb_bits(const float* z, float* z_hat, const float* mats, const float* biases, const float* factors,
                            const float* medians, int64_t npix, int C, double* acc, void* stream) {
  TDVC_REQUIRE(z && z_hat && mats && biases && factors && medians && acc && npix > 0 && C > 0, "eb_bits: bad args");
  const size_t smem = (size_t)(33 + 13 + 12 + 1) * C * sizeof(float);
  TDVC_REQUIRE(smem <= 48 * 1024, "eb_bits: C=%d too large", C);
  const int64_t total = npix * C;
  int grid = cdiv(total, 256);
  if (grid > kNumSMs * 4) grid = kNumSMs * 4;
  eb_bits_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(z, nullptr, z_hat, mats, biases, factors, medians, total, C, acc);
  TDVC_CHECK_LAUNCH("eb_bits");
  return TDVC_OK;
}

extern "C" int tdvc_eb_bits_noise(const float* z, const float* noise, float* z_tilde, const float* mats, const float* biases,
                                  const float* factors, int64_t npix, int C, double* acc, void* stream) {
  TDVC_REQUIRE(z && noise && z_tilde && mats && biases && factors && acc && npix > 0 && C > 0, "eb_bits_noise: bad args");
  const size_t smem = (size_t)(33 + 13 + 12 + 1) * C * sizeof(float);
  TDVC_REQUIRE(smem <= 48 * 1024, "eb_bits_noise: C=%d too large", C);
  const int64_t total = npix * C;
  int grid = cdiv(total, 256);
  if (grid > kNumSMs * 4) grid = kNumSMs * 4;
  eb_bits_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(z, noise, z_tilde, mats, biases, factors, biases /* unused */, total, C, acc);
  TDVC_CHECK_LAUNCH("eb_bits_noise");
  return TDVC_OK;
}

extern "C" int tdvc_eb_aux_loss(const float* mats, const float* biases, const float* factors, const float* quantiles,
                                const float* target3, int C, float* out, void* stream) {
  TDVC_REQUIRE(mats && biases && factors && quantiles && target3 && out && C > 0, "eb_aux_loss: bad args");
  const size_t smem = (size_t)((33 + 13 + 12) * C + 256) * sizeof(float);
  TDVC_REQUIRE(smem <= 48 * 1024, "eb_aux_loss: C=%d too large", C);
  eb_aux_loss_kernel<<<1, 256, smem, (cudaStream_t)stream>>>(mats, biases, factors, quantiles, target3, C, out);
  TDVC_CHECK_LAUNCH("eb_aux_loss");
  return TDVC_OK;
}

extern "C" int tdvc_eb_aux_loss_grad(const float* mats, const float* biases, const float* factors, const float* quantiles,
                                     const float* target3, const float* grad_out, int C, float* grad_quantiles, void* stream) {
  TDVC_REQUIRE(mats && biases && factors && quantiles && target3 && grad_out && grad_quantiles && C > 0, "eb_aux_loss_grad: bad args");
  eb_aux_loss_grad_kernel<<<cdiv(3 * C, 128), 128, 0, (cudaStream_t)stream>>>(mats, biases, factors, quantiles, target3, grad_out, C,
                                                                              grad_quantiles);
  TDVC_CHECK_LAUNCH("eb_aux_loss_grad");
  return TDVC_OK;
}

extern "C" int tdvc_uniform_noise(float* out, int64_t n, uint64_t seed, uint64_t stream_id, void* stream) {
  TDVC_REQUIRE(out && n >= 0, "uniform_noise: bad args");
  if (n == 0) return TDVC_OK;
  int grid = cdiv(n / 4 + 1, 256);
  if (grid > kNumSMs * 8) grid = kNumSMs * 8;
  uniform_noise_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(out, n, seed, stream_id);
  TDVC_CHECK_LAUNCH("uniform_noise");
  return TDVC_OK;
}

extern "C" int tdvc_gc_bits(const float* y, const float* params, int params_ld, int64_t npix, int C, double* acc, void* stream) {
  TDVC_REQUIRE(y && params && acc && npix > 0 && C > 0 && params_ld >= 2 * C, "gc_bits: bad args");
  const int64_t total = npix * C;
  int grid = cdiv(total, 256);
  if (grid > kNumSMs * 8) grid = kNumSMs * 8;
  gc_bits_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(y, nullptr, params, params_ld, total, C, acc);
  TDVC_CHECK_LAUNCH("gc_bits");
  return TDVC_OK;
}

extern "C" int tdvc_gc_bits_noise(const float* y, const float* noise, const float* params, int params_ld, int64_t npix, int C,
                                  double* acc, void* stream) {
  TDVC_REQUIRE(y && noise && params && acc && npix > 0 && C > 0 && params_ld >= 2 * C, "gc_bits_noise: bad args");
  const int64_t total = npix * C;
  int grid = cdiv(total, 256);
  if (grid > kNumSMs * 8) grid = kNumSMs * 8;
  gc_bits_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(y, noise, params, params_ld, total, C, acc);
  TDVC_CHECK_LAUNCH("gc_bits_noise");
  return TDVC_OK;
}

extern "C" int tdvc_eb_symbols(const float* z, int ld, const float* medians, int N, int HW, int C, int32_t* symbols,
                               int32_t* indexes, void* stream) {
  TDVC_REQUIRE(z && medians && symbols && indexes && N > 0 && HW > 0 && C > 0 && ld >= C, "eb_symbols: bad args");
  int grid = cdiv((int64_t)N * HW * C, 256);
  if (grid > kNumSMs * 8) grid = kNumSMs * 8;
  eb_symbols_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(z, ld, medians, N, HW, C, symbols, indexes);
  TDVC_CHECK_LAUNCH("eb_symbols");
  return TDVC_OK;
}
