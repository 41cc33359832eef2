// DCNv2 backward - the reference's `_ext.dcn_v2_backward` (reference main/utils/dcnv2/src/dcn_v2.h:48-92,
// src/cuda/dcn_v2_cuda.cu:97-216, src/cuda/dcn_v2_im2col_cuda.cu:55-123 (gradient / coordinate weights), :197-327 (col2im,
// col2im_coord)) for any kernel size / stride / padding / dilation / deformable-group width, contiguous NCHW fp32.
//
// The reference materialises grad_col = W^T * grad_output per image (`columns`, C*K x Ho*Wo), scatters it into grad_input with
// float atomicAdd (col2im: run-to-run different low bits), and re-reads it for grad_offset / grad_mask and, through a second
// im2col, for grad_weight.  Here nothing of size C*K*Ho*Wo is stored and every result is DETERMINISTIC:
//   dcn_bwd_sample_kernel  one thread per (n, group, tap, ho, wo): for each channel of the group it forms grad_col on the fly
//                          (a dot product over the output channels), accumulates grad_mask and both grad_offset components in
//                          registers (plain stores: each element has one owner) and adds the four bilinear corner terms of
//                          grad_input as 64-bit FIXED-POINT integers (atomicAdd on integers is associative, so the sum does not
//                          depend on the order in which threads arrive);
//   dcn_bwd_finish_kernel  fixed point -> fp32 grad_input;
//   dcn_im2col_kernel      the modulated samples as rows [position][c * K + k]; grad_weight is then the weight gradient of a
//                          1x1 convolution of those rows (tdvc_conv2d_wgrad: MMA contraction, fixed-order two-stage sum);
//   dcn_bwd_bias_kernel    grad_bias[o] = sum grad_output, same tree.
// The fixed-point scale is a power of two derived on the device from max|grad_output| and max|weight| (dcn_bwd_scale_kernel).
#include "common.cuh"

namespace tdvc {
namespace dcnb {

struct Geo {
  int N, C, O, H, W, Ho, Wo, kh, kw, sh, sw, ph, pw, dh, dw, dg;
};

// max |x| over a tensor -> *out (bits of a non-negative float order like ints)
__global__ void absmax_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ out) {
  float m = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) m = fmaxf(m, fabsf(x[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int*>(out), __float_as_int(m));
}

// scale[0] = 2^e with  O * max|w| * max|go| * 2^20 * 2^e <= 2^62 : a single grad_col term is at most O*max|w|*max|go|, and up to
// 2^20 terms of that size may meet in one input pixel before the 64-bit accumulator could overflow.  scale[1] = 1 / scale[0].
__global__ void scale_kernel(const float* __restrict__ maxes, int O, double* __restrict__ scale) {
  const double bound = fmax((double)O * (double)maxes[0] * (double)maxes[1], 1e-30);
  int e = 0;
  frexp(bound, &e);               // bound < 2^e
  const int k = 62 - 20 - e;
  scale[0] = ldexp(1.0, k);
  scale[1] = ldexp(1.0, -k);
}

__device__ __forceinline__ void fx_add(long long* acc, double v, double scale) {
  atomicAdd(reinterpret_cast<unsigned long long*>(acc), (unsigned long long)__double2ll_rn(v * scale));
}

// One CTA = 128 output positions of one (image, deformable group), all K taps.  grad_col(c, k, p) = sum_o weight[o][c][k] *
// grad_output[o][p] is a dot product over the output channels for every (channel of the group, tap): the thread keeps the
// OC_T grad_output values of its position in registers (read once per group instead of once per (channel, tap)) and the
// group's weight slice sits in shared memory as [tap][channel][o], read with 16-byte broadcast loads - 16 loads per 64 FMAs
// (the first version issued two global loads per FMA and spent 15.7 ms of a 110 ms training step on its load/store pipe).
// Output channels beyond OC_T are walked in further chunks.
constexpr int DB_OC = 64;     // output channels per register chunk
constexpr int DB_PX = 128;    // positions per CTA
__global__ void __launch_bounds__(DB_PX) dcn_bwd_sample_kernel(const float* __restrict__ input, const float* __restrict__ weight,
                                                               const float* __restrict__ offset, const float* __restrict__ mask,
                                                               const float* __restrict__ go, long long* __restrict__ gi_fx,
                                                               float* __restrict__ g_off, float* __restrict__ g_msk,
                                                               const double* __restrict__ scale_p, Geo q, int cpg_max) {
  extern __shared__ __align__(16) float wsm[];          // [K][cpg][DB_OC] for the current chunk of output channels
  const int K = q.kh * q.kw, cpg = q.C / q.dg;
  const int64_t hw = (int64_t)q.Ho * q.Wo, HW = (int64_t)q.H * q.W;
  const double scale = scale_p[0];
  const int g = blockIdx.y % q.dg, n = blockIdx.y / q.dg;
  const int64_t pix = (int64_t)blockIdx.x * DB_PX + threadIdx.x;
  const bool live = pix < hw;
  const int ho = live ? (int)(pix / q.Wo) : 0, wo = live ? (int)(pix - (int64_t)ho * q.Wo) : 0;
  const int n_chunks = (q.O + DB_OC - 1) / DB_OC;
  float gor[DB_OC];                                      // grad_output of this position, output channels of the current chunk
  if (n_chunks == 1 && live) {
    const float* go_n = go + (int64_t)n * q.O * hw + pix;
#pragma unroll
    for (int o = 0; o < DB_OC; ++o) gor[o] = o < q.O ? __ldg(go_n + (int64_t)o * hw) : 0.f;
  }
  for (int k = 0; k < K; ++k) {
    const int ki = k / q.kw, kj = k - ki * q.kw;
    const int64_t off_c = ((int64_t)n * q.dg * K + g * K + k) * 2;
    float dy = 0.f, dx = 0.f, m = 0.f;
    if (live) {
      dy = offset[off_c * hw + pix];
      dx = offset[(off_c + 1) * hw + pix];
      m = mask[((int64_t)n * q.dg * K + g * K + k) * hw + pix];
    }
    const float h_im = (float)(ho * q.sh - q.ph + ki * q.dh) + dy;
    const float w_im = (float)(wo * q.sw - q.pw + kj * q.dw) + dx;
    const bool inside = live && h_im > -1.f && w_im > -1.f && h_im < (float)q.H && w_im < (float)q.W;
    const float hf = floorf(h_im), wf = floorf(w_im);
    const int h_low = (int)hf, w_low = (int)wf, h_high = h_low + 1, w_high = w_low + 1;
    const float lh = h_im - hf, lw = w_im - wf, hh = 1.f - lh, hw_ = 1.f - lw;
    const bool t_ok = h_low >= 0, b_ok = h_high <= q.H - 1, l_ok = w_low >= 0, r_ok = w_high <= q.W - 1;
    float gm = 0.f, gh = 0.f, gw = 0.f;
    for (int cc0 = 0; cc0 < cpg; cc0 += cpg_max) {      // (cpg_max bounds the shared-memory slice; one pass for TDVC's 8)
      const int ncc = cpg - cc0 < cpg_max ? cpg - cc0 : cpg_max;
      float gc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) gc[j] = 0.f;
      for (int ch = 0; ch < n_chunks; ++ch) {
        const int o0 = ch * DB_OC;
        __syncthreads();
        // weight[o][c][k] -> wsm[cc][o - o0] for this tap (zero beyond O)
        for (int i = threadIdx.x; i < ncc * DB_OC; i += DB_PX) {
          const int cc = i / DB_OC, o = o0 + (i - cc * DB_OC);
          wsm[i] = o < q.O ? __ldg(weight + ((int64_t)o * q.C + g * cpg + cc0 + cc) * K + k) : 0.f;
        }
        __syncthreads();
        if (inside) {
          if (n_chunks > 1) {
            const float* go_n = go + ((int64_t)n * q.O + o0) * hw + pix;
#pragma unroll
            for (int o = 0; o < DB_OC; ++o) gor[o] = o0 + o < q.O ? __ldg(go_n + (int64_t)o * hw) : 0.f;
          }
          for (int cc = 0; cc < ncc; ++cc) {
            const float4* wv = reinterpret_cast<const float4*>(wsm + cc * DB_OC);
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
            for (int o4 = 0; o4 < DB_OC / 4; ++o4) {
              const float4 w4 = wv[o4];
              a0 = fmaf(w4.x, gor[4 * o4], a0);
              a1 = fmaf(w4.y, gor[4 * o4 + 1], a1);
              a2 = fmaf(w4.z, gor[4 * o4 + 2], a2);
              a3 = fmaf(w4.w, gor[4 * o4 + 3], a3);
            }
            const float v = (a0 + a1) + (a2 + a3);
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (j == cc) gc[j] += v;
          }
        }
      }
      if (inside) {
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) {
          if (cc >= ncc) break;
          const int c = g * cpg + cc0 + cc;
          const float gcv = gc[cc];
          const float* im = input + ((int64_t)n * q.C + c) * HW;
          const float v1 = (t_ok && l_ok) ? im[(int64_t)h_low * q.W + w_low] : 0.f;
          const float v2 = (t_ok && r_ok) ? im[(int64_t)h_low * q.W + w_high] : 0.f;
          const float v3 = (b_ok && l_ok) ? im[(int64_t)h_high * q.W + w_low] : 0.f;
          const float v4 = (b_ok && r_ok) ? im[(int64_t)h_high * q.W + w_high] : 0.f;
          // grad_mask: grad_col * bilinear(im)   (dcn_v2_im2col_cuda.cu:307-310)
          gm = fmaf(gcv, hh * hw_ * v1 + hh * lw * v2 + lh * hw_ * v3 + lh * lw * v4, gm);
          // grad_offset: grad_col * mask * d bilinear / d (h, w)   (dmcn_get_coordinate_weight_cuda, :82-123)
          gh = fmaf(gcv * m, -hw_ * v1 - lw * v2 + hw_ * v3 + lw * v4, gh);
          gw = fmaf(gcv * m, -hh * v1 + hh * v2 - lh * v3 + lh * v4, gw);
          // grad_input: the four corners get their bilinear weight of grad_col * mask   (col2im, :197-252)
          const double top = (double)gcv * (double)m;
          long long* gp = gi_fx + ((int64_t)n * q.C + c) * HW;
          if (t_ok && l_ok) fx_add(gp + (int64_t)h_low * q.W + w_low, top * (double)(hh * hw_), scale);
          if (t_ok && r_ok) fx_add(gp + (int64_t)h_low * q.W + w_high, top * (double)(hh * lw), scale);
          if (b_ok && l_ok) fx_add(gp + (int64_t)h_high * q.W + w_low, top * (double)(lh * hw_), scale);
          if (b_ok && r_ok) fx_add(gp + (int64_t)h_high * q.W + w_high, top * (double)(lh * lw), scale);
        }
      }
    }
    if (live) {
      g_off[off_c * hw + pix] = gh;
      g_off[(off_c + 1) * hw + pix] = gw;
      g_msk[((int64_t)n * q.dg * K + g * K + k) * hw + pix] = gm;
    }
  }
}

__global__ void dcn_bwd_finish_kernel(const long long* __restrict__ fx, float* __restrict__ out, int64_t n,
                                      const double* __restrict__ scale_p) {
  const double inv = scale_p[1];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = (float)((double)fx[i] * inv);
}

// grad_weight[o][c][k] = sum_{n, ho, wo} grad_output[n][o][ho][wo] * col[n, ho, wo][c][k], col = mask * bilinear(input[n][c],
// sample(n, g, k, ho, wo)): a GEMM over the output positions.  dcn_im2col_kernel writes col as rows [position][c * K + k] (one
// thread per (position, group, tap): the bilinear weights are formed once and applied to the channels of the group) and
// tdvc_conv2d_wgrad contracts it with the channels-last grad_output as a 1x1 "convolution" (csrc/conv_bwd.cu; deterministic
// two-stage sum).  (A first version - one CTA per (c, k) walking every position and re-reading grad_output - took 70 ms of a
// 440 ms training step: 576 CTAs each streamed the whole 134 MB gradient.)
__global__ void __launch_bounds__(256) dcn_im2col_kernel(const float* __restrict__ input, const float* __restrict__ offset,
                                                         const float* __restrict__ mask, float* __restrict__ col, int col_ld, Geo q) {
  const int K = q.kh * q.kw, cpg = q.C / q.dg;
  const int64_t hw = (int64_t)q.Ho * q.Wo, HW = (int64_t)q.H * q.W;
  const int64_t total = (int64_t)q.N * hw * q.dg * K;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int k = (int)(i % K);
    int64_t r = i / K;
    const int g = (int)(r % q.dg);
    const int64_t p = r / q.dg;                      // n * hw + pix
    const int n = (int)(p / hw);
    const int64_t pix = p - (int64_t)n * hw;
    const int ho = (int)(pix / q.Wo), wo = (int)(pix - (int64_t)ho * q.Wo);
    const int ki = k / q.kw, kj = k - ki * q.kw;
    const int64_t off_c = ((int64_t)n * q.dg * K + g * K + k) * 2;
    const float h_im = (float)(ho * q.sh - q.ph + ki * q.dh) + offset[off_c * hw + pix];
    const float w_im = (float)(wo * q.sw - q.pw + kj * q.dw) + offset[(off_c + 1) * hw + pix];
    float* dst = col + p * (int64_t)col_ld + (int64_t)g * cpg * K + k;
    if (!(h_im > -1.f && w_im > -1.f && h_im < (float)q.H && w_im < (float)q.W)) {
      for (int cc = 0; cc < cpg; ++cc) dst[cc * K] = 0.f;
      continue;
    }
    const float hf = floorf(h_im), wf = floorf(w_im);
    const int h_low = (int)hf, w_low = (int)wf, h_high = h_low + 1, w_high = w_low + 1;
    const float lh = h_im - hf, lw = w_im - wf, hh = 1.f - lh, hw_ = 1.f - lw;
    const float m = mask[((int64_t)n * q.dg * K + g * K + k) * hw + pix];
    const bool b1 = h_low >= 0 && w_low >= 0, b2 = h_low >= 0 && w_high <= q.W - 1, b3 = h_high <= q.H - 1 && w_low >= 0,
               b4 = h_high <= q.H - 1 && w_high <= q.W - 1;
    const float* im = input + ((int64_t)n * q.C + (int64_t)g * cpg) * HW;
    for (int cc = 0; cc < cpg; ++cc, im += HW) {
      const float v1 = b1 ? im[(int64_t)h_low * q.W + w_low] : 0.f;
      const float v2 = b2 ? im[(int64_t)h_low * q.W + w_high] : 0.f;
      const float v3 = b3 ? im[(int64_t)h_high * q.W + w_low] : 0.f;
      const float v4 = b4 ? im[(int64_t)h_high * q.W + w_high] : 0.f;
      dst[cc * K] = (hh * hw_ * v1 + hh * lw * v2 + lh * hw_ * v3 + lh * lw * v4) * m;
    }
  }
}

__global__ void __launch_bounds__(256) dcn_bwd_bias_kernel(const float* __restrict__ go, float* __restrict__ g_b, int N, int O, int64_t hw) {
  __shared__ float red[256];
  const int o = blockIdx.x;
  float s = 0.f;
  for (int64_t p = threadIdx.x; p < (int64_t)N * hw; p += blockDim.x) {
    const int64_t n = p / hw;
    s += go[(n * O + o) * hw + (p - n * hw)];
  }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int st = 128; st > 0; st >>= 1) {
    if (threadIdx.x < st) red[threadIdx.x] += red[threadIdx.x + st];
    __syncthreads();
  }
  if (threadIdx.x == 0) g_b[o] = red[0];
}

}  // namespace dcnb
}  // namespace tdvc

using namespace tdvc;

extern "C" size_t tdvc_conv2d_wgrad_workspace_bytes(int N, int Ho, int Wo, int cin, int cout, int k);
extern "C" int tdvc_conv2d_wgrad(const float* x, int x_ld, const float* grad_y, int g_ld, int N, int H, int W, int cin, int cout,
                                 int k, int stride, int pad, int in_square, int products, float* grad_w, float* grad_b_or_null,
                                 void* workspace, size_t workspace_bytes, void* stream);
extern "C" int tdvc_nchw_to_nhwc(const float* src, float* dst, int N, int C, int H, int W, int dst_ld, void* stream);

static size_t dcn_bwd_fixed_bytes(int N, int C, int H, int W) {
  return ((size_t)N * C * H * W * sizeof(long long) + 256 + 255) / 256 * 256;   // fixed-point grad_input + {max|go|, max|w|, scale, 1/scale}
}
static size_t align256(size_t b) { return (b + 255) / 256 * 256; }

// fixed-point grad_input | col rows [N*Ho*Wo][C*K] | channels-last grad_output | wgrad partial sums
extern "C" size_t tdvc_dcn_v2_backward_workspace_bytes(int N, int C, int O, int H, int W, int Ho, int Wo, int K) {
  if (N <= 0 || C <= 0 || O <= 0 || H <= 0 || W <= 0 || Ho <= 0 || Wo <= 0 || K <= 0) return 0;
  const size_t npos = (size_t)N * Ho * Wo;
  return dcn_bwd_fixed_bytes(N, C, H, W) + align256(npos * ((C * K + 3) / 4 * 4) * sizeof(float)) + align256(npos * ((O + 3) / 4 * 4) * sizeof(float)) +
         tdvc_conv2d_wgrad_workspace_bytes(N, Ho, Wo, C * K, O, 1);
}

extern "C" int tdvc_dcn_v2_backward(const float* input, const float* weight, const float* offset, const float* mask,
                                    const float* grad_output, float* grad_input, float* grad_offset, float* grad_mask,
                                    float* grad_weight, float* grad_bias, int N, int C, int O, int H, int W, int kh, int kw,
                                    int sh, int sw, int ph, int pw, int dh, int dw, int dg, void* workspace,
                                    size_t workspace_bytes, void* stream) {
  TDVC_REQUIRE(input && weight && offset && mask && grad_output && grad_input && grad_offset && grad_mask && grad_weight && grad_bias,
               "dcn_v2_backward: null pointer");
  TDVC_REQUIRE(N > 0 && C > 0 && O > 0 && H > 0 && W > 0 && dg > 0 && C % dg == 0, "dcn_v2_backward: bad shape");
  TDVC_REQUIRE(kh > 0 && kw > 0 && sh > 0 && sw > 0 && dh > 0 && dw > 0 && ph >= 0 && pw >= 0, "dcn_v2_backward: bad conv geometry");
  const int Ho = (H + 2 * ph - (dh * (kh - 1) + 1)) / sh + 1, Wo = (W + 2 * pw - (dw * (kw - 1) + 1)) / sw + 1;
  TDVC_REQUIRE(Ho > 0 && Wo > 0, "dcn_v2_backward: empty output");
  const size_t need = tdvc_dcn_v2_backward_workspace_bytes(N, C, O, H, W, Ho, Wo, kh * kw);
  TDVC_REQUIRE(workspace != nullptr && workspace_bytes >= need && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0,
               "dcn_v2_backward: workspace too small or misaligned (%zu < %zu)", workspace_bytes, need);

  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int64_t n_in = (int64_t)N * C * H * W;
  long long* fx = static_cast<long long*>(workspace);
  float* maxes = reinterpret_cast<float*>(fx + n_in);
  double* scale = reinterpret_cast<double*>(fx + n_in + 2);
  if (cudaMemsetAsync(workspace, 0, (size_t)n_in * sizeof(long long) + 64, st) != cudaSuccess) {
    set_error("dcn_v2_backward: cudaMemsetAsync failed");
    return TDVC_ECUDA;
  }
  const dcnb::Geo q{N, C, O, H, W, Ho, Wo, kh, kw, sh, sw, ph, pw, dh, dw, dg};
  const int64_t n_go = (int64_t)N * O * Ho * Wo, n_w = (int64_t)O * C * kh * kw;
  auto grid_for = [](int64_t n, int t) { int64_t b = (n + t - 1) / t; return (int)(b < 1 ? 1 : (b > kNumSMs * 16 ? kNumSMs * 16 : b)); };
  dcnb::absmax_kernel<<<grid_for(n_go, 256), 256, 0, st>>>(grad_output, n_go, maxes);
  dcnb::absmax_kernel<<<grid_for(n_w, 256), 256, 0, st>>>(weight, n_w, maxes + 1);
  dcnb::scale_kernel<<<1, 1, 0, st>>>(maxes, O, scale);
  {
    const int64_t tiles = ((int64_t)Ho * Wo + dcnb::DB_PX - 1) / dcnb::DB_PX;
    TDVC_REQUIRE(tiles < (1ll << 31) && (int64_t)N * dg < 65536, "dcn_v2_backward: too many positions / groups for one launch");
    const int cpg_max = 8;   // channels of a group per pass (the register accumulators of the sample kernel)
    const size_t smem = (size_t)cpg_max * dcnb::DB_OC * sizeof(float);
    dcnb::dcn_bwd_sample_kernel<<<dim3((unsigned)tiles, (unsigned)(N * dg)), dcnb::DB_PX, smem, st>>>(
        input, weight, offset, mask, grad_output, fx, grad_offset, grad_mask, scale, q, cpg_max);
  }
  TDVC_CHECK_LAUNCH("dcn_bwd_sample");
  dcnb::dcn_bwd_finish_kernel<<<grid_for(n_in, 256), 256, 0, st>>>(fx, grad_input, n_in, scale);
  {
    const size_t npos = (size_t)N * Ho * Wo;
    const int K = kh * kw, Op = (O + 3) / 4 * 4, col_ld = (C * K + 3) / 4 * 4;   // (the row tails beyond C * K are never read)
    char* base = static_cast<char*>(workspace) + dcn_bwd_fixed_bytes(N, C, H, W);
    float* col = reinterpret_cast<float*>(base);
    float* go_nhwc = reinterpret_cast<float*>(base + align256(npos * col_ld * sizeof(float)));
    char* wws = reinterpret_cast<char*>(go_nhwc) + align256(npos * Op * sizeof(float));
    dcnb::dcn_im2col_kernel<<<grid_for((int64_t)npos * dg * K, 256), 256, 0, st>>>(input, offset, mask, col, col_ld, q);
    TDVC_CHECK_LAUNCH("dcn_im2col");
    int rc = tdvc_nchw_to_nhwc(grad_output, go_nhwc, N, O, Ho, Wo, Op, st);
    if (rc != TDVC_OK) return rc;
    rc = tdvc_conv2d_wgrad(col, col_ld, go_nhwc, Op, N, Ho, Wo, C * K, O, 1, 1, 0, 0, 3, grad_weight, nullptr, wws,
                           tdvc_conv2d_wgrad_workspace_bytes(N, Ho, Wo, C * K, O, 1), st);
    if (rc != TDVC_OK) return rc;
  }
  dcnb::dcn_bwd_bias_kernel<<<O, 256, 0, st>>>(grad_output, grad_bias, N, O, (int64_t)Ho * Wo);
  TDVC_CHECK_LAUNCH("dcn_bwd");
  return TDVC_OK;
}
