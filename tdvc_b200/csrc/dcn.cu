// DCNv2 forward (modulated deformable 3x3 convolution) — the reference's only native op.
//
// Reference: main/utils/dcnv2/src/cuda/dcn_v2_cuda.cu:20-95 (alloc `columns` (B, C*9, H*W), im2col launch,
// at::matmul(weight_flat, columns) + bias) and src/cuda/dcn_v2_im2col_cuda.cu:25-54,125-195 (the bilinear
// rule).  At 1920x1024 the reference writes and re-reads a 4.5 GB `columns` tensor; here the modulated
// columns of a 64-pixel tile live only in shared memory and are contracted immediately:
//
//   CTA = 8x8 output pixels, 256 threads.  For each deformable group g (8 input channels):
//     gather : thread -> (pixel, tap); reads dy,dx,mask once, does the 4-corner bilinear gather for the
//              group's 8 channels as two float4 loads per corner (channels-last: 32 contiguous bytes),
//              writes col[c*9+tap][pixel] to smem;
//     GEMM   : acc[4 px][4 out] += col[k][px] * W[g*72+k][out], K = 72 per group, weights staged in smem.
//   Epilogue: + bias, optional fp16 rounding (`_DCNv2.forward` returns output.half(),
//   dcn_v2_amp.py:67-69), optional LeakyReLU evaluated like torch does on a Half tensor
//   (half(float(h) * slope)), store float4.
//
// The fused form reads offsets/mask logits straight from the conv_offset_mask output (dcn_v2_amp.py:219-234).
// A plain NCHW entry point with the exact `_ext.dcn_v2_forward` contract wraps it (tdvc_dcn_v2_forward);
// configurations outside the fast path (channels per group != 8, other kernel sizes/strides) run a
// direct one-thread-per-output kernel.
#include <string.h>
#include "common.cuh"

namespace tdvc {

constexpr int DCN_TP = 8;            // tile edge (pixels)
constexpr int DCN_PIX = DCN_TP * DCN_TP;
constexpr int DCN_THREADS = 256;
constexpr int DCN_KG = 72;           // 8 channels x 9 taps per group

__device__ __forceinline__ float sigmoid_exact(float x) { return 1.f / (1.f + expf(-x)); }

template <int OT>  // output channels per CTA pass (64)
__global__ void __launch_bounds__(DCN_THREADS) dcn_nhwc_kernel(const TdvcDcnParams p, int tiles_x, int tiles_y) {
  __shared__ __align__(16) float col[DCN_KG][DCN_PIX];   // 18 KB
  __shared__ __align__(16) float wsm[DCN_KG][OT];        // 18 KB
  const int tid = threadIdx.x;
  int t = blockIdx.x;
  const int tx = t % tiles_x; t /= tiles_x;
  const int ty = t % tiles_y;
  const int n = t / tiles_y;
  const int oc0 = blockIdx.y * OT;
  const int x0 = tx * DCN_TP, y0 = ty * DCN_TP;
  const int H = p.H, W = p.W;

  // GEMM mapping: 16 pixel-quads x 16 channel-quads
  const int pq = tid & 15, cq = tid >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int64_t img = (int64_t)n * H * W;
  for (int g = 0; g < p.dg; ++g) {
    __syncthreads();  // previous group's col/wsm consumed
    // ---- stage this group's weights: rows g*72 .. g*72+71, OT columns
    for (int idx = tid; idx < DCN_KG * OT / 4; idx += DCN_THREADS) {
      const int e = idx * 4;
      const int k = e / OT, co = e - k * OT;
      *reinterpret_cast<float4*>(&wsm[k][co]) =
          __ldg(reinterpret_cast<const float4*>(p.weight_packed + (int64_t)(g * DCN_KG + k) * p.O_pad + oc0 + co));
    }
    // ---- gather: 64 pixels x 9 taps
    for (int idx = tid; idx < DCN_PIX * 9; idx += DCN_THREADS) {
      const int px = idx & (DCN_PIX - 1);
      const int tap = idx >> 6;
      const int ly = px >> 3, lx = px & 7;
      const int y = y0 + ly, x = x0 + lx;
      float v[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) v[c] = 0.f;
      if (y < H && x < W) {
        const int64_t pix = img + (int64_t)y * W + x;
        const float* op = p.offset + pix * p.off_ld + g * 18 + 2 * tap;
        const float dy = __ldg(op), dx = __ldg(op + 1);
        float m = __ldg(p.mask + pix * p.mask_ld + g * 9 + tap);
        if (p.mask_is_logit) m = sigmoid_exact(m);
        const int ki = tap / 3, kj = tap - ki * 3;
        const float h_im = (float)(y - 1 + ki) + dy;
        const float w_im = (float)(x - 1 + kj) + dx;
        if (h_im > -1.f && w_im > -1.f && h_im < (float)H && w_im < (float)W) {
          const float hf = floorf(h_im), wf = floorf(w_im);
          const int h_low = (int)hf, w_low = (int)wf;
          const int h_high = h_low + 1, w_high = w_low + 1;
          const float lh = h_im - hf, lw = w_im - wf;
          const float hh = 1.f - lh, hw = 1.f - lw;
          const float w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
          const float* base = p.input + img * p.in_ld + g * 8;
          float4 a0, a1;
          const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
          float4 c1a = z, c1b = z, c2a = z, c2b = z, c3a = z, c3b = z, c4a = z, c4b = z;
          if (h_low >= 0 && w_low >= 0) {
            const float4* q = reinterpret_cast<const float4*>(base + ((int64_t)h_low * W + w_low) * p.in_ld);
            c1a = __ldg(q); c1b = __ldg(q + 1);
          }
          if (h_low >= 0 && w_high <= W - 1) {
            const float4* q = reinterpret_cast<const float4*>(base + ((int64_t)h_low * W + w_high) * p.in_ld);
            c2a = __ldg(q); c2b = __ldg(q + 1);
          }
          if (h_high <= H - 1 && w_low >= 0) {
            const float4* q = reinterpret_cast<const float4*>(base + ((int64_t)h_high * W + w_low) * p.in_ld);
            c3a = __ldg(q); c3b = __ldg(q + 1);
          }
          if (h_high <= H - 1 && w_high <= W - 1) {
            const float4* q = reinterpret_cast<const float4*>(base + ((int64_t)h_high * W + w_high) * p.in_ld);
            c4a = __ldg(q); c4b = __ldg(q + 1);
          }
          // same evaluation order as the reference: w1*v1 + w2*v2 + w3*v3 + w4*v4, then * mask
          a0.x = (w1 * c1a.x + w2 * c2a.x + w3 * c3a.x + w4 * c4a.x) * m;
          a0.y = (w1 * c1a.y + w2 * c2a.y + w3 * c3a.y + w4 * c4a.y) * m;
          a0.z = (w1 * c1a.z + w2 * c2a.z + w3 * c3a.z + w4 * c4a.z) * m;
          a0.w = (w1 * c1a.w + w2 * c2a.w + w3 * c3a.w + w4 * c4a.w) * m;
          a1.x = (w1 * c1b.x + w2 * c2b.x + w3 * c3b.x + w4 * c4b.x) * m;
          a1.y = (w1 * c1b.y + w2 * c2b.y + w3 * c3b.y + w4 * c4b.y) * m;
          a1.z = (w1 * c1b.z + w2 * c2b.z + w3 * c3b.z + w4 * c4b.z) * m;
          a1.w = (w1 * c1b.w + w2 * c2b.w + w3 * c3b.w + w4 * c4b.w) * m;
          v[0] = a0.x; v[1] = a0.y; v[2] = a0.z; v[3] = a0.w;
          v[4] = a1.x; v[5] = a1.y; v[6] = a1.z; v[7] = a1.w;
        }
      }
#pragma unroll
      for (int c = 0; c < 8; ++c) col[c * 9 + tap][px] = v[c];
    }
    __syncthreads();
    // ---- contraction over this group's 72 columns
#pragma unroll 8
    for (int k = 0; k < DCN_KG; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&col[k][pq * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&wsm[k][cq * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
  }

  // ---- epilogue
  const int co = oc0 + cq * 4;
  if (co >= p.O) return;
  float bz[4] = {0.f, 0.f, 0.f, 0.f};
  if (p.bias) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (co + j < p.O) bz[j] = __ldg(p.bias + co + j);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int px = pq * 4 + i;
    const int y = y0 + (px >> 3), x = x0 + (px & 7);
    if (y >= H || x >= W) continue;
    float r[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float v = acc[i][j] + bz[j];
      if (p.round_fp16) {
        __half hv = __float2half_rn(v);
        v = __half2float(hv);
        if (p.act == TDVC_ACT_LRELU) {
          if (v < 0.f) v = __half2float(__float2half_rn(v * p.slope));  // torch CPU Half leaky_relu
        } else {
          v = apply_act(v, p.act, p.slope);
        }
      } else {
        v = apply_act(v, p.act, p.slope);
      }
      r[j] = v;
    }
    float* o = p.out + (img + (int64_t)y * W + x) * p.out_ld + co;
    if (co + 3 < p.O && (p.out_ld & 3) == 0) {
      *reinterpret_cast<float4*>(o) = make_float4(r[0], r[1], r[2], r[3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (co + j < p.O) o[j] = r[j];
    }
  }
}

// Direct NCHW kernel for configurations outside the fast path (any kernel size / stride / pad / dilation /
// channels-per-group).  One thread per output element; same arithmetic rule.
__global__ void dcn_generic_nchw_kernel(const float* __restrict__ input, const float* __restrict__ weight,
                                        const float* __restrict__ bias, const float* __restrict__ offset,
                                        const float* __restrict__ mask, float* __restrict__ output,
                                        int N, int C, int O, int H, int W, int Ho, int Wo, int kh, int kw,
                                        int sh, int sw, int ph, int pw, int dh, int dw, int dg) {
  const int64_t total = (int64_t)N * O * Ho * Wo;
  const int cpg = C / dg;
  const int K = kh * kw;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int wo = (int)(i % Wo);
    int64_t r = i / Wo;
    const int ho = (int)(r % Ho); r /= Ho;
    const int o = (int)(r % O);
    const int n = (int)(r / O);
    float acc = bias ? bias[o] : 0.f;
    const int64_t hw = (int64_t)Ho * Wo;
    for (int c = 0; c < C; ++c) {
      const int g = c / cpg;
      const float* im = input + ((int64_t)n * C + c) * H * W;
      for (int k = 0; k < K; ++k) {
        const int ki = k / kw, kj = k - ki * kw;
        const float dy = offset[((int64_t)n * dg * 2 * K + (g * K + k) * 2) * hw + (int64_t)ho * Wo + wo];
        const float dx = offset[((int64_t)n * dg * 2 * K + (g * K + k) * 2 + 1) * hw + (int64_t)ho * Wo + wo];
        const float m = mask[((int64_t)n * dg * K + g * K + k) * hw + (int64_t)ho * Wo + wo];
        const float h_im = (float)(ho * sh - ph + ki * dh) + dy;
        const float w_im = (float)(wo * sw - pw + kj * dw) + dx;
        float val = 0.f;
        if (h_im > -1.f && w_im > -1.f && h_im < (float)H && w_im < (float)W) {
          const float hf = floorf(h_im), wf = floorf(w_im);
          const int h_low = (int)hf, w_low = (int)wf, h_high = h_low + 1, w_high = w_low + 1;
          const float lh = h_im - hf, lw = w_im - wf, hh = 1.f - lh, hw_ = 1.f - lw;
          const float v1 = (h_low >= 0 && w_low >= 0) ? im[(int64_t)h_low * W + w_low] : 0.f;
          const float v2 = (h_low >= 0 && w_high <= W - 1) ? im[(int64_t)h_low * W + w_high] : 0.f;
          const float v3 = (h_high <= H - 1 && w_low >= 0) ? im[(int64_t)h_high * W + w_low] : 0.f;
          const float v4 = (h_high <= H - 1 && w_high <= W - 1) ? im[(int64_t)h_high * W + w_high] : 0.f;
          val = hh * hw_ * v1 + hh * lw * v2 + lh * hw_ * v3 + lh * lw * v4;
        }
        acc = fmaf(weight[((int64_t)o * C + c) * K + k], val * m, acc);
      }
    }
    output[i] = acc;
  }
}

// weight (O, C, 3, 3) -> packed [C*9][O_pad], row = c*9 + tap, zero padded columns
__global__ void dcn_pack_weight_kernel(const float* __restrict__ w, float* __restrict__ packed, int O, int C, int O_pad) {
  const int total = C * 9 * O_pad;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int o = i % O_pad, row = i / O_pad;
    packed[i] = o < O ? w[(int64_t)o * C * 9 + row] : 0.f;
  }
}

static int dcn_nhwc_launch(const TdvcDcnParams& p, cudaStream_t st) {
  const int tiles_x = cdiv(p.W, DCN_TP), tiles_y = cdiv(p.H, DCN_TP);
  dim3 grid((unsigned)((int64_t)tiles_x * tiles_y * p.N), p.O_pad / 64);
  dcn_nhwc_kernel<64><<<grid, DCN_THREADS, 0, st>>>(p, tiles_x, tiles_y);
  TDVC_CHECK_LAUNCH("dcn_nhwc");
  return TDVC_OK;
}

int dcn_tc_supported(const TdvcDcnParams& p);
int dcn_tc(const TdvcDcnParams& p, cudaStream_t st);

}  // namespace tdvc

using namespace tdvc;

extern "C" int tdvc_dcn_nhwc(const TdvcDcnParams* p, void* stream) {
  TDVC_REQUIRE(p && (p->input || p->input_gp) && p->offset && p->mask && p->weight_packed && p->out, "dcn_nhwc: null pointer");
  TDVC_REQUIRE(p->N > 0 && p->H > 0 && p->W > 0 && p->dg > 0, "dcn_nhwc: empty input");
  TDVC_REQUIRE(p->C == 8 * p->dg, "dcn_nhwc: needs 8 channels per deformable group (C=%d dg=%d)", p->C, p->dg);
  TDVC_REQUIRE(p->O > 0 && p->O_pad % 64 == 0 && p->O_pad >= p->O, "dcn_nhwc: O=%d O_pad=%d", p->O, p->O_pad);
  TDVC_REQUIRE(p->params_planar == 0 || p->params_planar == 1, "dcn_nhwc: params_planar %d", p->params_planar);
  if (p->params_planar) {  // gather-friendly layouts: tensor-core kernel only
    TDVC_REQUIRE(p->impl != 1, "dcn_nhwc: planar offsets / group-planar input need the tcgen05 kernel (impl 0 or 2)");
    TDVC_REQUIRE(dcn_tc_supported(*p), "dcn_nhwc: tcgen05 kernel needs input_gp (32-byte aligned), weight_f16, O <= 64");
    TDVC_REQUIRE(p->off_ld >= 18 * p->dg && p->mask_ld >= 9 * p->dg, "dcn_nhwc: plane counts off_ld=%d mask_ld=%d", p->off_ld, p->mask_ld);
    return dcn_tc(*p, reinterpret_cast<cudaStream_t>(stream));
  }
  TDVC_REQUIRE(p->input != nullptr, "dcn_nhwc: channels-last offsets need the channels-last input");
  TDVC_REQUIRE(p->in_ld % 4 == 0 && (reinterpret_cast<uintptr_t>(p->input) & 15) == 0, "dcn_nhwc: input alignment");
  TDVC_REQUIRE((reinterpret_cast<uintptr_t>(p->weight_packed) & 15) == 0, "dcn_nhwc: weight alignment");
  TDVC_REQUIRE((reinterpret_cast<uintptr_t>(p->out) & 15) == 0 || p->out_ld % 4 != 0, "dcn_nhwc: out alignment");
  TDVC_REQUIRE((int64_t)cdiv(p->W, DCN_TP) * cdiv(p->H, DCN_TP) * p->N < (1ll << 31), "dcn_nhwc: grid too large");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  TDVC_REQUIRE(p->impl != 2, "dcn_nhwc: impl=2 (tcgen05) needs params_planar = 1 and input_gp");
  return dcn_nhwc_launch(*p, st);
}

extern "C" size_t tdvc_dcn_v2_workspace_bytes(int N, int C, int O, int H, int W, int dg) {
  if (N <= 0 || C <= 0 || O <= 0 || H <= 0 || W <= 0 || dg <= 0) return 0;
  const size_t px = (size_t)N * H * W;
  const size_t o_pad = (size_t)((O + 63) / 64) * 64;
  // NHWC copies of input, offset, mask, output + packed weights (+ alignment slack)
  return (px * ((size_t)C + 27 * (size_t)dg + (size_t)((O + 3) / 4 * 4)) + (size_t)C * 9 * o_pad) * sizeof(float) + 5 * 256;
}

extern "C" int tdvc_dcn_v2_forward(const float* input, const float* weight, const float* bias, const float* offset,
                                   const float* mask, float* output, int N, int C, int O, int H, int W,
                                   int kh, int kw, int sh, int sw, int ph, int pw, int dh, int dw, int dg,
                                   void* workspace, size_t workspace_bytes, void* stream) {
  TDVC_REQUIRE(input && weight && offset && mask && output, "dcn_v2_forward: null pointer");
  TDVC_REQUIRE(N > 0 && C > 0 && O > 0 && H > 0 && W > 0 && dg > 0 && C % dg == 0, "dcn_v2_forward: bad shape");
  TDVC_REQUIRE(kh > 0 && kw > 0 && sh > 0 && sw > 0 && dh > 0 && dw > 0 && ph >= 0 && pw >= 0, "dcn_v2_forward: bad conv geometry");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool fast = kh == 3 && kw == 3 && sh == 1 && sw == 1 && ph == 1 && pw == 1 && dh == 1 && dw == 1 && C == 8 * dg;
  if (!fast) {
    const int Ho = (H + 2 * ph - (dh * (kh - 1) + 1)) / sh + 1, Wo = (W + 2 * pw - (dw * (kw - 1) + 1)) / sw + 1;
    TDVC_REQUIRE(Ho > 0 && Wo > 0, "dcn_v2_forward: empty output");
    const int64_t total = (int64_t)N * O * Ho * Wo;
    int grid = cdiv(total, 128);
    if (grid > kNumSMs * 16) grid = kNumSMs * 16;
    dcn_generic_nchw_kernel<<<grid, 128, 0, st>>>(input, weight, bias, offset, mask, output, N, C, O, H, W, Ho, Wo,
                                                  kh, kw, sh, sw, ph, pw, dh, dw, dg);
    TDVC_CHECK_LAUNCH("dcn_generic");
    return TDVC_OK;
  }
  TDVC_REQUIRE(workspace != nullptr && workspace_bytes >= tdvc_dcn_v2_workspace_bytes(N, C, O, H, W, dg),
               "dcn_v2_forward: workspace too small (%zu < %zu)", workspace_bytes, tdvc_dcn_v2_workspace_bytes(N, C, O, H, W, dg));
  const size_t px = (size_t)N * H * W;
  const int o_pad = (O + 63) / 64 * 64;
  const int o_ld = (O + 3) / 4 * 4;
  auto align = [](char* q) { return reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(q) + 255) & ~(uintptr_t)255); };
  char* w = align(static_cast<char*>(workspace));
  float* in_nhwc = reinterpret_cast<float*>(w); w = align(w + px * C * sizeof(float));
  float* off_nhwc = reinterpret_cast<float*>(w); w = align(w + px * 18 * dg * sizeof(float));
  float* msk_nhwc = reinterpret_cast<float*>(w); w = align(w + px * 9 * dg * sizeof(float));
  float* out_nhwc = reinterpret_cast<float*>(w); w = align(w + px * o_ld * sizeof(float));
  float* wpk = reinterpret_cast<float*>(w);
  int rc;
  if ((rc = tdvc_nchw_to_nhwc(input, in_nhwc, N, C, H, W, C, stream)) != TDVC_OK) return rc;
  if ((rc = tdvc_nchw_to_nhwc(offset, off_nhwc, N, 18 * dg, H, W, 18 * dg, stream)) != TDVC_OK) return rc;
  if ((rc = tdvc_nchw_to_nhwc(mask, msk_nhwc, N, 9 * dg, H, W, 9 * dg, stream)) != TDVC_OK) return rc;
  dcn_pack_weight_kernel<<<cdiv((int64_t)C * 9 * o_pad, 256), 256, 0, st>>>(weight, wpk, O, C, o_pad);
  TDVC_CHECK_LAUNCH("dcn_pack_weight");
  TdvcDcnParams p;
  memset(&p, 0, sizeof(p));
  p.input = in_nhwc; p.in_ld = C;
  p.offset = off_nhwc; p.off_ld = 18 * dg;
  p.mask = msk_nhwc; p.mask_ld = 9 * dg; p.mask_is_logit = 0;
  p.weight_packed = wpk; p.bias = bias;
  p.out = out_nhwc; p.out_ld = o_ld;
  p.N = N; p.H = H; p.W = W; p.C = C; p.O = O; p.O_pad = o_pad; p.dg = dg;
  p.round_fp16 = 0; p.act = TDVC_ACT_NONE; p.slope = 0.f; p.impl = 1;  // exact fp32 contraction, like the reference
  if ((rc = tdvc_dcn_nhwc(&p, stream)) != TDVC_OK) return rc;
  return tdvc_nhwc_to_nchw(out_nhwc, o_ld, output, N, O, H, W, stream);
}
