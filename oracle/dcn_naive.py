"""ORACLE (test infrastructure): CPU restatements of the reference's only in-tree native op,
DCNv2 forward (reference main/utils/dcnv2/src/cuda/dcn_v2_cuda.cu:20-95 and
src/cuda/dcn_v2_im2col_cuda.cu:25-54,125-195).

Three implementations of the same arithmetic, cross-checked in tests/test_oracle.py:
  * dcn_v2_forward_naive       - element-by-element Python loops (tiny cases only)
  * dcn_v2_forward_vectorised  - the same rule, vectorised with torch index arithmetic
  * dcn_v2_forward             - torchvision.ops.deform_conv2d (same MXNet lineage; the fast path the
                                 oracle model and the `_ext` shim use)
Pinned by the reference's known-answer test `check_zero_offset`
(reference main/utils/dcnv2/testcuda.py:36-71): zero offsets, mask 0.5, identity-centre weights
=> 2*dcn(x) == x.

Layouts (contiguous NCHW fp32): offset (N, 2*DG*kh*kw, H, W) ordered [g][tap][dy,dx];
mask (N, DG*kh*kw, H, W) ordered [g][tap].
"""
import math

import torch


def _bilinear(img, H, W, h, w):
    """dmcn_im2col_bilinear_cuda (dcn_v2_im2col_cuda.cu:25-54): each corner zeroed individually."""
    h_low, w_low = math.floor(h), math.floor(w)
    h_high, w_high = h_low + 1, w_low + 1
    lh, lw = h - h_low, w - w_low
    hh, hw = 1 - lh, 1 - lw
    v1 = img[h_low][w_low] if (h_low >= 0 and w_low >= 0) else 0.0
    v2 = img[h_low][w_high] if (h_low >= 0 and w_high <= W - 1) else 0.0
    v3 = img[h_high][w_low] if (h_high <= H - 1 and w_low >= 0) else 0.0
    v4 = img[h_high][w_high] if (h_high <= H - 1 and w_high <= W - 1) else 0.0
    return hh * hw * v1 + hh * lw * v2 + lh * hw * v3 + lh * lw * v4


def dcn_v2_forward_naive(inp, weight, bias, offset, mask, dg, pad=1):
    """3x3, stride 1, dilation 1.  float64 accumulation of the exact reference rule."""
    N, C, H, W = inp.shape
    O = weight.shape[0]
    cpg = C // dg
    x = inp.double().tolist()
    off = offset.double().tolist()
    msk = mask.double().tolist()
    wgt = weight.double().tolist()
    out = torch.zeros(N, O, H, W, dtype=torch.float64)
    for n in range(N):
        for ho in range(H):
            for wo in range(W):
                col = [0.0] * (C * 9)
                for c in range(C):
                    g = c // cpg
                    for i in range(3):
                        for j in range(3):
                            k = i * 3 + j
                            oh = off[n][g * 18 + 2 * k][ho][wo]
                            ow = off[n][g * 18 + 2 * k + 1][ho][wo]
                            m = msk[n][g * 9 + k][ho][wo]
                            h_im = ho - pad + i + oh
                            w_im = wo - pad + j + ow
                            v = 0.0
                            if h_im > -1 and w_im > -1 and h_im < H and w_im < W:
                                v = _bilinear(x[n][c], H, W, h_im, w_im)
                            col[c * 9 + k] = v * m
                for o in range(O):
                    acc = float(bias[o])
                    wo_ = wgt[o]
                    for c in range(C):
                        for k in range(9):
                            acc += wo_[c][k // 3][k % 3] * col[c * 9 + k]
                    out[n, o, ho, wo] = acc
    return out.float()


def dcn_v2_forward_vectorised(inp, weight, bias, offset, mask, dg, pad=1):
    """Same rule in torch ops: build the modulated columns (N, C*9, H*W), then W_flat @ columns + bias
    (dcn_v2_cuda.cu:73-92)."""
    N, C, H, W = inp.shape
    O = weight.shape[0]
    cpg = C // dg
    dev = inp.device
    ys = torch.arange(H, device=dev, dtype=inp.dtype).view(1, 1, H, 1)
    xs = torch.arange(W, device=dev, dtype=inp.dtype).view(1, 1, 1, W)
    flat = inp.reshape(N, dg, cpg, H * W)
    cols = inp.new_zeros(N, dg, cpg, 9, H, W)
    for k in range(9):
        i, j = divmod(k, 3)
        oh = offset[:, 2 * k::18][:, :dg].reshape(N, dg, H, W)  # channel g*18 + 2k
        ow = offset[:, 2 * k + 1::18][:, :dg].reshape(N, dg, H, W)
        m = mask[:, k::9][:, :dg].reshape(N, dg, H, W)
        h_im = ys - pad + i + oh
        w_im = xs - pad + j + ow
        inside = (h_im > -1) & (w_im > -1) & (h_im < H) & (w_im < W)
        h_low, w_low = torch.floor(h_im), torch.floor(w_im)
        lh, lw = h_im - h_low, w_im - w_low
        hh, hw = 1 - lh, 1 - lw
        h_low, w_low = h_low.long(), w_low.long()
        h_high, w_high = h_low + 1, w_low + 1

        def corner(hi, wi, ok):
            idx = (hi.clamp(0, H - 1) * W + wi.clamp(0, W - 1)).view(N, dg, 1, H * W).expand(-1, -1, cpg, -1)
            v = torch.gather(flat, 3, idx).view(N, dg, cpg, H, W)
            return v * (ok & inside).unsqueeze(2).to(inp.dtype)

        v1 = corner(h_low, w_low, (h_low >= 0) & (w_low >= 0))
        v2 = corner(h_low, w_high, (h_low >= 0) & (w_high <= W - 1))
        v3 = corner(h_high, w_low, (h_high <= H - 1) & (w_low >= 0))
        v4 = corner(h_high, w_high, (h_high <= H - 1) & (w_high <= W - 1))
        val = ((hh * hw).unsqueeze(2) * v1 + (hh * lw).unsqueeze(2) * v2
               + (lh * hw).unsqueeze(2) * v3 + (lh * lw).unsqueeze(2) * v4)
        cols[:, :, :, k] = val * m.unsqueeze(2)
    cols = cols.reshape(N, C * 9, H * W)
    out = torch.matmul(weight.reshape(O, C * 9), cols).view(N, O, H, W)
    return out + bias.view(1, O, 1, 1)


def dcn_v2_forward(inp, weight, bias, offset, mask, dg, pad=1):
    """Fast path: torchvision.ops.deform_conv2d (identical rule; cross-checked in tests)."""
    import torchvision.ops
    assert offset.shape[1] == 2 * dg * 9 and mask.shape[1] == dg * 9
    return torchvision.ops.deform_conv2d(inp, offset, weight, bias, stride=(1, 1), padding=(pad, pad),
                                         dilation=(1, 1), mask=mask)
