"""Training-mode forward WITH autograd (`net.train()` + gradients enabled): the P-frame forward of reference
main/model/pnet.py:26-83 strung together from the autograd functions of `tdvc_b200.ops`, so that `rd_loss.backward()` of
reference tools/train.py:132-159 works on the module (BASELINE config 4).

What runs where:
* our kernels, forward and backward: every convolution (`ops.conv2d`: tcgen05 forward / dgrad, deterministic fp32 wgrad), GDN /
  IGDN (`ops.gdn`), the deformable convolution (`ops.dcn_v2_conv`), both likelihood-to-bits reductions
  (`ops.entropy_bottleneck_bits`, `ops.gaussian_conditional_bits`), the noise draws (`tdvc_uniform_noise`) and the aux losses;
  >= 99 % of the step's arithmetic (3.83 MMAC/px are convolutions, DCN and GDN);
* torch autograd is the tape, and torch's own element-wise / indexing operators are the GLUE between those functions
  (residual adds, concatenations, pixel shuffles, squeeze-excitation gates, average pooling, bilinear resizing, the SPyNet
  `grid_sample`, the FeatureFix block gather and cosine gate): their backward kernels are not built yet - each is listed in
  DESIGN.md.  The inference path (`model._Plan`) uses none of them.

Tensors are NCHW float32 as in the reference; the functions below take the parameter containers of `tdvc_b200.model`.
"""
import math

import torch
import torch.nn.functional as F

from tdvc_b200 import lib as L
from tdvc_b200 import ops

_LN2 = math.log(2.0)


# ----------------------------------------------------------------------------------------------- small pieces
def _cv(m, x, act=None, slope=0.01):
    """nn.Conv2d container -> ops.conv2d (+ fused activation)."""
    return ops.conv2d(x, m.weight, m.bias, m.stride[0], m.padding[0], act, slope)


def _cv3d(m, x, act=None, slope=0.01):
    """nn.Conv3d with a (1, 3, 3) kernel acts on every frame separately: a 2-D convolution of the frames (pnet.py:296-307)."""
    return ops.conv2d(x, m.weight.squeeze(2), m.bias, 1, 1, act, slope)


def _subpel(seq, x, act=None, slope=0.01):
    """compressai subpel_conv3x3: conv -> PixelShuffle(2); an element-wise activation commutes with the shuffle."""
    return F.pixel_shuffle(_cv(seq[0], x, act, slope), 2)


def _res_stack(stack, x):
    for blk in stack:   # x + conv2(relu(conv1 x))  (reference utils.py:43-56)
        x = x + _cv(blk.conv2, _cv(blk.conv1, x, "relu"))
    return x


def _se(m, x):
    """Squeeze-excitation (reference inflate.py:159-208): spatial mean and gated product on our kernels; the gate MLP acts on
    N x C numbers."""
    w1, w2 = m.conv1.conv.weight, m.conv2.conv.weight
    s = ops.channel_mean(x)                                             # (N, C)
    # (F.linear, not F.conv2d: cuDNN convolutions default to TF32, which would put 1e-3 into every gate)
    s = F.relu(F.linear(s, w1.view(w1.shape[0], -1), m.conv1.conv.bias))
    s = torch.sigmoid(F.linear(s, w2.view(w2.shape[0], -1), m.conv2.conv.bias))
    return ops.channel_scale(x, s)


class _LowerBound(torch.autograd.Function):
    """compressai LowerBound on PARAMETER tensors (GDN beta / gamma): max(x, bound), gradient passed where x >= bound or the
    gradient is negative."""

    @staticmethod
    def forward(ctx, x, bound):
        ctx.save_for_backward(x, bound)
        return torch.max(x, bound)

    @staticmethod
    def backward(ctx, g):
        x, bound = ctx.saved_tensors
        return ((x >= bound) | (g < 0)).type(g.dtype) * g, None


def _reparam(p, mod):
    """compressai NonNegativeParametrizer.forward: max(p, bound)^2 - pedestal."""
    return _LowerBound.apply(p, mod.lower_bound.bound) ** 2 - mod.pedestal


def _gdn(m, x):
    return ops.gdn(x, _reparam(m.beta, m.beta_reparam), _reparam(m.gamma, m.gamma_reparam), m.inverse, max_key=m.gamma)


# ----------------------------------------------------------------------------------------------- coders
def _rb_stride(m, x):
    return _gdn(m.gdn, _cv(m.conv2, _cv(m.conv1, x, "leaky_relu"))) + _cv(m.skip, x)


def _rb(m, x):
    return _cv(m.conv2, _cv(m.conv1, x, "leaky_relu"), "leaky_relu") + x


def _rb_up(m, x):
    return _gdn(m.igdn, _cv(m.conv, _subpel(m.subpel_conv, x, "leaky_relu"))) + _subpel(m.upsample, x)


def coder(cd, x, noise, num_pixels):
    """compressai Cheng2020Anchor.forward in training mode with the reference's g_a / g_s (encoder_v3.py:14-69).
    noise: {"z", "y", "y_lik"} uniform draws.  -> (x_hat, bpp of this coder as pnet.py:38-43 forms it)."""
    ga, gs = cd.g_a, cd.g_s
    a = _rb(ga[1], _rb_stride(ga[0], x))
    a = _rb(ga[4], _se(ga[3], _rb_stride(ga[2], a)))
    a = _rb(ga[6], _rb_stride(ga[5], a))
    y = _se(ga[8], _cv(ga[7], a))
    ha = cd.h_a
    h = _cv(ha[2], _cv(ha[0], y, "leaky_relu"), "leaky_relu")
    z = _cv(ha[8], _cv(ha[6], _cv(ha[4], h, "leaky_relu"), "leaky_relu"))
    eb = cd.entropy_bottleneck
    z_tilde, s_z = ops.entropy_bottleneck_bits(z, noise["z"], [getattr(eb, f"_matrix{i}") for i in range(5)],
                                               [getattr(eb, f"_bias{i}") for i in range(5)],
                                               [getattr(eb, f"_factor{i}") for i in range(4)])
    hs = cd.h_s
    s = _subpel(hs[2], _cv(hs[0], z_tilde, "leaky_relu"), "leaky_relu")
    params = _cv(hs[8], _subpel(hs[6], _cv(hs[4], s, "leaky_relu"), "leaky_relu"))
    y_hat = y + noise["y"]                                  # quantize(y, "noise")
    ctx = cd.context_prediction
    with torch.no_grad():   # compressai MaskedConv2d.forward masks the weight IN PLACE (`self.weight.data *= self.mask`) and then
        ctx.weight.mul_(ctx.mask)   # convolves with the parameter itself: the gradient it accumulates is therefore not masked
    ctx_p = ops.conv2d(y_hat, ctx.weight, ctx.bias, 1, 2)
    ep = cd.entropy_parameters
    gp = _cv(ep[4], _cv(ep[2], _cv(ep[0], torch.cat((params, ctx_p), 1), "leaky_relu"), "leaky_relu"))
    scales, means = gp.chunk(2, 1)
    s_y = ops.gaussian_conditional_bits(y, scales.contiguous(), means.contiguous(), noise["y_lik"])
    g = _rb_up(gs[2], _rb(gs[1], _se(gs[0], y_hat)))
    g = _se(gs[5], _rb_up(gs[4], _rb(gs[3], g)))
    g = _rb(gs[8], _rb_up(gs[7], _rb(gs[6], g)))
    x_hat = _subpel(gs[9], g)
    bpp = (s_y + s_z) / (-_LN2 * num_pixels)
    return x_hat, bpp


# ----------------------------------------------------------------------------------------------- motion estimation
def spynet(sp, ref, supp):
    """reference flownet.py:82-140 (6 levels, coarse to fine)."""
    n, _, h, w = ref.shape
    if h % 32 or w % 32:
        raise RuntimeError("SPyNet: H and W must be multiples of 32")
    refs, supps = [ref], [supp]
    for _ in range(5):
        refs.append(F.avg_pool2d(refs[-1], 2, 2, count_include_pad=False))
        supps.append(F.avg_pool2d(supps[-1], 2, 2, count_include_pad=False))
    refs, supps = refs[::-1], supps[::-1]
    flow = None
    for lvl in range(6):
        # [ref, warp(supp, up), up] with up = 2 * upsample_x2(flow): one kernel forward, two backward (ops.spynet_level_input)
        x8 = ops.spynet_level_input(refs[lvl], supps[lvl], flow)
        t = x8
        bm = sp.basic_module[lvl].basic_module
        for i in range(4):
            t = _cv(bm[i].conv, t, "relu")
        flow = x8[:, 6:8] + _cv(bm[4].conv, t)
    return flow


def motion_est(me, in_f, ref_f, in_img, ref_img):
    """reference pnet.py:131-167."""
    n = in_f.shape[0]
    both = torch.cat([in_f, ref_f], 0)
    l2 = _cv(me.conv_l2_2, _cv(me.conv_l2_1, both, "leaky_relu", 0.1), "leaky_relu", 0.1)
    l3 = _cv(me.conv_l3_2, _cv(me.conv_l3_1, l2, "leaky_relu", 0.1), "leaky_relu", 0.1)
    pyr_in, pyr_ref = [in_f, l2[:n], l3[:n]], [ref_f, l2[n:], l3[n:]]
    up = off = None
    for i in (3, 2, 1):
        lv = f"l{i}"
        o1 = _cv(me.offset_conv11[lv], torch.cat([pyr_in[i - 1], pyr_ref[i - 1]], 1), "leaky_relu", 0.1)
        o1 = _cv(me.offset_conv11_1[lv], o1, "leaky_relu", 0.1)
        if i == 3:
            off = _cv(me.offset_conv12[lv], o1, "leaky_relu", 0.1)
        else:
            off = _cv(me.feat_fusion[lv], torch.cat([up, o1], 1), "leaky_relu", 0.1)
        if i > 1:
            up = _cv(me.upsample_conv, F.interpolate(off, scale_factor=2, mode="bilinear", align_corners=False))
    flow = spynet(me.spynet, in_img, ref_img)
    off = off + flow.repeat(1, off.size(1) // 2, 1, 1)
    return _se(me.attn, _cv(me.feat_fusion_, off))


# ----------------------------------------------------------------------------------------------- motion compensation
def mcnet(mc, offset_feat, ref):
    """reference pnet.py:179-184 with the DCNv2 module of dcn_v2_amp.py:219-234 (output rounded to fp16, :67-69)."""
    d = mc.dconv
    o1, o2, m = torch.chunk(_cv(d.conv_offset_mask, offset_feat), 3, dim=1)
    res = ops.dcn_v2_conv(ref, torch.cat((o1, o2), 1).contiguous(), torch.sigmoid(m).contiguous(), d.weight, d.bias, 1, 1, 1, d.dg)
    out = F.leaky_relu(res.half(), 0.1).float()
    out2 = _cv(mc.conv, torch.cat([out, ref], 1), "leaky_relu", 0.1)
    return out + _res_stack(mc.recon_layer, out2)


# ----------------------------------------------------------------------------------------------- multi-frame fusion
def mcfilter(mf, pred, refer_frames):
    """reference LoopFilter (pnet.py:277-293) + Bottleneck3D (:309-317); frames are batched as (N * T, 64, H, W)."""
    r = refer_frames[:, 1:]
    n, m, _, h, w = r.shape
    r = _cv(mf.conv02, _cv(mf.conv01, r.reshape(n * m, 3, h, w), "leaky_relu", 0.1)).reshape(n, m, 64, h, w)
    T = m + 1
    x = torch.cat((r, pred.unsqueeze(1)), 1).reshape(n * T, 64, h, w)
    x = _cv3d(mf.conv1, x, "leaky_relu", 0.1)
    b3 = mf.layer1
    out = _cv3d(b3.spatial_conv3d, _cv3d(b3.conv1, x, "leaky_relu", 0.1)).reshape(n, T, 64, h, w)
    # temporal (3, 1, 1) kernel, stride 3, no bias: one output step from frames 0..2 = a 1x1 convolution of their channels
    wt = b3.temporal_conv3d.weight.squeeze(-1).squeeze(-1).permute(0, 2, 1).reshape(64, 192, 1, 1)
    tmp = ops.conv2d(out[:, :3].reshape(n, 192, h, w), wt, None, 1, 0, max_key=b3.temporal_conv3d.weight)
    out = F.leaky_relu(out + tmp.unsqueeze(1), 0.1).reshape(n * T, 64, h, w)
    x = (_cv3d(b3.conv3, out) + x).reshape(n, T, 64, h, w).reshape(n, T * 64, h, w)
    x = _se(mf.attn, _cv(mf.feat_fusion, x, "leaky_relu", 0.1))
    return pred + x


# ----------------------------------------------------------------------------------------------- in-loop filter
def _feature_extract(fe, x):
    x1 = _cv(fe.conv_first, x, "leaky_relu", 0.01)
    return _cv(fe.conv_last, _res_stack(fe.body, x1)) + x1


def loopfilter(lf, feat, refer_frames, training):
    """reference FeatureFix (pnet.py:213-263)."""
    n, c, h, w = feat.shape
    f_in = _feature_extract(lf.FeatureExtract_input, feat)
    f_ref = _feature_extract(lf.FeatureExtract_ref, refer_frames[:, 0])
    scale = 8 if training else int(h / 8)
    bs = 3 * scale
    with torch.no_grad():   # the patch match only yields indices (argmax): nothing to differentiate
        q = F.unfold(F.avg_pool2d(f_in, scale, scale), 3, padding=3, stride=3).transpose(2, 1)
        k = F.unfold(F.avg_pool2d(f_ref, scale, scale), 3, padding=3, stride=3).transpose(2, 1).reshape(n, -1, c * 9)
        ind = torch.bmm(F.normalize(q, dim=2), F.normalize(k.transpose(2, 1), dim=1)).max(dim=2, keepdim=True)[1]
    blocks = F.unfold(f_ref, bs, padding=bs, stride=bs).transpose(2, 1).reshape(n, -1, c * bs * bs)
    idx = ind.view(n, 1, -1).expand(-1, c * bs * bs, -1).permute(0, 2, 1)
    picked = torch.gather(blocks, 1, idx).view(n, -1, c, bs, bs).permute(0, 2, 3, 4, 1).reshape(n, -1, q.size(1))
    out = F.fold(picked, (h, w), bs, padding=bs, stride=bs)
    cor = torch.cosine_similarity(f_in, out).unsqueeze(1)
    out = _cv(lf.featfusion, torch.cat([f_in, out], 1) * cor, "leaky_relu", 0.1)
    out = F.leaky_relu(_se(lf.attn, _cv(lf.featfusion2, torch.cat([out, f_ref], 1))), 0.1)
    return _cv(lf.featdown, feat + _res_stack(lf.recon_layer, out)), ind


# ----------------------------------------------------------------------------------------------- the frame
_GRAPH_SEED = {}


def draw_noise(shapes, device):
    """Uniform(-0.5, 0.5) draws for the six noise quantisers from our Philox kernel, seeded from torch's CPU generator.  While
    a CUDA graph is being captured (a whole training step replayed as one graph) a host value would be frozen into the graph:
    the seed then lives in a device word that a captured add advances, so every replay draws fresh noise."""
    lib = L.load()
    capturing = torch.cuda.is_current_stream_capturing()
    if capturing:
        seed_dev = _GRAPH_SEED.get(device)
        if seed_dev is None:
            raise RuntimeError("run one eager training step on this device before capturing it in a CUDA graph")
        seed_dev.add_(0x9E3779B97F4A7C15 - (1 << 64))    # a captured operation: advances the seed at every replay
    else:
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        if device not in _GRAPH_SEED:
            _GRAPH_SEED[device] = torch.tensor([seed ^ 0x5DEECE66D], dtype=torch.int64, device=device)
    out = {}
    with torch.cuda.device(device):
        st = torch.cuda.current_stream(device).cuda_stream
        for i, (k, shp) in enumerate(shapes.items()):
            t = torch.empty(shp, device=device, dtype=torch.float32)
            if capturing:
                L.check(lib.tdvc_uniform_noise_dev(t.data_ptr(), t.numel(), seed_dev.data_ptr(), i, st), "uniform_noise_dev")
            else:
                L.check(lib.tdvc_uniform_noise(t.data_ptr(), t.numel(), seed, i, st), "uniform_noise")
            out[k] = t
    return out


def forward(net, input_image, refer_frames, noise=None):
    """`VideoCompressor.forward` in training mode with an autograd graph.  -> (recon, bpp_res (1,), bpp_mv (1,), ind)."""
    N, _, H, W = input_image.shape
    npx = N * H * W
    if noise is None:
        shapes = {}
        for nm in ("mv", "res"):
            shapes[f"{nm}.z"] = (N, 128, H // 64, W // 64)
            shapes[f"{nm}.y"] = shapes[f"{nm}.y_lik"] = (N, 128, H // 16, W // 16)
        noise = draw_noise(shapes, input_image.device)
    pick = lambda nm: {k: noise[f"{nm}.{k}"].detach().float().contiguous() for k in ("z", "y", "y_lik")}
    refer = refer_frames[:, -1].contiguous()
    fe = net.extra_fea
    in_f = _res_stack(fe.residual_layer, _cv(fe.conv_first, input_image, "leaky_relu", 0.1))
    ref_f = _res_stack(fe.residual_layer, _cv(fe.conv_first, refer, "leaky_relu", 0.1))
    estmv = motion_est(net.motion_est, in_f, ref_f, input_image, refer)
    mv_hat, bpp_mv = coder(net.mvCoder, estmv, pick("mv"), npx)
    pred1 = mcnet(net.mcnet, mv_hat, ref_f)
    pred = mcfilter(net.mcfilter, pred1, refer_frames)
    res_hat, bpp_res = coder(net.resCoder, in_f - pred, pick("res"), npx)
    recon, ind = loopfilter(net.loopfilter, pred + res_hat, refer_frames, True)
    return recon.clamp(0.0, 1.0), bpp_res.view(-1), bpp_mv.view(-1), ind
