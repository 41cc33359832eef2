"""ORACLE (test infrastructure, not product code): plain-PyTorch fp32 CPU restatement of TDVC's P-frame
coding forward pass, `VideoCompressor.forward` (reference main/model/pnet.py:26-83).

Every function cites the reference lines it follows.  The module tree reproduces the reference's
state_dict key set exactly (strict load both ways is tested in tests/test_oracle.py), and the result is
pinned against the reference's own unmodified code imported through oracle/ref_import.py
(tests/test_oracle.py, fixtures in tests/golden/ made by oracle/make_golden.py).

Parity status: the pnet/flownet/inflate/utils/dcn parts are PINNED to the reference source (bit-exact on
CPU against the verbatim import).  The coders' arithmetic lives in third-party CompressAI, absent from
/root/reference => that part is "parity unpinned" (oracle/compressai_port.py header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
anything under oracle/.  `forward(..., taps=dict)` additionally records per-stage tensors (SURVEY.md
App. D dump points) for stage-by-stage comparison with the CUDA path.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from oracle.compressai_port import (Cheng2020Anchor, ResidualBlock, ResidualBlockUpsample,
                                    ResidualBlockWithStride, conv3x3, subpel_conv3x3)
from oracle.dcn_naive import dcn_v2_forward


def _lrelu(x, slope):
    return F.leaky_relu(x, negative_slope=slope)


# --------------------------------------------------------------------------- small blocks
class ConvAct(nn.Module):
    """mmcv ConvModule stand-in: keys `<name>.conv.weight|bias` (reference flownet.py:187-227, inflate.py:189-202)."""

    def __init__(self, i, o, k, pad, act):
        super().__init__()
        self.conv = nn.Conv2d(i, o, k, 1, pad)
        self.act = act

    def forward(self, x):
        x = self.conv(x)
        if self.act == "relu":
            x = F.relu(x)
        elif self.act == "sigmoid":
            x = torch.sigmoid(x)
        return x


class SELayer(nn.Module):
    """Squeeze-and-excitation, ratio 16 (reference main/model/inflate.py:159-208)."""

    def __init__(self, channels, ratio=16):
        super().__init__()
        self.conv1 = ConvAct(channels, int(channels / ratio), 1, 0, "relu")
        self.conv2 = ConvAct(int(channels / ratio), channels, 1, 0, "sigmoid")

    def forward(self, x):
        s = F.adaptive_avg_pool2d(x, 1)
        return x * self.conv2(self.conv1(s))


class ResBlock(nn.Module):
    """x + conv2(relu(conv1(x)))  (reference main/utils/utils.py:43-56)."""

    def __init__(self, c=64):
        super().__init__()
        self.conv1 = nn.Conv2d(c, c, 3, 1, 1)
        self.conv2 = nn.Conv2d(c, c, 3, 1, 1)

    def forward(self, x):
        return x + self.conv2(F.relu(self.conv1(x)))


def _res_stack(n, c=64):
    return nn.Sequential(*[ResBlock(c) for _ in range(n)])


class FeaExtra(nn.Module):
    """reference pnet.py:86-96."""

    def __init__(self, num_block):
        super().__init__()
        self.conv_first = nn.Conv2d(3, 64, 3, 1, 1)
        self.residual_layer = _res_stack(num_block)

    def forward(self, x):
        return self.residual_layer(_lrelu(self.conv_first(x), 0.1))


# --------------------------------------------------------------------------- SPyNet
def flow_warp_border(x, flow_nhw2):
    """Bilinear backward warp, border padding, align_corners=True (reference flownet.py:8-48, 134-137).
    Keeps the reference's normalise -> grid_sample un-normalise coordinate round trip."""
    _, _, h, w = x.shape
    gy, gx = torch.meshgrid(torch.arange(0, h), torch.arange(0, w), indexing="ij")
    grid = torch.stack((gx, gy), 2).type_as(x)
    gf = grid + flow_nhw2
    nx = 2.0 * gf[..., 0] / max(w - 1, 1) - 1.0
    ny = 2.0 * gf[..., 1] / max(h - 1, 1) - 1.0
    return F.grid_sample(x, torch.stack((nx, ny), dim=3), mode="bilinear", padding_mode="border",
                         align_corners=True)


class SPyNetLevel(nn.Module):
    """5 x conv7x7: 8->32->64->32->16->2, ReLU between (reference flownet.py:178-238)."""

    def __init__(self):
        super().__init__()
        chans = [(8, 32), (32, 64), (64, 32), (32, 16), (16, 2)]
        self.basic_module = nn.Sequential(*[ConvAct(i, o, 7, 3, "relu" if n < 4 else None)
                                            for n, (i, o) in enumerate(chans)])

    def forward(self, x):
        return self.basic_module(x)


class SPyNet(nn.Module):
    """6-level coarse-to-fine flow (reference flownet.py:51-175); `mean`/`std` buffers exist but are unused."""

    def __init__(self):
        super().__init__()
        self.basic_module = nn.ModuleList([SPyNetLevel() for _ in range(6)])
        self.register_buffer("mean", torch.Tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1))
        self.register_buffer("std", torch.Tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1))

    def forward(self, ref, supp, taps=None):
        n, _, h, w = ref.shape
        # reference flownet.py:153-173 resizes to a multiple of 32 and back; identity for our shapes
        assert h % 32 == 0 and w % 32 == 0, "SPyNet resize path not restated (SURVEY App. C.15)"
        refs, supps = [ref], [supp]
        for _ in range(5):  # flownet.py:101-114
            refs.append(F.avg_pool2d(refs[-1], 2, 2, count_include_pad=False))
            supps.append(F.avg_pool2d(supps[-1], 2, 2, count_include_pad=False))
        refs, supps = refs[::-1], supps[::-1]
        flow = ref.new_zeros(n, 2, h // 32, w // 32)
        for lvl in range(6):  # flownet.py:118-138
            if lvl == 0:
                up = flow
            else:
                up = F.interpolate(flow, scale_factor=2, mode="bilinear", align_corners=True) * 2.0
            warped = flow_warp_border(supps[lvl], up.permute(0, 2, 3, 1))
            flow = up + self.basic_module[lvl](torch.cat([refs[lvl], warped, up], 1))
            if taps is not None:
                taps[f"spynet.flow{lvl}"] = flow
        return flow


# --------------------------------------------------------------------------- motion estimation
class OffsetGen(nn.Module):
    """3-level offset pyramid + SPyNet flow (reference pnet.py:99-167)."""

    def __init__(self, nf=64):
        super().__init__()
        self.offset_conv11 = nn.ModuleDict()
        self.offset_conv11_1 = nn.ModuleDict()
        self.offset_conv12 = nn.ModuleDict()
        self.feat_fusion = nn.ModuleDict()
        for i in (3, 2, 1):
            lv = f"l{i}"
            self.offset_conv11[lv] = nn.Conv2d(2 * nf, nf, 3, 1, 1)
            self.offset_conv11_1[lv] = nn.Conv2d(nf, nf, 3, 1, 1)
            self.offset_conv12[lv] = nn.Conv2d(nf, nf, 3, 1, 1)  # l1,l2 allocated but unused (pnet.py:112 vs 152-156)
            if i < 3:
                self.feat_fusion[lv] = nn.Conv2d(2 * nf, nf, 1, 1, 0)
        self.upsample_conv = nn.Conv2d(nf, nf, 3, 1, 1)
        self.conv_l2_1 = nn.Conv2d(nf, nf, 3, 2, 1)
        self.conv_l2_2 = nn.Conv2d(nf, nf, 3, 1, 1)
        self.conv_l3_1 = nn.Conv2d(nf, nf, 3, 2, 1)
        self.conv_l3_2 = nn.Conv2d(nf, nf, 3, 1, 1)
        self.spynet = SPyNet()
        self.attn = SELayer(64)
        self.feat_fusion_ = nn.Conv2d(nf, nf, 3, 1, 1)

    def forward(self, in_f, ref_f, in_img, ref_img, taps=None):
        both = torch.cat([in_f, ref_f], 0)  # batch the two images for the pyramid convs (pnet.py:132-140)
        l2 = _lrelu(self.conv_l2_2(_lrelu(self.conv_l2_1(both), 0.1)), 0.1)
        l3 = _lrelu(self.conv_l3_2(_lrelu(self.conv_l3_1(l2), 0.1)), 0.1)
        n = in_f.shape[0]
        pyr_in = [in_f, l2[:n], l3[:n]]
        pyr_ref = [ref_f, l2[n:], l3[n:]]
        up = None
        for i in (3, 2, 1):  # pnet.py:146-160
            lv = f"l{i}"
            o1 = torch.cat([pyr_in[i - 1], pyr_ref[i - 1]], 1)
            o1 = _lrelu(self.offset_conv11[lv](o1), 0.1)
            o1 = _lrelu(self.offset_conv11_1[lv](o1), 0.1)
            if i == 3:
                off = _lrelu(self.offset_conv12[lv](o1), 0.1)
            else:
                off = _lrelu(self.feat_fusion[lv](torch.cat([up, o1], 1)), 0.1)
            if i > 1:
                up = self.upsample_conv(F.interpolate(off, scale_factor=2, mode="bilinear", align_corners=False))
            if taps is not None:
                taps[f"motion_est.offset_{lv}"] = off
        flow = self.spynet(in_img, ref_img, taps)  # pnet.py:162
        off = off + flow.repeat(1, off.size(1) // 2, 1, 1)  # even ch += u, odd ch += v (pnet.py:163)
        return self.attn(self.feat_fusion_(off))


# --------------------------------------------------------------------------- motion compensation
class DCN(nn.Module):
    """DCNv2 module with its offset/mask head (reference main/utils/dcnv2/dcn_v2_amp.py:125-234).
    Output is rounded to fp16, as `_DCNv2.forward` does with its module-level use_amp=True
    (dcn_v2_amp.py:15,67-69)."""

    def __init__(self, cin, cout, dg):
        super().__init__()
        self.dg = dg
        self.weight = nn.Parameter(torch.Tensor(cout, cin, 3, 3))
        self.bias = nn.Parameter(torch.Tensor(cout))
        stdv = 1.0 / math.sqrt(cin * 9)
        self.weight.data.uniform_(-stdv, stdv)
        self.bias.data.zero_()
        self.conv_offset_mask = nn.Conv2d(cin, dg * 27, 3, 1, 1)
        self.conv_offset_mask.weight.data.zero_()
        self.conv_offset_mask.bias.data.zero_()

    def forward(self, x, y, taps=None):
        out = self.conv_offset_mask(y)
        o1, o2, m = torch.chunk(out, 3, dim=1)
        offset = torch.cat((o1, o2), dim=1)
        mask = torch.sigmoid(m)
        res = dcn_v2_forward(x.float(), self.weight.float(), self.bias.float(), offset.float(),
                                        mask.float(), self.dg)
        if taps is not None:
            taps["mcnet.dcn_offset"], taps["mcnet.dcn_mask"], taps["mcnet.dcn_raw_fp32"] = offset, mask, res
        return res.half()


class MCNet(nn.Module):
    """reference pnet.py:170-184."""

    def __init__(self, num_block):
        super().__init__()
        self.dconv = DCN(64, 64, 8)
        self.recon_layer = _res_stack(num_block)
        self.feat_down = nn.Conv2d(64, 3, 3, 1, 1)  # allocated, never executed (pnet.py:176)
        self.conv = nn.Conv2d(128, 64, 3, 1, 1)

    def forward(self, offset_feat, ref, taps=None):
        out = _lrelu(self.dconv(ref, offset_feat, taps), 0.1)  # LeakyReLU on the fp16 tensor
        out2 = _lrelu(self.conv(torch.cat([out, ref], 1)), 0.1)  # cat promotes back to fp32
        out2 = self.recon_layer(out2)
        return out + out2


# --------------------------------------------------------------------------- multi-frame fusion
class Bottleneck3D(nn.Module):
    """reference pnet.py:296-317."""

    def __init__(self):
        super().__init__()
        self.conv1 = nn.Conv3d(64, 64, (1, 3, 3), padding=(0, 1, 1))
        self.spatial_conv3d = nn.Conv3d(64, 64, (1, 3, 3), padding=(0, 1, 1))
        self.temporal_conv3d = nn.Conv3d(64, 64, (3, 1, 1), stride=(3, 1, 1), bias=False)
        self.conv3 = nn.Conv3d(64, 64, (1, 3, 3), padding=(0, 1, 1))

    def forward(self, x):
        out = self.spatial_conv3d(_lrelu(self.conv1(x), 0.1))
        out = out + self.temporal_conv3d(out)  # T=4 -> one slice from t=0..2, broadcast over all 4
        out = self.conv3(_lrelu(out, 0.1))
        return out + x


class LoopFilter(nn.Module):
    """`self.mcfilter` of the reference: multi-frame feature fusion (pnet.py:266-293)."""

    def __init__(self):
        super().__init__()
        self.conv01 = nn.Conv2d(3, 64, 3, 1, 1)
        self.conv02 = nn.Conv2d(64, 64, 3, 1, 1)
        self.conv1 = nn.Conv3d(64, 64, (1, 3, 3), padding=(0, 1, 1))
        self.layer1 = Bottleneck3D()
        self.attn = SELayer(64)
        self.feat_fusion = nn.Conv2d(256, 64, 1, 1)

    def forward(self, pred, refer_frames):
        r = refer_frames[:, 1:]
        n, m, _, h, w = r.shape
        r = self.conv02(_lrelu(self.conv01(r.reshape(n * m, 3, h, w)), 0.1)).view(n, m, 64, h, w)
        x = torch.cat((r, pred.unsqueeze(1)), 1).permute(0, 2, 1, 3, 4)  # (n,64,T=4,h,w)
        x = self.layer1(_lrelu(self.conv1(x), 0.1))
        x = x.permute(0, 2, 1, 3, 4).reshape(n, -1, h, w)  # channel = t*64 + c
        x = self.attn(_lrelu(self.feat_fusion(x), 0.1))
        return pred + x


# --------------------------------------------------------------------------- in-loop filter
class FeatureExtract(nn.Module):
    """reference pnet.py:320-332 (note F.leaky_relu default slope 0.01)."""

    def __init__(self, cin, mid, nb):
        super().__init__()
        self.conv_first = nn.Conv2d(cin, mid, 3, 1, 1)
        self.body = _res_stack(nb, mid)
        self.conv_last = nn.Conv2d(mid, mid, 3, 1, 1)

    def forward(self, x):
        x1 = _lrelu(self.conv_first(x), 0.01)
        return self.conv_last(self.body(x1)) + x1


class FeatureFix(nn.Module):
    """`self.loopfilter` of the reference: reference-based in-loop filter with patch matching
    (pnet.py:187-263)."""

    def __init__(self):
        super().__init__()
        self.FeatureExtract_input = FeatureExtract(64, 64, 2)
        self.FeatureExtract_ref = FeatureExtract(3, 64, 2)
        self.recon_layer = _res_stack(2)
        for nme, s in (("conv_10", 2), ("conv_11", 1), ("conv_12", 2), ("conv_13", 1)):  # never executed
            setattr(self, nme, nn.Conv2d(64, 64, 3, s, 1))
        self.featfusion = nn.Conv2d(128, 64, 3, 1, 1)
        self.featfusion2 = nn.Conv2d(128, 64, 3, 1, 1)
        self.featdown = nn.Conv2d(64, 3, 3, 1, 1)
        self.attn = SELayer(64)

    def forward(self, feat, refer_frames, taps=None):
        n, c, h, w = feat.shape
        iframe = refer_frames[:, 0]
        f_in = self.FeatureExtract_input(feat)
        f_ref = self.FeatureExtract_ref(iframe)
        scale = 8 if self.training else int(h / 8)  # pnet.py:220-223
        p_in = F.avg_pool2d(f_in, scale, scale)
        p_ref = F.avg_pool2d(f_ref, scale, scale)
        # 3x3 patches of the pooled maps, pad 3 / stride 3 (pnet.py:230-233)
        q = F.unfold(p_in, 3, padding=3, stride=3).transpose(2, 1)
        k = F.unfold(p_ref, 3, padding=3, stride=3).transpose(2, 1).reshape(n, -1, c * 9)
        sim = torch.bmm(F.normalize(q, dim=2), F.normalize(k.transpose(2, 1), dim=1))
        _, ind = sim.max(dim=2, keepdim=True)  # pnet.py:235-238
        # full-resolution blocks of 3*scale, moved by index; fold with stride==kernel is a placement
        bs = 3 * scale
        blocks = F.unfold(f_ref, bs, padding=bs, stride=bs).transpose(2, 1).reshape(n, -1, c * bs * bs)
        idx = ind.view(n, 1, -1).expand(-1, c * bs * bs, -1).permute(0, 2, 1)
        picked = torch.gather(blocks, 1, idx).view(n, -1, c, bs, bs).permute(0, 2, 3, 4, 1).reshape(n, -1, q.size(1))
        out = F.fold(picked, (h, w), bs, padding=bs, stride=bs) / 1.0  # div = (3/3)^2 (pnet.py:207-210)
        cor = torch.cosine_similarity(f_in, out).unsqueeze(1)  # pnet.py:255
        if taps is not None:
            taps.update({"loopfilter.f_in": f_in, "loopfilter.f_ref": f_ref, "loopfilter.pool_in": p_in,
                         "loopfilter.pool_ref": p_ref, "loopfilter.sim": sim, "loopfilter.ind": ind,
                         "loopfilter.gathered": out, "loopfilter.cor": cor})
        out = _lrelu(self.featfusion(torch.cat([f_in, out], 1) * cor), 0.1)
        out = _lrelu(self.attn(self.featfusion2(torch.cat([out, f_ref], 1))), 0.1)
        out = self.recon_layer(out)
        return self.featdown(feat + out)


# --------------------------------------------------------------------------- coders
def _coder_transforms(N):
    """g_a / g_s of MVCoder and ResCoder (identical; reference encoder_v3.py:17-40, 46-69)."""
    g_a = nn.Sequential(
        ResidualBlockWithStride(64, N, stride=2), ResidualBlock(N, N),
        ResidualBlockWithStride(N, N, stride=2), SELayer(N), ResidualBlock(N, N),
        ResidualBlockWithStride(N, N, stride=2), ResidualBlock(N, N),
        conv3x3(N, N, stride=2), SELayer(N))
    g_s = nn.Sequential(
        SELayer(N), ResidualBlock(N, N), ResidualBlockUpsample(N, N, 2),
        ResidualBlock(N, N), ResidualBlockUpsample(N, N, 2), SELayer(N),
        ResidualBlock(N, N), ResidualBlockUpsample(N, N, 2), ResidualBlock(N, N),
        subpel_conv3x3(N, 64, 2))
    return g_a, g_s


class Coder(Cheng2020Anchor):
    def __init__(self, N=128):
        super().__init__(N=N)
        self.g_a, self.g_s = _coder_transforms(N)


def _bpp(likelihoods, num_pixels):
    """reference pnet.py:38-43."""
    return sum(torch.log(l).sum() / (-math.log(2) * num_pixels) for l in likelihoods.values())


# --------------------------------------------------------------------------- top level
class VideoCompressor(nn.Module):
    """reference pnet.py:15-83.  Same constructor, forward signature, return tuple and state_dict keys."""

    def __init__(self):
        super().__init__()
        self.mvCoder = Coder(128)
        self.resCoder = Coder(128)
        self.extra_fea = FeaExtra(2)
        self.motion_est = OffsetGen()
        self.mcnet = MCNet(3)
        self.loopfilter = FeatureFix()
        self.mcfilter = LoopFilter()

    def forward(self, input_image, refer_frames, enabled_amp=False, is_compress=False, taps=None):
        self.last_coded = {}
        ref = refer_frames[:, -1]
        in_f = self.extra_fea(input_image)  # pnet.py:29-30
        ref_f = self.extra_fea(ref)
        estmv = self.motion_est(in_f, ref_f, input_image, ref, taps)  # :31
        mv = self.mvCoder.forward(estmv.float())  # :34
        mv_aux = self.mvCoder.aux_loss()
        n, _, h, w = input_image.shape
        npx = n * h * w
        bpp_mv = _bpp(mv["likelihoods"], npx)  # :38-43
        if is_compress:  # :45-49 (the reference computes ac_bpp_mv and drops it; kept here for the tests)
            self._code("mv", self.mvCoder, estmv.float(), npx, taps)
        pred1 = self.mcnet(mv["x_hat"], ref_f, taps)  # :52
        pred = self.mcfilter(pred1, refer_frames)  # :53
        resid = in_f - pred  # :55
        rs = self.resCoder.forward(resid.float())  # :58
        rs_aux = self.resCoder.aux_loss()
        bpp_res = _bpp(rs["likelihoods"], npx)  # :62-67
        if is_compress:  # :69-73
            self._code("res", self.resCoder, resid.float(), npx, taps)
        rec_f = pred + rs["x_hat"]  # :76
        recon = self.loopfilter(rec_f, refer_frames, taps).clamp(0.0, 1.0)  # :77-78
        if taps is not None:
            taps.update({"input_feat": in_f, "ref_feat": ref_f, "estmv": estmv, "prediction1": pred1,
                         "prediction": pred, "input_residual": resid, "recon_feat": rec_f,
                         "mv.x_hat": mv["x_hat"], "res.x_hat": rs["x_hat"]})
            for nme, d in (("mv", mv), ("res", rs)):
                for k, v in d["_taps"].items():
                    taps[f"{nme}.{k}"] = v
                taps[f"{nme}.lik_y"] = d["likelihoods"]["y"]
                taps[f"{nme}.lik_z"] = d["likelihoods"]["z"]
        if self.training:
            return recon, bpp_res.view(-1), bpp_mv.view(-1), mv_aux, rs_aux
        return recon, bpp_res.view(-1), bpp_mv.view(-1)

    def _code(self, name, coder, x, npx, taps):
        """reference pnet.py:45-49 / 69-73: eval(), update(force=True), compress(); coded bits of the FIRST batch item
        of each string list over N*H*W pixels, exactly as the reference writes it (`len(s[0])`)."""
        coder.eval()
        coder.update(force=True)
        t = {} if taps is not None else None
        out_enc = coder.compress(x, taps=t)
        ac_bpp = sum(len(s[0]) for s in out_enc["strings"]) * 8.0 / npx
        self.last_coded[name] = {"strings": out_enc["strings"], "shape": tuple(out_enc["shape"]), "ac_bpp": ac_bpp}
        if taps is not None:
            taps.update({f"{name}.ac.{k}": v for k, v in t.items()})
