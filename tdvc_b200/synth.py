"""Synthetic workload for parity tests and bench.py: temporally coherent frames and a conditioned
state_dict (SURVEY.md 8d "Synthetic inputs").

There is no network for datasets or checkpoints, and the reference's default random init is degenerate
(all latent symbols round to 0, DCN offsets are exactly 0), so:
  * frames   = smooth random texture + moving rectangles, each frame = previous one translated by a
               sub-pixel global motion plus 1 % noise (UVG-shaped: values in [0,1], (T,3,H,W));
  * weights  = module default init under seed 1111 (the reference's seed, reference tools/train.py:253-256)
               followed by `condition_state_dict`, a deterministic key-based rescaling that makes the
               latents span several quantisation bins, the DCN offsets a few pixels, SPyNet flow ~1 px,
               and the predicted scales cross the 0.11 floor.
Pure torch, no dependence on oracle/ or on the CUDA library.
"""
import math

import torch
import torch.nn.functional as F

SEED = 1111


def make_gop(h, w, gop=12, seed=0, device="cpu"):
    """Returns (gop, 3, h, w) float32 in [0,1]. Frame 0 plays the I-frame (taken raw: no BPG here)."""
    g = torch.Generator(device="cpu").manual_seed(10007 * (seed + 1))
    m = 32  # canvas margin for global motion
    H, W = h + 2 * m, w + 2 * m

    def field(div, amp):
        lo = torch.rand(1, 3, max(H // div, 2), max(W // div, 2), generator=g)
        return amp * F.interpolate(lo, size=(H, W), mode="bicubic", align_corners=False)

    canvas = (field(64, 0.55) + field(16, 0.3) + field(4, 0.12) + 0.03 * torch.rand(1, 3, H, W, generator=g))
    canvas = canvas.clamp(0, 1)
    nrect = 6
    rc = torch.rand(nrect, 8, generator=g)  # cx, cy, hw, hh, vx, vy, -, -
    col = torch.rand(nrect, 3, generator=g)
    vel = (torch.rand(2, generator=g) - 0.5) * 2 * 3.0  # global motion px/frame, |v| <= 3
    ys = torch.arange(h, dtype=torch.float32).view(h, 1)
    xs = torch.arange(w, dtype=torch.float32).view(1, w)
    frames = []
    for t in range(gop):
        ox, oy = m + vel[0] * t, m + vel[1] * t
        gx = (xs + ox) / (W - 1) * 2 - 1
        gy = (ys + oy) / (H - 1) * 2 - 1
        grid = torch.stack((gx.expand(h, w), gy.expand(h, w)), -1).unsqueeze(0)
        fr = F.grid_sample(canvas, grid, mode="bilinear", padding_mode="border", align_corners=True)[0]
        for r in range(nrect):
            cx = (rc[r, 0] * w + (rc[r, 4] - 0.5) * 16 * t) % w
            cy = (rc[r, 1] * h + (rc[r, 5] - 0.5) * 16 * t) % h
            hw_, hh_ = 4 + rc[r, 2] * w / 10, 4 + rc[r, 3] * h / 10
            a = (torch.sigmoid((hw_ - (xs - cx).abs()) * 1.5) * torch.sigmoid((hh_ - (ys - cy).abs()) * 1.5))
            fr = fr * (1 - a) + col[r].view(3, 1, 1) * a
        fr = fr + 0.01 * torch.randn(3, h, w, generator=g)
        frames.append(fr.clamp(0, 1))
    return torch.stack(frames).to(device)


def make_frame_pair(h, w, seed=0, device="cpu"):
    """(input (1,3,h,w), refer_frames (1,4,3,h,w)) — refs = [I, x(t-3), x(t-2), x(t-1)] raw frames,
    like the training loader (reference main/dataloader/dataset.py:225-245)."""
    f = make_gop(h, w, gop=5, seed=seed)
    x = f[4:5]
    refs = torch.stack([f[0], f[1], f[2], f[3]]).unsqueeze(0)
    return x.to(device), refs.to(device)


def condition_state_dict(sd, seed=SEED):
    """In-place, deterministic, key-based conditioning of a VideoCompressor state_dict (works for the
    oracle, the reference and the CUDA module alike — same keys).  Returns sd."""
    g = torch.Generator(device="cpu").manual_seed(seed + 7)

    def randn(like, std):
        return (torch.randn(like.shape, generator=g) * std).to(like.dtype)

    def rand(like):
        return torch.rand(like.shape, generator=g).to(like.dtype)

    # generic: default conv init shrinks activations layer by layer; give the backbone unit-ish gain
    for k, v in sd.items():
        if k.endswith(".weight") and v.dim() >= 4 and "spynet" not in k and "conv_offset_mask" not in k \
                and "Coder." not in k and not _is_se_key(k):
            v.mul_(GAIN_BACKBONE)
    # (i) DCN offsets ~ a few px, masks spread around 0.5
    k = "mcnet.dconv.conv_offset_mask."
    sd[k + "weight"].copy_(randn(sd[k + "weight"], GAIN_DCN_OFFSET_W))
    sd[k + "bias"].copy_(randn(sd[k + "bias"], 0.5))
    # (v) SPyNet: make each level contribute ~1 px of flow
    for lvl in range(6):
        p = f"motion_est.spynet.basic_module.{lvl}.basic_module."
        for i in range(5):
            sd[p + f"{i}.conv.weight"].mul_(GAIN_SPYNET)
        sd[p + "4.conv.bias"].copy_(randn(sd[p + "4.conv.bias"], 0.05))
    # final 64->3 projection: keep the reconstruction inside (0,1) so the clamp does not hide errors
    sd["loopfilter.featdown.weight"].mul_(GAIN_FEATDOWN / GAIN_BACKBONE)
    sd["loopfilter.featdown.bias"].fill_(0.5)
    for c in ("mvCoder.", "resCoder."):
        # (ii) latents y, z spanning several bins
        sd[c + "g_a.7.weight"].mul_(GAIN_Y[c])
        sd[c + "g_a.7.bias"].copy_(randn(sd[c + "g_a.7.bias"], 1.0))
        sd[c + "h_a.8.weight"].mul_(GAIN_Z)
        sd[c + "g_s.9.0.weight"].mul_(GAIN_XHAT[c])
        sd[c + "h_a.8.bias"].copy_(randn(sd[c + "h_a.8.bias"], 0.7))
        # (iii) exercise the reparametrisations and the per-channel medians
        q = sd[c + "entropy_bottleneck.quantiles"]
        q[:, :, 1].add_(randn(q[:, :, 1], 0.3))
        for i in range(4):
            f_ = sd[c + f"entropy_bottleneck._factor{i}"]
            f_.copy_(randn(f_, 0.3))
        for k2 in list(sd.keys()):
            if k2.startswith(c) and k2.endswith("gdn.beta"):
                sd[k2].add_(rand(sd[k2]) * 0.5)
            if k2.startswith(c) and (k2.endswith("gdn.gamma")):
                sd[k2].add_(rand(sd[k2]) * 0.02)
        # (iv) predicted scales spanning [0.05, 10], means ~ N(0,1)
        sd[c + "context_prediction.weight"].mul_(2.0)
        b = sd[c + "entropy_parameters.4.bias"]
        n = b.numel() // 2
        b[:n].copy_(torch.exp(rand(b[:n]) * (math.log(10.0) - math.log(0.05)) + math.log(0.05)))
        b[n:].copy_(randn(b[n:], 1.0))
        sd[c + "entropy_parameters.4.weight"].mul_(3.0)
    return sd


def _is_se_key(k):
    # SELayer convs are `<...>.conv1.conv.weight` / `conv2.conv.weight` (1x1, leave them alone)
    return k.endswith(".conv1.conv.weight") or k.endswith(".conv2.conv.weight")


# gains tuned once against the oracle at 256x256 (see oracle/make_golden.py --stats)
GAIN_BACKBONE = 1.6
GAIN_DCN_OFFSET_W = 0.1
GAIN_SPYNET = 1.7
GAIN_Y = {"mvCoder.": 150.0, "resCoder.": 150.0}
GAIN_Z = 10.0
GAIN_XHAT = {"mvCoder.": 2.0, "resCoder.": 0.3}
GAIN_FEATDOWN = 0.6


def state_checksum(sd):
    """Order-independent fp64 checksum of a state_dict (guards 'same weights' in golden tests)."""
    tot = 0.0
    for k in sorted(sd.keys()):
        v = sd[k]
        if v.numel() and v.dtype.is_floating_point:
            tot += float(v.double().abs().sum())
    return tot
