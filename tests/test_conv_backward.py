"""Convolution backward slice of the training step (the reference takes these gradients from cuDNN through
autograd, tools/train.py:125-159): `tdvc_b200.ops.conv2d` against torch's own float64 convolution and its autograd on the same
GPU.  grad_input runs on the forward kernels (tcgen05 hi/lo split) from the transposed, flipped weight; grad_weight / grad_bias
on the deterministic fp32-class `tdvc_conv2d_wgrad` (three TF32 products: tcgen05 for the stride-1 layers, warp-level MMAs else)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

CASES = [
    # N, cin, cout, k, stride, H, W, act
    (2, 64, 64, 3, 1, 40, 56, "relu"),          # Res_Block convolution (reference utils.py:43-56)
    (1, 64, 128, 3, 2, 64, 96, "leaky_relu"),   # ResidualBlockWithStride.conv1
    (1, 64, 128, 1, 2, 64, 96, None),           # ... .skip
    (2, 3, 64, 3, 1, 48, 40, "leaky_relu"),     # image-input layer (3 channels, stored as 4)
    (1, 8, 32, 7, 1, 32, 64, "relu"),           # SPyNet basic module
    (1, 64, 32, 7, 1, 20, 40, "relu"),          # ... its 64-channel layer (single-stage three-product weight gradient)
    (1, 128, 256, 5, 1, 24, 40, None),          # context model size
    (1, 426, 341, 1, 1, 16, 24, "leaky_relu"),  # entropy_parameters[2]: channel counts that are no multiple of 4
    (1, 64, 3, 3, 1, 32, 32, "clamp01"),        # image head
]


def _ref(x, w, b, stride, pad, act, slope):
    y = F.conv2d(x, w, b, stride, pad)
    if act == "relu":
        y = F.relu(y)
    elif act == "leaky_relu":
        y = F.leaky_relu(y, slope)
    elif act == "clamp01":
        y = y.clamp(0.0, 1.0)
    return y


@pytest.mark.parametrize("case", CASES, ids=[f"{c[1]}to{c[2]}k{c[3]}s{c[4]}" for c in CASES])
def test_conv2d_forward_backward_vs_torch_fp64(case):
    from tdvc_b200 import ops
    N, ci, co, k, s, H, W, act = case
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(k * 1000 + ci)
    x = torch.randn(N, ci, H, W, generator=g).to(dev)
    w = (torch.randn(co, ci, k, k, generator=g) / (ci * k * k) ** 0.5).to(dev)
    b = (torch.randn(co, generator=g) * 0.1 + (0.5 if act == "clamp01" else 0.0)).to(dev)
    gy = torch.randn(N, co, (H + 2 * (k // 2) - k) // s + 1, (W + 2 * (k // 2) - k) // s + 1, generator=g).to(dev)
    slope = 0.1
    xs, ws, bs = (t.clone().requires_grad_(True) for t in (x, w, b))
    y = ops.conv2d(xs, ws, bs, s, k // 2, act, slope)
    y.backward(gy)
    xd, wd, bd = (t.double().clone().requires_grad_(True) for t in (x, w, b))
    yd = _ref(xd, wd, bd, s, k // 2, act, slope)
    # the activation mask of the reference must be the one our forward saw (outputs within rounding of a kink flip it)
    yd.backward(gy.double())
    assert y.shape == yd.shape
    assert (y.double() - yd).abs().max().item() <= 2e-5 * max(1.0, yd.abs().max().item())
    pre = F.conv2d(xd.detach(), wd.detach(), bd.detach(), s, k // 2)
    kink = (pre.abs() < 1e-4) if act in ("relu", "leaky_relu") else torch.zeros_like(pre, dtype=torch.bool)
    if act == "clamp01":
        kink = (pre.abs() < 1e-4) | ((pre - 1).abs() < 1e-4)
    assert kink.float().mean().item() < 0.01
    tol = 1.0 if not kink.any() else 50.0   # a flipped mask element moves a few gradient entries by one term of the sum
    for name, a, r, rel in (("grad_input", xs.grad, xd.grad, 3e-5), ("grad_weight", ws.grad, wd.grad, 1e-4),
                            ("grad_bias", bs.grad, bd.grad, 1e-4)):
        assert a is not None and a.shape == r.shape, name
        err = (a.double() - r).abs().max().item()
        assert err <= tol * rel * max(1.0, r.abs().max().item()), (name, err, r.abs().max().item())


AMP_CASES = [
    # N, cin, cout, k, stride, H, W: the layer classes of the training step (weight gradient with one TF32 product, enabled_amp)
    (2, 64, 64, 3, 1, 36, 64),      # Res_Block: tcgen05 kernel, 64-channel tiles (2 kernel rows per M = 128 operand)
    (1, 128, 192, 3, 1, 12, 96),    # coder layers: several ci / co tiles
    (2, 3, 64, 3, 1, 21, 50),       # image input: one 32-channel group zero-filled past channel 3, ragged tiles
    (1, 64, 216, 3, 1, 16, 40),     # DCN offset / mask head: last co tile 24 channels wide
    (1, 64, 3, 3, 1, 16, 40),       # image head
    (1, 8, 32, 7, 1, 20, 40),       # SPyNet basic module, 4 kernel rows per operand
    (1, 32, 64, 7, 1, 20, 40),      # ... kernel columns split over two CTAs
    (1, 64, 32, 7, 1, 20, 40),      # ... two-row items
    (1, 16, 2, 7, 1, 20, 40),       # ... flow head
    (1, 192, 64, 1, 1, 20, 40),     # 1x1 fusion layer
    (1, 64, 128, 3, 2, 32, 64),     # strided layer: warp-level kernel
    (1, 128, 128, 3, 1, 8, 8),      # too small for the tensor-core items: warp-level kernel
]


@pytest.mark.parametrize("case", AMP_CASES, ids=[f"{c[1]}to{c[2]}k{c[3]}s{c[4]}@{c[5]}x{c[6]}" for c in AMP_CASES])
def test_conv2d_weight_gradient_one_tf32_product(case):
    """`enabled_amp=True` selects one TF32 product per MAC for the weight gradients (ops.WGRAD_PRODUCTS = 1: csrc/wgrad_tc.cu
    on tcgen05 where the shape allows, else the warp-level kernel): against float64 within TF32 operand rounding (2^-11 per
    operand; the tensor core truncates), bit-identical from run to run."""
    from tdvc_b200 import ops
    N, ci, co, k, s, H, W = case
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(7 * k + ci + co)
    x = torch.randn(N, ci, H, W, generator=g).to(dev)
    w = (torch.randn(co, ci, k, k, generator=g) / (ci * k * k) ** 0.5).to(dev)
    b = (torch.randn(co, generator=g) * 0.1).to(dev)
    Ho, Wo = (H + 2 * (k // 2) - k) // s + 1, (W + 2 * (k // 2) - k) // s + 1
    gy = torch.randn(N, co, Ho, Wo, generator=g).to(dev)
    saved, ops.WGRAD_PRODUCTS = ops.WGRAD_PRODUCTS, 1
    try:
        grads = []
        for _ in range(2):
            ws, bs = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
            ops.conv2d(x, ws, bs, s, k // 2).backward(gy)
            grads.append((ws.grad, bs.grad))
    finally:
        ops.WGRAD_PRODUCTS = saved
    wd, bd = w.double().requires_grad_(True), b.double().requires_grad_(True)
    F.conv2d(x.double(), wd, bd, s, k // 2).backward(gy.double())
    assert torch.equal(grads[0][0], grads[1][0]) and torch.equal(grads[0][1], grads[1][1])
    assert (grads[0][0].double() - wd.grad).abs().max().item() <= 2e-3 * wd.grad.abs().max().item()
    assert (grads[0][1].double() - bd.grad).abs().max().item() <= 1e-5 * max(1.0, bd.grad.abs().max().item())


def test_conv2d_backward_is_deterministic_and_composes():
    """Two backward passes give identical bits; a Res_Block (reference utils.py:43-56) built from ops.conv2d has the gradients of
    the same block in torch."""
    from tdvc_b200 import ops
    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    x = torch.randn(2, 64, 48, 64, device=dev)
    w1, w2 = (torch.randn(64, 64, 3, 3, device=dev) / 24.0 for _ in range(2))
    b1, b2 = (torch.randn(64, device=dev) * 0.1 for _ in range(2))

    def run(double):
        ps = [t.double() if double else t.clone() for t in (x, w1, b1, w2, b2)]
        ps = [t.requires_grad_(True) for t in ps]
        xx, a1, c1, a2, c2 = ps
        if double:
            out = xx + F.conv2d(F.relu(F.conv2d(xx, a1, c1, 1, 1)), a2, c2, 1, 1)
        else:
            out = xx + ops.conv2d(ops.conv2d(xx, a1, c1, 1, 1, "relu"), a2, c2, 1, 1)
        (out * out).sum().backward()
        return [p.grad for p in ps]

    ga, gb, gd = run(False), run(False), run(True)
    for a, b_, d in zip(ga, gb, gd):
        assert torch.equal(a, b_)
        assert (a.double() - d).abs().max().item() <= 2e-4 * d.abs().max().item()


def test_conv2d_rejects_bad_input():
    from tdvc_b200 import ops
    dev = torch.device("cuda:0")
    x, w = torch.zeros(1, 8, 16, 16, device=dev), torch.zeros(8, 8, 3, 3, device=dev)
    with pytest.raises(RuntimeError):
        ops.conv2d(x.cpu(), w.cpu())
    with pytest.raises(RuntimeError):
        ops.conv2d(x, w, None, 1, 0)          # padding != (k - 1) / 2
    with pytest.raises(RuntimeError):
        ops.conv2d(x, torch.zeros(8, 4, 3, 3, device=dev), None, 1, 1)


@pytest.mark.parametrize("inverse", [False, True], ids=["gdn", "igdn"])
def test_gdn_forward_backward_vs_torch_fp64(inverse):
    """compressai GDN / IGDN: out = x * (beta + gamma . x^2)^(-+1/2); forward on the fused tcgen05 kernel,
    backward through the two element-wise kernels + the 1x1 convolution's dgrad / wgrad, against float64 autograd."""
    from tdvc_b200 import ops
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(11 + inverse)
    C = 128
    x = (torch.randn(2, C, 24, 40, generator=g) * 3.0).to(dev)
    gamma = (0.1 * torch.eye(C) + 0.01 * torch.rand(C, C, generator=g)).to(dev)
    beta = (1.0 + 0.2 * torch.rand(C, generator=g)).to(dev)
    gy = torch.randn(2, C, 24, 40, generator=g).to(dev)
    xs, bs, gs = (t.clone().requires_grad_(True) for t in (x, beta, gamma))
    y = ops.gdn(xs, bs, gs, inverse)
    y.backward(gy)
    xd, bd, gd = (t.double().clone().requires_grad_(True) for t in (x, beta, gamma))
    norm = F.conv2d(xd * xd, gd.reshape(C, C, 1, 1), bd)
    yd = xd * (torch.sqrt(norm) if inverse else torch.rsqrt(norm))
    yd.backward(gy.double())
    assert (y.double() - yd).abs().max().item() <= 2e-5 * yd.abs().max().item()
    for name, a, r, rel in (("grad_input", xs.grad, xd.grad, 5e-5), ("grad_beta", bs.grad, bd.grad, 2e-4),
                            ("grad_gamma", gs.grad, gd.grad, 2e-4)):
        assert a is not None and a.shape == r.shape, name
        err = (a.double() - r).abs().max().item()
        assert err <= rel * r.abs().max().item(), (name, err, r.abs().max().item())
    # large activations: the squared operand is range-scaled instead of saturating fp16 (|x| > 255)
    xb = (x * 200.0).requires_grad_(True)
    yb = ops.gdn(xb, beta, gamma, inverse)
    nb = F.conv2d((x * 200.0).double() ** 2, gamma.double().reshape(C, C, 1, 1), beta.double())
    rb = (x * 200.0).double() * (torch.sqrt(nb) if inverse else torch.rsqrt(nb))
    assert (yb.double() - rb).abs().max().item() <= 5e-5 * rb.abs().max().item()


def test_gaussian_conditional_bits_backward_vs_oracle_fp64():
    """sum ln p of the conditional Gaussian in noise mode and its gradients against the oracle's GaussianConditional in float64
    (compressai's LowerBound gradient rule included: scales below 0.11, likelihoods below 1e-9, both signs of the upstream
    gradient)."""
    from oracle.compressai_port import GaussianConditional
    from tdvc_b200 import ops
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(21)
    shape = (2, 128, 8, 12)
    y = torch.randn(shape, generator=g) * 4
    means = torch.randn(shape, generator=g) * 2
    scales = torch.exp(torch.randn(shape, generator=g) * 1.5 - 0.5)       # spans both sides of the 0.11 bound
    y[0, :4] += 40.0                                                     # far tails: likelihood below the 1e-9 bound
    noise = torch.rand(shape, generator=g) - 0.5
    gc = GaussianConditional(None).double()
    for gs in (-0.37, 0.8):
        ys, ss, ms = (t.to(dev).requires_grad_(True) for t in (y, scales, means))
        S = ops.gaussian_conditional_bits(ys, ss, ms, noise.to(dev))
        (S * gs).backward()
        yd, sd, md = (t.double().requires_grad_(True) for t in (y, scales, means))
        lik = gc.likelihood_lower_bound(gc._likelihood(yd + noise.double(), sd, md))
        Sd = torch.log(lik).sum()
        (Sd * gs).backward()
        assert abs(S.item() - Sd.item()) <= 2e-5 * abs(Sd.item())
        for name, a, r in (("y", ys.grad, yd.grad), ("scales", ss.grad, sd.grad), ("means", ms.grad, md.grad)):
            a = a.cpu().double()
            # fp32 erfc differences are amplified where p is tiny (g / p): compare relative to the local gradient scale
            err = ((a - r).abs() / (r.abs() + 1e-3 * r.abs().max())).max().item()
            assert err <= 2e-3, (name, gs, err)
            # the bound rule zeroes the same entries (fp32 additionally underflows gradients below ~1e-38 in the far tails)
            assert (a[r == 0] == 0).all() and (a[r.abs() > 1e-12 * r.abs().max()] != 0).all(), name


def test_entropy_bottleneck_bits_backward_vs_oracle_fp64():
    from oracle.compressai_port import EntropyBottleneck
    from tdvc_b200 import ops
    dev = torch.device("cuda:0")
    torch.manual_seed(31)
    C = 128
    eb = EntropyBottleneck(C)
    with torch.no_grad():
        for n, p in eb.named_parameters():
            if n.startswith("_factor"):
                p.copy_(torch.randn_like(p) * 0.3)
            elif n.startswith("_matrix"):
                p.add_(torch.randn_like(p) * 0.2)
    ebd = EntropyBottleneck(C).double()
    ebd.load_state_dict(eb.state_dict())
    z = torch.randn(2, C, 4, 6) * 3
    z[0, :3] += 60.0                                  # likelihood below the bound
    noise = torch.rand(z.shape) - 0.5
    gzt = torch.randn(z.shape) * 0.01                 # gradient arriving at z~ from its consumers (h_s)
    for gs in (-0.37, 0.8):
        zs = z.to(dev).requires_grad_(True)
        params = {n: p.detach().to(dev).requires_grad_(True) for n, p in eb.named_parameters() if n != "quantiles"}
        zt, S = ops.entropy_bottleneck_bits(zs, noise.to(dev), [params[f"_matrix{i}"] for i in range(5)],
                                            [params[f"_bias{i}"] for i in range(5)], [params[f"_factor{i}"] for i in range(4)])
        (S * gs + (zt * gzt.to(dev)).sum()).backward()
        zd = z.double().requires_grad_(True)
        for p in ebd.parameters():
            p.grad = None
        v = (zd + noise.double()).permute(1, 0, 2, 3).reshape(C, 1, -1)
        lik = ebd.likelihood_lower_bound(ebd._likelihood(v))
        Sd = torch.log(lik).sum()
        (Sd * gs + ((zd + noise.double()) * gzt.double()).sum()).backward()
        assert (zt.cpu() - (z + noise)).abs().max().item() < 1e-6
        assert abs(S.item() - Sd.item()) <= 2e-5 * abs(Sd.item())
        pairs = [("z", zs.grad, zd.grad)] + [(n, params[n].grad, dict(ebd.named_parameters())[n].grad) for n in params]
        for name, a, r in pairs:
            assert a is not None, name
            a = a.cpu().double()
            err = (a - r).abs().max().item()
            assert err <= 1e-3 * max(r.abs().max().item(), 1e-6), (name, gs, err, r.abs().max().item())


def test_channel_mean_and_scale_vs_torch():
    """The squeeze-excitation pieces (reference inflate.py:159-208): spatial mean and gated product, forward and backward."""
    from tdvc_b200 import ops
    dev = torch.device("cuda:0")
    torch.manual_seed(41)
    x = torch.randn(2, 128, 24, 40, device=dev)
    s = torch.rand(2, 128, device=dev)
    gm, gy = torch.randn(2, 128, device=dev), torch.randn(2, 128, 24, 40, device=dev)
    xs, ss = x.clone().requires_grad_(True), s.clone().requires_grad_(True)
    m = ops.channel_mean(xs)
    y = ops.channel_scale(xs, ss)
    ((m * gm).sum() + (y * gy).sum()).backward()
    xd, sd = x.double().requires_grad_(True), s.double().requires_grad_(True)
    md = xd.mean((2, 3))
    yd = xd * sd[:, :, None, None]
    ((md * gm.double()).sum() + (yd * gy.double()).sum()).backward()
    assert (m.double() - md).abs().max().item() < 1e-6 and (y.double() - yd).abs().max().item() < 1e-6
    assert (xs.grad.double() - xd.grad).abs().max().item() < 1e-5
    assert (ss.grad.double() - sd.grad).abs().max().item() <= 1e-5 * sd.grad.abs().max().item()


def test_spynet_level_input_vs_torch():
    """One SPyNet level's input assembly (reference flownet.py:8-48, 116-138) and its gradient to the coarse flow against the
    torch formulation (interpolate x2 align_corners=True, * 2, grid_sample border / align_corners=True, cat)."""
    from oracle.model import flow_warp_border
    from tdvc_b200 import ops
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(51)
    N, h, w = 2, 32, 48
    ref, supp = torch.rand(N, 3, h, w, generator=g), torch.rand(N, 3, h, w, generator=g)
    flow = torch.randn(N, 2, h // 2, w // 2, generator=g) * 3.0
    flow[0, :, 0, :] = -40.0      # pushes samples beyond the border: clamped coordinates, zero derivative
    gy = torch.randn(N, 8, h, w, generator=g)
    fs = flow.to(dev).requires_grad_(True)
    y = ops.spynet_level_input(ref.to(dev), supp.to(dev), fs)
    (y * gy.to(dev)).sum().backward()
    fd = flow.double().requires_grad_(True)
    up = F.interpolate(fd, scale_factor=2, mode="bilinear", align_corners=True) * 2.0
    yd = torch.cat([ref.double(), flow_warp_border(supp.double(), up.permute(0, 2, 3, 1)), up], 1)
    (yd * gy.double()).sum().backward()
    # forward against the same fp32 arithmetic (white-noise images: a 1e-5 coordinate rounding moves a sample by 1e-5)
    up32 = F.interpolate(flow, scale_factor=2, mode="bilinear", align_corners=True) * 2.0
    y32 = torch.cat([ref, flow_warp_border(supp, up32.permute(0, 2, 3, 1)), up32], 1)
    assert (y.detach().cpu() - y32).abs().max().item() < 2e-5
    assert (y.detach().cpu().double() - yd).abs().max().item() < 2e-4
    err = (fs.grad.cpu().double() - fd.grad).abs().max().item()
    assert err <= 1e-3 * fd.grad.abs().max().item(), (err, fd.grad.abs().max().item())
    y0 = ops.spynet_level_input(ref.to(dev), supp.to(dev), None)      # coarsest level: zero flow
    assert (y0[:, 3:6].cpu() - supp).abs().max().item() < 1e-5 and y0[:, 6:].abs().max().item() == 0.0
