"""GPU tests (-m gpu): each CUDA kernel, called through the C-ABI, against the oracle / plain torch fp32 on the
CPU for the same seeded inputs.  Tolerances: fp32 kernels 2e-5 relative to the output scale (different
summation order only); the tcgen05 fp16 hi/lo-split convolution 2e-4 relative (fp16 residual
rounding, ~2^-16 per product); integer / index outputs bit-exact."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def plan(dev):
    from tdvc_b200.model import _Plan
    return _Plan(1, 64, 64, dev)


def _conv_case(plan, dev, cin_list, cout, k, stride, H, W, N=1, act=0, slope=0.0, shuffle=0, res=False, impl=1, seed=0,
               pad=None, products=0):
    from tdvc_b200.model import Act, pack_conv
    torch.manual_seed(seed)
    cin = sum(cin_list)
    conv = torch.nn.Conv2d(cin, cout, k, stride, k // 2 if pad is None else pad)
    xs = [torch.randn(N, c, H, W) for c in cin_list]
    if products == 1:   # the one-product scheme multiplies fp16(w) by fp16(x) exactly and accumulates in fp32
        want = F.conv2d(torch.cat(xs, 1).half().float(), conv.weight.half().float(), conv.bias, stride, k // 2 if pad is None else pad)
    else:
        want = conv(torch.cat(xs, 1))
    if act == 1:
        want = F.relu(want)
    elif act == 2:
        want = F.leaky_relu(want, slope)
    elif act == 3:
        want = want.clamp(0.0, 1.0)
    if shuffle == 2:
        want = F.pixel_shuffle(want, 2)
    r = torch.randn_like(want) if res else None
    if res:
        want = want + r
    layout = [(c, (c + 3) // 4 * 4) for c in cin_list]
    cw = pack_conv(conv.weight.to(dev), conv.bias.to(dev), src_layout=layout, shuffle=shuffle, pad=pad, stride=stride)
    from tdvc_b200 import tc
    tc.attach_f16({"w": cw}, one_product=products == 1)
    srcs = [Act.from_nchw(x.to(dev), ld=(x.shape[1] + 3) // 4 * 4) for x in xs]
    out = Act.alloc(N, want.shape[2], want.shape[3], want.shape[1], dev, ld=(want.shape[1] + 3) // 4 * 4 if want.shape[1] % 4 else None)
    r1 = Act.from_nchw(r.to(dev)) if res else None
    plan.conv(srcs, cw, out, stride=stride, act=act, slope=slope, res1=r1, impl=impl, products=products)
    torch.cuda.synchronize()
    return out.nchw().cpu(), want.detach()


CONV_CASES = [
    # cin_list, cout, k, stride, H, W, kwargs
    ([64], 64, 3, 1, 24, 40, dict(act=1)),
    ([64], 64, 3, 1, 17, 23, dict(act=2, slope=0.1, res=True)),       # ragged tile edges
    ([64, 64], 64, 3, 1, 16, 16, dict()),                              # dual-source (torch.cat replacement)
    ([3], 64, 3, 1, 20, 20, dict(N=2)),                                # image input, channel pad 3->4
    ([64], 128, 3, 2, 32, 32, dict(act=2, slope=0.01)),               # stride 2
    ([64], 128, 1, 2, 32, 32, dict(pad=0)),                            # 1x1 stride-2 skip
    ([128], 512, 3, 1, 8, 12, dict(shuffle=2, act=2, slope=0.01)),    # subpel conv (PixelShuffle folded in)
    ([128], 256, 5, 1, 8, 8, dict()),                                  # 5x5 (context model shape)
    ([8], 32, 7, 1, 16, 16, dict(act=1)),                              # SPyNet 7x7
    ([16], 2, 7, 1, 16, 16, dict(res=True)),                           # SPyNet last layer, cout 2
    ([64], 216, 3, 1, 16, 16, dict()),                                 # DCN offset/mask head
    ([64], 3, 3, 1, 16, 16, dict()),                                   # featdown
    ([64, 64, 64], 64, 1, 1, 8, 8, dict(pad=0)),                       # temporal (3,1,1) as 1x1 over 3 sources
    ([192], 768, 3, 1, 4, 4, dict(shuffle=2)),                         # h_s subpel
]


@pytest.mark.parametrize("idx", range(len(CONV_CASES)))
def test_conv2d_simt_vs_torch(plan, dev, idx):
    cin_list, cout, k, stride, H, W, kw = CONV_CASES[idx]
    got, want = _conv_case(plan, dev, cin_list, cout, k, stride, H, W, impl=1, seed=idx, **kw)
    assert got.shape == want.shape
    assert (got - want).abs().max() <= 2e-5 * max(1.0, want.abs().max().item())


TC_CASES = [
    # shapes with a tcgen05 path (csrc/conv_tc.cu `choose`): cin_list, cout, k, stride, H, W, kwargs
    ([64], 64, 3, 1, 40, 56, dict(act=1)),
    ([64], 64, 3, 1, 17, 23, dict(act=2, slope=0.1, res=True)),       # ragged tile edges
    ([64, 64], 64, 3, 1, 33, 16, dict()),                              # two K units from two sources
    ([128], 128, 3, 1, 24, 24, dict(act=2, slope=0.01)),              # two cout tiles
    ([128], 512, 3, 1, 8, 12, dict(shuffle=2, act=2, slope=0.01)),    # PixelShuffle store
    ([64], 216, 3, 1, 16, 16, dict()),                                 # cout not a multiple of the tile
    ([3], 64, 3, 1, 20, 36, dict(N=2, act=2, slope=0.1)),             # image input (CK=16)
    ([64], 3, 3, 1, 16, 40, dict()),                                   # featdown (NT=16)
    ([128], 128, 1, 1, 19, 21, dict(pad=0)),                           # 1x1
    ([64, 64, 64, 64], 64, 1, 1, 16, 16, dict(pad=0, act=2, slope=0.1)),  # 1x1 fusion over 4 sources
    ([256, 256], 426, 1, 1, 8, 8, dict(pad=0)),                        # entropy_parameters[0]
    ([428], 341, 1, 1, 8, 8, dict(pad=0)),                             # entropy_parameters[2] (cin % 64 != 0)
    ([128], 256, 5, 1, 16, 16, dict()),                                # context model 5x5
    ([8], 32, 7, 1, 16, 24, dict(act=1)),                              # SPyNet 7x7
    ([16], 2, 7, 1, 16, 16, dict(res=True)),
    ([32], 64, 7, 1, 32, 16, dict(act=1)),
    ([64], 32, 7, 1, 70, 20, dict(act=1)),                             # <= 32 output channels: two row phases per item
    ([32], 16, 7, 1, 33, 9, dict(act=1, N=2)),                         # odd height: the last row has phase 0 only
    ([8], 32, 7, 1, 129, 17, dict(act=2, slope=0.1, res=True)),        # three 64-row items, ragged
    ([64], 128, 3, 2, 32, 48, dict(act=2, slope=0.01)),               # stride 2
    ([128], 128, 3, 2, 18, 30, dict(N=2)),
    ([64], 128, 1, 2, 32, 32, dict(pad=0)),                            # 1x1 stride-2 skip
]


@pytest.mark.parametrize("idx", range(len(TC_CASES)))
def test_conv2d_tcgen05_vs_torch(plan, dev, idx):
    """impl=2 forces the tensor-core kernel (fp16 hi/lo split, fp32 accumulate): 2e-4 of the output scale."""
    cin_list, cout, k, stride, H, W, kw = TC_CASES[idx]
    got, want = _conv_case(plan, dev, cin_list, cout, k, stride, H, W, impl=2, seed=100 + idx, **kw)
    assert got.shape == want.shape
    assert (got - want).abs().max() <= 2e-4 * max(1.0, want.abs().max().item())


P1_CASES = [
    # one fp16 MMA product (TdvcConvParams::products = 1): cin_list, cout, k, stride, H, W, kwargs
    ([64], 64, 3, 1, 40, 56, dict(act=1)),                              # 64 channels x 2 row phases per item
    ([64], 64, 3, 1, 17, 23, dict(act=2, slope=0.1, res=True)),        # ragged edges, odd height: last row has phase 0 only
    ([64], 64, 3, 1, 129, 9, dict(res=True, N=2)),                      # three 64-row items
    ([64, 64], 64, 3, 1, 33, 16, dict(act=2, slope=0.1)),              # featfusion: four K units from two sources
    ([64], 48, 3, 1, 20, 20, dict()),                                   # fewer channels than a phase holds
    ([128], 128, 3, 1, 24, 24, dict(act=2, slope=0.01, res=True)),     # 128 channels per item, one phase
    ([128], 512, 3, 1, 8, 12, dict(shuffle=2, act=2, slope=0.01)),     # subpel conv
    ([128], 256, 3, 1, 16, 16, dict(shuffle=2)),                        # g_s[9]
    ([128], 128, 1, 1, 19, 21, dict(pad=0)),                            # 1x1
    ([64], 3, 3, 1, 37, 33, dict(act=3)),                               # featdown + clamp: 3 of the 128 rows carry data
]


@pytest.mark.parametrize("idx", range(len(P1_CASES)))
def test_conv2d_one_product_vs_torch(plan, dev, idx):
    """products = 1: against the fp32 convolution of the fp16-rounded operands (what one fp16 MMA with fp32 accumulation
    computes): only the summation order differs."""
    cin_list, cout, k, stride, H, W, kw = P1_CASES[idx]
    got, want = _conv_case(plan, dev, cin_list, cout, k, stride, H, W, impl=2, seed=400 + idx, products=1, **kw)
    assert got.shape == want.shape
    assert (got - want).abs().max() <= 3e-5 * max(1.0, want.abs().max().item())


@pytest.mark.parametrize("case", [(64, 64, 3, 2, 45, 70, 0, False), (128, 128, 3, 1, 21, 19, 0, True), (64, 64, 3, 1, 33, 40, 1, False),
                                  (128, 64, 1, 3, 16, 24, 0, False), (128, 128, 3, 1, 20, 12, 1, True)])
def test_conv2d_chan_sum_feeds_squeeze_excitation(plan, dev, case):
    """TdvcConvParams::chan_sum: the channel sums of a convolution's output, accumulated in its epilogue (every tile scheme:
    hi/lo rows, split, one product with and without row phases; batches), equal the sums over the stored tensor, and the
    squeeze-excitation built on them equals the one built on a separate pass over the tensor."""
    from oracle.model import SELayer
    from tdvc_b200.model import Act, pack_conv
    from tdvc_b200 import tc
    cin, cout, k, N, H, W, products, res = case
    torch.manual_seed(31)
    conv = torch.nn.Conv2d(cin, cout, k, 1, k // 2)
    x = torch.randn(N, cin, H, W)
    r = torch.randn(N, cout, H, W) if res else None
    cw = pack_conv(conv.weight.to(dev), conv.bias.to(dev), src_layout=[(cin, cin)])
    tc.attach_f16({"w": cw}, one_product=products == 1)
    out = Act.alloc(N, H, W, cout, dev)
    plan.conv([Act.from_nchw(x.to(dev))], cw, out, act=2, slope=0.1, res1=Act.from_nchw(r.to(dev)) if res else None, impl=2,
              products=products, csum=True)
    assert plan.last_csum is not None
    part, rows = plan.last_csum
    sums = part.view(rows, N, cout).double().sum(0).cpu()
    want = out.nchw().double().sum((2, 3)).cpu()
    assert (sums - want).abs().max().item() <= 1e-5 * max(1.0, want.abs().max().item()) * (H * W) ** 0.5
    se = SELayer(cout)
    w = [se.conv1.conv.weight.reshape(cout // 16, cout).contiguous().to(dev), se.conv1.conv.bias.to(dev),
         se.conv2.conv.weight.reshape(cout, cout // 16).contiguous().to(dev), se.conv2.conv.bias.to(dev)]
    o1, o2 = Act.alloc(N, H, W, cout, dev), Act.alloc(N, H, W, cout, dev)
    plan.se(out, w, o1, csum=plan.last_csum)
    plan.se(out, w, o2)
    ref = se(out.nchw().cpu())
    assert (o1.nchw().cpu() - ref).abs().max().item() < 2e-5 * max(1.0, ref.abs().max().item())
    assert (o1.nchw() - o2.nchw()).abs().max().item() < 2e-6 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("case", [(64, 64, 3, 1024, 1920, 1, 0), (64, 64, 3, 1024, 1920, 1, 1), (128, 128, 3, 512, 960, 1, 1),
                                  (4, 64, 3, 512, 960, 2, 0), (64, 32, 7, 512, 960, 1, 0), (128, 256, 5, 64, 120, 1, 0)])
def test_conv_tma_fed_equals_ldg_fed(plan, dev, case):
    """The two activation paths of conv_tc - tensor-map bulk copies into a staging ring (UTMALDG) and per-lane LDG - feed the same
    operand planes: identical bits, repeatedly, at sizes where every CTA walks dozens of items (a missing proxy fence between the
    converters' reads of a staged box and the next bulk copy into it once showed up only there)."""
    import os
    from tdvc_b200.model import Act, pack_conv
    from tdvc_b200 import tc
    cin, cout, k, H, W, N, products = case
    torch.manual_seed(17)
    conv = torch.nn.Conv2d(cin, cout, k, 1, k // 2).to(dev)
    cw = pack_conv(conv.weight, conv.bias, src_layout=[(cin, (cin + 3) // 4 * 4)])
    tc.attach_f16({"w": cw}, one_product=products == 1)
    x = Act.alloc(N, H, W, cin, dev, ld=(cin + 3) // 4 * 4, zero=True)
    x.t[..., :cin].normal_()
    saved = {k_: os.environ.pop(k_, None) for k_ in ("TDVC_B200_CONV_LDG", "TDVC_B200_CONV_TMA")}
    try:
        outs = []
        for var in ("TDVC_B200_CONV_LDG", "TDVC_B200_CONV_TMA", "TDVC_B200_CONV_TMA", "TDVC_B200_CONV_TMA"):
            os.environ[var] = "1"
            out = Act.alloc(N, H, W, cout, dev, ld=(cout + 3) // 4 * 4, zero=True)
            plan.conv([x], cw, out, impl=2, products=products, act=2, slope=0.1)
            torch.cuda.synchronize()
            outs.append(out.t)
            del os.environ[var]
    finally:
        for k_, v in saved.items():
            if v is not None:
                os.environ[k_] = v
    assert all(torch.equal(outs[0], o) for o in outs[1:])


def test_conv2d_out_absmax(plan, dev):
    """TdvcConvParams::out_absmax: max |v| over everything the layer stored, on the tensor-core and the SIMT kernel."""
    from tdvc_b200.model import Act, pack_conv
    from tdvc_b200 import tc
    torch.manual_seed(5)
    for cout, H, W in ((128, 21, 19), (64, 33, 10)):
        conv = torch.nn.Conv2d(64, cout, 3, 1, 1)
        x = torch.randn(2, 64, H, W)
        cw = pack_conv(conv.weight.to(dev), conv.bias.to(dev), src_layout=[(64, 64)])
        tc.attach_f16({"w": cw})
        out = Act.alloc(2, H, W, cout, dev)
        for impl in (1, 2):
            am = torch.zeros(1, device=dev)
            plan.conv([Act.from_nchw(x.to(dev))], cw, out, impl=impl, absmax_out=am.data_ptr())
            assert am.item() == out.nchw().abs().max().item()


@pytest.mark.parametrize("inverse", [False, True])
@pytest.mark.parametrize("amp", [300.0, 3000.0])
def test_gdn_large_activations(plan, dev, inverse, amp):
    """|x| > 255 made x*x saturate the fp16 operands of the tensor-core GDN.  With `in_absmax` (maintained by the
    producing layer's epilogue) x*x is pre-scaled by an exact power of two: the result follows the fp32 reference."""
    from oracle.compressai_port import GDN
    from tdvc_b200 import lib as L
    from tdvc_b200 import tc
    from tdvc_b200.model import Act, _reparam, pack_conv
    torch.manual_seed(9)
    g = GDN(128, inverse=inverse)
    g.beta.data.add_(torch.rand(128) * 0.5)
    g.gamma.data.add_(torch.rand(128, 128) * 0.02)
    x = torch.randn(1, 128, 19, 21) * amp
    want = g(x)
    cw = pack_conv(_reparam(g.gamma, g.gamma_reparam).reshape(128, 128, 1, 1), _reparam(g.beta, g.beta_reparam))
    cw.w, cw.b = cw.w.to(dev), cw.b.to(dev)
    tc.attach_f16({"w": cw})
    xa = Act.from_nchw(x.to(dev))
    am = x.abs().max().reshape(1).to(dev)
    out = Act.alloc(1, 19, 21, 128, dev)
    for impl, tol in ((1, 2e-5), (2, 1e-4)):
        plan.conv([xa], cw, out, in_square=True, post=L.POST_IGDN if inverse else L.POST_GDN, mul=xa, impl=impl,
                  absmax_in=am.data_ptr())
        assert (out.nchw().cpu() - want).abs().max() < tol * want.abs().max(), impl


SMALL_CASES = [
    # <= 4 output channels, exact fp32 SIMT kernel (csrc/conv_small.cu): cin_list, cout, k, stride, H, W, kwargs
    ([16], 2, 7, 1, 16, 16, dict(res=True)),                           # SPyNet flow head
    ([16], 2, 7, 1, 45, 70, dict(res=True, N=2)),                      # several tiles, ragged edges, batch
    ([64], 3, 3, 1, 16, 40, dict()),                                   # featdown
    ([64], 3, 3, 1, 37, 33, dict(act=2, slope=0.1, res=True)),
    ([32], 4, 3, 1, 9, 5, dict(act=1)),
    ([16], 1, 7, 1, 33, 8, dict()),
]


@pytest.mark.parametrize("idx", range(len(SMALL_CASES)))
def test_conv2d_small_cout_vs_torch(plan, dev, idx):
    cin_list, cout, k, stride, H, W, kw = SMALL_CASES[idx]
    got, want = _conv_case(plan, dev, cin_list, cout, k, stride, H, W, impl=3, seed=200 + idx, **kw)
    assert got.shape == want.shape
    assert (got - want).abs().max() <= 2e-5 * max(1.0, want.abs().max().item())


@pytest.mark.parametrize("shape", [(2, 37, 44), (1, 64, 96), (1, 19, 42)])
def test_conv2d_planar_output_vs_torch(plan, dev, shape):
    """out_planar = 1 (the DCN offset / mask head, 64 -> 216): NCHW planes written by the tensor-core kernel.  W % 4 == 0 takes
    the tensor-map TMA store (box clipped at the right / bottom edge and at the channel count), W % 4 != 0 the st.global path."""
    from tdvc_b200.model import Act, pack_conv
    from tdvc_b200 import tc
    N, H, W = shape
    torch.manual_seed(300 + H)
    conv = torch.nn.Conv2d(64, 216, 3, 1, 1)
    x = torch.randn(N, 64, H, W)
    want = conv(x).detach()
    cw = pack_conv(conv.weight.to(dev), conv.bias.to(dev), src_layout=[(64, 64)])
    tc.attach_f16({"w": cw})
    guard = 64
    flat = torch.full((N * 216 * H * W + 2 * guard,), 7.0, device=dev)
    out = flat[guard:guard + N * 216 * H * W].view(N, 216, H, W)
    plan.conv([Act.from_nchw(x.to(dev))], cw, out, planar=True, impl=2)
    torch.cuda.synchronize()
    assert (out.cpu() - want).abs().max() <= 2e-4 * max(1.0, want.abs().max().item())
    assert bool((flat[:guard] == 7.0).all()) and bool((flat[-guard:] == 7.0).all())   # nothing written outside the tensor


def test_conv2d_rejects_bad_arguments(plan, dev):
    from tdvc_b200 import lib as L
    p = L.ConvParams()
    p.n_src = 0
    assert L.load().tdvc_conv2d(p, None) == -1
    assert b"n_src" in L.load().tdvc_last_error()


def test_gdn_as_fused_conv(plan, dev):
    """compressai GDN / IGDN = 1x1 conv of x^2 with fused rsqrt/sqrt * x epilogue."""
    from oracle.compressai_port import GDN
    from tdvc_b200.model import Act, _reparam, pack_conv
    torch.manual_seed(3)
    for inverse in (False, True):
        g = GDN(128, inverse=inverse)
        g.beta.data.add_(torch.rand(128) * 0.5)
        g.gamma.data.add_(torch.rand(128, 128) * 0.02)
        x = torch.randn(1, 128, 9, 11) * 2
        idt = torch.randn(1, 128, 9, 11)
        want = g(x) + idt
        gd = g.to(dev)
        cw = pack_conv(_reparam(gd.gamma, gd.gamma_reparam).reshape(128, 128, 1, 1), _reparam(gd.beta, gd.beta_reparam))
        xa = Act.from_nchw(x.to(dev))
        out = Act.alloc(1, 9, 11, 128, dev)
        from tdvc_b200 import lib as L
        from tdvc_b200 import tc
        tc.attach_f16({"w": cw})
        for impl, tol in ((1, 2e-5), (2, 1e-4)):
            plan.conv([xa], cw, out, in_square=True, post=L.POST_IGDN if inverse else L.POST_GDN, mul=xa,
                      res1=Act.from_nchw(idt.to(dev)), impl=impl)
            assert (out.nchw().cpu() - want).abs().max() < tol * want.abs().max(), impl


@pytest.mark.parametrize("cout,H,W,products", [(64, 17, 23, 0), (216, 33, 9, 0), (128, 40, 31, 0), (32, 35, 12, 0),
                                               (64, 33, 23, 1), (48, 65, 9, 1), (128, 40, 31, 1), (216, 19, 12, 1)])
def test_conv_tc_writes_only_its_view(plan, dev, cout, H, W, products):
    """Guard bands: the tensor-core kernel stores through a channel view (ld > C, offset 4 floats... here 8 to keep the
    16-byte alignment) of a larger tensor with guard rows before and after; everything outside the view keeps its
    sentinel (ragged tiles in x, y and in the channel dimension)."""
    from tdvc_b200.model import Act, pack_conv
    from tdvc_b200 import tc
    torch.manual_seed(3)
    conv = torch.nn.Conv2d(64, cout, 3, 1, 1)
    x = torch.randn(1, 64, H, W)
    want = conv(x)
    cw = pack_conv(conv.weight.to(dev), conv.bias.to(dev), src_layout=[(64, 64)])
    tc.attach_f16({"w": cw}, one_product=products == 1)
    ld = cout + 16
    big = torch.full((H + 4, W, ld), 12345.0, device=dev)          # 2 guard rows above and below
    view = Act(big, big.data_ptr() + 4 * (2 * W * ld + 8), 1, H, W, cout, ld)
    # the channel-sum rows of the launch live in a guarded buffer too (TdvcConvParams::chan_sum)
    real_raw, guarded = plan.raw, {}

    def raw(name, shape, dtype=torch.float32):
        if isinstance(name, tuple) and name[0] == "se_csum":
            g = torch.full((shape[0] + 512,), 777.0, device=dev)
            guarded["buf"] = g
            return g[256:256 + shape[0]]
        return real_raw(name, shape, dtype)

    plan.raw = raw
    try:
        plan.conv([Act.from_nchw(x.to(dev))], cw, view, impl=2, products=products, csum=cout <= 128)
    finally:
        plan.raw = real_raw
    torch.cuda.synchronize()
    if cout <= 128 and (cout <= 64 or products == 1 or cout >= 96):
        assert plan.last_csum is not None
    if plan.last_csum is not None:
        g = guarded["buf"]
        assert (g[:256] == 777.0).all() and (g[-256:] == 777.0).all()
        part, rows = plan.last_csum
        sums = part.view(rows, 1, cout).double().sum(0).cpu()[0]
        got_sum = big[2:H + 2, :, 8:8 + cout].double().sum((0, 1)).cpu()
        assert (sums - got_sum).abs().max().item() <= 1e-4 * max(1.0, got_sum.abs().max().item())
    got = big[2:H + 2, :, 8:8 + cout].permute(2, 0, 1).unsqueeze(0).cpu()
    if products == 1:
        want = F.conv2d(x.half().float(), conv.weight.half().float(), conv.bias, 1, 1)
    assert (got - want).abs().max() < 1e-4 * want.abs().max()
    assert (big[:2] == 12345.0).all() and (big[H + 2:] == 12345.0).all()
    assert (big[2:H + 2, :, :8] == 12345.0).all() and (big[2:H + 2, :, 8 + cout:] == 12345.0).all()


# ------------------------------------------------------------------------------------------------ DCN
def test_dcn_zero_offset_known_answer(dev):
    """reference main/utils/dcnv2/testcuda.py:36-71 (check_zero_offset): 2*dcn(x) == x."""
    from tdvc_b200.ops import dcn_v2_forward
    torch.manual_seed(0)
    for (N, C, H, W, dg) in ((2, 64, 13, 19, 8), (2, 8, 7, 9, 2)):  # fast path (8 ch/group) and generic path
        x = torch.randn(N, C, H, W, device=dev)
        wgt = torch.zeros(C, C, 3, 3, device=dev)
        for c in range(C):
            wgt[c, c, 1, 1] = 1.0
        off = torch.zeros(N, dg * 18, H, W, device=dev)
        msk = torch.sigmoid(torch.zeros(N, dg * 9, H, W, device=dev))
        out = dcn_v2_forward(x, wgt, torch.zeros(C, device=dev), off, msk, 3, 3, 1, 1, 1, 1, 1, 1, dg)
        assert (2 * out - x).abs().max() < 1e-6


@pytest.mark.parametrize("shape", [(1, 64, 64, 21, 30, 8), (2, 16, 24, 9, 7, 2), (1, 6, 4, 5, 6, 3)])
def test_dcn_forward_vs_oracle(dev, shape):
    """Random offsets (sigma 3 px, many samples leave the image), random masks: against the oracle restatement of
    dcn_v2_im2col_cuda.cu:125-195 + dcn_v2_cuda.cu:73-92."""
    from oracle import dcn_naive
    from tdvc_b200.ops import dcn_v2_forward
    N, C, O, H, W, dg = shape
    torch.manual_seed(5)
    x = torch.randn(N, C, H, W)
    wgt = torch.randn(O, C, 3, 3) * 0.1
    b = torch.randn(O)
    off = torch.randn(N, dg * 18, H, W) * 3.0
    msk = torch.rand(N, dg * 9, H, W)
    want = dcn_naive.dcn_v2_forward(x, wgt, b, off, msk, dg)
    got = dcn_v2_forward(x.to(dev), wgt.to(dev), b.to(dev), off.to(dev), msk.to(dev), 3, 3, 1, 1, 1, 1, 1, 1, dg).cpu()
    assert (got - want).abs().max() < 2e-5 * max(1.0, want.abs().max().item())


@pytest.mark.parametrize("shape", [(1, 21, 30, 64), (2, 16, 48, 64), (1, 40, 33, 37)])
def test_dcn_tcgen05_vs_oracle(dev, shape):
    """csrc/dcn_tc.cu (gather producers + tcgen05 contraction, group-planar input, planar offsets / mask logits)
    against the oracle restatement, including the reference's fp16 rounding of the result
    (dcn_v2_amp.py:67-69) and the LeakyReLU evaluated on the Half tensor (pnet.py:180).  Tolerance: the fp16 hi/lo
    split keeps ~2^-22 per product; after the fp16 rounding at most one fp16 ulp may differ
    where the fp32 value sits on a rounding boundary."""
    from oracle import dcn_naive
    from tdvc_b200 import lib as L, tc
    from tdvc_b200.model import Act
    N, H, W, O = shape
    C, dg = 64, 8
    torch.manual_seed(11)
    x = torch.randn(N, C, H, W)
    wgt = torch.randn(O, C, 3, 3) * 0.1
    b = torch.randn(O)
    off = torch.randn(N, dg * 18, H, W) * 3.0
    logit = torch.randn(N, dg * 9, H, W)
    exact = dcn_naive.dcn_v2_forward(x, wgt, b, off, torch.sigmoid(logit), dg)
    want = F.leaky_relu(exact.half(), 0.1).float()
    lib = L.load()
    st = torch.cuda.current_stream(dev).cuda_stream
    opad = 64
    pk = torch.zeros(C * 9, opad, device=dev)
    pk[:, :O] = wgt.reshape(O, C * 9).t().to(dev)
    packed = tc.attach_dcn_f16({"w": pk.contiguous()}, "w", O, dg)
    xa = Act.from_nchw(x.to(dev))
    gp = torch.empty(N * dg, H, W, 8, device=dev)
    L.check(lib.tdvc_nhwc_to_group_planar(xa.ptr, xa.ld, gp.data_ptr(), N, H, W, C, st), "gp")
    assert torch.equal(gp.view(N, dg, H, W, 8), x.to(dev).view(N, dg, 8, H, W).permute(0, 1, 3, 4, 2))
    om = torch.cat([off, logit], 1).contiguous().to(dev)       # (N, 216, H, W) as the offset conv writes it
    out = Act.alloc(N, H, W, O, dev, ld=(O + 7) // 8 * 8)
    for round_fp16 in (0, 1):
        dp = L.DcnParams()
        dp.input_gp, dp.params_planar = gp.data_ptr(), 1
        dp.offset, dp.off_ld = om.data_ptr(), 27 * dg
        dp.mask, dp.mask_ld, dp.mask_is_logit = om.data_ptr() + 4 * 18 * dg * H * W, 27 * dg, 1
        bd = b.to(dev)
        dp.weight_packed, dp.bias, dp.weight_f16 = packed["w"].data_ptr(), bd.data_ptr(), packed["w_f16"].data_ptr()
        dp.out, dp.out_ld = out.ptr, out.ld
        dp.N, dp.H, dp.W, dp.C, dp.O, dp.O_pad, dp.dg = N, H, W, C, O, opad, dg
        dp.round_fp16, dp.act, dp.slope, dp.impl = round_fp16, (L.ACT_LRELU if round_fp16 else L.ACT_NONE), 0.1, 2
        L.check(lib.tdvc_dcn_nhwc(dp, st), "dcn_tc")
        torch.cuda.synchronize()
        got = out.nchw().cpu()
        if round_fp16:
            diff = (got - want).abs()
            # the fp32 value is within 2e-5 * scale of the oracle (previous case); rounding both to fp16 and re-rounding
            # after the 0.1 slope adds at most two fp16 ulps (2^-9 relative)
            assert (diff <= 2.0 ** -9 * want.abs() + 2e-5 * exact.abs().max()).all()
            assert (diff > 0).float().mean() < 1e-2          # only fp16 rounding-boundary cases may differ
        else:
            assert (got - exact).abs().max() < 2e-5 * max(1.0, exact.abs().max().item())


def test_dcn_generic_geometry_vs_torchvision(dev):
    """Outside the TDVC configuration (stride 2, 5x3 kernel, dilation) the `_ext` drop-in still answers."""
    import torchvision.ops
    from tdvc_b200.ops import dcn_v2_forward
    torch.manual_seed(7)
    N, C, O, H, W, dg = 1, 4, 6, 11, 12, 2
    kh, kw, sh, sw, ph, pw, dh, dw = 5, 3, 2, 1, 2, 1, 1, 2
    Ho = (H + 2 * ph - (dh * (kh - 1) + 1)) // sh + 1
    Wo = (W + 2 * pw - (dw * (kw - 1) + 1)) // sw + 1
    x, wgt, b = torch.randn(N, C, H, W), torch.randn(O, C, kh, kw) * 0.2, torch.randn(O)
    off, msk = torch.randn(N, dg * 2 * kh * kw, Ho, Wo) * 2, torch.rand(N, dg * kh * kw, Ho, Wo)
    want = torchvision.ops.deform_conv2d(x, off, wgt, b, stride=(sh, sw), padding=(ph, pw), dilation=(dh, dw), mask=msk)
    got = dcn_v2_forward(x.to(dev), wgt.to(dev), b.to(dev), off.to(dev), msk.to(dev), kh, kw, sh, sw, ph, pw, dh, dw, dg).cpu()
    assert (got - want).abs().max() < 2e-5 * max(1.0, want.abs().max().item())


@pytest.mark.parametrize("shape", [(2, 64, 64, 13, 17, 8, 3, 3, 1, 1, 1, 1, 1, 1), (1, 16, 24, 9, 7, 2, 3, 3, 1, 1, 1, 1, 1, 1),
                                   (2, 4, 6, 11, 12, 2, 5, 3, 2, 1, 2, 1, 1, 2),
                                   (1, 24, 70, 9, 7, 2, 3, 3, 1, 1, 1, 1, 1, 1)])   # 12 channels per group, > 64 outputs
def test_dcn_backward_vs_torchvision_autograd(dev, shape):
    """`dcn_v2_backward` (reference dcn_v2.h:48-92) against autograd through torchvision.ops.deform_conv2d on the CPU (the same
    MXNet-lineage arithmetic): all five gradients, offsets with many samples outside the image, TDVC's geometry
    and a generic one; and it is deterministic (the reference's col2im is not: float atomicAdd)."""
    import torchvision.ops
    from tdvc_b200.ops import dcn_v2_backward
    N, C, O, H, W, dg, kh, kw, sh, sw, ph, pw, dh, dw = shape
    torch.manual_seed(13)
    Ho = (H + 2 * ph - (dh * (kh - 1) + 1)) // sh + 1
    Wo = (W + 2 * pw - (dw * (kw - 1) + 1)) // sw + 1
    x = torch.randn(N, C, H, W, requires_grad=True)
    wgt = (torch.randn(O, C, kh, kw) * 0.1).requires_grad_()
    b = torch.randn(O, requires_grad=True)
    off = (torch.randn(N, dg * 2 * kh * kw, Ho, Wo) * 3.0).requires_grad_()
    msk = torch.rand(N, dg * kh * kw, Ho, Wo, requires_grad=True)
    go = torch.randn(N, O, Ho, Wo)
    out = torchvision.ops.deform_conv2d(x, off, wgt, b, stride=(sh, sw), padding=(ph, pw), dilation=(dh, dw), mask=msk)
    out.backward(go)
    args = [t.detach().to(dev) for t in (x, wgt, b, off, msk, go)] + [kh, kw, sh, sw, ph, pw, dh, dw, dg]
    got = dcn_v2_backward(*args)
    again = dcn_v2_backward(*args)
    for name, g, w, g2 in zip(("input", "offset", "mask", "weight", "bias"), got, (x.grad, off.grad, msk.grad, wgt.grad, b.grad), again):
        assert g.shape == w.shape, name
        assert (g.cpu() - w).abs().max().item() <= 2e-4 * max(1.0, w.abs().max().item()), name
        assert torch.equal(g, g2), f"grad_{name} is not deterministic"


def test_dcn_gradcheck_like_reference(dev):
    """reference main/utils/dcnv2/testcuda.py:73-101 (check_gradient_dconv): torch.autograd.gradcheck of the autograd function
    built on the two native ops, same sizes and tolerances."""
    from torch.autograd import gradcheck
    from tdvc_b200.ops import dcn_v2_conv
    torch.manual_seed(2)
    N, inC, inH, inW, outC, kH, kW, dg = 2, 2, 4, 4, 2, 3, 3, 1
    x = (torch.rand(N, inC, inH, inW, device=dev) * 0.01).requires_grad_()
    off = (torch.randn(N, dg * 2 * kW * kH, inH, inW, device=dev) * 2).requires_grad_()
    msk = torch.sigmoid(torch.rand(N, dg * kW * kH, inH, inW, device=dev)).requires_grad_()
    wgt = torch.randn(outC, inC, kH, kW, device=dev).requires_grad_()
    b = torch.rand(outC, device=dev).requires_grad_()
    assert gradcheck(dcn_v2_conv, (x, off, msk, wgt, b, 1, 1, 1, dg), eps=1e-3, atol=1e-4, rtol=1e-2, nondet_tol=0.0)


def test_dcn_errors_like_reference(dev):
    from tdvc_b200.ops import dcn_v2_forward
    x = torch.randn(1, 8, 4, 4, device=dev)
    w = torch.randn(8, 8, 3, 3, device=dev)
    with pytest.raises(RuntimeError):
        dcn_v2_forward(x.cpu(), w, torch.zeros(8, device=dev), torch.zeros(1, 18, 4, 4, device=dev),
                       torch.zeros(1, 9, 4, 4, device=dev), 3, 3, 1, 1, 1, 1, 1, 1, 1)
    with pytest.raises(RuntimeError):  # kernel size does not match the weight (reference AT_ASSERTM)
        dcn_v2_forward(x, w, torch.zeros(8, device=dev), torch.zeros(1, 18, 4, 4, device=dev),
                       torch.zeros(1, 9, 4, 4, device=dev), 5, 5, 1, 1, 1, 1, 1, 1, 1)


# ------------------------------------------------------------------------------------------------ SPyNet pieces
def test_spynet_prep_vs_oracle(plan, dev):
    """x2 flow upsample (align_corners=True, *2) + border warp with the reference's coordinate round trip
    (reference flownet.py:8-48,124-138)."""
    from oracle.model import flow_warp_border
    from tdvc_b200.model import Act
    torch.manual_seed(2)
    # several 64x8 tiles per image, ragged tile edges, batch; flows of sigma 6 (2 x 3) put most samples inside the staged
    # shared-memory tile (margin 8) and a good fraction outside it (the global-memory path), sigma 30 nearly all outside
    for (h, w, sig) in ((8, 10, 3.0), (32, 60, 3.0), (64, 64, 3.0), (70, 200, 3.0), (34, 150, 15.0)):
        ref, supp = torch.rand(1, 3, h, w), torch.rand(1, 3, h, w)
        flow = torch.randn(1, 2, h // 2, w // 2) * sig
        up = F.interpolate(flow, scale_factor=2, mode="bilinear", align_corners=True) * 2.0
        warped = flow_warp_border(supp, up.permute(0, 2, 3, 1))
        want = torch.cat([ref, warped, up], 1)
        r4, s4 = Act.from_nchw(ref.to(dev), ld=4), Act.from_nchw(supp.to(dev), ld=4)
        fl = Act.from_nchw(flow.to(dev))
        out = Act.alloc(1, h, w, 8, dev)
        plan.call("tdvc_spynet_prep", r4.ptr, s4.ptr, fl.ptr, out.ptr, 1, h, w)
        got = out.nchw().cpu()
        assert (got - want).abs().max() < 3e-6 * max(1.0, want.abs().max().item())
        plan.call("tdvc_spynet_prep", r4.ptr, s4.ptr, None, out.ptr, 1, h, w)  # level 0: zero flow => identity warp
        got0 = out.nchw().cpu()
        want0 = flow_warp_border(supp, torch.zeros(1, h, w, 2))  # the coordinate round trip is not exact
        assert (got0[:, 3:6] - want0).abs().max() < 3e-6 and got0[:, 6:].abs().max() == 0


def test_pool_upsample_flowadd(plan, dev):
    from tdvc_b200.model import Act
    torch.manual_seed(4)
    x = torch.rand(2, 3, 12, 20)
    a = Act.from_nchw(x.to(dev), ld=4)
    o = Act.alloc(2, 6, 10, 3, dev, ld=4)
    plan.call("tdvc_avgpool2x2", a.ptr, o.ptr, 2, 12, 20, 4)
    assert (o.nchw().cpu() - F.avg_pool2d(x, 2, 2)).abs().max() < 1e-7
    y = torch.randn(1, 64, 6, 9)
    u = Act.alloc(1, 12, 18, 64, dev)
    plan.call("tdvc_upsample2x", Act.from_nchw(y.to(dev)).ptr, u.ptr, 1, 6, 9, 64)
    want = F.interpolate(y, scale_factor=2, mode="bilinear", align_corners=False)
    assert (u.nchw().cpu() - want).abs().max() < 1e-6
    off, fl = torch.randn(1, 64, 5, 7), torch.randn(1, 2, 5, 7)
    o2 = Act.alloc(1, 5, 7, 64, dev)
    plan.call("tdvc_add_flow_tiled", Act.from_nchw(off.to(dev)).ptr, Act.from_nchw(fl.to(dev)).ptr, o2.ptr, 1, 5, 7, 64)
    assert torch.equal(o2.nchw().cpu(), off + fl.repeat(1, 32, 1, 1))


def test_layout_roundtrip(plan, dev):
    x = torch.randn(2, 3, 37, 41, device=dev)
    nhwc = torch.empty(2, 37, 41, 4, device=dev)
    back = torch.empty_like(x)
    plan.call("tdvc_nchw_to_nhwc", x.data_ptr(), nhwc.data_ptr(), 2, 3, 37, 41, 4)
    plan.call("tdvc_nhwc_to_nchw", nhwc.data_ptr(), 4, back.data_ptr(), 2, 3, 37, 41)
    assert torch.equal(back, x) and nhwc[..., 3].abs().max() == 0 and torch.equal(nhwc[..., :3], x.permute(0, 2, 3, 1))


# ------------------------------------------------------------------------------------------------ SE, element-wise
def test_se_layer_vs_oracle(plan, dev):
    from oracle.model import SELayer
    from tdvc_b200 import lib as L
    from tdvc_b200.model import Act
    torch.manual_seed(6)
    for C, H, W in ((64, 24, 40), (128, 5, 7)):
        se = SELayer(C)
        x, r = torch.randn(2, C, H, W), torch.randn(2, C, H, W)
        want = F.leaky_relu(se(x), 0.1) + r
        w = (se.conv1.conv.weight.reshape(C // 16, C).to(dev).contiguous(), se.conv1.conv.bias.to(dev),
             se.conv2.conv.weight.reshape(C, C // 16).to(dev).contiguous(), se.conv2.conv.bias.to(dev))
        out = Act.alloc(2, H, W, C, dev)
        plan.se(Act.from_nchw(x.to(dev)), w, out, act=L.ACT_LRELU, slope=0.1, res=Act.from_nchw(r.to(dev)))
        assert (out.nchw().cpu() - want.detach()).abs().max() < 1e-5


def test_elementwise(plan, dev):
    torch.manual_seed(8)
    a, b = torch.randn(1003, device=dev), torch.randn(1003, device=dev)
    o = torch.empty_like(a)
    plan.call("tdvc_axpby", a.data_ptr(), b.data_ptr(), o.data_ptr(), 1003, 1.0, -1.0)
    assert torch.equal(o, a - b)
    y = torch.tensor([0.5, 1.5, 2.5, -0.5, -1.5, 2.4999, -3.5001, 7.0], device=dev)
    o = torch.empty_like(y)
    plan.call("tdvc_round_half_even", y.data_ptr(), o.data_ptr(), 8)
    assert torch.equal(o, torch.round(y)) and o.tolist() == [0.0, 2.0, 2.0, -0.0, -2.0, 2.0, -4.0, 7.0]
    x, t = torch.randn(4, 640, device=dev), torch.randn(640, device=dev)
    o = torch.empty_like(x)
    plan.call("tdvc_bcast_add_lrelu", x.data_ptr(), t.data_ptr(), o.data_ptr(), 4, 640, 0.1)
    assert torch.equal(o, F.leaky_relu(x + t, 0.1))
    acc = torch.zeros(1, device=dev, dtype=torch.float64)
    plan.call("tdvc_sq_err_sum", a.data_ptr(), b.data_ptr(), 1003, acc.data_ptr())
    assert abs(acc.item() - ((a - b).double() ** 2).sum().item()) < 1e-9 * acc.item()


# ------------------------------------------------------------------------------------------------ entropy models
def test_entropy_bottleneck_bits_vs_oracle(plan, dev):
    from oracle.compressai_port import EntropyBottleneck
    torch.manual_seed(9)
    eb = EntropyBottleneck(128).eval()
    eb.quantiles.data[:, :, 1].add_(torch.randn(128, 1) * 0.3)
    for i in range(4):
        getattr(eb, f"_factor{i}").data.copy_(torch.randn(128, 3, 1) * 0.3)
    z = torch.randn(1, 128, 6, 10) * 4
    with torch.no_grad():
        zh, lik = eb(z)
    C = 128
    mats = torch.cat([F.softplus(getattr(eb, f"_matrix{i}").detach()).reshape(C, -1) for i in range(5)], 1).to(dev).contiguous()
    biases = torch.cat([getattr(eb, f"_bias{i}").detach().reshape(C, -1) for i in range(5)], 1).to(dev).contiguous()
    factors = torch.cat([torch.tanh(getattr(eb, f"_factor{i}").detach()).reshape(C, -1) for i in range(4)], 1).to(dev).contiguous()
    med = eb.quantiles.detach()[:, 0, 1].to(dev).contiguous()
    from tdvc_b200.model import Act
    za = Act.from_nchw(z.to(dev))
    zo = Act.alloc(1, 6, 10, 128, dev)
    acc = torch.zeros(1, device=dev, dtype=torch.float64)
    plan.call("tdvc_eb_bits", za.ptr, zo.ptr, mats.data_ptr(), biases.data_ptr(), factors.data_ptr(), med.data_ptr(),
              60, 128, acc.data_ptr())
    assert torch.equal(zo.nchw().cpu(), zh)  # quantised symbols bit-exact
    want = torch.log(lik).double().sum().item()
    assert abs(acc.item() - want) < 1e-4 * abs(want)


def test_gaussian_conditional_bits_vs_oracle(plan, dev):
    from oracle.compressai_port import GaussianConditional
    from tdvc_b200.model import Act
    torch.manual_seed(10)
    gc = GaussianConditional(None).eval()
    y = torch.randn(1, 128, 7, 9) * 5
    scales = torch.exp(torch.randn(1, 128, 7, 9) * 1.5) * 0.3  # crosses the 0.11 floor
    means = torch.randn(1, 128, 7, 9)
    with torch.no_grad():
        _, lik = gc(y, scales, means=means)
    gp = Act.from_nchw(torch.cat([scales, means], 1).to(dev))
    acc = torch.zeros(1, device=dev, dtype=torch.float64)
    plan.call("tdvc_gc_bits", Act.from_nchw(y.to(dev)).ptr, gp.ptr, gp.ld, 63, 128, acc.data_ptr())
    want = torch.log(lik).double().sum().item()
    assert (scales < 0.11).any() and (lik <= 1e-9).any()  # exercises both bounds
    assert abs(acc.item() - want) < 1e-4 * abs(want)


# ------------------------------------------------------------------------------------------------ in-loop filter pieces
@pytest.mark.parametrize("hw", [(64, 64), (128, 192), (64, 320)])
def test_featurefix_match_and_gather_vs_oracle(plan, dev, hw):
    """avg-pool -> unfold/normalise -> similarity/argmax -> block placement -> cosine gate
    (reference pnet.py:219-257) against the oracle's unfold/bmm/gather/fold restatement."""
    from tdvc_b200.model import Act
    H, W = hw
    torch.manual_seed(11)
    f_in, f_ref = torch.randn(1, 64, H, W), torch.randn(1, 64, H, W)
    scale = int(H / 8)
    p_in, p_ref = F.avg_pool2d(f_in, scale, scale), F.avg_pool2d(f_ref, scale, scale)
    q = F.unfold(p_in, 3, padding=3, stride=3).transpose(2, 1)
    k = F.unfold(p_ref, 3, padding=3, stride=3).transpose(2, 1).reshape(1, -1, 576)
    sim = torch.bmm(F.normalize(q, dim=2), F.normalize(k.transpose(2, 1), dim=1))
    ind = sim.max(dim=2)[1]
    bs = 3 * scale
    blocks = F.unfold(f_ref, bs, padding=bs, stride=bs).transpose(2, 1).reshape(1, -1, 64 * bs * bs)
    idx = ind.view(1, 1, -1).expand(-1, 64 * bs * bs, -1).permute(0, 2, 1)
    picked = torch.gather(blocks, 1, idx).view(1, -1, 64, bs, bs).permute(0, 2, 3, 4, 1).reshape(1, -1, q.size(1))
    gathered = F.fold(picked, (H, W), bs, padding=bs, stride=bs)
    cor = torch.cosine_similarity(f_in, gathered).unsqueeze(1)

    fa, fr = Act.from_nchw(f_in.to(dev)), Act.from_nchw(f_ref.to(dev))
    ph, pw = H // scale, W // scale
    P = q.size(1)
    pin, pref = torch.empty(1, ph, pw, 64, device=dev), torch.empty(1, ph, pw, 64, device=dev)
    plan.call("tdvc_avgpool_scale", fa.ptr, 64, pin.data_ptr(), 1, H, W, 64, scale)
    plan.call("tdvc_avgpool_scale", fr.ptr, 64, pref.data_ptr(), 1, H, W, 64, scale)
    assert (pin.permute(0, 3, 1, 2).cpu() - p_in).abs().max() < 1e-6
    din, dref = torch.empty(1, P, 576, device=dev), torch.empty(1, P, 576, device=dev)
    plan.call("tdvc_ff_descriptors", pin.data_ptr(), din.data_ptr(), 1, ph, pw, 64)
    plan.call("tdvc_ff_descriptors", pref.data_ptr(), dref.data_ptr(), 1, ph, pw, 64)
    gi = torch.empty(1, P, dtype=torch.int32, device=dev)
    gs = torch.empty(1, P, P, device=dev)
    plan.call("tdvc_ff_match", din.data_ptr(), dref.data_ptr(), gi.data_ptr(), gs.data_ptr(), 1, P, 576)
    assert (gs.cpu() - sim).abs().max() < 1e-5
    assert torch.equal(gi.cpu().long(), ind)  # indices bit-exact (incl. first-index tie-break on all-zero patches)
    oa, ob = Act.alloc(1, H, W, 64, dev), Act.alloc(1, H, W, 64, dev)
    og, oc = Act.alloc(1, H, W, 64, dev), torch.empty(1, 1, H, W, device=dev)
    plan.call("tdvc_ff_gather", fa.ptr, fr.ptr, gi.data_ptr(), oa.ptr, ob.ptr, og.ptr, oc.data_ptr(), 1, H, W, 64, scale)
    assert torch.equal(og.nchw().cpu(), gathered)  # pure data movement: bit-exact
    assert (oc.cpu() - cor).abs().max() < 1e-5
    assert (oa.nchw().cpu() - f_in * cor).abs().max() < 1e-4 and (ob.nchw().cpu() - gathered * cor).abs().max() < 1e-4


# ------------------------------------------------------------------------------------------------ MS-SSIM
def test_msssim_vs_oracle_and_reference_golden(dev):
    """tdvc_b200.metrics.ms_ssim (fused CUDA scale kernel) against oracle/ms_ssim.py and against the committed values of the
    reference's own function (tests/golden/msssim.json).  fp32 filter sums in a different order: 2e-5 absolute."""
    import json
    import os
    from conftest import ROOT
    from oracle import ms_ssim as O
    from oracle.make_golden import msssim_inputs
    from tdvc_b200.metrics import ms_ssim
    for c in json.load(open(os.path.join(ROOT, "tests", "golden", "msssim.json"))):
        x, y = msssim_inputs(c["seed"], tuple(c["shape"]), c["noise"])
        got = ms_ssim(x.to(dev), y.to(dev), data_range=1.0, size_average=False).cpu()
        assert torch.allclose(got, torch.tensor(c["per_image"]), atol=2e-5, rtol=0), (got, c["per_image"])
        assert torch.allclose(got, O.ms_ssim(x, y, data_range=1.0, size_average=False), atol=2e-5, rtol=0)
        assert abs(float(ms_ssim(x.to(dev), y.to(dev), data_range=1.0)) - c["mean"]) < 2e-5
    torch.manual_seed(9)
    x = torch.rand(1, 3, 163, 177)      # odd sizes: padded pooling between the scales
    y = (x + 0.1 * torch.randn_like(x)).clamp(0, 1)
    assert abs(float(ms_ssim(x.to(dev), y.to(dev), data_range=1.0)) - float(O.ms_ssim(x, y, data_range=1.0))) < 2e-5


def test_msssim_full_size_properties(dev):
    """1920x1024 (BASELINE config 2): identity gives 1, symmetry, monotone in the distortion, errors like the reference."""
    from tdvc_b200.metrics import ms_ssim
    torch.manual_seed(1)
    x = torch.rand(1, 3, 1024, 1920, device=dev)
    x = F.avg_pool2d(x, 5, 1, 2)
    assert abs(float(ms_ssim(x, x, data_range=1.0)) - 1.0) < 1e-6
    n = torch.randn_like(x)
    a = float(ms_ssim(x, (x + 0.02 * n).clamp(0, 1), data_range=1.0))
    b = float(ms_ssim(x, (x + 0.08 * n).clamp(0, 1), data_range=1.0))
    assert 0.0 < b < a < 1.0
    y = (x + 0.05 * n).clamp(0, 1)
    assert abs(float(ms_ssim(x, y, data_range=1.0)) - float(ms_ssim(y, x, data_range=1.0))) < 1e-6
    with pytest.raises(ValueError):
        ms_ssim(x, x[:, :, :512], data_range=1.0)
    with pytest.raises(RuntimeError):
        ms_ssim(x.cpu(), x.cpu(), data_range=1.0)
