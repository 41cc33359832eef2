"""Developer tool (GPU box): time one convolution shape through the C-ABI (CUDA events, L2-exceeding tensors)."""
import sys
import torch
sys.path.insert(0, ".")
from tdvc_b200.model import Act, _Plan, pack_conv
from tdvc_b200 import tc


def main(cin=64, cout=64, k=3, H=1024, W=1920, impl=2, iters=10, N=1, stride=1, planar=0, act=1, res=0, gdn=0, products=0):
    dev = torch.device("cuda:0")
    plan = _Plan(1, 64, 64, dev)
    torch.manual_seed(0)
    conv = torch.nn.Conv2d(cin, cout, k, stride, k // 2).to(dev)
    cw = pack_conv(conv.weight, conv.bias, src_layout=[(cin, (cin + 3) // 4 * 4)], stride=stride)
    tc.attach_f16({"w": cw}, one_product=products == 1)
    x = Act.alloc(N, H, W, cin, dev, ld=(cin + 3) // 4 * 4)
    x.t.normal_()
    Ho, Wo = (H + 2 * (k // 2) - k) // stride + 1, (W + 2 * (k // 2) - k) // stride + 1
    if planar:
        outs = [torch.empty(N, cout, Ho, Wo, device=dev) for _ in range(2)]
    else:
        outs = [Act.alloc(N, Ho, Wo, cout, dev, ld=(cout + 3) // 4 * 4) for _ in range(2)]
    r = Act.alloc(N, Ho, Wo, cout, dev, ld=(cout + 3) // 4 * 4) if res else None
    if r is not None:
        r.t.normal_()
    kw = dict(stride=stride, act=act, impl=impl, planar=bool(planar), res1=r, products=products)
    if gdn:   # GDN / IGDN as the fused 1x1: norm = conv(x^2), out = x * rsqrt(norm) | x * sqrt(norm)
        from tdvc_b200 import lib as L
        x.t.abs_()
        conv.weight.data.abs_()
        conv.bias.data.abs_().add_(1.0)
        cw = pack_conv(conv.weight, conv.bias, src_layout=[(cin, cin)], stride=stride)
        tc.attach_f16({"w": cw}, one_product=products == 1)
        kw.update(in_square=True, post=L.POST_GDN if gdn == 1 else L.POST_IGDN, mul=x)
    for i in range(3):
        plan.conv([x], cw, outs[i % 2], **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        plan.conv([x], cw, outs[i % 2], **kw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    macs = N * Ho * Wo * cin * cout * k * k
    print(f"conv{k}x{k}s{stride} {cin}->{cout} @{H}x{W} N{N} impl {impl} planar {planar} act {act} res {res} gdn {gdn} products {products}: {ms:.3f} ms  {2 * macs / ms / 1e9:.1f} TFLOP/s algorithmic "
          f"({6 * macs / ms / 1e9:.1f} MMA-equivalent), {(N * H * W * cin + N * Ho * Wo * cout) * 4 / ms / 1e6:.0f} GB/s")


if __name__ == "__main__":
    main(*[int(a) for a in sys.argv[1:]])
