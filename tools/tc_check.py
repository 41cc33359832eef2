"""Developer tool (GPU box): tcgen05 convolution vs the exact fp32 SIMT kernel on the same inputs."""
import sys
import time
import torch
sys.path.insert(0, ".")
from tdvc_b200.model import Act, _Plan, pack_conv
from tdvc_b200 import tc, lib as L


def case(plan, dev, cin_list, cout, k, H, W, N=1, act=0, slope=0.0, shuffle=0, res=False, seed=0):
    torch.manual_seed(seed)
    cin = sum(cin_list)
    conv = torch.nn.Conv2d(cin, cout, k, 1, k // 2).to(dev)
    xs = [torch.randn(N, c, H, W, device=dev) for c in cin_list]
    cw = pack_conv(conv.weight, conv.bias, src_layout=[(c, (c + 3) // 4 * 4) for c in cin_list], shuffle=shuffle)
    tc.attach_f16({"w": cw})
    assert cw.w_f16 is not None, "no tc path"
    srcs = [Act.from_nchw(x, ld=(x.shape[1] + 3) // 4 * 4) for x in xs]
    sh = 2 if shuffle == 2 else 1
    oc = cout // 4 if shuffle == 2 else cout
    ld = (oc + 3) // 4 * 4
    o1 = Act.alloc(N, H * sh, W * sh, oc, dev, ld=ld, zero=True)
    o2 = Act.alloc(N, H * sh, W * sh, oc, dev, ld=ld, zero=True)
    r = Act.from_nchw(torch.randn(N, oc, H * sh, W * sh, device=dev), ld=ld) if res else None
    plan.conv(srcs, cw, o1, act=act, slope=slope, res1=r, impl=1)
    torch.cuda.synchronize()
    t0 = time.time()
    plan.conv(srcs, cw, o2, act=act, slope=slope, res1=r, impl=2)
    torch.cuda.synchronize()
    a, b = o1.nchw(), o2.nchw()
    err = (a - b).abs().max().item()
    print(f"cin {cin_list} cout {cout} k{k} {H}x{W} N{N} shuffle {shuffle}: max|ref| {a.abs().max().item():.3f} "
          f"maxerr {err:.3e} rel {err / a.abs().max().item():.2e}  ({(time.time() - t0) * 1e3:.1f} ms)", flush=True)
    return err / a.abs().max().item()


def main():
    dev = torch.device("cuda:0")
    plan = _Plan(1, 64, 64, dev)
    worst = 0.0
    worst = max(worst, case(plan, dev, [64], 64, 3, 16, 16))
    worst = max(worst, case(plan, dev, [64], 64, 3, 32, 48, act=1))
    worst = max(worst, case(plan, dev, [64], 64, 3, 37, 53, act=2, slope=0.1, res=True))
    worst = max(worst, case(plan, dev, [64, 64], 64, 3, 40, 40))
    worst = max(worst, case(plan, dev, [128], 128, 3, 24, 56, N=2))
    worst = max(worst, case(plan, dev, [64], 216, 3, 32, 32))
    worst = max(worst, case(plan, dev, [128], 512, 3, 16, 24, shuffle=2, act=2, slope=0.01))
    worst = max(worst, case(plan, dev, [192], 768, 3, 8, 8, shuffle=2))
    worst = max(worst, case(plan, dev, [8], 32, 7, 32, 32, act=1))
    worst = max(worst, case(plan, dev, [32], 64, 7, 40, 24, act=1))
    worst = max(worst, case(plan, dev, [64], 32, 7, 32, 32, act=1))
    worst = max(worst, case(plan, dev, [32], 16, 7, 32, 32, act=1))
    worst = max(worst, case(plan, dev, [16], 2, 7, 32, 32, res=True))
    worst = max(worst, case(plan, dev, [64], 64, 3, 512, 960))
    print("worst rel err", worst)


if __name__ == "__main__":
    main()
