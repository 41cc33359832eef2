"""Developer tool (GPU box): the TMA-fed and the LDG-fed producers of conv_tc must give identical bits."""
import os
import sys
import torch
sys.path.insert(0, ".")
from tdvc_b200.model import Act, _Plan, pack_conv
from tdvc_b200 import tc


def run(cin, cout, k, H, W, N=1, products=0, reps=8):
    dev = torch.device("cuda:0")
    plan = _Plan(1, 64, 64, dev)
    torch.manual_seed(0)
    conv = torch.nn.Conv2d(cin, cout, k, 1, k // 2).to(dev)
    cw = pack_conv(conv.weight, conv.bias, src_layout=[(cin, (cin + 3) // 4 * 4)])
    tc.attach_f16({"w": cw}, one_product=products == 1)
    x = Act.alloc(N, H, W, cin, dev, ld=(cin + 3) // 4 * 4, zero=True)
    x.t[..., :cin].normal_()
    outs = {}
    for mode in ("ldg", "tma"):
        os.environ.pop("TDVC_B200_CONV_LDG", None)
        os.environ.pop("TDVC_B200_CONV_TMA", None)
        os.environ["TDVC_B200_CONV_LDG" if mode == "ldg" else "TDVC_B200_CONV_TMA"] = "1"
        res = []
        for r in range(reps):
            out = Act.alloc(N, H, W, cout, dev, ld=(cout + 3) // 4 * 4, zero=True)
            plan.conv([x], cw, out, impl=2, products=products, act=1)
            torch.cuda.synchronize()
            res.append(out.t.clone())
        outs[mode] = res
    a = outs["ldg"][0]
    msg = []
    for mode in ("ldg", "tma"):
        for r, t in enumerate(outs[mode]):
            bad = (t != a)
            if bad.any():
                idx = bad.nonzero()
                ys, xs = idx[:, 1], idx[:, 2]
                msg.append(f"{mode}[{r}]: {int(bad.sum())} mismatches, max {float((t - a).abs().max()):.3e}, y {int(ys.min())}-{int(ys.max())} x {int(xs.min())}-{int(xs.max())}, "
                           f"first {idx[0].tolist()}, distinct y%32 {sorted(set((ys % 32).tolist()))[:12]} x%8 {sorted(set((xs % 8).tolist()))}")
    print(f"conv{k}x{k} {cin}->{cout} @{H}x{W} N{N} products {products}: " + ("IDENTICAL" if not msg else "; ".join(msg)), flush=True)


if __name__ == "__main__":
    run(64, 64, 3, 128, 192)
    run(64, 64, 3, 1024, 1920)
    run(64, 64, 3, 1024, 1920, products=1)
    run(128, 128, 3, 512, 960)
    run(128, 128, 3, 512, 960, products=1)
    run(4, 64, 3, 1024, 1920, N=2)
    run(64, 32, 7, 512, 960)
    run(32, 64, 7, 512, 960)
