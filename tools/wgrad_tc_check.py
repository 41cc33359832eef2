"""Developer tool: the tcgen05 weight-gradient kernel (csrc/wgrad_tc.cu) against torch float64 and against the warp-level kernel,
with timings: python tools/wgrad_tc_check.py   (TDVC_B200_WGRAD_SIMT=1 / TDVC_B200_WGRAD_NOBASE=1 switch paths)."""
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
from tdvc_b200 import lib as L

lib = L.load()
dev = torch.device("cuda:0")


def wgrad(x, g, k, products):
    N, C, H, W = x.shape
    O = g.shape[1]
    xa = x.permute(0, 2, 3, 1).contiguous()
    ga = g.permute(0, 2, 3, 1).contiguous()
    gw = torch.empty(O, C, k, k, device=dev)
    gb = torch.empty(O, device=dev)
    nb = lib.tdvc_conv2d_wgrad_workspace_bytes(N, g.shape[2], g.shape[3], C, O, k)
    ws = torch.empty((nb + 3) // 4, device=dev)
    st = torch.cuda.current_stream().cuda_stream

    def run():
        L.check(lib.tdvc_conv2d_wgrad(xa.data_ptr(), C, ga.data_ptr(), O, N, H, W, C, O, k, 1, k // 2, 0, products, gw.data_ptr(),
                                      gb.data_ptr(), ws.data_ptr(), nb, st), "wgrad")
    run()
    torch.cuda.synchronize()
    return gw, gb, run


def check(N, C, O, H, W, seed=0, timing=False):
    g_ = torch.Generator().manual_seed(seed)
    x = torch.randn(N, C, H, W, generator=g_).to(dev)
    gy = torch.randn(N, O, H, W, generator=g_).to(dev)
    gw, gb, run = wgrad(x, gy, 3, 1)
    xd = x.double()
    wd = torch.zeros(O, C, 3, 3, device=dev, dtype=torch.float64, requires_grad=True)
    F.conv2d(xd, wd, None, 1, 1).backward(gy.double())
    ref = wd.grad
    scale = ref.abs().max().item()
    err = (gw.double() - ref).abs()
    per_tap = err.amax(dim=(0, 1)) / scale
    berr = (gb.double() - gy.double().sum((0, 2, 3))).abs().max().item() / gy.double().sum((0, 2, 3)).abs().max().item()
    msg = f"N{N} {C}->{O} {H}x{W}: max rel err {err.max().item() / scale:.2e} bias {berr:.2e} per tap {[f'{v:.1e}' for v in per_tap.flatten().tolist()]}"
    if timing:
        for _ in range(3):
            run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        msg += f" | {ms:.3f} ms = {2 * N * H * W * 9 * C * O / ms / 1e9:.1f} TFLOP/s"
    print(msg, flush=True)


check(1, 64, 64, 4, 32)
check(2, 64, 64, 8, 64, seed=1)
check(1, 128, 64, 16, 32, seed=2)
check(2, 128, 192, 12, 96, seed=3)
check(8, 64, 64, 256, 256, seed=4, timing=True)
check(8, 128, 128, 128, 128, seed=5, timing=True)
check(8, 128, 64, 256, 256, seed=6, timing=True)
check(8, 128, 512, 64, 64, seed=7, timing=True)
