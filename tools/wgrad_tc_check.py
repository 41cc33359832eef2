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
    Cp, Op = (C + 3) // 4 * 4, (O + 3) // 4 * 4     # channels-last rows padded to 16 bytes (garbage in the padding on purpose)
    xa = torch.full((N, H, W, Cp), 7.0, device=dev)
    xa[..., :C] = x.permute(0, 2, 3, 1)
    ga = torch.full((N, g.shape[2], g.shape[3], Op), 7.0, device=dev)
    ga[..., :O] = g.permute(0, 2, 3, 1)
    gw = torch.empty(O, C, k, k, device=dev)
    gb = torch.empty(O, device=dev)
    nb = lib.tdvc_conv2d_wgrad_workspace_bytes(N, g.shape[2], g.shape[3], C, O, k)
    ws = torch.empty((nb + 3) // 4, device=dev)
    st = torch.cuda.current_stream().cuda_stream

    def run():
        L.check(lib.tdvc_conv2d_wgrad(xa.data_ptr(), Cp, ga.data_ptr(), Op, N, H, W, C, O, k, 1, k // 2, 0, products, gw.data_ptr(),
                                      gb.data_ptr(), ws.data_ptr(), nb, st), "wgrad")
    run()
    torch.cuda.synchronize()
    return gw, gb, run


def check(N, C, O, H, W, k=3, seed=0, timing=False, products=1):
    g_ = torch.Generator().manual_seed(seed)
    x = torch.randn(N, C, H, W, generator=g_).to(dev)
    gy = torch.randn(N, O, H, W, generator=g_).to(dev)
    Cp, Op = (C + 3) // 4 * 4, (O + 3) // 4 * 4
    gw, gb, run = wgrad(x, gy, k, products)
    xd = x.double()
    wd = torch.zeros(O, C, k, k, device=dev, dtype=torch.float64, requires_grad=True)
    F.conv2d(xd, wd, None, 1, k // 2).backward(gy.double())
    ref = wd.grad
    scale = ref.abs().max().item()
    err = (gw.double() - ref).abs()
    bref = gy.double().sum((0, 2, 3))
    berr = (gb.double() - bref).abs().max().item() / bref.abs().max().item()
    msg = f"N{N} {C}->{O} k{k} {H}x{W} products {products}: max rel err {err.max().item() / scale:.2e} bias {berr:.2e}"
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
        msg += f" | {ms:.3f} ms = {2 * N * H * W * k * k * C * O / ms / 1e9:.1f} TFLOP/s"
    print(msg, flush=True)


if "small" in sys.argv:   # every kernel instantiation once on a small shape (what a compute-sanitizer run walks)
    for pr in (1, 3):
        for sh in ((1, 64, 64, 6, 40, 3), (1, 3, 64, 5, 33, 3), (1, 64, 3, 5, 33, 3), (1, 24, 20, 5, 33, 3), (1, 8, 32, 9, 40, 7),
                   (1, 32, 64, 9, 40, 7), (1, 64, 32, 9, 40, 7), (1, 192, 64, 6, 40, 1)):
            check(*sh[:5], k=sh[5], seed=31, products=pr)
    sys.exit(0)
if "exact" in sys.argv:   # the fp32-class three-product mode
    for sh in ((2, 64, 64, 9, 50, 3), (2, 128, 192, 12, 96, 3), (2, 3, 64, 12, 40, 3), (2, 64, 216, 12, 40, 3), (1, 8, 32, 20, 40, 7),
               (1, 32, 64, 20, 40, 7), (1, 192, 64, 20, 40, 1)):
        check(*sh[:5], k=sh[5], seed=21, products=3)
    for sh in ((8, 64, 64, 256, 256, 3), (8, 128, 128, 128, 128, 3), (8, 8, 32, 256, 256, 7), (8, 64, 32, 256, 256, 7)):
        check(*sh[:5], k=sh[5], seed=22, timing=True, products=3)
    sys.exit(0)
if "big" in sys.argv:   # the profiled case (tools/gpu_prof_wgrad.sh): the dominant layer of the training step
    check(8, 64, 64, 256, 256, seed=4, timing=True)
    sys.exit(0)
check(1, 64, 64, 4, 32)
check(2, 64, 64, 9, 50, seed=1)
check(1, 128, 64, 16, 32, seed=2)
check(2, 128, 192, 12, 96, seed=3)
check(2, 3, 64, 12, 40, seed=8)
check(2, 64, 216, 12, 40, seed=9)
check(2, 64, 3, 12, 40, seed=10)
check(1, 8, 32, 20, 40, k=7, seed=11)
check(1, 32, 64, 20, 40, k=7, seed=12)
check(1, 64, 32, 20, 40, k=7, seed=13)
check(1, 32, 16, 20, 40, k=7, seed=14)
check(1, 16, 2, 20, 40, k=7, seed=15)
check(1, 192, 64, 20, 40, k=1, seed=16)
check(8, 64, 64, 256, 256, seed=4, timing=True)
check(8, 128, 128, 128, 128, seed=5, timing=True)
check(8, 64, 216, 256, 256, seed=6, timing=True)
check(8, 3, 64, 256, 256, seed=7, timing=True)
check(8, 64, 3, 256, 256, seed=7, timing=True)
check(8, 8, 32, 256, 256, k=7, seed=7, timing=True)
check(8, 32, 64, 256, 256, k=7, seed=7, timing=True)
check(8, 64, 32, 256, 256, k=7, seed=7, timing=True)
check(8, 16, 2, 256, 256, k=7, seed=7, timing=True)
check(8, 256, 64, 256, 256, k=1, seed=7, timing=True)
check(8, 128, 512, 16, 16, seed=7, timing=True)
