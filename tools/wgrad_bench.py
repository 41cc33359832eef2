"""Developer tool: time / profile tdvc_conv2d_wgrad on one layer shape: python tools/wgrad_bench.py [N C O k H W]"""
import sys

import torch

sys.path.insert(0, ".")
from tdvc_b200 import lib as L

N, C, O, k, H, W = (int(a) for a in sys.argv[1:7]) if len(sys.argv) > 6 else (8, 64, 64, 3, 256, 256)
PR = int(sys.argv[7]) if len(sys.argv) > 7 else 3
dev = torch.device("cuda:0")
lib = L.load()
x = torch.randn(N, H, W, C, device=dev)
g = torch.randn(N, H, W, O, device=dev)
gw = torch.empty(O, C, k, k, device=dev)
gb = torch.empty(O, device=dev)
nb = lib.tdvc_conv2d_wgrad_workspace_bytes(N, H, W, C, O, k)
ws = torch.empty((nb + 3) // 4, device=dev)
st = torch.cuda.current_stream(dev).cuda_stream
for rep in range(4):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    L.check(lib.tdvc_conv2d_wgrad(x.data_ptr(), C, g.data_ptr(), O, N, H, W, C, O, k, 1, k // 2, 0, PR, gw.data_ptr(), gb.data_ptr(),
                                  ws.data_ptr(), nb, st), "wgrad")
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    print(f"wgrad {C}->{O} k{k} @ {N}x{H}x{W}: {ms:.3f} ms = {2e-9 * N * H * W * C * O * k * k / ms:.1f} TFLOP/s (algorithmic)")
ref = torch.nn.grad.conv2d_weight(x.permute(0, 3, 1, 2).double(), (O, C, k, k), g.permute(0, 3, 1, 2).double(), padding=k // 2)
print("max rel err vs fp64", ((gw.double() - ref).abs().max() / ref.abs().max()).item())
