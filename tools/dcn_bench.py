"""Developer tool (GPU box): time the tcgen05 DCNv2 kernel alone at 1920x1024 (offsets ~ N(0, sigma) px per pixel and tap)."""
import sys
import torch
sys.path.insert(0, ".")
from tdvc_b200 import lib as L, tc
from tdvc_b200.model import Act


def main(H=1024, W=1920, sigma=2.0, iters=10):
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    lib = L.load()
    C, dg, O = 64, 8, 64
    st = torch.cuda.current_stream(dev).cuda_stream
    gp = torch.randn(dg, H, W, 8, device=dev)
    om = torch.cat([torch.randn(1, 144, H, W, device=dev) * float(sigma), torch.randn(1, 72, H, W, device=dev)], 1).contiguous()
    wgt = torch.randn(O, C, 3, 3, device=dev) * 0.1
    pk = torch.zeros(C * 9, 64, device=dev)
    pk[:, :O] = wgt.reshape(O, C * 9).t()
    packed = tc.attach_dcn_f16({"w": pk.contiguous()}, "w", O, dg)
    bias = torch.randn(O, device=dev)
    out = Act.alloc(1, H, W, O, dev)
    dp = L.DcnParams()
    dp.input_gp, dp.params_planar = gp.data_ptr(), 1
    dp.offset, dp.off_ld = om.data_ptr(), 216
    dp.mask, dp.mask_ld, dp.mask_is_logit = om.data_ptr() + 4 * 144 * H * W, 216, 1
    dp.weight_packed, dp.bias, dp.weight_f16 = packed["w"].data_ptr(), bias.data_ptr(), packed["w_f16"].data_ptr()
    dp.out, dp.out_ld = out.ptr, out.ld
    dp.N, dp.H, dp.W, dp.C, dp.O, dp.O_pad, dp.dg = 1, H, W, C, O, 64, dg
    dp.round_fp16, dp.act, dp.slope, dp.impl = 1, L.ACT_LRELU, 0.1, 2
    for _ in range(3):
        L.check(lib.tdvc_dcn_nhwc(dp, st), "dcn")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        L.check(lib.tdvc_dcn_nhwc(dp, st), "dcn")
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"dcn_tc @{H}x{W} sigma {sigma}: {ms:.3f} ms, {H * W * 1376 / ms / 1e6:.0f} GB/s algorithmic, checksum {out.t.double().sum().item():.6e}")


if __name__ == "__main__":
    a = sys.argv[1:]
    main(*(int(a[0]), int(a[1]), float(a[2])) if len(a) >= 3 else ())
