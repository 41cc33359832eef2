"""Host side of the tcgen05 (fp16 hi/lo split) convolution: asks the C-ABI which packed convolutions have a
tensor-core path and lets it build their fp16 (hi, lo) weight blocks on the device (csrc/conv_tc.cu)."""
import math

import torch

from tdvc_b200 import lib as L


def attach_f16(packed):
    """packed: dict name -> ConvW (other value types are skipped).  Sets ConvW.w_f16 where supported."""
    lib = L.load()
    for cw in packed.values():
        if not hasattr(cw, "cout_pad") or cw.w is None or not cw.w.is_cuda:
            continue
        p = L.ConvParams()
        p.kh = p.kw = cw.k
        p.stride, p.pad = getattr(cw, "stride", 1), cw.pad
        p.cin, p.cin_pad, p.cout, p.cout_pad = cw.cin, cw.cin_pad, cw.cout, cw.cout_pad
        p.weight = cw.w.data_ptr()
        nb = lib.tdvc_conv2d_f16_bytes(p)
        if nb == 0:
            continue
        cw.w_shift = 0
        if lib.tdvc_conv2d_f16_is_split(p):
            # 3-product split scheme: weights are stored * 2^w_shift with max|w| * 2^w_shift in [2^13, 2^14), so that
            # w_lo = w - fp16(w) stays in the fp16 normal range for every weight above 2^-16 of the largest one
            m = float(cw.w.abs().max())
            cw.w_shift = 13 - math.frexp(m)[1] + 1 if m > 0 else 0
        p.w_shift = cw.w_shift
        with torch.cuda.device(cw.w.device):
            buf = torch.empty(nb, device=cw.w.device, dtype=torch.uint8)
            L.check(lib.tdvc_conv2d_pack_f16(p, buf.data_ptr(), torch.cuda.current_stream(cw.w.device).cuda_stream),
                    "conv2d_pack_f16")
        cw.w_f16 = buf
    return packed


def attach_dcn_f16(packed, key, O, dg):
    """fp16 (hi | lo) weight blocks of the tcgen05 DCN (csrc/dcn_tc.cu) from packed[key] ([C*9][O_pad] fp32)."""
    lib = L.load()
    w = packed[key]
    if not w.is_cuda:
        return packed
    with torch.cuda.device(w.device):
        buf = torch.empty(lib.tdvc_dcn_f16_bytes(dg), device=w.device, dtype=torch.uint8)
        L.check(lib.tdvc_dcn_pack_f16(w.data_ptr(), O, w.shape[1], dg, buf.data_ptr(),
                                      torch.cuda.current_stream(w.device).cuda_stream), "dcn_pack_f16")
    packed[key + "_f16"] = buf
    return packed
