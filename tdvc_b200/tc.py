"""Host side of the tcgen05 (fp16 hi/lo split) convolution: asks the C-ABI which packed convolutions have a
tensor-core path and lets it build their fp16 (hi, lo) weight blocks on the device (csrc/conv_tc.cu)."""
import math

import torch

from tdvc_b200 import lib as L


def _geometry(cw, products):
    p = L.ConvParams()
    p.kh = p.kw = cw.k
    p.stride, p.pad = getattr(cw, "stride", 1), cw.pad
    p.cin, p.cin_pad, p.cout, p.cout_pad = cw.cin, cw.cin_pad, cw.cout, cw.cout_pad
    p.weight = cw.w.data_ptr()
    p.products = products
    return p


def _pack(lib, cw, products):
    """fp16 blocks of one convolution for one scheme -> (uint8 device tensor or None, w_shift)."""
    p = _geometry(cw, products)
    nb = lib.tdvc_conv2d_f16_bytes(p)
    if nb == 0:
        return None, 0
    shift = 0
    if lib.tdvc_conv2d_f16_is_split(p):
        # weights are stored * 2^w_shift with max|w| * 2^w_shift in [2^13, 2^14): w_lo = w - fp16(w) (split scheme) and the small
        # weights of a one-product layer stay in the fp16 normal range for every weight above 2^-16 of the largest one
        m = getattr(cw, "wmax", None)
        m = m() if callable(m) else (m if m is not None else float(cw.w.abs().max()))
        shift = 13 - math.frexp(m)[1] + 1 if m > 0 else 0
    p.w_shift = shift
    with torch.cuda.device(cw.w.device):
        buf = torch.empty(nb, device=cw.w.device, dtype=torch.uint8)
        L.check(lib.tdvc_conv2d_pack_f16(p, buf.data_ptr(), torch.cuda.current_stream(cw.w.device).cuda_stream),
                "conv2d_pack_f16")
    return buf, shift


def attach_f16(packed, one_product=None):
    """packed: dict name -> ConvW (other value types are skipped).  Sets ConvW.w_f16 / w_shift (fp32-class split scheme) where the
    shape has a tcgen05 path, and ConvW.w_f16_p1 / w_shift_p1 (one fp16 product) for the layers flagged `p1_ok`
    (`one_product=True` forces it for every layer: kernel tests)."""
    lib = L.load()
    for cw in packed.values():
        if not hasattr(cw, "cout_pad") or cw.w is None or not cw.w.is_cuda:
            continue
        cw.w_f16, cw.w_shift = _pack(lib, cw, 0)
        if one_product or (one_product is None and getattr(cw, "p1_ok", False)):
            cw.w_f16_p1, cw.w_shift_p1 = _pack(lib, cw, 1)
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
