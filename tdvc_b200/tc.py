"""Host side of the tcgen05 (3xFP16 split) convolution: asks the C-ABI which packed convolutions have a
tensor-core path and lets it build their fp16 (hi, lo) weight blocks on the device (csrc/conv_tc.cu)."""
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
        with torch.cuda.device(cw.w.device):
            buf = torch.empty(nb, device=cw.w.device, dtype=torch.uint8)
            L.check(lib.tdvc_conv2d_pack_f16(p, buf.data_ptr(), torch.cuda.current_stream(cw.w.device).cuda_stream),
                    "conv2d_pack_f16")
        cw.w_f16 = buf
    return packed
