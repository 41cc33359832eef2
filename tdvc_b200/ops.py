"""Operator-level surface.  `dcn_v2_forward` is the drop-in for the reference's only native op,
`_ext.dcn_v2_forward` (reference main/utils/dcnv2/src/vision.cpp:5, src/dcn_v2.h:9-46): same argument list,
same tensor conventions (contiguous NCHW fp32 CUDA; offset (N, 2*dg*kh*kw, H, W) ordered [g][tap][dy,dx];
mask (N, dg*kh*kw, H, W)), a freshly allocated output, RuntimeError on invalid input (the reference raises
through AT_ASSERTM, src/cuda/dcn_v2_cuda.cu:38-62), launched on the current stream.  No CPU path."""
import torch

from tdvc_b200 import lib as L


def dcn_v2_forward(input, weight, bias, offset, mask, kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w,
                   dilation_h, dilation_w, deformable_group):
    for name, t in (("input", input), ("weight", weight), ("bias", bias), ("offset", offset), ("mask", mask)):
        if not t.is_cuda:
            raise RuntimeError(f"dcn_v2_forward: {name} must be a CUDA tensor (tdvc_b200 has no CPU path)")
        if t.dtype != torch.float32:
            raise RuntimeError(f"dcn_v2_forward: {name} must be float32")
    if input.dim() != 4 or weight.dim() != 4:
        raise RuntimeError("dcn_v2_forward: input and weight must be 4-D")
    N, C, H, W = input.shape
    O = weight.shape[0]
    if weight.shape[1] != C or weight.shape[2] != kernel_h or weight.shape[3] != kernel_w:
        raise RuntimeError(f"dcn_v2_forward: weight shape {tuple(weight.shape)} does not match input channels {C} "
                           f"and kernel ({kernel_h},{kernel_w})")
    Ho = (H + 2 * pad_h - (dilation_h * (kernel_h - 1) + 1)) // stride_h + 1
    Wo = (W + 2 * pad_w - (dilation_w * (kernel_w - 1) + 1)) // stride_w + 1
    K = kernel_h * kernel_w
    if tuple(offset.shape) != (N, 2 * deformable_group * K, Ho, Wo) or tuple(mask.shape) != (N, deformable_group * K, Ho, Wo):
        raise RuntimeError("dcn_v2_forward: offset / mask shape mismatch")
    if bias.numel() != O:
        raise RuntimeError("dcn_v2_forward: bias shape mismatch")
    lib = L.load()
    input, weight, bias, offset, mask = (t.contiguous() for t in (input, weight, bias, offset, mask))
    out = torch.empty((N, O, Ho, Wo), device=input.device, dtype=torch.float32)
    with torch.cuda.device(input.device):
        nb = lib.tdvc_dcn_v2_workspace_bytes(N, C, O, H, W, deformable_group)
        ws = torch.empty(nb, device=input.device, dtype=torch.uint8)
        rc = lib.tdvc_dcn_v2_forward(input.data_ptr(), weight.data_ptr(), bias.data_ptr(), offset.data_ptr(),
                                     mask.data_ptr(), out.data_ptr(), N, C, O, H, W, kernel_h, kernel_w, stride_h,
                                     stride_w, pad_h, pad_w, dilation_h, dilation_w, deformable_group,
                                     ws.data_ptr(), nb, torch.cuda.current_stream(input.device).cuda_stream)
        L.check(rc, "dcn_v2_forward")
        ws.record_stream(torch.cuda.current_stream(input.device))
    return out
