"""Operator-level surface: drop-ins for the reference's only native module, `_ext` (reference
main/utils/dcnv2/src/vision.cpp:4-9).  `dcn_v2_forward` / `dcn_v2_backward` take the reference's argument lists
(src/dcn_v2.h:9-46, :48-92) and tensor conventions (contiguous NCHW fp32 CUDA; offset (N, 2*dg*kh*kw, H, W) ordered
[g][tap][dy,dx]; mask (N, dg*kh*kw, H, W)), return freshly allocated tensors, raise RuntimeError on invalid input (the
reference raises through AT_ASSERTM, src/cuda/dcn_v2_cuda.cu:38-62) and launch on the current stream.  `dcn_v2_conv` is the
autograd function built on them (reference dcn_v2_amp.py:24-122, `_DCNv2.apply`).  No CPU path."""
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


def dcn_v2_backward(input, weight, bias, offset, mask, grad_output, kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w,
                    dilation_h, dilation_w, deformable_group):
    """-> [grad_input, grad_offset, grad_mask, grad_weight, grad_bias] (reference dcn_v2.h:48-92).  Deterministic: two calls on
    the same inputs return identical bits (the reference's col2im scatters with float atomicAdd)."""
    for name, t in (("input", input), ("weight", weight), ("bias", bias), ("offset", offset), ("mask", mask),
                    ("grad_output", grad_output)):
        if not t.is_cuda:
            raise RuntimeError(f"dcn_v2_backward: {name} must be a CUDA tensor (tdvc_b200 has no CPU path)")
        if t.dtype != torch.float32:
            raise RuntimeError(f"dcn_v2_backward: {name} must be float32")
    if input.dim() != 4 or weight.dim() != 4 or grad_output.dim() != 4:
        raise RuntimeError("dcn_v2_backward: input, weight and grad_output must be 4-D")
    N, C, H, W = input.shape
    O = weight.shape[0]
    if weight.shape[1] != C or weight.shape[2] != kernel_h or weight.shape[3] != kernel_w:
        raise RuntimeError(f"dcn_v2_backward: weight shape {tuple(weight.shape)} does not match input channels {C} "
                           f"and kernel ({kernel_h},{kernel_w})")
    Ho = (H + 2 * pad_h - (dilation_h * (kernel_h - 1) + 1)) // stride_h + 1
    Wo = (W + 2 * pad_w - (dilation_w * (kernel_w - 1) + 1)) // stride_w + 1
    K = kernel_h * kernel_w
    if tuple(offset.shape) != (N, 2 * deformable_group * K, Ho, Wo) or tuple(mask.shape) != (N, deformable_group * K, Ho, Wo):
        raise RuntimeError("dcn_v2_backward: offset / mask shape mismatch")
    if tuple(grad_output.shape) != (N, O, Ho, Wo) or bias.numel() != O:
        raise RuntimeError("dcn_v2_backward: grad_output / bias shape mismatch")
    lib = L.load()
    input, weight, offset, mask, grad_output = (t.contiguous() for t in (input, weight, offset, mask, grad_output))
    g_in, g_off, g_msk = torch.empty_like(input), torch.empty_like(offset), torch.empty_like(mask)
    g_w, g_b = torch.empty_like(weight), torch.empty(O, device=input.device, dtype=torch.float32)
    with torch.cuda.device(input.device):
        nb = lib.tdvc_dcn_v2_backward_workspace_bytes(N, C, H, W)
        ws = torch.empty(nb // 8 + 1, device=input.device, dtype=torch.int64)
        st = torch.cuda.current_stream(input.device)
        rc = lib.tdvc_dcn_v2_backward(input.data_ptr(), weight.data_ptr(), offset.data_ptr(), mask.data_ptr(),
                                      grad_output.data_ptr(), g_in.data_ptr(), g_off.data_ptr(), g_msk.data_ptr(),
                                      g_w.data_ptr(), g_b.data_ptr(), N, C, O, H, W, kernel_h, kernel_w, stride_h, stride_w,
                                      pad_h, pad_w, dilation_h, dilation_w, deformable_group, ws.data_ptr(), ws.numel() * 8,
                                      st.cuda_stream)
        L.check(rc, "dcn_v2_backward")
        ws.record_stream(st)
    return [g_in, g_off, g_msk, g_w, g_b]


class _DCNv2(torch.autograd.Function):
    """reference main/utils/dcnv2/dcn_v2_amp.py:24-122 (`_DCNv2`): forward / backward through the two native ops above.
    (The reference's forward returns `.half()` when its module-level `use_amp` is set; that rounding is a property of the
    caller's configuration and stays outside this function: VideoCompressor applies it in its fused kernel.)"""

    @staticmethod
    def forward(ctx, input, offset, mask, weight, bias, stride, padding, dilation, deformable_groups):
        pair = lambda v: (v, v) if isinstance(v, int) else tuple(v)
        ctx.stride, ctx.padding, ctx.dilation = pair(stride), pair(padding), pair(dilation)
        ctx.kernel_size = tuple(weight.shape[2:4])
        ctx.deformable_groups = deformable_groups
        out = dcn_v2_forward(input, weight, bias, offset, mask, ctx.kernel_size[0], ctx.kernel_size[1], ctx.stride[0],
                             ctx.stride[1], ctx.padding[0], ctx.padding[1], ctx.dilation[0], ctx.dilation[1], deformable_groups)
        ctx.save_for_backward(input, offset, mask, weight, bias)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_output):
        input, offset, mask, weight, bias = ctx.saved_tensors
        g = dcn_v2_backward(input, weight, bias, offset, mask, grad_output.float(), ctx.kernel_size[0], ctx.kernel_size[1],
                            ctx.stride[0], ctx.stride[1], ctx.padding[0], ctx.padding[1], ctx.dilation[0], ctx.dilation[1],
                            ctx.deformable_groups)
        return g[0], g[1], g[2], g[3], g[4], None, None, None, None


dcn_v2_conv = _DCNv2.apply
