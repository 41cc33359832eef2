"""Operator-level surface: drop-ins for the reference's only native module, `_ext` (reference
main/utils/dcnv2/src/vision.cpp:4-9).  `dcn_v2_forward` / `dcn_v2_backward` take the reference's argument lists
(src/dcn_v2.h:9-46, :48-92) and tensor conventions (contiguous NCHW fp32 CUDA; offset (N, 2*dg*kh*kw, H, W) ordered
[g][tap][dy,dx]; mask (N, dg*kh*kw, H, W)), return freshly allocated tensors, raise RuntimeError on invalid input (the
reference raises through AT_ASSERTM, src/cuda/dcn_v2_cuda.cu:38-62) and launch on the current stream.  `dcn_v2_conv` is the
autograd function built on them (reference dcn_v2_amp.py:24-122, `_DCNv2.apply`).  `conv2d` is `F.conv2d` (+ an optional
fused activation) with autograd through our own kernels: the convolution forward / dgrad on the tcgen05 kernels, wgrad and the
bias gradient on `tdvc_conv2d_wgrad` - the first convolution slice of the training step.  No CPU path."""
import weakref

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
        nb = lib.tdvc_dcn_v2_backward_workspace_bytes(N, C, O, H, W, Ho, Wo, K)
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


# ----------------------------------------------------------------------------------------------- conv2d with autograd
# fp16 / TF32 products of the weight-gradient MMAs: 3 = fp32-class (default), 1 = one TF32 product.  `VideoCompressor` sets it
# from the reference's own switch for the duration of a training forward (`enabled_amp=True`, the shipped cfg/train.yaml: the
# reference then trains with fp16 autocast products); every function captures the value at forward time.
WGRAD_PRODUCTS = 3

_ACTS = {None: L.ACT_NONE, "none": L.ACT_NONE, "relu": L.ACT_RELU, "leaky_relu": L.ACT_LRELU, "clamp01": L.ACT_CLAMP01}
_PACKS = {}


_WMAX = {}


def _weight_max(weight, key):
    """max |w| for the power-of-two weight scaling of the fp16 schemes (tc._pack), without stalling the host every step: the
    value comes from the copy issued at the PREVIOUS pack of the same parameter (`key`: a stable object, the parameter itself
    unless the caller names one) - one optimiser step old, and the scaling leaves a factor 4 of headroom; the first sight of
    a parameter pays one synchronous read."""
    ent = _WMAX.get(id(key))
    if ent is None or ent[0]() is not key:
        if len(_WMAX) >= 4096:
            _WMAX.clear()
        m = float(weight.detach().abs().max())
        pinned = torch.empty(1, dtype=torch.float32, pin_memory=True)
        pinned[0] = m
        ent = (weakref.ref(key), pinned, torch.cuda.Event(), [weight._version])
        _WMAX[id(key)] = ent
        return m
    if torch.cuda.is_current_stream_capturing():
        return float(ent[1][0])
    ent[2].synchronize()
    m = float(ent[1][0])
    if ent[3][0] != weight._version:      # one refresh per optimiser step (the forward and the dgrad pack both come here)
        ent[3][0] = weight._version
        ent[1].copy_(weight.detach().abs().max().reshape(1), non_blocking=True)
        ent[2].record(torch.cuda.current_stream(weight.device))
    return m


def forget_weight_maxima():
    """Drop the one-step-old maxima of `_weight_max`: call it after replacing weights wholesale (VideoCompressor.load_state_dict
    does) - a maximum remembered from other weights would scale the fp16 blocks of the first step after the swap wrongly."""
    _WMAX.clear()


def _packed(weight, bias, stride, pad, transposed, max_key=None):
    """ConvW (fp32 implicit-GEMM layout + tcgen05 fp16 blocks) of `weight`, or of its transposed, spatially flipped form
    (the dgrad operator), cached on the parameter's identity and version.  Packed by one kernel (tdvc_conv2d_pack_weight):
    the training step re-packs every weight after every optimiser step."""
    from tdvc_b200 import model as M, tc
    key = (weight.data_ptr(), weight._version, None if bias is None else (bias.data_ptr(), bias._version), stride, pad,
           transposed)
    ent = _PACKS.get(key)
    # an address can be handed to another tensor once its owner is freed: the entry also remembers WHICH tensors it packed
    cw = ent[0] if ent is not None and ent[1]() is weight and (bias is None or ent[2]() is bias) else None
    if cw is None:
        if len(_PACKS) >= 256:
            _PACKS.clear()
        w = weight.detach().float().contiguous()
        O, I, k, _ = w.shape
        if transposed:
            O, I = I, O
        dev = w.device
        cw = M.ConvW()
        cw.cin = M._r(I, 4)
        cw.cin_pad, cw.cout_pad = M._r(cw.cin, 8), M._r(O, 16)
        cw.cout, cw.k, cw.cin_real = O, k, I
        cw.w = torch.empty(k * k, cw.cin_pad, cw.cout_pad, device=dev, dtype=torch.float32)
        has_b = bias is not None and not transposed
        cw.b = torch.empty(cw.cout_pad, device=dev, dtype=torch.float32) if has_b else None
        b = bias.detach().float().contiguous() if has_b else None
        with torch.cuda.device(dev):
            L.check(L.load().tdvc_conv2d_pack_weight(w.data_ptr(), w.shape[0], w.shape[1], k, 1 if transposed else 0, cw.w.data_ptr(),
                                                     cw.cin_pad, cw.cout_pad, b.data_ptr() if has_b else None,
                                                     cw.b.data_ptr() if has_b else None,
                                                     torch.cuda.current_stream(dev).cuda_stream), "conv2d_pack_weight")
        cw.pad, cw.shuffle, cw.stride = pad, 0, stride
        cw.w_f16, cw.w_shift, cw.w_f16_p1, cw.w_shift_p1, cw.p1_ok = None, 0, None, 0, False
        if max_key is None:   # a view of a parameter (a squeezed Conv3d weight) is a new object every step: key on its base
            max_key = weight._base if weight._base is not None else weight
        cw.wmax = lambda: _weight_max(weight, max_key)    # asked for by the split / one-product schemes only (tc._pack)
        tc.attach_f16({"w": cw})
        cw.wmax = None     # (the closure would keep `weight` alive inside the cache)
        _PACKS[key] = (cw, weakref.ref(weight), weakref.ref(bias) if bias is not None else None)
    return cw


def _launch_conv(x, cw, out, stride, act, slope, impl):
    lib = L.load()
    p = L.ConvParams()
    p.src[0], p.src_c[0], p.src_ld[0], p.n_src = x.ptr, x.ld, x.ld, 1
    p.N, p.H, p.W = x.N, x.H, x.W
    p.Ho, p.Wo = out.H, out.W
    p.weight = cw.w.data_ptr()
    p.bias = cw.b.data_ptr() if cw.b is not None else None
    p.cin, p.cin_pad, p.cout, p.cout_pad = cw.cin, cw.cin_pad, cw.cout, cw.cout_pad
    p.kh = p.kw = cw.k
    p.stride, p.pad = stride, cw.pad
    p.act, p.slope = act, slope
    p.out, p.out_ld = out.ptr, out.ld
    p.impl = impl
    p.weight_f16 = cw.w_f16.data_ptr() if cw.w_f16 is not None else None
    p.w_shift = cw.w_shift
    L.check(lib.tdvc_conv2d(p, torch.cuda.current_stream(out.t.device).cuda_stream), "conv2d")


def _nhwc(t):
    """NCHW-shaped tensor -> channels-last Act.  A tensor that already lives in torch's channels_last memory format (what our
    own functions return) with a channel count that is a multiple of 4 is wrapped WITHOUT a copy - the Act then aliases it and
    must be treated as read-only; anything else goes through the layout kernel (channels padded to a multiple of 4 with zeros)."""
    from tdvc_b200.model import Act
    N, C, H, W = t.shape
    if C % 4 == 0 and t.dtype == torch.float32 and t.is_contiguous(memory_format=torch.channels_last) and t.data_ptr() % 16 == 0:
        v = t.permute(0, 2, 3, 1)
        return Act(v, v.data_ptr(), N, H, W, C, C)
    t = t.float().contiguous()
    a = Act.alloc(N, H, W, C, t.device, ld=(C + 3) // 4 * 4, zero=(C % 4 != 0))
    L.check(L.load().tdvc_nchw_to_nhwc(t.data_ptr(), a.ptr, N, C, H, W, a.ld, torch.cuda.current_stream(t.device).cuda_stream),
            "nchw_to_nhwc")
    return a


def _nchw(a):
    """Act -> NCHW-shaped tensor: a channels_last VIEW of the Act's buffer when it is dense (ld == C), a copy otherwise."""
    if a.ld == a.C and a.t.dim() == 4 and a.t.data_ptr() == a.ptr and tuple(a.t.shape) == (a.N, a.H, a.W, a.C):
        return a.t.permute(0, 3, 1, 2)
    out = torch.empty((a.N, a.C, a.H, a.W), device=a.t.device, dtype=torch.float32)
    L.check(L.load().tdvc_nhwc_to_nchw(a.ptr, a.ld, out.data_ptr(), a.N, a.C, a.H, a.W,
                                      torch.cuda.current_stream(a.t.device).cuda_stream), "nhwc_to_nchw")
    return out


class _Conv2d(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, stride, padding, act, slope, impl, max_key=None):
        from tdvc_b200.model import Act
        for name, t in (("input", x), ("weight", weight)) + ((("bias", bias),) if bias is not None else ()):
            if not t.is_cuda or t.dtype != torch.float32:
                raise RuntimeError(f"conv2d: {name} must be a float32 CUDA tensor (tdvc_b200 has no CPU path)")
        if x.dim() != 4 or weight.dim() != 4 or weight.shape[1] != x.shape[1] or weight.shape[2] != weight.shape[3]:
            raise RuntimeError(f"conv2d: input {tuple(x.shape)} / weight {tuple(weight.shape)} mismatch (square kernels, groups = 1)")
        if act not in _ACTS:
            raise RuntimeError(f"conv2d: unknown activation {act!r}")
        N, C, H, W = x.shape
        O, _, k, _ = weight.shape
        if 2 * padding != k - 1 and not (k == 1 and padding == 0):
            raise RuntimeError("conv2d: padding must be (k - 1) / 2 (every convolution of the reference is)")
        Ho, Wo = (H + 2 * padding - k) // stride + 1, (W + 2 * padding - k) // stride + 1
        with torch.cuda.device(x.device):
            xa = _nhwc(x.detach())
            cw = _packed(weight, bias, stride, padding, False, max_key)
            ya = Act.alloc(N, Ho, Wo, O, x.device, ld=(O + 3) // 4 * 4, zero=(O % 4 != 0))
            _launch_conv(xa, cw, ya, stride, _ACTS[act], slope, impl)
            y = _nchw(ya)
        ctx.geom = (stride, padding, _ACTS[act], slope, impl, bias is not None)
        ctx.max_key = max_key
        ctx.wg_products = WGRAD_PRODUCTS
        ctx.xa, ctx.ya = xa, (ya if _ACTS[act] != L.ACT_NONE else None)
        ctx.save_for_backward(weight)
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gy):
        from tdvc_b200.model import Act
        (weight,) = ctx.saved_tensors
        stride, padding, act, slope, impl, has_bias = ctx.geom
        xa = ctx.xa
        lib = L.load()
        O, C, k, _ = weight.shape
        dev = gy.device
        with torch.cuda.device(dev):
            st = torch.cuda.current_stream(dev).cuda_stream
            ga = _nhwc(gy)
            if act != L.ACT_NONE:   # g * f'(y), from the stored output (into a buffer of our own: `ga` may alias autograd's tensor)
                gp = Act.alloc(ga.N, ga.H, ga.W, ga.C, dev, ld=ga.ld)
                L.check(lib.tdvc_act_backward(ctx.ya.ptr, ga.ptr, gp.ptr, ga.N * ga.H * ga.W * ga.ld, act, slope, st), "act_backward")
                ga = gp
            gx = gw = gb = None
            if ctx.needs_input_grad[0]:
                cwt = _packed(weight, None, 1, k - 1 - padding, True, ctx.max_key)
                if stride > 1:
                    up = Act.alloc(xa.N, xa.H, xa.W, O, dev, ld=ga.ld)
                    L.check(lib.tdvc_zero_insert(ga.ptr, ga.ld, up.ptr, up.ld, xa.N, xa.H, xa.W, ga.H, ga.W, ga.ld, stride, st),
                            "zero_insert")
                else:
                    up = ga
                gxa = Act.alloc(xa.N, xa.H, xa.W, C, dev, ld=(C + 3) // 4 * 4, zero=(C % 4 != 0))
                _launch_conv(up, cwt, gxa, 1, L.ACT_NONE, 0.0, impl)
                gx = _nchw(gxa)
            if ctx.needs_input_grad[1] or (has_bias and ctx.needs_input_grad[2]):
                gw = torch.empty_like(weight)
                gb = torch.empty(O, device=dev, dtype=torch.float32) if has_bias else None
                nb = lib.tdvc_conv2d_wgrad_workspace_bytes(xa.N, ga.H, ga.W, C, O, k)
                ws = torch.empty((nb + 3) // 4, device=dev, dtype=torch.float32)
                L.check(lib.tdvc_conv2d_wgrad(xa.ptr, xa.ld, ga.ptr, ga.ld, xa.N, xa.H, xa.W, C, O, k, stride, padding, 0,
                                              ctx.wg_products, gw.data_ptr(), gb.data_ptr() if gb is not None else None, ws.data_ptr(), nb, st),
                        "conv2d_wgrad")
        return gx, gw, gb, None, None, None, None, None, None


def conv2d(input, weight, bias=None, stride=1, padding=0, act=None, slope=0.01, impl=L.IMPL_AUTO, max_key=None):
    """F.conv2d(input, weight, bias, stride, padding) followed by `act` (None | "relu" | "leaky_relu" | "clamp01"), NCHW float32
    CUDA tensors, with autograd: grad_input by the forward kernels on the transposed, flipped weight (fp32-class tcgen05 path
    included), grad_weight / grad_bias by the deterministic fp32 wgrad kernel.  impl: lib.IMPL_* (1 = exact fp32 SIMT forward
    and dgrad)."""
    return _Conv2d.apply(input, weight, bias, int(stride), int(padding), act, float(slope), int(impl), max_key)


# ----------------------------------------------------------------------------------------------- GDN / IGDN with autograd
def _launch_gdn(x, cw, out, inverse, norm_only, impl, absmax):
    """out = x * (beta + gamma . x^2)^(-+1/2) on the fused tensor-core kernel (1x1 convolution of x^2 with the GDN epilogue),
    or just the norm (norm_only: the backward pass recomputes it instead of keeping a second activation)."""
    lib = L.load()
    p = L.ConvParams()
    p.src[0], p.src_c[0], p.src_ld[0], p.n_src = x.ptr, x.ld, x.ld, 1
    p.N, p.H, p.W, p.Ho, p.Wo = x.N, x.H, x.W, x.H, x.W
    p.weight, p.bias = cw.w.data_ptr(), cw.b.data_ptr()
    p.cin, p.cin_pad, p.cout, p.cout_pad = cw.cin, cw.cin_pad, cw.cout, cw.cout_pad
    p.kh = p.kw = 1
    p.stride, p.pad = 1, 0
    p.in_square = 1
    if not norm_only:
        p.post = L.POST_IGDN if inverse else L.POST_GDN
        p.mul, p.mul_ld = x.ptr, x.ld
    p.out, p.out_ld = out.ptr, out.ld
    p.impl = impl
    p.weight_f16 = cw.w_f16.data_ptr() if cw.w_f16 is not None else None
    p.w_shift = cw.w_shift
    p.in_absmax = absmax.data_ptr()   # max |x|: x * x is pre-scaled by a power of two so that the fp16 operands cannot saturate
    L.check(lib.tdvc_conv2d(p, torch.cuda.current_stream(out.t.device).cuda_stream), "conv2d (gdn)")


class _GDN(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, beta, gamma, inverse, impl, max_key):
        from tdvc_b200.model import Act
        for name, t in (("input", x), ("beta", beta), ("gamma", gamma)):
            if not t.is_cuda or t.dtype != torch.float32:
                raise RuntimeError(f"gdn: {name} must be a float32 CUDA tensor (tdvc_b200 has no CPU path)")
        N, C, H, W = x.shape
        if tuple(gamma.shape) != (C, C) or tuple(beta.shape) != (C,) or C % 4:
            raise RuntimeError(f"gdn: expected beta ({C},) and gamma ({C},{C}), channels a multiple of 4")
        with torch.cuda.device(x.device):
            xa = _nhwc(x.detach())
            w4 = gamma.detach().reshape(C, C, 1, 1)
            cw = _packed(w4, beta.detach(), 1, 0, False, max_key)
            ya = Act.alloc(N, H, W, C, x.device)
            am = x.detach().abs().amax().reshape(1).float()   # (the inference path gets it from the producing layer's epilogue)
            _launch_gdn(xa, cw, ya, inverse, False, impl, am)
            y = _nchw(ya)
        ctx.xa, ctx.cw, ctx.w4, ctx.inverse, ctx.impl, ctx.am, ctx.max_key = xa, cw, w4, inverse, impl, am, max_key
        ctx.wg_products = WGRAD_PRODUCTS
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gy):
        from tdvc_b200.model import Act
        xa, cw, inverse, impl = ctx.xa, ctx.cw, ctx.inverse, ctx.impl
        lib = L.load()
        N, H, W, C = xa.N, xa.H, xa.W, xa.C
        dev = gy.device
        n = N * H * W * C
        with torch.cuda.device(dev):
            st = torch.cuda.current_stream(dev).cuda_stream
            ga = _nhwc(gy)
            norm = Act.alloc(N, H, W, C, dev)
            _launch_gdn(xa, cw, norm, inverse, True, impl, ctx.am)
            dxd, dn = Act.alloc(N, H, W, C, dev), Act.alloc(N, H, W, C, dev)
            L.check(lib.tdvc_gdn_backward_pre(xa.ptr, norm.ptr, ga.ptr, dxd.ptr, dn.ptr, n, 1 if inverse else 0, st), "gdn_backward_pre")
            cwt = _packed(ctx.w4, None, 1, 0, True, ctx.max_key)          # gamma^T: d(x^2) = dgrad of the 1x1 convolution
            dxsq = Act.alloc(N, H, W, C, dev)
            _launch_conv(dn, cwt, dxsq, 1, L.ACT_NONE, 0.0, impl)
            L.check(lib.tdvc_gdn_backward_post(dxd.ptr, xa.ptr, dxsq.ptr, dxd.ptr, n, st), "gdn_backward_post")
            gx = _nchw(dxd)
            gw = torch.empty((C, C, 1, 1), device=dev, dtype=torch.float32)
            gb = torch.empty(C, device=dev, dtype=torch.float32)
            nb = lib.tdvc_conv2d_wgrad_workspace_bytes(N, H, W, C, C, 1)
            ws = torch.empty((nb + 3) // 4, device=dev, dtype=torch.float32)
            L.check(lib.tdvc_conv2d_wgrad(xa.ptr, xa.ld, dn.ptr, dn.ld, N, H, W, C, C, 1, 1, 0, 1, ctx.wg_products, gw.data_ptr(), gb.data_ptr(),
                                          ws.data_ptr(), nb, st), "conv2d_wgrad (gdn)")
        return gx, gb, gw.view(C, C), None, None, None


def gdn(input, beta, gamma, inverse=False, impl=L.IMPL_AUTO, max_key=None):
    """compressai GDN / IGDN on NCHW float32 CUDA tensors: input * (beta + gamma . input^2)^(-1/2) (inverse: ^(+1/2)), with
    autograd.  beta (C,) and gamma (C, C) are the EFFECTIVE (reparametrised, non-negative) values - compressai's
    `beta_reparam(self.beta)`, `gamma_reparam(self.gamma)`, which stay ordinary torch parameter arithmetic on the caller's side.
    Forward: the fused tcgen05 kernel of the inference path; backward: two element-wise kernels around the 1x1 convolution's
    dgrad / wgrad."""
    return _GDN.apply(input, beta, gamma, bool(inverse), int(impl), max_key)


# ----------------------------------------------------------------------------------------------- likelihood -> bits with autograd
def _scalar_from_acc(acc):
    return acc.float().reshape(())


class _GcBits(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y, scales, means, noise):
        from tdvc_b200.model import Act
        for name, t in (("y", y), ("scales", scales), ("means", means), ("noise", noise)):
            if not t.is_cuda or t.dtype != torch.float32 or t.shape != y.shape:
                raise RuntimeError(f"gaussian_conditional_bits: {name} must be a float32 CUDA tensor of y's shape")
        N, C, H, W = y.shape
        lib = L.load()
        with torch.cuda.device(y.device):
            st = torch.cuda.current_stream(y.device).cuda_stream
            ya, na = _nhwc(y.detach()), _nhwc(noise.detach())
            pa = _nhwc(torch.cat((scales.detach(), means.detach()), 1))   # (scales | means), the layout of the entropy-parameter output
            acc = torch.zeros(1, device=y.device, dtype=torch.float64)
            L.check(lib.tdvc_gc_bits_noise(ya.ptr, na.ptr, pa.ptr, pa.ld, N * H * W, C, acc.data_ptr(), st), "gc_bits_noise")
        ctx.t = (ya, na, pa)
        return _scalar_from_acc(acc)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gS):
        from tdvc_b200.model import Act
        ya, na, pa = ctx.t
        N, H, W, C = ya.N, ya.H, ya.W, ya.C
        lib = L.load()
        dev = gS.device
        with torch.cuda.device(dev):
            st = torch.cuda.current_stream(dev).cuda_stream
            g = gS.detach().float().reshape(1).contiguous()
            gy, gp = Act.alloc(N, H, W, C, dev), Act.alloc(N, H, W, 2 * C, dev)
            L.check(lib.tdvc_gc_bits_backward(ya.ptr, na.ptr, pa.ptr, pa.ld, g.data_ptr(), gy.ptr, gp.ptr, N * H * W, C, st),
                    "gc_bits_backward")
            return _nchw(gy), _nchw(gp.chan(0, C)), _nchw(gp.chan(C, C)), None


def gaussian_conditional_bits(y, scales, means, noise):
    """sum ln max(p, 1e-9) of compressai's GaussianConditional in training mode - the likelihood of y + noise under N(means,
    max(scales, 0.11)) integrated over the quantisation bin - as a 0-d tensor with autograd to y, scales and means (NCHW float32
    CUDA tensors of one shape; reference pnet.py:38-43 divides the sum by -ln 2 * num_pixels)."""
    return _GcBits.apply(y, scales, means, noise)


class _EbBits(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, noise, mats, biases, factors):
        from tdvc_b200.model import Act
        N, C, H, W = z.shape
        if tuple(mats.shape) != (C, 33) or tuple(biases.shape) != (C, 13) or tuple(factors.shape) != (C, 12):
            raise RuntimeError("entropy_bottleneck_bits: expected filters (3, 3, 3, 3)")
        lib = L.load()
        with torch.cuda.device(z.device):
            st = torch.cuda.current_stream(z.device).cuda_stream
            za, na = _nhwc(z.detach()), _nhwc(noise.detach())
            zt = Act.alloc(N, H, W, C, z.device)
            m, b, f = (t.detach().float().contiguous() for t in (mats, biases, factors))
            acc = torch.zeros(1, device=z.device, dtype=torch.float64)
            L.check(lib.tdvc_eb_bits_noise(za.ptr, na.ptr, zt.ptr, m.data_ptr(), b.data_ptr(), f.data_ptr(), N * H * W, C,
                                           acc.data_ptr(), st), "eb_bits_noise")
            z_tilde = _nchw(zt)
        ctx.t = (zt, m, b, f)
        return z_tilde, _scalar_from_acc(acc)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_zt, gS):
        from tdvc_b200.model import Act
        zt, m, b, f = ctx.t
        N, H, W, C = zt.N, zt.H, zt.W, zt.C
        lib = L.load()
        dev = m.device
        with torch.cuda.device(dev):
            st = torch.cuda.current_stream(dev).cuda_stream
            g = gS.detach().float().reshape(1).contiguous()
            gz = Act.alloc(N, H, W, C, dev)
            gm, gb, gf = torch.empty_like(m), torch.empty_like(b), torch.empty_like(f)
            L.check(lib.tdvc_eb_bits_backward(zt.ptr, m.data_ptr(), b.data_ptr(), f.data_ptr(), g.data_ptr(), gz.ptr,
                                              gm.data_ptr(), gb.data_ptr(), gf.data_ptr(), N * H * W, C, st), "eb_bits_backward")
            if g_zt is not None:   # z~ = z + noise: what flows into z~ from its consumers flows into z unchanged
                ga = _nhwc(g_zt)
                L.check(lib.tdvc_axpby(gz.ptr, ga.ptr, gz.ptr, N * H * W * C, 1.0, 1.0, st), "axpby")
            return _nchw(gz), None, gm, gb, gf


def entropy_bottleneck_bits(z, noise, matrices, biases, factors):
    """compressai EntropyBottleneck in training mode: z~ = z + noise and sum ln max(p(z~), 1e-9) of the factorised prior
    (per-channel cumulative MLP, widths 1,3,3,3,3,1) -> (z~, 0-d sum), with autograd to z and to the RAW parameters
    (`_matrix0..4`, `_bias0..4`, `_factor0..3` as compressai stores them: the softplus / tanh reparametrisation is ordinary
    torch arithmetic on these small tensors, the element-wise work and its backward are kernels)."""
    C = z.shape[1]
    m = torch.cat([torch.nn.functional.softplus(t).reshape(C, -1) for t in matrices], 1)
    b = torch.cat([t.reshape(C, -1) for t in biases], 1)
    f = torch.cat([torch.tanh(t).reshape(C, -1) for t in factors], 1)
    return _EbBits.apply(z, noise, m, b, f)


# ----------------------------------------------------------------------------------------------- squeeze-excitation pieces
def _chan_dot(a, b, scale):
    """(N, C) = scale * sum over pixels of a * b (b None: of a); Acts with dense rows."""
    lib = L.load()
    dev = a.t.device
    out = torch.empty((a.N, a.C), device=dev, dtype=torch.float32)
    nb = lib.tdvc_chan_dot_workspace_bytes(a.N, a.H * a.W, a.C)
    ws = torch.empty((nb + 3) // 4, device=dev, dtype=torch.float32)
    L.check(lib.tdvc_chan_dot(a.ptr, b.ptr if b is not None else None, out.data_ptr(), a.N, a.H * a.W, a.C, scale, ws.data_ptr(), nb,
                              torch.cuda.current_stream(dev).cuda_stream), "chan_dot")
    return out


def _chan_affine(a, s, t, like):
    """Act out[n][p][c] = a * s[n][c] + t[n][c] (a / t may be None); `like` gives the shape."""
    from tdvc_b200.model import Act
    lib = L.load()
    dev = like.t.device
    out = Act.alloc(like.N, like.H, like.W, like.C, dev)
    L.check(lib.tdvc_chan_affine(a.ptr if a is not None else None, s.data_ptr(), t.data_ptr() if t is not None else None, out.ptr,
                                 like.N, like.H * like.W, like.C, torch.cuda.current_stream(dev).cuda_stream), "chan_affine")
    return out


def _dense(t):
    a = _nhwc(t)
    if a.ld != a.C:
        raise RuntimeError("channel count must be a multiple of 4")
    return a


class _ChannelMean(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        with torch.cuda.device(x.device):
            xa = _dense(x.detach())
            ctx.like = xa
            return _chan_dot(xa, None, 1.0 / (xa.H * xa.W))

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        like = ctx.like
        with torch.cuda.device(g.device):
            s = (g.float() * (1.0 / (like.H * like.W))).contiguous()
            return _nchw(_chan_affine(None, s, None, like))


class _ChannelScale(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, s):
        with torch.cuda.device(x.device):
            xa = _dense(x.detach())
            sc = s.detach().float().contiguous()
            ctx.t = (xa, sc)
            return _nchw(_chan_affine(xa, sc, None, xa))

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        xa, sc = ctx.t
        with torch.cuda.device(g.device):
            ga = _dense(g)
            return _nchw(_chan_affine(ga, sc, None, xa)), _chan_dot(ga, xa, 1.0)


def channel_mean(x):
    """(N, C, H, W) -> (N, C): the spatial mean (adaptive_avg_pool2d(x, 1)), with autograd; deterministic two-stage sum."""
    return _ChannelMean.apply(x)


def channel_scale(x, s):
    """x * s[:, :, None, None] for s of shape (N, C), with autograd to both (the gated product of a squeeze-excitation layer)."""
    return _ChannelScale.apply(x, s)


# ----------------------------------------------------------------------------------------------- SPyNet level input with autograd
def _img4(t):
    """(N, 3, h, w) image -> channels-last Act with 4 stored channels (the layout the SPyNet kernels read)."""
    from tdvc_b200.model import Act
    N, C, H, W = t.shape
    a = Act.alloc(N, H, W, C, t.device, ld=4, zero=True)
    tt = t.detach().float().contiguous()
    L.check(L.load().tdvc_nchw_to_nhwc(tt.data_ptr(), a.ptr, N, C, H, W, 4, torch.cuda.current_stream(t.device).cuda_stream), "nchw_to_nhwc")
    return a


class _SpynetLevelInput(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ref, supp, flow_prev):
        from tdvc_b200.model import Act
        N, _, h, w = ref.shape
        lib = L.load()
        with torch.cuda.device(ref.device):
            st = torch.cuda.current_stream(ref.device).cuda_stream
            ra, sa = _img4(ref), _img4(supp)
            fa = None
            if flow_prev is not None:
                if tuple(flow_prev.shape) != (N, 2, h // 2, w // 2):
                    raise RuntimeError("spynet_level_input: flow_prev must be (N, 2, h/2, w/2)")
                fa = Act.alloc(N, h // 2, w // 2, 2, ref.device, ld=2)
                fp = flow_prev.detach().float().contiguous()
                L.check(lib.tdvc_nchw_to_nhwc(fp.data_ptr(), fa.ptr, N, 2, h // 2, w // 2, 2, st), "nchw_to_nhwc")
            out = Act.alloc(N, h, w, 8, ref.device)
            L.check(lib.tdvc_spynet_prep(ra.ptr, sa.ptr, fa.ptr if fa is not None else None, out.ptr, N, h, w, st), "spynet_prep")
        ctx.t = (sa, fa, N, h, w)
        return _nchw(out)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        from tdvc_b200.model import Act
        sa, fa, N, h, w = ctx.t
        if fa is None:
            return None, None, None
        lib = L.load()
        with torch.cuda.device(g.device):
            st = torch.cuda.current_stream(g.device).cuda_stream
            ga = _nhwc(g)
            ws = torch.empty(N * h * w * 2, device=g.device, dtype=torch.float32)
            gfl = Act.alloc(N, h // 2, w // 2, 2, g.device, ld=2)
            L.check(lib.tdvc_spynet_prep_backward(sa.ptr, fa.ptr, ga.ptr, ws.data_ptr(), gfl.ptr, N, h, w, st), "spynet_prep_backward")
            out = torch.empty((N, 2, h // 2, w // 2), device=g.device, dtype=torch.float32)
            L.check(lib.tdvc_nhwc_to_nchw(gfl.ptr, 2, out.data_ptr(), N, 2, h // 2, w // 2, st), "nhwc_to_nchw")
        return None, None, out


def spynet_level_input(ref, supp, flow_prev=None):
    """One SPyNet level's input (reference flownet.py:116-138): cat([ref, flow_warp(supp, up), up], 1) with up = 2 x the x2
    bilinear upsample (align_corners=True) of the coarser level's flow (None at the coarsest level: zero flow) and flow_warp the
    border-clamped bilinear backward warp of reference flownet.py:8-48, in ONE kernel; autograd to flow_prev (the images are
    inputs of the network).  (N, 3, h, w) images, (N, 2, h/2, w/2) flow -> (N, 8, h, w)."""
    return _SpynetLevelInput.apply(ref, supp, flow_prev)
