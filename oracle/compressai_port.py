"""ORACLE (test infrastructure, not product code): CPU restatement of the CompressAI arithmetic
that TDVC's coders inherit.

The reference's `MVCoder`/`ResCoder` (reference main/model/encoder_v3.py:14-69) subclass
`compressai.models.waseda.Cheng2020Anchor`; `compressai` is a third-party dependency that is NOT in
/root/reference and is unpinned (reference requirement.txt:8, contemporaneous with torch 1.8 => 1.1.x).
This file restates the published CompressAI algorithm in plain PyTorch fp32.

PARITY UNPINNED: no reference test, golden vector or fixture pins anything at this boundary; this
restatement is the definition of truth for the build.  It is anchored only on the
reference's call sites: encoder_v3.py:3-11 (layer names), pnet.py:34-43,58-67 (forward / aux_loss /
likelihood use).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
anything under oracle/.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


# --------------------------------------------------------------------------- ops
class LowerBound(nn.Module):
    """max(x, bound) with CompressAI's pass-through gradient `(x >= bound) | (grad < 0)`."""

    class _Fn(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, bound):
            ctx.save_for_backward(x, bound)
            return torch.max(x, bound)

        @staticmethod
        def backward(ctx, g):
            x, bound = ctx.saved_tensors
            keep = (x >= bound) | (g < 0)
            return keep.type(g.dtype) * g, None

    def __init__(self, bound):
        super().__init__()
        self.register_buffer("bound", torch.Tensor([float(bound)]))

    def forward(self, x):
        return LowerBound._Fn.apply(x, self.bound)


class NonNegativeParametrizer(nn.Module):
    """reparam: forward(p) = max(p, sqrt(min + 2^-36))^2 - 2^-36 ; init(x) = sqrt(max(x + ped, ped))."""

    def __init__(self, minimum=0.0, reparam_offset=2 ** -18):
        super().__init__()
        self.minimum = float(minimum)
        self.reparam_offset = float(reparam_offset)
        pedestal = self.reparam_offset ** 2
        self.register_buffer("pedestal", torch.Tensor([pedestal]))
        self.lower_bound = LowerBound((self.minimum + self.reparam_offset ** 2) ** 0.5)

    def init(self, x):
        return torch.sqrt(torch.max(x + self.pedestal, self.pedestal))

    def forward(self, x):
        out = self.lower_bound(x)
        return out ** 2 - self.pedestal


class GDN(nn.Module):
    """y_i = x_i * (beta_i + sum_j gamma_ij x_j^2)^(-1/2)   (inverse: ^(+1/2))."""

    def __init__(self, in_channels, inverse=False, beta_min=1e-6, gamma_init=0.1):
        super().__init__()
        self.inverse = bool(inverse)
        self.beta_reparam = NonNegativeParametrizer(minimum=float(beta_min))
        self.beta = nn.Parameter(self.beta_reparam.init(torch.ones(in_channels)))
        self.gamma_reparam = NonNegativeParametrizer()
        self.gamma = nn.Parameter(self.gamma_reparam.init(float(gamma_init) * torch.eye(in_channels)))

    def forward(self, x):
        C = x.shape[1]
        beta = self.beta_reparam(self.beta)
        gamma = self.gamma_reparam(self.gamma).reshape(C, C, 1, 1)
        norm = F.conv2d(x ** 2, gamma, beta)
        norm = torch.sqrt(norm) if self.inverse else torch.rsqrt(norm)
        return x * norm


class MaskedConv2d(nn.Conv2d):
    """Mask-'A' autoregressive conv; the weight is masked IN PLACE on every forward."""

    def __init__(self, *args, mask_type="A", **kwargs):
        super().__init__(*args, **kwargs)
        self.register_buffer("mask", torch.ones_like(self.weight.data))
        _, _, h, w = self.mask.size()
        self.mask[:, :, h // 2, w // 2 + (mask_type == "B"):] = 0
        self.mask[:, :, h // 2 + 1:] = 0

    def forward(self, x):
        self.weight.data *= self.mask
        return super().forward(x)


def conv3x3(i, o, stride=1):
    return nn.Conv2d(i, o, kernel_size=3, stride=stride, padding=1)


def conv1x1(i, o, stride=1):
    return nn.Conv2d(i, o, kernel_size=1, stride=stride)


def subpel_conv3x3(i, o, r=1):
    return nn.Sequential(nn.Conv2d(i, o * r ** 2, kernel_size=3, padding=1), nn.PixelShuffle(r))


class ResidualBlockWithStride(nn.Module):
    def __init__(self, in_ch, out_ch, stride=2):
        super().__init__()
        self.conv1 = conv3x3(in_ch, out_ch, stride=stride)
        self.leaky_relu = nn.LeakyReLU(inplace=True)
        self.conv2 = conv3x3(out_ch, out_ch)
        self.gdn = GDN(out_ch)
        self.skip = conv1x1(in_ch, out_ch, stride=stride) if (stride != 1 or in_ch != out_ch) else None

    def forward(self, x):
        out = self.gdn(self.conv2(self.leaky_relu(self.conv1(x))))
        identity = x if self.skip is None else self.skip(x)
        return out + identity


class ResidualBlockUpsample(nn.Module):
    def __init__(self, in_ch, out_ch, upsample=2):
        super().__init__()
        self.subpel_conv = subpel_conv3x3(in_ch, out_ch, upsample)
        self.leaky_relu = nn.LeakyReLU(inplace=True)
        self.conv = conv3x3(out_ch, out_ch)
        self.igdn = GDN(out_ch, inverse=True)
        self.upsample = subpel_conv3x3(in_ch, out_ch, upsample)

    def forward(self, x):
        out = self.igdn(self.conv(self.leaky_relu(self.subpel_conv(x))))
        return out + self.upsample(x)


class ResidualBlock(nn.Module):
    def __init__(self, in_ch, out_ch):
        super().__init__()
        self.conv1 = conv3x3(in_ch, out_ch)
        self.leaky_relu = nn.LeakyReLU(inplace=True)
        self.conv2 = conv3x3(out_ch, out_ch)
        self.skip = conv1x1(in_ch, out_ch) if in_ch != out_ch else None

    def forward(self, x):
        out = self.leaky_relu(self.conv2(self.leaky_relu(self.conv1(x))))
        identity = x if self.skip is None else self.skip(x)
        return out + identity


# --------------------------------------------------------------------------- entropy models
class _EntropyModel(nn.Module):
    def __init__(self, likelihood_bound=1e-9):
        super().__init__()
        self.use_likelihood_bound = likelihood_bound > 0
        if self.use_likelihood_bound:
            self.likelihood_lower_bound = LowerBound(likelihood_bound)
        # CDF tables used only by compress()/decompress() (out of scope, kept for state_dict shape)
        self.register_buffer("_offset", torch.IntTensor())
        self.register_buffer("_quantized_cdf", torch.IntTensor())
        self.register_buffer("_cdf_length", torch.IntTensor())

    @staticmethod
    def quantize(inputs, mode, means=None):
        if mode == "noise":
            return inputs + torch.empty_like(inputs).uniform_(-0.5, 0.5)
        outputs = inputs.clone()
        if means is not None:
            outputs -= means
        outputs = torch.round(outputs)  # half-to-even
        if mode == "dequantize" and means is not None:
            outputs += means
        return outputs


class EntropyBottleneck(_EntropyModel):
    """Factorised prior: per-channel 5-layer MLP cumulative (widths 1,3,3,3,3,1)."""

    def __init__(self, channels, tail_mass=1e-9, init_scale=10, filters=(3, 3, 3, 3)):
        super().__init__()
        self.channels = int(channels)
        self.filters = tuple(int(f) for f in filters)
        self.init_scale = float(init_scale)
        self.tail_mass = float(tail_mass)
        filters = (1,) + self.filters + (1,)
        scale = self.init_scale ** (1 / (len(self.filters) + 1))
        for i in range(len(self.filters) + 1):
            init = math.log(math.expm1(1 / scale / filters[i + 1]))
            m = torch.Tensor(channels, filters[i + 1], filters[i])
            m.data.fill_(init)
            self.register_parameter(f"_matrix{i:d}", nn.Parameter(m))
            b = torch.Tensor(channels, filters[i + 1], 1)
            nn.init.uniform_(b, -0.5, 0.5)
            self.register_parameter(f"_bias{i:d}", nn.Parameter(b))
            if i < len(self.filters):
                f = torch.Tensor(channels, filters[i + 1], 1)
                nn.init.zeros_(f)
                self.register_parameter(f"_factor{i:d}", nn.Parameter(f))
        self.quantiles = nn.Parameter(torch.Tensor(channels, 1, 3))
        self.quantiles.data = torch.Tensor([-self.init_scale, 0, self.init_scale]).repeat(self.quantiles.size(0), 1, 1)
        target = math.log(2 / self.tail_mass - 1)
        self.register_buffer("target", torch.Tensor([-target, 0, target]))

    def _logits_cumulative(self, v, stop_gradient=False):
        for i in range(len(self.filters) + 1):
            m = getattr(self, f"_matrix{i:d}")
            b = getattr(self, f"_bias{i:d}")
            if stop_gradient:
                m, b = m.detach(), b.detach()
            v = torch.matmul(F.softplus(m), v) + b
            if i < len(self.filters):
                f = getattr(self, f"_factor{i:d}")
                if stop_gradient:
                    f = f.detach()
                v = v + torch.tanh(f) * torch.tanh(v)
        return v

    def loss(self):
        logits = self._logits_cumulative(self.quantiles, stop_gradient=True)
        return torch.abs(logits - self.target).sum()

    def _likelihood(self, x):
        lower = self._logits_cumulative(x - 0.5)
        upper = self._logits_cumulative(x + 0.5)
        sign = -torch.sign(lower + upper).detach()
        return torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower))

    def forward(self, x, training=None):
        training = self.training if training is None else training
        perm = (1, 0, 2, 3)  # channels first
        x = x.permute(*perm).contiguous()
        shape = x.size()
        v = x.reshape(x.size(0), 1, -1)
        med = self.quantiles[:, :, 1:2]
        out = self.quantize(v, "noise" if training else "dequantize", med)
        lik = self._likelihood(out)
        if self.use_likelihood_bound:
            lik = self.likelihood_lower_bound(lik)
        out = out.reshape(shape).permute(*perm).contiguous()
        lik = lik.reshape(shape).permute(*perm).contiguous()
        return out, lik


class GaussianConditional(_EntropyModel):
    def __init__(self, scale_table, scale_bound=0.11, tail_mass=1e-9):
        super().__init__()
        self.tail_mass = float(tail_mass)
        self.lower_bound_scale = LowerBound(scale_bound)
        self.register_buffer("scale_table", torch.Tensor(scale_table) if scale_table else torch.Tensor())
        self.register_buffer("scale_bound", torch.Tensor([float(scale_bound)]))

    @staticmethod
    def _standardized_cumulative(t):
        return 0.5 * torch.erfc(-(2 ** -0.5) * t)

    def _likelihood(self, x, scales, means=None):
        v = x - means if means is not None else x
        scales = self.lower_bound_scale(scales)
        v = torch.abs(v)
        upper = self._standardized_cumulative((0.5 - v) / scales)
        lower = self._standardized_cumulative((-0.5 - v) / scales)
        return upper - lower

    def forward(self, x, scales, means=None, training=None):
        training = self.training if training is None else training
        out = self.quantize(x, "noise" if training else "dequantize", means)
        lik = self._likelihood(out, scales, means)
        if self.use_likelihood_bound:
            lik = self.likelihood_lower_bound(lik)
        return out, lik


# --------------------------------------------------------------------------- model
class Cheng2020Anchor(nn.Module):
    """Joint autoregressive + hierarchical prior (Minnen 2018) with Cheng 2020 anchor transforms.

    g_a / g_s here are the stock ones; the reference overrides both (encoder_v3.py:17-40,46-69).
    """

    def __init__(self, N=192, **kwargs):
        super().__init__()
        M = N
        self.N, self.M = int(N), int(M)
        self.entropy_bottleneck = EntropyBottleneck(N)
        self.g_a = nn.Sequential(
            ResidualBlockWithStride(3, N, stride=2), ResidualBlock(N, N),
            ResidualBlockWithStride(N, N, stride=2), ResidualBlock(N, N),
            ResidualBlockWithStride(N, N, stride=2), ResidualBlock(N, N),
            conv3x3(N, N, stride=2))
        self.h_a = nn.Sequential(
            conv3x3(N, N), nn.LeakyReLU(inplace=True),
            conv3x3(N, N), nn.LeakyReLU(inplace=True),
            conv3x3(N, N, stride=2), nn.LeakyReLU(inplace=True),
            conv3x3(N, N), nn.LeakyReLU(inplace=True),
            conv3x3(N, N, stride=2))
        self.h_s = nn.Sequential(
            conv3x3(N, N), nn.LeakyReLU(inplace=True),
            subpel_conv3x3(N, N, 2), nn.LeakyReLU(inplace=True),
            conv3x3(N, N * 3 // 2), nn.LeakyReLU(inplace=True),
            subpel_conv3x3(N * 3 // 2, N * 3 // 2, 2), nn.LeakyReLU(inplace=True),
            conv3x3(N * 3 // 2, N * 2))
        self.g_s = nn.Sequential(
            ResidualBlock(N, N), ResidualBlockUpsample(N, N, 2),
            ResidualBlock(N, N), ResidualBlockUpsample(N, N, 2),
            ResidualBlock(N, N), ResidualBlockUpsample(N, N, 2),
            ResidualBlock(N, N), subpel_conv3x3(N, 3, 2))
        self.entropy_parameters = nn.Sequential(
            nn.Conv2d(M * 12 // 3, M * 10 // 3, 1), nn.LeakyReLU(inplace=True),
            nn.Conv2d(M * 10 // 3, M * 8 // 3, 1), nn.LeakyReLU(inplace=True),
            nn.Conv2d(M * 8 // 3, M * 6 // 3, 1))
        self.context_prediction = MaskedConv2d(M, 2 * M, kernel_size=5, padding=2, stride=1)
        self.gaussian_conditional = GaussianConditional(None)

    def aux_loss(self):
        return sum(m.loss() for m in self.modules() if isinstance(m, EntropyBottleneck))

    def forward(self, x):
        y = self.g_a(x)
        z = self.h_a(y)
        z_hat, z_lik = self.entropy_bottleneck(z)
        params = self.h_s(z_hat)
        y_hat = self.gaussian_conditional.quantize(y, "noise" if self.training else "dequantize")
        ctx = self.context_prediction(y_hat)
        gp = self.entropy_parameters(torch.cat((params, ctx), dim=1))
        scales_hat, means_hat = gp.chunk(2, 1)
        _, y_lik = self.gaussian_conditional(y, scales_hat, means=means_hat)
        x_hat = self.g_s(y_hat)
        # extras (y, z, y_hat, z_hat, scales, means) are for stage-by-stage parity checks only
        return {"x_hat": x_hat, "likelihoods": {"y": y_lik, "z": z_lik},
                "_taps": {"y": y, "z": z, "y_hat": y_hat, "z_hat": z_hat,
                          "scales_hat": scales_hat, "means_hat": means_hat}}
