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

from oracle import rans


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
        # CDF tables of compress()/decompress(), filled by update()
        self.register_buffer("_offset", torch.IntTensor())
        self.register_buffer("_quantized_cdf", torch.IntTensor())
        self.register_buffer("_cdf_length", torch.IntTensor())
        self.entropy_coder_precision = 16

    def _load_from_state_dict(self, state_dict, prefix, *args):
        """CompressAI resizes these buffers on load (`update_registered_buffers`): a state_dict saved after update() fits."""
        for name in ("_offset", "_quantized_cdf", "_cdf_length", "scale_table"):
            t = state_dict.get(prefix + name)
            buf = getattr(self, name, None)
            if t is not None and buf is not None and buf.shape != t.shape:
                setattr(self, name, torch.empty(t.shape, dtype=buf.dtype))
        super()._load_from_state_dict(state_dict, prefix, *args)

    @staticmethod
    def quantize(inputs, mode, means=None):
        if mode == "noise":
            return inputs + torch.empty_like(inputs).uniform_(-0.5, 0.5)
        outputs = inputs.clone()
        if means is not None:
            outputs -= means
        outputs = torch.round(outputs)  # half-to-even
        if mode == "dequantize":
            if means is not None:
                outputs += means
            return outputs
        assert mode == "symbols", mode
        return outputs.int()

    @staticmethod
    def dequantize(inputs, means=None):
        if means is not None:
            outputs = inputs.type_as(means)
            outputs += means
        else:
            outputs = inputs.float()
        return outputs

    def _pmf_to_cdf(self, pmf, tail_mass, pmf_length, max_length):
        """CompressAI `EntropyModel._pmf_to_cdf`: one quantised CDF row per channel / scale, zero-padded."""
        cdf = torch.zeros((len(pmf_length), max_length + 2), dtype=torch.int32)
        for i, p in enumerate(pmf):
            prob = torch.cat((p[: pmf_length[i]], tail_mass[i]), dim=0)
            row = rans.pmf_to_quantized_cdf(prob.tolist(), self.entropy_coder_precision)
            cdf[i, : len(row)] = torch.IntTensor(row)
        return cdf

    def _tables(self):
        return (self._quantized_cdf.tolist(), self._cdf_length.reshape(-1).int().tolist(),
                self._offset.reshape(-1).int().tolist())

    def compress(self, inputs, indexes, means=None):
        """CompressAI `EntropyModel.compress`: one rANS string per batch item, symbols in memory (NCHW) order."""
        symbols = self.quantize(inputs, "symbols", means)
        cdf, lens, offs = self._tables()
        return [rans.encode_with_indexes(symbols[i].reshape(-1).tolist(), indexes[i].reshape(-1).int().tolist(),
                                         cdf, lens, offs) for i in range(symbols.size(0))]

    def decompress(self, strings, indexes, means=None):
        cdf, lens, offs = self._tables()
        outputs = torch.empty(indexes.size(), dtype=torch.float32)
        for i, s in enumerate(strings):
            v = rans.decode_with_indexes(s, indexes[i].reshape(-1).int().tolist(), cdf, lens, offs)
            outputs[i] = torch.tensor(v, dtype=torch.float32).reshape(outputs[i].size())
        return self.dequantize(outputs, means)


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

    def update(self, force=False):
        """CompressAI `EntropyBottleneck.update`: per-channel pmf over [median - minima, median + maxima] from the
        learned quantiles, evaluated with the cumulative MLP, quantised to 16 bits."""
        if self._offset.numel() > 0 and not force:
            return False
        medians = self.quantiles[:, 0, 1]
        minima = torch.clamp(torch.ceil(medians - self.quantiles[:, 0, 0]).int(), min=0)
        maxima = torch.clamp(torch.ceil(self.quantiles[:, 0, 2] - medians).int(), min=0)
        self._offset = -minima
        pmf_start = medians - minima
        pmf_length = maxima + minima + 1
        max_length = pmf_length.max().item()
        samples = torch.arange(max_length)
        samples = samples[None, :] + pmf_start[:, None, None]
        lower = self._logits_cumulative(samples - 0.5, stop_gradient=True)
        upper = self._logits_cumulative(samples + 0.5, stop_gradient=True)
        sign = -torch.sign(lower + upper)
        pmf = torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower))
        pmf = pmf[:, 0, :]
        tail_mass = torch.sigmoid(lower[:, 0, :1]) + torch.sigmoid(-upper[:, 0, -1:])
        self._quantized_cdf = self._pmf_to_cdf(pmf.detach(), tail_mass.detach(), pmf_length, max_length)
        self._cdf_length = pmf_length + 2
        return True

    def _build_indexes(self, size):
        N, C = size[0], size[1]
        return torch.arange(C).view(1, -1, 1, 1).int().expand(N, C, *size[2:])

    def _medians(self, x):
        return self.quantiles[:, :, 1:2].detach().reshape(1, -1, 1, 1).expand(x.size(0), -1, 1, 1)

    def compress(self, x):
        return super().compress(x, self._build_indexes(x.size()), self._medians(x))

    def decompress(self, strings, size):
        out_size = (len(strings), self._quantized_cdf.size(0), size[0], size[1])
        idx = self._build_indexes(out_size)
        return super().decompress(strings, idx, self._medians(idx))


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

    def update_scale_table(self, scale_table, force=False):
        """CompressAI `GaussianConditional.update_scale_table` + `update`: one zero-mean Gaussian pmf per scale."""
        if self._offset.numel() > 0 and not force:
            return False
        self.scale_table = torch.Tensor(tuple(float(s) for s in scale_table))
        key = (tuple(self.scale_table.tolist()), self.tail_mass)
        if key in _GC_TABLE_CACHE:   # the table depends on nothing but (scale_table, tail_mass): built once per process
            self._quantized_cdf, self._offset, self._cdf_length = (t.clone() for t in _GC_TABLE_CACHE[key])
            return True
        multiplier = -standardized_quantile(self.tail_mass / 2)
        pmf_center = torch.ceil(self.scale_table * multiplier).int()
        pmf_length = 2 * pmf_center + 1
        max_length = torch.max(pmf_length).item()
        samples = torch.abs(torch.arange(max_length).int() - pmf_center[:, None]).float()
        samples_scale = self.scale_table.unsqueeze(1).float()
        upper = self._standardized_cumulative((0.5 - samples) / samples_scale)
        lower = self._standardized_cumulative((-0.5 - samples) / samples_scale)
        pmf = upper - lower
        tail_mass = 2 * lower[:, :1]
        self._quantized_cdf = self._pmf_to_cdf(pmf, tail_mass, pmf_length, max_length)
        self._offset = -pmf_center
        self._cdf_length = pmf_length + 2
        _GC_TABLE_CACHE[key] = (self._quantized_cdf.clone(), self._offset.clone(), self._cdf_length.clone())
        return True

    def build_indexes(self, scales):
        scales = self.lower_bound_scale(scales)
        indexes = scales.new_full(scales.size(), len(self.scale_table) - 1).int()
        for s in self.scale_table[:-1]:
            indexes -= (scales <= s).int()
        return indexes


SCALES_MIN, SCALES_MAX, SCALES_LEVELS = 0.11, 256, 64
_GC_TABLE_CACHE = {}


def get_scale_table(lo=SCALES_MIN, hi=SCALES_MAX, levels=SCALES_LEVELS):
    """CompressAI `models.priors.get_scale_table`."""
    return torch.exp(torch.linspace(math.log(lo), math.log(hi), levels))


def standardized_quantile(q):
    """scipy.stats.norm.ppf(q), as CompressAI's `_standardized_quantile`."""
    import scipy.stats
    return float(scipy.stats.norm.ppf(q))


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

    # ----- real entropy coding (reference pnet.py:45-49,69-73: `update(force=True)` then `compress(x)`)
    def update(self, scale_table=None, force=False):
        """CompressAI `JointAutoregressiveHierarchicalPriors.update` (-> MeanScaleHyperprior / CompressionModel)."""
        if scale_table is None:
            scale_table = get_scale_table()
        updated = self.gaussian_conditional.update_scale_table(scale_table, force=force)
        updated |= self.entropy_bottleneck.update(force=force)
        return updated

    def _ar_step(self, y_hat, params, h, w, masked_weight):
        """Gaussian parameters of latent position (h, w) from the causal 5x5 window of y_hat and the hyper-decoder output."""
        y_crop = y_hat[:, :, h:h + 5, w:w + 5]
        ctx_p = F.conv2d(y_crop, masked_weight, bias=self.context_prediction.bias)
        p = params[:, :, h:h + 1, w:w + 1]
        gp = self.entropy_parameters(torch.cat((p, ctx_p), dim=1)).squeeze(3).squeeze(2)
        scales_hat, means_hat = gp.chunk(2, 1)
        return y_crop, scales_hat, means_hat

    def ar_code(self, y, params):
        """CompressAI `_compress_ar` for every image of the batch: raster scan, each position quantised relative to the mean
        its causal 5x5 context predicts.  -> (strings, symbols, table indexes (lists in (h, w, c) order), y_hat)."""
        H, W = y.size(2), y.size(3)
        y_pad = F.pad(y, (2, 2, 2, 2))
        gc = self.gaussian_conditional
        cdf, lens, offs = gc._tables()
        masked_weight = self.context_prediction.weight * self.context_prediction.mask
        y_strings, all_syms, all_idx = [], [], []
        for i in range(y.size(0)):
            y_hat = y_pad[i:i + 1]
            syms, idxs = [], []
            for h in range(H):
                for w in range(W):
                    y_crop, scales_hat, means_hat = self._ar_step(y_hat, params[i:i + 1], h, w, masked_weight)
                    indexes = gc.build_indexes(scales_hat)
                    y_q = gc.quantize(y_crop[:, :, 2, 2], "symbols", means_hat)
                    y_hat[:, :, h + 2, w + 2] = y_q + means_hat
                    syms.extend(y_q.reshape(-1).tolist())
                    idxs.extend(indexes.reshape(-1).tolist())
            y_strings.append(rans.encode_with_indexes(syms, idxs, cdf, lens, offs))
            all_syms.append(syms)
            all_idx.append(idxs)
        return y_strings, all_syms, all_idx, y_pad[:, :, 2:-2, 2:-2].clone()

    def compress(self, x, taps=None):
        """CompressAI `JointAutoregressiveHierarchicalPriors.compress` / `_compress_ar`: z through the factorised
        prior; y raster-scanned, each position quantised RELATIVE to its context-predicted mean, all of an image's
        symbols ((h, w, c) order) in one rANS string."""
        y = self.g_a(x)
        z = self.h_a(y)
        z_strings = self.entropy_bottleneck.compress(z)
        z_hat = self.entropy_bottleneck.decompress(z_strings, z.size()[-2:])
        params = self.h_s(z_hat)
        y_strings, all_syms, all_idx, y_hat = self.ar_code(y, params)
        H, W = y.size(2), y.size(3)
        if taps is not None:  # (N, H, W, C) symbol / index arrays and the dequantised latent, for the parity tests
            C = y.size(1)
            taps.update({"y_symbols": torch.tensor(all_syms, dtype=torch.int32).reshape(-1, H, W, C),
                         "y_indexes": torch.tensor(all_idx, dtype=torch.int32).reshape(-1, H, W, C),
                         "y_hat": y_hat, "y": y, "z": z, "z_hat": z_hat, "params": params,
                         "z_symbols": self.entropy_bottleneck.quantize(
                             z, "symbols", self.entropy_bottleneck._medians(z))})
        return {"strings": [y_strings, z_strings], "shape": z.size()[-2:]}

    def decompress(self, strings, shape):
        """CompressAI `JointAutoregressiveHierarchicalPriors.decompress` / `_decompress_ar` (the reference never calls
        it - pnet.py only measures the coded size - it is here to prove that the strings decode)."""
        z_hat = self.entropy_bottleneck.decompress(strings[1], shape)
        params = self.h_s(z_hat)
        H, W = z_hat.size(2) * 4, z_hat.size(3) * 4
        gc = self.gaussian_conditional
        cdf, lens, offs = gc._tables()
        y_hat = torch.zeros((z_hat.size(0), self.M, H + 4, W + 4))
        masked_weight = self.context_prediction.weight * self.context_prediction.mask
        for i, s in enumerate(strings[0]):
            dec = rans.Decoder(s)
            for h in range(H):
                for w in range(W):
                    _, scales_hat, means_hat = self._ar_step(y_hat[i:i + 1], params[i:i + 1], h, w, masked_weight)
                    indexes = gc.build_indexes(scales_hat)
                    rv = dec.decode_stream(indexes.reshape(-1).tolist(), cdf, lens, offs)
                    rv = torch.tensor(rv, dtype=torch.float32).reshape(1, -1)
                    y_hat[i, :, h + 2, w + 2] = gc.dequantize(rv, means_hat)[0]
        y_hat = y_hat[:, :, 2:-2, 2:-2]
        return {"x_hat": self.g_s(y_hat), "y_hat": y_hat}
