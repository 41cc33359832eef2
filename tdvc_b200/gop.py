"""GOP evaluation driver — mirrors the reference's `validation` loop (reference tools/predict.py:43-100):
reference-frame buffer with warm-up duplication (:55-60), pad-to-64 / crop (reference
main/utils/utils.py:59-87), per-frame PSNR = 10 log10(1/MSE) (:87-88) and bpp = bpp_res + bpp_mv (:90).

Differences that do not change results: the per-frame `.cpu()` metric read-backs (one device sync per frame
per metric in the reference) become on-device accumulations (`tdvc_sq_err_sum`) read once per GOP, and GOPs
are sharded round-robin over ranks with one all-reduce of seven fp64 sums at the end.
"""
import math

import torch
import torch.nn.functional as F

from tdvc_b200 import lib as L


def pad(x, p=64):
    """reference main/utils/utils.py:59-73 (zero padding, centred)."""
    h, w = x.size(2), x.size(3)
    H, W = (h + p - 1) // p * p, (w + p - 1) // p * p
    left, top = (W - w) // 2, (H - h) // 2
    if H == h and W == w:
        return x
    return F.pad(x, (left, W - w - left, top, H - h - top), mode="constant", value=0)


def crop(x, size):
    """reference main/utils/utils.py:76-87."""
    H, W = x.size(2), x.size(3)
    h, w = size
    left, top = (W - w) // 2, (H - h) // 2
    if H == h and W == w:
        return x
    return x[:, :, top:top + h, left:left + w]


def reference_window(ref_list):
    """The 4 reference frames of the next P-frame: [I, x^(t-3), x^(t-2), x^(t-1)] with the reference's
    warm-up duplication (reference tools/predict.py:55-60).  Returns (N,4,3,H,W)."""
    n = len(ref_list)
    if n == 1:
        sel = [ref_list[0], ref_list[-1], ref_list[-1], ref_list[-1]]
    elif n == 2:
        sel = [ref_list[0], ref_list[-2], ref_list[-1], ref_list[-1]]
    else:
        sel = [ref_list[0], ref_list[-3], ref_list[-2], ref_list[-1]]
    return torch.stack(sel, dim=1)


class RefBuffer:
    """The reference-frame list of one GOP (reference tools/predict.py:51-68: `ref_image` grows by one reconstruction per
    P-frame) together with a host-side identity per frame.  `window()` returns the (N,4,3,H,W) tensor the codec takes and the four
    identities of its slices; passing those to `VideoCompressor.forward(ref_keys=...)` lets the per-GOP feature caches be keyed
    without hashing the frames on the device (one host synchronisation per frame less)."""
    _serial = 0

    def __init__(self, i_frame):
        self.frames, self.ids = [i_frame], [self._new_id()]

    @classmethod
    def _new_id(cls):
        cls._serial += 1
        return cls._serial

    def push(self, recon):
        self.frames.append(recon)
        self.ids.append(self._new_id())
        if len(self.frames) > 4:   # only the I-frame and the last three reconstructions are ever referenced again
            self.frames, self.ids = [self.frames[0]] + self.frames[-3:], [self.ids[0]] + self.ids[-3:]

    def window(self):
        n = len(self.frames)
        sel = (0, n - 1, n - 1, n - 1) if n == 1 else ((0, n - 2, n - 1, n - 1) if n == 2 else (0, n - 3, n - 2, n - 1))
        return torch.stack([self.frames[i] for i in sel], dim=1), tuple(self.ids[i] for i in sel)


def sq_err_sum(a, b, acc, slot):
    """acc[slot] += sum (a-b)^2 on the device (CUDA kernel; fp64 accumulator)."""
    a, b = a.contiguous(), b.contiguous()
    lib = L.load()
    L.check(lib.tdvc_sq_err_sum(a.data_ptr(), b.data_ptr(), a.numel(), acc.data_ptr() + 8 * slot,
                                torch.cuda.current_stream(a.device).cuda_stream), "sq_err_sum")


def code_gop(net, i_frame, p_frames, enable_amp=False, keep_recon=False, with_msssim=False):
    """Code one GOP.  i_frame (1,3,h,w): the (decoded) I-frame; p_frames (T,3,h,w): the frames to code.
    Returns dict with device tensors `bpp_mv`, `bpp_res` (T,), `sse` (T,) fp64 (sum of squared error over the
    cropped frame) and `numel`; plus `recon` (list of cropped reconstructions) when keep_recon."""
    dev = p_frames.device
    T, _, h, w = p_frames.shape
    refs = RefBuffer(pad(i_frame, 64))
    keyed = "ref_keys" in getattr(getattr(net, "forward", None), "__code__", type("c", (), {"co_varnames": ()})).co_varnames
    sse = torch.zeros(T, device=dev, dtype=torch.float64)
    bmv, bres, recons, mss = [], [], [], []
    for i in range(T):
        x = pad(p_frames[i:i + 1], 64)
        window, keys = refs.window()
        if keyed:   # tdvc_b200.VideoCompressor: the identities of the reference slices key its per-GOP feature caches
            recon, bpp_res, bpp_mv = net(x, window, enable_amp, ref_keys=keys)
        else:       # any module with the reference's signature (pnet.py:26)
            recon, bpp_res, bpp_mv = net(x, window, enable_amp)
        refs.push(recon)             # padded, clamped reconstruction feeds the next frame (predict.py:68)
        rc = crop(recon, (h, w))
        sq_err_sum(rc, p_frames[i:i + 1], sse, i)
        if with_msssim:              # tools/predict.py:93-94 (the enable_amp branch of the shipped cfg/predict.yaml)
            from tdvc_b200.metrics import ms_ssim
            mss.append(ms_ssim(rc.float(), p_frames[i:i + 1], data_range=1.0))
        bmv.append(bpp_mv.reshape(-1)[0])
        bres.append(bpp_res.reshape(-1)[0])
        if keep_recon:
            recons.append(rc.clone())
    out = {"bpp_mv": torch.stack(bmv), "bpp_res": torch.stack(bres), "sse": sse, "numel": 3 * h * w}
    if with_msssim:
        out["msssim"] = torch.stack(mss)
    if keep_recon:
        out["recon"] = recons
    return out


def validate(dataset, net, enable_amp=True, device=None, with_msssim=True, gops=None):
    """The reference's `validation` (reference tools/predict.py:35-110) over a `tdvc_b200.data.GopDataset`: every GOP is coded
    by `code_gop`; bpp / PSNR / MS-SSIM are averaged over ALL frames, the BPG-coded I-frame of each GOP included with the bpp /
    PSNR / MS-SSIM its loader reports (:47-50), while bpp_mv / bpp_res / mse average over the P-frames only (:95-97).
    Returns (bpp, bpp_mv, bpp_res, psnr, msssim, mse) like the reference (:110).  `gops`: indices to code (a rank's shard)."""
    device = device or next(net.parameters()).device
    tot = torch.zeros(7, device=device, dtype=torch.float64)
    n_i, i_bpp, i_psnr, i_ms = 0, 0.0, 0.0, 0.0
    for g in (range(len(dataset)) if gops is None else gops):
        p_frames, i_frame, bpp_i, psnr_i, _, raw = dataset[g]
        p_frames, i_frame = p_frames.to(device), i_frame.to(device).unsqueeze(0)
        res = code_gop(net, i_frame, p_frames, enable_amp=enable_amp, with_msssim=with_msssim)
        tot += gop_stats(res)
        n_i += 1
        i_bpp += float(bpp_i)
        i_psnr += float(psnr_i)
        if with_msssim:
            from tdvc_b200.metrics import ms_ssim
            i_ms += float(ms_ssim(raw[0:1].to(device), i_frame, data_range=1.0))
    return combine_with_iframes(summarise(tot), n_i, i_bpp, i_psnr, i_ms)


def combine_with_iframes(p_summary, n_i, sum_i_bpp, sum_i_psnr, sum_i_msssim):
    """Averages as the reference forms them (tools/predict.py:47-50, 86-110): the lists `bpps`, `psnrs`, `ssims` hold one entry per
    I-frame and per P-frame; `mvs`, `ress`, `mses` hold the P-frames only."""
    n_p = p_summary["frames"]
    n = max(n_p + n_i, 1)
    return ((p_summary["bpp"] * n_p + sum_i_bpp) / n, p_summary["bpp_mv"], p_summary["bpp_res"],
            (p_summary["psnr"] * n_p + sum_i_psnr) / n, (p_summary["msssim"] * n_p + sum_i_msssim) / n, p_summary["mse"])


def gop_stats(res):
    """Per-GOP sums in the reference's reporting units -> fp64 tensor
    [sum bpp, sum bpp_mv, sum bpp_res, sum psnr, sum msssim (0 unless code_gop(with_msssim=True)), sum mse, n_frames]."""
    mse = res["sse"] / res["numel"]
    psnr = 10.0 * torch.log10(1.0 / mse)
    bmv, bres = res["bpp_mv"].double(), res["bpp_res"].double()
    z = res["msssim"].double().sum() if "msssim" in res else torch.zeros((), device=mse.device, dtype=torch.float64)
    return torch.stack([(bmv + bres).sum(), bmv.sum(), bres.sum(), psnr.sum(), z, mse.sum(),
                        torch.tensor(float(mse.numel()), device=mse.device, dtype=torch.float64)])


def shard_gops(n_gops, rank, world):
    """GOPs are independent (each restarts from its own I-frame): rank r takes {g : g mod world == r}."""
    return [g for g in range(n_gops) if g % world == rank]


def reduce_stats(stats, group=None):
    """One all-reduce(sum) of the 7 fp64 sums over ranks (no-op without an initialised process group)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    return stats


def summarise(stats):
    s = [float(v) for v in stats.tolist()]
    n = max(s[6], 1.0)
    return {"bpp": s[0] / n, "bpp_mv": s[1] / n, "bpp_res": s[2] / n, "psnr": s[3] / n, "msssim": s[4] / n, "mse": s[5] / n,
            "frames": int(s[6])}


def psnr_from_mse(mse):
    return 10.0 * math.log10(1.0 / mse)
