"""Evaluation metrics on the device — drop-in for the reference's `main.model.ms_ssim_torch.ms_ssim`
(reference main/model/ms_ssim_torch.py:138-200, called from tools/predict.py:93-96 with data_range=1.0).

Same signature and the same arithmetic, including the reference's level weighting as written (`:193-196`: the last
level's SSIM term is broadcast over the four contrast-structure levels before the product, i.e. it enters with the
exponent 4*w_5).  (The other branch of the reference loop, `main.utils.utils.MS_SSIM` at tools/predict.py:96 - a 2-D
window with zero "same" padding - is a different estimator and is not provided here.)  Every scale is one fused CUDA kernel (`tdvc_ssim_level`: separable 11-tap Gaussian of x, y, x^2,
y^2, xy + SSIM / cs maps + their sums); nothing is read back to the host.  No CPU path."""
import ctypes as C
import math

import torch

from tdvc_b200 import lib as L

_WEIGHTS = (0.0448, 0.2856, 0.3001, 0.2363, 0.1333)


def _gauss_window(size, sigma):
    """reference ms_ssim_torch.py:5-18 (float32 arithmetic as torch does it)."""
    coords = torch.arange(size, dtype=torch.float) - size // 2
    g = torch.exp(-(coords ** 2) / (2 * sigma ** 2))
    return (g / g.sum()).tolist()


def ms_ssim(X, Y, win_size=11, win_sigma=1.5, win=None, data_range=255, size_average=True, full=False, weights=None):
    if X.dim() != 4:
        raise ValueError("Input images must 4-d tensor.")
    if X.type() != Y.type():
        raise ValueError("Input images must have the same dtype.")
    if X.shape != Y.shape:
        raise ValueError("Input images must have the same dimensions.")
    if not X.is_cuda:
        raise RuntimeError("tdvc_b200.metrics.ms_ssim runs on CUDA only (no CPU path)")
    if win is not None:
        g = [float(v) for v in win.reshape(win.shape[0], -1)[0].tolist()]
    else:
        if win_size % 2 != 1:
            raise ValueError("Window size must be odd.")
        g = _gauss_window(win_size, win_sigma)
    if len(g) != 11:
        raise RuntimeError("tdvc_b200.metrics.ms_ssim: the CUDA kernel is specialised for the reference's 11-tap window")
    w = list(_WEIGHTS) if weights is None else [float(v) for v in weights.tolist()]
    lib = L.load()
    dev = X.device
    N, Cc, H, W = X.shape
    c1, c2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    win_c = (C.c_float * 11)(*g)
    with torch.cuda.device(dev):
        st = torch.cuda.current_stream(dev).cuda_stream
        x, y = X.detach().float().contiguous(), Y.detach().float().contiguous()
        levels = len(w)
        acc = torch.zeros(levels, N, 2, device=dev, dtype=torch.float64)
        counts = []
        for lv in range(levels):
            h, wd = x.shape[2], x.shape[3]
            L.check(lib.tdvc_ssim_level(x.data_ptr(), y.data_ptr(), N, Cc, h, wd, win_c, c1, c2, acc[lv].data_ptr(), st),
                    "ssim_level")
            counts.append(Cc * (h - 10) * (wd - 10))
            if lv + 1 < levels:  # the reference also pools after the last level; the result is unused
                py, px = h % 2, wd % 2
                ho, wo = (h + 2 * py - 2) // 2 + 1, (wd + 2 * px - 2) // 2 + 1
                x2 = torch.empty(N, Cc, ho, wo, device=dev)
                y2 = torch.empty(N, Cc, ho, wo, device=dev)
                L.check(lib.tdvc_avgpool2_pad(x.data_ptr(), x2.data_ptr(), N * Cc, h, wd, st), "avgpool2_pad")
                L.check(lib.tdvc_avgpool2_pad(y.data_ptr(), y2.data_ptr(), N * Cc, h, wd, st), "avgpool2_pad")
                x, y = x2, y2
        cnt = torch.tensor(counts, device=dev, dtype=torch.float64).view(levels, 1, 1)
        mean = (acc / cnt).float()                    # (level, batch, [ssim, cs]) as fp32 like the reference's .mean()
        ssim_val = (mean[-1, :, 0] + 1) / 2            # "avoid Nan" (:79-81)
        mcs = (mean[:, :, 1] + 1) / 2                  # (level, batch)
        wt = torch.tensor(w, device=dev, dtype=torch.float32)
        val = torch.prod((mcs[:-1] ** wt[:-1].unsqueeze(1)) * (ssim_val ** wt[-1]), dim=0)   # (batch,)
    if size_average:
        val = val.mean()
    return val
