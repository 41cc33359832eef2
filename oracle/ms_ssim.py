"""ORACLE (test infrastructure): CPU restatement of the reference's MS-SSIM
(reference main/model/ms_ssim_torch.py:5-18 window, :21-87 `_ssim`, :138-200 `ms_ssim`), written with explicit
separable sums so that it does not share code with the reference's F.conv2d formulation.  Pinned bit-for-tolerance
against the reference's own function (imported verbatim from /root/reference) in tests/test_oracle.py and through the
committed fixture tests/golden/msssim.json (made by oracle/make_golden.py from the reference function)."""
import torch
import torch.nn.functional as F

WEIGHTS = (0.0448, 0.2856, 0.3001, 0.2363, 0.1333)


def gauss_window(size=11, sigma=1.5):
    coords = torch.arange(size, dtype=torch.float) - size // 2          # :13-14
    g = torch.exp(-(coords ** 2) / (2 * sigma ** 2))                   # :16
    return g / g.sum()                                                  # :17


def _filter(img, g):
    """valid-padding separable filter, horizontal then vertical (:31-33)."""
    k = g.numel()
    H, W = img.shape[-2:]
    hz = sum(g[i] * img[..., :, i:W - k + 1 + i] for i in range(k))
    return sum(g[i] * hz[..., i:H - k + 1 + i, :] for i in range(k))


def ssim_level(X, Y, g, data_range):
    """(ssim_val, cs) per batch element of one scale (:51-87 with size_average=False, full=True)."""
    C1, C2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    mu1, mu2 = _filter(X, g), _filter(Y, g)
    s1 = _filter(X * X, g) - mu1 * mu1
    s2 = _filter(Y * Y, g) - mu2 * mu2
    s12 = _filter(X * Y, g) - mu1 * mu2
    cs_map = (2 * s12 + C2) / (s1 + s2 + C2)
    ssim_map = ((2 * mu1 * mu2 + C1) / (mu1 * mu1 + mu2 * mu2 + C1)) * cs_map
    return (ssim_map.mean(dim=(1, 2, 3)) + 1) / 2, (cs_map.mean(dim=(1, 2, 3)) + 1) / 2


def ms_ssim(X, Y, data_range=255, size_average=True):
    g = gauss_window()
    w = torch.tensor(WEIGHTS, dtype=X.dtype)
    mcs = []
    ssim_val = None
    for _ in range(len(WEIGHTS)):
        ssim_val, cs = ssim_level(X, Y, g, data_range)
        mcs.append(cs)
        pad = (X.shape[2] % 2, X.shape[3] % 2)                          # :188-190
        X = F.avg_pool2d(X, kernel_size=2, padding=pad)
        Y = F.avg_pool2d(Y, kernel_size=2, padding=pad)
    mcs = torch.stack(mcs, dim=0)
    # :193-196 as written: the last level's SSIM term is broadcast over the four cs levels before the product
    val = torch.prod((mcs[:-1] ** w[:-1].unsqueeze(1)) * (ssim_val ** w[-1]), dim=0)
    return val.mean() if size_average else val
