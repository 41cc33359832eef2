"""Evaluation data pipeline — the GOP datasets of the reference (reference main/dataloader/dataset.py:16-98 `UVGDataSet`,
:100-193 `HEVCDataSet`) and its I+P averaging (reference tools/predict.py:43-110), without OpenCV / natsort:

    <root>/ori_img/<sequence>/im001.png ...                       the raw frames
    <root>/compress_img_bpg/<sequence>/<qp>/im001_<qp>.png|.txt   the BPG-coded I-frame of every GOP and its bpp

The I-frames are produced offline by the reference's preprocessing (tools/preprocess/02-04*.py: BPG at
qp 22/27/32/37 for lambda 4096/2048/1024/512, dataset.py:25-36); BPG itself is outside this package.
`GopDataset[i]` returns what the reference's loader returns for GOP i; `tdvc_b200.gop.validate` consumes it.
"""
import glob
import math
import os
import re

import torch

HEVC_CLASSES = {   # reference dataset.py:109-123
    "A": ("2560x1600", ["Traffic", "PeopleOnStreet"]),
    "B": ("1920x1080", ["ParkScene", "Kimono1", "Cactus", "BasketballDrive", "BQTerrace"]),
    "C": ("832x480", ["BasketballDrill", "BQMall", "PartyScene", "RaceHorses"]),
    "D": ("416x240", ["BasketballPass", "BQSquare", "BlowingBubbles", "RaceHorses"]),
    "E": ("1280x720", ["vidyo1", "vidyo3", "vidyo4"]),
}


def natural_key(s):
    """natsort-style key: digit runs compare as numbers (reference uses natsort.natsorted, dataset.py:41,144)."""
    return [int(t) if t.isdigit() else t for t in re.split(r"(\d+)", s)]


def qp_for_lambda(train_lambda):
    """reference dataset.py:25-36: the BPG quantiser of the I-frames that goes with a rate-distortion lambda."""
    for lams, qp in (((512, 16), 37), ((1024, 32), 32), ((2048, 64), 27), ((4096, 128), 22)):
        if train_lambda in lams:
            return qp
    raise ValueError(f"no BPG qp is defined for train_lambda={train_lambda} (reference dataset.py:25-36)")


def read_rgb(path):
    """(3,h,w) float32 in [0,1] (reference: cv2.imread + BGR2RGB, /255)."""
    from torchvision.io import ImageReadMode, read_image
    return read_image(path, ImageReadMode.RGB).float() / 255.0


class GopDataset(torch.utils.data.Dataset):
    """reference UVGDataSet (`hevc_class=None`) / HEVCDataSet (`hevc_class` in A..E) in evaluation mode (isTrain=False)."""

    def __init__(self, root, train_lambda, gop_size, testfull=False, hevc_class=None):
        self.input_path, self.ref_path = os.path.join(root, "ori_img"), os.path.join(root, "compress_img_bpg")
        self.gop_size = gop_size
        self.ref, self.refbpp, self.input = [], [], []
        qp = qp_for_lambda(train_lambda)
        folders = os.listdir(self.input_path)
        folders = sorted(folders, key=natural_key) if hevc_class is None else folders   # HEVCDataSet keeps listdir order
        for folder in folders:
            seq = folder.rstrip()
            if hevc_class is not None:
                res, names = HEVC_CLASSES[hevc_class]
                parts = seq.split("_")
                if len(parts) < 2 or parts[0] not in names or parts[1] != res:
                    continue
            imgs = sorted(glob.glob(os.path.join(self.input_path, seq, "*.png")), key=natural_key)
            n_gops = len(imgs) // gop_size if testfull else 8
            for i in range(n_gops):
                stem = os.path.join(self.ref_path, seq, str(qp), f"im{i * gop_size + 1:03d}_{qp}")
                with open(stem + ".txt", "r", encoding="utf-8") as f:
                    self.refbpp.append(float(f.read().splitlines()[0]))
                self.ref.append(stem + ".png")
                self.input.append([os.path.join(self.input_path, seq, f"im{i * gop_size + j + 1:03d}.png") for j in range(gop_size)])

    def __len__(self):
        return len(self.ref)

    def __getitem__(self, index):
        """-> (p_frames (T-1,3,h,w), i_frame (3,h,w) BPG-decoded, i_bpp, i_psnr, names, raw (T,3,h,w)).  The reference also returns
        the MS-SSIM of the I-frame, computed on the host inside the loader; here `gop.validate` computes it on the device."""
        i_frame = read_rgb(self.ref[index])
        h, w = i_frame.shape[1:]
        raw = torch.stack([read_rgb(p)[:, :h, :w] for p in self.input[index]])
        mse = ((raw[0] - i_frame) ** 2).mean().item()
        i_psnr = 10.0 * math.log10(1.0 / mse) if mse > 0 else 100.0    # reference CalcuPSNR
        return raw[1:], i_frame, self.refbpp[index], i_psnr, self.input[index], raw


class VimeoDataset(torch.utils.data.Dataset):
    """The reference's training set `DataSet` (reference main/dataloader/dataset.py:203-258) over a vimeo_septuplet tree
    `<root>/<dir>/<clip>/im1.png ... im7.png`, without OpenCV / natsort / albumentations.

    Samples per clip (`get_vimeo`, :210-247): for every target frame t = 2..n the references are
    `[im1, im(t-3), im(t-2), im(t-1)]` clipped at im1 - the first slot is always the clip's first frame (the I-frame FeatureFix
    matches against, reference pnet.py:212), a short history is filled by repeating its last entry - plus one long-range sample
    `[im1, im1, im3, im5] -> im7`.  `__getitem__` returns `(input (3,S,S), refs (4,3,S,S))` float32 in [0,1] after the
    augmentation of `imgauglist2` (reference main/dataloader/augmentation.py:29-85): horizontal flip p=0.5, vertical flip p=0.4,
    with p=0.5 one of {per-channel RGB shift of up to +-20/255, brightness/contrast change of up to +-20 %}, then with p=0.5 a
    random S x S crop, otherwise torchvision's RandomResizedCrop(S, scale=(0.5, 1)) - every draw shared by the five frames of a
    sample.  Same distributions as the reference's transforms, drawn from `torch` generators (not the same random stream).
    """

    def __init__(self, dataset_path, resize_size, generator=None):
        self.size = int(resize_size)
        self.generator = generator
        self.image_input_list, self.image_ref_list = self.get_vimeo(dataset_path)

    @staticmethod
    def get_vimeo(dataset_path):
        inputs, refs = [], []
        for d in sorted(os.listdir(dataset_path), key=natural_key):
            for clip in sorted(os.listdir(os.path.join(dataset_path, d)), key=natural_key):
                base = os.path.join(dataset_path, d, clip)
                name = lambda i: os.path.join(base, f"im{i}.png")
                n = len(glob.glob(os.path.join(base, "*.png")))
                for t in range(2, n + 1):
                    hist = [name(1)] + [name(i) for i in range(max(t - 3, 1), t)]
                    hist += [hist[-1]] * (4 - len(hist))
                    refs.append(hist)
                    inputs.append(name(t))
                refs.append([name(1), name(1), name(3), name(5)])
                inputs.append(name(7))
        return inputs, refs

    def __len__(self):
        return len(self.image_input_list)

    def _rand(self, *shape):
        return torch.rand(*shape, generator=self.generator)

    def augment(self, frames):
        """frames: (5,3,h,w) float32 in [0,1], target first -> (5,3,S,S)."""
        S = self.size
        if self._rand(1).item() < 0.5:
            frames = frames.flip(-1)
        if self._rand(1).item() < 0.4:
            frames = frames.flip(-2)
        if self._rand(1).item() < 0.5:
            if self._rand(1).item() < 0.5:     # albumentations RGBShift defaults: each channel shifted by U(-20, 20) of 255
                shift = (self._rand(3) * 40.0 - 20.0) / 255.0
                frames = (frames + shift.view(1, 3, 1, 1)).clamp(0.0, 1.0)
            else:                              # RandomBrightnessContrast defaults: x * (1 + U(-.2,.2)) + U(-.2,.2) * max
                alpha, beta = 1.0 + (self._rand(1).item() * 0.4 - 0.2), self._rand(1).item() * 0.4 - 0.2
                frames = (frames * alpha + beta).clamp(0.0, 1.0)
            frames = torch.round(frames * 255.0) / 255.0      # the reference augments uint8 images
        h, w = frames.shape[-2:]
        if self._rand(1).item() < 0.5:         # RandomSizedCrop([S,S], S, S): an S x S window at a random position
            if h < S or w < S:
                raise RuntimeError(f"frames of {h}x{w} are smaller than the {S}x{S} training crop")
            top = int(self._rand(1).item() * (h - S + 1))
            left = int(self._rand(1).item() * (w - S + 1))
            return frames[..., top:top + S, left:left + S].contiguous()
        # torchvision RandomResizedCrop((S,S), scale=(0.5,1.0)): area fraction U(0.5,1), log-uniform aspect in (3/4, 4/3), ten
        # attempts, then the central crop; bilinear resize of the stack (one window for the five frames)
        area = h * w
        for _ in range(10):
            target = area * (0.5 + 0.5 * self._rand(1).item())
            log_r = math.log(3.0 / 4.0) + self._rand(1).item() * (math.log(4.0 / 3.0) - math.log(3.0 / 4.0))
            ratio = math.exp(log_r)
            cw, ch = int(round(math.sqrt(target * ratio))), int(round(math.sqrt(target / ratio)))
            if 0 < cw <= w and 0 < ch <= h:
                top = int(self._rand(1).item() * (h - ch + 1))
                left = int(self._rand(1).item() * (w - cw + 1))
                break
        else:
            ratio = w / h
            if ratio < 3.0 / 4.0:
                cw, ch = w, int(round(w / (3.0 / 4.0)))
            elif ratio > 4.0 / 3.0:
                ch, cw = h, int(round(h * (4.0 / 3.0)))
            else:
                cw, ch = w, h
            top, left = (h - ch) // 2, (w - cw) // 2
        crop_ = frames[..., top:top + ch, left:left + cw]
        return torch.nn.functional.interpolate(crop_, size=(S, S), mode="bilinear", align_corners=False, antialias=True)

    def __getitem__(self, index):
        frames = torch.stack([read_rgb(self.image_input_list[index])] + [read_rgb(p) for p in self.image_ref_list[index]])
        out = self.augment(frames)
        return out[0], out[1:]
