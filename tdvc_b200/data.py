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
