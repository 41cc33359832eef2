"""ORACLE tool: print per-stage statistics of the conditioned synthetic workload (used once to tune the
gains in tdvc_b200/synth.py so the parity tests are not vacuous)."""
import sys
import time
import warnings

import torch

warnings.filterwarnings("ignore")


def build_oracle(conditioned=True):
    from oracle.model import VideoCompressor
    from tdvc_b200 import synth
    torch.manual_seed(synth.SEED)
    m = VideoCompressor().eval()
    if conditioned:
        sd = m.state_dict()
        synth.condition_state_dict(sd)
        m.load_state_dict(sd)
    return m


def main(h=256, w=256):
    from tdvc_b200 import synth
    m = build_oracle()
    x, refs = synth.make_frame_pair(h, w, seed=0)
    taps = {}
    t = time.time()
    with torch.no_grad():
        recon, bres, bmv = m(x, refs, False, taps=taps)
    print(f"forward {time.time() - t:.2f}s  bpp_res {bres.item():.4f} bpp_mv {bmv.item():.4f}")
    mse = ((recon - x) ** 2).mean().item()
    print("recon range", recon.min().item(), recon.max().item(), "mse vs input", mse,
          "frac clamped", ((recon <= 0) | (recon >= 1)).float().mean().item())
    for k in sorted(taps):
        v = taps[k].float()
        print(f"{k:28s} {tuple(v.shape)!s:22s} mean {v.mean().item():+.4f} std {v.std().item():.4f} "
              f"min {v.min().item():+.4f} max {v.max().item():+.4f}")
    for c in ("mv", "res"):
        yh, zh = taps[f"{c}.y_hat"], taps[f"{c}.z_hat"]
        print(c, "y_hat nonzero frac", (yh != 0).float().mean().item(), "unique", yh.unique().numel(),
              "| z_hat nonzero", (zh.round() != 0).float().mean().item(),
              "| scales<0.11 frac", (taps[f"{c}.scales_hat"] < 0.11).float().mean().item())


if __name__ == "__main__":
    main(*(int(a) for a in sys.argv[1:3]))
