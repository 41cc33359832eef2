"""ORACLE tool: 1920x1024 fixtures (BASELINE configs 2 and 5) from the reference's OWN unmodified model code
(oracle/ref_import.py), run once in the build container where /root/reference exists (~25 min on 8 threads).

    python -m oracle.make_golden_fullres [pair] [chain]

Two fixtures, both with the conditioned weights (seed 1111) the small fixtures use:

  tests/golden/p1024x1920_s0.npz      one P-frame with four DISTINCT raw reference frames (synth.make_frame_pair,
        seed 0): full reconstruction (uint16, x 65535), both bpp, every quantised symbol of both coders (int8), the
        FeatureFix match indices, and - for BASELINE config 5 (multi-frame fusion + in-loop filter at full size) -
        stride-32 samples and 64x64-tile means of prediction1 / prediction (mcfilter output) / recon_feat.
  tests/golden/chain1024x1920_s100.npz  six chained P-frames of the GOP bench.py codes first (synth.make_gop seed 100,
        reference window with warm-up duplication, reference tools/predict.py:51-68): per frame the symbols, indices, bpp,
        MSE against the source frame, a stride-4 sample of the reconstruction (uint16) and its 64x64-tile means.

Reconstruction and bpp come from the REFERENCE code; symbols / indices / stage tensors from the restatement
oracle/model.py run on the same inputs, which is asserted bit-exact with the reference on recon and bpp for every frame.
"""
import os
import sys
import time
import warnings

import numpy as np
import torch

warnings.filterwarnings("ignore")

H, W = 1024, 1920
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
STAGES = ("prediction1", "prediction", "recon_feat", "input_residual", "estmv", "mv.x_hat")


def _q16(x):
    return torch.round(x.clamp(0, 1) * 65535.0).to(torch.int32).numpy().astype(np.uint16)


def _tile_mean(x, t=64):
    return torch.nn.functional.avg_pool2d(x.double(), t).numpy()


def _symbols(d, taps, orc, prefix=""):
    for c in ("mv", "res"):
        yh = taps[f"{c}.y_hat"]
        med = getattr(orc, f"{c}Coder").entropy_bottleneck.quantiles[:, 0, 1].detach().view(1, -1, 1, 1)
        zq = torch.round(taps[f"{c}.z_hat"] - med)
        for nme, v in ((f"{c}_y_hat", yh), (f"{c}_z_hat_minus_med", zq)):
            assert v.abs().max() <= 32767
            dt = np.int8 if v.abs().max() <= 127 else np.int16
            d[prefix + nme] = v.numpy().astype(dt)
    d[prefix + "ind"] = taps["loopfilter.ind"].numpy().astype(np.int32)


def _run_both(ref, orc, x, refs):
    taps = {}
    with torch.no_grad():
        t0 = time.time()
        r = ref(x, refs, False)
        t1 = time.time()
        o = orc(x, refs, False, taps=taps)
        t2 = time.time()
    assert torch.equal(r[0], o[0]) and torch.equal(r[1], o[1]) and torch.equal(r[2], o[2]), \
        "oracle/model.py is not bit-exact with the reference code at 1920x1024"
    print(f"  reference {t1 - t0:.0f} s, restatement {t2 - t1:.0f} s, bpp_res {r[1].item():.5f} bpp_mv {r[2].item():.5f}", flush=True)
    return r, taps


def make_pair(ref, orc, csum):
    from tdvc_b200 import synth
    x, refs = synth.make_frame_pair(H, W, seed=0)
    (recon, bres, bmv), taps = _run_both(ref, orc, x, refs)
    d = dict(h=H, w=W, seed=0, state_checksum=csum, input_checksum=float(x.double().sum() + refs.double().sum()),
             recon_q16=_q16(recon), bpp_res=bres.numpy(), bpp_mv=bmv.numpy(),
             mse=float(((recon.double() - x.double()) ** 2).mean()))
    _symbols(d, taps, orc)
    for k in STAGES:
        v = taps[k]
        d["s32_" + k] = v[:, :, ::32, ::32].contiguous().numpy()
        d["tile_" + k] = _tile_mean(v).astype(np.float32)
        d["stat_" + k] = np.array([v.double().mean().item(), v.double().abs().mean().item(), v.abs().max().item()])
    np.savez_compressed(os.path.join(OUT, "p1024x1920_s0.npz"), **d)
    print("pair written", flush=True)


def make_chain(ref, orc, csum, n_p=6, seed=100):
    from tdvc_b200 import gop as G
    from tdvc_b200 import synth
    frames = synth.make_gop(H, W, gop=n_p + 1, seed=seed)
    d = dict(h=H, w=W, seed=seed, n_p=n_p, state_checksum=csum, input_checksum=float(frames.double().sum()))
    refs = [frames[0:1]]
    for t in range(1, n_p + 1):
        x = frames[t:t + 1]
        print(f"chain frame {t}", flush=True)
        (recon, bres, bmv), taps = _run_both(ref, orc, x, G.reference_window(refs))
        refs.append(recon)
        if len(refs) > 4:
            refs = [refs[0]] + refs[-3:]
        p = f"f{t}_"
        d[p + "bpp_res"], d[p + "bpp_mv"] = bres.numpy(), bmv.numpy()
        d[p + "mse"] = float(((recon.double() - x.double()) ** 2).mean())
        d[p + "recon_s4_q16"] = _q16(recon[:, :, ::4, ::4].contiguous())
        d[p + "recon_tile"] = _tile_mean(recon).astype(np.float32)
        _symbols(d, taps, orc, p)
        del taps
        np.savez_compressed(os.path.join(OUT, f"chain1024x1920_s{seed}.npz"), **d)   # rewritten after every frame
    print("chain written", flush=True)


def main(which):
    from oracle import ref_import
    from oracle.stats import build_oracle
    from tdvc_b200 import synth
    torch.set_num_threads(os.cpu_count() or 1)
    orc = build_oracle()
    ref = ref_import.reference_video_compressor().eval()
    ref.load_state_dict(orc.state_dict(), strict=True)
    csum = synth.state_checksum(orc.state_dict())
    os.makedirs(OUT, exist_ok=True)
    if not which or "pair" in which:
        make_pair(ref, orc, csum)
    if not which or "chain" in which:
        make_chain(ref, orc, csum)


if __name__ == "__main__":
    main(sys.argv[1:])
