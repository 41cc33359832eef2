"""ORACLE tool: generate the committed fixtures under tests/golden/ by running the reference's OWN
unmodified model code (oracle/ref_import.py) in this container, where /root/reference exists.

    python -m oracle.make_golden

For each case: synthetic frame pair (tdvc_b200.synth, seeded), conditioned weights (seed 1111), eval mode,
enabled_amp=False.  Stored: reconstruction, bpp_res, bpp_mv from the REFERENCE code; latent symbols
(y_hat, z_hat of both coders), the FeatureFix match indices and a few stage summaries from the
restatement oracle/model.py (which must agree bit-exactly with the reference on recon/bpp — asserted).
The GPU box has no /root/reference; tests there compare against these files and against oracle/model.py.
"""
import os
import warnings

import numpy as np
import torch

warnings.filterwarnings("ignore")

CASES = [("p64x64_s1", 64, 64, 1), ("p128x192_s2", 128, 192, 2)]
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main():
    from oracle import ref_import
    from oracle.stats import build_oracle
    from tdvc_b200 import synth
    torch.set_num_threads(1)  # fixed summation order inside MKL-DNN for reproducible fixtures
    orc = build_oracle()
    ref = ref_import.reference_video_compressor().eval()
    ref.load_state_dict(orc.state_dict(), strict=True)
    csum = synth.state_checksum(orc.state_dict())
    os.makedirs(OUT, exist_ok=True)
    for name, h, w, seed in CASES:
        x, refs = synth.make_frame_pair(h, w, seed=seed)
        taps = {}
        with torch.no_grad():
            r_recon, r_bres, r_bmv = ref(x, refs, False)
            o_recon, o_bres, o_bmv = orc(x, refs, False, taps=taps)
        assert torch.equal(r_recon, o_recon) and torch.equal(r_bres, o_bres) and torch.equal(r_bmv, o_bmv), \
            "oracle/model.py is not bit-exact with the reference code"
        d = dict(h=h, w=w, seed=seed, state_checksum=csum,
                 input_checksum=float(x.double().sum() + refs.double().sum()),
                 recon=r_recon.numpy(), bpp_res=r_bres.numpy(), bpp_mv=r_bmv.numpy(),
                 ind=taps["loopfilter.ind"].numpy().astype(np.int32))
        for c in ("mv", "res"):
            d[f"{c}_y_hat"] = taps[f"{c}.y_hat"].numpy().astype(np.int16)
            d[f"{c}_z_hat_minus_med"] = torch.round(taps[f"{c}.z_hat"] - getattr(
                orc, f"{c}Coder").entropy_bottleneck.quantiles[:, 0, 1].detach().view(1, -1, 1, 1)).numpy().astype(np.int16)
        for k in ("input_feat", "estmv", "mv.x_hat", "prediction1", "prediction", "input_residual",
                  "res.x_hat", "recon_feat", "spynet.flow5"):
            v = taps[k].double()
            d["stat_" + k] = np.array([v.mean().item(), v.abs().mean().item(), v.abs().max().item()])
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
        print(name, "bpp_res", r_bres.item(), "bpp_mv", r_bmv.item(), "recon mean", r_recon.mean().item())
    make_msssim_golden()


MSSSIM_CASES = [(0, (2, 3, 176, 200), 0.05), (1, (1, 3, 161, 175), 0.02), (2, (1, 3, 256, 320), 0.1)]


def msssim_inputs(seed, shape, noise):
    """Deterministic image pair in [0,1] for the MS-SSIM fixtures (also used by the tests)."""
    g = torch.Generator().manual_seed(4242 + seed)
    x = torch.rand(shape, generator=g)
    x = torch.nn.functional.avg_pool2d(x, 3, 1, 1)  # some spatial structure
    y = (x + noise * torch.randn(shape, generator=g)).clamp(0, 1)
    return x, y


def make_msssim_golden():
    """tests/golden/msssim.json: values of the REFERENCE's own ms_ssim (reference main/model/ms_ssim_torch.py:138-200)."""
    import importlib
    import json
    from oracle import ref_import
    ref_import.load_reference_pnet()
    R = importlib.import_module("main.model.ms_ssim_torch")
    out = []
    for seed, shape, noise in MSSSIM_CASES:
        x, y = msssim_inputs(seed, shape, noise)
        per = R.ms_ssim(x, y, data_range=1.0, size_average=False)
        out.append({"seed": seed, "shape": list(shape), "noise": noise, "per_image": [float(v) for v in per],
                    "mean": float(R.ms_ssim(x, y, data_range=1.0))})
        print("msssim", seed, shape, out[-1]["mean"])
    json.dump(out, open(os.path.join(OUT, "msssim.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
