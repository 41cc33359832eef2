"""Developer tool (GPU box): run the CUDA path and the oracle on the same seeded inputs and print the
per-stage max-abs / relative errors of every tap (SURVEY.md App. D dump points)."""
import sys
import time
import warnings

import torch

warnings.filterwarnings("ignore")
sys.path.insert(0, ".")


def main(h=64, w=64, seed=1, impl=1):
    from oracle.stats import build_oracle
    from tdvc_b200 import synth
    from tdvc_b200.model import VideoCompressor
    orc = build_oracle()
    net = VideoCompressor().eval()
    net.load_state_dict(orc.state_dict(), strict=True)
    net = net.cuda()
    net.conv_impl = impl
    x, refs = synth.make_frame_pair(h, w, seed=seed)
    ot, gt = {}, {}
    with torch.no_grad():
        t0 = time.time()
        o_recon, o_bres, o_bmv = orc(x, refs, False, taps=ot)
        t1 = time.time()
        g_recon, g_bres, g_bmv = net(x.cuda(), refs.cuda(), False, taps=gt)
        torch.cuda.synchronize()
    print(f"oracle {t1 - t0:.2f}s; launches {net.last_launches}")
    order = ["input_feat", "ref_feat", "motion_est.offset_l3", "motion_est.offset_l2", "motion_est.offset_l1"] + \
            [f"spynet.flow{i}" for i in range(6)] + ["estmv", "mv.y", "mv.z", "mv.z_hat", "mv.y_hat", "mv.scales_hat",
            "mv.means_hat", "mv.x_hat", "prediction1", "prediction", "input_residual", "res.y", "res.z", "res.z_hat",
            "res.y_hat", "res.scales_hat", "res.means_hat", "recon_feat", "loopfilter.f_in", "loopfilter.f_ref",
            "loopfilter.pool_in", "loopfilter.pool_ref", "loopfilter.sim", "loopfilter.ind", "loopfilter.gathered",
            "loopfilter.cor"]
    for k in order:
        if k not in ot or k not in gt:
            print(f"{k:28s} missing ({k in ot}, {k in gt})")
            continue
        a, b = ot[k].float(), gt[k].float().cpu()
        if a.shape != b.shape:
            print(f"{k:28s} SHAPE {tuple(a.shape)} vs {tuple(b.shape)}")
            continue
        d = (a - b).abs()
        extra = ""
        if k.endswith("_hat") and "y_hat" in k:
            extra = f" same {(a == b).float().mean().item():.6f}"
        if k.endswith("ind"):
            extra = f" same {(a == b).float().mean().item():.4f}"
        print(f"{k:28s} max|ref| {a.abs().max().item():10.4f} maxerr {d.max().item():.3e} meanerr {d.mean().item():.3e}{extra}")
    print("recon maxerr", (o_recon - g_recon.cpu()).abs().max().item())
    print("bpp_res", o_bres.item(), g_bres.item(), "bpp_mv", o_bmv.item(), g_bmv.item())


if __name__ == "__main__":
    a = [int(v) for v in sys.argv[1:]]
    main(*a)
