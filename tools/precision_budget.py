"""Developer tool (CPU, uses the oracle): per-stage precision budget of the tensor-core convolutions (SURVEY.md 7.2-1,
VERDICT r01 item 2).

Emulates a ONE-product fp16 MMA (both operands rounded to fp16, fp32 accumulation) inside the oracle for one group of
stages at a time - every conv2d / conv3d executed while a module of the group is running gets fp16-rounded inputs and
weights - and reports, against the exact fp32 oracle on the same frame pair:
  symbol flips of both coders (this frame), reconstruction max-abs / rms error, bpp relative error, FeatureFix index
  equality, and the symbol flips of the NEXT frame of a chain when this frame's reconstruction is its reference.

    python tools/precision_budget.py [H W [seed]]  > profiles/r02_precision_budget.txt
"""
import sys
import warnings

import torch
import torch.nn.functional as F

warnings.filterwarnings("ignore")
sys.path.insert(0, ".")

_state = {"on": 0, "mode": "1p"}
_conv2d, _conv3d = F.conv2d, F.conv3d


def _rnd(t):
    return t.half().float()


def _patched(orig):
    def f(x, w, *a, **kw):
        if _state["on"] > 0:
            if _state["mode"] == "1p":        # fp16(x) * fp16(w)
                x, w = _rnd(x), _rnd(w)
            elif _state["mode"] == "xhi":     # fp16(x) * w   (two products: w_hi + w_lo)
                x = _rnd(x)
            elif _state["mode"] == "bf16":
                x, w = x.bfloat16().float(), w.bfloat16().float()
        return orig(x, w, *a, **kw)
    return f


F.conv2d, F.conv3d = _patched(_conv2d), _patched(_conv3d)
torch.conv2d, torch.conv3d = F.conv2d, F.conv3d   # compressai_port / oracle call through F.*; nn.Conv2d uses F.conv2d


def _hook(mods):
    hs = []
    for m in mods:
        hs.append(m.register_forward_pre_hook(lambda *_: _state.__setitem__("on", _state["on"] + 1)))
        hs.append(m.register_forward_hook(lambda *_: _state.__setitem__("on", _state["on"] - 1)))
    return hs


def groups(m):
    lf = m.loopfilter
    g = {
        "res.g_s": [m.resCoder.g_s],
        "ff.post_argmax (featfusion, featfusion2, recon_layer, featdown)": [lf.featfusion, lf.featfusion2, lf.recon_layer, lf.featdown],
        "ff.FeatureExtract_input": [lf.FeatureExtract_input],
        "ff.FeatureExtract_ref": [lf.FeatureExtract_ref],
        "hyper tail (h_s, ctx, entropy_parameters; both coders)": [m.mvCoder.h_s, m.mvCoder.context_prediction, m.mvCoder.entropy_parameters,
                                                                  m.resCoder.h_s, m.resCoder.context_prediction, m.resCoder.entropy_parameters],
        "res.g_s + ff.post_argmax": [m.resCoder.g_s, lf.featfusion, lf.featfusion2, lf.recon_layer, lf.featdown],
        "all post-quantiser (res.g_s + whole FeatureFix)": [m.resCoder.g_s, lf],
        "mv.g_s": [m.mvCoder.g_s],
        "mcnet + mcfilter": [m.mcnet, m.mcfilter],
        "extra_fea": [m.extra_fea],
        "motion_est (OffsetGen + SPyNet)": [m.motion_est],
        "g_a (both coders)": [m.mvCoder.g_a, m.resCoder.g_a],
        "everything": [m],
    }
    return g


def run(m, x, refs):
    taps = {}
    with torch.no_grad():
        out = m(x, refs, False, taps=taps)
    return out, taps


def flips(a, b):
    return {k: (a[k] != b[k]).float().mean().item() for k in ("mv.y_hat", "mv.z_hat", "res.y_hat", "res.z_hat")}


def main(h=256, w=320, seed=3, modes=("1p", "xhi")):
    from oracle.stats import build_oracle
    from tdvc_b200 import gop as G
    from tdvc_b200 import synth
    torch.set_num_threads(8)
    m = build_oracle()
    frames = synth.make_gop(h, w, gop=6, seed=seed)
    x, refs = frames[4:5], torch.stack([frames[0], frames[1], frames[2], frames[3]]).unsqueeze(0)
    x2 = frames[5:6]
    (r0, bres0, bmv0), t0 = run(m, x, refs)
    refs2 = torch.stack([frames[0], frames[2], frames[3], r0[0]]).unsqueeze(0)
    (r1, _, _), t1 = run(m, x2, refs2)
    print(f"# frame {h}x{w} seed {seed}; exact oracle bpp_res {bres0.item():.5f} bpp_mv {bmv0.item():.5f}")
    print("# columns: flipped symbols this frame (mv.y mv.z res.y res.z) | recon max-abs, rms | bpp_res rel, bpp_mv rel | ind equal |"
          " next-frame flips with this recon as x^(t-1) (mv.y res.y) | next-frame recon max-abs")
    for mode in modes:
        _state["mode"] = mode
        print(f"## mode {mode}: " + {"1p": "fp16(x) * fp16(w), one MMA product", "xhi": "fp16(x) * w (two products)",
                                     "bf16": "bf16(x) * bf16(w)"}[mode])
        for name, mods in groups(m).items():
            hs = _hook(mods)
            (r, bres, bmv), t = run(m, x, refs)
            for hh in hs:
                hh.remove()
            assert _state["on"] == 0
            fl = flips(t0, t)
            d = (r - r0).abs()
            ind_eq = torch.equal(t["loopfilter.ind"], t0["loopfilter.ind"])
            # next frame, exact arithmetic, but with the perturbed reconstruction as reference
            refs2p = torch.stack([frames[0], frames[2], frames[3], r[0]]).unsqueeze(0)
            (r1p, _, _), t1p = run(m, x2, refs2p)
            fl2 = flips(t1, t1p)
            print(f"{name:70s} | {fl['mv.y_hat']:.5f} {fl['mv.z_hat']:.5f} {fl['res.y_hat']:.5f} {fl['res.z_hat']:.5f} | "
                  f"{d.max().item():.2e} {d.pow(2).mean().sqrt().item():.2e} | "
                  f"{abs(bres.item() - bres0.item()) / bres0.item():.2e} {abs(bmv.item() - bmv0.item()) / bmv0.item():.2e} | {ind_eq} | "
                  f"{fl2['mv.y_hat']:.5f} {fl2['res.y_hat']:.5f} | {(r1p - r1).abs().max().item():.2e}", flush=True)




def chain(h=256, w=320, seed=21, n_p=8, group="all post-quantiser (res.g_s + whole FeatureFix)", mode="1p"):
    """Free-running GOP chain: exact oracle vs the oracle with `group` emulated at `mode`; per frame the symbol flips,
    PSNR difference (dB) and bpp relative difference."""
    import math
    from oracle.stats import build_oracle
    from tdvc_b200 import gop as G
    from tdvc_b200 import synth
    torch.set_num_threads(8)
    m = build_oracle()
    frames = synth.make_gop(h, w, gop=n_p + 1, seed=seed)
    _state["mode"] = mode
    ra, rb = [frames[0:1]], [frames[0:1]]
    print(f"# chain {h}x{w} seed {seed}, group '{group}' at mode {mode}: frame | flips mv.y mv.z res.y res.z | dPSNR dB | bpp rel | recon max-abs, frac > 1e-3")
    for t in range(1, n_p + 1):
        x = frames[t:t + 1]
        (a, abres, abmv), ta = run(m, x, G.reference_window(ra))
        hs = _hook(groups(m)[group])
        (b, bbres, bbmv), tb = run(m, x, G.reference_window(rb))
        for hh in hs:
            hh.remove()
        ra.append(a)
        rb.append(b)
        fl = flips(ta, tb)
        pa = 10 * math.log10(1 / ((a - x) ** 2).mean().item())
        pb = 10 * math.log10(1 / ((b - x) ** 2).mean().item())
        bp_a, bp_b = abres.item() + abmv.item(), bbres.item() + bbmv.item()
        d = (a - b).abs()
        print(f"{t} | {fl['mv.y_hat']:.5f} {fl['mv.z_hat']:.5f} {fl['res.y_hat']:.5f} {fl['res.z_hat']:.5f} | {pb - pa:+.5f} | "
              f"{abs(bp_b - bp_a) / bp_a:.2e} | {d.max().item():.2e} {(d > 1e-3).float().mean().item():.4f} | ind equal "
              f"{torch.equal(ta['loopfilter.ind'], tb['loopfilter.ind'])}", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "chain":
        chain(*[int(v) for v in sys.argv[2:6]])
    else:
        main(*[int(v) for v in sys.argv[1:4]])
