"""GPU tests (-m gpu): end-to-end parity of the CUDA P-frame forward (through VideoCompressor.forward, i.e. the
C-ABI kernels) against the oracle (restatement pinned bit-exact to the reference code) and against the committed
golden fixtures produced by the reference's own unmodified code (tests/golden, oracle/make_golden.py).

Bars (BASELINE.json north_star): reconstruction <= 1e-3 max-abs, PSNR within 0.01 dB, bpp within 0.1 % relative,
>= 99.9 % of quantised latent symbols identical.  The exact-fp32 path (conv_impl=1) and the tensor-core path
(conv_impl=0: fp16 hi/lo-split tcgen05 MMA, fp32 accumulate) are both held to the same bars.
"""
import math

import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def net(oracle_model, dev):
    from tdvc_b200.model import VideoCompressor
    m = VideoCompressor().eval()
    m.load_state_dict(oracle_model.state_dict(), strict=True)
    return m.to(dev)


def _psnr(a, b):
    return 10.0 * math.log10(1.0 / ((a - b) ** 2).mean().item())


def _check_frame(net, oracle_model, dev, h, w, seed, impl):
    from tdvc_b200 import synth
    x, refs = synth.make_frame_pair(h, w, seed=seed)
    ot, gt = {}, {}
    net.conv_impl = impl
    with torch.no_grad():
        o_recon, o_bres, o_bmv = oracle_model(x, refs, False, taps=ot)
        g_recon, g_bres, g_bmv = net(x.to(dev), refs.to(dev), False, taps=gt)
    g_recon = g_recon.cpu()
    same = {}
    for c in ("mv", "res"):
        same[c + ".y"] = (ot[f"{c}.y_hat"] == gt[f"{c}.y_hat"].cpu()).float().mean().item()
        same[c + ".z"] = (ot[f"{c}.z_hat"] == gt[f"{c}.z_hat"].cpu()).float().mean().item()
    assert min(same.values()) >= 0.999, same
    # mismatches only at rounding ties: the pre-quantiser value must sit next to a half-integer
    for c in ("mv", "res"):
        bad = ot[f"{c}.y_hat"] != gt[f"{c}.y_hat"].cpu()
        if bad.any():
            frac = (ot[f"{c}.y"][bad] - torch.floor(ot[f"{c}.y"][bad]) - 0.5).abs()
            assert frac.max() < 5e-3, f"{c}: symbol mismatch away from a rounding tie ({frac.max().item()})"
    assert torch.equal(ot["loopfilter.ind"], gt["loopfilter.ind"].cpu())
    assert abs(o_bres.item() - g_bres.item()) <= 1e-3 * o_bres.item()
    assert abs(o_bmv.item() - g_bmv.item()) <= 1e-3 * o_bmv.item()
    if min(same.values()) == 1.0:
        assert (o_recon - g_recon).abs().max().item() <= 1e-3
    # PSNR of the reconstruction against the source frame: within 0.01 dB
    assert abs(_psnr(o_recon, x) - _psnr(g_recon, x)) <= 0.01
    assert g_recon.min() >= 0.0 and g_recon.max() <= 1.0
    return o_recon, g_recon, same


@pytest.mark.parametrize("impl", [1, 0])
@pytest.mark.parametrize("case", [(64, 64, 1), (128, 192, 2), (64, 256, 7)])
def test_pframe_forward_vs_oracle(net, oracle_model, dev, case, impl):
    _check_frame(net, oracle_model, dev, *case, impl)


def test_pframe_forward_vs_oracle_multi_tile(net, oracle_model, dev):
    """256x320: several 32- and 64-row items per column of tiles on every pyramid level (the row-phase 7x7 SPyNet layers
    walk four 64-row items), 128-channel split tiles at 1/2 scale, FeatureFix with more than one candidate block."""
    _check_frame(net, oracle_model, dev, 256, 320, 3, 0)


@pytest.mark.parametrize("impl", [1, 0])
@pytest.mark.parametrize("name", ["p64x64_s1", "p128x192_s2"])
def test_pframe_forward_vs_reference_golden(net, dev, name, impl):
    """Fixtures were produced by the reference's own code (oracle/make_golden.py); nothing from oracle/ runs here."""
    from tdvc_b200 import synth
    g = load_golden(name)
    assert abs(synth.state_checksum(net.state_dict()) - float(g["state_checksum"])) < 1e-6 * float(g["state_checksum"])
    x, refs = synth.make_frame_pair(int(g["h"]), int(g["w"]), seed=int(g["seed"]))
    taps = {}
    net.conv_impl = impl
    with torch.no_grad():
        recon, bres, bmv = net(x.to(dev), refs.to(dev), False, taps=taps)
    same = []
    for c in ("mv", "res"):
        same.append((taps[f"{c}.y_hat"].cpu().numpy().astype(np.int16) == g[f"{c}_y_hat"]).mean())
    assert min(same) >= 0.999, same
    assert (taps["loopfilter.ind"].cpu().numpy().astype(np.int32) == g["ind"]).all()
    assert abs(bres.item() - float(g["bpp_res"][0])) <= 1e-3 * float(g["bpp_res"][0])
    assert abs(bmv.item() - float(g["bpp_mv"][0])) <= 1e-3 * float(g["bpp_mv"][0])
    if min(same) == 1.0:
        assert np.abs(recon.cpu().numpy() - g["recon"]).max() <= 1e-3


@pytest.mark.parametrize("impl", [1, 0])
def test_batch_of_two_vs_oracle(net, oracle_model, dev, impl):
    """N = 2 (the reference's DataParallel / training batches, bpp aggregated over the batch: reference pnet.py:38-43,82-83):
    two different frame pairs in one call against the oracle on the same batch."""
    from tdvc_b200 import synth
    xa, ra = synth.make_frame_pair(64, 128, seed=11)
    xb, rb = synth.make_frame_pair(64, 128, seed=12)
    x, refs = torch.cat([xa, xb], 0), torch.cat([ra, rb], 0)
    ot, gt = {}, {}
    net.conv_impl = impl
    with torch.no_grad():
        o = oracle_model(x, refs, False, taps=ot)
        g = net(x.to(dev), refs.to(dev), False, taps=gt)
    net.conv_impl = 0
    for c in ("mv", "res"):
        assert (ot[f"{c}.y_hat"] == gt[f"{c}.y_hat"].cpu()).float().mean().item() >= 0.999
        assert (ot[f"{c}.z_hat"] == gt[f"{c}.z_hat"].cpu()).float().mean().item() >= 0.999
    assert torch.equal(ot["loopfilter.ind"], gt["loopfilter.ind"].cpu())
    assert (o[0] - g[0].cpu()).abs().max().item() <= 1e-3
    assert abs(o[1].item() - g[1].item()) <= 1e-3 * o[1].item() and abs(o[2].item() - g[2].item()) <= 1e-3 * o[2].item()
    # and the batch result equals the two single-frame results (recon) / their mean (bpp)
    with torch.no_grad():
        ga = net(xa.to(dev), ra.to(dev), False)
        gb = net(xb.to(dev), rb.to(dev), False)
    # (not bit-identical: the batched launches tile / reduce in a different order)
    assert (g[0][0:1] - ga[0]).abs().max().item() <= 1e-4 and (g[0][1:2] - gb[0]).abs().max().item() <= 1e-4
    assert abs(g[1].item() - 0.5 * (ga[1].item() + gb[1].item())) <= 1e-4 * g[1].item()


@pytest.mark.parametrize("impl", [1, 0])
def test_config5_fusion_and_inloop_filter_vs_oracle(net, oracle_model, dev, impl):
    """BASELINE config 5: multi-frame feature fusion + reference-based in-loop filter with 4 reference frames, fed with the
    oracle's own intermediate tensors (prediction1, recon_feat) of a synthetic frame pair."""
    from tdvc_b200 import synth
    x, refs = synth.make_frame_pair(128, 192, seed=5)
    taps = {}
    with torch.no_grad():
        oracle_model(x, refs, False, taps=taps)
        pred1, recf = taps["prediction1"], taps["recon_feat"]
        want_pred = oracle_model.mcfilter(pred1, refs)
        want_recon = oracle_model.loopfilter(recf, refs).clamp(0.0, 1.0)
    net.conv_impl = impl
    got_pred, got_recon = net.fusion_and_filter(pred1.to(dev), refs.to(dev), recf.to(dev))
    net.conv_impl = 0
    assert (got_pred.cpu() - want_pred).abs().max().item() <= 1e-4 * max(1.0, want_pred.abs().max().item())
    assert (got_recon.cpu() - want_recon).abs().max().item() <= 1e-4
    assert net.last_launches > 40


def test_default_init_degenerate_case(dev):
    """Module default init: all symbols 0, DCN offsets exactly 0 - still must agree."""
    from oracle.stats import build_oracle
    from tdvc_b200 import synth
    from tdvc_b200.model import VideoCompressor
    orc = build_oracle(conditioned=False)
    m = VideoCompressor().eval()
    m.load_state_dict(orc.state_dict(), strict=True)
    m = m.to(dev)
    x, refs = synth.make_frame_pair(64, 64, seed=3)
    with torch.no_grad():
        a = orc(x, refs, False)
        b = m(x.to(dev), refs.to(dev), False)
    assert (a[0] - b[0].cpu()).abs().max() < 1e-4
    assert abs(a[1].item() - b[1].item()) < 1e-3 * a[1].item() and abs(a[2].item() - b[2].item()) < 1e-3 * a[2].item()


def test_gop_chain_vs_oracle(net, oracle_model, dev):
    """Three chained P-frames through the GOP driver (reference tools/predict.py:51-68 semantics, incl. pad/crop
    for a frame size that is not a multiple of 64): PSNR within 0.01 dB and bpp within 0.1 % per frame."""
    from tdvc_b200 import gop as G
    from tdvc_b200 import synth
    net.conv_impl = 0
    frames = synth.make_gop(56, 120, gop=4, seed=21)
    res = G.code_gop(net, frames[0:1].to(dev), frames[1:].to(dev), keep_recon=True)
    refs = [G.pad(frames[0:1], 64)]
    for i in range(3):
        x = G.pad(frames[i + 1:i + 2], 64)
        with torch.no_grad():
            recon, bres, bmv = oracle_model(x, G.reference_window(refs), False)
        refs.append(recon)
        rc = G.crop(recon, (56, 120))
        mse = ((rc - frames[i + 1:i + 2]) ** 2).mean().item()
        g_mse = res["sse"][i].item() / res["numel"]
        assert abs(10 * math.log10(1 / mse) - 10 * math.log10(1 / g_mse)) <= 0.01
        assert abs(bres.item() - res["bpp_res"][i].item()) <= 1e-3 * bres.item()
        assert abs(bmv.item() - res["bpp_mv"][i].item()) <= 1e-3 * bmv.item()
    st = G.summarise(G.gop_stats(res))
    assert st["frames"] == 3 and st["bpp"] > 0 and 5 < st["psnr"] < 60


def test_cuda_graph_equals_eager_and_is_deterministic(net, dev):
    from tdvc_b200 import synth
    x, refs = synth.make_frame_pair(128, 128, seed=4)
    x, refs = x.to(dev), refs.to(dev)
    net.conv_impl = 0
    net.use_cuda_graph = False
    a = net(x, refs, False)
    b = net(x, refs, False)
    net.use_cuda_graph = True
    try:
        c = net(x, refs, False)
        d = net(x, refs, False)
    finally:
        net.use_cuda_graph = False
    for p, q, r, s in zip(a, b, c, d):
        assert torch.equal(p, q) and torch.equal(p, r) and torch.equal(p, s)
    assert net.last_launches > 200


def test_full_size_properties(net, dev):
    """1920x1024 (BASELINE config 2), where the oracle takes ~90 s: size-independent properties only -
    run-to-run bit-identical, tensor-core path vs exact-fp32 path within the parity bars, translation check of
    the in-loop filter indices' range, finite bpp."""
    from tdvc_b200 import synth
    g = synth.make_gop(1024, 1920, gop=2, seed=0).to(dev)
    x, refs = g[1:2], g[0:1].unsqueeze(1).expand(-1, 4, -1, -1, -1).contiguous()
    net.conv_impl = 0
    t0 = {}
    a = net(x, refs, False, taps=t0)
    b = net(x, refs, False)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and torch.equal(a[2], b[2])
    net.conv_impl = 1
    t1 = {}
    c = net(x, refs, False, taps=t1)
    net.conv_impl = 0
    for k in ("mv.y_hat", "res.y_hat", "mv.z_hat", "res.z_hat"):
        assert (t0[k] == t1[k]).float().mean().item() >= 0.999, k
    assert torch.equal(t0["loopfilter.ind"], t1["loopfilter.ind"])
    assert abs(a[1].item() - c[1].item()) <= 1e-3 * c[1].item() and abs(a[2].item() - c[2].item()) <= 1e-3 * c[2].item()
    assert abs(_psnr(a[0], x) - _psnr(c[0], x)) <= 0.01
    assert math.isfinite(a[1].item()) and a[1].item() > 0 and a[0].min() >= 0 and a[0].max() <= 1
