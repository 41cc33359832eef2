"""Developer tool (GPU box): code ONE 1920x1024 P-frame eagerly (no CUDA graph) inside a cudaProfilerStart/Stop
range after warm-up frames, so that `ncu --profile-from-start off` sees exactly one frame's launches.

    python tools/profile_frame.py [H W [conv_impl]]
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches.csv python tools/profile_frame.py
"""
import sys
import warnings

import torch

warnings.filterwarnings("ignore")
sys.path.insert(0, ".")


def main(h=1024, w=1920, impl=0, amp=1):
    from tdvc_b200 import gop as G
    from tdvc_b200 import synth
    from tdvc_b200.model import VideoCompressor
    dev = torch.device("cuda:0")
    torch.manual_seed(synth.SEED)
    net = VideoCompressor().eval()
    sd = net.state_dict()
    synth.condition_state_dict(sd)
    net.load_state_dict(sd)
    net = net.to(dev)
    net.conv_impl = impl
    # frame 7 of a GOP is the steady state: I-frame features and two of the three fusion fronts come from the per-GOP caches
    g = synth.make_gop(h, w, gop=8, seed=100).to(dev)
    refs = G.RefBuffer(g[0:1])
    with torch.no_grad():
        for t in range(1, 7):
            win, keys = refs.window()
            recon, _, _ = net(g[t:t + 1], win, bool(amp), ref_keys=keys)
            refs.push(recon)
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
        win, keys = refs.window()
        recon, bres, bmv = net(g[7:8], win, bool(amp), ref_keys=keys)
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
    print(f"profiled 1 P-frame {h}x{w}: launches {net.last_launches}, bpp_res {bres.item():.4f} bpp_mv {bmv.item():.4f}")


if __name__ == "__main__":
    main(*[int(a) for a in sys.argv[1:]])
