"""Developer tool (GPU box): BASELINE config 5 — multi-frame feature fusion + reference-based in-loop filter with 4
reference frames at 1920x1024 through VideoCompressor.fusion_and_filter (1,438,138 MAC/px = 5.65 TFLOP per call)."""
import sys
import torch
sys.path.insert(0, ".")
from tdvc_b200 import synth
from tdvc_b200.model import VideoCompressor


def main(h=1024, w=1920, iters=10):
    dev = torch.device("cuda:0")
    torch.manual_seed(synth.SEED)
    net = VideoCompressor().eval()
    sd = net.state_dict()
    synth.condition_state_dict(sd)
    net.load_state_dict(sd)
    net = net.to(dev)
    g = synth.make_gop(h, w, gop=4, seed=0).to(dev)
    refs = g.unsqueeze(0).contiguous()                       # (1, 4, 3, H, W)
    pred1 = torch.randn(1, 64, h, w, device=dev) * 0.5
    recf = torch.randn(1, 64, h, w, device=dev) * 0.5
    for _ in range(3):
        net.fusion_and_filter(pred1, refs, recf)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        net.fusion_and_filter(pred1, refs, recf)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    macs = 1438138 * h * w
    print(f"config 5 @{w}x{h}: {ms:.2f} ms per call, {2 * macs / ms / 1e9:.1f} TFLOP/s algorithmic, {net.last_launches} launches")


if __name__ == "__main__":
    main(*[int(a) for a in sys.argv[1:]])
