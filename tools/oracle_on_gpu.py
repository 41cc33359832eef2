"""Developer tool (GPU box): the oracle (plain PyTorch restatement of the reference forward) in eager mode on the B200,
fp32 with and without TF32 convolutions, and under autocast fp16 as the reference's predict.py runs it (enable_amp: True).
Reported in DESIGN.md as context only: it is neither the parity oracle's platform (CPU) nor the bench's reference arm."""
import sys
import time
import warnings

import torch

sys.path.insert(0, ".")
warnings.filterwarnings("ignore")


def main(h=1024, w=1920, iters=5):
    from oracle.stats import build_oracle
    from tdvc_b200 import synth
    dev = torch.device("cuda:0")
    orc = build_oracle().to(dev).eval()
    x, refs = synth.make_frame_pair(h, w, seed=3)
    x, refs = x.to(dev), refs.to(dev)
    for name, tf32, amp in (("fp32 (TF32 off)", False, False), ("fp32 convs with TF32", True, False), ("autocast fp16", True, True)):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        with torch.no_grad():
            for _ in range(2):
                with torch.autocast("cuda", enabled=amp):
                    orc(x, refs, amp)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(iters):
                with torch.autocast("cuda", enabled=amp):
                    out = orc(x, refs, amp)
            torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) / iters * 1e3
        print(f"oracle on B200, {name}: {ms:.1f} ms per {w}x{h} P-frame ({1e3 / ms:.2f} P-frames/s), bpp_res {out[1].item():.4f}")


if __name__ == "__main__":
    main(*[int(a) for a in sys.argv[1:]])
