"""Developer tool (GPU box): per-frame time of a resident GOP chain with (a) no feature cache, (b) cache keyed on the device
content hash (one host sync per frame), (c) cache keyed on caller-side identities (no sync)."""
import sys
import time
import warnings

import torch

warnings.filterwarnings("ignore")
sys.path.insert(0, ".")


def main(precision="exact", steps=33):
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
    net.precision, net.use_cuda_graph = precision, True
    gops = [synth.make_gop(1024, 1920, gop=12, seed=100 + i).to(dev) for i in range(2)]

    def chain(mode, n):
        net.cache_features = mode != "nocache"
        refs, ids = None, None
        k = 0
        for i in range(n):
            g = gops[(i // 11) % 2]
            t = i % 11 + 1
            if t == 1:
                refs, ids = [g[0:1]], [("I", i // 11)]
            keys = None
            if mode == "ids":
                n_ = len(ids)
                keys = [ids[0], ids[-1], ids[-1], ids[-1]] if n_ == 1 else ([ids[0], ids[-2], ids[-1], ids[-1]] if n_ == 2 else [ids[0], ids[-3], ids[-2], ids[-1]])
            recon, _, _ = net(g[t:t + 1], G.reference_window(refs), False, ref_keys=keys)
            refs.append(recon)
            ids.append(("P", i))
            if len(refs) > 4:
                refs, ids = [refs[0]] + refs[-3:], [ids[0]] + ids[-3:]

    for mode in ("nocache", "hash", "ids", "hash", "ids", "nocache"):
        chain(mode, 22)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        chain(mode, steps)
        e1.record()
        torch.cuda.synchronize()
        print(f"{precision} {mode:8s}: {e0.elapsed_time(e1) / steps:.3f} ms/frame (wall {(time.perf_counter() - t0) * 1e3 / steps:.3f})", flush=True)


if __name__ == "__main__":
    main(*sys.argv[1:2])
