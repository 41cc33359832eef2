"""Developer tool (GPU box): A/B of a plan-level switch inside ONE process - alternating blocks of GOP-chain frames, CUDA-event
timed - so that box-to-box and run-to-run drift (a few %) cancels.   python tools/ab_frame.py <attr> [blocks] [frames]"""
import sys
import warnings

import torch

warnings.filterwarnings("ignore")
sys.path.insert(0, ".")


def main(attr="overlap_cache", blocks=4, frames=44, amp=1):
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
    net.use_cuda_graph = True
    gops = [synth.make_gop(1024, 1920, gop=12, seed=100 + i).to(dev) for i in range(2)]
    plan = None

    def chain(n):
        refs = None
        for i in range(n):
            g = gops[(i // 11) % 2]
            t = i % 11 + 1
            if t == 1:
                refs = G.RefBuffer(g[0:1])
            win, keys = refs.window()
            recon, _, _ = net(g[t:t + 1], win, bool(amp), ref_keys=keys)
            refs.push(recon)

    chain(22)
    plan = net._plan(1, 1024, 1920, dev)
    res = {True: [], False: []}
    for b in range(2 * blocks):
        val = b % 2 == 0
        if attr.startswith("env:"):      # an environment switch read by the library at launch time: True = variable unset
            import os
            if val:
                os.environ.pop(attr[4:], None)
            else:
                os.environ[attr[4:]] = "1"
        else:
            setattr(plan, attr, val)
        plan.graphs = {}          # the switch changes the captured launch structure
        chain(22)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        chain(frames)
        e1.record()
        torch.cuda.synchronize()
        res[val].append(e0.elapsed_time(e1) / frames)
        print(f"{attr}={val}: {res[val][-1]:.3f} ms/frame", flush=True)
    for v in (True, False):
        print(f"{attr}={v}: mean {sum(res[v]) / len(res[v]):.3f} ms/frame over {len(res[v])} blocks")


if __name__ == "__main__":
    a = sys.argv[1:]
    main(a[0] if a else "overlap_cache", *[int(v) for v in a[1:]])
