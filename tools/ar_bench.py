"""Time / profile the autoregressive coding kernel alone at 1920x1024 (64x120 latent): python tools/ar_bench.py [cluster] [reps]."""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
from oracle.stats import build_oracle
from tdvc_b200 import coding, lib as L, synth
from tdvc_b200.model import VideoCompressor

cl = int(sys.argv[1]) if len(sys.argv) > 1 else 0
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda:0")
orc = build_oracle()
net = VideoCompressor().eval()
net.load_state_dict(orc.state_dict())
net = net.to(dev)
x, refs = synth.make_frame_pair(1024, 1920, seed=0)
with torch.no_grad():
    net(x.to(dev), refs.to(dev), False)
W = net._weights(dev)
plan = net._plan(1, 1024, 1920, dev)
tabs = coding.CoderTables(W, "mv", dev)
for rep in range(reps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    h = coding.launch_coding(plan, W, "mv", tabs, cluster=cl)
    b.record()
    torch.cuda.synchronize()
    print(f"cluster {cl}: eb_symbols + ar_code + D2H {a.elapsed_time(b):.2f} ms")
