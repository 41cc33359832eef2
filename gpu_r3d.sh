#!/bin/bash
mkdir -p gpurun_out
timeout 600 python - <<'PY' 2>&1 | grep -v Warn | tee gpurun_out/r3d.log
import torch, sys, numpy as np
sys.path.insert(0, '.')
from oracle.stats import build_oracle
from tdvc_b200 import synth
from tdvc_b200.model import VideoCompressor
dev = torch.device("cuda:0")
orc = build_oracle()
net = VideoCompressor().eval(); net.load_state_dict(orc.state_dict()); net = net.to(dev)
x, refs = synth.make_frame_pair(128, 192, seed=2)
gt = {}
with torch.no_grad():
    net(x.to(dev), refs.to(dev), False, is_compress=True, taps=gt)
plan = net._plan(1, 128, 192, dev)
out = {}
for nm, cn in (("mv", "mv"), ("res", "rs")):
    out[nm + "_y"] = plan.buf(f"{cn}.y", 1, 8, 12, 128).t.cpu().numpy()
    out[nm + "_params"] = plan.buf(f"{cn}.params", 1, 8, 12, 256).t.cpu().numpy()
    for k in ("y_symbols", "y_indexes", "y_hat"):
        out[nm + "_" + k] = gt[f"{nm}.ac.{k}"].cpu().numpy()
np.savez_compressed("gpurun_out/r3d_dump.npz", **out)
print("dumped")
PY
