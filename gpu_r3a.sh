#!/bin/bash
# round-2 session-2, call 1: coding tests on the GPU + timing of the AR pass
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_coding.py -x -q -m gpu 2>&1 | tail -25 > gpurun_out/r3a_coding.log
cat gpurun_out/r3a_coding.log
timeout 600 python - <<'PY' 2>&1 | grep -v Warn | tee gpurun_out/r3a_time.log
import time, torch, sys
sys.path.insert(0, '.')
from oracle.stats import build_oracle
from tdvc_b200 import synth, coding
from tdvc_b200.model import VideoCompressor
dev = torch.device("cuda:0")
orc = build_oracle()
net = VideoCompressor().eval(); net.load_state_dict(orc.state_dict()); net = net.to(dev)
x, refs = synth.make_frame_pair(1024, 1920, seed=0); x, refs = x.to(dev), refs.to(dev)
with torch.no_grad():
    net(x, refs, False); net(x, refs, False, is_compress=True)
    torch.cuda.synchronize(); t = time.time(); net(x, refs, False); torch.cuda.synchronize(); t0 = time.time() - t
    t = time.time(); net(x, refs, False, is_compress=True); torch.cuda.synchronize(); t1 = time.time() - t
print("forward %.1f ms, forward+coding %.1f ms" % (t0 * 1e3, t1 * 1e3), {k: (v["ac_bpp"], [len(s[0]) for s in v["strings"]]) for k, v in net.last_coded.items()})
W = net._weights(dev); plan = net._plan(1, 1024, 1920, dev)
for cl in (8, 16):
    for cn in ("mv", "rs"):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); tt = time.time(); e0.record()
        coding.code_latents(plan, W, cn, W["_tables"][cn], cluster=cl)
        e1.record(); torch.cuda.synchronize()
        print("cluster", cl, cn, "code_latents wall %.2f ms" % ((time.time() - tt) * 1e3))
    # kernel alone
    import ctypes as C
PY
