#!/bin/bash
mkdir -p gpurun_out
timeout 600 python - <<'PY' 2>&1 | grep -v Warn | tee gpurun_out/r3b_time.log
import time, torch, sys, numpy as np, ctypes as C
sys.path.insert(0, '.')
from oracle.stats import build_oracle
from tdvc_b200 import synth, coding, lib as L
from tdvc_b200.model import VideoCompressor
dev = torch.device("cuda:0")
orc = build_oracle()
net = VideoCompressor().eval(); net.load_state_dict(orc.state_dict()); net = net.to(dev)
W = net._weights(dev)
t = None
pass
x, refs = synth.make_frame_pair(1024, 1920, seed=0); x, refs = x.to(dev), refs.to(dev)
with torch.no_grad():
    net(x, refs, False, is_compress=True)
plan = net._plan(1, 1024, 1920, dev)
lib = L.load()
orig = lib.tdvc_ar_code
for cl in (8, 16):
    for cn in ("mv", "rs"):
        tabs = W["_tables"][cn]
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        # time pieces: monkeypatch-free: call code_latents pieces by hand
        st = torch.cuda.current_stream(dev).cuda_stream
        e0.record()
        t0 = time.time()
        out = coding.code_latents(plan, W, cn, tabs, cluster=cl, keep=True)
        t1 = time.time()
        sy, ix = out["y_symbols"].cpu().numpy(), out["y_indexes"].cpu().numpy()
        t2 = time.time(); s = coding.rans_encode(sy[0], ix[0], tabs.gc); t3 = time.time()
        print(f"cluster {cl} {cn}: code_latents {1e3*(t1-t0):.1f} ms, host rANS of y alone {1e3*(t3-t2):.1f} ms, {len(s)} bytes")
# kernel-only timing through events: launch ar_code directly
from tdvc_b200.coding import SCALES_LEVELS
for cl in (8, 16):
    cn = "mv"; tabs = W["_tables"][cn]
    hy, wy = 64, 120
    y, params = plan.buf(f"{cn}.y", 1, hy, wy, 128), plan.buf(f"{cn}.params", 1, hy, wy, 256)
    ctx, e0_, e2, e4 = W[f"{cn}.ctx"], W[f"{cn}.ep0"], W[f"{cn}.ep2"], W[f"{cn}.ep4"]
    y_hat = plan.raw(f"{cn}.ac.y_hat", (1, hy, wy, 128)); y_sym = plan.raw(f"{cn}.ac.y_sym", (1, hy, wy, 128), torch.int32); y_idx = plan.raw(f"{cn}.ac.y_idx", (1, hy, wy, 128), torch.int32)
    need = lib.tdvc_ar_code_workspace_bytes(1, e0_.cout_pad, e2.cout_pad)
    ws = plan.raw(f"{cn}.ac.ws", ((need + 3) // 4,))
    p = L.ArParams(y=y.ptr, y_ld=y.ld, params=params.ptr, params_ld=params.ld, w_ctx=ctx.w.data_ptr(), b_ctx=ctx.b.data_ptr(),
                   w1=e0_.w.data_ptr(), b1=e0_.b.data_ptr(), c1=e0_.cout, c1_pad=e0_.cout_pad, w2=e2.w.data_ptr(), b2=e2.b.data_ptr(), c2=e2.cout, c2_pad=e2.cout_pad,
                   w3=e4.w.data_ptr(), b3=e4.b.data_ptr(), scale_table=tabs.scale_table.data_ptr(), n_scales=SCALES_LEVELS,
                   y_hat=y_hat.data_ptr(), symbols=y_sym.data_ptr(), indexes=y_idx.data_ptr(), N=1, H=hy, W=wy, C=128, cluster=cl)
    st = torch.cuda.current_stream(dev).cuda_stream
    for rep in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); L.check(lib.tdvc_ar_code(C.byref(p), ws.data_ptr(), need, st), "ar"); b.record(); torch.cuda.synchronize()
        print(f"cluster {cl}: ar_code kernel {a.elapsed_time(b):.2f} ms")
PY
