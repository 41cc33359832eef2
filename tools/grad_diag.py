"""Developer tool: per-parameter gradient agreement of the autograd training path with the oracle (CPU): python tools/grad_diag.py [N H W]."""
import copy
import sys

import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from oracle.stats import build_oracle
from tdvc_b200 import synth
from tdvc_b200.model import VideoCompressor
import os
from test_training_step import _noise

WHICH = os.environ.get("LOSS", "all")


def _rd_loss(out, x):
    mse = torch.nn.MSELoss()(out[0], x)
    parts = {"mse": 2048 * mse, "bpp_res": out[1].mean(), "bpp_mv": out[2].mean()}
    return (sum(parts.values()) if WHICH == "all" else parts[WHICH]), mse

N, H, W = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (2, 64, 64)
dev = torch.device("cuda:0")
orc = copy.deepcopy(build_oracle()).train()
net = VideoCompressor()
net.load_state_dict(orc.state_dict())
net = net.to(dev).train()
xs, rs = zip(*[synth.make_frame_pair(H, W, seed=91 + i) for i in range(N)])
x, refs = torch.cat(xs, 0), torch.cat(rs, 0)
torch.manual_seed(77)
want = orc(x, refs, False)
lw, _ = _rd_loss(want, x)
lw.backward()
torch.manual_seed(77)
noise = {k: v.to(dev) for k, v in _noise(N, H, W).items()}
got = net._forward_training_autograd(x.to(dev), refs.to(dev), noise=noise)
lg, _ = _rd_loss(got, x.to(dev))
lg.backward()
print("loss", lw.item(), lg.item(), "recon err", (want[0] - got[0].detach().cpu()).abs().max().item())
ref = dict(orc.named_parameters())
rows = []
for name, p in net.named_parameters():
    r = ref[name].grad
    if r is None or p.grad is None:
        continue
    g = p.grad.cpu()
    scale = r.abs().max().item()
    err = (g - r).abs().max().item()
    cos = torch.nn.functional.cosine_similarity(g.reshape(1, -1), r.reshape(1, -1)).item()
    rel_l2 = ((g - r).norm() / (r.norm() + 1e-30)).item()
    rows.append((err / max(scale, 1e-30), rel_l2, cos, name, scale))
rows.sort(reverse=True)
print("params", len(rows), "median rel max", sorted(r[0] for r in rows)[len(rows) // 2], "median rel l2", sorted(r[1] for r in rows)[len(rows) // 2])
print('LOSS =', WHICH)
for r in rows[:int(os.environ.get('TOP', '25'))]:
    print(f"relmax {r[0]:.3e} rel_l2 {r[1]:.3e} cos {r[2]:.6f} scale {r[4]:.3e} {r[3]}")
