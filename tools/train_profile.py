"""Developer tool: where one training step (BASELINE config 4 shape) spends its GPU time: python tools/train_profile.py"""
import sys

import torch

sys.path.insert(0, ".")
from tdvc_b200 import synth
from tdvc_b200.model import VideoCompressor

dev = torch.device("cuda:0")
AMP = "--amp" in sys.argv   # enabled_amp of the step (cfg/train.yaml ships amp: True)
torch.manual_seed(synth.SEED)
net = VideoCompressor()
sd = net.state_dict()
synth.condition_state_dict(sd)
net.load_state_dict(sd)
net = net.to(dev).train()
params = [p for n, p in net.named_parameters() if not n.endswith(".quantiles")]
opt = torch.optim.Adam(params, lr=1e-4)
pairs = [synth.make_frame_pair(256, 256, seed=500 + i) for i in range(8)]
x = torch.cat([p[0] for p in pairs]).to(dev)
refs = torch.cat([p[1] for p in pairs]).to(dev)


def step():
    out = net(x, refs, AMP)
    loss = 2048 * torch.nn.MSELoss()(out[0], x) + out[1].mean() + out[2].mean()
    opt.zero_grad()
    loss.backward()
    torch.nn.utils.clip_grad_norm_(params, 2)
    opt.step()


for _ in range(2):
    step()
torch.cuda.synchronize()
import time
t = time.time()
step()
torch.cuda.synchronize()
print("wall per step %.1f ms" % ((time.time() - t) * 1e3))
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=70))
