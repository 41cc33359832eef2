"""Developer tool: per-shape time of the weight-gradient calls of a training step, from the log written by
`TDVC_B200_WGRAD_LOG=1 python tools/train_profile.py --amp 2> log` (events around every call): python tools/wgrad_census.py log"""
import collections
import re
import sys

ms, cnt = collections.Counter(), collections.Counter()
for line in open(sys.argv[1]):
    m = re.match(r"(wgrad .*) ms ([0-9.]+)", line)
    if m:
        ms[m.group(1)] += float(m.group(2))
        cnt[m.group(1)] += 1
steps = 4   # train_profile.py runs four steps
tot = sum(ms.values()) / steps
print(f"total {tot:.2f} ms per step over {sum(cnt.values()) // steps} calls")
for k, v in ms.most_common(40):
    print(f"{v / steps:7.3f} ms  x{cnt[k] // steps:3d}  {k}")
