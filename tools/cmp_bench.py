"""Developer tool: compare the per-kernel tables of two bench.py JSON lines (A/B of a kernel change)."""
import json
import sys


def load(p):
    return json.loads(open(p).read().strip().splitlines()[-1])


a, b = load(sys.argv[1]), load(sys.argv[2])
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.03
print(f"value {a['value']:.3f} -> {b['value']:.3f}   ms/step {a['ms_per_step']:.2f} -> {b['ms_per_step']:.2f}   "
      f"e2e {a['e2e']['value']:.3f} -> {b['e2e']['value']:.3f}   clocks {a['clocks']['sm_mhz']} -> {b['clocks']['sm_mhz']}")
ka, kb = a["kernels"], b["kernels"]
ta = sum(v["ms"] for v in ka.values())
tb = sum(v["ms"] for v in kb.values())
print(f"instrumented frame: {ta:.2f} -> {tb:.2f} ms")
for k in sorted(set(ka) | set(kb), key=lambda k: -max(ka.get(k, {}).get("ms", 0), kb.get(k, {}).get("ms", 0))):
    x, y = ka.get(k, {}).get("ms", 0.0), kb.get(k, {}).get("ms", 0.0)
    if abs(x - y) >= thr:
        print(f"  {k:40s} {x:8.3f} -> {y:8.3f}  ({y - x:+.3f})")
