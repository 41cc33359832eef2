"""Developer tool (build container): hottest SASS instructions of an .ncu-rep by warp-stall samples, with the stall
reasons, to see what a kernel waits on.  python tools/ncu_hot.py gpurun_out/x.ncu-rep [top]"""
import csv
import subprocess
import sys

path, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
body = [r for r in rows[2:] if len(r) == len(hdr)]
tot = sum(int(r[ix["# Samples"]] or 0) for r in body)
agg = {h: sum(int(r[ix[h]] or 0) for r in body) for h in stall_cols}
print(f"total samples {tot}; by reason: " + ", ".join(f"{k[6:]} {v * 100 // max(tot, 1)}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for n, r in enumerate(body):
    r.append(n)
for r in sorted(body, key=lambda r: -int(r[ix["# Samples"]] or 0))[:top]:
    s = int(r[ix["# Samples"]] or 0)
    reasons = sorted(((int(r[ix[h]] or 0), h[6:]) for h in stall_cols), reverse=True)[:3]
    print(f"{r[-1]:5d} {s * 100.0 / tot:5.1f}%  exec {r[ix['Instructions Executed']]:>9s}  {r[ix['Source']][:90]:90s} " +
          " ".join(f"{n}:{c}" for c, n in reasons if c))

# per execution-count class (= per warp role: instructions of one loop share their execution count)
cls = {}
for r in body:
    e = int(r[ix["Instructions Executed"]] or 0)
    c = cls.setdefault(e, {"n": 0, "s": 0, "r": {}})
    c["n"] += 1
    c["s"] += int(r[ix["# Samples"]] or 0)
    for h in stall_cols:
        c["r"][h] = c["r"].get(h, 0) + int(r[ix[h]] or 0)
print("\nby execution count (role):")
for e, c in sorted(cls.items(), key=lambda kv: -kv[1]["s"])[:12]:
    rs = sorted(c["r"].items(), key=lambda kv: -kv[1])[:5]
    print(f"  exec {e:>9d}: {c['n']:4d} instr, {c['s'] * 100.0 / tot:5.1f}% of samples: " + ", ".join(f"{k[6:]} {v}" for k, v in rs if v))
