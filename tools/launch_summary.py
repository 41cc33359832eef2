"""Developer tool (build container): per-kernel totals of an ncu launch list (`--metrics gpu__time_duration.sum --csv`).
    python tools/launch_summary.py gpurun_out/r01_launches.csv > profiles/r01_launches_summary.txt"""
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
agg = {}
for r in rows:
    name = r[4].replace("void ", "").split("(")[0]
    v = float(r[14].replace(",", ""))
    v = v / 1000.0 if r[13] in ("ns", "nsecond") else (v * 1000.0 if r[13] in ("ms", "msecond") else v)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print("# ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none : python tools/profile_frame.py")
print("# the launches of ONE steady-state 1920x1024 P-frame (frame 7 of a GOP, enabled_amp=True); per-launch times are cold-cache and serialised: compare SHARES")
print(f"# total {tot / 1000.0:.2f} ms over {len(rows)} launches")
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:88]:88s} {n:4d} launches {us:10.1f} us {100.0 * us / tot:5.1f}%")
