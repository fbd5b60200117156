"""Per-kernel time shares from an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X.csv ...`).
usage: python tools/launch_shares.py profiles/r2_ncu_launches_c4_enhanced.csv [title] > profiles/r2_ncu_launch_shares_c4_enhanced.txt"""
import collections
import csv
import re
import sys

path = sys.argv[1]
title = sys.argv[2] if len(sys.argv) > 2 else "kernel time shares"
rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
hdr, rows = rows[0], rows[1:]
iN, iV, iU = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = collections.defaultdict(float)
cnt = collections.Counter()
for r in rows:
    v = float(r[iV].replace(",", ""))
    v = v / 1e3 if r[iU] in ("ns", "nsecond") else v * 1e3 if r[iU] in ("ms", "msecond") else v   # -> us
    name = re.sub(r"\(.*", "", r[iN]).replace("vr::", "").strip()
    tot[name] += v
    cnt[name] += 1
total = sum(tot.values())
print(f"# {title}")
print(f"# from {path} (ncu --metrics gpu__time_duration.sum --clock-control none: serialised, cold cache -- shares, not bench values)")
for name, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{name:70s} n={cnt[name]:4d} {v / 1e3:8.2f} ms  {100 * v / total:5.1f} %  avg {v / cnt[name]:7.1f} us")
print(f"{'total':70s} n={sum(cnt.values()):4d} {total / 1e3:8.2f} ms")
