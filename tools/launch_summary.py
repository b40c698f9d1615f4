"""Per-kernel shares of an `ncu --metrics gpu__time_duration.sum --csv` launch list: python tools/launch_summary.py <csv> [title]"""
import csv, sys
from collections import defaultdict
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows[1:]:
    if len(r) <= vi or r[hdr.index("Metric Name")] != "gpu__time_duration.sum":
        continue
    v = float(r[vi].replace(",", ""))
    u = r[ui]
    ms = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v
    name = r[ki].split("(")[0]
    tot[name] += ms
    cnt[name] += 1
s = sum(tot.values()) or 1
print("# " + (sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]))
print("# (per-launch times under ncu are cold-cache and serialised: compare SHARES)")
for k in sorted(tot, key=lambda k: -tot[k]):
    print(f"{k[:60]:60s} launches={cnt[k]:4d} total_ms={tot[k]:10.3f} avg_ms={tot[k]/cnt[k]:9.3f} share={100*tot[k]/s:6.2f}%")
