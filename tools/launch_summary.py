"""Per-kernel shares of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
Usage: python tools/launch_summary.py gpurun_out/launches.csv "<command that was profiled>" > profiles/x_summary.txt"""
import csv
import re
import sys
from collections import defaultdict

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')) if len(r) > 14]
hdr, rows = rows[0], rows[1:]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows:
    name = re.sub(r"\(.*", "", r[ki])[:96]
    tot[name] += float(r[vi].replace(",", "")) / 1e3
    cnt[name] += 1
total = sum(tot.values())
print(f"# ncu --metrics gpu__time_duration.sum --clock-control none over `{sys.argv[2] if len(sys.argv) > 2 else '?'}` "
      f"(first {len(rows)} launches; cold-cache, serialised: compare SHARES)")
print(f"# {len(rows)} launches, {total / 1e3:.1f} ms")
for name, us in sorted(tot.items(), key=lambda kv: -kv[1])[:40]:
    print(f"{us:12.1f} us {100 * us / total:5.1f}% x{cnt[name]:4d} {name}")
