"""Sum an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name (microseconds)."""
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
ki, vi = rows[h].index("Kernel Name"), rows[h].index("Metric Value")
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
agg, cnt = {}, {}
for r in rows[h + 1 + skip:]:
    if len(r) > vi:
        name = re.sub(r"\(.*", "", r[ki]).split("::")[-1]
        agg[name] = agg.get(name, 0.0) + float(r[vi]) / 1e3
        cnt[name] = cnt.get(name, 0) + 1
for k in sorted(agg, key=agg.get, reverse=True):
    print(f"{agg[k]:10.1f} us  x{cnt[k]:4d}  {k}")
print(f"{sum(agg.values()):10.1f} us total")
