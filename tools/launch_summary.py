"""Aggregates an ncu launch list (--metrics gpu__time_duration.sum --csv) by kernel: count, total, share."""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    try:
        agg.setdefault(re.sub(r"\(.*", "", r[ki])[:72], []).append(float(r[vi].replace(",", "")))
    except Exception:  # noqa: BLE001
        pass
tot = sum(sum(v) for v in agg.values())
print(f"{sys.argv[1]}: {tot / 1e6:.3f} ms in {sum(len(v) for v in agg.values())} launches")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]))[: int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print(f"  {k:72s} n={len(v):5d} sum={sum(v) / 1e6:8.3f} ms {100 * sum(v) / tot:5.1f}%")
