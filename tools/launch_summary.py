"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name (microseconds)."""
import collections
import csv
import re
import sys


def main(path, top=25):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    h = rows[0]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        n = re.sub(r"\(.*", "", r[ki])
        v = float(r[vi].replace(",", ""))
        v = v / 1000 if r[ui] == "ns" else v * 1000 if r[ui] == "ms" else v
        agg[n][0] += 1
        agg[n][1] += v
    print("total us %.1f launches %d" % (sum(v[1] for v in agg.values()), sum(v[0] for v in agg.values())))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{v[1]:9.1f} us {v[0]:4d}  {k[:110]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
