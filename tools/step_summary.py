"""Per-kernel summary of the LAST step in an `ncu --metrics gpu__time_duration.sum --csv` launch list of
tools/train_bench.py (which runs nsteps identical steps: eager, warm-up, timed): time and launches per kernel name."""
import collections
import csv
import re
import sys


def main(path, nsteps=3, top=28):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    h = rows[0]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    n = (len(rows) - 1) // nsteps
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1 + (nsteps - 1) * n:]:
        k = re.sub(r"\(.*", "", r[ki])
        v = float(r[vi].replace(",", ""))
        v = v / 1000 if r[ui] == "ns" else v * 1000 if r[ui] == "ms" else v
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    print("last step: %.1f us in %d launches (cold-cache, serialised: shares, not absolutes)" % (tot, sum(v[0] for v in agg.values())))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{v[1]:9.1f} us {100 * v[1] / tot:5.1f}% {v[0]:4d}  {k[:100]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 3, int(sys.argv[3]) if len(sys.argv) > 3 else 28)
