"""Writes profiles/r1_ncu_dominant_kernel.json (read by bench.py's roofline.traffic) from an ncu_summary.py CSV of the
DRDB dilated-conv launches: dram__bytes_read.sum + dram__bytes_write.sum, averaged per launch.
usage: python tools/dominant_traffic.py profiles/r1_ncu_conv3x3_tc_v10_summary.csv"""
import csv
import json
import os
import sys

UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def nbytes(cell):
    v, u = cell.split()
    return float(v.replace(",", "")) * UNIT[u]


def main(path):
    rows = list(csv.DictReader(open(path)))
    per = [nbytes(r["dram_rd"]) + nbytes(r["dram_wr"]) for r in rows]
    out = {"source": os.path.basename(path), "kernel": rows[0]["kernel"], "launches": len(per),
           "traffic_bytes_per_launch": sum(per) / len(per), "per_launch": per,
           "note": "ncu --set full --clock-control none, one DRDB (layers 2-5, g-slab pull launches, B=8 480x640); "
                   "algorithmic bytes per launch = (Cin_g + 32 partial + 32 out) * 2 B * 2 457 600 px = 472 / 629 / 786 / 944 MB"}
    dst = os.path.join(os.path.dirname(os.path.abspath(path)), "r1_ncu_dominant_kernel.json")
    json.dump(out, open(dst, "w"), indent=1)
    print(dst, out["traffic_bytes_per_launch"])


if __name__ == "__main__":
    main(sys.argv[1])
