"""Writes profiles/r2_ncu_dominant_kernel.json (read by bench.py's roofline.traffic) from an ncu_summary.py CSV of the
DRDB dilated-conv launches: dram__bytes_read.sum + dram__bytes_write.sum, averaged per launch of the dominant family
(conv3x3_tc_kernel<32, 2, .> pulls and drdb_push_tc_kernel).
usage: python tools/dominant_traffic.py profiles/r2_ncu_drdb_summary.csv"""
import csv
import json
import os
import sys

UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def nbytes(cell):
    v, u = cell.split()
    return float(v.replace(",", "")) * UNIT[u]


def main(path):
    rows = [r for r in csv.DictReader(open(path)) if "conv3x3_tc_kernel<32, 2" in r["kernel"] or "drdb_push_tc_kernel" in r["kernel"]]
    per = [nbytes(r["dram_rd"]) + nbytes(r["dram_wr"]) for r in rows]
    out = {"source": os.path.basename(path), "kernel": rows[0]["kernel"], "launches": len(per),
           "traffic_bytes_per_launch": sum(per) / len(per), "per_launch": per,
           "kernels": [r["kernel"] for r in rows],
           "note": "ncu --set full --clock-control none at HEAD of round 2 (bench.py --profile --no-graph, B=8 480x640): two pull launches "
                   "(layers 4, 5: algorithmic (Cin_g + 32 partial + 32 out) * 2 B * 2 457 600 px = 786 / 944 MB) and the two x0 push "
                   "launches (algorithmic 64 in + 96 | 64 out channels: 787 / 629 MB)"}
    dst = os.path.join(os.path.dirname(os.path.abspath(path)), "r2_ncu_dominant_kernel.json")
    json.dump(out, open(dst, "w"), indent=1)
    print(dst, out["traffic_bytes_per_launch"])


if __name__ == "__main__":
    main(sys.argv[1])
