"""Condenses `ncu --set full` reports into the handful of numbers DESIGN.md / profiles/ quote.
usage: python tools/ncu_summary.py gpurun_out/ncu_x.ncu-rep [...]  (prints CSV to stdout)"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor_inst"),
    ("sm__pipe_tensor_subpipe_mma_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pct"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pct2"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"),
    ("smsp__inst_executed.sum", "warp_inst"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "smem_dyn"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conflicts"),
]


def rows_of(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(out)))
    hdr, units = rd[0], rd[1]
    for r in rd[2:]:
        yield {h: (v, u) for h, v, u in zip(hdr, r, units)}


def main():
    w = csv.writer(sys.stdout)
    w.writerow(["report", "kernel"] + [k[1] for k in KEYS] + ["stall_top3"])
    for p in sys.argv[1:]:
        for r in rows_of(p):
            name = r["Kernel Name"][0][:60]
            vals = []
            for full, _ in KEYS:
                v = r.get(full)
                vals.append(f"{v[0]} {v[1]}".strip() if v else "")
            stalls = sorted(((float(v[0].replace(",", "")), k.split("issue_stalled_")[-1].split("_per_warp")[0]) for k, v in r.items()
                             if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("_per_issue_active.ratio") and v[0] not in ("", "n/a")), reverse=True)[:3]
            w.writerow([p.split("/")[-1], name] + vals + [" ".join(f"{n}={x:.2f}" for x, n in stalls)])


if __name__ == "__main__":
    main()
