#!/usr/bin/env python
"""Summarise `.ncu-rep` files (ncu --set full) into the small CSV / JSON files kept under profiles/.

  python tools/ncu_summary.py gpurun_out/r2d_pip_stream.ncu-rep profiles/r2_pip_summary.csv [--traffic profiles/r2_pip_traffic.json --note "..."]

Reads the report with `ncu -i <rep> --page raw --csv` (works without a GPU) and keeps the counters DESIGN.md quotes:
duration, DRAM bytes, L2 sectors / hit rate, L1TEX and LTS throughput, issue-active, lanes per instruction, registers, stalls."""
import argparse
import csv
import io
import json
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]
BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("out_csv")
    ap.add_argument("--traffic")
    ap.add_argument("--note", default="")
    ap.add_argument("--row", type=int, default=0, help="which captured launch of the report")
    a = ap.parse_args()
    raw = subprocess.check_output(["ncu", "-i", a.rep, "--page", "raw", "--csv"], text=True, stderr=subprocess.DEVNULL)
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2 + a.row]
    kernel = vals[hdr.index("Kernel Name")]
    out = [("kernel", "", kernel)]
    got = {}
    for k in KEEP:
        if k in hdr:
            i = hdr.index(k)
            out.append((k, units[i], vals[i]))
            got[k] = (units[i], vals[i])
    with open(a.out_csv, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit", "value"])
        w.writerows(out)
    print(f"{a.out_csv}: {kernel[:80]} {got.get('gpu__time_duration.sum')}")
    if a.traffic:
        def by(name):
            u, v = got[name]
            return float(v.replace(",", "")) * BYTES[u]

        rd, wr = by("dram__bytes_read.sum"), by("dram__bytes_write.sum")
        json.dump({"kernel": kernel, "dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr,
                   "source": f"{a.out_csv} (ncu --set full --clock-control none; {a.note})"}, open(a.traffic, "w"), indent=1)


if __name__ == "__main__":
    sys.exit(main())
