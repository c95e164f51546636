#!/usr/bin/env python
"""Selected counters of an `ncu --set full` report as JSON: ncu_raw_summary.py report.ncu-rep "how it was captured" > out.json"""
import csv
import io
import json
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
           "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
           "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
           "sm__cycles_active.avg", "sm__cycles_elapsed.avg"]


def main():
    rep, how = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    out = {"source": how, "units": {}, "kernels": []}
    cols = [(m, hdr.index(m)) for m in ["Kernel Name"] + METRICS if m in hdr]
    for m, i in cols:
        out["units"][m] = units[i]
    for r in rows[2:]:
        out["kernels"].append({m: r[i] for m, i in cols})
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
