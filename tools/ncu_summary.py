"""Summarise one kernel of an ncu report as JSON (the file bench.py reads `traffic` and the
issue-slot roofline from).

    python tools/ncu_summary.py <report.ncu-rep> <windows> "<command that produced it>" > profiles/<name>.json

Uses `ncu -i <rep> --page raw --csv`; the first kernel in the report is summarised."""
import csv
import json
import subprocess
import sys

KEEP = [
    "Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "sm__inst_executed.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.per_cycle_active",
    "launch__block_size", "launch__grid_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    rep, windows, command = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    metrics = {}
    for h, u, v in zip(hdr, units, vals):
        if h in KEEP or h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
            metrics[h] = {"unit": u, "value": v}

    def num(name):
        m = metrics[name]
        return float(m["value"].replace(",", "")) * SCALE.get(m["unit"], 1.0)

    rd, wr, inst = num("dram__bytes_read.sum"), num("dram__bytes_write.sum"), num("smsp__inst_executed.sum")
    print(json.dumps({
        "command": command, "kernel": metrics["Kernel Name"]["value"], "windows": windows,
        "dram_bytes_read": rd, "dram_bytes_write": wr, "traffic_bytes": rd + wr,
        "warp_instructions": inst, "warp_instructions_per_window": inst / windows, "metrics": metrics}, indent=1))


if __name__ == "__main__":
    main()
