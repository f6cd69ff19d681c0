"""Summarise one kernel of an ncu report as JSON (what profiles/*.json hold and bench.py's `ncu_capture` reads).

    python tools/ncu_summary.py <report.ncu-rep> <units> "<command that produced it>" [unit-name [k]] > profiles/<name>.json

Uses `ncu -i <rep> --page raw --csv`; the k-th kernel of the report (default: the first) is summarised.  `units` = the
work units (windows, sequences, pairs ...) that launch processed, for the per-unit figures."""
import csv
import json
import subprocess
import sys

KEEP = [
    "Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
    "sm__inst_executed.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.per_cycle_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "launch__block_size", "launch__grid_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__cycles_elapsed.avg.per_second",
]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
TIME = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}


def main():
    rep, units, command = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    unit_name = sys.argv[4] if len(sys.argv) > 4 else "window"
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, unit_row, vals = rows[0], rows[1], rows[2 + (int(sys.argv[5]) if len(sys.argv) > 5 else 0)]
    metrics, stalls = {}, {}
    for h, u, v in zip(hdr, unit_row, vals):
        if h in KEEP:
            metrics[h] = {"unit": u, "value": v}
        if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio") and "not_issued" not in h:
            stalls[h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = round(float(v.replace(",", "")), 3)

    def num(name):
        m = metrics[name]
        return float(m["value"].replace(",", "")) * SCALE.get(m["unit"], 1.0)

    rd, wr, inst = num("dram__bytes_read.sum"), num("dram__bytes_write.sum"), num("smsp__inst_executed.sum")
    t = metrics["gpu__time_duration.sum"]
    ms = float(t["value"].replace(",", "")) * TIME.get(t["unit"], 1.0)
    top = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:6])
    print(json.dumps({
        "command": command, "kernel": metrics["Kernel Name"]["value"], "windows": units, "unit": unit_name,
        "kernel_ms": ms, "dram_bytes_read": rd, "dram_bytes_write": wr, "traffic_bytes": rd + wr,
        "dram_GBps": (rd + wr) / ms / 1e6,
        "warp_instructions": inst, "warp_instructions_per_window": inst / units,
        "stall_cycles_per_issue_top": top, "metrics": metrics}, indent=1))


if __name__ == "__main__":
    main()
