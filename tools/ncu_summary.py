#!/usr/bin/env python
"""Prints the key metrics of every launch in an `ncu --set full` report, with DRAM / L2 / L1 GB/s derived from the sector counters.

Usage: python tools/ncu_summary.py gpurun_out/prof_x.ncu-rep [peak_hbm_gbs] > profiles/rNN/ncu_x.txt
       python tools/ncu_summary.py gpurun_out/prof_x.ncu-rep --evidence workload:accel [launch index] [source note]
(run where `ncu` is on PATH; the .ncu-rep files themselves stay in gpurun_out/, the summaries are what is committed).
--evidence writes the measured fractions of ONE captured launch (default: the second, the bounce-1 launch) into profiles/ncu_evidence.json,
which bench.py attaches to its `roofline` object: what limits the kernel, as measured, next to the contract's algorithmic-bytes figure."""
import csv
import json
import os
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sectors.sum", "l1tex__t_sector_hit_rate.pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__maximum_warps_per_active_cycle_pct", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
]
UNIT_SCALE = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3, "second": 1.0, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def evidence(rep, key, index, note):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    peak = float(json.load(open(os.path.join(root, "MEASURED_PEAKS.json")))["hbm_gbs"])
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    head, units = rows[0], rows[1]
    col = {k: i for i, k in enumerate(head)}
    row = rows[2 + index]

    def val(key):
        return float(row[col[key]].replace(",", "")) * UNIT_SCALE.get(units[col[key]], 1.0)

    t = val("gpu__time_duration.sum")
    dram = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
    local = val("l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum") + val("l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum")
    e = {
        "source": note or os.path.basename(rep), "kernel": row[col["Kernel Name"]][:40], "launch": f"capture #{index} (bounce-{index} launch of the frame)",
        "duration_us_under_ncu": round(t * 1e6, 1), "registers": int(val("launch__registers_per_thread")),
        "dram_bytes_per_launch": int(dram), "dram_frac_of_measured_hbm_peak": round(dram / t / 1e9 / peak, 4),
        "l2_throughput_frac": round(val("lts__throughput.avg.pct_of_peak_sustained_elapsed") / 100, 4), "l2_hit_rate": round(val("lts__t_sector_hit_rate.pct") / 100, 4),
        "l1tex_throughput_frac": round(val("l1tex__throughput.avg.pct_of_peak_sustained_elapsed") / 100, 4), "l1_hit_rate": round(val("l1tex__t_sector_hit_rate.pct") / 100, 4),
        "l1_local_memory_sectors": int(local),
        "issue_active_frac": round(val("smsp__issue_active.avg.pct_of_peak_sustained_active") / 100, 4),
        "lanes_per_instruction": round(val("smsp__thread_inst_executed_per_inst_executed.ratio"), 2), "warp_instructions": int(val("smsp__inst_executed.sum")),
        "pipe_alu_frac": round(val("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active") / 100, 4),
        "pipe_fma_frac": round(val("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active") / 100, 4),
        "achieved_occupancy_frac": round(val("sm__warps_active.avg.pct_of_peak_sustained_active") / 100, 4),
        "stall_long_scoreboard_per_issue": round(val("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"), 2),
    }
    path = os.path.join(root, "profiles", "ncu_evidence.json")
    allv = json.load(open(path)) if os.path.exists(path) else {}
    allv[key] = e
    json.dump(allv, open(path, "w"), indent=1)
    print(json.dumps(e, indent=1))


def main():
    if len(sys.argv) > 3 and sys.argv[2] == "--evidence":
        evidence(sys.argv[1], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 1, sys.argv[5] if len(sys.argv) > 5 else "")
        return
    rep = sys.argv[1]
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    peak = float(sys.argv[2]) if len(sys.argv) > 2 else None
    if peak is None:
        try:
            peak = float(json.load(open(os.path.join(root, "MEASURED_PEAKS.json")))["hbm_gbs"])
        except Exception:
            peak = None
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    head, units = rows[0], rows[1]
    col = {k: i for i, k in enumerate(head)}

    def val(row, key):      # value in base units (seconds, bytes) where the report gives a unit
        v = float(row[col[key]].replace(",", ""))
        return v * UNIT_SCALE.get(units[col[key]], 1.0)

    for row in rows[2:]:
        print(row[col["Kernel Name"]][:100])
        for k in METRICS:
            if k in col and row[col[k]] != "":
                print(f"   {k:<84} {row[col[k]]} {units[col[k]]}")
        t = val(row, "gpu__time_duration.sum")
        dram = val(row, "dram__bytes_read.sum") + val(row, "dram__bytes_write.sum")
        line = f"   derived: DRAM {dram / t / 1e9:.0f} GB/s"
        if peak:
            line += f" ({100 * dram / t / 1e9 / peak:.1f} % of the {peak:.0f} GB/s measured peak)"
        line += f", L2 {val(row, 'lts__t_sectors.sum') * 32 / t / 1e9:.0f} GB/s"
        l1 = [k for k in col if k.endswith("l1tex__t_sectors.sum")]
        if l1:
            line += f", L1 {val(row, l1[0]) * 32 / t / 1e9:.0f} GB/s"
        print(line)


if __name__ == "__main__":
    main()
