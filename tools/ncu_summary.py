#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) into the few numbers DESIGN.md argues from.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [> profiles/rNN_<kernel>.txt]
"""
import csv
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_warps",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed.avg.per_cycle_elapsed", "sm__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    head, units = rows[0], rows[1]
    for row in rows[2:]:
        d = dict(zip(head, row))
        u = dict(zip(head, units))
        print(f"kernel: {d.get('Kernel Name')}  grid {d.get('Grid Size')} block {d.get('Block Size')}")
        for k in WANT:
            if k in d:
                print(f"  {k:70s} {d[k]} {u[k]}")
        stalls = [(float(d[k]), k) for k in head
                  if k.startswith("smsp__average_warp") and k.endswith("_per_issue_active.ratio")
                  or k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio")]
        stalls = [(v, k) for v, k in stalls if v > 0.05]
        for v, k in sorted(stalls, reverse=True)[:10]:
            print(f"  stall {k:64s} {v:.2f}")


if __name__ == "__main__":
    main()
