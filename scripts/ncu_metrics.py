#!/usr/bin/env python
"""Print a compact set of metrics for every kernel in an .ncu-rep (reads it with `ncu -i ... --page raw --csv`).

    python scripts/ncu_metrics.py gpurun_out/x.ncu-rep [kernel-substring]"""
import csv
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
    "lts__t_requests_srcunit_tex_op_red.sum", "lts__t_requests_srcunit_tex_op_read.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_red.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "sm__cycles_elapsed.max",
]
STALLS = ["long_scoreboard", "short_scoreboard", "wait", "barrier", "mio_throttle", "lg_throttle", "no_instruction",
          "math_pipe_throttle", "branch_resolving", "sleeping", "membar", "dispatch_stall", "drain", "imc_miss",
          "tex_throttle", "not_selected", "selected"]


def main():
    raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    flt = sys.argv[2] if len(sys.argv) > 2 else ""
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        if flt not in name:
            continue
        print("==", name.split("(")[0][-70:])
        for w in WANT:
            if w in hdr:
                print(f"   {w:75s} {r[hdr.index(w)]:>16s} {units[hdr.index(w)]}")
        st = []
        for k in STALLS:
            m = f"smsp__average_warps_issue_stalled_{k}_per_issue_active.ratio"
            if m in hdr:
                st.append((float(r[hdr.index(m)].replace(",", "")), k))
        print("   stalls/issue:", ", ".join(f"{k} {v:.2f}" for v, k in sorted(st, reverse=True)[:8]))


if __name__ == "__main__":
    main()
