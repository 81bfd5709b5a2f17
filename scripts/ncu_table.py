#!/usr/bin/env python
"""One line per kernel from `ncu -i rep --page raw --csv` output: python scripts/ncu_table.py file.csv [--md]"""
import csv
import sys

COLS = [("gpu__time_duration.sum", "us", 1), ("dram__bytes_read.sum", "rdMB", 1), ("dram__bytes_write.sum", "wrMB", 1),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%", 1),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts%", 1),
        ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1wf%", 1),
        ("lts__t_sector_hit_rate.pct", "l2hit%", 1), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%", 1),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%", 1),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%", 1),
        ("smsp__inst_executed.sum", "Minst", 1e-6), ("lts__t_requests_srcunit_tex_op_red.sum", "Mred", 1e-6),
        ("lts__t_requests_srcunit_tex_op_read.sum", "Mread", 1e-6), ("launch__registers_per_thread", "regs", 1),
        ("launch__occupancy_limit_shared_mem", "occ_smem", 1)]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    md = "--md" in sys.argv
    sep = " | " if md else " "
    head = sep.join(["%-26s" % "kernel"] + ["%8s" % s for _, s, _ in COLS])
    print(("| " + head + " |") if md else head)
    if md:
        print("|" + "---|" * (len(COLS) + 1))
    tot = 0.0
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0].split("::")[-1][:26]
        vals = []
        for c, _, k in COLS:
            try:
                v = float(r[hdr.index(c)].replace(",", "")) * k
                u = units[hdr.index(c)]
                if c.startswith("dram__bytes"):
                    v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
                if c == "gpu__time_duration.sum":
                    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1.0)
            except (ValueError, IndexError):
                v = float("nan")
            vals.append(v)
        tot += vals[0]
        line = sep.join(["%-26s" % name] + ["%8.1f" % v for v in vals])
        print(("| " + line + " |") if md else line)
    print("total us %.1f" % tot)


if __name__ == "__main__":
    main()
