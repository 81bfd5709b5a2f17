#!/usr/bin/env python
"""Turn ncu output brought back in gpurun_out/ into the small text summaries committed under profiles/.

    python scripts/summarize_ncu.py launches gpurun_out/launches_r1b.csv profiles/r1b_launches.md
    python scripts/summarize_ncu.py full     gpurun_out/prof_density_r1b.ncu-rep profiles/r1b_density_full.md
"""
import collections
import csv
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "l1tex__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.sum", "smsp__inst_executed.sum", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard_per_warp_active.pct",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
    "l1tex__t_set_accesses_pipe_lsu_mem_global_op_red.sum", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum",
    "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum",
    "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_red.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__cycles_elapsed.max", "smsp__issue_active.avg.pct_of_peak_sustained_active",
]


def launches(src, dst):
    with open(src) as fh:
        lines = [l for l in fh if not l.startswith("==")]
    agg = collections.OrderedDict()
    order = []
    for row in csv.DictReader(lines):
        name = row["Kernel Name"].split("(")[0].replace("ucsa::<unnamed>::", "")
        val = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        us = val / 1e3 if unit == "ns" else val * 1e3 if unit == "ms" else val
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += us
        order.append((name, us))
    total = sum(v[1] for v in agg.values())
    with open(dst, "w") as out:
        out.write(f"# ncu launch list summary ({src})\n\n")
        out.write("`ncu --metrics gpu__time_duration.sum --clock-control none` over a short bench.py run: per-launch "
                  "times are cold-cache and serialised -- compare SHARES, not absolutes.\n\n")
        out.write(f"total device time in the captured window: {total / 1e3:.2f} ms over {len(order)} launches\n\n")
        out.write("| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
        for name, (cnt, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
            out.write(f"| `{name[:80]}` | {cnt} | {us:.1f} | {100 * us / total:.1f}% |\n")
    print(open(dst).read())


def full(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as out:
        out.write(f"# ncu --set full summary ({src})\n\n")
        for row in rows[2:]:
            name = row[hdr.index("Kernel Name")].split("(")[0]
            out.write(f"## {name}\n\n| metric | value | unit |\n|---|---:|---|\n")
            for m in METRICS:
                if m in hdr:
                    i = hdr.index(m)
                    out.write(f"| {m} | {row[i]} | {units[i]} |\n")
            out.write("\n")
    print(open(dst).read())


def traffic(src, dst):
    """Per launch of every profiled kernel: DRAM bytes (read + write) -> bench.py's roofline.traffic, and the L2 request
    counters (reductions / reads issued by the SMs) -> bench.py's l2_request_roofline."""
    import json

    if src.endswith(".csv"):  # `ncu -i rep --page raw --csv` already run on the GPU box (reports can exceed the 64 MiB limit)
        raw = open(src).read()
    else:
        raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    req_metrics = {
        "lts_requests_red": "lts__t_requests_srcunit_tex_op_red.sum",
        "lts_requests_read": "lts__t_requests_srcunit_tex_op_read.sum",
        "l1_requests_red": "l1tex__t_requests_pipe_lsu_mem_global_op_red.sum",
        "l1_requests_ld": "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "lts_sectors_red": "lts__t_sectors_op_red.sum",
        "lts_hit_rate_pct": "lts__t_sector_hit_rate.pct",
    }
    acc, reqs = {}, {}
    for row in rows[2:]:
        name = row[hdr.index("Kernel Name")].split("(")[0].split("::")[-1].split("<")[0]
        total = 0.0
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = hdr.index(m)
            total += float(row[i].replace(",", "")) * scale[units[i]]
        acc.setdefault(name, []).append(total)
        for key, m in req_metrics.items():
            if m in hdr:
                try:
                    reqs.setdefault(name, {}).setdefault(key, []).append(float(row[hdr.index(m)].replace(",", "")))
                except ValueError:
                    pass
    out = {"source": src, "note": "dram__bytes_read.sum + dram__bytes_write.sum per launch (ncu --set full), mean over "
           "the profiled launches; l2_requests_per_launch: ncu request counters, same launches",
           "bytes_per_launch": {k: sum(v) / len(v) for k, v in acc.items()},
           "l2_requests_per_launch": {k: {m: sum(x) / len(x) for m, x in v.items()} for k, v in reqs.items()}}
    with open(dst, "w") as fh:
        json.dump(out, fh, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](sys.argv[2], sys.argv[3])
