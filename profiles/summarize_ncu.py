#!/usr/bin/env python3
"""Summarise ncu output into the text files kept under profiles/.

    python profiles/summarize_ncu.py launches gpurun_out/launches.csv   > profiles/rNN_launches.txt
    python profiles/summarize_ncu.py full     gpurun_out/prof.ncu-rep   > profiles/rNN_sw_full.txt
"""
import collections
import csv
import io
import subprocess
import sys

FULL_METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "l1tex__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio", "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
    "smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio", "smsp__average_warp_latency_issue_stalled_wait.ratio",
    "smsp__average_warp_latency_issue_stalled_not_selected.ratio", "smsp__average_warp_latency_issue_stalled_barrier.ratio",
    "smsp__average_warp_latency_issue_stalled_branch_resolving.ratio", "smsp__average_warp_latency_issue_stalled_dispatch_stall.ratio",
    "smsp__average_warp_latency_issue_stalled_lg_throttle.ratio", "smsp__average_warp_latency_issue_stalled_mio_throttle.ratio",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
]


def csv_rows(text):
    lines = [l for l in text.splitlines() if l.startswith('"')]
    return list(csv.reader(io.StringIO("\n".join(lines))))


def launches(path):
    rows = csv_rows(open(path).read())
    hdr = rows[0]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        name = r[kn].split("(")[0].replace("catt::<unnamed>::", "").replace("void ", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[mv].replace(",", ""))
    tot = sum(a[1] for a in agg.values())
    print(f"# {path}: {len(rows) - 1} launches, {tot / 1e6:.3f} ms of device time (ncu: cold-cache, serialised - compare shares)")
    print(f"{'launches':>8} {'total ms':>10} {'share':>7}  kernel")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{a[0]:8d} {a[1] / 1e6:10.3f} {100 * a[1] / tot:6.1f}%  {k}")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = csv_rows(out)
    hdr, units = rows[0], rows[1]
    kn = hdr.index("Kernel Name")
    print(f"# {path}: ncu --set full, selected raw metrics per profiled launch")
    for r in rows[2:]:
        print(f"\n== {r[kn]}")
        for m in FULL_METRICS:
            if m in hdr:
                i = hdr.index(m)
                print(f"  {m:80s} {r[i]:>18s} {units[i]}")


def traffic(path):
    """profiles/traffic.json: dram bytes (read + write) per launch of the dominant kernel, largest launch of each name."""
    import json
    import os
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = csv_rows(out)
    hdr, units = rows[0], rows[1]
    kn, rd, wr, du = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    res = {}
    for r in rows[2:]:
        name = r[kn].split("<")[0].split("::")[-1].split("(")[0]
        b = float(r[rd].replace(",", "")) * scale[units[rd]] + float(r[wr].replace(",", "")) * scale[units[wr]]
        if name not in res or b > res[name]:
            res[name] = b
    # the largest launch of the DP and seeding kernels: pipe utilisation, issue slots, warp instructions (per read for seeding:
    # argv[3] = reads in that launch)
    hdrs = {h: i for i, h in enumerate(hdr)}
    def num(r, m):
        return float(r[hdrs[m]].replace(",", ""))
    big = {}
    for r in rows[2:]:
        nm = r[kn].split("<")[0].split("::")[-1].split("(")[0]
        if nm not in big or num(r, "gpu__time_duration.sum") > num(big[nm], "gpu__time_duration.sum"):
            big[nm] = r
    if "sw_bulk2_kernel" in big:
        res["sw_bulk2_kernel_pipe_alu_pct"] = round(num(big["sw_bulk2_kernel"], "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"), 1)
        res["sw_bulk2_kernel_issue_active_pct"] = round(num(big["sw_bulk2_kernel"], "smsp__issue_active.avg.pct_of_peak_sustained_active"), 1)
    if "seed_kernel" in big:
        res["seed_kernel_issue_active_pct"] = round(num(big["seed_kernel"], "smsp__issue_active.avg.pct_of_peak_sustained_active"), 1)
        res["seed_kernel_lanes_per_instr"] = round(num(big["seed_kernel"], "smsp__thread_inst_executed_per_inst_executed.ratio"), 1)
        if len(sys.argv) > 3:
            res["seed_kernel_warp_instr_per_read"] = round(num(big["seed_kernel"], "smsp__inst_executed.sum") / float(sys.argv[3]))
    res["_source"] = os.path.basename(path) + " (ncu --set full; largest launch of each kernel" + (f", {sys.argv[3]} x 150 bp reads" if len(sys.argv) > 3 else "") + ")"
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "traffic.json"), "w") as fh:
        json.dump(res, fh, indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](sys.argv[2])
