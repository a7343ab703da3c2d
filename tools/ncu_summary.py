#!/usr/bin/env python
"""Condense `ncu -i X.ncu-rep --page raw --csv` into the small per-kernel table kept under
profiles/ (one column per captured launch).

    ncu_summary.py raw.csv > summary.csv
    ncu_summary.py raw.csv --register C2 --n 1000000 --summary profiles/X_summary.csv \
                   --hash <kernel source hash at capture time> [--launch K] [--match row_kernel]

--register also updates profiles/ncu_constants.json: DRAM bytes and warp instructions per launch
of the dominant kernel (summed over the launches of one step when --match selects several, e.g.
the main and the long-row pass of C3), tagged with the hash of the kernel sources the capture
was taken with (tools/srchash.py; tools/profile_capture.sh writes it next to the report).
bench.py reports these constants only while the sources still hash to the same value."""
import argparse
import csv
import json
import os
import sys

KEEP = [
    "Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum",
    "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
    "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__warps_eligible.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "sm__cycles_elapsed.max",
]


def num(x):
    return float(str(x).replace(",", "")) if x not in ("", None) else 0.0


def to_bytes(v, unit):
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return num(v) * scale.get(unit, 1.0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("raw")
    ap.add_argument("--register")
    ap.add_argument("--n", type=int, default=0)
    ap.add_argument("--summary", default="")
    ap.add_argument("--hash", default="")
    ap.add_argument("--sass-hash", default="", help="SASS digest of the workload's kernels at capture time")
    ap.add_argument("--match", default="row_kernel")
    ap.add_argument("--launches", default="", help="comma-separated launch indices forming one step")
    a = ap.parse_args()
    rows = list(csv.reader(open(a.raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    out = sys.stdout if not a.summary else open(a.summary, "w", newline="")
    w = csv.writer(out)
    w.writerow(["metric", "unit"] + ["launch%d" % i for i in range(len(data))])
    for k in KEEP:
        if k in hdr:
            i = hdr.index(k)
            w.writerow([k, units[i]] + [r[i] for r in data])
    if a.summary:
        out.close()
    if not a.register:
        return
    col = {k: hdr.index(k) for k in hdr}
    name_i = col["Kernel Name"]
    sel = [r for r in data if a.match in r[name_i]]
    if a.launches:
        idx = [int(x) for x in a.launches.split(",")]
        sel = [data[i] for i in idx]
    assert sel, "no launch matches"

    def total(metric, as_bytes=False):
        i = col[metric]
        return sum(to_bytes(r[i], units[i]) if as_bytes else num(r[i]) for r in sel)

    main_launch = max(sel, key=lambda r: num(r[col["gpu__time_duration.sum"]]))
    entry = {
        "summary": a.summary, "kernel_sources": a.hash, "kernel_sass": a.sass_hash, "N": a.n,
        "kernels": [r[name_i][:120] for r in sel],
        "dram_bytes": total("dram__bytes_read.sum", True) + total("dram__bytes_write.sum", True),
        "warp_instructions": total("smsp__inst_executed.sum"),
        "gpu_time_us_under_ncu": sum(
            num(r[col["gpu__time_duration.sum"]])
            * {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6}.get(units[col["gpu__time_duration.sum"]], 1.0)
            for r in sel),
        "l1_hit_pct": num(main_launch[col["l1tex__t_sector_hit_rate.pct"]]),
        "l2_hit_pct": num(main_launch[col["lts__t_sector_hit_rate.pct"]]),
        "issue_active_pct": num(main_launch[col["smsp__issue_active.avg.pct_of_peak_sustained_active"]]),
        "registers": num(main_launch[col["launch__registers_per_thread"]]),
    }
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles",
                        "ncu_constants.json")
    table = json.load(open(path)) if os.path.exists(path) else {}
    table[a.register] = entry
    json.dump(table, open(path, "w"), indent=1, sort_keys=True)
    print("registered", a.register, json.dumps(entry), file=sys.stderr)


if __name__ == "__main__":
    main()
