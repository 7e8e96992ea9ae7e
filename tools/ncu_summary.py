#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box): key raw metrics + hottest source lines.

usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [top_lines]
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__grid_size", "launch__block_size",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum.per_cycle_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum",
]
STALLS = "smsp__average_warps_issue_stalled_"


def run(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    rows = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        print("== kernel:", name)
        d = dict(zip(hdr, zip(units, vals)))
        for k in KEYS:
            if k in d:
                print("  %-70s %s %s" % (k, d[k][1], d[k][0]))
        st = sorted(((float(v[1]), k[len(STALLS):-len("_per_issue_active.ratio")]) for k, v in d.items()
                     if k.startswith(STALLS) and k.endswith("_per_issue_active.ratio")), reverse=True)
        print("  stalls per issue:", ", ".join("%s %.2f" % (n, x) for x, n in st[:8]))
    rows = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]))))
    hdr = None
    lines = []
    for r in rows:
        if r and r[0] == "Line No":
            hdr = r
            iS, iI = hdr.index("# Samples"), hdr.index("Instructions Executed")
            continue
        if hdr and r and r[0].strip().isdigit():
            try:
                lines.append((int(r[0]), r[1].strip()[:100], int(r[iS]), int(r[iI])))
            except ValueError:
                pass
    tS = sum(l[2] for l in lines) or 1
    tI = sum(l[3] for l in lines) or 1
    print("== hottest source lines (of %d samples, %d warp-instructions)" % (tS, tI))
    for l in sorted(lines, key=lambda l: -l[2])[:top]:
        print("  %4d  %5.2f%% samples  %5.2f%% inst   %s" % (l[0], 100.0 * l[2] / tS, 100.0 * l[3] / tI, l[1]))


if __name__ == "__main__":
    main()
