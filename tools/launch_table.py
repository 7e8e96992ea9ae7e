"""Per-kernel table out of an ncu launch list (--csv --metrics gpu__time_duration.sum,smsp__inst_executed.sum,
smsp__thread_inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum):

    python tools/launch_table.py LIST.csv [--last N] [--views V]

Sums per kernel name over the last N launches (default: the last render chain, found from its k_prepare), divided
by V views."""
import argparse
import csv
import sys
from collections import OrderedDict


def read(path):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    cur = None
    for r in rd:
        key = r["ID"]
        if cur is None or cur["id"] != key:
            cur = {"id": key, "name": r["Kernel Name"].split("(")[0].replace("void ", "").split("<")[0], "grid": r.get("Grid Size", ""), "m": {}}
            rows.append(cur)
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        unit = r["Metric Unit"]
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        cur["m"][r["Metric Name"]] = v * scale
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--last", type=int, default=0)
    ap.add_argument("--views", type=int, default=1)
    ap.add_argument("--each", action="store_true", help="list every launch instead of sums per kernel")
    ap.add_argument("--json", default=None, help="also write {warp_instructions_per_panorama, dram_bytes_per_panorama, ...}")
    ap.add_argument("--source", default=None, help="what the list is (goes into the json)")
    a = ap.parse_args()
    rows = read(a.csv)
    if a.last:
        rows = rows[-a.last:]
    else:
        k = max(i for i, r in enumerate(rows) if r["name"] == "k_prepare")
        rows = rows[k:]
    agg = OrderedDict()
    for r in rows:
        key = r["id"] if a.each else r["name"]
        e = agg.setdefault(key, {"name": r["name"], "n": 0, "us": 0.0, "inst": 0.0, "tinst": 0.0, "rd": 0.0, "wr": 0.0, "grid": r["grid"]})
        e["n"] += 1
        e["us"] += r["m"].get("gpu__time_duration.sum", 0.0)
        e["inst"] += r["m"].get("smsp__inst_executed.sum", 0.0)
        e["tinst"] += r["m"].get("smsp__thread_inst_executed.sum", 0.0)
        e["rd"] += r["m"].get("dram__bytes_read.sum", 0.0)
        e["wr"] += r["m"].get("dram__bytes_write.sum", 0.0)
    V = a.views
    tot = {k: sum(e[k] for e in agg.values()) for k in ("us", "inst", "rd", "wr", "n")}
    print("%-12s %4s %9s %7s %9s %8s %9s %9s  %s" % ("kernel", "n", "us", "share", "Minst", "thr/wrp", "rd_MB", "wr_MB", "grid"))
    for e in agg.values():
        print("%-12s %4d %9.1f %6.1f%% %9.3f %8.1f %9.2f %9.2f  %s" % (
            e["name"], e["n"], e["us"] / V, 100 * e["us"] / tot["us"], e["inst"] / 1e6 / V,
            e["tinst"] / e["inst"] if e["inst"] else 0, e["rd"] / 1e6 / V, e["wr"] / 1e6 / V, e["grid"]))
    print("%-12s %4d %9.1f %7s %9.3f %8s %9.2f %9.2f   (per view, %d view(s))" % (
        "total", tot["n"], tot["us"] / V, "", tot["inst"] / 1e6 / V, "", tot["rd"] / 1e6 / V, tot["wr"] / 1e6 / V, V))


    if a.json:
        import json
        json.dump({"warp_instructions_per_panorama": tot["inst"] / V, "dram_bytes_per_panorama": (tot["rd"] + tot["wr"]) / V,
                   "kernel_us_sum_per_panorama": tot["us"] / V, "launches": int(tot["n"]), "views_per_launch": V,
                   "per_kernel_Minst": {e["name"]: round(e["inst"] / 1e6 / V, 3) for e in agg.values()},
                   "source": a.source or a.csv}, open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()
