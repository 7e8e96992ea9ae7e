"""Cost of single renders of named views of the C2 context (lone-view latency, per-stage times, culling counters):

    python tools/view_probe.py [--out FILE.json] [--grid] [NAME ...]      # table of the named views (default: all)
    python tools/view_probe.py --ncu NAME [--reps 3]                      # just renders NAME a few times (for ncu)

Views: c2 (BASELINE configs[1]), eye3km / eye6km / eye12km (explicit eye height above sea level), zoom30 / zoom10 /
zoom5 / zoom2 (azimuth span in degrees around 45), gridworst / gridmedian (picked from the 8x8 viewpoint grid over the
central degree by --grid, else a recorded position).  --grid also lists min/median/max over the grid.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import horizonator_b200 as hz  # noqa: E402
from tools import synth  # noqa: E402

C2_LAT, C2_LON = 34.0 + 1.0 / 7200.0, -117.0 + 1.0 / 7200.0
FULL = (-180.05, 179.95)


def grid_views(g=8):
    return [(33.5 + (j + 0.5) / g + 1.0 / 7200.0, -117.5 + (i + 0.5) / g + 1.0 / 7200.0) for j in range(g) for i in range(g)]


VIEWS = {
    "c2": (C2_LAT, C2_LON, FULL[0], FULL[1], -1.),
    "eye3km": (C2_LAT, C2_LON, FULL[0], FULL[1], 3000.),
    "eye6km": (C2_LAT, C2_LON, FULL[0], FULL[1], 6000.),
    "eye12km": (C2_LAT, C2_LON, FULL[0], FULL[1], 12000.),
    "zoom30": (C2_LAT, C2_LON, 30., 60., -1.),
    "zoom10": (C2_LAT, C2_LON, 40., 50., -1.),
    "zoom5": (C2_LAT, C2_LON, 42.5, 47.5, -1.),
    "zoom2": (C2_LAT, C2_LON, 44., 46., -1.),
    # positions of the 8x8 grid (tools/view_probe.py --grid on B200 picks them again)
    "gridworst": (33.5 + 1.5 / 8 + 1.0 / 7200.0, -117.5 + 5.5 / 8 + 1.0 / 7200.0, FULL[0], FULL[1], -1.),
    "gridmedian": (33.5 + 1.5 / 8 + 1.0 / 7200.0, -117.5 + 1.5 / 8 + 1.0 / 7200.0, FULL[0], FULL[1], -1.),
}


class Probe:
    def __init__(self):
        tiles = synth.config2_tiles(os.environ.get("HZ_BENCH_TILES", "/tmp/hz_tiles_c2"))
        self.h = hz.horizonator(C2_LAT, C2_LON, 3600, 600, SRTM1=True, dir_dems=tiles, render_radius_m=150000.)
        self.h.set_zextents(100., 150000.)
        self.d_img = torch.empty((600, 3600, 3), dtype=torch.uint8, device="cuda")
        self.d_rng = torch.empty((600, 3600), dtype=torch.float32, device="cuda")
        self.st = torch.cuda.Stream()        # (stream 0 would make the batch call synchronous)
        torch.cuda.set_stream(self.st)

    def render(self, v, n=1):
        for _ in range(n):
            self.h.render_batch_device([v], self.d_img.data_ptr(), self.d_rng.data_ptr(), self.st.cuda_stream)

    def lone_ms(self, v, reps=10):
        self.render(v, 3)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(self.st)
        self.render(v, reps)
        e1.record(self.st)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    def details(self, v):
        h = self.h
        ms = self.lone_ms(v)
        h.profile(True); h.profile_read()
        self.render(v, 3)
        torch.cuda.synchronize()
        p = h.profile_read(); c = h.render_counters(); s = h.last_render_stats(); h.profile(False)
        return {"view": list(v), "lone_ms": round(ms, 4), "terrain_fraction": round(float((self.d_rng > 0).float().mean().item()), 4),
                "stage_us": {k: round(p[k] * 1e3, 1) for k in ("prepare", "near", "big_near", "march", "big_far", "resolve")},
                "counters": c, "big_entries": s["big_entries"], "launches": s["launches"]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--grid", action="store_true")
    ap.add_argument("--ncu", default=None)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("names", nargs="*")
    a = ap.parse_args()
    picks = os.path.join(ROOT, "gpurun_out", "grid_picks.json")
    if os.path.exists(picks) and not a.grid:
        for k, v in json.load(open(picks)).items():
            VIEWS[k] = tuple(v)
    p = Probe()
    if a.ncu:
        p.render(VIEWS[a.ncu], a.reps)
        torch.cuda.synchronize()
        return
    report = {}
    if a.grid:
        cost = [(p.lone_ms((la, lo) + FULL + (-1.,), 5), la, lo) for la, lo in grid_views()]
        cost.sort()
        ms = [c[0] for c in cost]
        VIEWS["gridworst"] = (cost[-1][1], cost[-1][2]) + FULL + (-1.,)
        VIEWS["gridmedian"] = (cost[len(cost) // 2][1], cost[len(cost) // 2][2]) + FULL + (-1.,)
        report["grid_8x8_lone_ms"] = {"min": round(ms[0], 4), "median": round(float(np.median(ms)), 4), "max": round(ms[-1], 4),
                                      "mean": round(float(np.mean(ms)), 4), "worst_at": list(cost[-1][1:]),
                                      "median_at": list(cost[len(cost) // 2][1:])}
        print("grid:", json.dumps(report["grid_8x8_lone_ms"]), flush=True)
        os.makedirs(os.path.dirname(picks), exist_ok=True)
        json.dump({k: list(VIEWS[k]) for k in ("gridworst", "gridmedian")}, open(picks, "w"))
    for name in (a.names or list(VIEWS)):
        report[name] = p.details(VIEWS[name])
        print(name, json.dumps(report[name]), flush=True)
    if a.out:
        json.dump(report, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
