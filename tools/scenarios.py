#!/usr/bin/env python
"""Timings of the BASELINE.json configs other than the benchmark one (bench.py measures configs[1]):

  C3  Python render() pan/zoom sweep on the C2 context: per-call latency (host numpy arrays out)
  C4  ultra-high-resolution 360-degree panorama 36000x4000 on one GPU, whole and as 8 sequential wedges
  C5  viewpoint batch on a grid over the central degree (device outputs), panoramas/s

Prints one JSON object; run on the GPU box:  python tools/scenarios.py > gpurun_out/scenarios.json
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

C2_LAT, C2_LON = 34.0 + 1.0 / 7200.0, -117.0 + 1.0 / 7200.0


def main():
    import torch
    import horizonator_b200 as hz
    from tools import synth
    tiles = synth.config2_tiles(os.environ.get("HZ_BENCH_TILES", "/tmp/hz_tiles_c2"))
    out = {}

    # ---------------------------------------------------------------- C3
    h = hz.horizonator(C2_LAT, C2_LON, 3600, 600, SRTM1=True, dir_dems=tiles, render_radius_m=150000.)
    calls = [(c - 45.0, c + 45.0) for c in np.linspace(0., 360., 60, endpoint=False)]
    calls += [(90.0 - s / 2, 90.0 + s / 2) for s in np.linspace(180., 10., 40)]
    for a0, a1 in calls[:3]:
        h.render(a0, a1, znear=100., zfar=150000.)
    lat = []
    for a0, a1 in calls:
        t0 = time.perf_counter()
        h.render(a0, a1, znear=100., zfar=150000.)
        lat.append((time.perf_counter() - t0) * 1e3)
    lat = np.array(lat)
    out["C3_pan_zoom_sweep"] = dict(calls=len(calls), ms_median=float(np.median(lat)), ms_p95=float(np.percentile(lat, 95)),
                                    ms_max=float(lat.max()), note="h.render() returning fresh numpy arrays (pooled page-locked blocks), 90-degree "
                                    "windows stepping round the circle then zooming 180 -> 10 degrees")

    # ---------------------------------------------------------------- C5
    g = 8
    views = [(33.5 + (j + 0.5) / g + 1.0 / 7200.0, -117.5 + (i + 0.5) / g + 1.0 / 7200.0, -180.05, 179.95)
             for j in range(g) for i in range(g)]
    h.set_zextents(100., 150000.)
    B = 16
    d_img = torch.empty((B, 600, 3600, 3), dtype=torch.uint8, device="cuda")
    d_rng = torch.empty((B, 600, 3600), dtype=torch.float32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    h.render_batch_device(views[:B], d_img.data_ptr(), d_rng.data_ptr(), st)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(0, len(views), B):
        h.render_batch_device(views[k:k + B], d_img.data_ptr(), d_rng.data_ptr(), st)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    per_view = []
    for v in views[::7]:
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        h.render_batch_device([v], d_img.data_ptr(), d_rng.data_ptr(), st)
        torch.cuda.synchronize()
        per_view.append((time.perf_counter() - t1) * 1e3)
    out["C5_viewpoint_grid"] = dict(viewpoints=len(views), panoramas_per_s=len(views) / dt,
                                    lone_render_ms_min=float(min(per_view)), lone_render_ms_max=float(max(per_view)),
                                    note="8x8 grid over the central degree, 3600x600 full circle each, device outputs, "
                                    "16 in flight")
    del h, d_img, d_rng

    # ---------------------------------------------------------------- C4
    W, H = 36000, 4000
    h = hz.horizonator(C2_LAT, C2_LON, W, H, SRTM1=True, dir_dems=tiles, render_radius_m=150000.)
    h.set_zextents(100., 150000.)
    h.pan_zoom(-180.005, 179.995)
    d_img = torch.empty((H, W, 3), dtype=torch.uint8, device="cuda")
    d_rng = torch.empty((H, W), dtype=torch.float32, device="cuda")
    v = [(C2_LAT, C2_LON, -180.005, 179.995)]
    for _ in range(2):
        h.render_batch_device(v, d_img.data_ptr(), d_rng.data_ptr(), st)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 5
    for _ in range(n):
        h.render_batch_device(v, d_img.data_ptr(), d_rng.data_ptr(), st)
    torch.cuda.synchronize()
    whole_ms = (time.perf_counter() - t0) / n * 1e3
    stats = h.last_render_stats()
    hit = float((d_rng > 0).float().mean().item())
    ref_sum = int(d_img.sum(dtype=torch.int64).item())
    G = 8
    edges = [W * k // G for k in range(G + 1)]
    slabs = [(torch.empty((H, edges[k + 1] - edges[k], 3), dtype=torch.uint8, device="cuda"),
              torch.empty((H, edges[k + 1] - edges[k]), dtype=torch.float32, device="cuda")) for k in range(G)]
    h.move(C2_LAT, C2_LON)
    wedge_ms = []
    for k in range(G):
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        h.render_wedge_device(edges[k], edges[k + 1], slabs[k][0].data_ptr(), slabs[k][1].data_ptr(), st)
        torch.cuda.synchronize()
        wedge_ms.append((time.perf_counter() - t1) * 1e3)
    stitched = torch.cat([s[0] for s in slabs], dim=1)
    same = bool(torch.equal(stitched, d_img)) and bool(torch.equal(torch.cat([s[1] for s in slabs], dim=1), d_rng))
    out["C4_36000x4000"] = dict(whole_ms=whole_ms, wedge_ms=wedge_ms, wedges_equal_whole=same, terrain_pixel_fraction=hit,
                                image_checksum=ref_sum, stats=stats,
                                note="one B200; wedge_ms = the 8 azimuth wedges one after the other on the same GPU "
                                "(what each of 8 GPUs would do in parallel before one all_gather)")
    print(json.dumps(out))


if __name__ == "__main__":
    main()
