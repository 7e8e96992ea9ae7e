#!/usr/bin/env python
"""BASELINE configs[0], the reference's own CPU-runnable case: standalone offscreen render, 1 viewpoint, synthetic 2x2
SRTM3 tiles, 3600x300 panorama + range image on the CPU via Mesa llvmpipe.  Times the unmodified reference on
llvmpipe (oracle/_ref/libhorizonator_mesa.so), the same sources on the oracle's GL restatement, the oracle port and --
where a CUDA device is present -- the product, all on the same inputs; checks every result against llvmpipe's.
TEST INFRASTRUCTURE (imports oracle/).  Prints one JSON line."""
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

C1_LAT, C1_LON = 35.0 + 1.0 / 2400.0, -117.0 + 1.0 / 2400.0
W, H, R = 3600, 300, 1200
VIEW = dict(znear=100., zfar=100000.)


def timed(render, n):
    render()
    t = []
    for _ in range(n):
        t0 = time.perf_counter(); out = render(); t.append(time.perf_counter() - t0)
    return sorted(t)[len(t) // 2], out


def main():
    from oracle import binding
    from tools import synth
    from compare import compare_renders
    tiles = synth.config1_tiles(os.path.join(tempfile.mkdtemp(prefix="hz_c0_"), "c1"))
    cores = os.cpu_count() or 1
    res = {"config": "BASELINE configs[0]: 3600x300, 2x2 synthetic SRTM3 tiles, R=1200 (11.5 M triangles), full circle",
           "host_cores": cores}
    m = binding.MesaReference(C1_LAT, C1_LON, W, H, dir_dems=tiles, render_radius_cells=R, threads=cores)
    t, (img_m, rng_m) = timed(lambda: m.render(-180.05, 179.95, **VIEW), 5)
    res["reference_on_llvmpipe"] = {"ms_per_render": t * 1e3, "gl": m.gl_strings()}
    m.close()
    for name, cls in (("reference_on_restated_gl", binding.Reference), ("oracle_port", binding.Oracle)):
        o = cls(C1_LAT, C1_LON, W, H, dir_dems=tiles, render_radius_cells=R, threads=cores)
        t, (img, rng) = timed(lambda: o.render(-180.05, 179.95, **VIEW), 5)
        s = compare_renders(img, rng, img_m, rng_m)
        res[name] = {"ms_per_render": t * 1e3, "vs_llvmpipe": {k: s[k] for k in ("coverage_agreement", "agreement", "range_mismatch", "off_silhouette", "ok")}}
        o.close()
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    if have_gpu:
        import horizonator_b200 as hz
        h = hz.horizonator(C1_LAT, C1_LON, W, H, dir_dems=tiles, render_radius_cells=R)
        t, (img, rng) = timed(lambda: h.render(-180.05, 179.95, **VIEW), 50)
        s = compare_renders(img, rng, img_m, rng_m)
        res["b200_python_render"] = {"ms_per_render": t * 1e3, "vs_llvmpipe": {k: s[k] for k in ("coverage_agreement", "agreement", "range_mismatch", "off_silhouette", "ok")}}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
