#!/usr/bin/env python
"""TEST INFRASTRUCTURE.  Random scenes rendered twice on the CPU -- by the unmodified reference on Mesa llvmpipe
(oracle/_ref/libhorizonator_mesa.so) and by the oracle (oracle/liboracle.so) -- and compared at the north_star
tolerances (tests/compare.py).  No GPU involved: this is how far the oracle's restated GL rules are from a real
OpenGL driver, over many more views than the committed fixtures hold.

    python tools/llvmpipe_sweep.py [--scenes 60] [--seed 1] [--out profiles/r01C_oracle_vs_llvmpipe_sweep.json]

The 15 fixed EDGE_CASES run first, then --scenes random ones.

Prints one JSON document: per scene the parameters and the comparison, then the totals.
"""
import argparse
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

C1_LAT, C1_LON = 35.0 + 1.0 / 2400.0, -117.0 + 1.0 / 2400.0


def random_scene(rs):
    W = int(rs.integers(48, 700))
    H = int(rs.integers(24, 260))
    R = int(rs.integers(40, 420))
    kind = rs.choice(["circle", "wide", "narrow", "seam"])
    if kind == "circle":
        az0 = float(rs.uniform(-360., 0.)); az1 = az0 + 360.
    elif kind == "wide":
        az0 = float(rs.uniform(-180., 180.)); az1 = az0 + float(rs.uniform(60., 220.))
    elif kind == "narrow":
        az0 = float(rs.uniform(-180., 180.)); az1 = az0 + float(rs.uniform(1.5, 30.))
    else:
        c = 180. + float(rs.uniform(-20., 20.)); half = float(rs.uniform(5., 60.)); az0, az1 = c - half, c + half
    # eye inside the loaded square, up to 30 % of the radius off its centre (cells of 1/1200 degree)
    off = 0.3 * R / 1200.
    lat = C1_LAT + float(rs.uniform(-off, off)) if rs.random() < 0.5 else None
    lon = C1_LON + float(rs.uniform(-off, off)) if lat is not None else None
    viewer_z = None if rs.random() < 0.6 else float(rs.uniform(800., 4500.))
    znear = float(rs.choice([10., 50., 100., 300.]))
    zfar = float(rs.choice([8000., 20000., 40000., 100000., 150000.]))
    znc = znear if rs.random() < 0.6 else float(rs.uniform(znear, 2000.))
    zfc = zfar if rs.random() < 0.6 else float(rs.uniform(3000., zfar))
    holes = bool(rs.random() < 0.2)
    return dict(W=W, H=H, R=R, kind=str(kind), az0=az0, az1=az1, lat=lat, lon=lon, viewer_z=viewer_z,
                znear=znear, zfar=zfar, znear_color=znc, zfar_color=zfc, holes=holes)


_BASE = dict(W=400, H=120, R=200, kind="edge", az0=-180.05, az1=179.95, lat=None, lon=None, viewer_z=None,
             znear=100., zfar=100000., znear_color=100., zfar_color=100000., holes=False)
# fixed scenes at the corners of the parameter space: windows wider than a circle (span 720: nothing is drawn, by
# either), near-plane clipping through the terrain, a far plane short of it, an eye outside the loaded square, extreme
# aspect ratios, one-pixel images, a two-cell mesh
EDGE_CASES = [dict(_BASE, kind=name, **upd) for name, upd in [
    ("span400", dict(az0=-200., az1=200.)),
    ("span720", dict(az0=-360., az1=360.)),
    ("znear2km", dict(znear=2000., znear_color=2000.)),
    ("znear5km", dict(znear=5000., znear_color=5000., az0=20., az1=50.)),
    ("zfar3km", dict(zfar=3000., zfar_color=3000.)),
    ("tall", dict(W=90, H=400, az0=10., az1=40.)),
    ("outside", dict(lat=C1_LAT + 0.25, lon=C1_LON + 0.2)),
    ("far_out", dict(lat=C1_LAT + 0.6, lon=C1_LON - 0.5, az0=180., az1=270.)),
    ("low_eye", dict(viewer_z=5.0)),
    ("W1", dict(W=1, H=50, az0=44., az1=46.)),
    ("H1", dict(W=300, H=1)),
    ("az<-360", dict(az0=-400., az1=-300.)),
    ("az>700", dict(az0=700., az1=800.)),
    ("R2", dict(R=2)),
    ("zoom1", dict(az0=100., az1=101., W=300, H=200)),
]]


def render_pair(scene, tiles, tiles_holes, threads):
    from oracle import binding
    d = tiles_holes if scene["holes"] else tiles
    kw_init = dict(dir_dems=d, render_radius_cells=scene["R"], viewer_z=scene["viewer_z"])
    kw = dict(znear=scene["znear"], zfar=scene["zfar"], znear_color=scene["znear_color"], zfar_color=scene["zfar_color"])
    if scene["lat"] is not None:
        kw.update(lat=scene["lat"], lon=scene["lon"])
    m = binding.MesaReference(C1_LAT, C1_LON, scene["W"], scene["H"], threads=threads, **kw_init)
    try:
        a = m.render(scene["az0"], scene["az1"], **kw)
        vz_m = m.viewer_z
    finally:
        m.close()
    o = binding.Oracle(C1_LAT, C1_LON, scene["W"], scene["H"], threads=threads, **kw_init)
    b = o.render(scene["az0"], scene["az1"], **kw)
    vz_o = o.viewer_z
    o.close()
    return a, b, vz_m, vz_o


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scenes", type=int, default=60)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--out", default=None)
    ap.add_argument("--no-edge-cases", action="store_true", help="random scenes only")
    ap.add_argument("--dump", default=None, metavar="DIR",
                    help="also save scenes.json and the llvmpipe renders (scene_%%03d.npz) there, for a third renderer "
                         "to be compared with (tests/test_zz_gpu_fullsize_llvmpipe.py does that with the CUDA path)")
    args = ap.parse_args()
    from oracle import binding
    from tools import synth
    from compare import compare_renders
    if not binding.have_mesa():
        raise SystemExit("oracle/_ref/libhorizonator_mesa.so (or the image's Mesa libGL) is absent")
    tmp = tempfile.mkdtemp(prefix="hz_sweep_")
    tiles = synth.config1_tiles(os.path.join(tmp, "c1"))
    holes = os.path.join(tmp, "holes")
    synth.write_tiles(holes, (34, 35), (-118, -117), seed=7, skip=((35, -118), (34, -117)))
    open(os.path.join(holes, synth.tile_name(34, -117)), "wb").close()

    if args.dump:
        os.makedirs(args.dump, exist_ok=True)
    rs = np.random.default_rng(args.seed)
    threads = min(8, os.cpu_count() or 1)
    scenes, tot = [], dict(pixels=0, terrain_pixels=0, coverage_mismatch=0, range_mismatch=0, off_silhouette=0,
                           range_bit_identical=0, both_hit=0, not_ok=0, eye_height_differs=0)
    worst = 1.0
    todo = ([] if args.no_edge_cases else list(EDGE_CASES)) + [random_scene(rs) for _ in range(args.scenes)]
    for k, sc in enumerate(todo):
        (img_m, rng_m), (img_o, rng_o), vz_m, vz_o = render_pair(sc, tiles, holes, threads)
        if args.dump:
            np.savez_compressed(os.path.join(args.dump, "scene_%03d.npz" % k), image=img_m, ranges=rng_m,
                                viewer_z=np.float32(vz_m))
        s = compare_renders(img_o, rng_o, img_m, rng_m)          # llvmpipe is the reference side
        both = (rng_m > 0) & (rng_o > 0)
        cov = int(((rng_m > 0) != (rng_o > 0)).sum())
        tot["pixels"] += rng_m.size; tot["terrain_pixels"] += int((rng_m > 0).sum()); tot["coverage_mismatch"] += cov
        tot["range_mismatch"] += s["range_mismatch"]; tot["off_silhouette"] += s["off_silhouette"]
        tot["range_bit_identical"] += int((rng_m[both] == rng_o[both]).sum()); tot["both_hit"] += int(both.sum())
        tot["not_ok"] += 0 if s["ok"] else 1
        tot["eye_height_differs"] += 0 if np.float32(vz_m) == np.float32(vz_o) else 1
        worst = min(worst, s["agreement"])
        scenes.append(dict(scene=sc, coverage_mismatch=cov, **{kk: s[kk] for kk in (
            "pixels", "hit_fraction_ref", "coverage_agreement", "range_mismatch", "agreement", "off_silhouette",
            "max_rel_range_err_where_agree", "red_max_diff_where_agree", "ok")}))
        sys.stderr.write("%3d %-7s %4dx%-4d R=%-4d hit %.3f  cov_mismatch %d  range_mismatch %d  off_sil %d  ok %s\n" % (
            k, sc["kind"], sc["W"], sc["H"], sc["R"], s["hit_fraction_ref"], cov, s["range_mismatch"],
            s["off_silhouette"], s["ok"]))
    tot["worst_agreement"] = worst
    tot["coverage_agreement_overall"] = 1.0 - tot["coverage_mismatch"] / tot["pixels"]
    doc = dict(what="oracle (oracle/liboracle.so) vs the unmodified reference on Mesa llvmpipe, random scenes on the "
                    "synthetic SRTM3 tiles: %d fixed edge cases + %d random (seed %d); tolerances of tests/compare.py"
                    % (0 if args.no_edge_cases else len(EDGE_CASES), args.scenes, args.seed),
               totals=tot, scenes=scenes)
    if args.dump:
        with open(os.path.join(args.dump, "scenes.json"), "w") as f:
            json.dump(dict(holes_skip=[[35, -118], [34, -117]], scenes=todo), f)
    text = json.dumps(doc, indent=1)
    if args.out:
        with open(args.out, "w") as f:
            f.write(text + "\n")
    print(json.dumps(tot))


if __name__ == "__main__":
    main()
