#!/usr/bin/env python
"""bench.py -- panoramas/s of the render hot path on N B200s (BASELINE.json metric), one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Workload (N=1 and per rank for N>1): BASELINE.json configs[1] ("C2") -- synthetic SRTM1 tiles N32..N35 x
W119..W116, viewer (34+1/7200, -117+1/7200), render_radius_m = 150 km (R = 5858 cells, 137 M vertices, 274 M
triangles), 3600x600 panorama + range image, az [-180.05, 179.95], znear 100 m, zfar 150 km.  A step is one
call of horizonator_render_batch_device() with --batch (default 16) panoramas per rank, which the library renders
concurrently on its render lanes (the kernels of one panorama are short and latency-bound; several in flight
fill the machine).  With N>1 every rank does the same against its own copy of the DEM (weak scaling, no
collective on the data path; NCCL only for the barrier/max of times).

value      device-resident throughput: K steps x batch panoramas / time; outputs stay in HBM, CUDA events on the
           stream the work is queued on, max over ranks
e2e        the same metric through the reference-facing call horizonator_render_offscreen(): one panorama per
           call into (page-locked) HOST buffers, device->host copy inside the timed region
roofline   achieved = algorithmic bytes per panorama (2*(2R)^2 + 7*W*H) x measured panoramas/s against
           MEASURED_PEAKS.json hbm_gbs; plus the per-stage CUDA-event times and the latency of a lone panorama
cpu_baseline  the CPU oracle timed on the host cores (rank 0, N=1), a bounded sample of the same workload

--impl reference: the reference's own horizonator-lib.c + dem.c (compiled unmodified, oracle/_ref) rendering
the same workload on the host cores through a software GL restatement (no GL driver can run in this image).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

C2 = dict(lat=34.0 + 1.0 / 7200.0, lon=-117.0 + 1.0 / 7200.0, W=3600, H=600, radius_m=150000.0,
          az0=-180.05, az1=179.95, znear=100.0, zfar=150000.0, R=5858)
TILES_DIR = os.environ.get("HZ_BENCH_TILES", "/tmp/hz_tiles_c2")
WARP_INST_PER_PANORAMA = 42.29e6    # ncu smsp__inst_executed.sum, all kernels of one C2 panorama (profiles/r01A_*)


def algorithmic_bytes(R, W, H):
    return 2 * (2 * R) ** 2 + 7 * W * H


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock + clock-event (throttle) reasons sampled DURING the timed region: NVML polled every ~2 ms from a
    thread (the region can be shorter than nvidia-smi's sampling period); nvidia-smi -lms as the fallback
    (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.samples = []          # (time, sm_mhz, sm_max_mhz, set of reasons)
        self.stop_flag = False
        self.thread = None
        self.how = None

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[self.gpu])
            except (ValueError, IndexError):
                pass
        return self.gpu

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            hnd = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
            smmax = float(pynvml.nvmlDeviceGetMaxClockInfo(hnd, pynvml.NVML_CLOCK_SM))
            names = (("hw_slowdown", pynvml.nvmlClocksEventReasonHwSlowdown),
                     ("hw_thermal_slowdown", pynvml.nvmlClocksEventReasonHwThermalSlowdown),
                     ("sw_thermal_slowdown", pynvml.nvmlClocksEventReasonSwThermalSlowdown),
                     ("sw_power_cap", pynvml.nvmlClocksEventReasonSwPowerCap))

            def poll():
                while not self.stop_flag:
                    try:
                        mhz = float(pynvml.nvmlDeviceGetClockInfo(hnd, pynvml.NVML_CLOCK_SM))
                        bits = pynvml.nvmlDeviceGetCurrentClocksEventReasons(hnd)
                        self.samples.append((time.time(), mhz, smmax, {n for n, b in names if bits & b}))
                    except Exception:
                        pass
                    time.sleep(0.002)
            self.thread = threading.Thread(target=poll, daemon=True)
            self.thread.start()
            self.how = "nvml"
            return
        except Exception:
            pass
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self._physical_index()), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
            self.how = "nvidia-smi"
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                mhz, smmax = float(f[1]), float(f[2])
            except ValueError:
                continue
            names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
            self.samples.append((time.time(), mhz, smmax,
                                 {n for n, v in zip(names, f[5:9]) if v.lower().startswith("active")}))

    def stop(self, t0, t1):
        if self.how is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no NVML and no nvidia-smi"]}
        self.stop_flag = True
        if self.proc is not None:
            time.sleep(0.05)
            self.proc.terminate()
        elif self.thread is not None:
            self.thread.join(timeout=1.0)
        inside = [s for s in self.samples if t0 <= s[0] <= t1]
        if not inside:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples in the timed region"], "how": self.how}
        sm = sorted(s[1] for s in inside)
        reasons = set()
        for s in inside:
            reasons |= s[3]
        return {"sm_mhz": sm[len(sm) // 2], "sm_min_mhz": sm[0], "sm_max_mhz": max(s[2] for s in inside),
                "reasons": sorted(reasons), "samples": len(inside), "how": self.how}


def ensure_tiles(rank, barrier):
    from tools import synth
    if rank == 0:
        synth.config2_tiles(TILES_DIR)
    barrier()
    return TILES_DIR


def viewpoints(n_ranks, rank, steps):
    """Every rank renders the C2 viewpoint (BASELINE configs[1]) against its own copy of the DEM: the same work per
    GPU, so that the N-GPU value measures how the machine scales and not how viewpoints differ."""
    return [(C2["lat"], C2["lon"])] * steps


# ------------------------------------------------------------------------------------------------ our arm

class StdoutToStderr:
    """stdout carries exactly one JSON line.  Libraries print there too (NCCL writes its version banner to file
    descriptor 1 when NCCL_DEBUG=VERSION): while this is active, descriptor 1 points at stderr; restore() puts it
    back right before the line is printed."""

    def __init__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def restore(self):
        if self.saved is not None:
            sys.stdout.flush()
            os.dup2(self.saved, 1)
            os.close(self.saved)
            self.saved = None


def run_b200(args):
    quiet = StdoutToStderr()
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the renderer has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()

    tiles = ensure_tiles(rank, barrier)
    import horizonator_b200 as hz

    t_init = time.time()
    h = hz.horizonator(C2["lat"], C2["lon"], C2["W"], C2["H"], SRTM1=True, dir_dems=tiles,
                       render_radius_m=C2["radius_m"])
    t_init = time.time() - t_init
    R = h.context.dems.radius_cells
    W, H = C2["W"], C2["H"]
    h.set_zextents(C2["znear"], C2["zfar"])
    K, Wm = args.steps, args.warmup
    pts = viewpoints(world, rank, K + Wm)
    views = [(la, lo, C2["az0"], C2["az1"]) for la, lo in pts]

    B = args.batch                      # panoramas per step: rendered concurrently on the library's render lanes
    d_img = torch.empty((B, H, W, 3), dtype=torch.uint8, device="cuda")
    d_rng = torch.empty((B, H, W), dtype=torch.float32, device="cuda")
    stream = torch.cuda.current_stream()

    # ---- latency of one panorama at a time ----
    for k in range(Wm):
        h.render_batch_device(views[k:k + 1], d_img.data_ptr(), d_rng.data_ptr(), stream.cuda_stream)
    torch.cuda.synchronize()
    n_lat = min(K, 50)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for k in range(n_lat):
        h.render_batch_device(views[Wm:Wm + 1], d_img.data_ptr(), d_rng.data_ptr(), stream.cuda_stream)
    ev1.record(stream)
    torch.cuda.synchronize()
    latency_ms = ev0.elapsed_time(ev1) / n_lat
    # ... and the same with CUDA events around every stage and the culling counters switched on
    h.profile(True)
    h.profile_read()
    for k in range(20):
        h.render_batch_device(views[Wm:Wm + 1], d_img.data_ptr(), d_rng.data_ptr(), stream.cuda_stream)
    torch.cuda.synchronize()
    prof = h.profile_read()
    h.profile(False)
    stats = h.last_render_stats()
    counters = h.render_counters()

    # ---- device-resident throughput: K steps of B panoramas ----
    step_views = [views[Wm]] * B
    for k in range(Wm):
        h.render_batch_device(step_views, d_img.data_ptr(), d_rng.data_ptr(), stream.cuda_stream)
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.05)
    barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    wall0 = time.time()
    ev0.record(stream)
    for k in range(K):
        h.render_batch_device(step_views, d_img.data_ptr(), d_rng.data_ptr(), stream.cuda_stream)
    ev1.record(stream)
    host_enqueue_s = time.time() - wall0
    torch.cuda.synchronize()
    barrier()
    wall1 = time.time()
    ms_total = ev0.elapsed_time(ev1)
    clocks = sampler.stop(wall0, wall1) if rank == 0 else None

    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total_max = float(t.item())
    value = world * K * B / (ms_total_max / 1e3)

    # ---- end to end through the reference-facing call, host buffers ----
    import numpy as np
    import ctypes as C
    # host result buffers: page-locked (horizonator_host_alloc), reused across steps
    img = hz.pinned_array((H, W, 3), np.uint8)
    rng = hz.pinned_array((H, W), np.float32)
    ctx = C.byref(h.context)

    def e2e_step(la, lo):
        # what horizonator-pywrap.c's render() does per call, minus the numpy allocation
        assert hz.lib.horizonator_pan_zoom(ctx, C2["az0"], C2["az1"])
        assert hz.lib.horizonator_move(ctx, None, la, lo)
        assert hz.lib.horizonator_set_zextents(ctx, C2["znear"], C2["zfar"], C2["znear"], C2["zfar"])
        assert hz.lib.horizonator_render_offscreen(ctx, img.ctypes.data, rng.ctypes.data)

    for k in range(Wm):
        e2e_step(*pts[k])
    barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(K):
        e2e_step(*pts[Wm + k])
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * K / float(t.item())
    hit_fraction = float((rng > 0).mean())

    # same call into ordinary pageable numpy arrays (what an unmodified caller of the reference passes)
    img_p = np.empty((H, W, 3), np.uint8)
    rng_p = np.empty((H, W), np.float32)
    img_p[:] = 0; rng_p[:] = 0
    for k in range(2):
        hz.lib.horizonator_render_offscreen(ctx, img_p.ctypes.data, rng_p.ctypes.data)
    t0 = time.perf_counter()
    for k in range(K):
        hz.lib.horizonator_render_offscreen(ctx, img_p.ctypes.data, rng_p.ctypes.data)
    e2e_pageable = K / (time.perf_counter() - t0)

    # the additive host-pointer batch call: renders and device->host copies of different views overlap
    bimg = hz.pinned_array((B, H, W, 3), np.uint8)
    brng = hz.pinned_array((B, H, W), np.float32)
    varr = h._views(step_views)
    for k in range(2):
        assert hz.lib.horizonator_render_batch(ctx, B, varr, bimg.ctypes.data, brng.ctypes.data)
    n_b = max(3, K // B)
    t0 = time.perf_counter()
    for k in range(n_b):
        assert hz.lib.horizonator_render_batch(ctx, B, varr, bimg.ctypes.data, brng.ctypes.data)
    e2e_batch = n_b * B / (time.perf_counter() - t0)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak_gbs()
    alg = algorithmic_bytes(R, W, H)
    stages = {"prepare": "k_prepare", "near": "k_near+k_raster", "big_near": "k_big",
              "march": "bands: (k_tiles, k_blocks, k_mesh, k_raster) x 2", "big_far": "k_big", "resolve": "k_resolve"}
    # achieved: algorithmic bytes per panorama x panoramas/s of the whole device-resident job (kernels of
    # concurrent panoramas overlap, so a single kernel's duration no longer measures the machine)
    achieved = alg * (K * B / (ms_total / 1e3)) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "dram_bytes_per_panorama.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("dram_bytes_per_panorama")
        except Exception:
            traffic = None

    mosaic_ms = h.time_mosaic(5)
    out = {
        "metric": "panoramas/sec (SRTM1, 3600x600 px)",
        "value": value, "unit": "panoramas/s",
        "n_gpus": world, "steps": K, "warmup": Wm,
        "ms_per_step": ms_total_max / K,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {
            "workload": "BASELINE configs[1]: single viewpoint per GPU, synthetic SRTM1 4x4 tiles, 150 km radius "
                        "(R=%d cells, %d triangles), 3600x600 panorama + range image" % (R, h.context.Ntriangles),
            "az_deg": [C2["az0"], C2["az1"]], "znear_m": C2["znear"], "zfar_m": C2["zfar"],
            "panoramas_per_step_per_gpu": B,
            "concurrency": "the %d panoramas of a step render concurrently on up to 16 render lanes (one CUDA stream "
                           "and scratch set each) of one context; same viewpoint, nothing cached between them" % B,
            "l2": "no explicit flush; inputs larger than L2: the int16 DEM square is %.0f MB and its culling pyramid "
                  "%.0f MB, and the %d panoramas in flight cycle %d visibility buffers of %.0f MB each through the 126 MB "
                  "L2 (hierarchical culling makes one panorama touch only a few tens of MB of the DEM, see "
                  "roofline.traffic)" % (2 * (2 * R) ** 2 / 1e6, 4 * ((2 * R) // 4) ** 2 / 1e6, B, min(B, 16), 8 * W * H / 1e6),
            "parallelism": "viewpoint batch, %d panoramas per GPU per step, DEM replicated, no data-path collective" % B,
        },
        "e2e": {"value": e2e_value, "unit": "panoramas/s",
                "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 7 * W * H,
                "note": "horizonator_pan_zoom+move+set_zextents+render_offscreen into page-locked host buffers; "
                        "per-step inputs are 7 scalars passed as kernel arguments (no H2D copy), outputs 7*W*H "
                        "bytes D2H; one panorama per call, calls back to back",
                "pageable_host_buffers_value": e2e_pageable, "batch_call_value": e2e_batch,
                "batch_call_note": "horizonator_render_batch() (additive API): %d views per call into page-locked host "
                                   "memory, copies overlapping the next views' kernels" % B},
        "gpu_launches": K * B * stats["launches"],
        "roofline": {"bound": "hbm", "kernel": "whole panorama: 14 kernels replayed as one CUDA graph "
                               "(k_prepare, k_near, k_raster, k_big, 2 x (k_tiles, k_blocks, k_mesh, k_raster), k_big, k_resolve)",
                     "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg,
                     "stage_ms_single_panorama": {"%s [%s]" % (k, stages[k]): prof[k] for k in stages},
                     "latency_ms_single_panorama": latency_ms,
                     # the resource that actually binds in batch mode: instruction issue.  Instructions per panorama
                     # from the same ncu launch list as `traffic`; peak = SMs x 4 schedulers x SM clock
                     "issue": {"warp_instructions_per_panorama": WARP_INST_PER_PANORAMA,
                               "peak_warp_instructions_per_s": 148 * 4 * (clocks.get("sm_mhz") or 1965.0) * 1e6,
                               "frac": value / world * WARP_INST_PER_PANORAMA
                                       / (148 * 4 * (clocks.get("sm_mhz") or 1965.0) * 1e6),
                               "source": "smsp__inst_executed.sum over the 14 kernels, profiles/r01A_kernels_per_panorama.txt"},
                     "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum summed over the 14 kernels of one "
                                     "panorama (one CUDA-graph launch), from the ncu capture summarised in "
                                     "profiles/r01A_kernels_per_panorama.txt",
                     "note": "achieved = algorithmic bytes per panorama (one read of the int16 DEM square + one write of "
                             "image and range, SURVEY 8d) x measured panoramas/s over the whole timed region.  Hierarchical culling makes the kernels read far less DRAM than the "
                             "algorithmic figure (see traffic) and leaves them latency-bound, which is why several "
                             "panoramas are kept in flight"},
        "clocks": clocks,
        "aux": {"init_s": t_init, "mosaic_decode_ms": mosaic_ms,
                "mosaic_decode_gbs": 4 * (2 * R) ** 2 / (mosaic_ms / 1e3) / 1e9,
                "triangles_rasterised": stats["triangles_rasterised"], "big_entries": stats["big_entries"],
                "terrain_pixel_fraction": hit_fraction, "culling": counters,
                "host_enqueue_us_per_panorama": host_enqueue_s / (K * B) * 1e6},
    }
    if world == 1 and not args.no_cpu_baseline:
        port = cpu_baseline(tiles, use_ref=False, steps=3)
        real = llvmpipe_baseline()
        out["cpu_baseline"] = real or port
        if real:
            out["cpu_baseline_port"] = port
    quiet.restore()
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------ CPU legs

def cpu_baseline(tiles, use_ref, steps=1, warmup=0, budget_s=25.0):
    """Times the CPU oracle (or the reference build) on the C2 workload with all host threads.
    Bounded: renders are timed until `steps` are done or the budget is spent."""
    from oracle import binding
    cores = os.cpu_count() or 1
    cls = binding.Reference if use_ref else binding.Oracle
    o = cls(C2["lat"], C2["lon"], C2["W"], C2["H"], SRTM1=True, dir_dems=tiles,
            render_radius_m=C2["radius_m"], threads=cores)
    for _ in range(warmup):
        o.render(C2["az0"], C2["az1"], znear=C2["znear"], zfar=C2["zfar"])
    times = []
    t_start = time.perf_counter()
    for _ in range(steps):
        t0 = time.perf_counter()
        o.render(C2["az0"], C2["az1"], znear=C2["znear"], zfar=C2["zfar"])
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_start > budget_s:
            break
    mean = sum(times) / len(times)
    return {"value": 1.0 / mean, "unit": "panoramas/s", "cores": cores,
            "kind": "reference" if use_ref else "port",
            "sample": "%d full C2 panorama(s) (3600x600, 274 M triangles), steady state (DEM loaded), %.2f s each; "
                      "%s, OpenMP over %d threads" %
                      (len(times), mean,
                       "reference horizonator-lib.c+dem.c unmodified on the software-GL restatement (oracle/_ref)"
                       if use_ref else "CPU restatement of the reference GL path (oracle/)", cores),
            "ms_per_render": mean * 1e3}


def _timed_renders(o, steps, warmup, budget_s):
    """warm-up renders (at most a quarter of the budget), then up to `steps` timed ones until the budget is spent"""
    t_begin = time.perf_counter()
    done_w = 0
    for _ in range(warmup):
        o.render(C2["az0"], C2["az1"], znear=C2["znear"], zfar=C2["zfar"])
        done_w += 1
        if time.perf_counter() - t_begin > budget_s / 4:
            break
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        o.render(C2["az0"], C2["az1"], znear=C2["znear"], zfar=C2["zfar"])
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_begin > budget_s:
            break
    return times, done_w


def llvmpipe_child(args):
    """Child process of the reference arm: the reference's horizonator-lib.c + dem.c, unmodified, on Mesa llvmpipe
    (oracle/_ref/libhorizonator_mesa.so).  Own process because a GL driver owns process-wide state (JIT, thread pool)
    and because the parent must survive whatever the driver does.  Prints one JSON line."""
    from oracle import binding
    cores = os.cpu_count() or 1
    t0 = time.perf_counter()
    o = binding.MesaReference(C2["lat"], C2["lon"], C2["W"], C2["H"], SRTM1=True, dir_dems=TILES_DIR,
                              render_radius_m=C2["radius_m"], threads=cores)
    init_s = time.perf_counter() - t0
    version, renderer = o.gl_strings()
    times, done_w = _timed_renders(o, args.steps, args.warmup, float(os.environ.get("HZ_REF_BUDGET_S", "120")))
    print("LLVMPIPE " + json.dumps({"times": times, "warmup": done_w, "gl_version": version, "gl_renderer": renderer,
                                    "init_s": init_s, "lp_threads": min(cores, 16)}), flush=True)


def _llvmpipe_run(steps, warmup, budget_s):
    """The llvmpipe child process; returns its report or None (stderr says why)."""
    import subprocess
    from oracle import binding
    if not binding.have_mesa() or os.environ.get("HZ_REF_GL", "llvmpipe") != "llvmpipe":
        return None
    try:
        p = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--llvmpipe-child",
                            "--steps", str(steps), "--warmup", str(warmup)],
                           env=dict(os.environ, HZ_REF_BUDGET_S=str(budget_s)),
                           capture_output=True, text=True, timeout=2.5 * budget_s + 300)
        lines = [l for l in p.stdout.splitlines() if l.startswith("LLVMPIPE ")]
        if p.returncode == 0 and lines:
            return json.loads(lines[-1][len("LLVMPIPE "):])
        sys.stderr.write("bench: llvmpipe child failed (rc %d): %s\n" % (p.returncode, p.stderr[-1000:]))
    except Exception as e:      # timeout, missing interpreter, ...
        sys.stderr.write("bench: llvmpipe child failed: %r\n" % (e,))
    return None


def llvmpipe_baseline():
    """cpu_baseline of the B200 arm: ONE full benchmark panorama by the unmodified reference on Mesa llvmpipe on this
    box's host cores (~30 s of CPU work), or None where that build is absent."""
    r = _llvmpipe_run(steps=1, warmup=0, budget_s=1.0)
    if r is None:
        return None
    cores = os.cpu_count() or 1
    t = r["times"][0]
    return {"value": 1.0 / t, "unit": "panoramas/s", "cores": cores, "kind": "reference",
            "sample": "1 full C2 panorama (3600x600, 274 M triangles), first frame after init, %.1f s; the reference's "
                      "horizonator-lib.c + dem.c + GLSL shaders compiled UNMODIFIED on a real OpenGL driver: %s, %s "
                      "(oracle/_ref/libhorizonator_mesa.so); llvmpipe rasterises on %d threads, its vertex and geometry "
                      "stages run on one" % (t, r["gl_renderer"], r["gl_version"], r["lp_threads"]),
            "ms_per_render": t * 1e3}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    from tools import synth
    from oracle import binding
    synth.config2_tiles(TILES_DIR)
    if not os.path.exists(binding.ORACLE_SO):
        binding.build(ref=False)
    cores = os.cpu_count() or 1
    budget = float(os.environ.get("HZ_REF_BUDGET_S", "120"))

    # (1) the real thing (north_star: "the reference llvmpipe render timed on the box's host cores"): the unmodified
    #     reference on Mesa llvmpipe, in a child process
    llvmpipe = _llvmpipe_run(args.steps, args.warmup, budget)

    # (2) the same reference sources on the software-GL restatement (oracle/_ref), or the oracle port where the
    #     reference was never compiled: the whole arm when (1) is unavailable, a short second opinion when it is not
    use_ref = binding.have_ref()
    cls = binding.Reference if use_ref else binding.Oracle
    o = cls(C2["lat"], C2["lon"], C2["W"], C2["H"], SRTM1=True, dir_dems=TILES_DIR,
            render_radius_m=C2["radius_m"], threads=cores)
    if llvmpipe is None:
        times_r, done_w_r = _timed_renders(o, args.steps, args.warmup, budget)
    else:
        times_r, done_w_r = _timed_renders(o, min(args.steps, 3), 1, 20.0)
    mean_r = sum(times_r) / len(times_r)
    what_r = ("the reference's horizonator-lib.c + dem.c compiled unmodified (oracle/_ref) on a software-GL "
              "restatement of the driver (oracle/gl_pipeline.c)" if use_ref
              else "CPU restatement of the reference GL path (oracle/)")
    restated = {"value": 1.0 / mean_r, "unit": "panoramas/s", "cores": cores, "kind": "reference" if use_ref else "port",
                "sample": "%d full C2 panorama(s), %.2f s each; %s; OpenMP over %d host threads" %
                          (len(times_r), mean_r, what_r, cores)}

    if llvmpipe is not None:
        times, done_w = llvmpipe["times"], llvmpipe["warmup"]
        mean = sum(times) / len(times)
        kind = "reference"
        sample = ("%d of the requested %d steps timed (budget %.0f s), each one full C2 panorama (3600x600, R=5858, "
                  "274 M triangles) through horizonator_render_offscreen() into host buffers, %.1f s each; the "
                  "reference's horizonator-lib.c + dem.c compiled UNMODIFIED, running its own GLSL shaders on a real "
                  "OpenGL driver: %s, %s (oracle/_ref/libhorizonator_mesa.so; context on GLX pbuffers, no X server); "
                  "llvmpipe rasterises on %d threads, its vertex and geometry stages run on one; %d host cores; "
                  "1 process regardless of --gpus" %
                  (len(times), args.steps, budget, mean, llvmpipe["gl_renderer"], llvmpipe["gl_version"],
                   llvmpipe["lp_threads"], cores))
    else:
        times, done_w, mean, kind = times_r, done_w_r, mean_r, restated["kind"]
        sample = ("%d of the requested %d steps timed (budget %.0f s), each one full C2 panorama (3600x600, R=5858, "
                  "274 M triangles) into host buffers, %.2f s each; %s (Mesa llvmpipe build not available here); "
                  "OpenMP over %d host threads; 1 process regardless of --gpus" %
                  (len(times), args.steps, budget, mean, what_r, cores))
    value = 1.0 / mean
    out = {
        "impl": "reference",
        "metric": "panoramas/sec (SRTM1, 3600x600 px)",
        "value": value, "unit": "panoramas/s",
        "n_gpus": world, "steps": len(times), "warmup": done_w,
        "ms_per_step": mean * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "BASELINE configs[1]: single viewpoint, synthetic SRTM1 4x4 tiles, 150 km radius "
                               "(R=5858 cells, 274482450 triangles), 3600x600 panorama + range image",
                   "az_deg": [C2["az0"], C2["az1"]], "znear_m": C2["znear"], "zfar_m": C2["zfar"]},
        "cpu_baseline": {"value": value, "unit": "panoramas/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "panoramas/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if llvmpipe is not None:
        out["gl"] = {"version": llvmpipe["gl_version"], "renderer": llvmpipe["gl_renderer"], "init_s": llvmpipe["init_s"]}
        out["restated_gl"] = restated     # the same sources on the oracle's GL restatement: a much faster CPU rasteriser
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--batch", type=int, default=16, help="panoramas per step (rendered concurrently)")
    ap.add_argument("--llvmpipe-child", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.llvmpipe_child:
        return llvmpipe_child(args)
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
