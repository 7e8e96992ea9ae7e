#!/usr/bin/env python
"""BASELINE config 4: one ultra-high-resolution 360-degree panorama (36000 x 4000) partitioned by azimuth wedge over
the ranks (torchrun, one process per GPU), gathered with one NCCL all_gather per output; checked against the
unsharded render of rank 0 by checksum.  Prints one JSON line on rank 0."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

C2_LAT, C2_LON = 34.0 + 1.0 / 7200.0, -117.0 + 1.0 / 7200.0


def main():
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    os.environ["HORIZONATOR_DEVICE"] = str(local)
    from bench import StdoutToStderr
    quiet = StdoutToStderr()     # NCCL prints its version banner on descriptor 1
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import horizonator_b200 as hz
    from horizonator_b200 import sharding
    from tools import synth
    if rank == 0:
        synth.config2_tiles(os.environ.get("HZ_BENCH_TILES", "/tmp/hz_tiles_c2"))
    if world > 1:
        dist.barrier()
    tiles = os.environ.get("HZ_BENCH_TILES", "/tmp/hz_tiles_c2")
    W, H = int(os.environ.get("HZ_WEDGE_W", "36000")), int(os.environ.get("HZ_WEDGE_H", "4000"))
    h = hz.horizonator(C2_LAT, C2_LON, W, H, SRTM1=True, dir_dems=tiles, render_radius_m=150000.)
    h.set_zextents(100., 150000.)
    az0, az1 = -180.0 + 180.0 / W, 180.0 - 180.0 / W       # pixel centres on multiples of 360/W, not exactly 360 wide
    h.pan_zoom(az0, az1)
    h.move(C2_LAT, C2_LON)

    def sharded():
        return sharding.render_wedges(h)

    for _ in range(2):
        img, rng = sharded()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    n = 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        img, rng = sharded()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    sharded_ms = float(t.item())
    # the same with the exchange fused into the resolve kernel (peer-memory stores over NVLink)
    pp = sharding.PeerPanorama(h)
    for _ in range(2):
        pimg, prng = pp.render()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        pimg, prng = pp.render()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    peer_ms = float(t.item())
    peer_same = bool(torch.equal(pimg, img)) and bool(torch.equal(prng, rng))
    # ... and with rank 0 as the only destination
    for _ in range(2):
        pp.render(root=0)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        pp.render(root=0)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    peer_root_ms = float(t.item())
    if rank == 0:
        peer_same = peer_same and bool(torch.equal(pp.image, img)) and bool(torch.equal(pp.ranges, rng))
    # render-only part (no gather) of this rank's wedge
    edges = sharding.wedge_edges(W, world)
    x0, x1 = edges[rank], edges[rank + 1]
    di = torch.empty((H, x1 - x0, 3), dtype=torch.uint8, device="cuda"); dr = torch.empty((H, x1 - x0), dtype=torch.float32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    h.render_wedge_device(x0, x1, di.data_ptr(), dr.data_ptr(), st); torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        h.render_wedge_device(x0, x1, di.data_ptr(), dr.data_ptr(), st)
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    wedge_ms = float(t.item())
    ck = (int(img.sum(dtype=torch.int64).item()), float(rng.double().sum().item()))
    if rank == 0:
        # the unsharded render on this GPU
        fi = torch.empty((H, W, 3), dtype=torch.uint8, device="cuda"); fr = torch.empty((H, W), dtype=torch.float32, device="cuda")
        h.render_wedge_device(0, W, fi.data_ptr(), fr.data_ptr(), st); torch.cuda.synchronize()
        e0.record()
        for _ in range(3):
            h.render_wedge_device(0, W, fi.data_ptr(), fr.data_ptr(), st)
        e1.record(); torch.cuda.synchronize()
        whole_ms = e0.elapsed_time(e1) / 3
        same = bool(torch.equal(fi, img)) and bool(torch.equal(fr, rng))
        quiet.restore()
        print(json.dumps({"config": "BASELINE configs[3]: %dx%d full circle, C2 DEM (R=5858), azimuth wedges" % (W, H),
                          "n_gpus": world, "sharded_ms_per_panorama_incl_gather": sharded_ms, "wedge_render_ms_max_over_ranks": wedge_ms,
                          "peer_store_ms_per_panorama": peer_ms, "peer_store_to_rank0_only_ms": peer_root_ms,
                          "peer_store_equals_gathered": peer_same,
                          "unsharded_ms_one_gpu": whole_ms, "speedup_allgather": whole_ms / sharded_ms,
                          "speedup_peer_store": whole_ms / peer_ms, "gathered_equals_unsharded": same,
                          "gather_bytes_total": 7 * W * H, "checksum": ck}), flush=True)
    assert pp.timeouts() == 0
    pp.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
