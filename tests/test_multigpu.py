"""NCCL path on real GPUs (skipped with fewer than 2): azimuth-wedge render gathered over NCCL == the unsharded
render, bit for bit; sharded viewpoint batch == the same views rendered by one GPU."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

LAT, LON = 35.0 + 1.0 / 2400.0, -117.0 + 1.0 / 2400.0


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, tiles, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    os.environ["HORIZONATOR_DEVICE"] = str(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import horizonator_b200 as hz
        from horizonator_b200 import sharding
        W, H, R = 1202, 160, 300
        h = hz.horizonator(LAT, LON, W, H, dir_dems=tiles, render_radius_cells=R)
        h.set_zextents(100., 100000.)
        h.pan_zoom(-180.05, 179.95)
        full_i, full_r = h.render(-180.05, 179.95, zfar=100000.)
        gi, gr = sharding.render_wedges(h)
        torch.cuda.synchronize()
        ok = np.array_equal(gi.cpu().numpy(), full_i) and np.array_equal(gr.cpu().numpy(), full_r)

        # the same panorama assembled by the resolve kernels themselves in peer memory (no gather collective)
        pp = sharding.PeerPanorama(h)
        for _ in range(2):
            pi, pr = pp.render()
            torch.cuda.synchronize()
            ok = ok and np.array_equal(pi.cpu().numpy(), full_i) and np.array_equal(pr.cpu().numpy(), full_r)
        # ... and only into rank 0's buffers
        pp.image.zero_(); pp.ranges.zero_()
        torch.cuda.synchronize()
        pi, pr = pp.render(root=0)
        torch.cuda.synchronize()
        if rank == 0:
            ok = ok and np.array_equal(pi.cpu().numpy(), full_i) and np.array_equal(pr.cpu().numpy(), full_r)
        else:
            ok = ok and not pi.any().item()
        ok = ok and pp.timeouts() == 0
        dist.barrier()
        pp.close()

        # ... and delivered to ONE host buffer that all ranks share, each rank copying its own wedge (its own PCIe link)
        hp = sharding.HostPanorama(h)
        for _ in range(2):
            hi_, hr_ = hp.render()
            ok = ok and np.array_equal(hi_, full_i) and np.array_equal(hr_, full_r)
        hp.close()
        # strict mode of the peer barrier: a time-out would raise instead of returning a stale panorama
        pp2 = sharding.PeerPanorama(h)
        pi, pr = pp2.render(strict=True)
        ok = ok and np.array_equal(pi.cpu().numpy(), full_i)
        dist.barrier()
        pp2.close()
        views = [(LAT + 0.01 * k, LON - 0.008 * k, -180.05, 179.95) for k in range(5)]
        img, rng, (lo, hi), prof = sharding.render_batch_sharded(h, views)
        torch.cuda.synchronize()
        for k in range(lo, hi):
            i1, r1 = h.render(-180.05, 179.95, lat=views[k][0], lon=views[k][1], zfar=100000.)
            ok = ok and np.array_equal(img[k - lo].cpu().numpy(), i1) and np.array_equal(rng[k - lo].cpu().numpy(), r1)
            rows, pr = sharding.horizon_profile(torch.from_numpy(r1))
            ok = ok and torch.equal(prof[0][k].cpu(), rows) and torch.equal(prof[1][k].cpu(), pr)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_nccl_wedges_and_batch(tiles_c1):
    import torch
    import torch.multiprocessing as mp
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = min(n, 4)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, tiles_c1, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    got = sorted(q.get(timeout=5) for _ in range(world))
    assert got == [(r, True) for r in range(world)]
