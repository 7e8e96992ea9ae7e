"""Multi-process path on CPU: world_size-2 (and 3) gloo runs of exactly the tensor code that runs over NCCL on
the GPUs (horizonator_b200/sharding.py): wedge gather + placement, block partition + gather of reduced products,
the horizon-profile reduction.  The renders themselves are faked by slicing a known panorama -- the claim tested
is "sharded == unsharded", not the renderer."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from horizonator_b200 import sharding


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _panorama(H=40, W=301, seed=0):
    rs = np.random.default_rng(seed)
    img = torch.from_numpy(rs.integers(0, 256, (H, W, 3), dtype=np.uint8))
    rng = torch.from_numpy(rs.uniform(100, 1e5, (H, W)).astype(np.float32))
    rng[rs.uniform(size=(H, W)) < 0.6] = -1.0
    return img, rng


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ok = True
        # ---- wedges: widths differ by one column (301 columns over 2 or 3 ranks)
        img, rng = _panorama()
        edges = sharding.wedge_edges(img.shape[1], world)
        x0, x1 = edges[rank], edges[rank + 1]
        full_i = sharding.gather_wedges(img[:, x0:x1].contiguous(), edges)
        full_r = sharding.gather_wedges(rng[:, x0:x1].contiguous(), edges)
        ok &= bool(torch.equal(full_i, img) and torch.equal(full_r, rng))
        # ---- viewpoint batch: 7 views over the ranks, profiles gathered, images stay local
        n = 7
        batch = torch.stack([_panorama(seed=k)[1] for k in range(n)])
        lo, hi = sharding.block_partition(n, world, rank)
        rows, r = sharding.horizon_profile(batch[lo:hi])
        all_rows = sharding.gather_blocks(rows, n)
        all_r = sharding.gather_blocks(r, n)
        want_rows, want_r = sharding.horizon_profile(batch)
        ok &= bool(torch.equal(all_rows, want_rows) and torch.equal(all_r, want_r))
        # max-over-ranks timing as bench.py does it
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ok &= t.item() == world
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_sharded_equals_unsharded(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    got = sorted(q.get(timeout=5) for _ in range(world))
    assert got == [(r, True) for r in range(world)]


def test_partitions_cover_everything_once():
    for n in (0, 1, 7, 4096):
        for world in (1, 2, 3, 8):
            parts = [sharding.block_partition(n, world, r) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            assert max(h - l for l, h in parts) - min(h - l for l, h in parts) <= 1
    for W in (8, 3600, 36000, 301):
        for G in (1, 2, 3, 8):
            e = sharding.wedge_edges(W, G)
            assert e[0] == 0 and e[-1] == W and all(b > a for a, b in zip(e, e[1:]))
    with pytest.raises(ValueError):
        sharding.wedge_edges(4, 5)
    with pytest.raises(ValueError):
        sharding.block_partition(4, 2, 2)


def test_horizon_profile_known_answer():
    r = torch.full((5, 4), -1.0)
    r[3, 0] = 700.0; r[4, 0] = 300.0         # column 0: topmost terrain in row 3
    r[0, 2] = 9000.0                         # column 2: terrain in the top row
    rows, rng = sharding.horizon_profile(r)
    assert rows.tolist() == [3, -1, 0, -1]
    assert rng.tolist() == [700.0, -1.0, 9000.0, -1.0]
    rows2, rng2 = sharding.horizon_profile(torch.stack([r, r.flip(0)]))
    assert rows2.shape == (2, 4) and rows2[1].tolist() == [0, -1, 4, -1] and rng2[1, 0].item() == 300.0


def test_single_process_paths_need_no_process_group():
    img, rng = _panorama()
    e = sharding.wedge_edges(img.shape[1], 1)
    assert torch.equal(sharding.gather_wedges(img, e), img)
    assert torch.equal(sharding.gather_blocks(rng, rng.shape[0]), rng)
