"""world_size-2 gloo tests of the multi-GPU plumbing (geopolars_b200/dist.py) on CPU: row-range
sharding, polygon broadcast, count all-reduce, ragged gather.  The per-rank compute is injected
(here: the oracle) so the collective logic runs without a device."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist

    from geopolars_b200 import GeoArrowArray, dist as gd, synth
    from oracle import oracle as og

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = 20_001  # odd on purpose: ranks get different row counts
        lo, hi = gd.shard_rows(n, world, rank)
        pts = synth.uniform_points(hi - lo, first=lo, scale=100.0)
        polys = None
        if rank == 0:
            xy, ro, go = synth.star_polygons(100, 10)
            polys = GeoArrowArray.polygons(xy, ro, go)

        def local_join(p, x):
            o = og.OGArray(int(p.type), p.xy, geom_off=p.geom_off, ring_off=p.ring_off)
            return og.contains_join(o, x, use_grid=True)[0]

        first, total, gathered = gd.contains_join_sharded(pts, polys, local_join, src=0, gather_to=0)
        out = {"rank": rank, "n_local": len(first), "total": total}
        if rank == 0:
            out["gathered"] = gathered
        q.put(out)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_broadcast_join_world2_gloo(og):
    from geopolars_b200 import GeoArrowArray, synth

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 200)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=100) for _ in procs]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    res = {r["rank"]: r for r in res}
    n = 20_001
    assert res[0]["n_local"] + res[1]["n_local"] == n and res[0]["n_local"] == 10_001
    xy, ro, go = synth.star_polygons(100, 10)
    want, _ = og.contains_join(og.OGArray(og.POLYGON, xy, geom_off=go, ring_off=ro), synth.uniform_points(n, scale=100.0), use_grid=True)
    assert np.array_equal(res[0]["gathered"], want)  # rank order == row order
    want_counts = np.bincount(want[want >= 0], minlength=100)
    assert np.array_equal(res[0]["total"], want_counts) and np.array_equal(res[1]["total"], want_counts)


def test_shard_helpers():
    from geopolars_b200 import GeoArrowArray, GeometryType, dist as gd

    assert [gd.shard_rows(10, 4, r) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert gd.shard_rows(0, 2, 1) == (0, 0)
    # coordinate-balanced boundaries: one huge ring and many small ones
    rings = [[(0, 0)] * 1000] + [[(0, 0)] * 10 for _ in range(100)]
    arr = GeoArrowArray.from_shapes(GeometryType.POLYGON, [[r] for r in rings])
    b = gd.shard_rows_by_coords(arr, 2)
    assert b[0] == 0 and b[-1] == 101 and b[1] in (1, 2)  # the big row alone already holds half the coordinates
    parts = [arr.take_rows(b[i], b[i + 1]) for i in range(2)]
    assert sum(p.n_coords for p in parts) == arr.n_coords and all(p.ring_off[0] == 0 for p in parts)
