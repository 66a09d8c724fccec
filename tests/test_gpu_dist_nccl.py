"""2-GPU NCCL run of the broadcast join (skipped on single-GPU boxes): same helper as the gloo test, but the
per-rank compute is the CUDA engine and the collectives run over NVLink."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from geopolars_b200 import GeoArrowArray, dist as gd, synth
    from geopolars_b200 import engine as E

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        ctx = E.Context(rank)
        n = 400_003
        lo, hi = gd.shard_rows(n, world, rank)
        pts = synth.uniform_points(hi - lo, first=lo, scale=100.0)
        polys = None
        if rank == 0:
            xy, ro, go = synth.star_polygons(100, 10)
            polys = GeoArrowArray.polygons(xy, ro, go)
        first, total, gathered = gd.contains_join_sharded(pts, polys, lambda p, x: E.PipIndex(ctx.upload(p)).query(x), src=0, gather_to=0)
        q.put({"rank": rank, "total": total, "gathered": gathered if rank == 0 else None})
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_broadcast_join_two_gpus_nccl(og):
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from geopolars_b200 import synth

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 200)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = {r["rank"]: r for r in (q.get(timeout=240) for _ in procs)}
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    n = 400_003
    xy, ro, go = synth.star_polygons(100, 10)
    want, _ = og.contains_join(og.OGArray(og.POLYGON, xy, geom_off=go, ring_off=ro), synth.uniform_points(n, scale=100.0), use_grid=True, threads=0)
    assert np.array_equal(res[0]["gathered"], want)
    assert np.array_equal(res[0]["total"], np.bincount(want[want >= 0], minlength=100))
    assert np.array_equal(res[1]["total"], res[0]["total"])
