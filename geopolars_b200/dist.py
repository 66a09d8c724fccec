"""Multi-GPU plumbing for the GeoSeries hot path: one process per GPU, torch.distributed (NCCL on the
B200s, gloo in CPU tests).  SURVEY.md §8e:

  * every op is independent per row, so unary / row-wise ops shard by contiguous ROW RANGES balanced by
    coordinate count — no data-path collective at all;
  * the points-in-polygons join is a broadcast join: the big side (points) is partitioned by row range
    and never moves, the small side (polygons, ~10 MB) is BROADCAST once from the rank that holds it;
    results stay sharded like the input (a row-partitioned column), only the per-polygon hit counts are
    all-reduced, and `gather_rows` collects a sharded result on one rank when the caller wants it there
    (config 4: "NCCL gather").

The reference has no distributed backend: this module has no reference counterpart beyond
the join semantics of geopolars/src/spatial_index.rs:37-204.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import numpy as np

from .geoarrow import GeoArrowArray, GeometryType


def shard_rows(n: int, world: int, rank: int) -> Tuple[int, int]:
    """contiguous balanced row range [lo, hi) of rank `rank` (first n % world ranks get one extra row)"""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_rows_by_coords(arr: GeoArrowArray, world: int) -> List[int]:
    """row boundaries [b_0=0, ..., b_world=n] such that every rank gets ~ the same number of COORDINATES
    (the unit of HBM traffic for area/centroid/affine/hull), found on the prefix of the offsets."""
    n = len(arr)
    if arr.type == GeometryType.POINT:
        return [shard_rows(n, world, r)[0] for r in range(world)] + [n]
    g = arr.geom_off
    if arr.type in (GeometryType.LINESTRING, GeometryType.MULTIPOINT):
        cum = g
    elif arr.type in (GeometryType.POLYGON, GeometryType.MULTILINESTRING):
        cum = arr.ring_off[g]
    else:
        cum = arr.ring_off[arr.part_off[g]]
    total = int(cum[-1])
    bounds = [0]
    for r in range(1, world):
        bounds.append(int(np.searchsorted(cum, total * r / world, side="left")))
    bounds.append(n)
    return [min(max(b, bounds[i - 1] if i else 0), n) for i, b in enumerate(bounds)]


def _dist():
    import torch.distributed as dist

    return dist


def _device_for_backend():
    import torch

    dist = _dist()
    if dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def broadcast_geoarrow(arr: Optional[GeoArrowArray], src: int = 0) -> GeoArrowArray:
    """Broadcast the small side of a join from `src` to every rank: one header broadcast (type + sizes),
    then one broadcast per buffer (coords, offsets, validity).  10.4 MB for BASELINE config 2's polygons:
    latency bound over NVLink/NVSwitch."""
    import torch

    dist = _dist()
    dev = _device_for_backend()
    rank = dist.get_rank()
    hdr = torch.zeros(8, dtype=torch.int64, device=dev)
    if rank == src:
        assert arr is not None
        hdr[0] = int(arr.type)
        hdr[1] = arr.n_coords
        hdr[2] = 0 if arr.geom_off is None else len(arr.geom_off)
        hdr[3] = 0 if arr.part_off is None else len(arr.part_off)
        hdr[4] = 0 if arr.ring_off is None else len(arr.ring_off)
        hdr[5] = 0 if arr.valid is None else len(arr.valid)
    dist.broadcast(hdr, src=src)
    t, nc, ng, npart, nr, nv = [int(v) for v in hdr[:6].tolist()]

    def bc(a: Optional[np.ndarray], n: int, dtype) -> Optional[np.ndarray]:
        if n == 0:
            return None
        ten = torch.empty(n, dtype=dtype, device=dev)
        if rank == src:
            ten.copy_(torch.from_numpy(np.ascontiguousarray(a).reshape(-1)))
        dist.broadcast(ten, src=src)
        return ten.cpu().numpy()

    xy = bc(arr.xy if rank == src else None, nc * 2, torch.float64)
    xy = np.zeros((0, 2)) if xy is None else xy.reshape(-1, 2)
    geom = bc(arr.geom_off if rank == src else None, ng, torch.int64)
    part = bc(arr.part_off if rank == src else None, npart, torch.int64)
    ring = bc(arr.ring_off if rank == src else None, nr, torch.int64)
    valid = bc(arr.valid.astype(np.uint8) if rank == src and arr.valid is not None else None, nv, torch.uint8)
    return GeoArrowArray(GeometryType(t), xy, geom_off=geom, part_off=part, ring_off=ring, valid=None if valid is None else valid.astype(bool))


def allreduce_counts(counts: np.ndarray) -> np.ndarray:
    """sum per-polygon hit counts over ranks (config 4: M x 8 B all-reduce)"""
    import torch

    dist = _dist()
    t = torch.from_numpy(np.ascontiguousarray(counts, dtype=np.int64)).to(_device_for_backend())
    dist.all_reduce(t)
    return t.cpu().numpy()


def gather_rows(local: np.ndarray, dst: int = 0) -> Optional[np.ndarray]:
    """collect a row-partitioned result column on rank `dst` in rank order (ranks may hold different
    numbers of rows): all-gather of the counts, then an all-gather of padded buffers trimmed on `dst`."""
    import torch

    dist = _dist()
    dev = _device_for_backend()
    world, rank = dist.get_world_size(), dist.get_rank()
    n = torch.tensor([local.shape[0]], dtype=torch.int64, device=dev)
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    m = max(sizes) if sizes else 0
    buf = torch.zeros((m,) + local.shape[1:], dtype=torch.from_numpy(local[:0]).dtype, device=dev)
    buf[: local.shape[0]] = torch.from_numpy(np.ascontiguousarray(local)).to(dev)
    outs = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(outs, buf)
    if rank != dst:
        return None
    return np.concatenate([o[:s].cpu().numpy() for o, s in zip(outs, sizes)], axis=0)


def contains_join_sharded(points_local: np.ndarray, polygons: Optional[GeoArrowArray], local_join: Callable, src: int = 0,
                          gather_to: Optional[int] = None):
    """Broadcast join over the process group.

    points_local : this rank's row range of the point column, (n_local, 2) f64
    polygons     : the polygon column on rank `src` (None elsewhere)
    local_join   : (polygons: GeoArrowArray, points: ndarray) -> first_id int32[n_local]; on a GPU rank this is
                   `lambda polys, pts: PipIndex(ctx.upload(polys)).query(pts)` (see bench.py); CPU tests inject
                   the oracle so the collective logic is exercised without a device.
    returns (first_id_local, global per-polygon hit counts, gathered ids on `gather_to` or None)
    """
    polys = broadcast_geoarrow(polygons, src=src)
    first = np.asarray(local_join(polys, points_local), dtype=np.int32)
    counts = np.bincount(first[first >= 0], minlength=len(polys)).astype(np.int64)
    total = allreduce_counts(counts)
    gathered = gather_rows(first, dst=gather_to) if gather_to is not None else None
    return first, total, gathered
