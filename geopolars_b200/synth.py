"""Synthetic workloads for BASELINE.json's configs, host side (numpy).

Counter-based splitmix64 so that CPU oracle, numpy and the CUDA generator kernels
(csrc/synth.cu) produce bit-identical coordinates from (stream, element) alone.
Anything involving cos/sin (polygon vertices) is generated HERE on the host and uploaded, because
libm and CUDA sincos differ in the last ulp; uniform points and random walks are pure
integer/IEEE mul-add and are regenerated bit-exactly on the device.
"""
from __future__ import annotations

import numpy as np

_U64 = np.uint64
SEED_BASE = 0xB2000000


def splitmix_u(stream: int, counter: np.ndarray) -> np.ndarray:
    """u in [0,1): z=seed+G*(i+1); two xorshift-multiply rounds; top 53 bits."""
    c = np.asarray(counter).astype(np.uint64)
    with np.errstate(over="ignore"):
        z = _U64(SEED_BASE + stream) + _U64(0x9E3779B97F4A7C15) * (c + _U64(1))
        z = (z ^ (z >> _U64(30))) * _U64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> _U64(27))) * _U64(0x94D049BB133111EB)
        z = z ^ (z >> _U64(31))
    return (z >> _U64(11)).astype(np.float64) * 2.0**-53


def uniform_points(n: int, first: int = 0, stream: int = 2, scale: float = 1000.0) -> np.ndarray:
    """config 2/4 points: x = scale*u(i*4+0), y = scale*u(i*4+1)."""
    i = np.arange(first, first + n, dtype=np.uint64)
    xy = np.empty((n, 2), dtype=np.float64)
    xy[:, 0] = scale * splitmix_u(stream, i * _U64(4))
    xy[:, 1] = scale * splitmix_u(stream, i * _U64(4) + _U64(1))
    return xy


def star_polygons(m: int = 10_000, grid: int = 100, cell: float = 10.0, nvert: int = 64, stream: int = 1):
    """config 2/4 polygons: polygon j sits at the centre of cell j of a grid x grid lattice,
    nvert distinct vertices at theta_k = 2*pi*k/nvert, radius (0.20 + 0.28*u(j*nvert+k))*cell,
    CCW, explicitly closed -> nvert+1 coords.  Returns (xy (m*(nvert+1),2), ring_off, geom_off)."""
    j = np.arange(m, dtype=np.uint64)
    k = np.arange(nvert, dtype=np.uint64)
    u = splitmix_u(stream, j[:, None] * _U64(nvert) + k[None, :])
    r = (0.20 + 0.28 * u) * cell
    th = 2.0 * np.pi * k.astype(np.float64) / nvert
    cx = ((j % _U64(grid)).astype(np.float64) + 0.5) * cell
    cy = ((j // _U64(grid)).astype(np.float64) + 0.5) * cell
    x = cx[:, None] + r * np.cos(th)[None, :]
    y = cy[:, None] + r * np.sin(th)[None, :]
    xy = np.empty((m, nvert + 1, 2), dtype=np.float64)
    xy[:, :nvert, 0] = x
    xy[:, :nvert, 1] = y
    xy[:, nvert, :] = xy[:, 0, :]
    ring_off = np.arange(m + 1, dtype=np.int64) * (nvert + 1)
    geom_off = np.arange(m + 1, dtype=np.int64)
    return xy.reshape(-1, 2), ring_off, geom_off


def walk_linestrings(n: int, k: int = 16, first: int = 0, stream: int = 3, other_of: int | None = None):
    """config 3 linestrings: random walk of k coords from (1000u,1000u), steps uniform in [-1,1]^2.
    With other_of=s the start is instead the start of stream s's walk shifted by (4u-2, 4u-2), so that
    about half of the (A_i, B_i) pairs intersect.  Counter layout: element i uses i*2k + c."""
    i = np.arange(first, first + n, dtype=np.uint64)
    per = _U64(2 * k)
    base = i * per
    if other_of is None:
        sx = 1000.0 * splitmix_u(stream, base)
        sy = 1000.0 * splitmix_u(stream, base + _U64(1))
    else:
        ax = 1000.0 * splitmix_u(other_of, base)
        ay = 1000.0 * splitmix_u(other_of, base + _U64(1))
        sx = ax + (4.0 * splitmix_u(stream, base) - 2.0)
        sy = ay + (4.0 * splitmix_u(stream, base + _U64(1)) - 2.0)
    xy = np.empty((n, k, 2), dtype=np.float64)
    xy[:, 0, 0] = sx
    xy[:, 0, 1] = sy
    for s in range(1, k):  # sequential prefix sum: same rounding order as the device generator
        dx = 2.0 * splitmix_u(stream, base + _U64(2 * s)) - 1.0
        dy = 2.0 * splitmix_u(stream, base + _U64(2 * s + 1)) - 1.0
        xy[:, s, 0] = xy[:, s - 1, 0] + dx
        xy[:, s, 1] = xy[:, s - 1, 1] + dy
    geom_off = np.arange(n + 1, dtype=np.int64) * k
    return xy.reshape(-1, 2), geom_off


def blob_polygons(n: int, nvert: int = 256, first: int = 0, stream: int = 5):
    """config 5 polygons: centre (1000u,1000u), nvert vertices, r_k = 0.5+0.5u, closed (nvert+1 coords).
    Counter layout: element i uses i*(nvert+2) + c (c=0,1 centre; 2+k radius)."""
    i = np.arange(first, first + n, dtype=np.uint64)
    per = _U64(nvert + 2)
    base = i * per
    cx = 1000.0 * splitmix_u(stream, base)
    cy = 1000.0 * splitmix_u(stream, base + _U64(1))
    k = np.arange(nvert, dtype=np.uint64)
    r = 0.5 + 0.5 * splitmix_u(stream, base[:, None] + _U64(2) + k[None, :])
    th = 2.0 * np.pi * k.astype(np.float64) / nvert
    xy = np.empty((n, nvert + 1, 2), dtype=np.float64)
    xy[:, :nvert, 0] = cx[:, None] + r * np.cos(th)[None, :]
    xy[:, :nvert, 1] = cy[:, None] + r * np.sin(th)[None, :]
    xy[:, nvert, :] = xy[:, 0, :]
    ring_off = np.arange(n + 1, dtype=np.int64) * (nvert + 1)
    geom_off = np.arange(n + 1, dtype=np.int64)
    return xy.reshape(-1, 2), ring_off, geom_off
