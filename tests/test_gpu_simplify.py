"""GeoSeries::simplify (geopolars/geopolars-geo/src/geoseries.rs:108-116) on the GPU against the oracle's
recursive restatement of geo 0.27's Ramer-Douglas-Peucker: the retained coordinates are bit-identical, in
order, including exact ties and the minimum-size guard (decisions depend on the depth-first traversal order)."""
import numpy as np
import pytest

from geopolars_b200 import GeoArrowArray, GeometryType, engine, synth
from geopolars_b200._lib import MismatchedGeometry

pytestmark = pytest.mark.gpu


def check(ctx, og, conv, arr: GeoArrowArray, eps: float):
    keep = og.simplify_mask(conv(arr), eps, threads=0)
    got = engine.simplify(ctx.upload(arr), eps).to_host()
    assert int(got.type) == int(arr.type)
    assert np.array_equal(got.xy, arr.xy[keep])
    inner = "geom_off" if arr.type == GeometryType.LINESTRING else "ring_off"
    pos = np.concatenate([[0], np.cumsum(keep)])
    assert np.array_equal(getattr(got, inner), pos[getattr(arr, inner)])
    for name in ("geom_off", "part_off"):
        if name != inner and getattr(arr, name) is not None:
            assert np.array_equal(getattr(got, name), getattr(arr, name))
    return keep


def walks(rng, n, kmax, lattice=False):
    lens = rng.integers(0, kmax, n)
    off = np.concatenate([[0], np.cumsum(lens)])
    steps = rng.integers(-2, 3, (off[-1], 2)).astype(float) if lattice else rng.normal(0, 1, (off[-1], 2))
    xy = np.cumsum(steps, axis=0)
    return xy, off


@pytest.mark.parametrize("eps", [0.05, 0.7, 3.0, 50.0])
def test_linestrings(ctx, og, conv, eps):
    rng = np.random.default_rng(int(eps * 100))
    xy, off = walks(rng, 4000, 300)
    keep = check(ctx, og, conv, GeoArrowArray.linestrings(xy, off), eps)
    assert 0 < keep.sum() < len(keep)


def test_ties_collinear_runs_and_duplicates(ctx, og, conv):
    rng = np.random.default_rng(5)
    xy, off = walks(rng, 3000, 120, lattice=True)  # integer steps: equal distances, zero-length segments, collinear runs
    for eps in (0.5, 1.0, 2.0):
        check(ctx, og, conv, GeoArrowArray.linestrings(xy, off), eps)
    check(ctx, og, conv, GeoArrowArray.linestrings(xy, off), 0.0)  # epsilon <= 0: identity
    check(ctx, og, conv, GeoArrowArray.linestrings(xy, off), -1.0)


def test_polygons_keep_at_least_four_coordinates(ctx, og, conv):
    xy, ro, go = synth.star_polygons(2500, 50)
    arr = GeoArrowArray.polygons(xy, ro, go)
    for eps in (0.01, 0.5, 2.0, 100.0):
        keep = check(ctx, og, conv, arr, eps)
    sizes = np.diff(np.concatenate([[0], np.cumsum(keep)])[ro])
    assert sizes.min() >= 4  # eps = 100 collapses every ring as far as the guard allows
    # small rings around the guard: 4..7 coordinates, tiny deviations
    rng = np.random.default_rng(2)
    shapes_ = []
    for i in range(2000):
        k = int(rng.integers(3, 7))
        th = np.sort(rng.uniform(0, 2 * np.pi, k))
        ring = [(round(np.cos(t), 2), round(0.02 * np.sin(t), 3)) for t in th]
        shapes_.append([ring + [ring[0]]] + ([[(5, 5), (6, 5), (6, 6), (5, 5)]] if i % 5 == 0 else []))
    small = GeoArrowArray.from_shapes(GeometryType.POLYGON, shapes_)
    for eps in (0.01, 0.05, 0.5):
        check(ctx, og, conv, small, eps)


def test_multi_types_share_outer_offsets(ctx, og, conv):
    rng = np.random.default_rng(8)
    ml = []
    for i in range(800):
        ml.append([[tuple(p) for p in np.cumsum(rng.normal(0, 1, (int(rng.integers(0, 40)), 2)), 0)] for _ in range(int(rng.integers(0, 4)))])
    ml[3] = None
    check(ctx, og, conv, GeoArrowArray.from_shapes(GeometryType.MULTILINESTRING, ml), 0.8)
    mp = []
    for i in range(500):
        polys = []
        for _ in range(int(rng.integers(0, 3))):
            k = int(rng.integers(3, 30))
            th = np.sort(rng.uniform(0, 2 * np.pi, k))
            r = rng.uniform(0.5, 1.0, k)
            ring = [(float(r[j] * np.cos(th[j])), float(r[j] * np.sin(th[j]))) for j in range(k)]
            polys.append([ring + [ring[0]]])
        mp.append(polys)
    check(ctx, og, conv, GeoArrowArray.from_shapes(GeometryType.MULTIPOLYGON, mp), 0.15)


def test_simplify_rejects_points(ctx):
    with pytest.raises(MismatchedGeometry):
        engine.simplify(ctx.upload(GeoArrowArray.points(np.zeros((3, 2)))), 1.0)
