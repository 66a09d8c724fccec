"""convex_hull parity: vertex order must be geo's quick_hull order (index output => bit-exact)."""
import numpy as np
import pytest

from geopolars_b200 import GeoArrowArray, GeometryType, synth

pytestmark = pytest.mark.gpu


def _check(ctx, og, conv, arr):
    from geopolars_b200 import engine as E

    want_off, want_xy = og.convex_hull(conv(arr), threads=0)
    got = E.convex_hull(ctx.upload(arr)).to_host()
    assert got.type == GeometryType.POLYGON
    assert np.array_equal(got.ring_off, want_off)
    assert np.array_equal(got.geom_off, np.arange(len(arr) + 1))
    assert np.array_equal(got.xy, want_xy)  # same vertices, same order, same bits
    return got


def test_blob_polygons_config5_shape(ctx, og, conv):
    xy, ro, go = synth.blob_polygons(3000, 256)
    got = _check(ctx, og, conv, GeoArrowArray.polygons(xy, ro, go))
    sizes = np.diff(got.ring_off)
    assert sizes.min() >= 4 and sizes.max() < 257
    # rings are closed and counter-clockwise
    first = got.xy[got.ring_off[:-1]]
    last = got.xy[got.ring_off[1:] - 1]
    assert np.array_equal(first, last)


def test_star_polygons_and_random_clouds(ctx, og, conv):
    xy, ro, go = synth.star_polygons(2000, 50)
    _check(ctx, og, conv, GeoArrowArray.polygons(xy, ro, go))
    rng = np.random.default_rng(7)
    sizes = rng.integers(1, 700, size=400)  # ragged, includes > 512 coords (several staging chunks)
    off = np.concatenate([[0], np.cumsum(sizes)])
    pts = rng.normal(size=(off[-1], 2)) * rng.uniform(0.1, 1e6, size=(off[-1], 1))
    _check(ctx, og, conv, GeoArrowArray(GeometryType.MULTIPOINT, pts, geom_off=off))
    _check(ctx, og, conv, GeoArrowArray.linestrings(pts, off))


def test_degenerate_hulls(ctx, og, conv):
    shapes = [
        [[(0, 0), (4, 0), (4, 4), (0, 4), (0, 0)]],  # rectangle
        [[(0, 0), (1, 1), (2, 2), (3, 3), (0, 0)]],  # all collinear
        [[(5, 5), (5, 5), (5, 5), (5, 5)]],  # identical points
        [[(0, 0), (1, 0), (0, 0)]],  # fewer than four coords
        [[(2, 2)]],
        [],  # empty polygon
        [[(0, 0), (2, 1), (4, 0), (5, 2), (4, 4), (2, 5), (0, 4), (-1, 2), (0, 0)]],  # SURVEY.md §8a worked example
        [[(0, 0), (3, 0), (3, 3), (1, 1), (0, 3), (0, 0)], [(1, 1), (2, 1), (2, 2), (1, 1)]],  # concave + hole (holes ignored)
    ]
    arr = GeoArrowArray.from_shapes(GeometryType.POLYGON, shapes)
    got = _check(ctx, og, conv, arr)
    # the documented order for the 9-point example: lower chain after min ..., max, upper chain ..., min, first again
    ex = GeoArrowArray.from_shapes(GeometryType.POLYGON, [[[(0, 0), (2, -1), (4, 0), (5, 2), (4, 4), (2, 5), (0, 4), (-1, 2), (0, 0)]]])
    g2 = _check(ctx, og, conv, ex)
    assert g2.xy.tolist() == [[0, 0], [2, -1], [4, 0], [5, 2], [4, 4], [2, 5], [0, 4], [-1, 2], [0, 0]]


def test_multipolygon_points_and_nulls(ctx, og, conv):
    sq = lambda x, y, s: [(x, y), (x + s, y), (x + s, y + s), (x, y + s), (x, y)]
    mp = GeoArrowArray.from_shapes(GeometryType.MULTIPOLYGON, [[[sq(0, 0, 2)], [sq(10, 10, 4), sq(11, 11, 1)]], [[sq(-5, -5, 1)]], [], None])
    _check(ctx, og, conv, mp)
    _check(ctx, og, conv, GeoArrowArray.points(synth.uniform_points(50)))


def test_hull_vertex_set_against_exact_referee(ctx):
    """independent check of the vertex SET with exact rational arithmetic (oracle/exact.py)"""
    from geopolars_b200 import engine as E
    from oracle import exact

    xy, ro, go = synth.blob_polygons(40, 64)
    got = E.convex_hull(ctx.upload(GeoArrowArray.polygons(xy, ro, go))).to_host()
    for i in range(40):
        ring = got.xy[got.ring_off[i] : got.ring_off[i + 1]]
        want = exact.convex_hull_vertices(xy[ro[i] : ro[i + 1]].tolist())
        assert sorted(map(tuple, ring[:-1].tolist())) == sorted(want)
        # counter-clockwise: every consecutive triple turns left, exactly
        n = len(ring) - 1
        for k in range(n):
            assert exact.orient_sign(tuple(ring[k]), tuple(ring[(k + 1) % n]), tuple(ring[(k + 2) % n])) > 0


def test_exact_ties_follow_geo_slice_order(ctx, og, conv):
    """integer lattices, regular polygons with duplicated and collinear points: many EXACT ties for the farthest
    point.  The result then depends on geo's in-place slice permutations (Hoare partition, swap_remove_to_first,
    max_by = last maximum); the kernel reproduces them, so the rings must still be identical to the oracle's."""
    rng = np.random.default_rng(11)
    shapes = []
    for n in (3, 4, 5, 7, 9, 12):
        g = np.stack(np.meshgrid(np.arange(n), np.arange(n)), -1).reshape(-1, 2).astype(float)
        for rep in range(6):
            shapes.append(rng.permutation(g).tolist())
    for k in range(40):  # random small-integer clouds with repeats
        m = int(rng.integers(4, 90))
        shapes.append(rng.integers(-3, 4, size=(m, 2)).astype(float).tolist())
    # points on a circle of radius 5 with integer coordinates + midpoints (collinear runs on the hull)
    circ = [(5, 0), (4, 3), (3, 4), (0, 5), (-3, 4), (-4, 3), (-5, 0), (-4, -3), (-3, -4), (0, -5), (3, -4), (4, -3)]
    mids = [((a[0] + b[0]) / 2, (a[1] + b[1]) / 2) for a, b in zip(circ, circ[1:] + circ[:1])]
    for rep in range(10):
        shapes.append(rng.permutation(np.array(circ + mids + circ, float)).tolist())
    arr = GeoArrowArray.from_shapes(GeometryType.MULTIPOINT, shapes)
    _check(ctx, og, conv, arr)
