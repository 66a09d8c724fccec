"""Parity of the CUDA points-in-polygons join against the oracle (bit-exact: index outputs).
Reference semantics: geopolars/src/spatial_index.rs:89-96 (`poly.contains(point)`), golden vector
:432-484."""
import numpy as np
import pytest

from geopolars_b200 import GeoArrowArray, GeometryType, synth

pytestmark = pytest.mark.gpu


def test_golden_vector_square(ctx):
    """the reference's only numeric pin: 9 points x square -> exactly (1,1) and (10,1) are contained"""
    from geopolars_b200.engine import PipIndex

    sq = GeoArrowArray.from_shapes(GeometryType.POLYGON, [[[(0, 0), (20, 0), (20, 20), (0, 20), (0, 0)]]])
    pts = np.array([(0, 10), (1, 1), (10, 1), (1, -1), (0, -10), (-1, -1), (-10, 0), (-1, 1), (0, 10)], float)
    idx = PipIndex(ctx.upload(sq))
    first, cnt = idx.query(pts, with_count=True)
    assert first.tolist() == [-1, 0, 0, -1, -1, -1, -1, -1, -1]
    assert cnt.tolist() == [0, 1, 1, 0, 0, 0, 0, 0, 0]
    lhs, rhs = idx.pairs(pts)
    assert lhs.tolist() == [1, 2] and rhs.tolist() == [0, 0]  # inner join shape (2, ..) as in spatial_join_test


def test_unclosed_ring_is_closed_like_polygon_new(ctx):
    from geopolars_b200.engine import PipIndex

    sq = GeoArrowArray.from_shapes(GeometryType.POLYGON, [[[(0, 0), (20, 0), (20, 20), (0, 20)]]])
    pts = np.array([(0, 10), (1, 1), (10, 1), (19.5, 19.5), (20, 20), (0, 0), (-1, 5)], float)
    first = PipIndex(ctx.upload(sq)).query(pts)
    assert first.tolist() == [-1, 0, 0, 0, -1, -1, -1]


@pytest.mark.parametrize("m,grid,n", [(100, 10, 200_000), (2500, 50, 300_000)])
def test_star_polygons_vs_oracle(ctx, og, conv, m, grid, n):
    from geopolars_b200.engine import PipIndex

    xy, ro, go = synth.star_polygons(m, grid)
    polys = GeoArrowArray.polygons(xy, ro, go)
    pts = synth.uniform_points(n, scale=grid * 10.0)
    # adversarial additions: polygon vertices, edge midpoints, bbox corners, points far outside
    extra = np.concatenate([xy[:500], 0.5 * (xy[:500] + xy[1:501]), np.array([[-5.0, -5.0], [1e9, 1e9], [np.nan, 1.0]])])
    pts = np.concatenate([pts, extra])
    want_first, want_cnt = og.contains_join(conv(polys), pts, use_grid=True, threads=0)
    idx = PipIndex(ctx.upload(polys))
    first, cnt = idx.query(pts, with_count=True)
    assert np.array_equal(first, want_first)
    assert np.array_equal(cnt, want_cnt)
    lhs, rhs = idx.pairs(pts)
    hit = np.nonzero(want_first >= 0)[0]
    assert np.array_equal(lhs, hit.astype(np.uint64))
    assert np.array_equal(rhs, want_first[hit].astype(np.uint64))


def _holes_and_multis():
    # polygon with a hole, overlapping polygons, a multipolygon with two parts, a null row, an empty polygon
    outer = [(0, 0), (10, 0), (10, 10), (0, 10), (0, 0)]
    hole = [(4, 4), (4, 6), (6, 6), (6, 4), (4, 4)]
    tri = [(5, 5), (15, 5), (10, 15), (5, 5)]
    far1 = [(20, 20), (22, 20), (22, 22), (20, 22), (20, 20)]
    far2 = [(21, 21), (25, 21), (25, 25), (21, 25), (21, 21)]
    return [[[outer, hole]], [[tri]], [[far1], [far2]], None, []]


def test_holes_multipolygons_overlaps_nulls(ctx, og, conv):
    from geopolars_b200.engine import PipIndex

    polys = GeoArrowArray.from_shapes(GeometryType.MULTIPOLYGON, _holes_and_multis())
    g = np.linspace(-1, 26, 109)
    gx, gy = np.meshgrid(g, g)
    pts = np.stack([gx.ravel(), gy.ravel()], 1)  # lattice hits vertices, edges, hole borders exactly
    want_first, want_cnt = og.contains_join(conv(polys), pts, use_grid=False, threads=0)
    idx = PipIndex(ctx.upload(polys))
    first, cnt = idx.query(pts, with_count=True)
    assert np.array_equal(first, want_first)
    assert np.array_equal(cnt, want_cnt)
    assert cnt.max() == 2  # the overlap region really is exercised
    lhs, rhs = idx.pairs(pts)
    assert len(lhs) == int(want_cnt.sum())
    # pair list is sorted by (point, polygon) and consistent with per-point counts
    assert np.all(np.diff(lhs.astype(np.int64)) >= 0)
    assert np.array_equal(np.bincount(lhs.astype(np.int64), minlength=len(pts)), want_cnt)


def test_near_degenerate_points_use_exact_predicate(ctx, og, conv):
    """points a few ulps off long edges: the fast orient2d filter cannot decide, the adaptive stage must"""
    from geopolars_b200.engine import PipIndex

    big = [(1e6 + 0.1, 2e6 + 0.3), (1e6 + 1000.7, 2e6 + 500.9), (1e6 + 300.2, 2e6 + 900.4), (1e6 + 0.1, 2e6 + 0.3)]
    polys = GeoArrowArray.from_shapes(GeometryType.POLYGON, [[big]])
    a, b = np.array(big[0]), np.array(big[1])
    t = np.linspace(0.01, 0.99, 4000)
    on = a[None, :] + t[:, None] * (b - a)[None, :]
    pts = np.concatenate([on, np.nextafter(on, 1e9), np.nextafter(on, -1e9)])
    before = og.adapt_calls()
    want_first, want_cnt = og.contains_join(conv(polys), pts, use_grid=False, threads=1)
    assert og.adapt_calls() > before  # the oracle needed the adaptive stage, so the case is real
    first, cnt = PipIndex(ctx.upload(polys)).query(pts, with_count=True)
    assert np.array_equal(first, want_first)
    assert 0 < (want_first >= 0).sum() < len(pts)


def test_empty_inputs(ctx):
    from geopolars_b200.engine import PipIndex

    polys = GeoArrowArray.polygons(np.zeros((0, 2)), np.zeros(1, np.int64), np.zeros(1, np.int64))
    idx = PipIndex(ctx.upload(polys))
    assert idx.query(np.array([[1.0, 1.0]])).tolist() == [-1]
    xy, ro, go = synth.star_polygons(4, 2)
    idx2 = PipIndex(ctx.upload(GeoArrowArray.polygons(xy, ro, go)))
    assert idx2.query(np.zeros((0, 2))).shape == (0,)
