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


def test_large_rings_overlapping_boxes_and_deferred_overflow(ctx, og, conv):
    """POLYGON rows that are not eligible for the FP32 fast table (buckets longer than 63 edges), cells with
    several candidates (overlapping boxes), and a point set made almost entirely of boundary points so that
    the deferred list overflows and the exact kernel falls back to scanning the id column."""
    from geopolars_b200.engine import PipIndex

    k = np.arange(3000)
    th = 2 * np.pi * k / 3000
    r = 10 + 3 * np.sin(40 * th)
    big = np.stack([r * np.cos(th), r * np.sin(th)], 1)
    big = np.concatenate([big, big[:1]])
    sq = lambda x, y, s: [(x, y), (x + s, y), (x + s, y + s), (x, y + s), (x, y)]
    shapes = [[big.tolist()], [sq(-3, -3, 6)], [sq(-1, -1, 9)], [sq(20, 20, 1)]]
    polys = GeoArrowArray.from_shapes(GeometryType.POLYGON, shapes)
    pts = np.concatenate([synth.uniform_points(150_000, scale=30.0) - 14.0, big[:-1], 0.5 * (big[:-1] + big[1:])])
    want_first, want_cnt = og.contains_join(conv(polys), pts, use_grid=True, threads=0)
    idx = PipIndex(ctx.upload(polys))
    first, cnt = idx.query(pts, with_count=True)
    assert np.array_equal(first, want_first) and np.array_equal(cnt, want_cnt)
    assert cnt.max() == 3
    # every point on a boundary: all deferred (list capacity is max(n/16, 65536) -> overflow -> scan path)
    xy, ro, go = synth.star_polygons(400, 20)
    stars = GeoArrowArray.polygons(xy, ro, go)
    t = np.linspace(0.0, 1.0, 41)[None, :, None]
    on_edges = (xy[:-1, None, :] * (1 - t) + xy[1:, None, :] * t).reshape(-1, 2)[:1_200_000]
    wf, wc = og.contains_join(conv(stars), on_edges, use_grid=True, threads=0)
    f2, c2 = PipIndex(ctx.upload(stars)).query(on_edges, with_count=True)
    assert np.array_equal(f2, wf) and np.array_equal(c2, wc)


def test_empty_polygon_next_to_polygon_with_hole(ctx, og, conv):
    """ring count == polygon count although a hole exists: the hole table must still be built"""
    from geopolars_b200.engine import PipIndex

    sq = [(0, 0), (10, 0), (10, 10), (0, 10), (0, 0)]
    hole = [(4, 4), (4, 6), (6, 6), (6, 4), (4, 4)]
    polys = GeoArrowArray.from_shapes(GeometryType.POLYGON, [[], [sq, hole]])
    assert polys.n_rings == len(polys) == 2
    g = np.linspace(-1, 11, 49)
    pts = np.stack(np.meshgrid(g, g), -1).reshape(-1, 2)
    want, wc = og.contains_join(conv(polys), pts, use_grid=False)
    first, cnt = PipIndex(ctx.upload(polys)).query(pts, with_count=True)
    assert np.array_equal(first, want) and np.array_equal(cnt, wc) and (want == 1).any() and (want == -1).any()
