"""Row-wise intersects / contains over every GeoArrow type pair against the oracle (bit-exact booleans).
Reference call sites: geopolars/src/spatial_index.rs:89-137 (type-pair dispatch of the join), geo 0.27
Intersects / Contains (recalled; restated in oracle/geo_oracle.c)."""
import zlib

import numpy as np
import pytest

import shapes
from geopolars_b200 import GeoArrowArray, GeometryType, engine
from conftest import rel_close
from geopolars_b200._lib import MismatchedGeometry

pytestmark = pytest.mark.gpu

T = {
    "point": GeometryType.POINT,
    "multipoint": GeometryType.MULTIPOINT,
    "linestring": GeometryType.LINESTRING,
    "multilinestring": GeometryType.MULTILINESTRING,
    "polygon": GeometryType.POLYGON,
    "multipolygon": GeometryType.MULTIPOLYGON,
}


@pytest.mark.parametrize("ka", shapes.KINDS)
@pytest.mark.parametrize("kb", shapes.KINDS)
def test_intersects_every_type_pair(ctx, og, conv, ka, kb):
    rng = np.random.default_rng(zlib.crc32(f"{ka}x{kb}".encode()))
    n = 700
    rb = shapes.random_rows(rng, kb, n)
    ra = shapes.plant_touching(rng, ka, shapes.random_rows(rng, ka, n), kb, rb)
    ra[5] = None  # null row -> false
    ga, gb = GeoArrowArray.from_shapes(T[ka], ra), GeoArrowArray.from_shapes(T[kb], rb)
    want = og.intersects_rowwise(conv(ga), conv(gb), threads=0)
    got = engine.intersects(ctx.upload(ga), ctx.upload(gb))
    assert np.array_equal(got, want)
    assert 0 < want.sum() < n and not got[5]


@pytest.mark.parametrize("ka", shapes.KINDS)
@pytest.mark.parametrize("kb", shapes.KINDS)
def test_distance_every_type_pair(ctx, og, conv, ka, kb):
    """GeoSeries::distance (geoseries.rs:141-146) over every pair of GeoArrow types; Multi* = minimum over the members
    (geo's impl for iterable geometries), empty collection -> f64::MAX, geo-would-panic rows -> null.  The reference's own
    datasets (nybb, naturalearth_lowres) are MultiPolygon columns."""
    rng = np.random.default_rng(zlib.crc32(f"d{ka}x{kb}".encode()))
    n = 500
    rb = shapes.random_rows(rng, kb, n)
    ra = shapes.plant_touching(rng, ka, shapes.random_rows(rng, ka, n), kb, rb)
    ra[5] = None
    ga, gb = GeoArrowArray.from_shapes(T[ka], ra), GeoArrowArray.from_shapes(T[kb], rb)
    want = og.distance_rowwise(conv(ga), conv(gb), threads=0)
    got, valid = engine.distance(ctx.upload(ga), ctx.upload(gb))
    ok = ~np.isnan(want)
    assert np.array_equal(valid, ok) and not valid[5]
    assert rel_close(got[ok], want[ok], 1e-9)
    assert (want[ok] == 0.0).any() and (want[ok] > 0.0).any()
    zero = ok & (want == 0.0)
    assert (got[zero] == 0.0).all()  # intersecting rows are exactly 0, not merely small


def test_distance_multipolygon_datasets(ctx, og, conv):
    """the reference's bundled (Multi)Polygon columns: every row against the centroid of the NEXT row (MultiPolygon x Point)
    and against itself (MultiPolygon x MultiPolygon -> 0: a geometry intersects itself)"""
    from geopolars_b200 import engine as E
    from test_oracle_cpu import load

    for name in ("nybb", "naturalearth_lowres"):
        arr, _ = load(name)
        d = ctx.upload(arr)
        cw, cv = og.centroid(conv(arr))
        assert cv.all()
        pts = GeoArrowArray.points(np.roll(cw, 1, axis=0))
        want = og.distance_rowwise(conv(arr), conv(pts), threads=0)
        got, valid = E.distance(d, ctx.upload(pts))
        assert valid.all() and rel_close(got, want, 1e-9) and (want > 0).sum() > len(arr) // 2
        got2, valid2 = E.distance(ctx.upload(pts), d)  # symmetric impl
        assert valid2.all() and rel_close(got2, want, 1e-9)
        self_d, self_ok = E.distance(d, d)
        want_self = og.distance_rowwise(conv(arr), conv(arr), threads=0)
        assert np.array_equal(self_ok, ~np.isnan(want_self)) and (self_d[self_ok] == 0.0).all()


def test_polygon_pairs_against_exact_rational_referee(ctx):
    from oracle import exact

    rng = np.random.default_rng(11)
    A, B = shapes.random_rows(rng, "polygon", 500), shapes.random_rows(rng, "polygon", 500)
    got = engine.intersects(ctx.upload(GeoArrowArray.from_shapes(T["polygon"], A)), ctx.upload(GeoArrowArray.from_shapes(T["polygon"], B)))
    want = np.array([exact.polygons_intersect(a, b) for a, b in zip(A, B)])
    assert np.array_equal(got, want) and want.any() and not want.all()


def test_intersects_degenerate_and_invalid_polygons(ctx, og, conv):
    sq = [(0, 0), (10, 0), (10, 10), (0, 10), (0, 0)]
    hole = [(4, 4), (4, 6), (6, 6), (6, 4), (4, 4)]
    stray_hole = [(20, 20), (20, 22), (22, 22), (22, 20), (20, 20)]  # invalid: hole outside the exterior
    small = [(4.5, 4.5), (5.5, 4.5), (5.5, 5.5), (4.5, 5.5), (4.5, 4.5)]  # strictly inside `hole`
    touch = [(10, 10), (12, 10), (12, 12), (10, 12), (10, 10)]  # shares one corner with sq
    far = [(21, 21), (21.5, 21), (21.5, 21.5), (21, 21)]  # inside stray_hole's box only
    big = [(-5, -5), (30, -5), (30, 30), (-5, 30), (-5, -5)]
    unclosed = [(1, 1), (3, 1), (3, 3), (1, 3)]
    A = [[sq, hole], [sq, hole], [sq], [sq, stray_hole], [sq, stray_hole], [sq, hole], [[]], [sq], [[(1, 1)]], [unclosed], [sq, hole]]
    B = [[small], [big], [touch], [far], [[(19, 19), (23, 19), (23, 23), (19, 23), (19, 19)]], [hole], [sq], [], [sq], [[(2, 2), (2.5, 2), (2.5, 2.5)]], [[(5, 5), (9, 5), (9, 9), (5, 5)]]]
    ga, gb = GeoArrowArray.from_shapes(T["polygon"], A), GeoArrowArray.from_shapes(T["polygon"], B)
    want = og.intersects_rowwise(conv(ga), conv(gb))
    got = engine.intersects(ctx.upload(ga), ctx.upload(gb))
    assert np.array_equal(got, want), (got, want)
    assert want[:3].tolist() == [False, True, True] and not want[6] and not want[7]
    # both argument orders
    assert np.array_equal(engine.intersects(ctx.upload(gb), ctx.upload(ga)), og.intersects_rowwise(conv(gb), conv(ga)))


def test_linestring_contains_point(ctx, og, conv):
    rng = np.random.default_rng(3)
    n = 600
    L = shapes.random_rows(rng, "linestring", n)
    P = shapes.plant_touching(rng, "point", shapes.random_rows(rng, "point", n), "linestring", L, every=2)
    L[0], P[0] = [(0, 0), (2, 0), (2, 2), (0, 0)], (0, 0)  # closed: the end point is interior
    L[1], P[1] = [(0, 0), (2, 0), (2, 2)], (0, 0)  # open: boundary -> false
    L[2], P[2] = [(0, 0), (2, 0), (2, 2)], (2, 0)  # interior vertex -> true
    L[3], P[3] = [(1, 1), (1, 1)], (1, 1)
    L[4], P[4] = [], (1, 1)
    gl = GeoArrowArray.from_shapes(T["linestring"], L)
    pts = np.array(P, float)
    want = og.contains_rowwise(conv(gl), pts)
    got = engine.contains(ctx.upload(gl), ctx.upload(GeoArrowArray.points(pts)))
    assert np.array_equal(got, want)
    assert want[:5].tolist() == [True, False, True, True, False] and want[5:].any()
    M = shapes.random_rows(rng, "multilinestring", n)
    PM = shapes.plant_touching(rng, "point", shapes.random_rows(rng, "point", n), "multilinestring", M, every=2)
    gm = GeoArrowArray.from_shapes(T["multilinestring"], M)
    pm = np.array(PM, float)
    assert np.array_equal(engine.contains(ctx.upload(gm), ctx.upload(GeoArrowArray.points(pm))), og.contains_rowwise(conv(gm), pm))


def test_contains_rejects_unsupported_pairs(ctx):
    a = ctx.upload(GeoArrowArray.from_shapes(T["polygon"], [[[(0, 0), (1, 0), (1, 1), (0, 0)]]]))
    assert engine.contains(a, a).tolist() == [True]  # Polygon contains Polygon: geo's relate semantics (test_gpu_join.py)
    ls = ctx.upload(GeoArrowArray.from_shapes(T["linestring"], [[(0, 0), (1, 1)]]))
    with pytest.raises(MismatchedGeometry):
        engine.contains(a, ls)  # not a pair the reference's join dispatches (spatial_index.rs:89-137)
    with pytest.raises(MismatchedGeometry):
        engine.contains(ls, a)


@pytest.mark.parametrize("ka,kb", [("linestring", "polygon"), ("polygon", "linestring"), ("polygon", "polygon")])
def test_distance_linestring_polygon_pairs(ctx, og, conv, ka, kb):
    rng = np.random.default_rng(zlib.crc32(f"d{ka}x{kb}".encode()))
    n = 600
    ra, rb = shapes.random_rows(rng, ka, n, span=14.0), shapes.random_rows(rng, kb, n, span=14.0)
    sq = [(0, 0), (10, 0), (10, 10), (0, 10), (0, 0)]
    hole = [(2, 2), (2, 8), (8, 8), (8, 2), (2, 2)]
    inner = [(4, 4), (6, 4), (6, 6), (4, 6), (4, 4)]
    # a geometry sitting in the other's hole: measured against the hole ring (distance 2)
    if ka == "polygon":
        ra[0] = [sq, hole]
        rb[0] = [inner] if kb == "polygon" else inner[:3]
    else:
        ra[0], rb[0] = inner[:3], [sq, hole]
    ra[1] = None
    ga, gb = GeoArrowArray.from_shapes(T[ka], ra), GeoArrowArray.from_shapes(T[kb], rb)
    want = og.distance_rowwise(conv(ga), conv(gb), threads=0)
    got, valid = engine.distance(ctx.upload(ga), ctx.upload(gb))
    assert want[0] == 2.0 and np.isnan(want[1]) and not valid[1]
    assert rel_close(got, want, 1e-9)
    assert np.array_equal(valid, ~np.isnan(want))
    zero = og.intersects_rowwise(conv(ga), conv(gb))
    assert np.array_equal(got[valid] == 0.0, zero[valid]) and zero.any() and not zero.all()
