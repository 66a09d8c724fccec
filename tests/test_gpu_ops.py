"""Parity of the unary/binary GeoSeries kernels against the oracle.
f64 outputs: 1e-9 relative (north_star); affine: bit-exact; booleans/indices: bit-exact."""
import math

import numpy as np
import pytest

from conftest import rel_close
from geopolars_b200 import GeoArrowArray, GeometryType, synth

pytestmark = pytest.mark.gpu

TOL = 1e-9  # relative tolerance for f64 area/distance/centroid stated by BASELINE.json north_star


def _mixed_polygons():
    """stars + degenerate shapes + holes"""
    xy, ro, go = synth.star_polygons(300, 20)
    base = GeoArrowArray.polygons(xy, ro, go)
    sq = [(0, 0), (4, 0), (4, 4), (0, 4), (0, 0)]
    hole = [(1, 1), (1, 2), (2, 2), (2, 1), (1, 1)]
    shapes = [
        [sq, hole],
        [sq],
        [[(0, 0), (1, 1), (2, 2), (0, 0)]],  # zero area: centroid falls back to the linestring
        [[(3, 3), (3, 3), (3, 3), (3, 3)]],  # all-identical: point fallback
        [[(5, 5)]],  # 1-coord ring
        [],  # empty polygon -> null centroid
        [sq, sq],  # hole covers the exterior: weight 0 -> exterior-as-linestring fallback
        [[(0, 0), (4, 0), (4, 4), (0, 4)]],  # not closed: area 0 in geo
    ]
    extra = GeoArrowArray.from_shapes(GeometryType.POLYGON, shapes)
    return base, extra


def test_area_centroid_envelope_length_polygons(ctx, og, conv):
    from geopolars_b200 import engine as E

    for arr in _mixed_polygons():
        d = ctx.upload(arr)
        o = conv(arr)
        assert rel_close(E.area(d), og.area(o), TOL)
        want_c, want_v = og.centroid(o)
        got = E.centroid(d).to_host()
        got_v = np.ones(len(arr), bool) if got.valid is None else got.valid
        assert np.array_equal(got_v, want_v)
        assert rel_close(got.xy[want_v], want_c[want_v], TOL)
        wb, wv = og.envelope(o)
        gb = E.bounds(d)
        assert np.array_equal(np.isnan(gb[:, 0]), ~wv)
        assert np.array_equal(gb[wv], wb[wv])  # min/max: bit-exact
        assert rel_close(E.euclidean_length(d), og.euclidean_length(o), TOL)
        env = E.envelope(d).to_host()
        assert env.type == GeometryType.POLYGON and env.n_coords == 5 * len(arr)


def test_multipolygon_and_linestring_measures(ctx, og, conv):
    from geopolars_b200 import engine as E

    sq = lambda x, y, s: [(x, y), (x + s, y), (x + s, y + s), (x, y + s), (x, y)]
    mp = GeoArrowArray.from_shapes(
        GeometryType.MULTIPOLYGON,
        [[[sq(0, 0, 2)], [sq(10, 10, 4), sq(11, 11, 1)]], [[sq(-5, -5, 1)]], [], None],
    )
    for arr in (mp,):
        d, o = ctx.upload(arr), conv(arr)
        assert rel_close(E.area(d), og.area(o), TOL)
        wc, wv = og.centroid(o)
        got = E.centroid(d).to_host()
        gv = np.ones(len(arr), bool) if got.valid is None else got.valid
        assert np.array_equal(gv, wv)
        assert rel_close(got.xy[wv], wc[wv], TOL)
    xy, off = synth.walk_linestrings(5000, 16)
    ls = GeoArrowArray.linestrings(xy, off)
    d, o = ctx.upload(ls), conv(ls)
    assert rel_close(E.euclidean_length(d), og.euclidean_length(o), TOL)
    wc, wv = og.centroid(o)
    got = E.centroid(d).to_host()
    assert wv.all() and got.valid is None
    assert rel_close(got.xy, wc, TOL)
    assert np.array_equal(E.area(d), np.zeros(len(ls)))
    # degenerate linestrings: repeated points only, single point, empty
    deg = GeoArrowArray.from_shapes(GeometryType.LINESTRING, [[(1, 1), (1, 1), (1, 1)], [(2, 3)], [], [(0, 0), (2, 0)]])
    wc, wv = og.centroid(conv(deg))
    got = E.centroid(ctx.upload(deg)).to_host()
    gv = np.ones(len(deg), bool) if got.valid is None else got.valid
    assert np.array_equal(gv, wv) and rel_close(got.xy[wv], wc[wv], TOL)


def test_points_config1_semantics(ctx, og, conv):
    """BASELINE config 1 semantics: centroid(points) == points bit-exactly, area == 0"""
    from geopolars_b200 import engine as E

    pts = GeoArrowArray.points(synth.uniform_points(1000, scale=360.0) - 180.0)
    d = ctx.upload(pts)
    assert np.array_equal(E.centroid(d).to_host().xy, pts.xy)
    assert np.array_equal(E.area(d), np.zeros(1000))
    assert np.array_equal(E.x(d), pts.xy[:, 0]) and np.array_equal(E.y(d), pts.xy[:, 1])
    assert (E.geom_type(d) == 0).all()


def test_affine_bit_exact(ctx, og):
    from geopolars_b200 import engine as E

    xy, ro, go = synth.blob_polygons(2000, 256)
    arr = GeoArrowArray.polygons(xy, ro, go)
    d = ctx.upload(arr)
    m = (0.8, -0.6, 10.0, 0.6, 0.8, -5.0)  # BASELINE config 5's matrix
    got = E.affine_transform(d, m).to_host()
    assert np.array_equal(got.xy, og.affine_transform(arr.xy, m))  # separately rounded mul/add: bit-exact
    assert np.array_equal(got.ring_off, arr.ring_off) and np.array_equal(got.geom_off, arr.geom_off)
    t = E.translate(d, 3.5, -1.25).to_host()
    assert np.array_equal(t.xy, og.affine_transform(arr.xy, (1.0, 0.0, 3.5, 0.0, 1.0, -1.25)))
    # fixed-point origin: coefficients follow geo's AffineTransform::{scale,rotate,skew}
    ox, oy = 2.0, -3.0
    s = E.scale(d, 2.0, 0.5, origin=(ox, oy)).to_host()
    assert np.array_equal(s.xy, og.affine_transform(arr.xy, (2.0, 0.0, ox - ox * 2.0, 0.0, 0.5, oy - oy * 0.5)))
    rad = 30.0 * (math.pi / 180.0)
    c, sn = math.cos(rad), math.sin(rad)
    r = E.rotate(d, 30.0, origin={"x": ox, "y": oy}).to_host()
    assert rel_close(r.xy, og.affine_transform(arr.xy, (c, -sn, ox - ox * c + oy * sn, sn, c, oy - ox * sn - oy * c)), 1e-12)
    tx, ty = math.tan(10.0 * (math.pi / 180.0)), math.tan(20.0 * (math.pi / 180.0))
    k = E.skew(d, 10.0, 20.0, origin=(ox, oy)).to_host()
    assert rel_close(k.xy, og.affine_transform(arr.xy, (1.0, tx, -oy * tx, ty, 1.0, -ox * ty)), 1e-12)


def test_affine_about_per_geometry_origin(ctx, og, conv):
    from geopolars_b200 import engine as E

    xy, ro, go = synth.blob_polygons(500, 64)
    arr = GeoArrowArray.polygons(xy, ro, go)
    d = ctx.upload(arr)
    cen, _ = og.centroid(conv(arr))
    bb, _ = og.envelope(conv(arr))
    ctr = np.stack([(bb[:, 2] + bb[:, 0]) / 2.0, (bb[:, 3] + bb[:, 1]) / 2.0], 1)
    for origin, org in (("centroid", cen), ("center", ctr)):
        got = E.scale(d, 2.0, 3.0, origin=origin).to_host().xy.reshape(500, 65, 2)
        src = arr.xy.reshape(500, 65, 2)
        want = np.empty_like(src)
        want[:, :, 0] = 2.0 * src[:, :, 0] + (org[:, 0] - org[:, 0] * 2.0)[:, None]
        want[:, :, 1] = 3.0 * src[:, :, 1] + (org[:, 1] - org[:, 1] * 3.0)[:, None]
        assert rel_close(got, want, 1e-9)
    with pytest.raises(ValueError):
        E.rotate(d, 10.0, origin="nonsense")


def test_linestring_pairs_intersects_and_distance(ctx, og, conv):
    from geopolars_b200 import engine as E

    n = 20000
    axy, aoff = synth.walk_linestrings(n, 16, stream=3)
    bxy, boff = synth.walk_linestrings(n, 16, stream=4, other_of=3)
    A, B = GeoArrowArray.linestrings(axy, aoff), GeoArrowArray.linestrings(bxy, boff)
    da, db = ctx.upload(A), ctx.upload(B)
    want_i = og.intersects_rowwise(conv(A), conv(B), threads=0)
    got_i = E.intersects(da, db)
    assert np.array_equal(got_i, want_i)
    assert 0.2 < want_i.mean() < 0.8  # the workload really mixes both outcomes
    want_d = og.distance_rowwise(conv(A), conv(B), threads=0)
    got_d, valid = E.distance(da, db)
    assert valid.all()
    assert np.array_equal(got_d == 0.0, want_d == 0.0)
    assert rel_close(got_d, want_d, TOL)


def test_ragged_and_degenerate_linestring_pairs(ctx, og, conv):
    from geopolars_b200 import engine as E

    a = [
        [(0, 0), (10, 0)],
        [(0, 0), (10, 0)],
        [(0, 0), (10, 0)],  # collinear overlapping
        [(0, 0), (10, 0)],  # collinear disjoint
        [(0, 0), (10, 0)],  # touching at an endpoint
        [(0, 0), (0, 0)],  # degenerate segment on the other line
        [(0, 0), (1, 1), (2, 0), (3, 1), (4, 0), (5, 1)] * 30,  # long: exceeds the shared-memory staging
        [(0, 0)],
        [(0.5, 0.5), (2, 2)],
    ]
    b = [
        [(5, -5), (5, 5)],
        [(5, 1), (5, 5)],
        [(5, 0), (15, 0)],
        [(11, 0), (15, 0)],
        [(10, 0), (10, 7)],
        [(-1, 0), (1, 0)],
        [(0, 0.5), (200, 0.5)],
        [(1, 1), (2, 2)],
        [(0, 1), (1, 0), (3, 3)],
    ]
    A = GeoArrowArray.from_shapes(GeometryType.LINESTRING, a)
    B = GeoArrowArray.from_shapes(GeometryType.LINESTRING, b)
    want_i = og.intersects_rowwise(conv(A), conv(B))
    got_i = E.intersects(ctx.upload(A), ctx.upload(B))
    assert np.array_equal(got_i, want_i)
    assert want_i.tolist()[:6] == [True, False, True, False, True, True]
    want_d = og.distance_rowwise(conv(A), conv(B))
    got_d, valid = E.distance(ctx.upload(A), ctx.upload(B))
    ok = ~np.isnan(want_d)
    assert np.array_equal(valid, ok)
    assert rel_close(got_d[ok], want_d[ok], TOL)


def test_point_distances_and_rowwise_contains(ctx, og, conv):
    from geopolars_b200 import engine as E

    n = 5000
    P = GeoArrowArray.points(synth.uniform_points(n, scale=50.0))
    Q = GeoArrowArray.points(synth.uniform_points(n, first=n, scale=50.0))
    lxy, loff = synth.walk_linestrings(n, 16)
    L = GeoArrowArray.linestrings(lxy * 0.05, loff)
    xy, ro, go = synth.star_polygons(n, 100)
    S = GeoArrowArray.polygons(xy * 0.05, ro, go)
    dP, dQ, dL, dS = (ctx.upload(v) for v in (P, Q, L, S))
    for (x, dx), (y, dy) in [((P, dP), (Q, dQ)), ((P, dP), (L, dL)), ((L, dL), (P, dP)), ((P, dP), (S, dS)), ((S, dS), (P, dP))]:
        want = og.distance_rowwise(conv(x), conv(y), threads=0)
        got, valid = E.distance(dx, dy)
        assert valid.all() and rel_close(got, want, TOL)
    pts_in = GeoArrowArray.points(0.5 * (xy.reshape(n, 65, 2)[:, 0] + xy.reshape(n, 65, 2)[:, 32]) * 0.05)
    got = E.contains(dS, ctx.upload(pts_in))
    want = np.array([og.contains_point(conv(S), i, *pts_in.xy[i]) for i in range(n)])
    assert np.array_equal(got, want) and want.any()
    from geopolars_b200 import ShapeError

    assert (E.distance(dS, dS)[0] == 0.0).all()  # Polygon x Polygon is on the path (k_distance_generic)
    # Multi* operands: the minimum over the members (test_gpu_predicates.py::test_distance_every_type_pair)
    dmp, ok = E.distance(dS, ctx.upload(GeoArrowArray.from_shapes(GeometryType.MULTIPOINT, [[(0.0, 0.0)]] * n)))
    assert ok.all() and rel_close(dmp, og.distance_rowwise(conv(S), conv(GeoArrowArray.points(np.zeros((n, 2)))), threads=0), TOL)
    with pytest.raises(ShapeError):
        E.distance(dP, ctx.upload(GeoArrowArray.points(np.zeros((3, 2)))))


def test_secondary_ops(ctx):
    from geopolars_b200 import engine as E

    sq = [(0, 0), (4, 0), (4, 4), (0, 4), (0, 0)]
    hole = [(1, 1), (1, 2), (2, 2), (2, 1), (1, 1)]
    polys = GeoArrowArray.from_shapes(GeometryType.POLYGON, [[sq, hole], [], None, [sq]])
    d = ctx.upload(polys)
    assert E.geom_type(d).tolist() == [3, 3, -1, 3]
    assert E.is_empty(d).tolist()[:2] == [False, True]
    ext = E.exterior(d).to_host()
    assert ext.type == GeometryType.LINESTRING and ext.geom_off.tolist() == [0, 5, 5, 5, 10]
    assert np.array_equal(ext.xy[:5], np.array(sq, float))
    ls = GeoArrowArray.from_shapes(GeometryType.LINESTRING, [sq, sq[:3], []])
    assert E.is_ring(ctx.upload(ls)).tolist() == [True, False, True]
    mp = GeoArrowArray.from_shapes(GeometryType.MULTIPOLYGON, [[[sq], [hole]], [[sq, hole]]])
    ex = E.explode(ctx.upload(mp)).to_host()
    assert ex.type == GeometryType.POLYGON and len(ex) == 3 and ex.geom_off.tolist() == [0, 1, 2, 4]
    sep = ctx.upload_separated(GeometryType.POLYGON, np.array(sq, float)[:, 0], np.array(sq, float)[:, 1], geom_off=[0, 1], ring_off=[0, 5])
    assert E.area(sep).tolist() == [16.0]
