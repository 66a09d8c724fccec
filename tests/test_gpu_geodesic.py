"""GeoSeries::geodesic_length (geopolars/geopolars-geo/src/geoseries.rs:52-58; methods named in
py-geopolars/src/geo.rs:64-67 and georust/geoseries.py:128-166) on the GPU against the oracle, 1e-9 relative
(CUDA math vs libm differ in the last ulps of sin/cos/atan2)."""
import numpy as np
import pytest

from conftest import rel_close
from geopolars_b200 import GeoArrowArray, GeometryType, engine

pytestmark = pytest.mark.gpu

METHODS = ["geodesic", "haversine", "vincenty"]


def lonlat_walks(rng, n, kmax, step):
    lens = rng.integers(0, kmax, n)
    off = np.concatenate([[0], np.cumsum(lens)])
    start = np.stack([rng.uniform(-170, 170, n), rng.uniform(-80, 80, n)], 1)
    xy = np.empty((off[-1], 2))
    for i in range(n):
        k = lens[i]
        if k:
            p = start[i] + np.cumsum(rng.uniform(-step, step, (k, 2)), 0)
            p[:, 1] = np.clip(p[:, 1], -89.5, 89.5)
            xy[off[i] : off[i + 1]] = p
    return xy, off


@pytest.mark.parametrize("method", METHODS)
@pytest.mark.parametrize("step", [0.001, 0.5, 20.0])
def test_linestring_lengths(ctx, og, conv, method, step):
    rng = np.random.default_rng(int(step * 1000) + len(method))
    xy, off = lonlat_walks(rng, 3000, 40, step)
    arr = GeoArrowArray.linestrings(xy, off)
    want = og.geodesic_length(conv(arr), method, threads=0)
    got, valid = engine.geodesic_length(ctx.upload(arr), method)
    assert np.array_equal(valid, ~np.isnan(want))
    assert rel_close(got[valid], want[valid], 1e-9)
    assert want[valid].max() > 0


def test_known_answers_and_hard_pairs(ctx, og, conv):
    """Karney 2013 examples, the WGS84 quarter meridian, the equator, antipodal and near-antipodal pairs"""
    pairs = [
        ((0, 0), (0, 90)),  # quarter meridian 10 001 965.729 m
        ((0, 0), (90, 0)),  # a * pi / 2
        ((0, 0), (180, 0)),  # over the pole: 2 quarter meridians
        ((0, -30), (179.8, 29.9)),  # Karney 2013 section 7: 19 989 832.827 61 m
        ((0, 40), (137.84490004377, 41.79331020506)),  # Karney 2013 direct example: 10 000 km
        ((0, -90), (0, 90)),
        ((10, 20), (10, 20)),  # coincident
        ((179.5, 0.5), (-0.3, -0.5)),
        ((-73.8, 40.6), (104, 1.4)),
        ((5, 0), (5.000001, 0.0000001)),
        ((170, 10), (-170, -10)),  # across the antimeridian
    ]
    shapes = [[a, b] for a, b in pairs]
    arr = GeoArrowArray.from_shapes(GeometryType.LINESTRING, shapes)
    dev = ctx.upload(arr)
    want = og.geodesic_length(conv(arr), "geodesic")
    got, valid = engine.geodesic_length(dev, "geodesic")
    assert valid.all() and rel_close(got, want, 1e-9)
    assert abs(got[0] - 10001965.729) < 1e-3 and abs(got[1] - 6378137.0 * np.pi / 2) < 1e-6
    assert abs(got[2] - 2 * got[0]) < 1e-6 and abs(got[3] - 19989832.82761) < 1e-4 and abs(got[4] - 1e7) < 1e-4
    assert got[6] == 0.0
    # Vincenty: antipodal pairs do not converge -> null rows exactly where the oracle reports failure
    wv = og.geodesic_length(conv(arr), "vincenty")
    gv, vv = engine.geodesic_length(dev, "vincenty")
    assert np.array_equal(vv, ~np.isnan(wv)) and not vv.all() and rel_close(gv[vv], wv[vv], 1e-9)
    gh, vh = engine.geodesic_length(dev, "haversine")
    assert vh.all() and rel_close(gh, og.geodesic_length(conv(arr), "haversine"), 1e-9)


def test_near_antipodal_cloud(ctx, og, conv):
    rng = np.random.default_rng(12)
    n = 4000
    lo1, la1 = rng.uniform(-180, 180, n), rng.uniform(-70, 70, n)
    lo2, la2 = lo1 + 180 + rng.uniform(-1.5, 1.5, n), -la1 + rng.uniform(-1.5, 1.5, n)
    xy = np.stack([np.stack([lo1, la1], 1), np.stack([lo2, la2], 1)], 1).reshape(-1, 2)
    arr = GeoArrowArray.linestrings(xy, np.arange(0, 2 * n + 1, 2))
    want = og.geodesic_length(conv(arr), "geodesic", threads=0)
    got, valid = engine.geodesic_length(ctx.upload(arr), "geodesic")
    assert valid.all() and rel_close(got, want, 1e-9) and want.min() > 1.9e7


def test_polygons_use_the_exterior_ring_and_points_are_zero(ctx, og, conv):
    sq = [(0, 0), (1, 0), (1, 1), (0, 1), (0, 0)]
    hole = [(0.2, 0.2), (0.2, 0.4), (0.4, 0.4), (0.2, 0.2)]
    poly = GeoArrowArray.from_shapes(GeometryType.POLYGON, [[sq, hole], [sq], [], None])
    for m in METHODS:
        got, valid = engine.geodesic_length(ctx.upload(poly), m)
        want = og.geodesic_length(conv(poly), m)
        assert rel_close(got, want, 1e-9) and got[0] == got[1] and got[2] == 0.0
    mp = GeoArrowArray.from_shapes(GeometryType.MULTIPOLYGON, [[[sq, hole], [[(5, 5), (6, 5), (6, 6), (5, 5)]]], []])
    assert rel_close(engine.geodesic_length(ctx.upload(mp), "geodesic")[0], og.geodesic_length(conv(mp), "geodesic"), 1e-9)
    pts = GeoArrowArray.points(np.array([[1.0, 2.0], [3.0, 4.0]]))
    assert engine.geodesic_length(ctx.upload(pts), "haversine")[0].tolist() == [0.0, 0.0]
    with pytest.raises(ValueError):
        engine.geodesic_length(ctx.upload(pts), "euclid")
