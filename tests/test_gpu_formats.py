"""WKB / Arrow C Data Interface / GeoSeries accessor on the GPU path, on the reference's own fixtures.
BASELINE config 1: data/cities.arrow centroid() + area() through the full boundary."""
import os

import numpy as np
import pyarrow as pa
import pytest

from conftest import rel_close
from geopolars_b200 import GeoArrowArray, GeometryType
from wkbutil import column_to_shapes

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def wkb_column(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    arr = pa.Array.from_buffers(pa.binary(), len(z["offsets"]) - 1, [None, pa.py_buffer(z["offsets"].tobytes()), pa.py_buffer(z["bytes"].tobytes())])
    return arr, z


def test_config1_cities_through_the_accessor(ctx):
    """gs.geo.centroid / gs.geo.area on data/cities.arrow: centroid == input points bit-exactly, area == 0"""
    from geopolars_b200 import geoseries as G

    G.set_context(ctx)
    col, z = wkb_column("cities")
    gs = G.from_arrow(col)
    assert len(gs) == 202
    area = gs.geo.area
    assert area.to_pylist() == [0.0] * 202
    cen = gs.geo.centroid
    _, shapes = column_to_shapes(z["offsets"], z["bytes"])
    want = np.array(shapes, dtype=np.float64)
    got = cen.device.to_host()
    assert got.type == GeometryType.POINT and np.array_equal(got.xy, want)
    assert np.array_equal(np.asarray(gs.geo.x), want[:, 0]) and np.array_equal(np.asarray(gs.geo.y), want[:, 1])
    assert gs.geo.geom_type.to_pylist() == [0] * 202
    # WKB in == WKB out, byte for byte (first row: 01 01000000 4933fe4722e82840 80fe1ec09ef34440)
    back = gs.to_wkb()
    assert back.equals(col)
    assert back[0].as_py().hex() == "01010000004933fe4722e8284080fe1ec09ef34440"


@pytest.mark.parametrize("name", ["naturalearth_lowres", "nybb"])
def test_wkb_decode_and_measures_on_reference_datasets(ctx, og, conv, name):
    from geopolars_b200 import engine as E

    col, z = wkb_column(name)
    d = ctx.import_arrow(col)
    code, shapes = column_to_shapes(z["offsets"], z["bytes"])
    want = GeoArrowArray.from_shapes(GeometryType(code), shapes)
    got = d.to_host()
    assert got.type == want.type
    for k in ("xy", "geom_off", "part_off", "ring_off"):
        assert np.array_equal(getattr(got, k), getattr(want, k)), k
    o = conv(want)
    assert rel_close(E.area(d), og.area(o), 1e-9)
    wc, wv = og.centroid(o)
    gc = E.centroid(d).to_host()
    assert wv.all() and rel_close(gc.xy, wc, 1e-9)
    wb, _ = og.envelope(o)
    assert np.array_equal(E.bounds(d), wb)
    assert rel_close(E.euclidean_length(d), og.euclidean_length(o), 1e-9)
    off, hxy = og.convex_hull(o, threads=0)
    hull = E.convex_hull(d).to_host()
    assert np.array_equal(hull.ring_off, off) and np.array_equal(hull.xy, hxy)
    if name == "nybb":  # skewed rings (largest: 16 051 coords) + the Shape_Area attribute as a soft check
        assert np.all(np.abs(E.area(d) / z["Shape_Area"] - 1) < 2e-6)
        # points-in-boroughs join on the real multipolygons
        b = E.bounds(d)
        rng = np.random.default_rng(3)
        pts = np.stack([rng.uniform(b[:, 0].min(), b[:, 2].max(), 20000), rng.uniform(b[:, 1].min(), b[:, 3].max(), 20000)], 1)
        first, cnt = E.PipIndex(d).query(pts, with_count=True)
        wf, wc2 = og.contains_join(o, pts, use_grid=True, threads=0)
        assert np.array_equal(first, wf) and np.array_equal(cnt, wc2) and (wf >= 0).mean() > 0.1
    # WKB round trip: decode -> encode -> decode gives identical buffers
    again = ctx.import_arrow(d.to_wkb()).to_host()
    assert np.array_equal(again.xy, got.xy) and np.array_equal(again.ring_off, got.ring_off)


def test_arrow_c_data_interface_layouts(ctx):
    from geopolars_b200 import engine as E

    sq = [(0.0, 0.0), (4.0, 0.0), (4.0, 4.0), (0.0, 4.0), (0.0, 0.0)]
    # interleaved FixedSizeList coords, List<List<>> polygon, with a null row
    coord_t = pa.list_(pa.float64(), 2)
    poly = pa.array([[sq], None, [sq, [(1.0, 1.0), (1.0, 2.0), (2.0, 2.0), (1.0, 1.0)]]], type=pa.list_(pa.list_(coord_t)))
    d = ctx.import_arrow(poly)
    assert d.type == GeometryType.POLYGON
    assert E.area(d).tolist() == [16.0, 0.0, 15.5]
    assert E.geom_type(d).tolist() == [3, -1, 3]
    # separated Struct{x,y} coords (what py-geopolars builds, internals/geoseries.py:87-107), LargeList offsets
    st = pa.struct([("x", pa.float64()), ("y", pa.float64())])
    ls = pa.array([[{"x": 0.0, "y": 0.0}, {"x": 3.0, "y": 4.0}], [{"x": 1.0, "y": 1.0}, {"x": 1.0, "y": 2.0}, {"x": 3.0, "y": 2.0}]],
                  type=pa.large_list(st))
    dl = ctx.import_arrow(ls)
    assert dl.type == GeometryType.LINESTRING and E.euclidean_length(dl).tolist() == [5.0, 3.0]
    # sliced arrays carry a non-zero Arrow offset
    dsl = ctx.import_arrow(ls.slice(1, 1))
    assert len(dsl) == 1 and E.euclidean_length(dsl).tolist() == [3.0]
    pts = pa.array([(1.0, 2.0), (3.0, 4.0)], type=coord_t)
    dp = ctx.import_arrow(pts)
    assert dp.type == GeometryType.POINT and E.x(dp).tolist() == [1.0, 3.0]
    # export: geoarrow extension metadata + same values back
    out = d.to_arrow()
    assert out.type == pa.list_(pa.list_(coord_t)) or str(out.type).startswith("list<")
    assert out.null_count == 1 and out[0].as_py() == [[list(c) for c in sq]]
    again = ctx.import_arrow(out)
    assert E.area(again).tolist() == [16.0, 0.0, 15.5]
    hull = E.convex_hull(d).to_arrow()
    assert len(hull) == 3
    with pytest.raises(Exception):
        ctx.import_arrow(pa.array([1, 2, 3]))


def test_geoseries_accessor_surface(ctx):
    """every GeoRustSeries entry point of the reference (georust/geoseries.py:22-320) is callable"""
    from geopolars_b200 import geoseries as G

    G.set_context(ctx)
    col, _ = wkb_column("naturalearth_lowres")
    gs = G.GeoSeries(col)
    geo = gs.geo
    assert len(geo.area) == 177 and len(geo.euclidean_length()) == 177
    assert len(geo.centroid) == 177 and len(geo.convex_hull()) == 177 and len(geo.envelope()) == 177
    assert len(geo.explode()) >= 177
    assert set(geo.geom_type.to_pylist()) == {6}
    assert geo.is_empty().to_pylist().count(True) == 0
    m = [0.8, -0.6, 10.0, 0.6, 0.8, -5.0]  # geo's AffineTransform::from([a, b, xoff, d, e, yoff]), see test_affine_matrix_order_is_geos
    t = geo.affine_transform(m).device.to_host()
    src = gs.device.to_host()
    assert np.array_equal(t.xy[:, 0], 0.8 * src.xy[:, 0] + -0.6 * src.xy[:, 1] + 10.0)
    assert np.array_equal(t.xy[:, 1], 0.6 * src.xy[:, 0] + 0.8 * src.xy[:, 1] + -5.0)
    for g2 in (geo.translate(1.0, 2.0), geo.rotate(30.0), geo.rotate(30.0, origin="centroid"), geo.scale(2.0, 2.0, origin=(0.0, 0.0)), geo.skew(5.0, 0.0)):
        assert len(g2) == 177
    # rotation about the centroid preserves area and centroid
    r = geo.rotate(90.0, origin="centroid")
    assert rel_close(np.asarray(r.geo.area), np.asarray(geo.area), 1e-9)
    d = geo.centroid.geo.distance(r.geo.centroid)
    assert np.nanmax(np.asarray(d)) < 1e-6
    cen = geo.centroid
    inside = geo.contains(cen).to_pylist()
    assert 100 < sum(inside) <= 177  # most country centroids fall inside the country
    lhs, rhs = G.spatial_join(cen, gs, how="inner")
    assert len(lhs) >= sum(inside)
    l2, r2 = G.spatial_join(cen, gs, how="left")
    assert set(l2.tolist()) == set(range(177))
    simp = geo.simplify(0.1)
    assert len(simp) == 177 and simp.device.view().n_coords < gs.device.view().n_coords
    assert (np.asarray(simp.geo.area) > 0).all()
    lengths = {m: np.asarray(geo.geodesic_length(m)) for m in ("geodesic", "haversine", "vincenty")}
    ok = ~np.isnan(lengths["vincenty"].astype(float))  # country outlines: lon/lat degrees -> metres
    assert ok.sum() >= 170 and rel_close(lengths["geodesic"][ok], lengths["vincenty"][ok], 1e-5)
    assert rel_close(lengths["geodesic"], lengths["haversine"], 1e-2) and lengths["geodesic"].min() > 1e4
    with pytest.raises(ValueError):
        geo.geodesic_length("nope")


def test_config1_through_read_dataset(ctx):
    """BASELINE configs[0], literally: `read_dataset("cities").geometry.geo.centroid / .area` (datasets/__init__.py:39-42):
    centroid(points) == points bit for bit, area == 0, 202 rows; and the Shape_Area attribute of nybb against area()"""
    from geopolars_b200 import datasets
    from geopolars_b200 import geoseries as G

    G.set_context(ctx)
    gdf = datasets.read_dataset("cities")
    assert gdf.shape[0] == 202 and "geometry" in gdf.columns
    geo = gdf.geometry.geo
    assert np.asarray(geo.area).tolist() == [0.0] * 202
    cen = geo.centroid.device.to_host()
    src = gdf.geometry.device.to_host()
    assert np.array_equal(cen.xy, src.xy)
    assert len(datasets.read_dataset("naturalearth_cities")) == 243  # py-geopolars/tests/unit/internals/test_geoseries.py:4-5
    nybb = datasets.read_dataset("nybb")
    a = np.asarray(nybb.geometry.geo.area)
    assert rel_close(a, np.asarray(nybb["Shape_Area"]), 2e-6)
    t = nybb.to_arrow()
    assert isinstance(G.from_arrow(t), G.GeoDataFrame) and isinstance(G.from_arrow(t.column("geometry")), G.GeoSeries)
    with pytest.raises(ValueError):
        datasets.get_path("atlantis")


def test_affine_matrix_order_is_geos(ctx):
    """The reference binding hands the Python list to geo untouched (py-geopolars/src/geo.rs:10-13:
    `series.affine_transform(transform)` with `transform: [f64; 6]`, and geo's `From<[T; 6]>` is
    `AffineTransform::new(a, b, xoff, d, e, yoff)`), so the third number is the x offset — not `d` as the
    docstring at georust/geoseries.py:33 says.  Parity follows the executable path."""
    from geopolars_b200 import geoseries as G

    G.set_context(ctx)
    coord_t = pa.list_(pa.float64(), 2)
    pts = G.GeoSeries(pa.array([(1.0, 2.0), (-3.0, 0.5)], type=coord_t))
    out = pts.geo.affine_transform([2.0, 3.0, 100.0, 5.0, 7.0, 1000.0]).device.to_host().xy
    assert out.tolist() == [[2.0 * 1.0 + 3.0 * 2.0 + 100.0, 5.0 * 1.0 + 7.0 * 2.0 + 1000.0],
                            [2.0 * -3.0 + 3.0 * 0.5 + 100.0, 5.0 * -3.0 + 7.0 * 0.5 + 1000.0]]


def test_reference_join_shapes(ctx):
    """spatial_join_test (spatial_index.rs:432-484): inner -> 2 rows, left -> 9 rows"""
    from geopolars_b200 import geoseries as G

    G.set_context(ctx)
    z = np.load(os.path.join(GOLD, "contains_golden.npz"))
    coord_t = pa.list_(pa.float64(), 2)
    pts = G.GeoSeries(pa.array([tuple(p) for p in z["points"]], type=coord_t))
    sq = [tuple(p) for p in z["square"]] + [tuple(z["square"][0])]
    polys = G.GeoSeries(pa.array([[sq]], type=pa.list_(pa.list_(coord_t))))
    lhs, rhs = G.spatial_join(pts, polys, how="inner")
    assert lhs.tolist() == [1, 2] and rhs.tolist() == [0, 0]
    l2, r2 = G.spatial_join(pts, polys, how="left")
    assert len(l2) == 9 and (r2 >= 0).sum() == 2
