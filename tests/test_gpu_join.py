"""General spatial join and (Multi)Polygon.contains(Polygon) against the oracle and the exact-rational referee.
Reference: geopolars/src/spatial_index.rs:37-157 (candidate generation + type-pair dispatch + pair list)."""
import zlib

import numpy as np
import pytest

import shapes
from geopolars_b200 import GeoArrowArray, GeometryType, engine

pytestmark = pytest.mark.gpu

T = {
    "point": GeometryType.POINT,
    "multipoint": GeometryType.MULTIPOINT,
    "linestring": GeometryType.LINESTRING,
    "multilinestring": GeometryType.MULTILINESTRING,
    "polygon": GeometryType.POLYGON,
    "multipolygon": GeometryType.MULTIPOLYGON,
}


def test_polygon_contains_polygon_rowwise(ctx, og, conv):
    from oracle import exact

    rng = np.random.default_rng(5)
    A, B = shapes.contains_cases(rng, 400)
    ga, gb = GeoArrowArray.from_shapes(T["polygon"], A + [None]), GeoArrowArray.from_shapes(T["polygon"], B + [B[0]])
    want = og.contains_polygon_rowwise(conv(ga), conv(gb), threads=0)
    got = engine.contains(ctx.upload(ga), ctx.upload(gb))
    assert np.array_equal(got, want) and not got[-1]
    assert want[:18].tolist() == [True, True, True, False, False, False, True, True, True, True, True, True, False, False, False, True, True, True]
    ref = np.array([exact.region_contains_polygon([a], b) for a, b in zip(A[:150], B[:150])])
    assert np.array_equal(got[:150], ref)
    assert 50 < want.sum() < len(want) - 50
    MA, MB = shapes.multi_contains_cases()
    ma, mb = GeoArrowArray.from_shapes(T["multipolygon"], MA), GeoArrowArray.from_shapes(T["polygon"], MB)
    wantm = og.contains_polygon_rowwise(conv(ma), conv(mb), threads=0)
    gotm = engine.contains(ctx.upload(ma), ctx.upload(mb))
    assert np.array_equal(gotm, wantm) and wantm.tolist() == [True, True, False, True, True, False, False]
    assert wantm.tolist() == [exact.region_contains_polygon(a, b) for a, b in zip(MA, MB)]


def _expected_pairs(og, conv, ka, A, kb, B, predicate):
    """the reference's dispatch (spatial_index.rs:89-137) over ALL pairs, through the oracle's row-wise functions"""
    n, m = len(A), len(B)
    At = [A[i] for i in range(n) for _ in range(m)]
    Bt = [B[j] for _ in range(n) for j in range(m)]
    ga, gb = GeoArrowArray.from_shapes(T[ka], At), GeoArrowArray.from_shapes(T[kb], Bt)
    area, lines = ("polygon", "multipolygon"), ("linestring", "multilinestring")
    if ka == "point" and kb in area + lines:
        hit = og.contains_rowwise(conv(gb), ga.xy, threads=0)
    elif kb == "point" and ka in area + lines:
        hit = og.contains_rowwise(conv(ga), gb.xy, threads=0)
    elif ka in area and kb == "polygon":
        hit = og.contains_polygon_rowwise(conv(ga), conv(gb), threads=0) if predicate == "contains" else og.intersects_rowwise(conv(ga), conv(gb), threads=0)
    elif ka == "polygon" and kb == "multipolygon" and predicate == "intersects":
        hit = og.intersects_rowwise(conv(ga), conv(gb), threads=0)
    else:
        hit = np.zeros(n * m, dtype=bool)
    k = np.nonzero(hit)[0]
    return (k // m).astype(np.uint64), (k % m).astype(np.uint64)


@pytest.mark.parametrize("ka,kb", [("polygon", "polygon"), ("multipolygon", "polygon"), ("polygon", "multipolygon"), ("point", "polygon"),
                                   ("polygon", "point"), ("point", "multipolygon"), ("linestring", "point"), ("point", "multilinestring"),
                                   ("multilinestring", "point"), ("linestring", "linestring"), ("multipolygon", "multipolygon"),
                                   ("multipoint", "polygon")])
@pytest.mark.parametrize("predicate", ["intersects", "contains"])
def test_spatial_join_every_dispatch_arm(ctx, og, conv, ka, kb, predicate):
    rng = np.random.default_rng(zlib.crc32(f"j{ka}x{kb}".encode()))
    n, m = 45, 38
    A = shapes.random_rows(rng, ka, n, span=12.0)
    B = shapes.random_rows(rng, kb, m, span=12.0)
    A = shapes.plant_touching(rng, ka, A, kb, (B * 2)[:n])
    if ka in ("polygon",) and kb == "polygon":  # nested pairs so that `contains` has hits
        for i in range(0, 12):
            B[i] = [shapes.valid_star(rng, 6, 6, 1.0, 5)]
            A[i] = [shapes.sq(2, 2, 8)]
    if ka == "multipolygon" and kb == "polygon":
        for i in range(0, 8):
            B[i] = [shapes.sq(3, 3, 1)]
            A[i] = [[shapes.sq(2, 2, 4)], [shapes.sq(8, 8, 2)]]
    A[3] = None
    want_l, want_r = _expected_pairs(og, conv, ka, A, kb, B, predicate)
    ga, gb = GeoArrowArray.from_shapes(T[ka], A), GeoArrowArray.from_shapes(T[kb], B)
    got_l, got_r = engine.spatial_join(ctx.upload(ga), ctx.upload(gb), predicate)
    assert np.array_equal(got_l, want_l) and np.array_equal(got_r, want_r)
    dispatched = not ((ka, kb) in (("linestring", "linestring"), ("multipolygon", "multipolygon"), ("multipoint", "polygon"))
                      or ((ka, kb) == ("polygon", "multipolygon") and predicate == "contains"))
    if not dispatched:
        assert len(want_l) == 0  # the reference's `_ => false` arm
    elif "linestring" not in ka + kb:
        assert len(want_l) > 0


def test_spatial_join_larger_and_accessor(ctx, og, conv):
    """2 000 x 1 500 polygons (candidate grid with many cells, long candidate lists) and the GeoSeries-level join types"""
    from geopolars_b200 import geoseries as G
    import pyarrow as pa

    rng = np.random.default_rng(77)
    A = [[shapes.valid_star(rng, rng.uniform(0, 100), rng.uniform(0, 100), rng.uniform(0.5, 6), int(rng.integers(3, 9)))] for _ in range(2000)]
    B = [[shapes.valid_star(rng, rng.uniform(0, 100), rng.uniform(0, 100), rng.uniform(0.3, 3), int(rng.integers(3, 7)))] for _ in range(1500)]
    ga, gb = GeoArrowArray.from_shapes(T["polygon"], A), GeoArrowArray.from_shapes(T["polygon"], B)
    da, db = ctx.upload(ga), ctx.upload(gb)
    for predicate in ("intersects", "contains"):
        l, r = engine.spatial_join(da, db, predicate)
        # verify every reported pair and a sample of unreported envelope-intersecting pairs row-wise through the oracle
        fn = og.intersects_rowwise if predicate == "intersects" else og.contains_polygon_rowwise
        sub_a = GeoArrowArray.from_shapes(T["polygon"], [A[int(i)] for i in l])
        sub_b = GeoArrowArray.from_shapes(T["polygon"], [B[int(j)] for j in r])
        assert len(l) > (500 if predicate == "intersects" else 20) and fn(conv(sub_a), conv(sub_b), threads=0).all()
        ba, _ = og.envelope(conv(ga))
        bb, _ = og.envelope(conv(gb))
        ii, jj = np.nonzero((ba[:, None, 0] <= bb[None, :, 2]) & (bb[None, :, 0] <= ba[:, None, 2]) & (ba[:, None, 1] <= bb[None, :, 3]) & (bb[None, :, 1] <= ba[:, None, 3]))
        cand_a = GeoArrowArray.from_shapes(T["polygon"], [A[int(i)] for i in ii])
        cand_b = GeoArrowArray.from_shapes(T["polygon"], [B[int(j)] for j in jj])
        hit = fn(conv(cand_a), conv(cand_b), threads=0)
        assert np.array_equal(np.stack([ii[hit], jj[hit]], 1), np.stack([l.astype(np.int64), r.astype(np.int64)], 1))
    G.set_context(ctx)
    sa, sb = G.GeoSeries(_device=da), G.GeoSeries(_device=db)
    li, ri = G.spatial_join(sa, sb, how="inner", predicate="intersects")
    ll, rl = G.spatial_join(sa, sb, how="left", predicate="intersects")
    assert set(ll.tolist()) == set(range(2000)) and (rl >= 0).sum() == len(li) and len(ll) == len(li) + (2000 - len(set(li.tolist())))
