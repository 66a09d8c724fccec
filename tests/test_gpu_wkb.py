"""The GPU WKB codec (replaces the per-row / per-op decode of
geopolars/geopolars-geo/src/util.rs:27-37 and the re-encode of :11-24).  Checked against an independent
pure-Python struct reader / writer (tests/wkbutil.py): decoded buffers are bit-identical, re-encoded bytes are
byte-identical to canonical little-endian ISO WKB."""
import numpy as np
import pytest

import shapes
import wkbutil
from geopolars_b200 import GeoArrowArray, GeometryType
from geopolars_b200._lib import GeopolarsError, MismatchedGeometry

pytestmark = pytest.mark.gpu

T = dict(zip(shapes.KINDS, [GeometryType.POINT, GeometryType.MULTIPOINT, GeometryType.LINESTRING, GeometryType.MULTILINESTRING,
                            GeometryType.POLYGON, GeometryType.MULTIPOLYGON]))
WKB = {"point": 1, "linestring": 2, "polygon": 3, "multipoint": 4, "multilinestring": 5, "multipolygon": 6}


def same_buffers(dev, want: GeoArrowArray):
    got = dev.to_host()
    assert int(got.type) == int(want.type)
    assert np.array_equal(got.xy, want.xy, equal_nan=True)
    for name in ("geom_off", "part_off", "ring_off"):
        a, b = getattr(got, name), getattr(want, name)
        assert (a is None) == (b is None), name
        if a is not None:
            assert np.array_equal(a, b), name
    wv = np.ones(len(want), bool) if want.valid is None else want.valid
    gv = np.ones(len(want), bool) if got.valid is None else got.valid
    assert np.array_equal(gv, wv)


@pytest.mark.parametrize("kind", shapes.KINDS)
def test_decode_encode_every_type(ctx, kind):
    rng = np.random.default_rng(WKB[kind])
    n = 3000
    rows = shapes.random_rows(rng, kind, n)
    rows[7] = None
    if kind != "point":
        rows[9] = []  # empty geometry: valid, zero coordinates
    # input flavours: little endian, big endian, EWKB with SRID
    raw = [None if r is None else wkbutil.dump(WKB[kind], r, big_endian=(i % 3 == 1), srid=(4326 if i % 5 == 2 else None))
           for i, r in enumerate(rows)]
    off, data, valid = wkbutil.column(raw)
    dev = ctx.decode_wkb(data, off.astype(np.int32), valid)
    same_buffers(dev, GeoArrowArray.from_shapes(T[kind], rows))
    # canonical re-encoding, both offset widths
    canon = [None if r is None else wkbutil.dump(WKB[kind], r) for r in rows]
    coff, cdata, _ = wkbutil.column(canon)
    for width in (32, 64):
        eo, eb = dev.encode_wkb(width)
        assert np.array_equal(eo, coff) and np.array_equal(eb, cdata)
    # the validity bitmap alone marks nulls too (row bytes present but masked)
    masked = valid.copy()
    masked[11] = False
    dev2 = ctx.decode_wkb(data, off, masked)
    rows2 = list(rows)
    rows2[11] = None
    same_buffers(dev2, GeoArrowArray.from_shapes(T[kind], rows2))


@pytest.mark.parametrize("single,multi", [("point", "multipoint"), ("linestring", "multilinestring"), ("polygon", "multipolygon")])
def test_single_rows_are_promoted_when_mixed_with_multi_rows(ctx, single, multi):
    rng = np.random.default_rng(5)
    a, b = shapes.random_rows(rng, single, 400), shapes.random_rows(rng, multi, 400)
    raw, want = [], []
    for i in range(400):
        if i % 2:
            raw.append(wkbutil.dump(WKB[single], a[i])), want.append([a[i]])
        else:
            raw.append(wkbutil.dump(WKB[multi], b[i])), want.append(b[i])
    off, data, _ = wkbutil.column(raw)
    same_buffers(ctx.decode_wkb(data, off), GeoArrowArray.from_shapes(T[multi], want))


def test_sliced_offsets_and_large_binary(ctx):
    rng = np.random.default_rng(9)
    rows = shapes.random_rows(rng, "polygon", 500)
    off, data, _ = wkbutil.column([wkbutil.dump(3, r) for r in rows])
    dev = ctx.decode_wkb(data, off[100:301])  # int64 offsets starting at off[100]
    same_buffers(dev, GeoArrowArray.from_shapes(T["polygon"], rows[100:300]))


def test_malformed_input_is_reported_with_its_row(ctx):
    rng = np.random.default_rng(2)
    rows = shapes.random_rows(rng, "linestring", 50)
    raw = [wkbutil.dump(2, r) for r in rows]
    raw[17] = raw[17][:-5]  # truncated payload
    off, data, _ = wkbutil.column(raw)
    with pytest.raises(GeopolarsError, match="row 17"):
        ctx.decode_wkb(data, off)
    raw[17] = b"\x01" + (1002).to_bytes(4, "little") + raw[17][5:]  # LineString Z
    off, data, _ = wkbutil.column(raw)
    with pytest.raises(GeopolarsError, match="row 17"):
        ctx.decode_wkb(data, off)
    mixed = [wkbutil.dump(1, (1.0, 2.0)), wkbutil.dump(2, [(0, 0), (1, 1)])]
    off, data, _ = wkbutil.column(mixed)
    with pytest.raises(MismatchedGeometry):
        ctx.decode_wkb(data, off)
    off, data, _ = wkbutil.column(raw[:10])
    bad = off.copy()
    bad[4], bad[5] = bad[5], bad[4]  # offsets out of order (row 4 ends before it starts): reported, never dereferenced
    with pytest.raises(GeopolarsError, match="row 4"):
        ctx.decode_wkb(data, bad)
    empty = ctx.decode_wkb(np.zeros(0, np.uint8), np.zeros(1, np.int32))
    assert len(empty) == 0


def test_device_resident_columns(ctx):
    torch = pytest.importorskip("torch")
    n = 200_000
    xy = np.random.default_rng(4).uniform(-180, 180, (n, 2))
    # fixed-size Point rows built with numpy: 1 + 4 + 16 bytes
    rec = np.zeros((n, 21), np.uint8)
    rec[:, 0] = 1
    rec[:, 1] = 1
    rec[:, 5:] = xy.view(np.uint8).reshape(n, 16)
    data = torch.from_numpy(rec.reshape(-1)).cuda()
    off = torch.arange(0, 21 * (n + 1), 21, dtype=torch.int32).cuda()
    torch.cuda.synchronize()
    dev = ctx.decode_wkb(data.data_ptr(), off.data_ptr(), None, n=n, offset_width=32, device=True)
    assert np.array_equal(dev.to_host().xy, xy)
    out_b = torch.zeros(21 * n + 64, dtype=torch.uint8, device="cuda")
    out_o = torch.zeros(n + 1, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    used = dev.encode_wkb(64, device_out=(out_o.data_ptr(), out_b.data_ptr(), out_b.numel()))
    ctx.synchronize()
    assert used == 21 * n
    assert torch.equal(out_b[:used].cpu(), torch.from_numpy(rec.reshape(-1)))
    assert torch.equal(out_o.cpu(), torch.arange(0, 21 * (n + 1), 21, dtype=torch.int64))


def test_round_trip_at_size(ctx):
    """1 M polygons x 65 coordinates (1.05 GB of WKB): decode -> encode is the identity, the decoded
    coordinates are the ones written"""
    m, k = 1_000_000, 65
    rng = np.random.default_rng(8)
    ring = rng.uniform(0, 1000, (m, k, 2))
    ring[:, -1] = ring[:, 0]
    rec = np.zeros((m, 13 + 16 * k), np.uint8)
    rec[:, 0] = 1
    rec[:, 1] = 3
    rec[:, 5] = 1  # one ring
    rec[:, 9] = k
    rec[:, 13:] = ring.reshape(m, -1).view(np.uint8)
    data = rec.reshape(-1)
    off = np.arange(0, (m + 1) * rec.shape[1], rec.shape[1], dtype=np.int64)
    dev = ctx.decode_wkb(data, off)
    host = dev.to_host()
    assert np.array_equal(host.xy, ring.reshape(-1, 2))
    assert np.array_equal(host.ring_off, np.arange(0, (m + 1) * k, k)) and np.array_equal(host.geom_off, np.arange(m + 1))
    eo, eb = dev.encode_wkb(64)
    assert np.array_equal(eo, off) and np.array_equal(eb, data)
