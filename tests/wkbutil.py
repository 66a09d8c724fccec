"""Tiny pure-Python ISO-WKB reader used ONLY by tests to turn golden fixtures into oracle inputs
(independent of the library's C++ decoder, so that decoder gets checked against something)."""
import struct

import numpy as np


def parse(buf: bytes):
    """-> (type_code 1..6, nested python lists of (x, y))"""
    pos = 0

    def rd(fmt):
        nonlocal pos
        v = struct.unpack_from(fmt, buf, pos)
        pos += struct.calcsize(fmt)
        return v

    def geom():
        (bo,) = rd("<B")
        e = "<" if bo else ">"
        (t,) = rd(e + "I")
        if t == 1:
            return t, rd(e + "dd")
        if t == 2:
            (n,) = rd(e + "I")
            return t, [rd(e + "dd") for _ in range(n)]
        if t == 3:
            (nr,) = rd(e + "I")
            rings = []
            for _ in range(nr):
                (n,) = rd(e + "I")
                rings.append([rd(e + "dd") for _ in range(n)])
            return t, rings
        (n,) = rd(e + "I")
        return t, [geom()[1] for _ in range(n)]

    return geom()


def column_to_shapes(offsets: np.ndarray, data: np.ndarray):
    """WKB column -> (reference GeometryType code, shapes for GeoArrowArray.from_shapes), promoting
    Polygon rows to MultiPolygon when both occur (same rule as gpl_array_from_wkb)."""
    rows = [parse(bytes(data[offsets[i] : offsets[i + 1]])) for i in range(len(offsets) - 1)]
    types = {t for t, _ in rows}
    code = {1: 0, 2: 1, 3: 3, 4: 4, 5: 5, 6: 6}
    if types == {3, 6}:
        return 6, [([g] if t == 3 else g) for t, g in rows]
    assert len(types) == 1, types
    t = types.pop()
    return code[t], [g for _, g in rows]


def dump(wkb_type: int, shape, big_endian: bool = False, srid=None) -> bytes:
    """nested python lists -> ISO WKB (optionally big endian / with the EWKB SRID flag) — test input writer"""
    e = ">" if big_endian else "<"

    def head(t, top=False):
        flag = 0x20000000 if (top and srid is not None) else 0
        b = struct.pack("<B", 0 if big_endian else 1) + struct.pack(e + "I", t | flag)
        if flag:
            b += struct.pack(e + "I", srid)
        return b

    def pts(seq):
        return b"".join(struct.pack(e + "dd", float(x), float(y)) for x, y in seq)

    def line(seq):
        return struct.pack(e + "I", len(seq)) + pts(seq)

    def poly(rings):
        return struct.pack(e + "I", len(rings)) + b"".join(line(r) for r in rings)

    if wkb_type == 1:
        return head(1, True) + pts([shape])
    if wkb_type == 2:
        return head(2, True) + line(shape)
    if wkb_type == 3:
        return head(3, True) + poly(shape)
    if wkb_type == 4:
        return head(4, True) + struct.pack(e + "I", len(shape)) + b"".join(head(1) + pts([p]) for p in shape)
    if wkb_type == 5:
        return head(5, True) + struct.pack(e + "I", len(shape)) + b"".join(head(2) + line(l) for l in shape)
    if wkb_type == 6:
        return head(6, True) + struct.pack(e + "I", len(shape)) + b"".join(head(3) + poly(p) for p in shape)
    raise ValueError(wkb_type)


def column(rows):
    """list of bytes|None -> (int32 offsets, uint8 data, bool valid)"""
    off = np.zeros(len(rows) + 1, dtype=np.int64)
    for i, r in enumerate(rows):
        off[i + 1] = off[i] + (0 if r is None else len(r))
    data = np.frombuffer(b"".join(r for r in rows if r is not None), dtype=np.uint8).copy()
    return off, data, np.array([r is not None for r in rows])
