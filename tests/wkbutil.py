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
