"""Random small geometries on a coarse lattice (so touching / collinear / shared-vertex cases occur) for the
row-wise predicate tests."""
import numpy as np


def star_ring(rng, cx, cy, r, n, q=1):
    th = np.sort(rng.uniform(0, 2 * np.pi, n))
    rr = r * rng.uniform(0.4, 1.0, n)
    pts = [(round(cx + rr[i] * np.cos(th[i]), q), round(cy + rr[i] * np.sin(th[i]), q)) for i in range(n)]
    return pts + [pts[0]]


def random_polygon(rng, span=10.0, hole=False, q=1):
    cx, cy = rng.uniform(0, span, 2)
    ext = star_ring(rng, cx, cy, rng.uniform(0.3, 4), int(rng.integers(3, 10)), q)
    rings = [ext]
    if hole:
        m = np.mean(np.array(ext[:-1]), 0)
        rings.append(star_ring(rng, m[0], m[1], 0.25, int(rng.integers(3, 6)), 2)[::-1])
    return rings


def random_linestring(rng, span=10.0, q=1):
    n = int(rng.integers(2, 8))
    p = rng.uniform(0, span, 2)
    out = []
    for _ in range(n):
        out.append((round(p[0], q), round(p[1], q)))
        p = p + rng.uniform(-2, 2, 2)
    return out


def random_point(rng, span=10.0, q=1):
    return (round(rng.uniform(0, span), q), round(rng.uniform(0, span), q))


def random_rows(rng, kind, n, span=10.0):
    """kind: one of point, multipoint, linestring, multilinestring, polygon, multipolygon -> list of shapes in the
    nesting GeoArrowArray.from_shapes expects"""
    rows = []
    for i in range(n):
        if kind == "point":
            rows.append(random_point(rng, span))
        elif kind == "multipoint":
            rows.append([random_point(rng, span) for _ in range(int(rng.integers(0, 4)))])
        elif kind == "linestring":
            rows.append(random_linestring(rng, span))
        elif kind == "multilinestring":
            rows.append([random_linestring(rng, span) for _ in range(int(rng.integers(0, 3)))])
        elif kind == "polygon":
            rows.append(random_polygon(rng, span, hole=(i % 3 == 0)))
        elif kind == "multipolygon":
            rows.append([random_polygon(rng, span, hole=(i % 4 == 0)) for _ in range(int(rng.integers(0, 3)))])
        else:
            raise ValueError(kind)
    return rows


KINDS = ["point", "multipoint", "linestring", "multilinestring", "polygon", "multipolygon"]


def first_vertices(kind, row):
    """some coordinates of a row (to plant touching cases)"""
    if row is None:
        return []
    if kind == "point":
        return [row]
    if kind in ("multipoint", "linestring"):
        return list(row)
    if kind in ("multilinestring", "polygon"):
        return [c for part in row for c in part]
    return [c for poly in row for ring in poly for c in ring]


def plant_touching(rng, kind_a, rows_a, kind_b, rows_b, every=4):
    """make every `every`-th row of A share a vertex (or a segment midpoint) with the row of B"""
    out = list(rows_a)
    for i in range(0, len(out), every):
        vs = first_vertices(kind_b, rows_b[i])
        if not vs:
            continue
        j = int(rng.integers(0, len(vs)))
        v = vs[j]
        if i % (2 * every) == 0 and j + 1 < len(vs):  # lattice midpoint: exactly on the segment when representable
            w = vs[j + 1]
            v = ((v[0] + w[0]) / 2, (v[1] + w[1]) / 2)
        if kind_a == "point":
            out[i] = v
        elif kind_a == "multipoint":
            out[i] = list(out[i]) + [v]
        elif kind_a == "linestring":
            out[i] = [v] + list(out[i])
        elif kind_a == "multilinestring":
            out[i] = [[v] + random_linestring(rng)] + list(out[i])
    return out
