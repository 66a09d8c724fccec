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


def _simple_ring(ring):
    """no two non-adjacent edges of the closed ring meet, adjacent ones only at their shared vertex (exact)"""
    from oracle import exact

    pts = ring[:-1]
    n = len(pts)
    if n < 3 or len(set(pts)) != n:
        return False
    for i in range(n):
        a0, a1 = pts[i], pts[(i + 1) % n]
        for j in range(i + 1, n):
            b0, b1 = pts[j], pts[(j + 1) % n]
            adjacent = j == i + 1 or (i == 0 and j == n - 1)
            if adjacent:
                shared = a1 if j == i + 1 else a0
                other_a = a0 if j == i + 1 else a1
                other_b = b1 if j == i + 1 else b0
                # collinear overlap of adjacent edges (a spike) makes the ring invalid
                if exact.orient_sign(other_a, shared, other_b) == 0 and exact._on_segment(other_b, other_a, shared) or \
                        exact.orient_sign(other_a, shared, other_b) == 0 and exact._on_segment(other_a, other_b, shared):
                    return False
                continue
            if exact.segments_intersect(a0, a1, b0, b1):
                return False
    return True


def valid_star(rng, cx, cy, r, n, q=1):
    """a VALID (simple) lattice polygon, star-shaped about (cx, cy); retries until the rounded ring is simple"""
    for _ in range(50):
        th = np.sort(rng.uniform(0, 2 * np.pi, n))
        rr = r * rng.uniform(0.5, 1.0, n)
        pts = [(round(float(cx + rr[i] * np.cos(th[i])), q), round(float(cy + rr[i] * np.sin(th[i])), q)) for i in range(n)]
        ring = [pts[0]]
        for p in pts[1:]:
            if p != ring[-1]:
                ring.append(p)
        if len(ring) > 1 and ring[0] == ring[-1]:
            ring.pop()
        ring = ring + [ring[0]]
        if len(ring) >= 4 and _simple_ring(ring):
            return ring
    s = round(float(r), q) or 1.0
    return [(cx, cy), (cx + s, cy), (cx + s, cy + s), (cx, cy + s), (cx, cy)]


def sq(x, y, s):
    return [(x, y), (x + s, y), (x + s, y + s), (x, y + s), (x, y)]


def contains_cases(rng, n_random=300):
    """(A, B) polygon pairs for (Multi)Polygon.contains(Polygon): crafted touching / equal / hole / notch cases plus random
    valid lattice polygons (B inside A, B made of A's vertices, B off to the side).  A entries are lists of rings."""
    A, B = [], []

    def add(a, b):
        A.append(a), B.append(b)

    big, hole = sq(0, 0, 10), sq(4, 4, 2)[::-1]
    notch = [(0, 0), (10, 0), (10, 10), (5, 4), (0, 10), (0, 0)]
    add([big], [sq(2, 2, 3)])                                   # strictly inside
    add([big], [big])                                           # equal
    add([big], [sq(0, 0, 5)])                                   # shares two edges
    add([big], [sq(5, 5, 10)])                                  # overlaps
    add([big, hole], [sq(3, 3, 4)])                             # covers the hole
    add([big, hole], [sq(4, 4, 2)])                             # IS the hole
    add([big, hole], [sq(1, 1, 2)])                             # inside, away from the hole
    add([big, hole], [sq(3, 3, 4), hole])                       # has the same hole
    add([big], [[(0, 0), (10, 0), (5, 5), (0, 0)]])             # triangle on the bottom edge
    add([big], [[(0, 0), (10, 10), (0, 10), (0, 0)]])           # half along the diagonal
    add([big], [[(5, 0), (10, 5), (5, 10), (0, 5), (5, 0)]])    # diamond touching all four edges
    add([notch], [[(0, 10), (5, 4), (10, 10), (10, 9), (5, 3), (0, 9), (0, 10)]])  # follows the notch from inside
    add([notch], [sq(1, 1, 8)[:2] + [(9, 8), (1, 8), (1, 1)]])  # spans the notch -> leaves A
    add([big], [sq(10, 0, 5)])                                  # touching from outside
    add([big], [sq(20, 20, 1)])                                 # far away
    add([sq(0, 0, 10)[::-1]], [sq(2, 2, 3)[::-1]])              # both clockwise
    add([big], [sq(2, 2, 3)[::-1]])                             # B clockwise
    add([[(0, 0), (10, 0), (10, 10), (0, 10)]], [[(2, 2), (5, 2), (5, 5), (2, 5)]])  # unclosed rings (Polygon::new closes)
    for k in range(n_random):
        a = valid_star(rng, 5, 5, rng.uniform(2, 5), int(rng.integers(4, 12)))
        mode = k % 4
        if mode == 0:
            b = valid_star(rng, 5, 5, rng.uniform(0.5, 3), int(rng.integers(3, 8)))
        elif mode == 1:  # a sub-polygon of A's vertices (shares vertices and possibly edges)
            idx = sorted(rng.choice(len(a) - 1, size=min(len(a) - 1, int(rng.integers(3, 6))), replace=False).tolist())
            b = [a[i] for i in idx] + [a[idx[0]]]
            if not _simple_ring(b):
                continue
        elif mode == 2:  # one edge of A and the centre
            i = int(rng.integers(0, len(a) - 2))
            b = [a[i], a[i + 1], (5.0, 5.0), a[i]]
            if not _simple_ring(b):
                continue
        else:
            b = valid_star(rng, rng.uniform(3, 7), rng.uniform(3, 7), rng.uniform(0.5, 2), int(rng.integers(3, 7)))
        add([a], [b])
    return A, B


def multi_contains_cases():
    """MultiPolygon A (members touching in points only, or disjoint) x Polygon B"""
    left, right = sq(0, 0, 10), sq(10, 10, 10)  # touch at the corner (10, 10)
    A, B = [], []
    A.append([[left], [right]]), B.append([sq(2, 2, 3)])                       # inside member 0
    A.append([[left], [right]]), B.append([sq(12, 12, 3)])                     # inside member 1
    A.append([[left], [right]]), B.append([sq(8, 8, 4)])                       # across the touching corner: leaves A
    A.append([[left], [right]]), B.append([[(5, 5), (10, 10), (5, 10), (5, 5)]])   # inside member 0, touching the shared corner
    A.append([[left], [sq(20, 0, 5)]]), B.append([sq(21, 1, 2)])               # disjoint members
    A.append([[left], [sq(20, 0, 5)]]), B.append([sq(12, 1, 2)])               # in the gap
    A.append([]), B.append([sq(0, 0, 1)])                                      # empty multipolygon
    return A, B
