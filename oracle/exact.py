"""Exact-rational referee for the boolean / index outputs — TEST INFRASTRUCTURE ONLY.

Independent of oracle/geo_oracle.c: it evaluates the *mathematical* definitions with
fractions.Fraction (every f64 is an exact rational), using formulations different from the
winding/orient2d ones the oracle restates, so agreement is evidence and not tautology.
Pure-Python loops: small cases only.
"""
from __future__ import annotations

from fractions import Fraction as F
from typing import List, Sequence, Tuple

Coord = Tuple[float, float]


def orient_sign(a: Coord, b: Coord, c: Coord) -> int:
    """sign of det [[ax-cx, ay-cy],[bx-cx, by-cy]] computed exactly."""
    d = (F(a[0]) - F(c[0])) * (F(b[1]) - F(c[1])) - (F(a[1]) - F(c[1])) * (F(b[0]) - F(c[0]))
    return (d > 0) - (d < 0)


def _on_segment(p: Coord, s: Coord, e: Coord) -> bool:
    px, py, sx, sy, ex, ey = map(F, (*p, *s, *e))
    if (sx, sy) == (ex, ey):
        return (px, py) == (sx, sy)
    cross = (ex - sx) * (py - sy) - (ey - sy) * (px - sx)
    if cross != 0:
        return False
    dot = (px - sx) * (ex - sx) + (py - sy) * (ey - sy)
    return 0 <= dot <= (ex - sx) ** 2 + (ey - sy) ** 2


def ring_position(p: Coord, ring: Sequence[Coord]) -> int:
    """0 outside / 1 boundary / 2 inside (non-zero winding) for a closed ring, by exact ray casting."""
    n = len(ring)
    if n == 0:
        return 0
    if n == 1:
        return 1 if tuple(ring[0]) == tuple(p) else 0
    pts = list(ring)
    if tuple(pts[0]) != tuple(pts[-1]):
        pts.append(pts[0])
    for s, e in zip(pts[:-1], pts[1:]):
        if _on_segment(p, s, e):
            return 1
    px, py = F(p[0]), F(p[1])
    wn = 0
    for s, e in zip(pts[:-1], pts[1:]):
        sx, sy, ex, ey = map(F, (*s, *e))
        if sy <= py < ey:  # upward
            xint = sx + (py - sy) * (ex - sx) / (ey - sy)
            if xint > px:
                wn += 1
        elif ey <= py < sy:  # downward
            xint = sx + (py - sy) * (ex - sx) / (ey - sy)
            if xint > px:
                wn -= 1
    return 2 if wn != 0 else 0


def polygon_contains(p: Coord, rings: Sequence[Sequence[Coord]]) -> bool:
    """interior-only containment: inside exterior, strictly outside every hole."""
    if not rings or len(rings[0]) == 0:
        return False
    if ring_position(p, rings[0]) != 2:
        return False
    return all(ring_position(p, h) == 0 for h in rings[1:])


def segments_intersect(s0: Coord, e0: Coord, s1: Coord, e1: Coord) -> bool:
    ax, ay, bx, by, cx, cy, dx, dy = map(F, (*s0, *e0, *s1, *e1))
    rx, ry = bx - ax, by - ay
    qx, qy = dx - cx, dy - cy
    den = rx * qy - ry * qx
    wx, wy = cx - ax, cy - ay
    if den != 0:
        t = (wx * qy - wy * qx) / den
        u = (wx * ry - wy * rx) / den
        return 0 <= t <= 1 and 0 <= u <= 1
    # parallel or degenerate
    if (rx, ry) == (0, 0):
        return _on_segment(s0, s1, e1)
    if (qx, qy) == (0, 0):
        return _on_segment(s1, s0, e0)
    if wx * ry - wy * rx != 0:
        return False
    rr = rx * rx + ry * ry
    t0 = (wx * rx + wy * ry) / rr
    t1 = t0 + (qx * rx + qy * ry) / rr
    lo, hi = min(t0, t1), max(t0, t1)
    return hi >= 0 and lo <= 1


def linestrings_intersect(a: Sequence[Coord], b: Sequence[Coord]) -> bool:
    for i in range(len(a) - 1):
        for j in range(len(b) - 1):
            if segments_intersect(a[i], a[i + 1], b[j], b[j + 1]):
                return True
    return False


def convex_hull_vertices(pts: Sequence[Coord]) -> List[Coord]:
    """strictly convex hull, CCW, starting at the lexicographic minimum (Andrew's chain, exact)."""
    P = sorted(set((float(x), float(y)) for x, y in pts))
    if len(P) <= 2:
        return P

    def cross(o, a, b):
        return (F(a[0]) - F(o[0])) * (F(b[1]) - F(o[1])) - (F(a[1]) - F(o[1])) * (F(b[0]) - F(o[0]))

    lower: List[Coord] = []
    for p in P:
        while len(lower) >= 2 and cross(lower[-2], lower[-1], p) <= 0:
            lower.pop()
        lower.append(p)
    upper: List[Coord] = []
    for p in reversed(P):
        while len(upper) >= 2 and cross(upper[-2], upper[-1], p) <= 0:
            upper.pop()
        upper.append(p)
    return lower[:-1] + upper[:-1]


def closed_polygon_has(p: Coord, rings: Sequence[Sequence[Coord]]) -> bool:
    """p belongs to the closed region: not outside the exterior, not strictly inside a hole"""
    if not rings or len(rings[0]) == 0:
        return False
    if ring_position(p, rings[0]) == 0:
        return False
    return all(ring_position(p, h) != 2 for h in rings[1:])


def _closed(ring: Sequence[Coord]) -> List[Coord]:
    r = list(ring)
    if len(r) >= 2 and tuple(r[0]) != tuple(r[-1]):
        r.append(r[0])
    return r


def polygons_intersect(a: Sequence[Sequence[Coord]], b: Sequence[Sequence[Coord]]) -> bool:
    """set definition for VALID polygons (closed regions share a point): some boundary segments meet, or a
    vertex of one lies in the other's closed region."""
    ra, rb = [_closed(r) for r in a], [_closed(r) for r in b]
    if not ra or not rb or len(ra[0]) < 2 or len(rb[0]) < 2:
        return False
    for x in ra:
        for y in rb:
            if linestrings_intersect(x, y):
                return True
    return closed_polygon_has(rb[0][0], ra) or closed_polygon_has(ra[0][0], rb)


def linestring_intersects_polygon(ls: Sequence[Coord], rings: Sequence[Sequence[Coord]]) -> bool:
    rr = [_closed(r) for r in rings]
    if len(ls) < 2 or not rr:
        return False
    return any(linestrings_intersect(ls, r) for r in rr) or closed_polygon_has(ls[0], rr)


# ---- (Multi)Polygon contains Polygon: the closed-region definition on the exact arrangement -------------------
def _region_pos(p, members) -> int:
    """0 exterior / 1 boundary / 2 interior of the union of polygons `members` (each a list of rings); p may be rational"""
    boundary = False
    for rings in members:
        rings = [_closed(r) for r in rings if len(r) > 0]
        if not rings:
            continue
        pe = ring_position(p, rings[0])
        if pe == 0:
            continue
        if pe == 1:
            boundary = True
            continue
        in_hole = False
        for h in rings[1:]:
            ph = ring_position(p, h)
            if ph == 2:
                in_hole = True
                break
            if ph == 1:
                in_hole = boundary = True
                break
        if not in_hole:
            return 2
    return 1 if boundary else 0


def _edges(members):
    for rings in members:
        for r in rings:
            c = _closed(r)
            for s, e in zip(c[:-1], c[1:]):
                if tuple(s) != tuple(e):
                    yield s, e


def _cut_params(u, v, others):
    """parameters in [0,1] at which the segment u-v meets the segments `others` (crossings and overlap ends), exact"""
    ux, uy, vx, vy = map(F, (*u, *v))
    rx, ry = vx - ux, vy - uy
    rr = rx * rx + ry * ry
    ts = {F(0), F(1)}
    for s, e in others:
        sx, sy, ex, ey = map(F, (*s, *e))
        qx, qy = ex - sx, ey - sy
        den = rx * qy - ry * qx
        wx, wy = sx - ux, sy - uy
        if den != 0:
            t = (wx * qy - wy * qx) / den
            w = (wx * ry - wy * rx) / den
            if 0 <= t <= 1 and 0 <= w <= 1:
                ts.add(t)
        elif wx * ry - wy * rx == 0:  # collinear: the ends of the other segment
            for px, py in ((sx, sy), (ex, ey)):
                t = ((px - ux) * rx + (py - uy) * ry) / rr
                if 0 <= t <= 1:
                    ts.add(t)
    return sorted(ts)


def _pieces_midpoints(u, v, others):
    ts = _cut_params(u, v, others)
    ux, uy, vx, vy = map(F, (*u, *v))
    for a, b in zip(ts[:-1], ts[1:]):
        m = (a + b) / 2
        yield (ux + m * (vx - ux), uy + m * (vy - uy))


def region_contains_polygon(a_members, b_rings) -> bool:
    """closure(A) contains B and the interiors meet (DE-9IM T*****FF*) for VALID operands, on the exact arrangement:
    every piece of boundary(B) between consecutive meetings with boundary(A) has its midpoint in closure(A); no piece of
    boundary(A) has its midpoint strictly inside B; and a point just beside boundary(B), on B's interior side, is
    strictly inside A."""
    b_members = [b_rings]
    if not a_members or not b_rings or len(_closed(b_rings[0])) < 4:
        return False
    a_edges, b_edges = list(_edges(a_members)), list(_edges(b_members))
    if not a_edges or not b_edges:
        return False
    for u, v in b_edges:
        for m in _pieces_midpoints(u, v, a_edges):
            if _region_pos(m, a_members) == 0:
                return False
    for s, e in a_edges:
        for m in _pieces_midpoints(s, e, b_edges):
            if _region_pos(m, b_members) == 2:
                return False
    u, v = b_edges[0]
    (mx, my), = list(_pieces_midpoints(u, v, []))
    dx, dy = F(v[0]) - F(u[0]), F(v[1]) - F(u[1])
    tiny = F(1, 2 ** 400)
    for side in (1, -1):
        q = (mx - side * tiny * dy, my + side * tiny * dx)
        if _region_pos(q, b_members) == 2:
            return _region_pos(q, a_members) == 2
    return False
