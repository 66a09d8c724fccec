"""An executable model of the LEVEL-WISE quick hull of geopolars_b200/csrc/k_hull.cu (`k_hull_fast`) in Python — the same
float expressions for the arg-max (two rounded products, one rounded sum; no FMA: the library is compiled -fmad=false), exact
orientation signs for the partitions (what robust's orient2d returns) — compared with the CPU oracle's depth-first restatement of
geo's `quick_hull` (oracle/geo_oracle.c), ring for ring, coordinate for coordinate.

What it establishes, independently of any GPU: whenever the level-wise form does NOT flag a geometry (no tie between points with
different coordinates, no point strictly left of both child segments, finite input, min != max, >= 4 coordinates), processing all
calls of one recursion depth together yields exactly the ring geo's depth-first recursion emits — same vertices, same order, same
closing rule; and that the flag conditions are what separates the two (lattice inputs with exact ties must flag, random data
must not)."""
import struct

import numpy as np
import pytest

from geopolars_b200 import GeoArrowArray, GeometryType
from oracle import exact


def key_bits(d):
    """hull_ord: order-preserving unsigned key of a double"""
    (b,) = struct.unpack("<Q", struct.pack("<d", d))
    return (~b & 0xFFFFFFFFFFFFFFFF) if b >> 63 else (b | 0x8000000000000000)


def is_ccw(a, b, c):
    return exact.orient_sign(a, b, c) > 0


def level_wise_hull(pts):
    """pts: list of (x, y) doubles incl. the ring's closing duplicate.  Returns the hull ring, or None when the kernel would flag
    the geometry (and hand it to the order-exact kernel)."""
    n = len(pts)
    if n < 4 or not all(np.isfinite(p[0]) and np.isfinite(p[1]) for p in pts):
        return None
    mi = min(range(n), key=lambda i: (pts[i][0], pts[i][1], i))
    xi = min(range(n), key=lambda i: (-pts[i][0], -pts[i][1], i))
    mn, mx = pts[mi], pts[xi]
    if mn == mx:
        return None
    chain = [xi, mi]  # emission order: [.., max, .., min]
    live = []  # (point index, call)
    for i, q in enumerate(pts):
        s = exact.orient_sign(mx, mn, q)
        if s:
            live.append((i, 0 if s > 0 else 1))
    while live:
        m = len(chain)
        best = {}
        keys = []
        for i, sg in live:
            a, b, p = pts[chain[sg]], pts[chain[sg - 1 if sg else m - 1]], pts[i]
            ox, oy, dx, dy = a[1] - b[1], b[0] - a[0], p[0] - a[0], p[1] - a[1]
            k = key_bits(ox * dx + oy * dy)
            keys.append(k)
            best[sg] = max(best.get(sg, 0), k)
        far = {}
        for (i, sg), k in zip(live, keys):
            if k == best[sg]:
                if sg in far and pts[far[sg]] != pts[i]:
                    return None  # a different point with the same value: geo's slice order would decide
                far.setdefault(sg, i)
        # chain update: the far point of every live call goes in front of the call's `a`
        new_chain, new_index = [], {}
        for j in range(m):
            if j in far:
                new_chain.append(far[j])
            new_index[j] = len(new_chain)
            new_chain.append(chain[j])
        if len(new_chain) > 128:
            return None
        nxt = []
        for i, sg in live:
            a, b, f, p = pts[chain[sg]], pts[chain[sg - 1 if sg else m - 1]], pts[far[sg]], pts[i]
            if p == f:
                continue
            t1, t2 = is_ccw(f, b, p), is_ccw(a, f, p)
            if t1 and t2:
                return None
            if t1 or t2:
                nxt.append((i, new_index[sg] - 1 if t1 else new_index[sg]))
        chain, live = new_chain, nxt
    ring = [pts[j] for j in chain]
    if ring[0] != mn:
        ring.append(ring[0])
    return ring


def oracle_rings(og, conv, shapes):
    arr = GeoArrowArray.from_shapes(GeometryType.MULTIPOINT, shapes)
    off, hxy = og.convex_hull(conv(arr))
    return [list(map(tuple, hxy[off[i]:off[i + 1]].tolist())) for i in range(len(shapes))]


def test_key_encoding_is_order_preserving():
    vals = [-np.inf, -1e300, -1.5, -5e-324, -0.0, 0.0, 5e-324, 1.0, 1.0000000000000002, 1e300, np.inf]
    ks = [key_bits(v) for v in vals]
    assert ks == sorted(ks) and len(set(ks)) == len(ks) and min(ks) > 0


def test_random_rings_never_flag_and_match_geo(og, conv):
    """stars and blobs of 5..300 points at several magnitudes: the level-wise ring IS geo's ring"""
    rng = np.random.default_rng(5)
    shapes = []
    for k in range(300):
        n = int(rng.integers(5, 300))
        cx, cy, sc = [(0.0, 0.0, 1.0), (9.8e5, 1.9e5, 100.0), (-3e-7, 2e-7, 1e-9)][k % 3]
        if k % 2:
            th = np.sort(rng.uniform(0, 2 * np.pi, n))
            r = rng.uniform(0.2, 1.0, n)
            p = np.stack([cx + sc * r * np.cos(th), cy + sc * r * np.sin(th)], 1)
        else:
            p = np.stack([cx + sc * rng.normal(size=n), cy + sc * rng.normal(size=n)], 1)
        ring = p.tolist() + [p[0].tolist()]
        shapes.append([tuple(c) for c in ring])
    want = oracle_rings(og, conv, shapes)
    flagged = 0
    for s, w in zip(shapes, want):
        got = level_wise_hull(s)
        if got is None:
            flagged += 1
            continue
        assert got == w
    assert flagged == 0


def test_lattices_and_repeats_flag_or_match(og, conv):
    """integer lattices, small-integer clouds with repeats, points on a circle with collinear midpoints: exact ties abound.
    Whatever is not flagged must still be geo's ring; and the flag is raised by most of these inputs (they are the reason
    the order-exact kernel is kept)."""
    rng = np.random.default_rng(11)
    shapes = []
    for n in (3, 4, 5, 7, 9):
        g = np.stack(np.meshgrid(np.arange(n), np.arange(n)), -1).reshape(-1, 2).astype(float)
        for rep in range(6):
            shapes.append([tuple(c) for c in rng.permutation(g).tolist()])
    for k in range(60):
        m = int(rng.integers(4, 90))
        shapes.append([tuple(c) for c in rng.integers(-3, 4, size=(m, 2)).astype(float).tolist()])
    circ = [(5, 0), (4, 3), (3, 4), (0, 5), (-3, 4), (-4, 3), (-5, 0), (-4, -3), (-3, -4), (0, -5), (3, -4), (4, -3)]
    mids = [((a[0] + b[0]) / 2, (a[1] + b[1]) / 2) for a, b in zip(circ, circ[1:] + circ[:1])]
    for rep in range(10):
        shapes.append([tuple(c) for c in rng.permutation(np.array(circ + mids + circ, float)).tolist()])
    want = oracle_rings(og, conv, shapes)
    flagged = matched = 0
    for s, w in zip(shapes, want):
        got = level_wise_hull(s)
        if got is None:
            flagged += 1
        else:
            assert got == w
            matched += 1
    assert flagged > 20 and matched > 20, (flagged, matched)


def test_equal_coordinates_tying_are_harmless(og, conv):
    """every ring carries its closing duplicate; duplicated hull vertices tie with themselves for the maximum — no flag, same ring"""
    rng = np.random.default_rng(3)
    shapes = []
    for k in range(40):
        n = int(rng.integers(6, 40))
        th = np.sort(rng.uniform(0, 2 * np.pi, n))
        p = np.stack([np.cos(th) * rng.uniform(0.5, 1.0, n), np.sin(th) * rng.uniform(0.5, 1.0, n)], 1).tolist()
        dup = [p[int(j)] for j in rng.integers(0, n, 5)]
        ring = p + dup + [p[0]]
        shapes.append([tuple(c) for c in ring])
    want = oracle_rings(og, conv, shapes)
    for s, w in zip(shapes, want):
        assert level_wise_hull(s) == w
