"""An executable model of the FP32 filter of the points-in-polygons kernel (geopolars_b200/csrc/k_pip.cu:
fast_edge_rule / fast_walk / k_fast_fill) in numpy float32 arithmetic — same operations, same roundings
(no FMA: the library is compiled with -fmad=false) — refereed by exact rational arithmetic.

What it establishes, independently of any GPU: whenever the filter returns a DEFINITE answer for a (point,
polygon) pair, that answer is geo's `Polygon::contains` (interior only) — including the two one-sided edge
lists (the left one stored mirrored), the y-buckets, and points placed on / next to edges and vertices; and
that it answers definitely for almost all random points (the rest goes to the exact kernel)."""
from fractions import Fraction as F

import numpy as np
import pytest

from oracle import exact

f32 = np.float32


def rd(x):  # __double2float_rd
    y = f32(x)
    return y if float(y) <= x else np.nextafter(y, f32(-np.inf))


def ru(x):  # __double2float_ru
    y = f32(x)
    return y if float(y) >= x else np.nextafter(y, f32(np.inf))


class FastPart:
    """build side: what k_part_headers / k_buckets / k_fast_fill store for one plain ring"""

    def __init__(self, ring, slots_x100=300):
        self.ring = ring  # closed, list of (x, y) doubles
        xs, ys = [c[0] for c in ring], [c[1] for c in ring]
        self.xmin, self.xmax, self.ymin, self.ymax = min(xs), max(xs), min(ys), max(ys)
        self.xminf, self.xmaxf, self.yminf = rd(self.xmin), ru(self.xmax), rd(self.ymin)
        n_slots = len(ring)
        self.nb = max(1, (n_slots * 100 + slots_x100 - 1) // slots_x100)
        h_ext = self.ymax - float(self.yminf)
        self.inv_hf = f32(self.nb / h_ext) if h_ext > 0 else f32(0)
        self.xm = 0.5 * (float(self.xminf) + float(self.xmaxf))
        ox, mx, oy = float(self.xminf), float(self.xmaxf), float(self.yminf)
        self.lists = {}
        for s, e in zip(ring[:-1], ring[1:]):
            ylo, yhi = min(s[1], e[1]), max(s[1], e[1])
            for b in range(self.bucket(ylo), self.bucket(yhi) + 1):
                if max(s[0], e[0]) >= self.xm:
                    self.lists.setdefault((b, 0), []).append((f32(s[0] - ox), f32(s[1] - oy), f32(e[0] - ox), f32(e[1] - oy)))
                if min(s[0], e[0]) <= self.xm:
                    self.lists.setdefault((b, 1), []).append((f32(mx - s[0]), f32(s[1] - oy), f32(mx - e[0]), f32(e[1] - oy)))

    def bucket(self, y):  # mono_index: double arithmetic on float-valued parameters
        t = np.floor((y - float(self.yminf)) * float(self.inv_hf))
        return 0 if not t > 0 else int(min(t, self.nb - 1))

    def query(self, px, py):
        """fast_walk: returns True / False (definite) or None (undecided -> exact kernel)"""
        xlo, xhi = float(self.xminf), float(self.xmaxf)
        if not (xlo <= px <= xhi):
            return False
        yrel = py - float(self.yminf)
        t = yrel * float(self.inv_hf)
        if t < 0.0 or not t < self.nb + 1:
            return False
        b = min(int(t), self.nb - 1)
        right = px >= 0.5 * (xlo + xhi)
        qx, qy = f32(px - xlo if right else xhi - px), f32(yrel)
        height = f32(self.nb) / self.inv_hf if self.inv_hf > 0 else f32(0)
        R = max(self.xmaxf - self.xminf, height)
        eta = f32(9.5367431640625e-07) * R
        B = f32(1.01) * (f32(8.5) * eta * R + f32(1.9073486328125e-06) * R * R)
        wn, und = 0, False
        for ex, ey, ez, ew in self.lists.get((b, 0 if right else 1), []):
            u, w, z, v = ex - qx, ey - qy, ez - qx, ew - qy
            cert_y = abs(w) > eta and abs(v) > eta
            wl, vl = w < 0, v < 0
            act = wl != vl
            det = u * v - w * z  # float32 products and difference, three roundings
            und = und or (not cert_y) or (act and not abs(det) > B)
            wn += int(act and wl and det > 0) - int(act and (not wl) and det < 0)
        return None if und else wn != 0


def star(rng, cx, cy, n, r0, r1):
    th = 2 * np.pi * np.arange(n) / n
    r = rng.uniform(r0, r1, n)
    pts = [(float(cx + r[i] * np.cos(th[i])), float(cy + r[i] * np.sin(th[i]))) for i in range(n)]
    return pts + [pts[0]]


@pytest.mark.parametrize("seed,n,cx,cy", [(1, 64, 505.0, 495.0), (2, 17, -3.25, 1e4), (3, 200, 0.0, 0.0), (4, 5, 123456.789, -98765.4321)])
def test_definite_answers_of_the_fp32_filter_are_exact(seed, n, cx, cy):
    rng = np.random.default_rng(seed)
    ring = star(rng, cx, cy, n, 2.0, 4.8)
    part = FastPart(ring)
    pts = [(float(x), float(y)) for x, y in zip(rng.uniform(part.xmin - 0.5, part.xmax + 0.5, 1500), rng.uniform(part.ymin - 0.5, part.ymax + 0.5, 1500))]
    # adversarial: vertices, edge midpoints, points a few float/double ulps off them, vertex ordinates
    for (sx, sy), (ex, ey) in zip(ring[:-1], ring[1:]):
        mx_, my_ = 0.5 * (sx + ex), 0.5 * (sy + ey)
        pts += [(sx, sy), (mx_, my_), (np.nextafter(mx_, np.inf), my_), (mx_ + 1e-6, my_), (mx_ - 3e-7, my_ + 2e-7),
                (cx, sy), (cx, np.nextafter(sy, np.inf)), (cx, sy - 4e-7), (sx + 1e-5, sy), (part.xm, my_), (np.nextafter(part.xm, -np.inf), my_)]
    definite = wrong = 0
    for px, py in pts:
        got = part.query(px, py)
        if got is None:
            continue
        definite += 1
        want = exact.polygon_contains((px, py), [ring])
        wrong += got != want
    assert wrong == 0
    assert definite > 0.9 * 1500  # the filter decides nearly every random point on its own


def test_points_on_the_boundary_are_never_decided_by_the_filter():
    sq = [(0.0, 0.0), (8.0, 0.0), (8.0, 8.0), (0.0, 8.0), (0.0, 0.0)]
    part = FastPart(sq)
    for p in [(0.0, 4.0), (8.0, 1.0), (3.0, 0.0), (5.0, 8.0), (0.0, 0.0), (8.0, 8.0)]:
        assert part.query(*p) is None  # boundary => not contained: only the exact kernel may say so
    assert part.query(4.0, 4.0) is True and part.query(9.0, 4.0) is False and part.query(1e-3, 7.999) is True
