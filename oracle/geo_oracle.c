/*
 * geo_oracle.c — CPU oracle for the GeoSeries hot path.  TEST INFRASTRUCTURE ONLY (see geo_oracle.h).
 *
 * PARITY: pinned for contains() by spatial_index.rs:432-484 only; everything else PARITY UNPINNED.
 * Each function cites (a) the reference call site whose behaviour it stands in for and (b) the
 * third-party algorithm it restates ("recalled": restated from the published crate, source absent).
 *
 * Build: gcc -O2 -std=c11 -ffp-contract=off -fopenmp -fPIC -shared  (see oracle/Makefile).
 * -ffp-contract=off mirrors Rust, which never contracts a*b+c into an FMA.
 */
#include "geo_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------- */
/* robust 1.1.0 `orient2d` (a port of Shewchuk's predicates.c) — recalled.                      */
/* Reference call sites: geo's RobustKernel, reached from spatial_index.rs:91-135 (contains /   */
/* intersects) and geoseries.rs:26 (convex_hull).                                               */
/* ------------------------------------------------------------------------------------------- */
#define OG_EPS 1.1102230246251565e-16 /* 2^-53 */
static const double SPLITTER = 134217729.0; /* 2^27 + 1 */
static const double RESULTERRBOUND = (3.0 + 8.0 * OG_EPS) * OG_EPS;
static const double CCWERRBOUND_A = (3.0 + 16.0 * OG_EPS) * OG_EPS;
static const double CCWERRBOUND_B = (2.0 + 12.0 * OG_EPS) * OG_EPS;
static const double CCWERRBOUND_C = (9.0 + 64.0 * OG_EPS) * OG_EPS * OG_EPS;

static int64_t g_adapt_calls = 0;

static inline void two_sum(double a, double b, double *x, double *y) {
    double s = a + b;
    double bvirt = s - a;
    double avirt = s - bvirt;
    double bround = b - bvirt;
    double around = a - avirt;
    *x = s;
    *y = around + bround;
}
static inline void fast_two_sum(double a, double b, double *x, double *y) {
    double s = a + b;
    double bvirt = s - a;
    *x = s;
    *y = b - bvirt;
}
static inline double two_diff_tail(double a, double b, double x) {
    double bvirt = a - x;
    double avirt = x + bvirt;
    double bround = bvirt - b;
    double around = a - avirt;
    return around + bround;
}
static inline void two_diff(double a, double b, double *x, double *y) {
    double d = a - b;
    *x = d;
    *y = two_diff_tail(a, b, d);
}
static inline void split(double a, double *hi, double *lo) {
    double c = SPLITTER * a;
    double abig = c - a;
    *hi = c - abig;
    *lo = a - *hi;
}
static inline void two_product(double a, double b, double *x, double *y) {
    double p = a * b;
    double ahi, alo, bhi, blo;
    split(a, &ahi, &alo);
    split(b, &bhi, &blo);
    double err1 = p - (ahi * bhi);
    double err2 = err1 - (alo * bhi);
    double err3 = err2 - (ahi * blo);
    *x = p;
    *y = (alo * blo) - err3;
}
/* (a1,a0) - (b1,b0) -> x[3..0] */
static inline void two_two_diff(double a1, double a0, double b1, double b0, double x[4]) {
    double i, j, z, t;
    /* Two_One_Diff(a1, a0, b0, j, z, x0) */
    two_diff(a0, b0, &i, &x[0]);
    two_sum(a1, i, &j, &z);
    /* Two_One_Diff(j, z, b1, x3, x2, x1) */
    two_diff(z, b1, &i, &x[1]);
    two_sum(j, i, &x[3], &t);
    x[2] = t;
}
static int fast_expansion_sum_zeroelim(int elen, const double *e, int flen, const double *f, double *h) {
    double Q, Qnew, hh;
    int eindex = 0, findex = 0, hindex = 0;
    double enow = e[0], fnow = f[0];
    if ((fnow > enow) == (fnow > -enow)) {
        Q = enow;
        ++eindex;
        enow = eindex < elen ? e[eindex] : 0.0;
    } else {
        Q = fnow;
        ++findex;
        fnow = findex < flen ? f[findex] : 0.0;
    }
    if ((eindex < elen) && (findex < flen)) {
        if ((fnow > enow) == (fnow > -enow)) {
            fast_two_sum(enow, Q, &Qnew, &hh);
            ++eindex;
            enow = eindex < elen ? e[eindex] : 0.0;
        } else {
            fast_two_sum(fnow, Q, &Qnew, &hh);
            ++findex;
            fnow = findex < flen ? f[findex] : 0.0;
        }
        Q = Qnew;
        if (hh != 0.0) h[hindex++] = hh;
        while ((eindex < elen) && (findex < flen)) {
            if ((fnow > enow) == (fnow > -enow)) {
                two_sum(Q, enow, &Qnew, &hh);
                ++eindex;
                enow = eindex < elen ? e[eindex] : 0.0;
            } else {
                two_sum(Q, fnow, &Qnew, &hh);
                ++findex;
                fnow = findex < flen ? f[findex] : 0.0;
            }
            Q = Qnew;
            if (hh != 0.0) h[hindex++] = hh;
        }
    }
    while (eindex < elen) {
        two_sum(Q, enow, &Qnew, &hh);
        ++eindex;
        enow = eindex < elen ? e[eindex] : 0.0;
        Q = Qnew;
        if (hh != 0.0) h[hindex++] = hh;
    }
    while (findex < flen) {
        two_sum(Q, fnow, &Qnew, &hh);
        ++findex;
        fnow = findex < flen ? f[findex] : 0.0;
        Q = Qnew;
        if (hh != 0.0) h[hindex++] = hh;
    }
    if ((Q != 0.0) || (hindex == 0)) h[hindex++] = Q;
    return hindex;
}

static double orient2d_adapt(double ax, double ay, double bx, double by, double cx, double cy, double detsum) {
    double acx = ax - cx, bcx = bx - cx, acy = ay - cy, bcy = by - cy;
    double detleft, detlefttail, detright, detrighttail;
    double B[4], u[4], C1[8], C2[12], D[16];
    two_product(acx, bcy, &detleft, &detlefttail);
    two_product(acy, bcx, &detright, &detrighttail);
    two_two_diff(detleft, detlefttail, detright, detrighttail, B);
    double det = B[0] + B[1] + B[2] + B[3];
    double errbound = CCWERRBOUND_B * detsum;
    if ((det >= errbound) || (-det >= errbound)) return det;

    double acxtail = two_diff_tail(ax, cx, acx);
    double bcxtail = two_diff_tail(bx, cx, bcx);
    double acytail = two_diff_tail(ay, cy, acy);
    double bcytail = two_diff_tail(by, cy, bcy);
    if ((acxtail == 0.0) && (acytail == 0.0) && (bcxtail == 0.0) && (bcytail == 0.0)) return det;

    errbound = CCWERRBOUND_C * detsum + RESULTERRBOUND * fabs(det);
    det += (acx * bcytail + bcy * acxtail) - (acy * bcxtail + bcx * acytail);
    if ((det >= errbound) || (-det >= errbound)) return det;

    double s1, s0, t1, t0;
    two_product(acxtail, bcy, &s1, &s0);
    two_product(acytail, bcx, &t1, &t0);
    two_two_diff(s1, s0, t1, t0, u);
    int c1len = fast_expansion_sum_zeroelim(4, B, 4, u, C1);

    two_product(acx, bcytail, &s1, &s0);
    two_product(acy, bcxtail, &t1, &t0);
    two_two_diff(s1, s0, t1, t0, u);
    int c2len = fast_expansion_sum_zeroelim(c1len, C1, 4, u, C2);

    two_product(acxtail, bcytail, &s1, &s0);
    two_product(acytail, bcxtail, &t1, &t0);
    two_two_diff(s1, s0, t1, t0, u);
    int dlen = fast_expansion_sum_zeroelim(c2len, C2, 4, u, D);
    return D[dlen - 1];
}

double og_orient2d(double ax, double ay, double bx, double by, double cx, double cy) {
    double detleft = (ax - cx) * (by - cy);
    double detright = (ay - cy) * (bx - cx);
    double det = detleft - detright;
    double detsum;
    if (detleft > 0.0) {
        if (detright <= 0.0) return det;
        detsum = detleft + detright;
    } else if (detleft < 0.0) {
        if (detright >= 0.0) return det;
        detsum = -detleft - detright;
    } else {
        return det;
    }
    double errbound = CCWERRBOUND_A * detsum;
    if ((det >= errbound) || (-det >= errbound)) return det;
#ifdef _OPENMP
#pragma omp atomic
#endif
    g_adapt_calls++;
    return orient2d_adapt(ax, ay, bx, by, cx, cy, detsum);
}
int64_t og_orient2d_adapt_calls(void) { return g_adapt_calls; }

/* geo Orientation: +1 CounterClockwise, -1 Clockwise, 0 Collinear */
static inline int orient_sign(const double *a, const double *b, const double *c) {
    double d = og_orient2d(a[0], a[1], b[0], b[1], c[0], c[1]);
    return (d > 0.0) - (d < 0.0);
}

static int resolve_threads(int threads) {
#ifdef _OPENMP
    if (threads <= 0) return omp_get_max_threads();
    return threads;
#else
    (void)threads;
    return 1;
#endif
}
int og_max_threads(void) { return resolve_threads(0); }

static inline int is_valid(const og_array *a, int64_t i) {
    return a->valid == NULL || ((a->valid[i >> 3] >> (i & 7)) & 1);
}

/* ------------------------------------------------------------------------------------------- */
/* affine — GeoSeries::affine_transform geoseries.rs:11-12 ; geo AffineTransform::apply recalled */
/* ------------------------------------------------------------------------------------------- */
void og_affine_transform(const double *xy, int64_t n, double a, double b, double xoff, double d, double e,
                         double yoff, double *out, int threads) {
    int nt = resolve_threads(threads);
    (void)nt;
#pragma omp parallel for num_threads(nt) schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        double x = xy[2 * i], y = xy[2 * i + 1];
        out[2 * i] = a * x + b * y + xoff;
        out[2 * i + 1] = d * x + e * y + yoff;
    }
}

/* ------------------------------------------------------------------------------------------- */
/* area — GeoSeries::area geoseries.rs:14-16 ; geo algorithm/area.rs recalled                     */
/* ------------------------------------------------------------------------------------------- */
static double twice_signed_ring_area(const double *xy, int64_t n) {
    if (n < 3) return 0.0;
    if (xy[0] != xy[2 * (n - 1)] || xy[1] != xy[2 * (n - 1) + 1]) return 0.0;
    double sx = xy[0], sy = xy[1];
    double tmp = 0.0;
    for (int64_t i = 0; i + 1 < n; ++i) {
        double x0 = xy[2 * i] - sx, y0 = xy[2 * i + 1] - sy;
        double x1 = xy[2 * i + 2] - sx, y1 = xy[2 * i + 3] - sy;
        tmp = tmp + (x0 * y1 - y0 * x1);
    }
    return tmp;
}
static inline double ring_area(const double *xy, int64_t n) { return twice_signed_ring_area(xy, n) / 2.0; }

/* Polygon::signed_area over rings [r0, r1) */
static double polygon_signed_area(const og_array *a, int64_t r0, int64_t r1) {
    if (r1 <= r0) return 0.0;
    const int64_t *ro = a->ring_off;
    double area = ring_area(a->xy + 2 * ro[r0], ro[r0 + 1] - ro[r0]);
    int neg = area < 0.0;
    double tot = fabs(area);
    for (int64_t r = r0 + 1; r < r1; ++r) tot = tot - fabs(ring_area(a->xy + 2 * ro[r], ro[r + 1] - ro[r]));
    return neg ? -tot : tot;
}

void og_area(const og_array *a, double *out, int threads) {
    int nt = resolve_threads(threads);
    (void)nt;
#pragma omp parallel for num_threads(nt) schedule(static)
    for (int64_t i = 0; i < a->n; ++i) {
        double v = 0.0;
        if (a->type == OG_POLYGON) {
            v = fabs(polygon_signed_area(a, a->geom_off[i], a->geom_off[i + 1]));
        } else if (a->type == OG_MULTIPOLYGON) {
            for (int64_t p = a->geom_off[i]; p < a->geom_off[i + 1]; ++p)
                v = v + fabs(polygon_signed_area(a, a->part_off[p], a->part_off[p + 1]));
        }
        out[i] = v;
    }
}

/* ------------------------------------------------------------------------------------------- */
/* centroid — GeoSeries::centroid geoseries.rs:18-21 ; geo algorithm/centroid.rs recalled         */
/* ------------------------------------------------------------------------------------------- */
typedef struct {
    int dim; /* -1 none (Empty), 0,1,2 */
    double w, ax, ay;
} wc_t;

static void wc_add(wc_t *s, wc_t b) { /* add_weighted_centroid + WeightedCentroid::add_assign */
    if (b.dim < 0) return;
    if (s->dim < 0) {
        *s = b;
        return;
    }
    if (s->dim < b.dim) {
        *s = b;
    } else if (s->dim == b.dim) {
        s->ax = s->ax + b.ax;
        s->ay = s->ay + b.ay;
        s->w = s->w + b.w;
    }
}
static void wc_add_centroid(wc_t *s, int dim, double cx, double cy, double w) {
    wc_t b = {dim, w, cx * w, cy * w};
    wc_add(s, b);
}
static void wc_add_line(wc_t *s, const double *p0, const double *p1) {
    if (p0[0] == p1[0] && p0[1] == p1[1]) {
        wc_add_centroid(s, 0, p0[0], p0[1], 1.0);
    } else {
        double len = hypot(p1[0] - p0[0], p1[1] - p0[1]);
        wc_add_centroid(s, 1, (p1[0] + p0[0]) / 2.0, (p1[1] + p0[1]) / 2.0, len);
    }
}
static void wc_add_line_string(wc_t *s, const double *xy, int64_t n) {
    if (s->dim > 1) return;
    if (n == 1) {
        wc_add_centroid(s, 0, xy[0], xy[1], 1.0);
        return;
    }
    for (int64_t i = 0; i + 1 < n; ++i) wc_add_line(s, xy + 2 * i, xy + 2 * i + 2);
}
static void wc_add_ring(wc_t *s, const double *xy, int64_t n) {
    double area = ring_area(xy, n);
    if (area == 0.0) {
        if (n == 0) return;
        if (n == 1) {
            wc_add_centroid(s, 0, xy[0], xy[1], 1.0);
            return;
        }
        wc_add_line_string(s, xy, n);
        return;
    }
    double sx = xy[0], sy = xy[1];
    double accx = 0.0, accy = 0.0;
    for (int64_t i = 0; i + 1 < n; ++i) {
        double x0 = xy[2 * i] - sx, y0 = xy[2 * i + 1] - sy;
        double x1 = xy[2 * i + 2] - sx, y1 = xy[2 * i + 3] - sy;
        double tmp = x0 * y1 - y0 * x1;
        accx = accx + (x1 + x0) * tmp;
        accy = accy + (y1 + y0) * tmp;
    }
    double cx = accx / (6.0 * area) + sx;
    double cy = accy / (6.0 * area) + sy;
    wc_add_centroid(s, 2, cx, cy, fabs(area));
}
static void wc_add_polygon(wc_t *s, const og_array *a, int64_t r0, int64_t r1) {
    if (r1 <= r0) return;
    const int64_t *ro = a->ring_off;
    wc_t ext = {-1, 0, 0, 0}, itr = {-1, 0, 0, 0};
    wc_add_ring(&ext, a->xy + 2 * ro[r0], ro[r0 + 1] - ro[r0]);
    for (int64_t r = r0 + 1; r < r1; ++r) wc_add_ring(&itr, a->xy + 2 * ro[r], ro[r + 1] - ro[r]);
    if (ext.dim >= 0) {
        wc_t poly = ext;
        if (itr.dim >= 0) {
            if (poly.dim == itr.dim) { /* WeightedCentroid::sub_assign */
                poly.ax = poly.ax - itr.ax;
                poly.ay = poly.ay - itr.ay;
                poly.w = poly.w - itr.w;
            }
            if (poly.w == 0.0) {
                wc_add_line_string(s, a->xy + 2 * ro[r0], ro[r0 + 1] - ro[r0]);
                return;
            }
        }
        wc_add(s, poly);
    }
}

void og_centroid(const og_array *a, double *out_xy, uint8_t *out_valid, int threads) {
    int nt = resolve_threads(threads);
    (void)nt;
#pragma omp parallel for num_threads(nt) schedule(static)
    for (int64_t i = 0; i < a->n; ++i) {
        wc_t s = {-1, 0, 0, 0};
        int ok = is_valid(a, i);
        if (ok) {
            switch (a->type) {
            case OG_POINT:
                /* geo: Point::centroid == self; an empty (NaN) point has no centroid */
                if (!(isnan(a->xy[2 * i]) && isnan(a->xy[2 * i + 1]))) wc_add_centroid(&s, 0, a->xy[2 * i], a->xy[2 * i + 1], 1.0);
                break;
            case OG_MULTIPOINT:
                for (int64_t c = a->geom_off[i]; c < a->geom_off[i + 1]; ++c)
                    wc_add_centroid(&s, 0, a->xy[2 * c], a->xy[2 * c + 1], 1.0);
                break;
            case OG_LINESTRING:
                wc_add_line_string(&s, a->xy + 2 * a->geom_off[i], a->geom_off[i + 1] - a->geom_off[i]);
                break;
            case OG_MULTILINESTRING:
                for (int64_t l = a->geom_off[i]; l < a->geom_off[i + 1]; ++l)
                    wc_add_line_string(&s, a->xy + 2 * a->ring_off[l], a->ring_off[l + 1] - a->ring_off[l]);
                break;
            case OG_POLYGON:
                wc_add_polygon(&s, a, a->geom_off[i], a->geom_off[i + 1]);
                break;
            case OG_MULTIPOLYGON:
                for (int64_t p = a->geom_off[i]; p < a->geom_off[i + 1]; ++p)
                    wc_add_polygon(&s, a, a->part_off[p], a->part_off[p + 1]);
                break;
            default:
                break;
            }
        }
        if (s.dim >= 0) {
            if (a->type == OG_POINT) { /* exact pass-through, no w*x/w round trip needed: 1.0 weights are exact */
                out_xy[2 * i] = s.ax / s.w;
                out_xy[2 * i + 1] = s.ay / s.w;
            } else {
                out_xy[2 * i] = s.ax / s.w;
                out_xy[2 * i + 1] = s.ay / s.w;
            }
            out_valid[i] = 1;
        } else {
            out_xy[2 * i] = NAN;
            out_xy[2 * i + 1] = NAN;
            out_valid[i] = 0;
        }
    }
}

/* ------------------------------------------------------------------------------------------- */
/* envelope / length — geoseries.rs:28-41 ; geo bounding_rect.rs, euclidean_length.rs recalled    */
/* ------------------------------------------------------------------------------------------- */
typedef struct {
    int has;
    double x0, y0, x1, y1;
} bb_t;
static inline void bb_add(bb_t *b, const double *xy, int64_t n) {
    for (int64_t i = 0; i < n; ++i) {
        double px = xy[2 * i], py = xy[2 * i + 1];
        if (!b->has) {
            b->has = 1;
            b->x0 = b->x1 = px;
            b->y0 = b->y1 = py;
            continue;
        }
        if (px > b->x1) b->x1 = px;
        else if (px < b->x0) b->x0 = px;
        if (py > b->y1) b->y1 = py;
        else if (py < b->y0) b->y0 = py;
    }
}
static bb_t geom_bbox(const og_array *a, int64_t i) {
    bb_t b = {0, 0, 0, 0, 0};
    const int64_t *go = a->geom_off, *ro = a->ring_off, *po = a->part_off;
    switch (a->type) {
    case OG_POINT:
        if (!(isnan(a->xy[2 * i]) && isnan(a->xy[2 * i + 1]))) bb_add(&b, a->xy + 2 * i, 1);
        break;
    case OG_LINESTRING:
    case OG_MULTIPOINT:
        bb_add(&b, a->xy + 2 * go[i], go[i + 1] - go[i]);
        break;
    case OG_MULTILINESTRING:
        for (int64_t l = go[i]; l < go[i + 1]; ++l) bb_add(&b, a->xy + 2 * ro[l], ro[l + 1] - ro[l]);
        break;
    case OG_POLYGON: /* exterior only */
        if (go[i + 1] > go[i]) bb_add(&b, a->xy + 2 * ro[go[i]], ro[go[i] + 1] - ro[go[i]]);
        break;
    case OG_MULTIPOLYGON:
        for (int64_t p = go[i]; p < go[i + 1]; ++p)
            if (po[p + 1] > po[p]) bb_add(&b, a->xy + 2 * ro[po[p]], ro[po[p] + 1] - ro[po[p]]);
        break;
    default:
        break;
    }
    return b;
}
void og_envelope(const og_array *a, double *out4, uint8_t *out_valid, int threads) {
    int nt = resolve_threads(threads);
    (void)nt;
#pragma omp parallel for num_threads(nt) schedule(static)
    for (int64_t i = 0; i < a->n; ++i) {
        bb_t b = {0, 0, 0, 0, 0};
        if (is_valid(a, i)) b = geom_bbox(a, i);
        out_valid[i] = (uint8_t)b.has;
        out4[4 * i] = b.has ? b.x0 : NAN;
        out4[4 * i + 1] = b.has ? b.y0 : NAN;
        out4[4 * i + 2] = b.has ? b.x1 : NAN;
        out4[4 * i + 3] = b.has ? b.y1 : NAN;
    }
}
static double ls_length(const double *xy, int64_t n) {
    double s = 0.0;
    for (int64_t i = 0; i + 1 < n; ++i) s = s + hypot(xy[2 * i + 2] - xy[2 * i], xy[2 * i + 3] - xy[2 * i + 1]);
    return s;
}
void og_euclidean_length(const og_array *a, double *out, int threads) {
    int nt = resolve_threads(threads);
    (void)nt;
    const int64_t *go = a->geom_off, *ro = a->ring_off, *po = a->part_off;
#pragma omp parallel for num_threads(nt) schedule(static)
    for (int64_t i = 0; i < a->n; ++i) {
        double v = 0.0;
        switch (a->type) {
        case OG_LINESTRING:
            v = ls_length(a->xy + 2 * go[i], go[i + 1] - go[i]);
            break;
        case OG_MULTILINESTRING:
            for (int64_t l = go[i]; l < go[i + 1]; ++l) v = v + ls_length(a->xy + 2 * ro[l], ro[l + 1] - ro[l]);
            break;
        case OG_POLYGON: /* geoseries.rs:35-41: "for Polygon it's the length of the exterior ring" */
            if (go[i + 1] > go[i]) v = ls_length(a->xy + 2 * ro[go[i]], ro[go[i] + 1] - ro[go[i]]);
            break;
        case OG_MULTIPOLYGON:
            for (int64_t p = go[i]; p < go[i + 1]; ++p)
                if (po[p + 1] > po[p]) v = v + ls_length(a->xy + 2 * ro[po[p]], ro[po[p] + 1] - ro[po[p]]);
            break;
        default:
            break;
        }
        out[i] = v;
    }
}

/* ------------------------------------------------------------------------------------------- */
/* contains — spatial_index.rs:89-96 `poly.contains(point)` ; geo coordinate_position.rs recalled */
/* ------------------------------------------------------------------------------------------- */
static inline int value_in_between(double v, double b1, double b2) {
    if (b1 < b2) return v >= b1 && v <= b2;
    return v >= b2 && v <= b1;
}
static inline int point_in_rect(const double *v, const double *b1, const double *b2) {
    return value_in_between(v[0], b1[0], b2[0]) && value_in_between(v[1], b1[1], b2[1]);
}
enum { POS_OUTSIDE = 0, POS_BOUNDARY = 1, POS_INSIDE = 2 };

/* coord_pos_relative_to_ring.  Rings in GeoArrow are explicitly closed; geo's Polygon::new closes an
 * open ring by appending the first coord, which `closing` emulates. */
static int coord_pos_ring(double px, double py, const double *xy, int64_t n) {
    if (n == 0) return POS_OUTSIDE;
    if (n == 1) return (px == xy[0] && py == xy[1]) ? POS_BOUNDARY : POS_OUTSIDE;
    int closing = !(xy[0] == xy[2 * (n - 1)] && xy[1] == xy[2 * (n - 1) + 1]);
    int64_t nseg = n - 1 + closing;
    int wn = 0;
    double p[2] = {px, py};
    for (int64_t i = 0; i < nseg; ++i) {
        const double *s = xy + 2 * i;
        const double *e = (i + 1 < n) ? xy + 2 * (i + 1) : xy;
        if (s[1] <= py) {
            if (e[1] >= py) {
                int o = orient_sign(s, e, p);
                if (o > 0 && e[1] != py) wn += 1;
                else if (o == 0 && value_in_between(px, s[0], e[0])) return POS_BOUNDARY;
            }
        } else if (e[1] <= py) {
            int o = orient_sign(s, e, p);
            if (o < 0) wn -= 1;
            else if (o == 0 && value_in_between(px, s[0], e[0])) return POS_BOUNDARY;
        }
    }
    return wn == 0 ? POS_OUTSIDE : POS_INSIDE;
}
/* Polygon::calculate_coordinate_position over rings [r0,r1) */
static void polygon_coord_pos(const og_array *a, int64_t r0, int64_t r1, double px, double py, int *is_inside,
                              int64_t *boundary_count) {
    const int64_t *ro = a->ring_off;
    if (r1 <= r0) return;
    if (ro[r0 + 1] - ro[r0] == 0) return;
    int pos = coord_pos_ring(px, py, a->xy + 2 * ro[r0], ro[r0 + 1] - ro[r0]);
    if (pos == POS_OUTSIDE) return;
    if (pos == POS_BOUNDARY) {
        *boundary_count += 1;
        return;
    }
    for (int64_t r = r0 + 1; r < r1; ++r) {
        int hp = coord_pos_ring(px, py, a->xy + 2 * ro[r], ro[r + 1] - ro[r]);
        if (hp == POS_BOUNDARY) {
            *boundary_count += 1;
            return;
        }
        if (hp == POS_INSIDE) return;
    }
    *is_inside = 1;
}
int og_coord_position(const og_array *a, int64_t i, double px, double py) {
    int inside = 0;
    int64_t bc = 0;
    if (a->type == OG_POLYGON) {
        polygon_coord_pos(a, a->geom_off[i], a->geom_off[i + 1], px, py, &inside, &bc);
    } else if (a->type == OG_MULTIPOLYGON) {
        for (int64_t p = a->geom_off[i]; p < a->geom_off[i + 1]; ++p)
            polygon_coord_pos(a, a->part_off[p], a->part_off[p + 1], px, py, &inside, &bc);
    }
    if (bc % 2 == 1) return POS_BOUNDARY;
    return inside ? POS_INSIDE : POS_OUTSIDE;
}
int og_contains_point(const og_array *a, int64_t i, double px, double py) {
    if (!is_valid(a, i)) return 0;
    if (a->type == OG_POLYGON) return og_coord_position(a, i, px, py) == POS_INSIDE;
    if (a->type == OG_MULTIPOLYGON) { /* MultiPolygon::contains(coord) = any polygon contains */
        for (int64_t p = a->geom_off[i]; p < a->geom_off[i + 1]; ++p) {
            int inside = 0;
            int64_t bc = 0;
            polygon_coord_pos(a, a->part_off[p], a->part_off[p + 1], px, py, &inside, &bc);
            if (bc % 2 == 0 && inside) return 1;
        }
    }
    return 0;
}

/* broadcast join.  The grid is only a candidate filter (closed-interval bbox overlap like rstar's
 * AABB test, spatial_index.rs:74-76, 289); the exact test decides. */
void og_contains_join(const og_array *polys, const double *pts, int64_t n_pts, int32_t *first_id, int32_t *count,
                      int use_grid, int threads) {
    int nt = resolve_threads(threads);
    (void)nt;
    int64_t m = polys->n;
    double *bb = (double *)malloc(sizeof(double) * 4 * (size_t)(m > 0 ? m : 1));
    uint8_t *bbv = (uint8_t *)malloc((size_t)(m > 0 ? m : 1));
    double gx0 = INFINITY, gy0 = INFINITY, gx1 = -INFINITY, gy1 = -INFINITY;
    for (int64_t j = 0; j < m; ++j) {
        bb_t b = {0, 0, 0, 0, 0};
        if (is_valid(polys, j)) b = geom_bbox(polys, j);
        bbv[j] = (uint8_t)b.has;
        bb[4 * j] = b.x0, bb[4 * j + 1] = b.y0, bb[4 * j + 2] = b.x1, bb[4 * j + 3] = b.y1;
        if (b.has) {
            if (b.x0 < gx0) gx0 = b.x0;
            if (b.y0 < gy0) gy0 = b.y0;
            if (b.x1 > gx1) gx1 = b.x1;
            if (b.y1 > gy1) gy1 = b.y1;
        }
    }
    int G = 1;
    int64_t *cell_start = NULL;
    int32_t *cell_items = NULL;
    double inv_w = 0, inv_h = 0;
    if (use_grid && m > 0 && gx1 >= gx0) {
        G = (int)ceil(sqrt((double)m));
        if (G < 1) G = 1;
        if (G > 2048) G = 2048;
        inv_w = (gx1 > gx0) ? G / (gx1 - gx0) : 0.0;
        inv_h = (gy1 > gy0) ? G / (gy1 - gy0) : 0.0;
        cell_start = (int64_t *)calloc((size_t)G * G + 1, sizeof(int64_t));
#define CELL(v, lo, inv) ({ double t_ = floor(((v) - (lo)) * (inv)); int c_ = t_ < 0 ? 0 : (t_ >= G ? G - 1 : (int)t_); c_; })
        for (int pass = 0; pass < 2; ++pass) {
            if (pass == 1) {
                int64_t acc = 0;
                for (int64_t c = 0; c < (int64_t)G * G; ++c) {
                    int64_t t = cell_start[c];
                    cell_start[c] = acc;
                    acc += t;
                }
                cell_start[(int64_t)G * G] = acc;
                cell_items = (int32_t *)malloc(sizeof(int32_t) * (size_t)(acc > 0 ? acc : 1));
            }
            int64_t *fill = NULL;
            if (pass == 1) {
                fill = (int64_t *)malloc(sizeof(int64_t) * (size_t)G * G);
                memcpy(fill, cell_start, sizeof(int64_t) * (size_t)G * G);
            }
            for (int64_t j = 0; j < m; ++j) {
                if (!bbv[j]) continue;
                int cx0 = CELL(bb[4 * j], gx0, inv_w), cx1 = CELL(bb[4 * j + 2], gx0, inv_w);
                int cy0 = CELL(bb[4 * j + 1], gy0, inv_h), cy1 = CELL(bb[4 * j + 3], gy0, inv_h);
                for (int cy = cy0; cy <= cy1; ++cy)
                    for (int cx = cx0; cx <= cx1; ++cx) {
                        int64_t c = (int64_t)cy * G + cx;
                        if (pass == 0) cell_start[c]++;
                        else cell_items[fill[c]++] = (int32_t)j;
                    }
            }
            free(fill);
        }
    }
#pragma omp parallel for num_threads(nt) schedule(static)
    for (int64_t p = 0; p < n_pts; ++p) {
        double px = pts[2 * p], py = pts[2 * p + 1];
        int32_t first = -1, cnt = 0;
        if (cell_start) {
            if (px >= gx0 && px <= gx1 && py >= gy0 && py <= gy1) {
                int cx = CELL(px, gx0, inv_w), cy = CELL(py, gy0, inv_h);
                int64_t c = (int64_t)cy * G + cx;
                for (int64_t k = cell_start[c]; k < cell_start[c + 1]; ++k) {
                    int32_t j = cell_items[k];
                    if (px < bb[4 * j] || px > bb[4 * j + 2] || py < bb[4 * j + 1] || py > bb[4 * j + 3]) continue;
                    if (og_contains_point(polys, j, px, py)) {
                        if (first < 0 || j < first) first = j;
                        cnt++;
                    }
                }
            }
        } else {
            for (int64_t j = 0; j < m; ++j) {
                if (!bbv[j]) continue;
                if (og_contains_point(polys, j, px, py)) {
                    if (first < 0) first = (int32_t)j;
                    cnt++;
                }
            }
        }
        first_id[p] = first;
        if (count) count[p] = cnt;
    }
    free(bb);
    free(bbv);
    free(cell_start);
    free(cell_items);
}

/* ------------------------------------------------------------------------------------------- */
/* intersects — spatial_index.rs:102-104 semantics donor ; geo intersects/{line,line_string}.rs   */
/* ------------------------------------------------------------------------------------------- */
static int line_intersects_coord(const double *s, const double *e, const double *c) {
    return orient_sign(s, e, c) == 0 && point_in_rect(c, s, e);
}
/* impl Intersects<Line> for Line: `self`=(s0,e0), `line`=(s1,e1) */
static int line_intersects_line(const double *s0, const double *e0, const double *s1, const double *e1) {
    if (s0[0] == e0[0] && s0[1] == e0[1]) return line_intersects_coord(s1, e1, s0);
    int c11 = orient_sign(s0, e0, s1);
    int c12 = orient_sign(s0, e0, e1);
    if (c11 != c12) {
        int c21 = orient_sign(s1, e1, s0);
        int c22 = orient_sign(s1, e1, e0);
        return c21 != c22;
    } else if (c11 == 0) {
        return point_in_rect(s1, s0, e0) || point_in_rect(e1, s0, e0) || point_in_rect(e0, s1, e1) ||
               point_in_rect(e0, s1, e1);
    }
    return 0;
}
static int ls_bbox(const double *xy, int64_t n, double *b) {
    bb_t t = {0, 0, 0, 0, 0};
    bb_add(&t, xy, n);
    b[0] = t.x0, b[1] = t.y0, b[2] = t.x1, b[3] = t.y1;
    return t.has;
}
static int ls_intersects_ls(const double *a, int64_t na, const double *b, int64_t nb) {
    double ba[4], bbx[4];
    int ha = ls_bbox(a, na, ba), hb = ls_bbox(b, nb, bbx);
    /* has_disjoint_bboxes: only when both have a bbox */
    if (ha && hb) {
        if (ba[0] > bbx[2] || bbx[0] > ba[2] || ba[1] > bbx[3] || bbx[1] > ba[3]) return 0;
    }
    for (int64_t i = 0; i + 1 < na; ++i)
        for (int64_t j = 0; j + 1 < nb; ++j)
            if (line_intersects_line(b + 2 * j, b + 2 * j + 2, a + 2 * i, a + 2 * i + 2)) return 1;
    return 0;
}
/* ---- every type pair: geo 0.27 intersects/{coordinate,line,line_string,polygon,collections}.rs (recalled) ----
 * Written with geo's own nesting (every vertex tested, the polygon-level bounding-box rejects taken on the
 * exterior ring exactly where geo takes them); the CUDA kernel uses a shorter but equivalent procedure. */
typedef struct chain_t { /* linestring, or ring with Polygon::new's closing coord emulated */
    const double *xy;
    int64_t n;
    int closing;
} chain_t;
static inline int64_t chain_coords(const chain_t *c) { return c->n + c->closing; }
static inline const double *chain_at(const chain_t *c, int64_t i) { return c->xy + 2 * (i < c->n ? i : 0); }
static chain_t mk_line(const double *xy, int64_t c0, int64_t c1) {
    chain_t c = {xy + 2 * c0, c1 - c0, 0};
    return c;
}
static chain_t mk_ring(const double *xy, int64_t c0, int64_t c1) {
    chain_t c = {xy + 2 * c0, c1 - c0, 0};
    if (c.n >= 2) c.closing = !(c.xy[0] == c.xy[2 * (c.n - 1)] && c.xy[1] == c.xy[2 * (c.n - 1) + 1]);
    return c;
}
static int chain_bbox(const chain_t *c, double *b) { return ls_bbox(c->xy, c->n, b); }
static int bbox_disjoint(int ha, const double *a, int hb, const double *b) {
    return ha && hb && (a[0] > b[2] || b[0] > a[2] || a[1] > b[3] || b[1] > a[3]);
}
/* impl Intersects<Coord> for Polygon: exterior != Outside && all interiors != Inside */
static int polygon_intersects_coord(const og_array *a, int64_t r0, int64_t r1, const double *p) {
    const int64_t *ro = a->ring_off;
    if (r1 <= r0) return 0;
    if (coord_pos_ring(p[0], p[1], a->xy + 2 * ro[r0], ro[r0 + 1] - ro[r0]) == POS_OUTSIDE) return 0;
    for (int64_t h = r0 + 1; h < r1; ++h)
        if (coord_pos_ring(p[0], p[1], a->xy + 2 * ro[h], ro[h + 1] - ro[h]) == POS_INSIDE) return 0;
    return 1;
}
/* blanket impl Intersects<Line> for LineString: bbox reject, then lines().any() */
static int chain_intersects_line(const chain_t *c, const double *s, const double *e) {
    double bc[4], bl[4] = {fmin(s[0], e[0]), fmin(s[1], e[1]), fmax(s[0], e[0]), fmax(s[1], e[1])};
    int hc = chain_bbox(c, bc);
    if (bbox_disjoint(hc, bc, 1, bl)) return 0;
    for (int64_t i = 0; i + 1 < chain_coords(c); ++i)
        if (line_intersects_line(chain_at(c, i), chain_at(c, i + 1), s, e)) return 1;
    return 0;
}
/* impl Intersects<Line> for Polygon */
static int polygon_intersects_line(const og_array *a, int64_t r0, int64_t r1, const double *s, const double *e) {
    for (int64_t r = r0; r < r1; ++r) {
        chain_t ring = mk_ring(a->xy, a->ring_off[r], a->ring_off[r + 1]);
        if (chain_intersects_line(&ring, s, e)) return 1;
    }
    return polygon_intersects_coord(a, r0, r1, s) || polygon_intersects_coord(a, r0, r1, e);
}
/* symmetric impl of the blanket LineString impl: has_disjoint_bboxes(linestring, polygon), then lines().any() */
static int polygon_intersects_chain(const og_array *a, int64_t r0, int64_t r1, const chain_t *c) {
    if (r1 <= r0) return 0;
    double bc[4], bp[4];
    int hc = chain_bbox(c, bc);
    int hp = ls_bbox(a->xy + 2 * a->ring_off[r0], a->ring_off[r0 + 1] - a->ring_off[r0], bp); /* Polygon::bounding_rect = exterior's */
    if (bbox_disjoint(hc, bc, hp, bp)) return 0;
    for (int64_t i = 0; i + 1 < chain_coords(c); ++i)
        if (polygon_intersects_line(a, r0, r1, chain_at(c, i), chain_at(c, i + 1))) return 1;
    return 0;
}
/* impl Intersects<Polygon> for Polygon: self = (s, s0..s1), polygon = (o, o0..o1) */
static int polygon_intersects_polygon(const og_array *s, int64_t s0, int64_t s1, const og_array *o, int64_t o0, int64_t o1) {
    if (s1 <= s0 || o1 <= o0) return 0;
    double bs[4], bo[4];
    int hs = ls_bbox(s->xy + 2 * s->ring_off[s0], s->ring_off[s0 + 1] - s->ring_off[s0], bs);
    int ho = ls_bbox(o->xy + 2 * o->ring_off[o0], o->ring_off[o0 + 1] - o->ring_off[o0], bo);
    if (bbox_disjoint(hs, bs, ho, bo)) return 0;
    for (int64_t r = o0; r < o1; ++r) { /* polygon.exterior(), then polygon.interiors() */
        chain_t ring = mk_ring(o->xy, o->ring_off[r], o->ring_off[r + 1]);
        if (polygon_intersects_chain(s, s0, s1, &ring)) return 1;
    }
    chain_t sext = mk_ring(s->xy, s->ring_off[s0], s->ring_off[s0 + 1]);
    return polygon_intersects_chain(o, o0, o1, &sext);
}
static int geom_class(int t) { return (t == OG_POINT || t == OG_MULTIPOINT) ? 0 : (t == OG_LINESTRING || t == OG_MULTILINESTRING) ? 1 : 2; }
static void sub_range(const og_array *a, int64_t r, int64_t *lo, int64_t *hi) {
    if (a->type == OG_POINT) *lo = r, *hi = r + 1;
    else if (a->type == OG_LINESTRING || a->type == OG_POLYGON) *lo = 0, *hi = 1;
    else *lo = a->geom_off[r], *hi = a->geom_off[r + 1];
}
static chain_t sub_chain(const og_array *a, int64_t r, int64_t k) {
    if (a->type == OG_LINESTRING) return mk_line(a->xy, a->geom_off[r], a->geom_off[r + 1]);
    return mk_line(a->xy, a->ring_off[k], a->ring_off[k + 1]);
}
static void sub_part(const og_array *a, int64_t r, int64_t q, int64_t *r0, int64_t *r1) {
    if (a->type == OG_POLYGON) *r0 = a->geom_off[r], *r1 = a->geom_off[r + 1];
    else *r0 = a->part_off[q], *r1 = a->part_off[q + 1];
}
static int row_intersects(const og_array *a, int64_t ra, const og_array *b, int64_t rb) {
    int ca = geom_class(a->type), cb = geom_class(b->type);
    int64_t a0, a1, b0, b1;
    if (ca == 2 && cb == 2) {
        sub_range(a, ra, &a0, &a1);
        sub_range(b, rb, &b0, &b1);
        /* Multi*: self.iter().any(|p| p.intersects(rhs)); Polygon x MultiPolygon is the symmetric impl, so the
         * MultiPolygon's members are `self` */
        int self_is_a = b->type == OG_POLYGON;
        for (int64_t p = a0; p < a1; ++p) {
            int64_t s0, s1;
            sub_part(a, ra, p, &s0, &s1);
            for (int64_t q = b0; q < b1; ++q) {
                int64_t o0, o1;
                sub_part(b, rb, q, &o0, &o1);
                if (self_is_a ? polygon_intersects_polygon(a, s0, s1, b, o0, o1) : polygon_intersects_polygon(b, o0, o1, a, s0, s1)) return 1;
            }
        }
        return 0;
    }
    if (ca > cb) {
        const og_array *t = a; a = b; b = t;
        int64_t tr = ra; ra = rb; rb = tr;
        int tc = ca; ca = cb; cb = tc;
    }
    sub_range(a, ra, &a0, &a1);
    sub_range(b, rb, &b0, &b1);
    if (ca == 0) {
        for (int64_t i = a0; i < a1; ++i) {
            const double *p = a->xy + 2 * i;
            if (cb == 0) {
                for (int64_t j = b0; j < b1; ++j)
                    if (b->xy[2 * j] == p[0] && b->xy[2 * j + 1] == p[1]) return 1;
            } else if (cb == 1) {
                for (int64_t k = b0; k < b1; ++k) {
                    chain_t c = sub_chain(b, rb, k);
                    for (int64_t j = 0; j + 1 < c.n; ++j)
                        if (line_intersects_coord(c.xy + 2 * j, c.xy + 2 * j + 2, p)) return 1;
                }
            } else {
                for (int64_t q = b0; q < b1; ++q) {
                    int64_t r0, r1;
                    sub_part(b, rb, q, &r0, &r1);
                    if (polygon_intersects_coord(b, r0, r1, p)) return 1;
                }
            }
        }
        return 0;
    }
    for (int64_t k = a0; k < a1; ++k) {
        chain_t c = sub_chain(a, ra, k);
        if (cb == 1) {
            for (int64_t m = b0; m < b1; ++m) {
                chain_t d = sub_chain(b, rb, m);
                if (ls_intersects_ls(c.xy, c.n, d.xy, d.n)) return 1;
            }
        } else {
            for (int64_t q = b0; q < b1; ++q) {
                int64_t r0, r1;
                sub_part(b, rb, q, &r0, &r1);
                if (polygon_intersects_chain(b, r0, r1, &c)) return 1;
            }
        }
    }
    return 0;
}
void og_intersects_rowwise(const og_array *a, const og_array *b, uint8_t *out, int threads) {
    int nt = resolve_threads(threads);
    (void)nt;
#pragma omp parallel for num_threads(nt) schedule(dynamic, 1024)
    for (int64_t i = 0; i < a->n; ++i) {
        uint8_t r = 0;
        if (is_valid(a, i) && is_valid(b, i)) {
            if (a->type == OG_LINESTRING && b->type == OG_LINESTRING)
                r = (uint8_t)ls_intersects_ls(a->xy + 2 * a->geom_off[i], a->geom_off[i + 1] - a->geom_off[i],
                                             b->xy + 2 * b->geom_off[i], b->geom_off[i + 1] - b->geom_off[i]);
            else
                r = (uint8_t)row_intersects(a, i, b, i);
        }
        out[i] = r;
    }
}

/* row-wise contains of a point: (Multi)Polygon -> og_contains_point; (Multi)LineString -> geo 0.27
 * contains/line_string.rs + contains/line.rs (recalled; spatial_index.rs:125-135 are the call sites) */
static int line_contains_coord(const double *s, const double *e, const double *c) {
    if (s[0] == e[0] && s[1] == e[1]) return s[0] == c[0] && s[1] == c[1];
    return !(c[0] == s[0] && c[1] == s[1]) && !(c[0] == e[0] && c[1] == e[1]) && line_intersects_coord(s, e, c);
}
static int linestring_contains_coord(const double *xy, int64_t n, const double *c) {
    if (n == 0) return 0;
    const double *f = xy, *l = xy + 2 * (n - 1);
    if ((c[0] == f[0] && c[1] == f[1]) || (c[0] == l[0] && c[1] == l[1])) return f[0] == l[0] && f[1] == l[1]; /* is_closed() */
    for (int64_t i = 0; i + 1 < n; ++i) {
        const double *s = xy + 2 * i, *e = xy + 2 * i + 2;
        if (line_contains_coord(s, e, c) || (i > 0 && c[0] == s[0] && c[1] == s[1])) return 1;
    }
    return 0;
}
void og_contains_rowwise(const og_array *a, const double *pts_xy, const uint8_t *pts_valid, uint8_t *out, int threads) {
    int nt = resolve_threads(threads);
    (void)nt;
#pragma omp parallel for num_threads(nt) schedule(dynamic, 1024)
    for (int64_t i = 0; i < a->n; ++i) {
        uint8_t r = 0;
        int pv = pts_valid ? ((pts_valid[i >> 3] >> (i & 7)) & 1) : 1;
        if (is_valid(a, i) && pv) {
            const double *p = pts_xy + 2 * i;
            if (a->type == OG_POLYGON || a->type == OG_MULTIPOLYGON) r = (uint8_t)og_contains_point(a, i, p[0], p[1]);
            else if (a->type == OG_LINESTRING) r = (uint8_t)linestring_contains_coord(a->xy + 2 * a->geom_off[i], a->geom_off[i + 1] - a->geom_off[i], p);
            else if (a->type == OG_MULTILINESTRING)
                for (int64_t k = a->geom_off[i]; k < a->geom_off[i + 1] && !r; ++k)
                    r = (uint8_t)linestring_contains_coord(a->xy + 2 * a->ring_off[k], a->ring_off[k + 1] - a->ring_off[k], p);
        }
        out[i] = r;
    }
}

/* ------------------------------------------------------------------------------------------- */
/* Polygon / MultiPolygon contains Polygon — spatial_index.rs:99-110 (`poly_lhs.contains(poly_rhs)`)        */
/* ------------------------------------------------------------------------------------------- */
/* geo 0.27 contains/polygon.rs (recalled): `impl_contains_from_relate!(Polygon<T>, [.. Polygon ..])`, i.e.
 * self.relate(rhs).is_contains() = DE-9IM [T*****FF*]: the interiors meet and no point of rhs (interior or boundary)
 * lies in the exterior of self.  For VALID operands (rings simple, holes inside, members of a MultiPolygon touching in
 * points only) that is:   B subset of closure(A)   and   interior(A) meets interior(B).   geo evaluates it on its
 * topology graph; this restatement decides the same three conditions with exact orientation signs only:
 *   C1  boundary(B) inside closure(A):  no B edge leaves A.  A straight B edge can only leave closure(A) (i) at its start,
 *       (ii) by crossing an A edge properly (both interiors) or (iii) at an A vertex lying inside it; so it suffices that
 *       the edge, just after its start and just after every A vertex on it, is not in the exterior, and that no proper
 *       crossing exists (unless an A vertex sits exactly on the crossing — members touching there — which (iii) judges).
 *   C2  no point of boundary(A) is strictly inside B (otherwise a hole of A, or a gap between touching members, lies in B):
 *       the same three events with the roles swapped.
 *   C3  anchor: with C1 and C2 the connected interior of B is entirely inside or entirely outside A; one point just
 *       beside the first edge of B, on B's own interior side, decides.
 * "Just after / just beside" are SYMBOLIC points q = x + eps (v - x) + eps^2 side L(v - x), L = left normal, eps an
 * infinitesimal: every comparison geo's coord_pos_relative_to_ring makes is evaluated on x first and, on a tie, on the
 * eps and eps^2 terms, whose signs are exact (coordinate comparisons and orient2d of input vertices).
 * Parity unpinned (the reference holds no vector for it); refereed by oracle/exact.py's rational arrangement form. */
typedef struct {
    double xx, xy, vx, vy;
    int side; /* 0: on the open segment x->v just after x; +1 / -1: just left / right of it */
} symq_t;
static inline int cmpd(double a, double b) { return (a > b) - (a < b); }
static inline int sgnd(double a) { return (a > 0.0) - (a < 0.0); }
/* sign of (cy - q.y),  q.y = x.y + eps (v.y - x.y) + eps^2 side (v.x - x.x) */
static int sym_cmp_y(const symq_t *q, double cy) {
    int c = cmpd(cy, q->xy);
    if (c) return c;
    c = -cmpd(q->vy, q->xy);
    if (c) return c;
    return -q->side * cmpd(q->vx, q->xx);
}
/* sign of (cx - q.x),  q.x = x.x + eps (v.x - x.x) - eps^2 side (v.y - x.y) */
static int sym_cmp_x(const symq_t *q, double cx) {
    int c = cmpd(cx, q->xx);
    if (c) return c;
    c = -cmpd(q->vx, q->xx);
    if (c) return c;
    return q->side * cmpd(q->vy, q->xy);
}
/* sign of orient2d(s, e, q): affine in q; the eps term is orient2d(s,e,v) once orient2d(s,e,x) = 0, the eps^2 term is
 * side * dot(e - s, v - x), and with s, e, x, v collinear the dot product's sign is a product of coordinate orders */
static int sym_orient(const double *s, const double *e, const symq_t *q) {
    int o = sgnd(og_orient2d(s[0], s[1], e[0], e[1], q->xx, q->xy));
    if (o) return o;
    o = sgnd(og_orient2d(s[0], s[1], e[0], e[1], q->vx, q->vy));
    if (o || !q->side) return o;
    int d1, d2;
    if (e[0] != s[0]) d1 = cmpd(e[0], s[0]), d2 = cmpd(q->vx, q->xx);
    else d1 = cmpd(e[1], s[1]), d2 = cmpd(q->vy, q->xy);
    return q->side * d1 * d2;
}
static int sym_between_x(const symq_t *q, double b1, double b2) { /* value_in_between(q.x, b1, b2) */
    if (b1 < b2) return sym_cmp_x(q, b1) <= 0 && sym_cmp_x(q, b2) >= 0;
    return sym_cmp_x(q, b2) <= 0 && sym_cmp_x(q, b1) >= 0;
}
/* coord_pos_relative_to_ring (the loop of coord_pos_ring above) for a symbolic point */
static int sym_ring_pos(const symq_t *q, const double *xy, int64_t n) {
    if (n < 2) return POS_OUTSIDE; /* q never equals a vertex */
    chain_t c = mk_ring(xy, 0, n);
    int64_t m = chain_coords(&c);
    int wn = 0;
    for (int64_t i = 0; i + 1 < m; ++i) {
        const double *s = chain_at(&c, i), *e = chain_at(&c, i + 1);
        int cs = sym_cmp_y(q, s[1]), ce = sym_cmp_y(q, e[1]); /* signs of s.y - q.y, e.y - q.y */
        if (cs <= 0) {
            if (ce >= 0) {
                int o = sym_orient(s, e, q);
                if (o > 0 && ce != 0) wn += 1;
                else if (o == 0 && sym_between_x(q, s[0], e[0])) return POS_BOUNDARY;
            }
        } else if (ce <= 0) {
            int o = sym_orient(s, e, q);
            if (o < 0) wn -= 1;
            else if (o == 0 && sym_between_x(q, s[0], e[0])) return POS_BOUNDARY;
        }
    }
    return wn ? POS_INSIDE : POS_OUTSIDE;
}
/* position of q in the union of the polygon members [m0, m1) of the POLYGON view `a` */
static int sym_region_pos(const og_array *a, int64_t m0, int64_t m1, const symq_t *q) {
    int boundary = 0;
    for (int64_t m = m0; m < m1; ++m) {
        int64_t r0 = a->geom_off[m], r1 = a->geom_off[m + 1];
        if (r1 <= r0) continue;
        int pe = sym_ring_pos(q, a->xy + 2 * a->ring_off[r0], a->ring_off[r0 + 1] - a->ring_off[r0]);
        if (pe == POS_OUTSIDE) continue;
        if (pe == POS_BOUNDARY) {
            boundary = 1;
            continue;
        }
        int in_hole = 0;
        for (int64_t r = r0 + 1; r < r1 && !in_hole; ++r) {
            int ph = sym_ring_pos(q, a->xy + 2 * a->ring_off[r], a->ring_off[r + 1] - a->ring_off[r]);
            if (ph == POS_INSIDE) in_hole = 1;
            else if (ph == POS_BOUNDARY) in_hole = 1, boundary = 1;
        }
        if (!in_hole) return POS_INSIDE;
    }
    return boundary ? POS_BOUNDARY : POS_OUTSIDE;
}
/* c strictly between u and v on their common line (c collinear with u, v by the caller's orient2d == 0) */
static int strictly_between(const double *c, const double *u, const double *v) {
    if (u[0] != v[0]) return (c[0] > fmin(u[0], v[0])) && (c[0] < fmax(u[0], v[0]));
    return (c[1] > fmin(u[1], v[1])) && (c[1] < fmax(u[1], v[1]));
}
/* is some vertex of the members [m0,m1) of `a` on both lines u-v and s-e (i.e. exactly at their crossing)? */
static int vertex_at_crossing(const og_array *a, int64_t m0, int64_t m1, const double *u, const double *v, const double *s, const double *e) {
    for (int64_t r = a->geom_off[m0]; r < a->geom_off[m1]; ++r)
        for (int64_t c = a->ring_off[r]; c < a->ring_off[r + 1]; ++c) {
            const double *w = a->xy + 2 * c;
            if (og_orient2d(u[0], u[1], v[0], v[1], w[0], w[1]) == 0.0 && og_orient2d(s[0], s[1], e[0], e[1], w[0], w[1]) == 0.0) return 1;
        }
    return 0;
}
/* the edges of region X = members [x0,x1) of view x must not reach into FORBIDDEN (POS_OUTSIDE or POS_INSIDE) of region
 * Y = members [y0,y1) of view y; check_cross: proper crossings make it fail (C1) */
static int edges_avoid(const og_array *x, int64_t x0, int64_t x1, const og_array *y, int64_t y0, int64_t y1, int forbidden, int check_cross) {
    for (int64_t rx = x->geom_off[x0]; rx < x->geom_off[x1]; ++rx) {
        chain_t cx = mk_ring(x->xy, x->ring_off[rx], x->ring_off[rx + 1]);
        int64_t nx = chain_coords(&cx);
        for (int64_t i = 0; i + 1 < nx; ++i) {
            const double *u = chain_at(&cx, i), *v = chain_at(&cx, i + 1);
            if (u[0] == v[0] && u[1] == v[1]) continue;
            symq_t q = {u[0], u[1], v[0], v[1], 0};
            if (sym_region_pos(y, y0, y1, &q) == forbidden) return 0;
            for (int64_t ry = y->geom_off[y0]; ry < y->geom_off[y1]; ++ry) {
                chain_t cy = mk_ring(y->xy, y->ring_off[ry], y->ring_off[ry + 1]);
                int64_t ny = chain_coords(&cy);
                for (int64_t j = 0; j + 1 < ny; ++j) {
                    const double *s = chain_at(&cy, j), *e = chain_at(&cy, j + 1);
                    int o1 = sgnd(og_orient2d(u[0], u[1], v[0], v[1], s[0], s[1]));
                    if (check_cross && !(s[0] == e[0] && s[1] == e[1])) {
                        int o2 = sgnd(og_orient2d(u[0], u[1], v[0], v[1], e[0], e[1]));
                        if (o1 * o2 < 0) {
                            int o3 = sgnd(og_orient2d(s[0], s[1], e[0], e[1], u[0], u[1])), o4 = sgnd(og_orient2d(s[0], s[1], e[0], e[1], v[0], v[1]));
                            if (o3 * o4 < 0 && !vertex_at_crossing(y, y0, y1, u, v, s, e)) return 0;
                        }
                    }
                    if (o1 == 0 && strictly_between(s, u, v)) { /* a vertex of Y inside this edge: the part after it */
                        symq_t q2 = {s[0], s[1], v[0], v[1], 0};
                        if (sym_region_pos(y, y0, y1, &q2) == forbidden) return 0;
                    }
                }
            }
        }
    }
    return 1;
}
/* A = members [a0,a1) of the POLYGON view `a` (one for Polygon, the parts of a MultiPolygon); B = polygon b of view `bv` */
static int region_contains_polygon(const og_array *a, int64_t a0, int64_t a1, const og_array *bv, int64_t b) {
    int64_t br0 = bv->geom_off[b], br1 = bv->geom_off[b + 1];
    if (a1 <= a0 || br1 <= br0 || bv->ring_off[br0 + 1] - bv->ring_off[br0] < 3) return 0;
    if (!edges_avoid(bv, b, b + 1, a, a0, a1, POS_OUTSIDE, 1)) return 0; /* C1 */
    if (!edges_avoid(a, a0, a1, bv, b, b + 1, POS_INSIDE, 0)) return 0;  /* C2 */
    chain_t ce = mk_ring(bv->xy, bv->ring_off[br0], bv->ring_off[br0 + 1]);
    int64_t n = chain_coords(&ce);
    for (int64_t i = 0; i + 1 < n; ++i) { /* C3: first non-degenerate edge of B's exterior */
        const double *u = chain_at(&ce, i), *v = chain_at(&ce, i + 1);
        if (u[0] == v[0] && u[1] == v[1]) continue;
        symq_t q = {u[0], u[1], v[0], v[1], 1};
        if (sym_region_pos(bv, b, b + 1, &q) != POS_INSIDE) {
            q.side = -1;
            if (sym_region_pos(bv, b, b + 1, &q) != POS_INSIDE) return 0; /* no interior beside its own boundary: degenerate */
        }
        return sym_region_pos(a, a0, a1, &q) == POS_INSIDE;
    }
    return 0;
}
/* row-wise: a POLYGON or MULTIPOLYGON, b POLYGON (the pairs spatial_index.rs:99-115 dispatches with Predicate::Contains) */
int og_contains_polygon_rowwise(const og_array *a, const og_array *b, uint8_t *out, int threads) {
    int nt = resolve_threads(threads);
    (void)nt;
    if (a->n != b->n || (a->type != OG_POLYGON && a->type != OG_MULTIPOLYGON) || b->type != OG_POLYGON) return -1;
    og_array va = *a;
    if (a->type == OG_MULTIPOLYGON) va.type = OG_POLYGON, va.geom_off = a->part_off, va.part_off = NULL;
#pragma omp parallel for num_threads(nt) schedule(dynamic, 64)
    for (int64_t i = 0; i < a->n; ++i) {
        uint8_t r = 0;
        if (is_valid(a, i) && is_valid(b, i)) {
            int64_t m0 = a->type == OG_MULTIPOLYGON ? a->geom_off[i] : i, m1 = a->type == OG_MULTIPOLYGON ? a->geom_off[i + 1] : i + 1;
            r = (uint8_t)region_contains_polygon(&va, m0, m1, b, i);
        }
        out[i] = r;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------- */
/* distance — GeoSeries::distance geoseries.rs:141-146 ; geo euclidean_distance.rs +              */
/* geo-types private_utils.rs recalled                                                          */
/* ------------------------------------------------------------------------------------------- */
static inline double pt_dist(const double *a, const double *b) { return hypot(b[0] - a[0], b[1] - a[1]); }
static double line_segment_distance(const double *p, const double *s, const double *e) {
    if (s[0] == e[0] && s[1] == e[1]) return pt_dist(p, s);
    double dx = e[0] - s[0], dy = e[1] - s[1];
    double d2 = dx * dx + dy * dy;
    double r = ((p[0] - s[0]) * dx + (p[1] - s[1]) * dy) / d2;
    if (r <= 0.0) return pt_dist(p, s);
    if (r >= 1.0) return pt_dist(p, e);
    double sv = ((s[1] - p[1]) * dx - (s[0] - p[0]) * dy) / d2;
    return fabs(sv) * hypot(dx, dy);
}
/* geo-types private_utils::line_string_contains_point (inexact, epsilon based) */
static int line_string_contains_point(const double *xy, int64_t n, const double *p) {
    if (n == 0) return 0;
    if (n == 1) {
        float d = (float)pt_dist(xy, p);
        return d <= 1.1920929e-07f; /* approx::relative_eq!(d, 0.0) on f32 */
    }
    for (int64_t i = 0; i < n; ++i)
        if (xy[2 * i] == p[0] && xy[2 * i + 1] == p[1]) return 1;
    for (int64_t i = 0; i + 1 < n; ++i) {
        const double *s = xy + 2 * i, *e = xy + 2 * i + 2;
        double dx = e[0] - s[0], dy = e[1] - s[1];
        int hx = dx != 0.0, hy = dy != 0.0;
        double tx = hx ? (p[0] - s[0]) / dx : 0.0;
        double ty = hy ? (p[1] - s[1]) / dy : 0.0;
        int c;
        if (!hx && !hy) c = (p[0] == s[0] && p[1] == s[1]);
        else if (hx && !hy) c = (p[1] == s[1] && 0.0 <= tx && tx <= 1.0);
        else if (!hx && hy) c = (p[0] == s[0] && 0.0 <= ty && ty <= 1.0);
        else c = (fabs(tx - ty) <= 2.220446049250313e-16 && 0.0 <= tx && tx <= 1.0);
        if (c) return 1;
    }
    return 0;
}
static double point_ls_distance(const double *p, const double *xy, int64_t n) {
    if (line_string_contains_point(xy, n, p) || n == 0) return 0.0;
    double acc = 1.7976931348623157e308;
    for (int64_t i = 0; i + 1 < n; ++i) {
        double v = line_segment_distance(p, xy + 2 * i, xy + 2 * i + 2);
        acc = fmin(acc, v);
    }
    return acc;
}
static double ls_ls_distance(const double *a, int64_t na, const double *b, int64_t nb, int *ok) {
    if (ls_intersects_ls(a, na, b, nb)) return 0.0;
    if (na < 2 || nb < 2) { /* reference: nearest_neighbor(..).unwrap() panics on an empty tree */
        *ok = 0;
        return NAN;
    }
    double m1 = 1.7976931348623157e308, m2 = 1.7976931348623157e308;
    for (int64_t j = 0; j < nb; ++j) { /* every point of B against lines of A */
        double best = 1.7976931348623157e308;
        for (int64_t i = 0; i + 1 < na; ++i) best = fmin(best, line_segment_distance(b + 2 * j, a + 2 * i, a + 2 * i + 2));
        m1 = fmin(m1, best);
    }
    for (int64_t i = 0; i < na; ++i) {
        double best = 1.7976931348623157e308;
        for (int64_t j = 0; j + 1 < nb; ++j) best = fmin(best, line_segment_distance(a + 2 * i, b + 2 * j, b + 2 * j + 2));
        m2 = fmin(m2, best);
    }
    return fmin(m1, m2);
}
static double point_polygon_distance(const double *p, const og_array *a, int64_t i) {
    const int64_t *go = a->geom_off, *ro = a->ring_off;
    int64_t r0 = go[i], r1 = go[i + 1];
    if (r1 <= r0 || ro[r0 + 1] - ro[r0] == 0) return 0.0;
    if (og_coord_position(a, i, p[0], p[1]) != POS_OUTSIDE) return 0.0;
    double acc = 1.7976931348623157e308;
    for (int64_t r = r0 + 1; r < r1; ++r) acc = fmin(acc, point_ls_distance(p, a->xy + 2 * ro[r], ro[r + 1] - ro[r]));
    double ext = 1.7976931348623157e308;
    const double *xy = a->xy + 2 * ro[r0];
    for (int64_t k = 0; k + 1 < ro[r0 + 1] - ro[r0]; ++k) ext = fmin(ext, line_segment_distance(p, xy + 2 * k, xy + 2 * k + 2));
    return fmin(acc, ext);
}
/* nearest_neighbour_distance(a, b): min over {point of a -> lines of b} U {point of b -> lines of a}; geo finds
 * the nearest line through an rstar tree and evaluates the same point-line formula */
static double chain_nn_distance(const chain_t *a, const chain_t *b) {
    double m = 1.7976931348623157e308;
    int64_t na = chain_coords(a), nb = chain_coords(b);
    for (int64_t j = 0; j < nb; ++j)
        for (int64_t i = 0; i + 1 < na; ++i) m = fmin(m, line_segment_distance(chain_at(b, j), chain_at(a, i), chain_at(a, i + 1)));
    for (int64_t i = 0; i < na; ++i)
        for (int64_t j = 0; j + 1 < nb; ++j) m = fmin(m, line_segment_distance(chain_at(a, i), chain_at(b, j), chain_at(b, j + 1)));
    return m;
}
/* ring_contains_point: strictly inside the exterior ring */
static int exterior_contains(const og_array *p, int64_t r0, const double *c) {
    return coord_pos_ring(c[0], c[1], p->xy + 2 * p->ring_off[r0], p->ring_off[r0 + 1] - p->ring_off[r0]) == POS_INSIDE;
}
/* impl EuclideanDistance<Polygon> for LineString (geo 0.27 euclidean_distance.rs, recalled) */
static double ls_polygon_distance(const chain_t *ls, const og_array *p, int64_t i, int *ok) {
    int64_t r0 = p->geom_off[i], r1 = p->geom_off[i + 1];
    if (ls->n < 2 || r1 <= r0 || p->ring_off[r0 + 1] - p->ring_off[r0] < 2) { /* empty rstar tree -> unwrap() panics */
        *ok = 0;
        return NAN;
    }
    if (polygon_intersects_chain(p, r0, r1, ls)) return 0.0;
    if (r1 - r0 > 1 && exterior_contains(p, r0, ls->xy)) {
        double m = 1.7976931348623157e308;
        for (int64_t r = r0 + 1; r < r1; ++r) {
            chain_t ring = mk_ring(p->xy, p->ring_off[r], p->ring_off[r + 1]);
            m = fmin(m, chain_nn_distance(ls, &ring));
        }
        return m;
    }
    chain_t ext = mk_ring(p->xy, p->ring_off[r0], p->ring_off[r0 + 1]);
    return chain_nn_distance(ls, &ext);
}
/* impl EuclideanDistance<Polygon> for Polygon.  geo switches to rotating calipers (min_poly_dist) when both
 * polygons are convex; that path computes the same minimum with different roundings and is restated here by
 * the nearest-neighbour form (tolerance 1e-9 relative covers it). */
static double polygon_polygon_distance(const og_array *a, int64_t ia, const og_array *b, int64_t ib, int *ok) {
    int64_t a0 = a->geom_off[ia], a1 = a->geom_off[ia + 1], b0 = b->geom_off[ib], b1 = b->geom_off[ib + 1];
    if (a1 <= a0 || b1 <= b0 || a->ring_off[a0 + 1] - a->ring_off[a0] < 2 || b->ring_off[b0 + 1] - b->ring_off[b0] < 2) {
        *ok = 0;
        return NAN;
    }
    if (polygon_intersects_polygon(a, a0, a1, b, b0, b1)) return 0.0;
    chain_t ea = mk_ring(a->xy, a->ring_off[a0], a->ring_off[a0 + 1]), eb = mk_ring(b->xy, b->ring_off[b0], b->ring_off[b0 + 1]);
    if (a1 - a0 > 1 && exterior_contains(a, a0, eb.xy)) {
        double m = 1.7976931348623157e308;
        for (int64_t r = a0 + 1; r < a1; ++r) {
            chain_t ring = mk_ring(a->xy, a->ring_off[r], a->ring_off[r + 1]);
            m = fmin(m, chain_nn_distance(&eb, &ring));
        }
        return m;
    }
    if (b1 - b0 > 1 && exterior_contains(b, b0, ea.xy)) {
        double m = 1.7976931348623157e308;
        for (int64_t r = b0 + 1; r < b1; ++r) {
            chain_t ring = mk_ring(b->xy, b->ring_off[r], b->ring_off[r + 1]);
            m = fmin(m, chain_nn_distance(&ea, &ring));
        }
        return m;
    }
    return chain_nn_distance(&ea, &eb);
}
/* distance between two SINGLE geometries (Point / LineString / Polygon views, member indices ia / ib) */
static double single_distance(const og_array *a, int64_t ia, const og_array *b, int64_t ib, int *ok) {
    int ta = a->type, tb = b->type;
    if (ta == OG_POINT && tb == OG_POINT) return pt_dist(a->xy + 2 * ia, b->xy + 2 * ib);
    if (ta == OG_POINT && tb == OG_LINESTRING) return point_ls_distance(a->xy + 2 * ia, b->xy + 2 * b->geom_off[ib], b->geom_off[ib + 1] - b->geom_off[ib]);
    if (ta == OG_LINESTRING && tb == OG_POINT) return point_ls_distance(b->xy + 2 * ib, a->xy + 2 * a->geom_off[ia], a->geom_off[ia + 1] - a->geom_off[ia]);
    if (ta == OG_LINESTRING && tb == OG_LINESTRING)
        return ls_ls_distance(a->xy + 2 * a->geom_off[ia], a->geom_off[ia + 1] - a->geom_off[ia], b->xy + 2 * b->geom_off[ib],
                              b->geom_off[ib + 1] - b->geom_off[ib], ok);
    if (ta == OG_POINT && tb == OG_POLYGON) return point_polygon_distance(a->xy + 2 * ia, b, ib);
    if (ta == OG_POLYGON && tb == OG_POINT) return point_polygon_distance(b->xy + 2 * ib, a, ia);
    if (ta == OG_LINESTRING && tb == OG_POLYGON) {
        chain_t ls = mk_line(a->xy, a->geom_off[ia], a->geom_off[ia + 1]);
        return ls_polygon_distance(&ls, b, ib, ok);
    }
    if (ta == OG_POLYGON && tb == OG_LINESTRING) {
        chain_t ls = mk_line(b->xy, b->geom_off[ib], b->geom_off[ib + 1]);
        return ls_polygon_distance(&ls, a, ia, ok);
    }
    return polygon_polygon_distance(a, ia, b, ib, ok);
}
/* Members of a Multi* row as single geometries: a VIEW of the same buffers whose geom_off is the next offset level
 * (MultiPoint -> points, MultiLineString -> its lines via ring_off, MultiPolygon -> its polygons via part_off). */
static og_array member_view(const og_array *a) {
    og_array v = *a;
    v.valid = NULL;
    switch (a->type) {
    case OG_MULTIPOINT: v.type = OG_POINT, v.geom_off = NULL; break;
    case OG_MULTILINESTRING: v.type = OG_LINESTRING, v.geom_off = a->ring_off, v.ring_off = NULL; break;
    case OG_MULTIPOLYGON: v.type = OG_POLYGON, v.geom_off = a->part_off, v.part_off = NULL; break;
    default: break;
    }
    return v;
}
static void member_range(const og_array *a, int64_t i, int64_t *lo, int64_t *hi) {
    if (a->type == OG_MULTIPOINT || a->type == OG_MULTILINESTRING || a->type == OG_MULTIPOLYGON) *lo = a->geom_off[i], *hi = a->geom_off[i + 1];
    else *lo = i, *hi = i + 1;
}
/* GeoSeries::distance (geoseries.rs:141-146) row-wise.  Multi* operands: geo 0.27 euclidean_distance.rs (recalled)
 * `impl_euclidean_distance_for_iter_geometry!`: self.iter().map(|g| g.euclidean_distance(target))
 * .fold(T::max_value(), |acc, v| acc.min(v)) — the minimum over the members, f64::MAX for an empty collection; a member
 * pair on which geo would panic (nearest_neighbor on an empty tree) makes the row invalid (NaN). */
int og_distance_rowwise(const og_array *a, const og_array *b, double *out, int threads) {
    int nt = resolve_threads(threads);
    (void)nt;
    if (a->n != b->n) return -1;
    const og_array va = member_view(a), vb = member_view(b);
#pragma omp parallel for num_threads(nt) schedule(dynamic, 1024)
    for (int64_t i = 0; i < a->n; ++i) {
        double v = NAN;
        if (is_valid(a, i) && is_valid(b, i)) {
            int ok = 1;
            int64_t a0, a1, b0, b1;
            member_range(a, i, &a0, &a1);
            member_range(b, i, &b0, &b1);
            v = 1.7976931348623157e308;
            for (int64_t p = a0; p < a1; ++p)
                for (int64_t q = b0; q < b1; ++q) {
                    double d = single_distance(&va, p, &vb, q, &ok);
                    v = fmin(v, d);
                }
            if (!ok) v = NAN;
        }
        out[i] = v;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------- */
/* convex hull — GeoSeries::convex_hull geoseries.rs:23-26 ; geo convex_hull/{mod,qhull}.rs       */
/* recalled, including the in-place slice permutations that fix the output vertex order.        */
/* ------------------------------------------------------------------------------------------- */
typedef struct {
    double x, y;
} coord_t;
static inline int lex_cmp(const coord_t *a, const coord_t *b) {
    if (a->x < b->x) return -1;
    if (a->x > b->x) return 1;
    if (a->y < b->y) return -1;
    if (a->y > b->y) return 1;
    return 0;
}
static inline int is_ccw(coord_t a, coord_t b, coord_t c) { return og_orient2d(a.x, a.y, b.x, b.y, c.x, c.y) > 0.0; }
static inline void cswap(coord_t *a, coord_t *b) {
    coord_t t = *a;
    *a = *b;
    *b = t;
}
/* utils::partition_slice (Hoare style): returns number of elements satisfying pred, moved to front */
static int64_t partition_ccw(coord_t *d, int64_t len, coord_t pa, coord_t pb) {
    if (len == 0) return 0;
    int64_t l = 0, r = len - 1;
    for (;;) {
        while (l < len && is_ccw(pa, pb, d[l])) l++;
        while (r > 0 && !is_ccw(pa, pb, d[r])) r--;
        if (l >= r) return l;
        cswap(&d[l], &d[r]);
    }
}
typedef struct {
    coord_t *v;
    int64_t n, cap;
} cvec;
static void cpush(cvec *h, coord_t c) {
    if (h->n == h->cap) {
        h->cap = h->cap ? h->cap * 2 : 64;
        h->v = (coord_t *)realloc(h->v, sizeof(coord_t) * (size_t)h->cap);
    }
    h->v[h->n++] = c;
}
static void hull_set(coord_t pa, coord_t pb, coord_t *set, int64_t len, cvec *hull) {
    if (len == 0) return;
    if (len == 1) {
        cpush(hull, set[0]);
        return;
    }
    double ox = pa.y - pb.y, oy = pb.x - pa.x;
    int64_t fi = 0;
    double fv = 0;
    for (int64_t i = 0; i < len; ++i) { /* Iterator::max_by — last maximal element wins */
        double dx = set[i].x - pa.x, dy = set[i].y - pa.y;
        double v = ox * dx + oy * dy;
        if (i == 0 || !(v < fv)) {
            fi = i;
            fv = v;
        }
    }
    cswap(&set[0], &set[fi]); /* swap_remove_to_first */
    coord_t far = set[0];
    set += 1;
    len -= 1;
    int64_t k = partition_ccw(set, len, far, pb);
    hull_set(far, pb, set, k, hull);
    cpush(hull, far);
    k = partition_ccw(set, len, pa, far);
    hull_set(pa, far, set, k, hull);
}
static void trivial_hull(coord_t *pts, int64_t n, cvec *hull) {
    coord_t ls[4];
    int64_t m = n;
    for (int64_t i = 0; i < n; ++i) ls[i] = pts[i];
    /* sort_unstable_by(lex_cmp) on <= 3 items */
    for (int64_t i = 1; i < m; ++i)
        for (int64_t j = i; j > 0 && lex_cmp(&ls[j], &ls[j - 1]) < 0; --j) cswap(&ls[j], &ls[j - 1]);
    if (m == 3 && og_orient2d(ls[0].x, ls[0].y, ls[1].x, ls[1].y, ls[2].x, ls[2].y) == 0.0) {
        ls[1] = ls[2];
        m = 2;
    }
    if (m == 1) ls[m++] = ls[0];
    if (m == 0) return;
    /* close */
    if (!(ls[0].x == ls[m - 1].x && ls[0].y == ls[m - 1].y)) ls[m++] = ls[0];
    /* make_ccw_winding: only a closed ring with >= 4 coords has a winding order */
    if (m >= 4) {
        /* least index = 0 after the sort; next = 1; prev = m-2 (m-1 duplicates index 0) */
        if (og_orient2d(ls[m - 2].x, ls[m - 2].y, ls[0].x, ls[0].y, ls[1].x, ls[1].y) < 0.0) {
            for (int64_t i = 0, j = m - 1; i < j; ++i, --j) cswap(&ls[i], &ls[j]);
        }
    }
    for (int64_t i = 0; i < m; ++i) cpush(hull, ls[i]);
}
static void quick_hull(coord_t *pts, int64_t n, cvec *hull) {
    if (n < 4) {
        trivial_hull(pts, n, hull);
        return;
    }
    int64_t min_idx = 0, max_idx = 0;
    for (int64_t i = 1; i < n; ++i) { /* utils::least_and_greatest_index: first least, first greatest */
        if (lex_cmp(&pts[i], &pts[min_idx]) < 0) min_idx = i;
        if (lex_cmp(&pts[i], &pts[max_idx]) > 0) max_idx = i;
    }
    cswap(&pts[0], &pts[min_idx]);
    coord_t mn = pts[0];
    pts += 1;
    n -= 1;
    if (max_idx == 0) max_idx = min_idx;
    max_idx = max_idx > 0 ? max_idx - 1 : 0;
    cswap(&pts[0], &pts[max_idx]);
    coord_t mx = pts[0];
    pts += 1;
    n -= 1;
    int64_t k = partition_ccw(pts, n, mx, mn);
    hull_set(mx, mn, pts, k, hull);
    cpush(hull, mx);
    k = partition_ccw(pts, n, mn, mx);
    hull_set(mn, mx, pts, k, hull);
    cpush(hull, mn);
    if (!(hull->v[0].x == hull->v[hull->n - 1].x && hull->v[0].y == hull->v[hull->n - 1].y)) cpush(hull, hull->v[0]);
}
/* CoordsIter::exterior_coords_iter */
static int64_t gather_exterior(const og_array *a, int64_t i, coord_t *dst) {
    const int64_t *go = a->geom_off, *ro = a->ring_off, *po = a->part_off;
    int64_t n = 0;
#define PUSH_RANGE(c0, c1)                                                                                   \
    for (int64_t c_ = (c0); c_ < (c1); ++c_) {                                                               \
        if (dst) { dst[n].x = a->xy[2 * c_]; dst[n].y = a->xy[2 * c_ + 1]; }                                  \
        n++;                                                                                                 \
    }
    switch (a->type) {
    case OG_POINT:
        PUSH_RANGE(i, i + 1);
        break;
    case OG_LINESTRING:
    case OG_MULTIPOINT:
        PUSH_RANGE(go[i], go[i + 1]);
        break;
    case OG_MULTILINESTRING:
        for (int64_t l = go[i]; l < go[i + 1]; ++l) PUSH_RANGE(ro[l], ro[l + 1]);
        break;
    case OG_POLYGON:
        if (go[i + 1] > go[i]) PUSH_RANGE(ro[go[i]], ro[go[i] + 1]);
        break;
    case OG_MULTIPOLYGON:
        for (int64_t p = go[i]; p < go[i + 1]; ++p)
            if (po[p + 1] > po[p]) PUSH_RANGE(ro[po[p]], ro[po[p] + 1]);
        break;
    default:
        break;
    }
    return n;
}
int64_t og_convex_hull_one(const og_array *a, int64_t i, double *out_xy, int64_t cap) {
    int64_t n = gather_exterior(a, i, NULL);
    coord_t *pts = (coord_t *)malloc(sizeof(coord_t) * (size_t)(n > 0 ? n : 1));
    gather_exterior(a, i, pts);
    cvec hull = {NULL, 0, 0};
    quick_hull(pts, n, &hull);
    int64_t m = hull.n;
    if (out_xy)
        for (int64_t k = 0; k < m && k < cap; ++k) {
            out_xy[2 * k] = hull.v[k].x;
            out_xy[2 * k + 1] = hull.v[k].y;
        }
    free(hull.v);
    free(pts);
    return m;
}
int64_t og_convex_hull(const og_array *a, int64_t *out_off, double *out_xy, int threads) {
    int nt = resolve_threads(threads);
    (void)nt;
    int64_t n = a->n;
    int64_t *sizes = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n + 1));
    if (out_xy == NULL) {
#pragma omp parallel for num_threads(nt) schedule(dynamic, 256)
        for (int64_t i = 0; i < n; ++i) sizes[i] = is_valid(a, i) ? og_convex_hull_one(a, i, NULL, 0) : 0;
        int64_t acc = 0;
        for (int64_t i = 0; i < n; ++i) {
            out_off[i] = acc;
            acc += sizes[i];
        }
        out_off[n] = acc;
        free(sizes);
        return acc;
    }
#pragma omp parallel for num_threads(nt) schedule(dynamic, 256)
    for (int64_t i = 0; i < n; ++i)
        if (is_valid(a, i)) og_convex_hull_one(a, i, out_xy + 2 * out_off[i], out_off[i + 1] - out_off[i]);
    free(sizes);
    return out_off[n];
}

/* ------------------------------------------------------------------------------------------- */
/* geodesic_length — GeoSeries::geodesic_length geoseries.rs:52-58 (impl :216-218); methods        */
/* Haversine / Vincenty / Geodesic named by py-geopolars/src/geo.rs:64-67 and the accessor        */
/* docstring (georust/geoseries.py:128-166).  geo 0.27 haversine_distance.rs, vincenty_distance.rs */
/* (recalled) and, for `geodesic`, Karney's inverse algorithm (arXiv:1109.4448) as implemented by  */
/* geographiclib (order-6 series), which geographiclib-rs 0.2 ports and geo calls.                */
/* ------------------------------------------------------------------------------------------- */
#define OG_MEAN_EARTH_RADIUS 6371008.8
#define OG_WGS84_A 6378137.0
#define OG_WGS84_B 6356752.314245 /* geo's POLAR_EARTH_RADIUS (Vincenty) */
static const double DEG = 0.017453292519943295769;
static double haversine_distance(const double *p, const double *q) {
    double theta1 = p[1] * DEG, theta2 = q[1] * DEG;
    double delta_theta = (q[1] - p[1]) * DEG, delta_lambda = (q[0] - p[0]) * DEG;
    double s1 = sin(delta_theta / 2.0), s2 = sin(delta_lambda / 2.0);
    double a = s1 * s1 + cos(theta1) * cos(theta2) * (s2 * s2);
    double c = 2.0 * asin(sqrt(a));
    return OG_MEAN_EARTH_RADIUS * c;
}
/* Ok(distance) -> 1, Err(FailedToConvergeError) -> 0 */
static int vincenty_distance(const double *p, const double *q, double *out) {
    const double a = OG_WGS84_A, b = OG_WGS84_B, f = (OG_WGS84_A - OG_WGS84_B) / OG_WGS84_A;
    double L = (q[0] - p[0]) * DEG;
    double U1 = atan((1.0 - f) * tan(p[1] * DEG)), U2 = atan((1.0 - f) * tan(q[1] * DEG));
    double sinU1 = sin(U1), cosU1 = cos(U1), sinU2 = sin(U2), cosU2 = cos(U2);
    double cosSqAlpha, sinSigma, cos2SigmaM, cosSigma, sigma;
    double lambda = L, lambdaP;
    int iterLimit = 100;
    for (;;) {
        double sinLambda = sin(lambda), cosLambda = cos(lambda);
        double t1 = cosU2 * sinLambda, t2 = cosU1 * sinU2 - sinU1 * cosU2 * cosLambda;
        sinSigma = sqrt(t1 * t1 + t2 * t2);
        if (sinSigma == 0.0) {
            if (p[0] == q[0] && p[1] == q[1]) { /* coincident points */
                *out = 0.0;
                return 1;
            }
            return 0; /* antipodal */
        }
        cosSigma = sinU1 * sinU2 + cosU1 * cosU2 * cosLambda;
        sigma = atan2(sinSigma, cosSigma);
        double sinAlpha = cosU1 * cosU2 * sinLambda / sinSigma;
        cosSqAlpha = 1.0 - sinAlpha * sinAlpha;
        if (cosSqAlpha == 0.0) cos2SigmaM = 0.0; /* equatorial geodesics */
        else cos2SigmaM = cosSigma - 2.0 * sinU1 * sinU2 / cosSqAlpha;
        double C = f / 16.0 * cosSqAlpha * (4.0 + f * (4.0 - 3.0 * cosSqAlpha));
        lambdaP = lambda;
        lambda = L + (1.0 - C) * f * sinAlpha * (sigma + C * sinSigma * (cos2SigmaM + C * cosSigma * (-1.0 + 2.0 * cos2SigmaM * cos2SigmaM)));
        if (fabs(lambda - lambdaP) <= 1e-12) break;
        iterLimit -= 1;
        if (iterLimit == 0) break;
    }
    if (iterLimit == 0) return 0;
    double uSq = cosSqAlpha * (a * a - b * b) / (b * b);
    double A = 1.0 + uSq / 16384.0 * (4096.0 + uSq * (-768.0 + uSq * (320.0 - 175.0 * uSq)));
    double B = uSq / 1024.0 * (256.0 + uSq * (-128.0 + uSq * (74.0 - 47.0 * uSq)));
    double deltaSigma = B * sinSigma *
                        (cos2SigmaM + B / 4.0 * (cosSigma * (-1.0 + 2.0 * cos2SigmaM * cos2SigmaM) -
                                                 B / 6.0 * cos2SigmaM * (-3.0 + 4.0 * sinSigma * sinSigma) * (-3.0 + 4.0 * cos2SigmaM * cos2SigmaM)));
    *out = b * A * (sigma - deltaSigma);
    return 1;
}

/* ---- Karney inverse (distance only), WGS84 ---- */
typedef struct kg {
    double a, f, f1, e2, ep2, n, b, etol2;
} kg_t;
static const double K_TINY = 1.4916681462400413e-154; /* sqrt(DBL_MIN) */
static const double K_TOL0 = 2.220446049250313e-16;
#define K_TOL1 (200 * K_TOL0)
#define K_TOL2 1.4901161193847656e-08 /* sqrt(tol0) */
#define K_TOLB (K_TOL0 * K_TOL2)
#define K_XTHRESH (1000 * K_TOL2)
#define K_MAXIT1 20
#define K_MAXIT2 (K_MAXIT1 + 53 + 10)
static double ksq(double x) { return x * x; }
static void knorm2(double *x, double *y) {
    double r = hypot(*x, *y);
    *x /= r;
    *y /= r;
}
static double ksum(double u, double v, double *t) {
    volatile double s = u + v;
    volatile double up = s - v;
    volatile double vpp = s - up;
    up -= u;
    vpp -= v;
    *t = s != 0 ? 0.0 - (up + vpp) : s;
    return s;
}
static double kang_round(double x) {
    const double z = 1.0 / 16.0;
    volatile double y = fabs(x);
    volatile double w = z - y;
    y = w > 0 ? z - w : y;
    return copysign(y, x);
}
static double kang_diff(double x, double y, double *e) {
    double t, d = ksum(remainder(-x, 360.0), remainder(y, 360.0), &t);
    d = ksum(remainder(d, 360.0), t, &t);
    if (d == 0 || fabs(d) == 180.0) d = copysign(d, t == 0 ? y - x : -t);
    *e = t;
    return d;
}
static void ksincosd_q(double r, int q, double *sinx, double *cosx) {
    double s = sin(r), c = cos(r);
    switch ((unsigned)q & 3U) {
    case 0U: *sinx = s, *cosx = c; break;
    case 1U: *sinx = c, *cosx = -s; break;
    case 2U: *sinx = -s, *cosx = -c; break;
    default: *sinx = -c, *cosx = s; break;
    }
    *cosx += 0.0;
}
static void ksincosd(double x, double *sinx, double *cosx) {
    int q = 0;
    double r = remquo(x, 90.0, &q);
    ksincosd_q(r * DEG, q, sinx, cosx);
    if (*sinx == 0) *sinx = copysign(*sinx, x);
}
static void ksincosde(double x, double t, double *sinx, double *cosx) {
    int q = 0;
    double r = remquo(x, 90.0, &q);
    r = kang_round(r + t);
    ksincosd_q(r * DEG, q, sinx, cosx);
    if (*sinx == 0) *sinx = copysign(*sinx, x);
}
/* sum_{l=1..n} c[l] sin(2 l x), Clenshaw */
static double ksin_series(double sinx, double cosx, const double *c, int n) {
    const double *p = c + n + 1;
    double ar = 2 * (cosx - sinx) * (cosx + sinx), y0 = (n & 1) ? *--p : 0, y1 = 0;
    n /= 2;
    while (n--) {
        y1 = ar * y0 - y1 + *--p;
        y0 = ar * y1 - y0 + *--p;
    }
    return 2 * sinx * cosx * y0;
}
static double kA1m1f(double eps) {
    double e2 = eps * eps, t = e2 * (64 + e2 * (4 + e2)) / 256;
    return (t + eps) / (1 - eps);
}
static void kC1f(double eps, double *c) {
    double e2 = eps * eps, d = eps;
    c[1] = d * (-16 + e2 * (6 - e2)) / 32;
    d *= eps;
    c[2] = d * (-128 + e2 * (64 - 9 * e2)) / 2048;
    d *= eps;
    c[3] = d * (9 * e2 - 16) / 768;
    d *= eps;
    c[4] = d * (3 * e2 - 5) / 512;
    d *= eps;
    c[5] = -7 * d / 1280;
    d *= eps;
    c[6] = -7 * d / 2048;
}
static double kA2m1f(double eps) {
    double e2 = eps * eps, t = e2 * (-192 + e2 * (-28 - 11 * e2)) / 256;
    return (t - eps) / (1 + eps);
}
static void kC2f(double eps, double *c) {
    double e2 = eps * eps, d = eps;
    c[1] = d * (16 + e2 * (2 + e2)) / 32;
    d *= eps;
    c[2] = d * (384 + e2 * (64 + 35 * e2)) / 2048;
    d *= eps;
    c[3] = d * (80 + 15 * e2) / 768;
    d *= eps;
    c[4] = d * (35 + 7 * e2) / 512;
    d *= eps;
    c[5] = 63 * d / 1280;
    d *= eps;
    c[6] = 77 * d / 2048;
}
static double kA3f(const kg_t *g, double eps) {
    double n = g->n;
    double c1 = (n - 1) / 2, c2 = (n * (3 * n - 1) - 2) / 8, c3 = ((-n - 3) * n - 1) / 16, c4 = (-2 * n - 3) / 64, c5 = -3.0 / 128;
    return 1 + eps * (c1 + eps * (c2 + eps * (c3 + eps * (c4 + eps * c5))));
}
static void kC3f(const kg_t *g, double eps, double *c) {
    double n = g->n, n2 = n * n, d = eps;
    c[1] = d * ((1 - n) / 4 + eps * ((1 - n2) / 8 + eps * ((3 + 3 * n - n2) / 64 + eps * ((5 + 2 * n) / 128 + eps * (3.0 / 128)))));
    d *= eps;
    c[2] = d * ((2 - 3 * n + n2) / 32 + eps * ((3 - 2 * n - 3 * n2) / 64 + eps * ((3 + n) / 128 + eps * (5.0 / 256))));
    d *= eps;
    c[3] = d * ((5 - 9 * n + 5 * n2) / 192 + eps * ((9 - 10 * n) / 384 + eps * (7.0 / 512)));
    d *= eps;
    c[4] = d * ((7 - 14 * n) / 512 + eps * (7.0 / 512));
    d *= eps;
    c[5] = d * (21.0 / 2560);
}
/* ps12b and/or pm12b may be NULL */
static void klengths(double eps, double sig12, double ssig1, double csig1, double dn1, double ssig2, double csig2, double dn2,
                     double *ps12b, double *pm12b) {
    double Ca[7], Cb[7], m0 = 0, J12 = 0, A1, A2 = 0;
    A1 = kA1m1f(eps);
    kC1f(eps, Ca);
    if (pm12b) {
        A2 = kA2m1f(eps);
        kC2f(eps, Cb);
        m0 = A1 - A2;
        A2 = 1 + A2;
    }
    A1 = 1 + A1;
    if (ps12b) {
        double B1 = ksin_series(ssig2, csig2, Ca, 6) - ksin_series(ssig1, csig1, Ca, 6);
        *ps12b = A1 * (sig12 + B1);
        if (pm12b) {
            double B2 = ksin_series(ssig2, csig2, Cb, 6) - ksin_series(ssig1, csig1, Cb, 6);
            J12 = m0 * sig12 + (A1 * B1 - A2 * B2);
        }
    } else if (pm12b) {
        for (int l = 1; l <= 6; ++l) Cb[l] = A1 * Ca[l] - A2 * Cb[l];
        J12 = m0 * sig12 + (ksin_series(ssig2, csig2, Cb, 6) - ksin_series(ssig1, csig1, Cb, 6));
    }
    if (pm12b) *pm12b = dn2 * (csig1 * ssig2) - dn1 * (ssig1 * csig2) - csig1 * csig2 * J12;
}
static double kastroid(double x, double y) {
    double k, p = ksq(x), q = ksq(y), r = (p + q - 1) / 6;
    if (!(q == 0 && r <= 0)) {
        double S = p * q / 4, r2 = ksq(r), r3 = r * r2, disc = S * (S + 2 * r3), u = r;
        if (disc >= 0) {
            double T3 = S + r3;
            T3 += T3 < 0 ? -sqrt(disc) : sqrt(disc);
            double T = cbrt(T3);
            u += T + (T != 0 ? r2 / T : 0);
        } else {
            double ang = atan2(sqrt(-disc), -(S + r3));
            u += 2 * r * cos(ang / 3);
        }
        double v = sqrt(ksq(u) + q), uv = u < 0 ? q / (v - u) : u + v, w = (uv - q) / (2 * v);
        k = uv / (sqrt(uv + ksq(w)) + w);
    } else
        k = 0;
    return k;
}
static double kinverse_start(const kg_t *g, double sbet1, double cbet1, double dn1, double sbet2, double cbet2, double dn2, double lam12,
                             double slam12, double clam12, double *psalp1, double *pcalp1, double *psalp2, double *pcalp2, double *pdnm) {
    (void)dn1, (void)dn2;
    double sig12 = -1, salp1, calp1, salp2 = 0, calp2 = 0, dnm = 0;
    double sbet12 = sbet2 * cbet1 - cbet2 * sbet1, cbet12 = cbet2 * cbet1 + sbet2 * sbet1;
    double sbet12a = sbet2 * cbet1 + cbet2 * sbet1;
    int shortline = cbet12 >= 0 && sbet12 < 0.5 && cbet2 * lam12 < 0.5;
    double somg12, comg12, ssig12, csig12;
    if (shortline) {
        double sbetm2 = ksq(sbet1 + sbet2), omg12;
        sbetm2 /= sbetm2 + ksq(cbet1 + cbet2);
        dnm = sqrt(1 + g->ep2 * sbetm2);
        omg12 = lam12 / (g->f1 * dnm);
        somg12 = sin(omg12);
        comg12 = cos(omg12);
    } else {
        somg12 = slam12;
        comg12 = clam12;
    }
    salp1 = cbet2 * somg12;
    calp1 = comg12 >= 0 ? sbet12 + cbet2 * sbet1 * ksq(somg12) / (1 + comg12) : sbet12a - cbet2 * sbet1 * ksq(somg12) / (1 - comg12);
    ssig12 = hypot(salp1, calp1);
    csig12 = sbet1 * sbet2 + cbet1 * cbet2 * comg12;
    if (shortline && ssig12 < g->etol2) {
        salp2 = cbet1 * somg12;
        calp2 = sbet12 - cbet1 * sbet2 * (comg12 >= 0 ? ksq(somg12) / (1 + comg12) : 1 - comg12);
        knorm2(&salp2, &calp2);
        sig12 = atan2(ssig12, csig12);
    } else if (fabs(g->n) > 0.1 || csig12 >= 0 || ssig12 >= 6 * fabs(g->n) * M_PI * ksq(cbet1)) {
        /* zeroth order spherical approximation is OK */
    } else {
        double x, y, lamscale, betscale;
        double lam12x = atan2(-slam12, -clam12); /* lam12 - pi */
        {
            double k2 = ksq(sbet1) * g->ep2, eps = k2 / (2 * (1 + sqrt(1 + k2)) + k2);
            lamscale = g->f * cbet1 * kA3f(g, eps) * M_PI;
        }
        betscale = lamscale * cbet1;
        x = lam12x / lamscale;
        y = sbet12a / betscale;
        if (y > -K_TOL1 && x > -1 - K_XTHRESH) { /* strip near cut */
            salp1 = fmin(1.0, -x);
            calp1 = -sqrt(1 - ksq(salp1));
        } else {
            double k = kastroid(x, y);
            double omg12a = lamscale * (-x * k / (1 + k));
            somg12 = sin(omg12a);
            comg12 = -cos(omg12a);
            salp1 = cbet2 * somg12;
            calp1 = sbet12a - cbet2 * sbet1 * ksq(somg12) / (1 - comg12);
        }
    }
    if (!(salp1 <= 0)) knorm2(&salp1, &calp1);
    else {
        salp1 = 1;
        calp1 = 0;
    }
    *psalp1 = salp1;
    *pcalp1 = calp1;
    if (shortline) *pdnm = dnm;
    if (sig12 >= 0) {
        *psalp2 = salp2;
        *pcalp2 = calp2;
    }
    return sig12;
}
static double klambda12(const kg_t *g, double sbet1, double cbet1, double dn1, double sbet2, double cbet2, double dn2, double salp1,
                        double calp1, double slam120, double clam120, double *psig12, double *pssig1, double *pcsig1, double *pssig2,
                        double *pcsig2, double *peps, int diffp, double *pdlam12) {
    double calp2, sig12, ssig1, csig1, ssig2, csig2, eps, domg12, dlam12 = 0;
    double salp0, calp0, somg1, comg1, somg2, comg2, somg12, comg12, lam12, B312, eta, k2, Ca[7];
    if (sbet1 == 0 && calp1 == 0) calp1 = -K_TINY;
    salp0 = salp1 * cbet1;
    calp0 = hypot(calp1, salp1 * sbet1);
    ssig1 = sbet1;
    somg1 = salp0 * sbet1;
    csig1 = comg1 = calp1 * cbet1;
    knorm2(&ssig1, &csig1);
    calp2 = cbet2 != cbet1 || fabs(sbet2) != -sbet1
                ? sqrt(ksq(calp1 * cbet1) + (cbet1 < -sbet1 ? (cbet2 - cbet1) * (cbet1 + cbet2) : (sbet1 - sbet2) * (sbet1 + sbet2))) / cbet2
                : fabs(calp1);
    ssig2 = sbet2;
    somg2 = salp0 * sbet2;
    csig2 = comg2 = calp2 * cbet2;
    knorm2(&ssig2, &csig2);
    sig12 = atan2(fmax(0.0, csig1 * ssig2 - ssig1 * csig2) + 0.0, csig1 * csig2 + ssig1 * ssig2);
    somg12 = fmax(0.0, comg1 * somg2 - somg1 * comg2) + 0.0;
    comg12 = comg1 * comg2 + somg1 * somg2;
    eta = atan2(somg12 * clam120 - comg12 * slam120, comg12 * clam120 + somg12 * slam120);
    k2 = ksq(calp0) * g->ep2;
    eps = k2 / (2 * (1 + sqrt(1 + k2)) + k2);
    kC3f(g, eps, Ca);
    B312 = ksin_series(ssig2, csig2, Ca, 5) - ksin_series(ssig1, csig1, Ca, 5);
    domg12 = -g->f * kA3f(g, eps) * salp0 * (sig12 + B312);
    lam12 = eta + domg12;
    if (diffp) {
        if (calp2 == 0) dlam12 = -2 * g->f1 * dn1 / sbet1;
        else {
            klengths(eps, sig12, ssig1, csig1, dn1, ssig2, csig2, dn2, NULL, &dlam12);
            dlam12 *= g->f1 / (calp2 * cbet2);
        }
    }
    *psig12 = sig12, *pssig1 = ssig1, *pcsig1 = csig1, *pssig2 = ssig2, *pcsig2 = csig2, *peps = eps;
    *pdlam12 = dlam12;
    return lam12;
}
static kg_t kwgs84(void) {
    kg_t g;
    g.a = OG_WGS84_A;
    g.f = 1.0 / 298.257223563;
    g.f1 = 1 - g.f;
    g.e2 = g.f * (2 - g.f);
    g.ep2 = g.e2 / ksq(g.f1);
    g.n = g.f / (2 - g.f);
    g.b = g.a * g.f1;
    g.etol2 = 0.1 * K_TOL2 / sqrt(fmax(0.001, fabs(g.f)) * fmin(1.0, 1 - g.f / 2) / 2);
    return g;
}
/* Geodesic::wgs84().inverse(lat1, lon1, lat2, lon2) -> s12 (metres) */
static double karney_distance(const double *p, const double *q) {
    const kg_t G = kwgs84();
    const kg_t *g = &G;
    double lat1 = p[1], lon1 = p[0], lat2 = q[1], lon2 = q[0];
    double lon12s, lon12 = kang_diff(lon1, lon2, &lon12s);
    int lonsign = signbit(lon12) ? -1 : 1;
    lon12 *= lonsign;
    lon12s *= lonsign;
    double lam12 = lon12 * DEG, slam12, clam12;
    ksincosde(lon12, lon12s, &slam12, &clam12);
    lon12s = (180.0 - lon12) - lon12s;
    lat1 = kang_round(fabs(lat1) > 90 ? NAN : lat1);
    lat2 = kang_round(fabs(lat2) > 90 ? NAN : lat2);
    if (fabs(lat1) < fabs(lat2) || lat2 != lat2) {
        double t = lat1;
        lat1 = lat2;
        lat2 = t;
    }
    int latsign = signbit(lat1) ? 1 : -1;
    lat1 *= latsign;
    lat2 *= latsign;
    double sbet1, cbet1, sbet2, cbet2, s12x = 0, m12x = 0, sig12;
    ksincosd(lat1, &sbet1, &cbet1);
    sbet1 *= g->f1;
    knorm2(&sbet1, &cbet1);
    cbet1 = fmax(K_TINY, cbet1);
    ksincosd(lat2, &sbet2, &cbet2);
    sbet2 *= g->f1;
    knorm2(&sbet2, &cbet2);
    cbet2 = fmax(K_TINY, cbet2);
    if (cbet1 < -sbet1) {
        if (cbet2 == cbet1) sbet2 = copysign(sbet1, sbet2);
    } else {
        if (fabs(sbet2) == -sbet1) cbet2 = cbet1;
    }
    double dn1 = sqrt(1 + g->ep2 * ksq(sbet1)), dn2 = sqrt(1 + g->ep2 * ksq(sbet2));
    int meridian = lat1 == -90 || slam12 == 0;
    if (meridian) {
        double calp1 = clam12, calp2 = 1;
        double ssig1 = sbet1, csig1 = calp1 * cbet1, ssig2 = sbet2, csig2 = calp2 * cbet2;
        sig12 = atan2(fmax(0.0, csig1 * ssig2 - ssig1 * csig2) + 0.0, csig1 * csig2 + ssig1 * ssig2);
        klengths(g->n, sig12, ssig1, csig1, dn1, ssig2, csig2, dn2, &s12x, &m12x);
        if (sig12 < 1 || m12x >= 0) {
            if (sig12 < 3 * K_TINY || (sig12 < K_TOL0 && (s12x < 0 || m12x < 0))) sig12 = m12x = s12x = 0;
            s12x *= g->b;
        } else
            meridian = 0;
    }
    if (!meridian && sbet1 == 0 && (g->f <= 0 || lon12s >= g->f * 180.0)) {
        s12x = g->a * lam12; /* geodesic runs along the equator */
    } else if (!meridian) {
        double salp1, calp1, salp2 = 0, calp2 = 0, dnm = 0;
        sig12 = kinverse_start(g, sbet1, cbet1, dn1, sbet2, cbet2, dn2, lam12, slam12, clam12, &salp1, &calp1, &salp2, &calp2, &dnm);
        if (sig12 >= 0) {
            s12x = sig12 * g->b * dnm; /* short lines */
        } else {
            double ssig1 = 0, csig1 = 0, ssig2 = 0, csig2 = 0, eps = 0;
            unsigned numit = 0;
            double salp1a = K_TINY, calp1a = 1, salp1b = K_TINY, calp1b = -1;
            int tripn = 0, tripb = 0;
            for (;; ++numit) {
                double dv = 0, v = klambda12(g, sbet1, cbet1, dn1, sbet2, cbet2, dn2, salp1, calp1, slam12, clam12, &sig12, &ssig1, &csig1,
                                             &ssig2, &csig2, &eps, numit < K_MAXIT1, &dv);
                if (tripb || !(fabs(v) >= (tripn ? 8 : 1) * K_TOL0) || numit == K_MAXIT2) break;
                if (v > 0 && (numit > K_MAXIT1 || calp1 / salp1 > calp1b / salp1b)) {
                    salp1b = salp1;
                    calp1b = calp1;
                } else if (v < 0 && (numit > K_MAXIT1 || calp1 / salp1 < calp1a / salp1a)) {
                    salp1a = salp1;
                    calp1a = calp1;
                }
                if (numit < K_MAXIT1 && dv > 0) {
                    double dalp1 = -v / dv;
                    if (fabs(dalp1) < M_PI) {
                        double sdalp1 = sin(dalp1), cdalp1 = cos(dalp1), nsalp1 = salp1 * cdalp1 + calp1 * sdalp1;
                        if (nsalp1 > 0) {
                            calp1 = calp1 * cdalp1 - salp1 * sdalp1;
                            salp1 = nsalp1;
                            knorm2(&salp1, &calp1);
                            tripn = fabs(v) <= 16 * K_TOL0;
                            continue;
                        }
                    }
                }
                salp1 = (salp1a + salp1b) / 2;
                calp1 = (calp1a + calp1b) / 2;
                knorm2(&salp1, &calp1);
                tripn = 0;
                tripb = (fabs(salp1a - salp1) + (calp1a - calp1) < K_TOLB || fabs(salp1 - salp1b) + (calp1 - calp1b) < K_TOLB);
            }
            klengths(eps, sig12, ssig1, csig1, dn1, ssig2, csig2, dn2, &s12x, NULL);
            s12x *= g->b;
        }
    }
    return 0.0 + s12x;
}

/* length of one chain under a geodesic metric; *ok cleared when Vincenty fails to converge */
static double chain_geodesic_length(const double *xy, int64_t n, int method, int *ok) {
    double s = 0.0;
    for (int64_t i = 0; i + 1 < n; ++i) {
        const double *p = xy + 2 * i, *q = p + 2;
        double d = 0.0;
        if (method == 1) d = haversine_distance(p, q);
        else if (method == 2) {
            if (!vincenty_distance(p, q, &d)) {
                *ok = 0;
                d = NAN;
            }
        } else
            d = karney_distance(p, q);
        s = s + d;
    }
    return s;
}
/* method: 0 geodesic (Karney), 1 haversine, 2 vincenty (py-geopolars/src/geo.rs:64-67).  Same geometry rules as
 * euclidean_length (exterior ring for polygons).  A row whose Vincenty iteration fails is NaN. */
int og_geodesic_length(const og_array *a, int method, double *out, int threads) {
    int nt = resolve_threads(threads);
    (void)nt;
    if (method < 0 || method > 2) return -1;
    const int64_t *go = a->geom_off, *ro = a->ring_off, *po = a->part_off;
#pragma omp parallel for num_threads(nt) schedule(dynamic, 64)
    for (int64_t i = 0; i < a->n; ++i) {
        double v = 0.0;
        int ok = 1;
        switch (a->type) {
        case OG_LINESTRING: v = chain_geodesic_length(a->xy + 2 * go[i], go[i + 1] - go[i], method, &ok); break;
        case OG_MULTILINESTRING:
            for (int64_t l = go[i]; l < go[i + 1]; ++l) v = v + chain_geodesic_length(a->xy + 2 * ro[l], ro[l + 1] - ro[l], method, &ok);
            break;
        case OG_POLYGON:
            if (go[i + 1] > go[i]) v = chain_geodesic_length(a->xy + 2 * ro[go[i]], ro[go[i] + 1] - ro[go[i]], method, &ok);
            break;
        case OG_MULTIPOLYGON:
            for (int64_t p = go[i]; p < go[i + 1]; ++p)
                if (po[p + 1] > po[p]) v = v + chain_geodesic_length(a->xy + 2 * ro[po[p]], ro[po[p] + 1] - ro[po[p]], method, &ok);
            break;
        default: break;
        }
        out[i] = ok ? v : NAN;
    }
    return 0;
}
/* single pair, for known-answer tests: (lon, lat) degrees */
double og_geodesic_distance(int method, double lon1, double lat1, double lon2, double lat2) {
    double p[2] = {lon1, lat1}, q[2] = {lon2, lat2}, d = NAN;
    if (method == 1) return haversine_distance(p, q);
    if (method == 2) return vincenty_distance(p, q, &d) ? d : NAN;
    return karney_distance(p, q);
}

/* ------------------------------------------------------------------------------------------- */
/* simplify — GeoSeries::simplify geoseries.rs:108-116 ; geo 0.27 simplify.rs (recalled):        */
/* recursive Ramer-Douglas-Peucker with the INITIAL_MIN guard carried through the recursion     */
/* ------------------------------------------------------------------------------------------- */
typedef struct rdp_state {
    const double *xy; /* the chain's coordinates */
    uint8_t *keep;
    double eps;
    int64_t simplified_len;
    int64_t min_len;
} rdp_state;
/* marks what compute_rdp(&rdp_indices[lo..=hi]) returns */
static void compute_rdp(rdp_state *s, int64_t lo, int64_t hi) {
    int64_t len = hi - lo + 1;
    s->keep[lo] = 1;
    s->keep[hi] = 1;
    if (len == 2) return;
    const double *first = s->xy + 2 * lo, *last = s->xy + 2 * hi;
    int64_t farthest_index = 0;
    double farthest_distance = 0.0;
    for (int64_t i = lo + 1; i < hi; ++i) { /* skip(1).take(len - 1): interior points only */
        double d = line_segment_distance(s->xy + 2 * i, first, last);
        if (d >= farthest_distance) {
            farthest_index = i;
            farthest_distance = d;
        }
    }
    if (farthest_distance > s->eps) {
        compute_rdp(s, lo, farthest_index);
        compute_rdp(s, farthest_index, hi);
        return;
    }
    int64_t number_culled = len - 2;
    int64_t new_length = s->simplified_len - number_culled;
    if (new_length < s->min_len) { /* would drop below the minimum: return the slice untouched */
        for (int64_t i = lo; i <= hi; ++i) s->keep[i] = 1;
        return;
    }
    s->simplified_len = new_length;
}
static void rdp_chain(const double *xy, int64_t n, double eps, int64_t min_len, uint8_t *keep) {
    if (n <= 0) return;
    if (!(eps > 0.0) || n <= 2) { /* epsilon <= 0 returns the input; n == 1 panics (debug) / misbehaves in geo: kept as is */
        memset(keep, 1, (size_t)n);
        return;
    }
    memset(keep, 0, (size_t)n);
    rdp_state s = {xy, keep, eps, n, min_len};
    compute_rdp(&s, 0, n - 1);
}
int og_simplify_mask(const og_array *a, double eps, uint8_t *keep, int threads) {
    int nt = resolve_threads(threads);
    (void)nt;
    const int64_t *off;
    int64_t chains, min_len;
    switch (a->type) {
    case OG_LINESTRING: off = a->geom_off, chains = a->n, min_len = 2; break;
    case OG_MULTILINESTRING: off = a->ring_off, chains = a->geom_off[a->n], min_len = 2; break;
    case OG_POLYGON: off = a->ring_off, chains = a->geom_off[a->n], min_len = 4; break;
    case OG_MULTIPOLYGON: off = a->ring_off, chains = a->part_off[a->geom_off[a->n]], min_len = 4; break;
    default: return -1;
    }
#pragma omp parallel for num_threads(nt) schedule(dynamic, 256)
    for (int64_t k = 0; k < chains; ++k) rdp_chain(a->xy + 2 * off[k], off[k + 1] - off[k], eps, min_len, keep + off[k]);
    return 0;
}

/* ------------------------------------------------------------------------------------------- */
/* synthetic data — SURVEY.md §8d RNG                                                          */
/* ------------------------------------------------------------------------------------------- */
double og_splitmix_u(uint64_t seed, uint64_t counter) {
    uint64_t z = seed + 0x9E3779B97F4A7C15ULL * (counter + 1ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z ^= z >> 31;
    return (double)(z >> 11) * 0x1.0p-53;
}
void og_gen_uniform_points(uint64_t stream, int64_t first, int64_t n, double scale, double *out_xy) {
    uint64_t seed = 0xB2000000ULL + stream;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        uint64_t g = (uint64_t)(first + i);
        out_xy[2 * i] = scale * og_splitmix_u(seed, g * 4ULL + 0ULL);
        out_xy[2 * i + 1] = scale * og_splitmix_u(seed, g * 4ULL + 1ULL);
    }
}
