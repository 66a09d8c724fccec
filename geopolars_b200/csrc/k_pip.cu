// k_pip.cu — points-in-polygons broadcast join: the north-star path
// ("100M random points .contains() against 10k 64-vertex polygons", BASELINE.json configs[1]).
//
// Reference semantics (file:line under /root/reference):
//   * candidate generation = bbox overlap with closed intervals (rstar AABB,
//     geopolars/src/spatial_index.rs:74-76, 206-312),
//   * exact test = `poly.contains(point)` for every predicate (spatial_index.rs:89-96), i.e. geo 0.27
//     coordinate_position == Inside: winding number with exact orient2d, boundary => NOT contained
//     (pinned by the 9-point vector at spatial_index.rs:432-484),
//   * output = (lhs_index, rhs_index) pairs (spatial_index.rs:139-157).
//
// B200 design.  The polygon side is small (10 MB) and lives in the 126 MB L2; the point side is a
// 1.6 GB stream.  So the kernel is "one thread per point, streaming LDG.128 of the point, gathers
// from an L2-resident index":
//   grid cell (arithmetic)  -> candidate polygon parts (sorted, so the first hit is the lowest row)
//   part header (64 B)      -> bbox reject, y-bucket parameters
//   y-bucket (arithmetic)   -> short list of 32-byte edge records whose closed y-range can contain p.y
//   exact winding test over that list only.
// Exactness of the pruning: geo's loop only ever acts on an edge when p.y lies in the edge's closed
// y-range.  Buckets are assigned with f(y) = clamp(floor((y - ymin) * inv_h)), a monotone
// non-decreasing function of y in IEEE arithmetic (subtraction, multiplication by a positive constant,
// floor and clamp are all monotone), and an edge is listed in buckets f(ylo)..f(yhi); hence
// ylo <= p.y <= yhi implies the edge is in bucket f(p.y): the bucket list is a superset of the edges
// the reference would act on, and every listed edge is evaluated with the reference's own rule.
// The same argument covers the grid cells (bbox filter).  No epsilon anywhere.
#include <math.h>

#include "common.cuh"
#include "scan.cuh"

namespace gpl {

struct __align__(16) PartHeader {  // 64 bytes, two L2 sectors
    double xmin, ymin, xmax, ymax;
    double inv_h;          // n_buckets / (ymax - ymin), 0 when degenerate
    int32_t n_buckets;
    int32_t bucket_base;   // first bucket of this part in bucket_start[]
    int32_t geom;          // parent row in the polygon array
    int32_t flags;         // bit0: has holes (entries carry ring ids), bit1: valid
    double pad;
};
static_assert(sizeof(PartHeader) == 64, "PartHeader must be 64 bytes");

struct __align__(32) EdgeRec {  // one L2 sector
    double sx, sy, ex, ey;
};

struct GridParams {
    double x0, y0, x1, y1;  // union bbox of valid parts
    double inv_cw, inv_ch;
    int32_t gx, gy;
};

}  // namespace gpl

struct gpl_pip_index {
    gpl_ctx *ctx = nullptr;
    const gpl_array *polys = nullptr;
    int64_t n_parts = 0, n_geoms = 0, n_buckets = 0, n_entries = 0, n_cell_items = 0;
    int32_t gx = 0, gy = 0;
    gpl::PartHeader *parts = nullptr;
    gpl::GridParams *grid = nullptr;  // device copy
    int32_t *cell_start = nullptr;    // gx*gy + 1
    int32_t *cell_items = nullptr;    // part ids, ascending per cell
    int32_t *bucket_start = nullptr;  // n_buckets + 1
    gpl::EdgeRec *entries = nullptr;
    int32_t *entry_ring = nullptr;  // ring index within part (0 = exterior), only if any part has holes
    bool any_holes = false;
    int64_t bytes = 0;
};

namespace gpl {

// ------------------------------------------------------------------------------------------------
// index build
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void part_rings(int type, int64_t part, const int64_t *geom_off, const int64_t *part_off,
                                           int64_t &r0, int64_t &r1) {
    if (type == GPL_POLYGON) {
        r0 = geom_off[part], r1 = geom_off[part + 1];
    } else {
        r0 = part_off[part], r1 = part_off[part + 1];
    }
}

// parent geometry of each part (MULTIPOLYGON): geom g owns parts [geom_off[g], geom_off[g+1])
__global__ void k_part_parent(int64_t n_geoms, const int64_t *__restrict__ geom_off, int32_t *__restrict__ parent) {
    int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (g >= n_geoms) return;
    for (int64_t p = geom_off[g]; p < geom_off[g + 1]; ++p) parent[p] = (int32_t)g;
}

// one warp per part: bbox of the exterior ring, number of buckets
__global__ void __launch_bounds__(256) k_part_headers(int type, int64_t n_parts, const double2 *__restrict__ xy,
                                                      const int64_t *__restrict__ geom_off,
                                                      const int64_t *__restrict__ part_off,
                                                      const int64_t *__restrict__ ring_off,
                                                      const uint8_t *__restrict__ validity,
                                                      const int32_t *__restrict__ parent, PartHeader *__restrict__ parts,
                                                      int32_t *__restrict__ nb_out) {
    const int lane = threadIdx.x & 31;
    int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    for (int64_t p = warp; p < n_parts; p += nwarps) {
        int64_t r0, r1;
        part_rings(type, p, geom_off, part_off, r0, r1);
        int32_t g = parent ? parent[p] : (int32_t)p;
        bool valid = bit_get(validity, g) && r1 > r0;
        double x0 = inf, y0 = inf, x1 = -inf, y1 = -inf;
        int64_t n_slots = 0;
        if (valid) {
            int64_t c0 = ring_off[r0], c1 = ring_off[r0 + 1];
            if (c1 <= c0) valid = false;  // empty exterior: Outside for every point
            for (int64_t c = c0 + lane; c < c1; c += 32) {
                double2 q = xy[c];
                x0 = fmin(x0, q.x), y0 = fmin(y0, q.y), x1 = fmax(x1, q.x), y1 = fmax(y1, q.y);
            }
            n_slots = ring_off[r1] - ring_off[r0];
        }
        x0 = warp_min(x0), y0 = warp_min(y0), x1 = warp_max(x1), y1 = warp_max(y1);
        if (!(x0 <= x1 && y0 <= y1)) valid = false;  // NaN-only exterior
        if (lane == 0) {
            PartHeader h;
            h.xmin = x0, h.ymin = y0, h.xmax = x1, h.ymax = y1;
            // ~2 edge slots per bucket: measured on the config-2 stars this gives ~6.7 listed edges
            // per point against 64 for a brute-force ring walk, at 3.3 records per edge of memory.
            int64_t nb = valid ? (n_slots + 1) / 2 : 0;
            if (nb < 1) nb = valid ? 1 : 0;
            if (nb > 4096) nb = 4096;
            double h_ext = y1 - y0;
            h.inv_h = (valid && h_ext > 0.0 && isfinite(h_ext)) ? (double)nb / h_ext : 0.0;
            if (!isfinite(h.inv_h)) h.inv_h = 0.0;
            h.n_buckets = (int32_t)nb;
            h.bucket_base = 0;
            h.geom = g;
            h.flags = (valid ? 2 : 0) | ((r1 - r0 > 1) ? 1 : 0);
            h.pad = 0.0;
            parts[p] = h;
            nb_out[p] = (int32_t)nb;
        }
    }
}

// single CTA: union bbox of valid parts + grid parameters
__global__ void __launch_bounds__(1024) k_grid_params(const PartHeader *__restrict__ parts, int64_t n_parts, int32_t gx,
                                                      int32_t gy, GridParams *__restrict__ out) {
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    double x0 = inf, y0 = inf, x1 = -inf, y1 = -inf;
    for (int64_t p = threadIdx.x; p < n_parts; p += blockDim.x) {
        PartHeader h = parts[p];
        if (h.flags & 2) x0 = fmin(x0, h.xmin), y0 = fmin(y0, h.ymin), x1 = fmax(x1, h.xmax), y1 = fmax(y1, h.ymax);
    }
    x0 = warp_min(x0), y0 = warp_min(y0), x1 = warp_max(x1), y1 = warp_max(y1);
    __shared__ double s[4][32];
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) s[0][wid] = x0, s[1][wid] = y0, s[2][wid] = x1, s[3][wid] = y1;
    __syncthreads();
    if (wid == 0) {
        x0 = s[0][lane], y0 = s[1][lane], x1 = s[2][lane], y1 = s[3][lane];
        x0 = warp_min(x0), y0 = warp_min(y0), x1 = warp_max(x1), y1 = warp_max(y1);
        if (lane == 0) {
            GridParams g;
            g.x0 = x0, g.y0 = y0, g.x1 = x1, g.y1 = y1;
            double w = x1 - x0, h = y1 - y0;
            g.inv_cw = (w > 0.0 && isfinite(w)) ? (double)gx / w : 0.0;
            g.inv_ch = (h > 0.0 && isfinite(h)) ? (double)gy / h : 0.0;
            if (!isfinite(g.inv_cw)) g.inv_cw = 0.0;
            if (!isfinite(g.inv_ch)) g.inv_ch = 0.0;
            g.gx = gx, g.gy = gy;
            *out = g;
        }
    }
}

// monotone cell / bucket functions (see the exactness argument at the top of the file)
__device__ __forceinline__ int32_t mono_index(double v, double lo, double inv, int32_t n) {
    double t = floor((v - lo) * inv);
    // NaN never reaches here (callers reject points outside the closed bbox first)
    if (!(t > 0.0)) return 0;
    if (t >= (double)n) return n - 1;
    return (int32_t)t;
}

// pass 0 counts, pass 1 fills: one thread per part walks the cells its bbox overlaps
template <int PASS>
__global__ void k_cells(const PartHeader *__restrict__ parts, int64_t n_parts, const GridParams *__restrict__ gp,
                        int32_t *__restrict__ cell_count_or_cursor, int32_t *__restrict__ cell_items) {
    int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= n_parts) return;
    PartHeader h = parts[p];
    if (!(h.flags & 2)) return;
    GridParams g = *gp;
    int32_t cx0 = mono_index(h.xmin, g.x0, g.inv_cw, g.gx), cx1 = mono_index(h.xmax, g.x0, g.inv_cw, g.gx);
    int32_t cy0 = mono_index(h.ymin, g.y0, g.inv_ch, g.gy), cy1 = mono_index(h.ymax, g.y0, g.inv_ch, g.gy);
    for (int32_t cy = cy0; cy <= cy1; ++cy)
        for (int32_t cx = cx0; cx <= cx1; ++cx) {
            int64_t c = (int64_t)cy * g.gx + cx;
            if (PASS == 0) {
                atomicAdd(&cell_count_or_cursor[c], 1);
            } else {
                int32_t pos = atomicAdd(&cell_count_or_cursor[c], 1);
                cell_items[pos] = (int32_t)p;
            }
        }
}

// edge slot c of ring [c0,c1): (c -> c+1), or the implicit closing edge geo's Polygon::new would add
// for an open ring, or the degenerate edge of a 1-coordinate ring.  Returns false for "no edge".
__device__ __forceinline__ bool edge_of_slot(const double2 *__restrict__ xy, int64_t c, int64_t c0, int64_t c1, double2 &s,
                                             double2 &e) {
    s = xy[c];
    if (c + 1 < c1) {
        e = xy[c + 1];
        return true;
    }
    double2 first = xy[c0];
    if (c1 - c0 == 1) {
        e = s;
        return true;
    }
    if (first.x == s.x && first.y == s.y) return false;  // ring already closed
    e = first;
    return true;
}

// one warp per part; pass 0 counts bucket entries, pass 1 writes edge ids (global coord index)
template <int PASS>
__global__ void __launch_bounds__(256) k_buckets(int type, int64_t n_parts, const double2 *__restrict__ xy,
                                                 const int64_t *__restrict__ geom_off,
                                                 const int64_t *__restrict__ part_off,
                                                 const int64_t *__restrict__ ring_off,
                                                 const PartHeader *__restrict__ parts,
                                                 int32_t *__restrict__ count_or_cursor, int64_t *__restrict__ entry_edge) {
    const int lane = threadIdx.x & 31;
    int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t p = warp; p < n_parts; p += nwarps) {
        PartHeader h = parts[p];
        if (!(h.flags & 2)) continue;
        int64_t r0, r1;
        part_rings(type, p, geom_off, part_off, r0, r1);
        for (int64_t r = r0; r < r1; ++r) {
            int64_t c0 = ring_off[r], c1 = ring_off[r + 1];
            for (int64_t c = c0 + lane; c < c1; c += 32) {
                double2 s, e;
                if (!edge_of_slot(xy, c, c0, c1, s, e)) continue;
                // an edge with a NaN ordinate never satisfies geo's comparisons: it contributes nothing
                if (isnan(s.y) || isnan(e.y)) continue;
                double ylo = fmin(s.y, e.y), yhi = fmax(s.y, e.y);
                // holes may stick out of the exterior's bbox: only the part of their y-range inside the
                // bucketed span [ymin,ymax] can hold a queried p.y (queries are bbox-filtered first)
                if (yhi < h.ymin || ylo > h.ymax) continue;
                int32_t b0 = mono_index(fmax(ylo, h.ymin), h.ymin, h.inv_h, h.n_buckets);
                int32_t b1 = mono_index(fmin(yhi, h.ymax), h.ymin, h.inv_h, h.n_buckets);
                for (int32_t b = b0; b <= b1; ++b) {
                    if (PASS == 0) {
                        atomicAdd(&count_or_cursor[h.bucket_base + b], 1);
                    } else {
                        int32_t pos = atomicAdd(&count_or_cursor[h.bucket_base + b], 1);
                        entry_edge[pos] = c;
                    }
                }
            }
        }
    }
}

__global__ void k_set_bucket_base(PartHeader *__restrict__ parts, const int32_t *__restrict__ base, int64_t n_parts) {
    int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p < n_parts) parts[p].bucket_base = base[p];
}

// one thread per segment: insertion sort (lists are a handful of items) — makes cell candidate lists
// ascending by part id and bucket lists ascending by edge id (=> grouped by ring, deterministic).
template <typename T>
__global__ void k_sort_segments(T *__restrict__ items, const int32_t *__restrict__ seg_start, int64_t n_seg) {
    int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (s >= n_seg) return;
    int32_t a = seg_start[s], b = seg_start[s + 1];
    for (int32_t i = a + 1; i < b; ++i) {
        T v = items[i];
        int32_t j = i - 1;
        while (j >= a && items[j] > v) {
            items[j + 1] = items[j];
            --j;
        }
        items[j + 1] = v;
    }
}

// materialise sorted edge ids into 32-byte records (+ ring index within the part when holes exist)
__global__ void k_materialise(int type, int64_t n_parts, const double2 *__restrict__ xy, const int64_t *__restrict__ geom_off,
                              const int64_t *__restrict__ part_off, const int64_t *__restrict__ ring_off,
                              const PartHeader *__restrict__ parts, const int32_t *__restrict__ bucket_start,
                              const int64_t *__restrict__ entry_edge, EdgeRec *__restrict__ entries,
                              int32_t *__restrict__ entry_ring) {
    // one warp per part, lanes over the part's entries
    const int lane = threadIdx.x & 31;
    int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t p = warp; p < n_parts; p += nwarps) {
        PartHeader h = parts[p];
        if (!(h.flags & 2)) continue;
        int64_t r0, r1;
        part_rings(type, p, geom_off, part_off, r0, r1);
        int32_t e0 = bucket_start[h.bucket_base], e1 = bucket_start[h.bucket_base + h.n_buckets];
        for (int32_t k = e0 + lane; k < e1; k += 32) {
            int64_t c = entry_edge[k];
            // ring of coordinate c: rings of a part are few; linear search from the exterior
            int64_t r = r0;
            while (r + 1 < r1 && ring_off[r + 1] <= c) ++r;
            double2 s, e;
            edge_of_slot(xy, c, ring_off[r], ring_off[r + 1], s, e);
            EdgeRec rec;
            rec.sx = s.x, rec.sy = s.y, rec.ex = e.x, rec.ey = e.y;
            entries[k] = rec;
            if (entry_ring) entry_ring[k] = (int32_t)(r - r0);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// the query kernel
// ------------------------------------------------------------------------------------------------
struct IndexView {
    const PartHeader *parts;
    const GridParams *grid;
    const int32_t *cell_start, *cell_items, *bucket_start;
    const EdgeRec *entries;
    const int32_t *entry_ring;
};

// geo coord_pos_relative_to_ring edge rule; returns true when p is on this edge (boundary)
__device__ __forceinline__ bool edge_rule(const EdgeRec &ed, double px, double py, int &wn) {
    if (ed.sy <= py) {
        if (ed.ey >= py) {
            double o = orient2d(ed.sx, ed.sy, ed.ex, ed.ey, px, py);
            if (o > 0.0 && ed.ey != py) wn += 1;
            else if (o == 0.0 && value_in_between(px, ed.sx, ed.ex)) return true;
        }
    } else if (ed.ey <= py) {
        double o = orient2d(ed.sx, ed.sy, ed.ex, ed.ey, px, py);
        if (o < 0.0) wn -= 1;
        else if (o == 0.0 && value_in_between(px, ed.sx, ed.ex)) return true;
    }
    return false;
}

__device__ __forceinline__ EdgeRec load_edge(const EdgeRec *p) {
    const double2 *q = reinterpret_cast<const double2 *>(p);
    double2 a = __ldg(q), b = __ldg(q + 1);
    EdgeRec r;
    r.sx = a.x, r.sy = a.y, r.ex = b.x, r.ey = b.y;
    return r;
}

// Polygon::contains(coord) over the entries of one bucket.
// No holes: Inside <=> not on any listed edge and winding != 0.
// Holes: entries are grouped by ring (ascending); exterior must wind, every hole must not, and no
// ring may have p on its boundary (geo: boundary of exterior or of a hole => not Inside).
__device__ __forceinline__ bool part_contains(const IndexView &ix, const PartHeader &h, double px, double py) {
    int32_t b = mono_index(py, h.ymin, h.inv_h, h.n_buckets);
    int32_t e0 = __ldg(ix.bucket_start + h.bucket_base + b), e1 = __ldg(ix.bucket_start + h.bucket_base + b + 1);
    if (e0 == e1) return false;
    int wn = 0;
    if (!(h.flags & 1)) {
        for (int32_t k = e0; k < e1; ++k) {
            EdgeRec ed = load_edge(ix.entries + k);
            if (edge_rule(ed, px, py, wn)) return false;
        }
        return wn != 0;
    }
    int32_t cur = __ldg(ix.entry_ring + e0);
    if (cur != 0) return false;  // no exterior edge near p.y: winding 0 => Outside
    for (int32_t k = e0; k < e1; ++k) {
        int32_t ring = __ldg(ix.entry_ring + k);
        if (ring != cur) {
            if (cur == 0 ? (wn == 0) : (wn != 0)) return false;
            cur = ring;
            wn = 0;
        }
        EdgeRec ed = load_edge(ix.entries + k);
        if (edge_rule(ed, px, py, wn)) return false;
    }
    return cur == 0 ? (wn != 0) : (wn == 0);
}

// MODE 0: first_id (+ optional count)   MODE 1: write every (point, polygon) pair at pair_off[i]
template <int MODE>
__global__ void __launch_bounds__(256, 3) k_pip_query(IndexView ix, const double2 *__restrict__ pts,
                                                   const uint8_t *__restrict__ pts_validity, int64_t n_pts,
                                                   int32_t *__restrict__ first_id, int32_t *__restrict__ count,
                                                   const int64_t *__restrict__ pair_off, uint64_t *__restrict__ lhs,
                                                   uint64_t *__restrict__ rhs, int64_t point_base) {
    const GridParams g = *ix.grid;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_pts; i += stride) {
        double2 p = __ldcs(pts + i);  // read-once stream: do not pollute L1/L2 residency of the index
        int32_t first = -1, cnt = 0;
        int64_t w = MODE == 1 ? pair_off[i] : 0;
        bool ok = p.x >= g.x0 && p.x <= g.x1 && p.y >= g.y0 && p.y <= g.y1;  // false for NaN (empty point)
        if (ok && pts_validity) ok = bit_get(pts_validity, i);
        if (ok) {
            int32_t cx = mono_index(p.x, g.x0, g.inv_cw, g.gx), cy = mono_index(p.y, g.y0, g.inv_ch, g.gy);
            int64_t c = (int64_t)cy * g.gx + cx;
            int32_t k0 = __ldg(ix.cell_start + c), k1 = __ldg(ix.cell_start + c + 1);
            int32_t last_geom = -1;
            for (int32_t k = k0; k < k1; ++k) {
                int32_t part = __ldg(ix.cell_items + k);
                const double2 *hp = reinterpret_cast<const double2 *>(ix.parts + part);
                double2 lo = __ldg(hp), hi = __ldg(hp + 1);  // bbox: first sector of the header
                if (p.x < lo.x || p.x > hi.x || p.y < lo.y || p.y > hi.y) continue;
                double2 h1 = __ldg(hp + 2);
                double h2 = __ldg(reinterpret_cast<const double *>(hp + 3));
                PartHeader h;
                h.xmin = lo.x, h.ymin = lo.y, h.xmax = hi.x, h.ymax = hi.y;
                h.inv_h = h1.x;
                h.n_buckets = (int32_t)(__double_as_longlong(h1.y) & 0xffffffffLL);
                h.bucket_base = (int32_t)(__double_as_longlong(h1.y) >> 32);
                h.geom = (int32_t)(__double_as_longlong(h2) & 0xffffffffLL);
                h.flags = (int32_t)(__double_as_longlong(h2) >> 32);
                if (h.geom == last_geom) continue;  // MultiPolygon::contains = any part; count rows once
                if (part_contains(ix, h, p.x, p.y)) {
                    last_geom = h.geom;
                    if (MODE == 1) {
                        lhs[w] = (uint64_t)(point_base + i);
                        rhs[w] = (uint64_t)h.geom;
                        ++w;
                    } else {
                        if (first < 0) first = h.geom;
                        ++cnt;
                        if (count == nullptr) break;  // only the first hit is wanted
                    }
                }
            }
        }
        if (MODE == 0) {
            __stcs(first_id + i, first);
            if (count) __stcs(count + i, cnt);
        }
    }
}

__global__ void k_histogram(const int32_t *__restrict__ ids, int64_t n, unsigned long long *__restrict__ counts,
                            int64_t n_polys) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) {
        int32_t id = ids[i];
        if (id >= 0 && id < n_polys) atomicAdd(&counts[id], 1ULL);
    }
}

static IndexView view_of(const gpl_pip_index *idx) {
    IndexView v;
    v.parts = idx->parts, v.grid = idx->grid;
    v.cell_start = idx->cell_start, v.cell_items = idx->cell_items, v.bucket_start = idx->bucket_start;
    v.entries = idx->entries, v.entry_ring = idx->entry_ring;
    return v;
}

static int query_grid(int64_t n) {
    // persistent-style: 148 SMs x 8 resident CTAs of 256 threads, grid-stride over the point stream
    int64_t want = ceil_div(n, 256);
    return (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)kSMs * 8));
}

int pip_query(gpl_ctx *ctx, const gpl_pip_index *idx, const double *pts_dev, const uint8_t *validity_dev, int64_t n,
              int32_t *first_dev, int32_t *count_dev, cudaStream_t stream) {
    if (n == 0) return GPL_OK;
    k_pip_query<0><<<query_grid(n), 256, 0, stream>>>(view_of(idx), reinterpret_cast<const double2 *>(pts_dev), validity_dev,
                                                       n, first_dev, count_dev, nullptr, nullptr, nullptr, 0);
    ctx->launches++;
    GPL_CUDA(cudaGetLastError());
    return GPL_OK;
}

}  // namespace gpl

using namespace gpl;

extern "C" void gpl_pip_index_free(gpl_pip_index *idx) {
    if (!idx) return;
    gpl_ctx *c = idx->ctx;
    c->release(idx->parts);
    c->release(idx->grid);
    c->release(idx->cell_start);
    c->release(idx->cell_items);
    c->release(idx->bucket_start);
    c->release(idx->entries);
    c->release(idx->entry_ring);
    delete idx;
}
extern "C" int64_t gpl_pip_index_bytes(const gpl_pip_index *idx) { return idx ? idx->bytes : 0; }

extern "C" int gpl_pip_index_build(gpl_ctx *ctx, const gpl_array *polys, gpl_pip_index **out) {
    GPL_REQUIRE(ctx && polys && out, GPL_ERR_INVALID_ARG, "gpl_pip_index_build: NULL argument");
    GPL_REQUIRE(polys->type == GPL_POLYGON || polys->type == GPL_MULTIPOLYGON, GPL_ERR_INVALID_TYPE,
                "Expected Polygon or MultiPolygon (found geometry type %d)", polys->type);
    GPL_CUDA(cudaSetDevice(ctx->device));
    const int type = polys->type;
    const int64_t P = type == GPL_POLYGON ? polys->n_geoms : polys->n_parts;
    GPL_REQUIRE(P < (1LL << 31) && polys->n_coords < (1LL << 31), GPL_ERR_UNSUPPORTED,
                "polygon side of a broadcast join is limited to 2^31 parts/coords");
    gpl_pip_index *idx = new gpl_pip_index();
    idx->ctx = ctx, idx->polys = polys, idx->n_parts = P, idx->n_geoms = polys->n_geoms;
    auto fail = [&](int rc) {
        gpl_pip_index_free(idx);
        return rc;
    };
#define TRYF(expr)                         \
    do {                                   \
        int rc__ = (expr);                 \
        if (rc__ != GPL_OK) return fail(rc__); \
    } while (0)
#define CUDAF(expr)                                                                     \
    do {                                                                                \
        cudaError_t e__ = (expr);                                                       \
        if (e__ != cudaSuccess) return fail(cuda_fail(e__, #expr, __FILE__, __LINE__)); \
    } while (0)

    const double2 *xy = reinterpret_cast<const double2 *>(polys->xy);
    int64_t Pa = P > 0 ? P : 1;
    void *q;
    TRYF(ctx->alloc(sizeof(PartHeader) * Pa, &q));
    idx->parts = (PartHeader *)q;
    TRYF(ctx->alloc(sizeof(GridParams), &q));
    idx->grid = (GridParams *)q;

    Scratch<int32_t> parent, nb, base;
    TRYF(nb.get(ctx, (size_t)Pa + 1));
    TRYF(base.get(ctx, (size_t)Pa + 1));
    const int32_t *parent_p = nullptr;
    if (type == GPL_MULTIPOLYGON) {
        TRYF(parent.get(ctx, (size_t)Pa));
        if (polys->n_geoms > 0) {
            k_part_parent<<<(int)ceil_div(polys->n_geoms, 256), 256, 0, ctx->stream>>>(polys->n_geoms, polys->geom_off, parent.p);
            ctx->launches++;
        }
        parent_p = parent.p;
    }
    int wgrid = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(Pa, 8), (int64_t)kSMs * 8));
    k_part_headers<<<wgrid, 256, 0, ctx->stream>>>(type, P, xy, polys->geom_off, polys->part_off, polys->ring_off,
                                                  polys->validity, parent_p, idx->parts, nb.p);
    ctx->launches++;
    CUDAF(cudaGetLastError());

    // grid: about one cell per part, capped so the cell table stays L2-resident
    int64_t G = (int64_t)ceil(sqrt((double)Pa));
    if (G < 1) G = 1;
    if (G > 2048) G = 2048;
    idx->gx = idx->gy = (int32_t)G;
    k_grid_params<<<1, 1024, 0, ctx->stream>>>(idx->parts, P, idx->gx, idx->gy, idx->grid);
    ctx->launches++;

    // cell lists: count -> scan -> fill -> sort
    int64_t n_cells = G * G;
    TRYF(ctx->alloc(sizeof(int32_t) * (n_cells + 1), &q));
    idx->cell_start = (int32_t *)q;
    Scratch<int32_t> cursor;
    TRYF(cursor.get(ctx, (size_t)n_cells + 1));
    CUDAF(cudaMemsetAsync(cursor.p, 0, sizeof(int32_t) * (n_cells + 1), ctx->stream));
    if (P > 0) {
        k_cells<0><<<(int)ceil_div(P, 128), 128, 0, ctx->stream>>>(idx->parts, P, idx->grid, cursor.p, nullptr);
        ctx->launches++;
    }
    Scratch<int64_t> totals;
    TRYF(totals.get(ctx, 4));
    TRYF((exclusive_scan<int32_t, int32_t>(ctx, cursor.p, n_cells, idx->cell_start, totals.p)));
    // bucket bases (no data dependence on the host: sum of n_buckets is bounded by P + n_coords/2)
    TRYF((exclusive_scan<int32_t, int32_t>(ctx, nb.p, P, base.p, totals.p + 1)));
    if (P > 0) {
        k_set_bucket_base<<<(int)ceil_div(P, 256), 256, 0, ctx->stream>>>(idx->parts, base.p, P);
        ctx->launches++;
    }
    int64_t NB_cap = Pa + polys->n_coords / 2 + 1;
    TRYF(ctx->alloc(sizeof(int32_t) * (NB_cap + 1), &q));
    idx->bucket_start = (int32_t *)q;
    Scratch<int32_t> bcursor;
    TRYF(bcursor.get(ctx, (size_t)NB_cap + 1));
    CUDAF(cudaMemsetAsync(bcursor.p, 0, sizeof(int32_t) * (NB_cap + 1), ctx->stream));
    if (P > 0) {
        k_buckets<0><<<wgrid, 256, 0, ctx->stream>>>(type, P, xy, polys->geom_off, polys->part_off, polys->ring_off, idx->parts,
                                                    bcursor.p, nullptr);
        ctx->launches++;
    }
    // total bucket count lives on the device; scanning the capped array is equivalent (tail counts are 0)
    TRYF((exclusive_scan<int32_t, int32_t>(ctx, bcursor.p, NB_cap, idx->bucket_start, totals.p + 2)));

    // the two data-dependent sizes: one small D2H (index build is once per join, not per point)
    int64_t h_tot[3];
    CUDAF(cudaMemcpyAsync(h_tot, totals.p, sizeof(int64_t) * 3, cudaMemcpyDeviceToHost, ctx->stream));
    CUDAF(cudaStreamSynchronize(ctx->stream));
    idx->n_cell_items = h_tot[0];
    idx->n_buckets = h_tot[1];
    idx->n_entries = h_tot[2];
    if (idx->n_entries >= (1LL << 31) || idx->n_cell_items >= (1LL << 31)) {
        set_error("join index too large (%lld edge records, %lld cell items)", (long long)idx->n_entries,
                  (long long)idx->n_cell_items);
        return fail(GPL_ERR_UNSUPPORTED);
    }

    TRYF(ctx->alloc(sizeof(int32_t) * (idx->n_cell_items + 1), &q));
    idx->cell_items = (int32_t *)q;
    TRYF(ctx->alloc(sizeof(EdgeRec) * (idx->n_entries + 1), &q));
    idx->entries = (EdgeRec *)q;
    Scratch<int64_t> entry_edge;
    TRYF(entry_edge.get(ctx, (size_t)idx->n_entries + 1));

    if (P > 0) {
        CUDAF(cudaMemcpyAsync(cursor.p, idx->cell_start, sizeof(int32_t) * n_cells, cudaMemcpyDeviceToDevice, ctx->stream));
        k_cells<1><<<(int)ceil_div(P, 128), 128, 0, ctx->stream>>>(idx->parts, P, idx->grid, cursor.p, idx->cell_items);
        k_sort_segments<int32_t><<<(int)ceil_div(n_cells, 128), 128, 0, ctx->stream>>>(idx->cell_items, idx->cell_start, n_cells);
        CUDAF(cudaMemcpyAsync(bcursor.p, idx->bucket_start, sizeof(int32_t) * NB_cap, cudaMemcpyDeviceToDevice, ctx->stream));
        k_buckets<1><<<wgrid, 256, 0, ctx->stream>>>(type, P, xy, polys->geom_off, polys->part_off, polys->ring_off, idx->parts,
                                                    bcursor.p, entry_edge.p);
        int64_t nbk = idx->n_buckets > 0 ? idx->n_buckets : 1;
        k_sort_segments<int64_t><<<(int)ceil_div(nbk, 128), 128, 0, ctx->stream>>>(entry_edge.p, idx->bucket_start, idx->n_buckets);
        // holes anywhere?  (n_rings > n_parts)
        idx->any_holes = polys->n_rings > P;
        if (idx->any_holes) {
            TRYF(ctx->alloc(sizeof(int32_t) * (idx->n_entries + 1), &q));
            idx->entry_ring = (int32_t *)q;
        }
        k_materialise<<<wgrid, 256, 0, ctx->stream>>>(type, P, xy, polys->geom_off, polys->part_off, polys->ring_off, idx->parts,
                                                     idx->bucket_start, entry_edge.p, idx->entries, idx->entry_ring);
        ctx->launches += 5;
        CUDAF(cudaGetLastError());
    }
    idx->bytes = (int64_t)(sizeof(PartHeader) * Pa + sizeof(int32_t) * (n_cells + 1 + idx->n_cell_items + NB_cap + 1) +
                           sizeof(EdgeRec) * idx->n_entries + (idx->any_holes ? sizeof(int32_t) * idx->n_entries : 0));
    // scratch used by the kernels above goes back to the cache only after they have run
    CUDAF(cudaStreamSynchronize(ctx->stream));
    *out = idx;
    return GPL_OK;
#undef TRYF
#undef CUDAF
}

extern "C" int gpl_contains_join(gpl_ctx *ctx, const gpl_pip_index *idx, const double *points_xy, int64_t n_points,
                                 int32_t *first_id, int32_t *count, int mem) {
    GPL_REQUIRE(ctx && idx && first_id && (points_xy || n_points == 0), GPL_ERR_INVALID_ARG, "gpl_contains_join: NULL argument");
    GPL_CUDA(cudaSetDevice(ctx->device));
    if (n_points == 0) return GPL_OK;
    if (mem == GPL_DEVICE) return pip_query(ctx, idx, points_xy, nullptr, n_points, first_id, count, ctx->stream);
    // host buffers: plain (unpipelined) path — copy in, run, copy out.  gpl_contains_join_host overlaps them.
    Scratch<double> pts;
    Scratch<int32_t> ids, cnt;
    GPL_TRY(pts.get(ctx, (size_t)n_points * 2));
    GPL_TRY(ids.get(ctx, (size_t)n_points));
    if (count) GPL_TRY(cnt.get(ctx, (size_t)n_points));
    GPL_CUDA(cudaMemcpyAsync(pts.p, points_xy, sizeof(double) * 2 * n_points, cudaMemcpyHostToDevice, ctx->stream));
    GPL_TRY(pip_query(ctx, idx, pts.p, nullptr, n_points, ids.p, count ? cnt.p : nullptr, ctx->stream));
    GPL_CUDA(cudaMemcpyAsync(first_id, ids.p, sizeof(int32_t) * n_points, cudaMemcpyDeviceToHost, ctx->stream));
    if (count) GPL_CUDA(cudaMemcpyAsync(count, cnt.p, sizeof(int32_t) * n_points, cudaMemcpyDeviceToHost, ctx->stream));
    GPL_CUDA(cudaStreamSynchronize(ctx->stream));
    return GPL_OK;
}

extern "C" int gpl_contains_join_array(gpl_ctx *ctx, const gpl_pip_index *idx, const gpl_array *points, int32_t *first_id,
                                       int32_t *count, int mem) {
    GPL_REQUIRE(ctx && idx && points && first_id, GPL_ERR_INVALID_ARG, "gpl_contains_join_array: NULL argument");
    GPL_REQUIRE(points->type == GPL_POINT, GPL_ERR_INVALID_TYPE, "Expected Point (found geometry type %d)", points->type);
    GPL_CUDA(cudaSetDevice(ctx->device));
    int64_t n = points->n_geoms;
    if (n == 0) return GPL_OK;
    if (mem == GPL_DEVICE) return pip_query(ctx, idx, points->xy, points->validity, n, first_id, count, ctx->stream);
    Scratch<int32_t> ids, cnt;
    GPL_TRY(ids.get(ctx, (size_t)n));
    if (count) GPL_TRY(cnt.get(ctx, (size_t)n));
    GPL_TRY(pip_query(ctx, idx, points->xy, points->validity, n, ids.p, count ? cnt.p : nullptr, ctx->stream));
    GPL_CUDA(cudaMemcpyAsync(first_id, ids.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, ctx->stream));
    if (count) GPL_CUDA(cudaMemcpyAsync(count, cnt.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, ctx->stream));
    GPL_CUDA(cudaStreamSynchronize(ctx->stream));
    return GPL_OK;
}

extern "C" int gpl_contains_join_pairs(gpl_ctx *ctx, const gpl_pip_index *idx, const double *points_xy, int64_t n_points,
                                       uint64_t *lhs, uint64_t *rhs, int64_t *n_pairs, int mem) {
    GPL_REQUIRE(ctx && idx && n_pairs, GPL_ERR_INVALID_ARG, "gpl_contains_join_pairs: NULL argument");
    GPL_CUDA(cudaSetDevice(ctx->device));
    if (n_points == 0) {
        *n_pairs = 0;
        return GPL_OK;
    }
    Scratch<double> pts;
    const double *pdev = points_xy;
    if (mem == GPL_HOST) {
        GPL_TRY(pts.get(ctx, (size_t)n_points * 2));
        GPL_CUDA(cudaMemcpyAsync(pts.p, points_xy, sizeof(double) * 2 * n_points, cudaMemcpyHostToDevice, ctx->stream));
        pdev = pts.p;
    }
    Scratch<int32_t> ids, cnt;
    Scratch<int64_t> off;
    GPL_TRY(ids.get(ctx, (size_t)n_points));
    GPL_TRY(cnt.get(ctx, (size_t)n_points));
    GPL_TRY(off.get(ctx, (size_t)n_points + 2));
    GPL_TRY(pip_query(ctx, idx, pdev, nullptr, n_points, ids.p, cnt.p, ctx->stream));
    GPL_TRY((exclusive_scan<int32_t, int64_t>(ctx, cnt.p, n_points, off.p, off.p + n_points + 1)));
    int64_t total = 0;
    GPL_CUDA(cudaMemcpyAsync(&total, off.p + n_points + 1, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
    GPL_CUDA(cudaStreamSynchronize(ctx->stream));
    if (lhs == nullptr || rhs == nullptr) {
        *n_pairs = total;
        return GPL_OK;
    }
    GPL_REQUIRE(*n_pairs >= total, GPL_ERR_INVALID_ARG, "pair buffers too small: need %lld, have %lld", (long long)total,
                (long long)*n_pairs);
    *n_pairs = total;
    if (total == 0) return GPL_OK;
    Scratch<uint64_t> dl, dr;
    uint64_t *pl = lhs, *pr = rhs;
    if (mem == GPL_HOST) {
        GPL_TRY(dl.get(ctx, (size_t)total));
        GPL_TRY(dr.get(ctx, (size_t)total));
        pl = dl.p, pr = dr.p;
    }
    k_pip_query<1><<<query_grid(n_points), 256, 0, ctx->stream>>>(view_of(idx), reinterpret_cast<const double2 *>(pdev), nullptr,
                                                                  n_points, nullptr, nullptr, off.p, pl, pr, 0);
    ctx->launches++;
    GPL_CUDA(cudaGetLastError());
    if (mem == GPL_HOST) {
        GPL_CUDA(cudaMemcpyAsync(lhs, pl, sizeof(uint64_t) * total, cudaMemcpyDeviceToHost, ctx->stream));
        GPL_CUDA(cudaMemcpyAsync(rhs, pr, sizeof(uint64_t) * total, cudaMemcpyDeviceToHost, ctx->stream));
    }
    GPL_CUDA(cudaStreamSynchronize(ctx->stream));
    return GPL_OK;
}

// End-to-end path: host points -> HBM -> ids -> host, chunked and double buffered so that the H2D of
// chunk k+1, the kernel of chunk k and the D2H of chunk k-1 overlap (PCIe Gen5 is full duplex).
extern "C" int gpl_contains_join_host(gpl_ctx *ctx, const gpl_pip_index *idx, const double *points_xy_host, int64_t n_points,
                                      int32_t *first_id_host, int64_t chunk_points) {
    GPL_REQUIRE(ctx && idx && (n_points == 0 || (points_xy_host && first_id_host)), GPL_ERR_INVALID_ARG,
                "gpl_contains_join_host: NULL argument");
    GPL_CUDA(cudaSetDevice(ctx->device));
    if (n_points == 0) return GPL_OK;
    if (chunk_points <= 0) chunk_points = 4 << 20;  // 64 MiB of points per chunk
    if (chunk_points > n_points) chunk_points = n_points;
    if (!ctx->copy_in) GPL_CUDA(cudaStreamCreateWithFlags(&ctx->copy_in, cudaStreamNonBlocking));
    if (!ctx->copy_out) GPL_CUDA(cudaStreamCreateWithFlags(&ctx->copy_out, cudaStreamNonBlocking));
    constexpr int NBUF = 3;
    Scratch<double> pts[NBUF];
    Scratch<int32_t> ids[NBUF];
    cudaEvent_t in_done[NBUF], k_done[NBUF], out_done[NBUF];
    for (int b = 0; b < NBUF; ++b) {
        GPL_TRY(pts[b].get(ctx, (size_t)chunk_points * 2));
        GPL_TRY(ids[b].get(ctx, (size_t)chunk_points));
        GPL_CUDA(cudaEventCreateWithFlags(&in_done[b], cudaEventDisableTiming));
        GPL_CUDA(cudaEventCreateWithFlags(&k_done[b], cudaEventDisableTiming));
        GPL_CUDA(cudaEventCreateWithFlags(&out_done[b], cudaEventDisableTiming));
    }
    // the scratch blocks may have been used by earlier work on ctx->stream: order the copy streams after it
    cudaEvent_t start;
    GPL_CUDA(cudaEventCreateWithFlags(&start, cudaEventDisableTiming));
    GPL_CUDA(cudaEventRecord(start, ctx->stream));
    GPL_CUDA(cudaStreamWaitEvent(ctx->copy_in, start, 0));
    GPL_CUDA(cudaStreamWaitEvent(ctx->copy_out, start, 0));
    int rc = GPL_OK;
    int64_t n_chunks = ceil_div(n_points, chunk_points);
    for (int64_t c = 0; c < n_chunks && rc == GPL_OK; ++c) {
        int b = (int)(c % NBUF);
        int64_t lo = c * chunk_points, n = std::min(chunk_points, n_points - lo);
        if (c >= NBUF) {
            cudaStreamWaitEvent(ctx->copy_in, k_done[b], 0);    // points buffer free once its kernel ran
            cudaStreamWaitEvent(ctx->stream, out_done[b], 0);   // ids buffer free once its D2H ran
        }
        cudaMemcpyAsync(pts[b].p, points_xy_host + 2 * lo, sizeof(double) * 2 * n, cudaMemcpyHostToDevice, ctx->copy_in);
        cudaEventRecord(in_done[b], ctx->copy_in);
        cudaStreamWaitEvent(ctx->stream, in_done[b], 0);
        rc = pip_query(ctx, idx, pts[b].p, nullptr, n, ids[b].p, nullptr, ctx->stream);
        cudaEventRecord(k_done[b], ctx->stream);
        cudaStreamWaitEvent(ctx->copy_out, k_done[b], 0);
        cudaMemcpyAsync(first_id_host + lo, ids[b].p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, ctx->copy_out);
        cudaEventRecord(out_done[b], ctx->copy_out);
    }
    cudaError_t e1 = cudaStreamSynchronize(ctx->copy_out);
    cudaError_t e2 = cudaStreamSynchronize(ctx->stream);
    cudaError_t e3 = cudaStreamSynchronize(ctx->copy_in);
    for (int b = 0; b < NBUF; ++b) {
        cudaEventDestroy(in_done[b]);
        cudaEventDestroy(k_done[b]);
        cudaEventDestroy(out_done[b]);
    }
    cudaEventDestroy(start);
    if (rc != GPL_OK) return rc;
    GPL_CUDA(e1);
    GPL_CUDA(e2);
    GPL_CUDA(e3);
    return GPL_OK;
}

extern "C" int gpl_join_histogram(gpl_ctx *ctx, const int32_t *first_id, int64_t n_points, uint64_t *counts, int64_t n_polygons,
                                  int mem) {
    GPL_REQUIRE(ctx && counts && (first_id || n_points == 0), GPL_ERR_INVALID_ARG, "gpl_join_histogram: NULL argument");
    GPL_REQUIRE(mem == GPL_DEVICE, GPL_ERR_UNSUPPORTED, "gpl_join_histogram: device buffers only");
    GPL_CUDA(cudaSetDevice(ctx->device));
    if (n_points == 0) return GPL_OK;
    GPL_LAUNCH(ctx, k_histogram, query_grid(n_points), 256, 0, first_id, n_points, reinterpret_cast<unsigned long long *>(counts),
               n_polygons);
    return GPL_OK;
}
