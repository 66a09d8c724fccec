/*
 * geo_oracle.h — CPU oracle for the GeoSeries hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is a plain-C restatement of the arithmetic the reference delegates to the un-vendored
 * crates geo 0.27.0 / geo-types 0.7.12 / robust 1.1.0 (Cargo.lock:986-988, 1003-1004, 2251-2252 of
 * the reference).  Their source is NOT under /root/reference and no Rust toolchain exists in this
 * image, so every formula is restated from the crates' published algorithms ("recalled").
 *
 * PARITY STATUS: pinned only for contains() (the 9-point vector in the reference's
 * geopolars/src/spatial_index.rs:432-484).  All other ops: PARITY UNPINNED — the reference holds no
 * numeric assertion for them.  Boolean/index results are additionally refereed by
 * an exact-rational Python checker (oracle/exact.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * link or call this file.  The product (geopolars_b200/) never does.
 *
 * Layout: GeoArrow nesting with int64 offsets, interleaved xy doubles.
 *   POINT            xy[n]
 *   LINESTRING       geom_off[n+1] -> coords
 *   POLYGON          geom_off[n+1] -> rings ; ring_off[] -> coords
 *   MULTIPOINT       geom_off[n+1] -> coords
 *   MULTILINESTRING  geom_off[n+1] -> lines ; ring_off[] -> coords
 *   MULTIPOLYGON     geom_off[n+1] -> polygons ; part_off[] -> rings ; ring_off[] -> coords
 * Type codes follow py-geopolars/python/geopolars/enums.py:4-15.
 */
#ifndef GEO_ORACLE_H
#define GEO_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    OG_MISSING = -1,
    OG_POINT = 0,
    OG_LINESTRING = 1,
    OG_LINEARRING = 2,
    OG_POLYGON = 3,
    OG_MULTIPOINT = 4,
    OG_MULTILINESTRING = 5,
    OG_MULTIPOLYGON = 6
};

typedef struct og_array {
    int32_t type;
    int64_t n;               /* geometries */
    const double *xy;        /* interleaved */
    const int64_t *geom_off; /* see header comment */
    const int64_t *part_off;
    const int64_t *ring_off;
    const uint8_t *valid;    /* Arrow LSB bitmap or NULL */
} og_array;

/* GeoSeries::geodesic_length (geoseries.rs:52-58): method 0 geodesic (Karney), 1 haversine, 2 vincenty; metres;
 * coordinates are (lon, lat) degrees.  NaN where Vincenty does not converge. */
int og_geodesic_length(const og_array *arr, int method, double *out, int threads);
double og_geodesic_distance(int method, double lon1, double lat1, double lon2, double lat2);

/* GeoSeries::simplify (geoseries.rs:108-116): keep[c] = 1 for the coordinates geo's RDP retains (LineString /
 * MultiLineString minimum 2 points, polygon rings minimum 4); returns -1 for other types */
int og_simplify_mask(const og_array *arr, double eps, uint8_t *keep, int threads);

/* robust 1.1.0 orient2d (Shewchuk adaptive).  Only the sign is consumed anywhere on the path. */
double og_orient2d(double ax, double ay, double bx, double by, double cx, double cy);
/* statistics: how many calls fell through to the adaptive stage since process start */
int64_t og_orient2d_adapt_calls(void);

/* geo AffineTransform::apply — x' = a*x + b*y + xoff ; y' = d*x + e*y + yoff (no FMA) */
void og_affine_transform(const double *xy, int64_t n_coords, double a, double b, double xoff, double d,
                         double e, double yoff, double *out_xy, int threads);

/* geo Area::unsigned_area per geometry */
void og_area(const og_array *arr, double *out, int threads);
/* geo Centroid per geometry; out_valid[i]=0 when the geometry has no centroid (empty) */
void og_centroid(const og_array *arr, double *out_xy, uint8_t *out_valid, int threads);
/* geo BoundingRect; out = minx,miny,maxx,maxy per geometry; out_valid 0 for empty */
void og_envelope(const og_array *arr, double *out4, uint8_t *out_valid, int threads);
/* geo EuclideanLength (LineString / MultiLineString; polygons: exterior ring per geoseries.rs:35-41) */
void og_euclidean_length(const og_array *arr, double *out, int threads);

/* coordinate position of (px,py) w.r.t. geometry i of a POLYGON / MULTIPOLYGON array:
 * 0 outside, 1 on boundary, 2 inside (geo CoordinatePosition) */
int og_coord_position(const og_array *polys, int64_t i, double px, double py);
/* geo Contains<Coord> for Polygon / MultiPolygon: inside only */
int og_contains_point(const og_array *polys, int64_t i, double px, double py);

/* points x polygons broadcast join (semantics of spatial_index.rs:89-96):
 * first_id[p] = lowest polygon row containing point p or -1 ; count[p] = number of containing rows.
 * use_grid!=0 prunes candidates with a uniform bbox grid (index + exact test, as the reference's
 * r-tree does); result is identical either way. */
void og_contains_join(const og_array *polys, const double *pts_xy, int64_t n_pts, int32_t *first_id,
                      int32_t *count, int use_grid, int threads);

/* row-wise intersects for every pair of Point / MultiPoint / LineString / MultiLineString / Polygon /
 * MultiPolygon arrays (geo Intersects; spatial_index.rs:102-123 call sites) -> 0/1 bytes */
void og_intersects_rowwise(const og_array *a, const og_array *b, uint8_t *out, int threads);
/* row-wise `a[i] contains point i` for (Multi)Polygon and (Multi)LineString rows (spatial_index.rs:91-96,125-135) */
/* (Multi)Polygon.contains(Polygon) row-wise (spatial_index.rs:99-110; geo relate().is_contains(), recalled): 0/1 bytes */
int og_contains_polygon_rowwise(const og_array *a, const og_array *b, uint8_t *out, int threads);
void og_contains_rowwise(const og_array *a, const double *pts_xy, const uint8_t *pts_valid, uint8_t *out, int threads);
/* row-wise euclidean distance (geo EuclideanDistance) for Point/LineString/Polygon pairs */
int og_distance_rowwise(const og_array *a, const og_array *b, double *out, int threads);

/* geo ConvexHull (quick_hull) of geometry i: writes closed ring into out_xy (capacity cap coords),
 * returns number of coords written (or needed if > cap). */
int64_t og_convex_hull_one(const og_array *arr, int64_t i, double *out_xy, int64_t cap);
/* all geometries: out_off[n+1] filled; out_xy may be NULL to size only. returns total coords */
int64_t og_convex_hull(const og_array *arr, int64_t *out_off, double *out_xy, int threads);

/* synthetic-data generator shared by tests and bench: splitmix64 uniform */
double og_splitmix_u(uint64_t seed, uint64_t counter);
void og_gen_uniform_points(uint64_t stream, int64_t first, int64_t n, double scale, double *out_xy);

int og_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
