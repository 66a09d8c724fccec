// arrow_wkb.cu — Arrow C Data Interface import/export, the same structs the reference moves across its own
// FFI (py-geopolars/src/ffi.rs:14-49).  geoarrow nested layouts are re-based on the host and uploaded with one
// H2D copy per buffer; WKB `binary` / `large_binary` columns are handed to the GPU codec (k_wkb.cu).
// Pure host C++ (compiled by nvcc for convenience).
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "common.cuh"

// ---- Arrow C Data Interface (https://arrow.apache.org/docs/format/CDataInterface.html) ---------------
struct ArrowSchema {
    const char *format;
    const char *name;
    const char *metadata;
    int64_t flags;
    int64_t n_children;
    struct ArrowSchema **children;
    struct ArrowSchema *dictionary;
    void (*release)(struct ArrowSchema *);
    void *private_data;
};
struct ArrowArray {
    int64_t length;
    int64_t null_count;
    int64_t offset;
    int64_t n_buffers;
    int64_t n_children;
    const void **buffers;
    struct ArrowArray **children;
    struct ArrowArray *dictionary;
    void (*release)(struct ArrowArray *);
    void *private_data;
};

namespace gpl {

struct HostGeo {  // GeoArrow buffers on the host, int64 offsets, interleaved xy
    int type = GPL_MISSING;
    std::vector<double> xy;
    std::vector<int64_t> geom, part, ring;
    std::vector<uint8_t> valid;  // bitmap, empty = all valid
    int64_t n = 0;
};

static int upload(gpl_ctx *ctx, const HostGeo &h, gpl_array **out) {
    gpl_buffers b;
    memset(&b, 0, sizeof(b));
    b.geom_type = h.type;
    b.offset_width = 64;
    b.mem = GPL_HOST;
    b.n_geoms = h.n;
    b.n_parts = h.part.empty() ? 0 : (int64_t)h.part.size() - 1;
    b.n_rings = h.ring.empty() ? 0 : (int64_t)h.ring.size() - 1;
    b.n_coords = (int64_t)h.xy.size() / 2;
    static const double dummy[2] = {0, 0};
    b.x = h.xy.empty() ? dummy : h.xy.data();
    b.geom_offsets = h.geom.empty() ? nullptr : h.geom.data();
    b.part_offsets = h.part.empty() ? nullptr : h.part.data();
    b.ring_offsets = h.ring.empty() ? nullptr : h.ring.data();
    b.validity = h.valid.empty() ? nullptr : h.valid.data();
    return gpl_array_from_buffers(ctx, &b, out);
}

static int fetch(gpl_ctx *ctx, const gpl_array *a, HostGeo &h);

}  // namespace gpl

using namespace gpl;

namespace gpl {
static int fetch(gpl_ctx *ctx, const gpl_array *a, HostGeo &h) {
    h.type = a->type;
    h.n = a->n_geoms;
    h.xy.resize((size_t)a->n_coords * 2);
    if (a->geom_off) h.geom.resize((size_t)a->n_geoms + 1);
    if (a->part_off) h.part.resize((size_t)a->n_parts + 1);
    if (a->ring_off) h.ring.resize((size_t)a->n_rings + 1);
    if (a->validity) h.valid.resize((size_t)(a->n_geoms + 7) / 8);
    return gpl_array_copy_out(ctx, a, h.xy.data(), h.geom.empty() ? nullptr : h.geom.data(), h.part.empty() ? nullptr : h.part.data(),
                              h.ring.empty() ? nullptr : h.ring.data(), h.valid.empty() ? nullptr : h.valid.data(), GPL_HOST);
}
}  // namespace gpl

// ---- Arrow C Data Interface: import ------------------------------------------------------------------
namespace gpl {

static std::string meta_value(const char *md, const char *key) {
    if (!md) return "";
    const char *p = md;
    int32_t n;
    memcpy(&n, p, 4);
    p += 4;
    for (int32_t i = 0; i < n; ++i) {
        int32_t kl, vl;
        memcpy(&kl, p, 4);
        p += 4;
        std::string k(p, (size_t)kl);
        p += kl;
        memcpy(&vl, p, 4);
        p += 4;
        std::string v(p, (size_t)vl);
        p += vl;
        if (k == key) return v;
    }
    return "";
}
static bool is_list(const char *f) { return f && f[0] == '+' && (f[1] == 'l' || f[1] == 'L') && f[2] == 0; }
static bool is_coord(const ArrowSchema *s) {
    if (!s || !s->format) return false;
    if (strcmp(s->format, "+w:2") == 0) return s->n_children == 1 && s->children[0]->format && strcmp(s->children[0]->format, "g") == 0;
    if (strcmp(s->format, "+s") == 0)
        return s->n_children >= 2 && strcmp(s->children[0]->format, "g") == 0 && strcmp(s->children[1]->format, "g") == 0;
    return false;
}
static bool bit(const uint8_t *bm, int64_t i) { return !bm || ((bm[i >> 3] >> (i & 7)) & 1); }

// read offsets [lo, lo+len] of a list level as int64, rebased to start at 0; returns child range
static void read_offsets(const ArrowSchema *s, const ArrowArray *a, int64_t lo, int64_t len, std::vector<int64_t> &out, int64_t &c0,
                         int64_t &c1) {
    out.resize((size_t)len + 1);
    const int64_t base = a->offset + lo;
    if (s->format[1] == 'l') {
        const int32_t *o = static_cast<const int32_t *>(a->buffers[1]) + base;
        c0 = len ? o[0] : (a->buffers[1] ? o[0] : 0);
        for (int64_t i = 0; i <= len; ++i) out[i] = (int64_t)o[i] - c0;
        c1 = c0 + out[len];
    } else {
        const int64_t *o = static_cast<const int64_t *>(a->buffers[1]) + base;
        c0 = o[0];
        for (int64_t i = 0; i <= len; ++i) out[i] = o[i] - c0;
        c1 = c0 + out[len];
    }
}
static void read_coords(const ArrowSchema *s, const ArrowArray *a, int64_t c0, int64_t c1, std::vector<double> &xy) {
    const int64_t n = c1 - c0;
    xy.resize((size_t)n * 2);
    if (s->format[1] == 'w') {
        const ArrowArray *v = a->children[0];
        const double *d = static_cast<const double *>(v->buffers[1]) + v->offset + 2 * (a->offset + c0);
        if (n) memcpy(xy.data(), d, sizeof(double) * 2 * n);
    } else {
        const ArrowArray *ax = a->children[0], *ay = a->children[1];
        const double *x = static_cast<const double *>(ax->buffers[1]) + ax->offset + a->offset + c0;
        const double *y = static_cast<const double *>(ay->buffers[1]) + ay->offset + a->offset + c0;
        for (int64_t i = 0; i < n; ++i) xy[2 * i] = x[i], xy[2 * i + 1] = y[i];
    }
}

}  // namespace gpl

extern "C" int gpl_array_import_arrow(gpl_ctx *ctx, const void *array, const void *schema, gpl_array **out) {
    GPL_REQUIRE(ctx && array && schema && out, GPL_ERR_INVALID_ARG, "gpl_array_import_arrow: NULL argument");
    const ArrowArray *a = static_cast<const ArrowArray *>(array);
    const ArrowSchema *s = static_cast<const ArrowSchema *>(schema);
    GPL_REQUIRE(s->format, GPL_ERR_INVALID_ARG, "schema has no format");
    const int64_t n = a->length;
    // validity of the top level, re-based to bit 0
    std::vector<uint8_t> valid;
    const uint8_t *bm = a->n_buffers > 0 ? static_cast<const uint8_t *>(a->buffers[0]) : nullptr;
    if (bm && a->null_count != 0) {
        valid.assign((size_t)(n + 7) / 8, 0);
        for (int64_t i = 0; i < n; ++i)
            if (bit(bm, a->offset + i)) valid[i >> 3] |= (uint8_t)(1u << (i & 7));
    }
    // WKB column ("z" binary / "Z" large binary): what every bundled fixture of the reference holds
    // (decoded on the GPU: k_wkb.cu)
    if (strcmp(s->format, "z") == 0 || strcmp(s->format, "Z") == 0) {
        const uint8_t *data = static_cast<const uint8_t *>(a->buffers[2]);
        const bool wide = s->format[0] == 'Z';
        const void *o = wide ? static_cast<const void *>(static_cast<const int64_t *>(a->buffers[1]) + a->offset)
                             : static_cast<const void *>(static_cast<const int32_t *>(a->buffers[1]) + a->offset);
        return gpl_wkb_decode(ctx, data, o, wide ? 64 : 32, valid.empty() ? nullptr : valid.data(), n, GPL_HOST, out);
    }
    // geoarrow nested layouts: walk the list levels down to the coordinate array
    const ArrowSchema *ls[3];
    const ArrowArray *la[3];
    int depth = 0;
    const ArrowSchema *cs = s;
    const ArrowArray *ca = a;
    while (is_list(cs->format)) {
        GPL_REQUIRE(depth < 3 && cs->n_children == 1, GPL_ERR_INVALID_TYPE, "unsupported nesting in geometry column");
        ls[depth] = cs, la[depth] = ca;
        ++depth;
        cs = cs->children[0];
        ca = ca->children[0];
    }
    GPL_REQUIRE(is_coord(cs), GPL_ERR_INVALID_TYPE,
                "Expected a geoarrow layout (FixedSizeList<f64>[2] or Struct{x,y} coordinates) or WKB binary (found format '%s')",
                cs->format ? cs->format : "?");
    std::string ext = meta_value(s->metadata, "ARROW:extension:name");
    int type;
    if (depth == 0) type = GPL_POINT;
    else if (depth == 1) {
        bool mp = ext == "geoarrow.multipoint" || (ext.empty() && s->children[0]->name && strcmp(s->children[0]->name, "points") == 0);
        type = mp ? GPL_MULTIPOINT : GPL_LINESTRING;
    } else if (depth == 2) {
        bool ml = ext == "geoarrow.multilinestring" ||
                  (ext.empty() && s->children[0]->name && strcmp(s->children[0]->name, "linestrings") == 0);
        type = ml ? GPL_MULTILINESTRING : GPL_POLYGON;
    } else type = GPL_MULTIPOLYGON;

    HostGeo h;
    h.type = type;
    h.n = n;
    h.valid = valid;
    int64_t c0 = 0, c1 = n;
    std::vector<int64_t> *levels[3] = {&h.geom, nullptr, nullptr};
    if (depth == 2) levels[1] = &h.ring;
    if (depth == 3) levels[1] = &h.part, levels[2] = &h.ring;
    for (int d = 0; d < depth; ++d) {
        int64_t n0, n1;
        read_offsets(ls[d], la[d], c0, c1 - c0, *levels[d], n0, n1);
        c0 = n0, c1 = n1;
    }
    read_coords(cs, ca, c0, c1, h.xy);
    return upload(ctx, h, out);
}

// ---- Arrow C Data Interface: export ------------------------------------------------------------------
namespace gpl {

struct Owned {  // private_data of exported arrays/schemas: everything is freed by the release callback
    std::vector<void *> mallocs;
    std::vector<ArrowArray *> child_arrays;
    std::vector<ArrowSchema *> child_schemas;
    std::vector<const void *> buffers;
    std::vector<ArrowArray *> children_ptrs;
    std::vector<ArrowSchema *> schema_children_ptrs;
    std::string format, name, metadata;
};
static void release_array(ArrowArray *a) {
    if (!a || !a->release) return;
    Owned *o = static_cast<Owned *>(a->private_data);
    for (ArrowArray *c : o->child_arrays) {
        if (c->release) c->release(c);
        free(c);
    }
    for (void *m : o->mallocs) free(m);
    delete o;
    a->release = nullptr;
}
static void release_schema(ArrowSchema *s) {
    if (!s || !s->release) return;
    Owned *o = static_cast<Owned *>(s->private_data);
    for (ArrowSchema *c : o->child_schemas) {
        if (c->release) c->release(c);
        free(c);
    }
    delete o;
    s->release = nullptr;
}
static void make_schema(ArrowSchema *s, const char *format, const char *name, const std::string &metadata,
                        std::vector<ArrowSchema *> children, bool nullable) {
    Owned *o = new Owned();
    o->format = format, o->name = name ? name : "", o->metadata = metadata;
    o->child_schemas = children;
    o->schema_children_ptrs = children;
    memset(s, 0, sizeof(*s));
    s->format = o->format.c_str();
    s->name = o->name.c_str();
    s->metadata = o->metadata.empty() ? nullptr : o->metadata.data();
    s->flags = nullable ? 2 : 0;  // ARROW_FLAG_NULLABLE
    s->n_children = (int64_t)children.size();
    s->children = o->schema_children_ptrs.empty() ? nullptr : o->schema_children_ptrs.data();
    s->release = release_schema;
    s->private_data = o;
}
static void make_array(ArrowArray *a, int64_t length, int64_t null_count, std::vector<const void *> buffers, std::vector<void *> owned,
                       std::vector<ArrowArray *> children) {
    Owned *o = new Owned();
    o->buffers = buffers;
    o->mallocs = owned;
    o->child_arrays = children;
    o->children_ptrs = children;
    memset(a, 0, sizeof(*a));
    a->length = length;
    a->null_count = null_count;
    a->n_buffers = (int64_t)o->buffers.size();
    a->buffers = o->buffers.data();
    a->n_children = (int64_t)children.size();
    a->children = o->children_ptrs.empty() ? nullptr : o->children_ptrs.data();
    a->release = release_array;
    a->private_data = o;
}
static std::string ext_metadata(const char *ext_name) {
    std::string md;
    auto put32 = [&](int32_t v) { md.append(reinterpret_cast<const char *>(&v), 4); };
    auto put = [&](const std::string &k, const std::string &v) {
        put32((int32_t)k.size());
        md += k;
        put32((int32_t)v.size());
        md += v;
    };
    put32(2);
    put("ARROW:extension:name", ext_name);
    put("ARROW:extension:metadata", "{}");
    return md;
}
template <typename T>
static T *dup(const T *src, size_t n) {
    T *p = static_cast<T *>(malloc(sizeof(T) * (n ? n : 1)));
    if (n) memcpy(p, src, sizeof(T) * n);
    return p;
}
static int64_t count_nulls(const std::vector<uint8_t> &bm, int64_t n) {
    if (bm.empty()) return 0;
    int64_t c = 0;
    for (int64_t i = 0; i < n; ++i) c += !((bm[i >> 3] >> (i & 7)) & 1);
    return c;
}

}  // namespace gpl

extern "C" int gpl_array_export_arrow(gpl_ctx *ctx, const gpl_array *a, void *out_array, void *out_schema) {
    GPL_REQUIRE(ctx && a && out_array && out_schema, GPL_ERR_INVALID_ARG, "gpl_array_export_arrow: NULL argument");
    HostGeo h;
    GPL_TRY(fetch(ctx, a, h));
    const int64_t nc = (int64_t)h.xy.size() / 2;
    // coordinates: FixedSizeList<f64>[2] named "xy"
    ArrowArray *vals = static_cast<ArrowArray *>(calloc(1, sizeof(ArrowArray)));
    double *xy = dup(h.xy.data(), h.xy.size());
    make_array(vals, nc * 2, 0, {nullptr, xy}, {xy}, {});
    ArrowSchema *vals_s = static_cast<ArrowSchema *>(calloc(1, sizeof(ArrowSchema)));
    make_schema(vals_s, "g", "xy", "", {}, false);
    ArrowArray *cur = static_cast<ArrowArray *>(calloc(1, sizeof(ArrowArray)));
    ArrowSchema *cur_s = static_cast<ArrowSchema *>(calloc(1, sizeof(ArrowSchema)));
    int64_t cur_len = nc;
    // levels from the innermost list outwards
    struct Level {
        const std::vector<int64_t> *off;
        const char *child_name;
    };
    std::vector<Level> lv;
    const char *ext = "geoarrow.point";
    switch (h.type) {  // (push_back rather than brace assignment: GCC 13 reports a spurious -Wnonnull on the latter)
    case GPL_POINT: break;
    case GPL_LINESTRING: lv.push_back({&h.geom, "vertices"}), ext = "geoarrow.linestring"; break;
    case GPL_MULTIPOINT: lv.push_back({&h.geom, "points"}), ext = "geoarrow.multipoint"; break;
    case GPL_POLYGON: lv.push_back({&h.ring, "vertices"}), lv.push_back({&h.geom, "rings"}), ext = "geoarrow.polygon"; break;
    case GPL_MULTILINESTRING:
        lv.push_back({&h.ring, "vertices"}), lv.push_back({&h.geom, "linestrings"}), ext = "geoarrow.multilinestring";
        break;
    case GPL_MULTIPOLYGON:
        lv.push_back({&h.ring, "vertices"}), lv.push_back({&h.part, "rings"}), lv.push_back({&h.geom, "polygons"});
        ext = "geoarrow.multipolygon";
        break;
    default:
        set_error("cannot export geometry type %d", h.type);
        return GPL_ERR_INVALID_TYPE;
    }
    const bool top_is_coord = lv.empty();
    const int64_t nulls = count_nulls(h.valid, h.n);
    uint8_t *vbm = h.valid.empty() ? nullptr : dup(h.valid.data(), h.valid.size());
    {
        std::vector<void *> owned;
        if (top_is_coord && vbm) owned.push_back(vbm);
        make_array(cur, cur_len, top_is_coord ? nulls : 0, {top_is_coord ? (const void *)vbm : nullptr}, owned, {vals});
        make_schema(cur_s, "+w:2", lv.empty() ? "geometry" : lv[0].child_name, top_is_coord ? ext_metadata(ext) : "", {vals_s},
                    top_is_coord && nulls > 0);
    }
    for (size_t d = 0; d < lv.size(); ++d) {
        const bool top = d + 1 == lv.size();
        const std::vector<int64_t> &off = *lv[d].off;
        const int64_t len = (int64_t)off.size() - 1;
        const bool small = off.back() < (1LL << 31);
        void *obuf;
        if (small) {
            int32_t *o32 = static_cast<int32_t *>(malloc(sizeof(int32_t) * off.size()));
            for (size_t i = 0; i < off.size(); ++i) o32[i] = (int32_t)off[i];
            obuf = o32;
        } else {
            obuf = dup(off.data(), off.size());
        }
        ArrowArray *na = static_cast<ArrowArray *>(calloc(1, sizeof(ArrowArray)));
        ArrowSchema *ns = static_cast<ArrowSchema *>(calloc(1, sizeof(ArrowSchema)));
        std::vector<void *> owned = {obuf};
        if (top && vbm) owned.push_back(vbm);
        make_array(na, len, top ? nulls : 0, {top ? (const void *)vbm : nullptr, obuf}, owned, {cur});
        make_schema(ns, small ? "+l" : "+L", top ? "geometry" : lv[d + 1].child_name, top ? ext_metadata(ext) : "", {cur_s},
                    top && nulls > 0);
        cur = na, cur_s = ns;
    }
    // move the top-level structs into the caller's memory
    memcpy(out_array, cur, sizeof(ArrowArray));
    memcpy(out_schema, cur_s, sizeof(ArrowSchema));
    free(cur);
    free(cur_s);
    return GPL_OK;
}

extern "C" int gpl_export_f64_arrow(const double *values, const uint8_t *validity, int64_t n, void *out_array, void *out_schema) {
    GPL_REQUIRE((values || n == 0) && out_array && out_schema, GPL_ERR_INVALID_ARG, "gpl_export_f64_arrow: NULL argument");
    double *v = dup(values, (size_t)n);
    uint8_t *bm = validity ? dup(validity, (size_t)(n + 7) / 8) : nullptr;
    int64_t nulls = 0;
    if (bm)
        for (int64_t i = 0; i < n; ++i) nulls += !((bm[i >> 3] >> (i & 7)) & 1);
    std::vector<void *> owned = {v};
    if (bm) owned.push_back(bm);
    make_array(static_cast<ArrowArray *>(out_array), n, nulls, {bm, v}, owned, {});
    make_schema(static_cast<ArrowSchema *>(out_schema), "g", "", "", {}, nulls > 0);
    return GPL_OK;
}
extern "C" int gpl_export_bool_arrow(const uint8_t *bitmap, const uint8_t *validity, int64_t n, void *out_array, void *out_schema) {
    GPL_REQUIRE((bitmap || n == 0) && out_array && out_schema, GPL_ERR_INVALID_ARG, "gpl_export_bool_arrow: NULL argument");
    uint8_t *v = dup(bitmap, (size_t)(n + 7) / 8);
    uint8_t *bm = validity ? dup(validity, (size_t)(n + 7) / 8) : nullptr;
    int64_t nulls = 0;
    if (bm)
        for (int64_t i = 0; i < n; ++i) nulls += !((bm[i >> 3] >> (i & 7)) & 1);
    std::vector<void *> owned = {v};
    if (bm) owned.push_back(bm);
    make_array(static_cast<ArrowArray *>(out_array), n, nulls, {bm, v}, owned, {});
    make_schema(static_cast<ArrowSchema *>(out_schema), "b", "", "", {}, nulls > 0);
    return GPL_OK;
}
