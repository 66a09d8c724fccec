// k_simplify.cu — GeoSeries::simplify (geoseries.rs:108-116, impl :240-242): Ramer-Douglas-Peucker as in
// geo 0.27 simplify.rs (recalled; restated in oracle/geo_oracle.c `og_simplify_mask`).
//
// geo's compute_rdp is recursive and carries ONE running counter through the depth-first traversal
// (`simplified_len`): a slice whose farthest point is within epsilon is culled to its end points only while the
// whole line keeps at least INITIAL_MIN points (2 for linestrings, 4 for polygon rings), otherwise that slice is
// returned untouched.  The outcome therefore depends on the traversal order, which is reproduced: one warp per
// line / ring runs the recursion as an explicit stack machine (right half pushed first, left half on top), the
// lanes split each slice's farthest-point search (point-to-SEGMENT distance, the fold's `>=` keeps the LAST
// maximum, NaN distances never win).  Output = keep flag per coordinate -> scan -> gather; the outer offset levels
// and the validity bitmap are shared with the input.
#include "common.cuh"
#include "scan.cuh"

namespace gpl {

constexpr int kRdpWarps = 8;

__global__ void __launch_bounds__(kRdpWarps * 32) k_rdp(int64_t n_chains, const double2 *__restrict__ xy, const int64_t *__restrict__ off,
                                                       double eps, int min_len, int2 *__restrict__ stack_ws,
                                                       uint8_t *__restrict__ keep) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t k = warp; k < n_chains; k += nwarps) {
        const int64_t c0 = off[k], n = off[k + 1] - c0;
        if (n <= 2 || !(eps > 0.0)) {  // nothing to cull (epsilon <= 0 returns the input)
            for (int64_t i = lane; i < n; i += 32) keep[c0 + i] = 1;
            continue;
        }
        for (int64_t i = lane; i < n; i += 32) keep[c0 + i] = 0;
        __syncwarp();
        const double2 *P = xy + c0;
        uint8_t *K = keep + c0;
        int2 *stack = stack_ws + c0;  // at most n pending slices
        int64_t simplified_len = n;
        int sp = 0;
        if (lane == 0) stack[0] = make_int2(0, (int)(n - 1));
        sp = 1;
        __syncwarp();
        while (sp > 0) {
            const int2 f = stack[--sp];
            const int lo = f.x, hi = f.y, len = hi - lo + 1;
            __syncwarp();
            if (lane == 0) K[lo] = 1, K[hi] = 1;
            if (len == 2) continue;
            const double2 a = P[lo], b = P[hi];
            double best = 0.0;
            int best_i = 0;
            for (int i = lo + 1 + lane; i < hi; i += 32) {
                const double d = line_segment_distance(P[i], a, b);
                if (d >= best) best = d, best_i = i;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double od = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
                if (od > best || (od == best && oi > best_i)) best = od, best_i = oi;
            }
            if (best > eps) {
                if (lane == 0) {
                    stack[sp] = make_int2(best_i, hi);
                    stack[sp + 1] = make_int2(lo, best_i);
                }
                sp += 2;
                __syncwarp();
            } else {
                const int64_t new_len = simplified_len - (len - 2);
                if (new_len < min_len) {
                    for (int i = lo + 1 + lane; i < hi; i += 32) K[i] = 1;  // slice returned untouched
                } else {
                    simplified_len = new_len;
                }
            }
        }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(256) k_rdp_gather(int64_t nc, const double2 *__restrict__ xy, const uint8_t *__restrict__ keep,
                                                    const int64_t *__restrict__ pos, double2 *__restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nc; i += stride)
        if (keep[i]) out[pos[i]] = xy[i];
}
__global__ void __launch_bounds__(256) k_rdp_offsets(int64_t n_chains, const int64_t *__restrict__ off, const int64_t *__restrict__ pos,
                                                     int64_t *__restrict__ new_off) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i <= n_chains) new_off[i] = pos[off[i]];  // pos has nc + 1 entries: pos[nc] = total
}

}  // namespace gpl

using namespace gpl;

extern "C" int gpl_simplify(gpl_ctx *ctx, const gpl_array *in, double tolerance, gpl_array **out) {
    GPL_REQUIRE(ctx && in && out, GPL_ERR_INVALID_ARG, "gpl_simplify: NULL argument");
    const bool line = in->type == GPL_LINESTRING, rings = in->type == GPL_MULTILINESTRING || in->type == GPL_POLYGON || in->type == GPL_MULTIPOLYGON;
    GPL_REQUIRE(line || rings, GPL_ERR_INVALID_TYPE, "Expected LineString, MultiLineString, Polygon or MultiPolygon (found geometry type %d)",
                in->type);
    GPL_REQUIRE(in->n_coords < (1LL << 31), GPL_ERR_UNSUPPORTED, "simplify: more than 2^31 coordinates in one array");
    GPL_CUDA(cudaSetDevice(ctx->device));
    const int64_t nc = in->n_coords;
    const int64_t n_chains = line ? in->n_geoms : in->n_rings;
    const int64_t *off = line ? in->geom_off : in->ring_off;
    const int min_len = (in->type == GPL_POLYGON || in->type == GPL_MULTIPOLYGON) ? 4 : 2;
    Scratch<uint8_t> keep;
    Scratch<int2> stack;
    Scratch<int64_t> pos, new_off, total;
    GPL_TRY(keep.get(ctx, (size_t)nc));
    GPL_TRY(stack.get(ctx, (size_t)nc));
    GPL_TRY(pos.get(ctx, (size_t)nc + 1));
    GPL_TRY(new_off.get(ctx, (size_t)n_chains + 1));
    GPL_TRY(total.get(ctx, 1));
    const double2 *xy = reinterpret_cast<const double2 *>(in->xy);
    if (n_chains > 0) {
        const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(n_chains, kRdpWarps), (int64_t)kSMs * 8));
        GPL_LAUNCH(ctx, k_rdp, grid, kRdpWarps * 32, 0, n_chains, xy, off, tolerance, min_len, stack.p, keep.p);
    }
    stack.reset();
    GPL_TRY((exclusive_scan<uint8_t, int64_t>(ctx, keep.p, nc, pos.p, total.p)));
    int64_t h_total = 0;
    GPL_CUDA(cudaMemcpyAsync(&h_total, total.p, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
    GPL_CUDA(cudaStreamSynchronize(ctx->stream));
    Scratch<double> oxy;
    GPL_TRY(oxy.get(ctx, (size_t)h_total * 2));
    if (nc > 0) {
        const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(nc, 256), (int64_t)kSMs * 16));
        GPL_LAUNCH(ctx, k_rdp_gather, grid, 256, 0, nc, xy, keep.p, pos.p, reinterpret_cast<double2 *>(oxy.p));
    }
    if (off) GPL_LAUNCH(ctx, k_rdp_offsets, (int)ceil_div(n_chains + 1, 256), 256, 0, n_chains, off, pos.p, new_off.p);
    else GPL_CUDA(cudaMemsetAsync(new_off.p, 0, sizeof(int64_t), ctx->stream));  // an empty array may carry no offsets buffer
    gpl_array *o = array_new(ctx, in->type);
    o->n_geoms = in->n_geoms, o->n_parts = in->n_parts, o->n_rings = in->n_rings, o->n_coords = h_total;
    o->xy = oxy.take(), o->own_xy = true;
    if (line) {
        o->geom_off = new_off.take(), o->own_geom = true;
    } else {
        o->ring_off = new_off.take(), o->own_ring = true;
        o->geom_off = in->geom_off, o->part_off = in->part_off;  // outer levels shared with the input
    }
    o->validity = in->validity;
    o->parent = const_cast<gpl_array *>(in);
    array_retain(o->parent);
    *out = o;
    return GPL_OK;
}
