// synth.cu — device-side generators for BASELINE.json's synthetic workloads, bit-identical to
// geopolars_b200/synth.py and oracle og_gen_uniform_points (counter-based splitmix64).
// Benchmark/test utilities, not part of the reference's surface.
#include <math.h>

#include "common.cuh"

namespace gpl {

__device__ __forceinline__ double splitmix_u(uint64_t seed, uint64_t counter) {
    uint64_t z = seed + 0x9E3779B97F4A7C15ULL * (counter + 1ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z ^= z >> 31;
    return (double)(z >> 11) * 0x1.0p-53;
}

__global__ void k_gen_points(uint64_t seed, int64_t first, int64_t n, double scale, double2 *__restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) {
        uint64_t g = (uint64_t)(first + i);
        out[i] = make_double2(scale * splitmix_u(seed, g * 4ULL), scale * splitmix_u(seed, g * 4ULL + 1ULL));
    }
}

// random walk of k coords; one thread per linestring (sequential prefix sum = synth.py's rounding order)
__global__ void k_gen_walks(uint64_t seed, uint64_t other_seed, int has_other, int64_t first, int64_t n, int32_t k,
                            double2 *__restrict__ out, int64_t *__restrict__ off) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i > n) return;
    off[i] = i * (int64_t)k;
    if (i == n) return;
    uint64_t base = (uint64_t)(first + i) * (uint64_t)(2 * k);
    double x, y;
    if (!has_other) {
        x = 1000.0 * splitmix_u(seed, base);
        y = 1000.0 * splitmix_u(seed, base + 1ULL);
    } else {
        double ax = 1000.0 * splitmix_u(other_seed, base);
        double ay = 1000.0 * splitmix_u(other_seed, base + 1ULL);
        x = ax + (4.0 * splitmix_u(seed, base) - 2.0);
        y = ay + (4.0 * splitmix_u(seed, base + 1ULL) - 2.0);
    }
    double2 *o = out + i * (int64_t)k;
    o[0] = make_double2(x, y);
    for (int32_t s = 1; s < k; ++s) {
        double dx = 2.0 * splitmix_u(seed, base + (uint64_t)(2 * s)) - 1.0;
        double dy = 2.0 * splitmix_u(seed, base + (uint64_t)(2 * s + 1)) - 1.0;
        x = x + dx;
        y = y + dy;
        o[s] = make_double2(x, y);
    }
}

// star-shaped blobs (config 5).  NOT bit-identical to numpy (device sincos): used at sizes where the
// tests rely on size-independent properties, never for oracle comparison.
__global__ void k_gen_blobs(uint64_t seed, int64_t first, int64_t n, int32_t nvert, double2 *__restrict__ out,
                            int64_t *__restrict__ ring_off, int64_t *__restrict__ geom_off) {
    const int lane = threadIdx.x & 31;
    int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp; i <= n; i += nwarps) {
        if (lane == 0) {
            ring_off[i] = i * (int64_t)(nvert + 1);
            geom_off[i] = i;
        }
        if (i == n) break;
        uint64_t base = (uint64_t)(first + i) * (uint64_t)(nvert + 2);
        double cx = 1000.0 * splitmix_u(seed, base), cy = 1000.0 * splitmix_u(seed, base + 1ULL);
        double2 *o = out + i * (int64_t)(nvert + 1);
        for (int32_t v = lane; v <= nvert; v += 32) {
            int32_t k = v == nvert ? 0 : v;
            double r = 0.5 + 0.5 * splitmix_u(seed, base + 2ULL + (uint64_t)k);
            double th = 2.0 * 3.14159265358979323846 * (double)k / (double)nvert;
            double sn, cs;
            sincos(th, &sn, &cs);
            o[v] = make_double2(cx + r * cs, cy + r * sn);
        }
    }
}

}  // namespace gpl

using namespace gpl;

extern "C" int gpl_gen_uniform_points(gpl_ctx *ctx, uint64_t stream_id, int64_t first, int64_t n, double scale, double *out) {
    GPL_REQUIRE(ctx && (out || n == 0), GPL_ERR_INVALID_ARG, "gpl_gen_uniform_points: NULL argument");
    GPL_CUDA(cudaSetDevice(ctx->device));
    if (n == 0) return GPL_OK;
    int grid = (int)std::min<int64_t>(ceil_div(n, 256), (int64_t)kSMs * 16);
    GPL_LAUNCH(ctx, k_gen_points, grid, 256, 0, 0xB2000000ULL + stream_id, first, n, scale, reinterpret_cast<double2 *>(out));
    return GPL_OK;
}
extern "C" int gpl_gen_walk_linestrings(gpl_ctx *ctx, uint64_t stream_id, int64_t other_of, int64_t first, int64_t n, int32_t k,
                                        double *out_xy, int64_t *out_off) {
    GPL_REQUIRE(ctx && out_xy && out_off && k >= 1, GPL_ERR_INVALID_ARG, "gpl_gen_walk_linestrings: bad argument");
    GPL_CUDA(cudaSetDevice(ctx->device));
    GPL_LAUNCH(ctx, k_gen_walks, (int)ceil_div(n + 1, 128), 128, 0, 0xB2000000ULL + stream_id,
               0xB2000000ULL + (uint64_t)(other_of >= 0 ? other_of : 0), other_of >= 0 ? 1 : 0, first, n, k,
               reinterpret_cast<double2 *>(out_xy), out_off);
    return GPL_OK;
}
extern "C" int gpl_gen_blob_polygons(gpl_ctx *ctx, uint64_t stream_id, int64_t first, int64_t n, int32_t nvert, double *out_xy,
                                     int64_t *out_ring_off, int64_t *out_geom_off) {
    GPL_REQUIRE(ctx && out_xy && out_ring_off && out_geom_off && nvert >= 3, GPL_ERR_INVALID_ARG, "gpl_gen_blob_polygons: bad argument");
    GPL_CUDA(cudaSetDevice(ctx->device));
    int grid = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(n + 1, 8), (int64_t)kSMs * 16));
    GPL_LAUNCH(ctx, k_gen_blobs, grid, 256, 0, 0xB2000000ULL + stream_id, first, n, nvert, reinterpret_cast<double2 *>(out_xy),
               out_ring_off, out_geom_off);
    return GPL_OK;
}
