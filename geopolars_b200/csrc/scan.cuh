// scan.cuh — exclusive prefix sum (count -> offsets) used by every variable-length output on the path:
// grid cell lists and edge buckets of the join index, join pair lists, convex-hull ring offsets.
// Three launches (tile sums, scan of tile sums, tile re-scan) — the inputs are read twice from L2/HBM,
// which is irrelevant next to the kernels that produce them.
#pragma once

#include "common.cuh"

namespace gpl {

constexpr int kScanThreads = 512;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ int64_t block_exclusive_scan(int64_t v, int64_t &total) {
    // exclusive scan of one value per thread across a kScanThreads block; total = block sum
    constexpr int NW = kScanThreads / 32;
    __shared__ int64_t warp_off[NW];
    __shared__ int64_t s_total;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int64_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int64_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_off[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        int64_t w = lane < NW ? warp_off[lane] : 0;
        int64_t winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int64_t t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        if (lane < NW) warp_off[lane] = winc - w;
        if (lane == NW - 1) s_total = winc;
    }
    __syncthreads();
    int64_t excl = warp_off[wid] + inc - v;
    total = s_total;
    __syncthreads();  // shared scratch is reused by the next call
    return excl;
}

template <typename In>
__global__ void __launch_bounds__(kScanThreads) k_scan_tile_sums(const In *__restrict__ in, int64_t n,
                                                                int64_t *__restrict__ tile_sums) {
    int64_t base = (int64_t)blockIdx.x * kScanTile;
    int64_t s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        int64_t i = base + (int64_t)k * kScanThreads + threadIdx.x;
        if (i < n) s += (int64_t)in[i];
    }
    int64_t total;
    (void)block_exclusive_scan(s, total);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// single block: in-place exclusive scan of tile sums; also writes the grand total to *total_out
static __global__ void __launch_bounds__(kScanThreads) k_scan_sums(int64_t *__restrict__ sums, int64_t n_tiles,
                                                           int64_t *__restrict__ total_out) {
    int64_t carry = 0;
    for (int64_t base = 0; base < n_tiles; base += kScanThreads) {
        int64_t i = base + threadIdx.x;
        int64_t v = i < n_tiles ? sums[i] : 0;
        int64_t total;
        int64_t e = block_exclusive_scan(v, total);
        if (i < n_tiles) sums[i] = carry + e;
        carry += total;
    }
    if (threadIdx.x == 0 && total_out) *total_out = carry;
}

template <typename In, typename Out>
__global__ void __launch_bounds__(kScanThreads) k_scan_apply(const In *__restrict__ in, int64_t n,
                                                            const int64_t *__restrict__ tile_offsets,
                                                            Out *__restrict__ out /* n+1 */, const int64_t *total) {
    // thread t owns kScanItems consecutive elements so the per-thread partials scan in order
    int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
    int64_t vals[kScanItems];
    int64_t s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        int64_t i = base + k;
        vals[k] = i < n ? (int64_t)in[i] : 0;
        s += vals[k];
    }
    int64_t tot;
    int64_t e = block_exclusive_scan(s, tot) + tile_offsets[blockIdx.x];
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        int64_t i = base + k;
        if (i < n) out[i] = (Out)e;
        e += vals[k];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = (Out)(*total);
}

// out[0..n] = exclusive prefix sums of in[0..n-1]; out[n] = total.  total_dev (optional) receives the
// grand total as int64 on the device.  `in` and `out` may alias only if sizeof(In)==sizeof(Out).
template <typename In, typename Out>
int exclusive_scan(gpl_ctx *ctx, const In *in, int64_t n, Out *out, int64_t *total_dev) {
    int64_t tiles = ceil_div(n > 0 ? n : 1, kScanTile);
    Scratch<int64_t> sums;
    GPL_TRY(sums.get(ctx, (size_t)tiles + 1));
    int64_t *tot = total_dev ? total_dev : sums.p + tiles;
    GPL_LAUNCH(ctx, (k_scan_tile_sums<In>), (int)tiles, kScanThreads, 0, in, n, sums.p);
    GPL_LAUNCH(ctx, k_scan_sums, 1, kScanThreads, 0, sums.p, tiles, tot);
    GPL_LAUNCH(ctx, (k_scan_apply<In, Out>), (int)tiles, kScanThreads, 0, in, n, sums.p, out, tot);
    return GPL_OK;
}

}  // namespace gpl
