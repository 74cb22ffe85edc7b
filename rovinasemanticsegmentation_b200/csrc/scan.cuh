// Device-wide exclusive prefix sum over uint32 (stream compaction of samples, lattice vertex numbering,
// CSR row offsets).  Three phases: per-tile sums -> scan of tile sums (recursive) -> per-tile rescan.
// Deterministic (integer), any n.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace rss {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 16;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;  // 4096 elements per CTA

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// exclusive scan of one value per thread across the CTA; returns the exclusive prefix, *total = CTA sum
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* total) {
    __shared__ uint32_t wsum[SCAN_THREADS / 32];
    __shared__ uint32_t btotal;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t inc = warp_incl_scan(v, lane);
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    if (w == 0) {
        uint32_t s = lane < SCAN_THREADS / 32 ? wsum[lane] : 0;
        uint32_t si = warp_incl_scan(s, lane);
        if (lane < SCAN_THREADS / 32) wsum[lane] = si - s;
        if (lane == SCAN_THREADS / 32 - 1) btotal = si;
    }
    __syncthreads();
    uint32_t r = inc - v + wsum[w];
    *total = btotal;
    __syncthreads();
    return r;
}

// phase 1: tile sums
static __global__ void __launch_bounds__(SCAN_THREADS) scan_tile_sums(const uint32_t* __restrict__ in, size_t n,
                                                               uint32_t* __restrict__ sums) {
    const size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++)
        if (base + k < n) s += in[base + k];
    uint32_t tot;
    block_excl_scan(s, &tot);
    if (threadIdx.x == 0) sums[blockIdx.x] = tot;
}

// phase 3 (and the whole job when one tile suffices): rescan a tile with its offset
// `in` may alias `out` (in-place scans of rank / deg / nseg): no __restrict__ on the two, so the loads stay coherent
static __global__ void __launch_bounds__(SCAN_THREADS) scan_tile_apply(const uint32_t* in, size_t n,
                                                                const uint32_t* __restrict__ tile_off,
                                                                uint32_t* out,
                                                                uint32_t* __restrict__ total_out) {
    const size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        v[k] = base + k < n ? in[base + k] : 0;
        s += v[k];
    }
    uint32_t tot;
    uint32_t pre = block_excl_scan(s, &tot);
    const uint32_t off = tile_off ? tile_off[blockIdx.x] : 0;
    pre += off;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        if (base + k < n) out[base + k] = pre;
        pre += v[k];
    }
    if (total_out && blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) *total_out = off + tot;
}

inline size_t scan_tmp_elems(size_t n) {
    size_t t = 0;
    while (n > (size_t)SCAN_TILE) {
        n = (n + SCAN_TILE - 1) / SCAN_TILE;
        t += n + 1;
    }
    return t + 2;
}

// in may alias out.  tmp: scan_tmp_elems(n) uint32.  total_out (device, optional) receives the grand total.
inline void exclusive_scan_u32(const uint32_t* in, uint32_t* out, size_t n, uint32_t* tmp, uint32_t* total_out,
                               cudaStream_t st, uint64_t* launches) {
    if (n == 0) {
        if (total_out) cudaMemsetAsync(total_out, 0, sizeof(uint32_t), st);
        return;
    }
    const size_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    if (tiles == 1) {
        scan_tile_apply<<<1, SCAN_THREADS, 0, st>>>(in, n, nullptr, out, total_out);
        (*launches)++;
        return;
    }
    uint32_t* sums = tmp;
    scan_tile_sums<<<(unsigned)tiles, SCAN_THREADS, 0, st>>>(in, n, sums);
    (*launches)++;
    exclusive_scan_u32(sums, sums, tiles, tmp + tiles + 1, nullptr, st, launches);
    scan_tile_apply<<<(unsigned)tiles, SCAN_THREADS, 0, st>>>(in, n, sums, out, total_out);
    (*launches)++;
}

}  // namespace rss
