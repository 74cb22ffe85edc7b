// Fused mean-field iteration for point sets whose order is spatially coherent (image raster order: the keyframe path,
// DenseCRF2D; any point set whose consecutive points share lattice vertices).  Per iteration there are two launches:
//   meanfield_point_kernel  slice of the previous iteration's blurred value tables + Potts + unary + per-layer soft-max
//                           (+ gated argmax / Q store on the last pass) AND the splat of the new marginals;
//   blur_multi_coop_kernel  every axis of every lattice of the CRF, one cooperative launch with grid barriers.
// Reference: third-party/densecrf/src/densecrf.cpp:98-131 (expAndNormalize, inference), pairwise.cpp:63-80
// (DenseKernel::filter), labelcompatibility.cpp:46-48 (Potts), permutohedral.cpp:529-589 (sseCompute).
//
// A tile (one CTA of the point kernel) is a block of 256 points - a 32 x 8 pixel block for image CRFs, else a run of
// consecutive points - and ONE THREAD OWNS ONE POINT with all its label channels in registers.
//   staging  asynchronous copies bring in, without touching a register: the tile's unary rows (TMA bulk copies,
//            cp.async.bulk + mbarrier, into the shared Q tile: a thread reads its unary row, later overwrites it with its
//            marginals), the tile's splat lists (segment metadata + pairs, TMA), and the DISTINCT value rows of each
//            lattice that the tile's points reference (cp.async 16 B each, a few KB to ~13 KB per lattice, completion
//            counted on the same mbarrier with cp.async.mbarrier.arrive).
//   phase 1  t = -unary + sum over lattices and simplex corners of w * row[slot]: the weight already contains the
//            barycentric coordinate, the post-normalisation, the Potts weight and the slice scale, and the row comes out
//            of SHARED MEMORY by its tile-local slot - d+1 LDS.128 x G per lattice instead of L1/L2 gathers.  Soft-max
//            per label layer in registers (no shuffles), marginals to the shared Q tile.
//   phase 2  thread = one segment of <= TILE_SEG (point, weight) pairs of one vertex: sums w * Q[point] for all channels
//            out of shared memory and issues G red.global.add.v4.f32.  Global atomics per iteration = (distinct vertices
//            per tile) * G instead of (nonzeros) * G, Q is never re-read from L2, and nothing depends on how noisy the
//            pixel order is (+-8 colour noise sends 64 % of consecutive pixels to a different bilateral vertex).
//            The 32 segments a warp walks are stored column-major (conflict-free pair reads), and a point's row in the Q
//            tile is rotated by 3 x its warp index (meanfield.cuh: q_row), so that narrow vertex blobs still spread over
//            all eight 16-byte bank groups.
// Value tables rotate through three buffers per lattice: `res` (blurred result being sliced), `tgt` (all zero, receives
// the splat) and `spare`; the blur ping-pongs between tgt and spare and clears the old `res`, the next tgt.
// The point kernel itself lives in meanfield_point.inl and is compiled twice (this file: up to 85 registers, the build
// for a context that is alone on the GPU; meanfield_shared.cu: 64 registers, for SMs shared with other keyframes' blur).
//
// Compile-time tuning knobs (build.py: RSS_NVCC_DEFS="-DNAME=value"), defaults measured on B200:
#include <algorithm>

#include "kernels.hpp"
#include "lattice.cuh"
#include "meanfield.cuh"

#ifndef RSS_POINT_MAXNREG
#define RSS_POINT_MAXNREG 0  // > 0: register cap of the "alone" build of the point kernel instead of RSS_TILE_MINB
#endif
#ifndef RSS_BLUR_SHARED_T
#define RSS_BLUR_SHARED_T 256  // CTA size of the cooperative blur while several keyframes share the GPU (128: 1027, 512: 1075
                               // keyframes/s against 1100 at 256)
#endif
#ifndef RSS_BLUR_ALONE_T
#define RSS_BLUR_ALONE_T 768   // ... and alone on the GPU: 26.1 us at 512, 23.4 us at 768, 25.2 us at 1024 threads per SM
#endif
#ifndef RSS_BLUR_SHARED_GRID_DIV
#define RSS_BLUR_SHARED_GRID_DIV 1  // the shared-GPU blur runs on sm_count / this many SMs
#endif
#ifndef RSS_BLUR_U
#define RSS_BLUR_U 2      // independent (vertex, channel group) items a blur thread keeps in flight
#endif
#ifndef RSS_BLUR_MAXT
#define RSS_BLUR_MAXT 1024  // launch bound of the cooperative blur = its register cap (64): a 256-thread CTA must fit next to
                            // three 64-register point CTAs (a bound of 768 lets the compiler take more, which cost 7 % of the
                            // shared-GPU throughput); the phases are L2-throughput-bound once >= 512 threads per SM keep loads
                            // in flight (tools/micro/blur_bench.cu: 35 us at 128, 25 us at 512 and 1024)
#endif
#ifndef RSS_BLUR_FUSE
#define RSS_BLUR_FUSE 2  // lattice axes blurred per phase of the cooperative blur (1 = one grid barrier per axis); measured on
                         // the keyframe workload: 30.4 us (1), 28.0 us (2), 29.9 us (3) per launch
#endif
#ifndef RSS_BLUR_FU
#define RSS_BLUR_FU 1  // independent items per thread and trip in a 2-axis phase
#endif
#ifndef RSS_BLUR_FUSE3_ITEMS
#define RSS_BLUR_FUSE3_ITEMS (400u << 10)  // float4 items (vcap * G) up to which 3 axes are fused (27 row reads per item) ...
#endif
#ifndef RSS_BLUR_FUSE2_ITEMS
#define RSS_BLUR_FUSE2_ITEMS (1u << 20)    // ... and 2 axes (9 row reads per item)
#endif

#if RSS_POINT_MAXNREG > 0
#define RSS_POINT_BOUNDS __maxnreg__(RSS_POINT_MAXNREG)
#else
#define RSS_POINT_BOUNDS __launch_bounds__(TILE_POINTS, RSS_TILE_MINB)
#endif
#define RSS_POINT_NS point_alone
#define RSS_POINT_ENTRY launch_meanfield_fused_alone
#include "meanfield_point.inl"

namespace rss {
using point_alone::launch_meanfield_fused_alone;

template <int D1>
__device__ __forceinline__ void load_row_i(const int* __restrict__ p, int (&o)[D1]) {
    if constexpr (D1 % 4 == 0) {
#pragma unroll
        for (int k = 0; k < D1 / 4; k++) {
            const int4 v = __ldg(reinterpret_cast<const int4*>(p) + k);
            o[4 * k] = v.x; o[4 * k + 1] = v.y; o[4 * k + 2] = v.z; o[4 * k + 3] = v.w;
        }
    } else if constexpr (D1 % 2 == 0) {
#pragma unroll
        for (int k = 0; k < D1 / 2; k++) {
            const int2 v = __ldg(reinterpret_cast<const int2*>(p) + k);
            o[2 * k] = v.x; o[2 * k + 1] = v.y;
        }
    } else {
#pragma unroll
        for (int k = 0; k < D1; k++) o[k] = __ldg(p + k);
    }
}
template <int D1>
__device__ __forceinline__ void load_row_f(const float* __restrict__ p, float (&o)[D1]) {
    if constexpr (D1 % 4 == 0) {
#pragma unroll
        for (int k = 0; k < D1 / 4; k++) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(p) + k);
            o[4 * k] = v.x; o[4 * k + 1] = v.y; o[4 * k + 2] = v.z; o[4 * k + 3] = v.w;
        }
    } else if constexpr (D1 % 2 == 0) {
#pragma unroll
        for (int k = 0; k < D1 / 2; k++) {
            const float2 v = __ldg(reinterpret_cast<const float2*>(p) + k);
            o[2 * k] = v.x; o[2 * k + 1] = v.y;
        }
    } else {
#pragma unroll
        for (int k = 0; k < D1; k++) o[k] = __ldg(p + k);
    }
}
// ---------------------------------------------------------------------------------------------------------------
// Tile data of one lattice (see FusedLat in meanfield.cuh).  Built once per lattice by one CTA per tile with a
// shared-memory hash table (native 32-bit CAS / integer adds only): insert keys -> count -> compact + scan -> fill.
// ---------------------------------------------------------------------------------------------------------------

__device__ __forceinline__ int3 block_excl_scan3(int a, int b, int c, int3* total) {  // 256 threads
    __shared__ int3 wsum[8];
    __shared__ int3 btot;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int ia = a, ib = b, ic = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int ta = __shfl_up_sync(0xffffffffu, ia, o), tb = __shfl_up_sync(0xffffffffu, ib, o),
                  tc = __shfl_up_sync(0xffffffffu, ic, o);
        if (lane >= o) { ia += ta; ib += tb; ic += tc; }
    }
    if (lane == 31) wsum[w] = make_int3(ia, ib, ic);
    __syncthreads();
    if (threadIdx.x == 0) {
        int sa = 0, sb = 0, sc = 0;
        for (int k = 0; k < 8; k++) {
            const int3 v = wsum[k];
            wsum[k] = make_int3(sa, sb, sc);
            sa += v.x; sb += v.y; sc += v.z;
        }
        btot = make_int3(sa, sb, sc);
    }
    __syncthreads();
    const int3 r = make_int3(ia - a + wsum[w].x, ib - b + wsum[w].y, ic - c + wsum[w].z);
    *total = btot;
    __syncthreads();
    return r;
}

template <int D1>
__global__ void __launch_bounds__(256) tile_csr_build_kernel(const int* __restrict__ offsets, const float* __restrict__ bary,
                                                             const float* __restrict__ norm, int pre, int post,
                                                             float slice_scale, const TileMap tm, int row_bytes, int HC,
                                                             const uint32_t* __restrict__ counts, const TileCsrOut out) {
    __shared__ int hist[TILE_SEG + 1], binstart[TILE_SEG + 1], cumstart[TILE_SEG + 1];
    __shared__ int colstart[8 * 32];  // per warp: column starts of the segment group it is ordering
    extern __shared__ int tile_smem[];  // above the 48 KB static limit for large d
    int* hkeys = tile_smem;  // HC slots (power of two, > pairs per tile)
    // per slot: pair count in the low half, fill cursor of the slot's pair list in the high half (both <= TP * D1 <
    // 65536) - one word instead of two, and 16-bit row slots: 10 instead of 16 bytes per hash slot, which is what lets
    // 5 (d = 5) / 6 (d = 3) CTAs share an SM instead of 4 - 1200 tiles are then two waves of CTAs instead of three
    unsigned* hcc = reinterpret_cast<unsigned*>(tile_smem + HC);
    unsigned short* hrow = reinterpret_cast<unsigned short*>(tile_smem + 2 * HC);
    unsigned short* pslot = hrow + HC;
    uint2* spairs = reinterpret_cast<uint2*>(pslot + TILE_POINTS * D1);  // the tile's pair lists, ordered here, written out at the end
    int* segs = reinterpret_cast<int*>(spairs + TILE_POINTS * D1);      // start | length << 16 of every segment
    const int hshift = 32 - (31 - __clz(HC));
    const int tile = blockIdx.x;
    constexpr int TP = TILE_POINTS;
    const TileOrigin org = tile_origin(tm, tile);
    const int npairs = TP * D1;  // slots; points outside the image / beyond N are skipped
    const size_t tb = (size_t)tile * TP * D1;
    if (counts[1]) {
        if (threadIdx.x == 0) out.tile_info[tile] = make_int2(0, 0);
        return;
    }
    for (int i = threadIdx.x; i < HC; i += 256) { hkeys[i] = -1; hcc[i] = 0u; }
    if (threadIdx.x <= TILE_SEG) hist[threadIdx.x] = 0;
    __syncthreads();
    // pair i = (corner j, local point lp), lp fastest: the point-major outputs are written coalesced
    for (int i = threadIdx.x; i < npairs; i += 256) {
        const int j = i / TP, lp = i - j * TP, p = tile_point(tm, org, lp);
        if (p < 0) continue;
        const int key = offsets[(size_t)p * D1 + j];
        unsigned h = ((unsigned)key * 2654435761u) >> hshift;
        for (;;) {
            const int cur = hkeys[h];
            if (cur == key) break;
            if (cur == -1) {
                const int old = atomicCAS(&hkeys[h], -1, key);
                if (old == -1 || old == key) break;
            }
            h = (h + 1) & (HC - 1);
        }
        pslot[i] = (unsigned short)h;
        atomicAdd(&hcc[h], 1u);
    }
    __syncthreads();
    // Lists are cut into segments of at most TILE_SEG pairs and the segments are ordered by length (longest first), so
    // that the lanes of a warp of the gather walk lists of (nearly) equal length.  Thread t owns HC / 256 hash slots.
    int nseg = 0, ncnt = 0, nvert = 0;
    const int spt = HC / 256, s0 = threadIdx.x * spt;  // slots per thread
#pragma unroll 4
    for (int k = 0; k < spt; k++) {
        const int c = (int)(hcc[s0 + k] & 0xffffu);
        if (c > 0) {
            const int full = c / TILE_SEG, rem = c - full * TILE_SEG;
            if (full) atomicAdd(&hist[TILE_SEG], full);
            if (rem) atomicAdd(&hist[rem], 1);
            nseg += full + (rem ? 1 : 0);
            ncnt += c;
            nvert++;
        }
    }
    int3 tot;
    int3 pre3 = block_excl_scan3(nseg, ncnt, nvert, &tot);  // ends with a barrier: hist is complete
    if (threadIdx.x == 0) {
        int pos = 0, cpos = 0;  // segments / pairs before the first segment of this length, in the sorted segment order
        for (int len = TILE_SEG; len >= 1; len--) {
            binstart[len] = pos; cumstart[len] = cpos;
            pos += hist[len]; cpos += hist[len] * len;
            hist[len] = 0;
        }
        out.tile_info[tile] = make_int2(tot.x, tot.z);
    }
    __syncthreads();
    for (int k = 0; k < spt; k++) {
        const int c = (int)(hcc[s0 + k] & 0xffffu);
        if (c > 0) {
            hrow[s0 + k] = (unsigned short)pre3.z;  // the vertex's row slot in the tile
            out.tile_vert[tb + pre3.z] = hkeys[s0 + k];
            pre3.z++;
            hcc[s0 + k] = (unsigned)c | ((unsigned)pre3.y << 16);  // fill cursor of the slot's pair list
            pre3.y += c;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < npairs; i += 256) {
        const int j = i / TP, lp = i - j * TP, p = tile_point(tm, org, lp);
        float ws = 0.f;
        int slot = 0;
        if (p >= 0) {
            const int h = pslot[i];
            const int pos = (int)(atomicAdd(&hcc[h], 0x10000u) >> 16);
            const float b = bary[(size_t)p * D1 + j];
            const float nv = (pre | post) ? norm[p] : 1.f;
            spairs[pos] = make_uint2((unsigned)q_row(lp), __float_as_uint(pre ? __fmul_rn(b, nv) : b));  // row in the Q tile
            ws = __fmul_rn(post ? __fmul_rn(b, nv) : b, slice_scale);
            slot = hrow[h];
        }
        out.pt_w[tb + pt_index(D1, j, lp)] = ws;
        out.pt_slot[tb + pt_index(D1, j, lp)] = (uint16_t)slot;
    }
    __syncthreads();
    // Segments, and the ORDER of the pairs inside a segment.  In the gather, thread e walks segment e; at step k the eight
    // lanes of a quarter-warp read one 16-byte unit of eight different Q rows (an LDS.128 is served per quarter-warp).
    // Row lp, unit g lies in bank group (G * lp + g) mod 8, so the eight reads are conflict-free when the eight lp differ
    // mod 8 (G odd).  Every list is therefore arranged so that position k of segment e holds a point of class
    // (e + k) mod 8 whenever the list still has one - a vertex covers a blob of pixels, so all classes occur about equally.
    const int G = row_bytes / 16;
    for (int k = 0; k < spt; k++) {
        const unsigned cc = hcc[s0 + k];
        const int c = (int)(cc & 0xffffu);
        if (c > 0) {
            const int key = hkeys[s0 + k], start = (int)(cc >> 16) - c;
            for (int o = 0; o < c; o += TILE_SEG) {
                const int len = min(TILE_SEG, c - o);
                const int idx = binstart[len] + atomicAdd(&hist[len], 1);
                const int m = (start + o) | (len << 16);
                reinterpret_cast<int*>(out.ent_meta + tb + idx)[1] = key;  // .x follows in the ordering loop
                segs[idx] = m;
            }
        }
    }
    __syncthreads();
    // one segment per thread: the classes of its <= 32 pairs packed as nibbles in two 64-bit registers (unused nibbles are
    // 0xF, which matches no class), so "the pairs of class c" is a zero-nibble mask, not a scan.
    // STORAGE: the 32 consecutive segments a warp of the gather walks form a group, stored COLUMN-major - first pair 0 of
    // every segment of the group, then pair 1 of every segment that has one, ... (segments are ordered longest first, so
    // the segments that still have a pair q are a prefix of the group and a column is compact).  At step q the lanes of a
    // warp then read consecutive 8-byte pairs: no bank conflicts (row-major lists with arbitrary starts: 3.6 wavefronts
    // per 8-byte read).  ent_meta.x = first pair of the segment's GROUP | length << 16.
    static_assert(RSS_BLUR_ALONE_T <= RSS_BLUR_MAXT && RSS_BLUR_SHARED_T <= RSS_BLUR_MAXT, "blur CTA sizes exceed the launch bound");
    static_assert(TILE_SEG <= 32, "the segment ordering packs 32 classes into 128 bits");
    const int lane = threadIdx.x & 31;
    for (int i0 = threadIdx.x - lane; i0 < tot.x; i0 += 256) {
        const int idx = i0 + lane;
        int len = 0, start = 0;
        if (idx < tot.x) { const int m = segs[idx]; len = m >> 16; start = m & 0xffff; }
        const int len0 = __shfl_sync(0xffffffffu, len, 0);  // the group's longest segment
        int col = cumstart[len0] + (i0 - binstart[len0]) * len0;  // pairs before the group
        if (idx < tot.x) reinterpret_cast<int*>(out.ent_meta + tb + idx)[0] = col | (len << 16);
        const uint2* lst = spairs + start;
        // column starts of the group: position q of every segment lies at colw[q] + lane (lane q keeps the start it sees)
        int mycol = 0;
        for (int q = 0; q < len0; q++) {
            if (lane == q) mycol = col;
            col += __popc(__ballot_sync(0xffffffffu, q < len));
        }
        int* colw = colstart + (threadIdx.x >> 5) * 32;
        __syncwarp();
        colw[lane] = mycol;
        __syncwarp();
        if (len == 0) continue;  // (no warp-level operation below)
        // the classes of the pairs as nibbles in two 64-bit registers (never indexed dynamically: that would be local memory)
        unsigned long long c0 = ~0ull, c1 = ~0ull;
        for (int f = 0; f < len; f++) {
            const int sh = 4 * (f & 15);
            const unsigned long long keep = ~(0xFull << sh), val = (unsigned long long)((G * lst[f].x) & 7u) << sh;
            if (f < 16) c0 = (c0 & keep) | val;
            else c1 = (c1 & keep) | val;
        }
        // the positions that want class c: want(q) = G (idx + q) mod 8 has period 8 in q
        unsigned posmask[8];
#pragma unroll
        for (int c = 0; c < 8; c++) posmask[c] = 0u;
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const unsigned w = (unsigned)(G * (idx + q)) & 7u;
#pragma unroll
            for (int c = 0; c < 8; c++)
                if (w == (unsigned)c) posmask[c] |= 0x01010101u << q;
        }
        const unsigned lenmask = len >= 32 ? 0xffffffffu : ((1u << len) - 1u);
        unsigned filled = 0u;
        unsigned long long l0 = 0ull, l1 = 0ull;  // members that found no position of their class (bit 4 f' + 3 of word f / 16)
        auto place = [&](int f, int q) {
            const uint2 v = lst[f];
            out.pairs[tb + colw[q] + lane] = make_uint2(v.x * (unsigned)row_bytes, v.y);
        };
        // class by class (static unroll: every mask stays in a register): the k-th pair of the class takes the k-th position
        // that wants it
#pragma unroll
        for (int c = 0; c < 8; c++) {
            const unsigned long long want = (unsigned long long)c * 0x1111111111111111ull;
            const unsigned long long x0 = c0 ^ want, x1 = c1 ^ want;  // zero nibble <=> pair of class c
            // EXACT zero-nibble mask (bit 3 of every zero nibble): no carries between nibbles, unlike the (x - 0x11..) & ~x
            // test, which is only right about the LOWEST zero nibble
            constexpr unsigned long long N7 = 0x7777777777777777ull, N8 = 0x8888888888888888ull;
            unsigned long long z0 = ~(((x0 & N7) + N7) | x0) & N8;
            unsigned long long z1 = ~(((x1 & N7) + N7) | x1) & N8;
            unsigned pos = posmask[c] & lenmask;
            while (pos && (z0 | z1)) {
                int f;
                if (z0) { f = (__ffsll((long long)z0) - 1) >> 2; z0 &= z0 - 1; }
                else { f = 16 + ((__ffsll((long long)z1) - 1) >> 2); z1 &= z1 - 1; }
                const int q = __ffs(pos) - 1;
                pos &= pos - 1;
                filled |= 1u << q;
                place(f, q);
            }
            l0 |= z0; l1 |= z1;
        }
        // the classes that ran out leave holes; the left-over pairs fill them in order
        unsigned holes = lenmask & ~filled;
        while (holes) {
            int f;
            if (l0) { f = (__ffsll((long long)l0) - 1) >> 2; l0 &= l0 - 1; }
            else { f = 16 + ((__ffsll((long long)l1) - 1) >> 2); l1 &= l1 - 1; }
            const int q = __ffs(holes) - 1;
            holes &= holes - 1;
            place(f, q);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Blur of every lattice of the CRF in ONE cooperative launch (permutohedral.cpp:555-569).  Phase j blurs axis j of
// each lattice that has one; phases are separated by a grid barrier on an L2 counter (lattice.cuh).  A phase is
// L2-throughput-bound (random 16*G-byte rows): tools/micro/blur_bench.cu.  A MISSING neighbour is the zero row
// (index vcap): its load is skipped - on sparse lattices a large share of the neighbours is missing, and reading them
// would send all those requests to the one L2 slice that holds the zero row.
// Phase 0 additionally clears `zero` (the table the point kernel sliced from one iteration ago), which becomes the
// next splat target - so no phase follows the last axis.
// ---------------------------------------------------------------------------------------------------------------
// one axis of one lattice: items are (vertex, channel group); U independent items per thread are in flight at once
template <int U>
__device__ __forceinline__ void blur_axis(const float4* __restrict__ src, float4* __restrict__ dst, const int2* __restrict__ nb_j,
                                          uint32_t items, int G, int vcap, uint32_t tid, uint32_t nthr) {
    for (uint32_t base = tid; base < items; base += U * nthr) {
        int2 nb[U];
        uint32_t gq[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint32_t it = base + u * nthr;
            if (it < items) {
                const uint32_t v = it / (uint32_t)G;
                gq[u] = it - v * (uint32_t)G;
                nb[u] = __ldg(nb_j + v);
            }
        }
        float4 o[U], x[U], y[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint32_t it = base + u * nthr;
            x[u] = y[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (it < items) {
                o[u] = __ldcg(src + it);
                if (nb[u].x != vcap) x[u] = __ldcg(src + (size_t)nb[u].x * G + gq[u]);
                if (nb[u].y != vcap) y[u] = __ldcg(src + (size_t)nb[u].y * G + gq[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint32_t it = base + u * nthr;
            if (it < items) __stcg(dst + it, blur_item(o[u], x[u], y[u]));
        }
    }
}
// F consecutive axes in ONE phase: the value of vertex v after axes j0 .. j0+F-1 is computed recursively from the 3^F raw
// values of its neighbourhood, with exactly the arithmetic of F separate passes (every intermediate is the same
// blur_item of the same operands, so the result is bit-identical); a missing vertex (index vcap) is not part of the
// lattice and contributes 0 whatever its own neighbours are.  This trades L2 reads (3^F instead of 3 F rows per item,
// the L2 is at 15 % of its throughput in the one-axis-per-phase kernel) for grid barriers, each of which costs a full
// store -> fence -> atomic -> poll -> load round trip of ~5 us on 148 SMs.
template <int F>
__device__ __forceinline__ float4 blur_value(const float4* __restrict__ src, const int2* __restrict__ nbr, int j0, int vcap, int G,
                                             int v, int g) {
    if (v == vcap) return make_float4(0.f, 0.f, 0.f, 0.f);
    if constexpr (F == 0) {
        return __ldcg(src + (size_t)v * G + g);
    } else {
        const int2 nb = __ldg(nbr + (size_t)(j0 + F - 1) * vcap + v);
        const float4 o = blur_value<F - 1>(src, nbr, j0, vcap, G, v, g);
        const float4 x = blur_value<F - 1>(src, nbr, j0, vcap, G, nb.x, g);
        const float4 y = blur_value<F - 1>(src, nbr, j0, vcap, G, nb.y, g);
        return blur_item(o, x, y);
    }
}
template <int F, int U>
__device__ __forceinline__ void blur_axes(const float4* __restrict__ src, float4* __restrict__ dst, const int2* __restrict__ nbr,
                                          int j0, uint32_t items, int G, int vcap, uint32_t tid, uint32_t nthr) {
    for (uint32_t base = tid; base < items; base += U * nthr) {  // U independent items per trip: their load chains overlap
        float4 r[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint32_t it = base + u * nthr;
            if (it < items) {
                const uint32_t v = it / (uint32_t)G, g = it - v * (uint32_t)G;
                r[u] = blur_value<F>(src, nbr, j0, vcap, G, (int)v, (int)g);
            }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint32_t it = base + u * nthr;
            if (it < items) __stcg(dst + it, r[u]);
        }
    }
}
__global__ void __launch_bounds__(RSS_BLUR_MAXT) blur_multi_coop_kernel(const __grid_constant__ BlurMultiArgs a, int G,
                                                              unsigned int* barrier, unsigned int barrier_base) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, nthr = gridDim.x * blockDim.x;
    uint32_t V[FUSED_MAX_LAT];
    int j0[FUSED_MAX_LAT];
    for (int k = 0; k < FUSED_MAX_LAT; k++) {
        V[k] = k < a.K && !a.counts[k][1] ? a.counts[k][0] : 0u;
        j0[k] = 0;
    }
    for (int p = 0; p < a.phases; p++) {
        for (int k = 0; k < a.K; k++) {
            const int f = a.fuse[k][p];
            if (f == 0) continue;
            const float4* src = (p & 1) ? a.pong[k] : a.ping[k];
            float4* dst = (p & 1) ? a.ping[k] : a.pong[k];
            const uint32_t items = V[k] * (uint32_t)G;
            if (f == 1) blur_axis<RSS_BLUR_U>(src, dst, a.nbr[k] + (size_t)j0[k] * a.vcap[k], items, G, (int)a.vcap[k], tid, nthr);
            else if (f == 2) blur_axes<2, RSS_BLUR_FU>(src, dst, a.nbr[k], j0[k], items, G, (int)a.vcap[k], tid, nthr);
            else blur_axes<3, 1>(src, dst, a.nbr[k], j0[k], items, G, (int)a.vcap[k], tid, nthr);
            j0[k] += f;
            if (p == 0 && a.zero[k])
                for (uint32_t it = tid; it < items; it += nthr) __stcg(a.zero[k] + it, make_float4(0.f, 0.f, 0.f, 0.f));
        }
        if (p + 1 < a.phases) grid_barrier(barrier, barrier_base + (unsigned int)(p + 1) * gridDim.x);
    }
}

// splat of the all-ones vector for the normalisation (pairwise.cpp:44): values[v][0] += sum of barycentric weights.
// Run-accumulation over consecutive points, one thread per chunk, rows of 4 floats with channel 0 live.
template <int D1>
__global__ void __launch_bounds__(256) splat_ones_runs_kernel(const int* __restrict__ offsets, const float* __restrict__ bary,
                                                              int N, int Pc, const uint32_t* __restrict__ counts,
                                                              float* __restrict__ values) {
    if (counts[1]) return;
    const long long chunk = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (chunk * Pc >= N) return;
    const int p0 = (int)(chunk * Pc), p1 = min(N, p0 + Pc);
    int cur[D1];
    float acc[D1];
#pragma unroll
    for (int j = 0; j < D1; j++) { cur[j] = -1; acc[j] = 0.f; }
    for (int p = p0; p < p1; p++) {
        int key[D1];
        float w[D1];
        load_row_i<D1>(offsets + (size_t)p * D1, key);
        load_row_f<D1>(bary + (size_t)p * D1, w);
#pragma unroll
        for (int j = 0; j < D1; j++) {
            if (key[j] != cur[j]) {
                if (cur[j] >= 0) atomicAdd(values + (size_t)cur[j] * 4, acc[j]);
                cur[j] = key[j];
                acc[j] = 0.f;
            }
            acc[j] += w[j];
        }
    }
#pragma unroll
    for (int j = 0; j < D1; j++)
        if (cur[j] >= 0) atomicAdd(values + (size_t)cur[j] * 4, acc[j]);
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
TileMap fused_tile_map(int N, int W, int H) {
    TileMap m;
    m.N = N;
    m.TP = TILE_POINTS;
    m.perm = nullptr;
    if (W > 0 && H > 0 && (long long)W * H == N) {
        m.W = W; m.H = H; m.TW = TILE_W; m.TH = TILE_H;
        m.tiles_x = (W + m.TW - 1) / m.TW;
        m.ntiles = m.tiles_x * ((H + m.TH - 1) / m.TH);
    } else {
        m.W = m.H = 0; m.TW = m.TH = 0; m.tiles_x = 0;
        m.ntiles = (int)(((long long)N + m.TP - 1) / m.TP);
    }
    return m;
}

bool fused_group_supported(int G) { return G >= 1 && G <= 6; }
bool fused_signature_supported(int G, int d1a, int d1b) {
    if (!fused_group_supported(G)) return false;
    switch (d1a * 16 + d1b) {
        case 0x46: case 0x36: case 0x70: case 0x60: case 0x40: case 0x30: return true;
        default: return false;
    }
}

cudaError_t launch_meanfield_fused(rss_ctx* c, cudaStream_t st, const FusedArgs& a, int d1a, int d1b, const float* unary, float* Q,
                                   uint8_t* labels, const TileMap& tm, int G, const FusedLayers& ls, int mode) {
    // several keyframes in flight on this GPU: the 64-register build, which shares an SM with a blur CTA (meanfield_point.inl)
    if (live_contexts(c->device).load() > 1)
        return launch_meanfield_fused_shared(c, st, a, d1a, d1b, unary, Q, labels, tm, G, ls, mode);
    return launch_meanfield_fused_alone(c, st, a, d1a, d1b, unary, Q, labels, tm, G, ls, mode);
}

void launch_tile_csr_build(rss_ctx* c, cudaStream_t st, const int* offsets, const float* bary, const float* norm, bool pre,
                           bool post, float slice_scale, const TileMap& tm, int d1, int row_bytes, const uint32_t* counts,
                           const TileCsrOut& out) {
    const int grid = tm.ntiles, TP = TILE_POINTS;
    int HC = 1024;  // power of two with load factor <= 0.8 even when every pair of the tile hits a different vertex
    while (HC * 4 < TP * d1 * 5) HC *= 2;
    const size_t tsm = (size_t)HC * (4 + 4 + 2) + (size_t)TP * d1 * (2 + 8 + 4);
    const int ipre = pre ? 1 : 0, ipost = post ? 1 : 0;
#define RSS_TCB(D)                                                                                                      \
    do {                                                                                                                \
        if (c->smem_attr_done.insert((const void*)tile_csr_build_kernel<D>).second) { /* once per context (= per device) */ \
            cudaFuncSetAttribute(tile_csr_build_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);     \
            cudaFuncSetAttribute(tile_csr_build_kernel<D>, cudaFuncAttributePreferredSharedMemoryCarveout,               \
                                 cudaSharedmemCarveoutMaxShared);                                                       \
        }                                                                                                               \
        RSS_LAUNCH(c, tile_csr_build_kernel<D>, grid, 256, tsm, st, offsets, bary, norm, ipre, ipost, slice_scale, tm,  \
                   row_bytes, HC, counts, out);                                                                         \
    } while (0)
    switch (d1) {
        case 2: RSS_TCB(2); break;
        case 3: RSS_TCB(3); break;
        case 4: RSS_TCB(4); break;
        case 5: RSS_TCB(5); break;
        case 6: RSS_TCB(6); break;
        case 7: RSS_TCB(7); break;
        default: RSS_TCB(8); break;
    }
#undef RSS_TCB
}

// Launch shape of the cooperative blur.  Alone on the GPU: one CTA of RSS_BLUR_ALONE_T threads per SM.  Sharing the GPU with
// other keyframes in flight: RSS_BLUR_SHARED_T threads per CTA on sm_count / RSS_BLUR_SHARED_GRID_DIV SMs - the blur is
// resident most of the time then, and what it costs the other keyframes' kernels is the CTA slots its registers block.
BlurShape blur_multi_shape(const rss_ctx* c) {
    BlurShape s;
    if (live_contexts(c->device).load() > 1) {
        s.grid = std::max(1, c->sm_count / RSS_BLUR_SHARED_GRID_DIV);
        s.block = RSS_BLUR_SHARED_T;
    } else {
        s.grid = c->sm_count;
        s.block = RSS_BLUR_ALONE_T;
    }
    return s;
}
// Splits the d+1 axes of every lattice into phases of up to `RSS_BLUR_FUSE` fused axes (fewer grid barriers, more L2
// reads): 3 axes per phase for tables of at most RSS_BLUR_FUSE3_ITEMS float4 items, 2 up to RSS_BLUR_FUSE2_ITEMS, else 1.
// Returns the number of phases of the launch; phases_of[k] = phases lattice k takes part in (the blurred table is `ping`
// when that is even, else `pong`).
int blur_multi_plan(BlurMultiArgs& a, int G, int* phases_of) {
    // phases of the launch = the most any lattice needs at its own fusion limit ...
    int fmax[FUSED_MAX_LAT] = {0};
    a.phases = 0;
    for (int k = 0; k < a.K; k++) {
        const size_t items = (size_t)a.vcap[k] * G;
        fmax[k] = std::min(RSS_BLUR_FUSE, items <= RSS_BLUR_FUSE3_ITEMS ? 3 : (items <= RSS_BLUR_FUSE2_ITEMS ? 2 : 1));
        a.phases = std::max(a.phases, (a.d1[k] + fmax[k] - 1) / fmax[k]);
    }
    // ... and every lattice spreads its axes over ALL of them, as evenly as possible with the larger chunks first: a lattice
    // with fewer axes then fuses fewer per phase (4 axes over 3 phases: 2 + 1 + 1 instead of 2 + 2 + idle), which reads
    // 3^F rows per item less often - the phases are L2-throughput-bound - and evens the phases out.
    for (int k = 0; k < FUSED_MAX_LAT; k++) {
        for (int p = 0; p < BLUR_MAX_PHASES; p++) a.fuse[k][p] = 0;
        if (k >= a.K) continue;
        const int np = std::min(a.phases, a.d1[k]);
        for (int p = 0, left = a.d1[k]; p < np; p++) {
            const int f = (left + (np - p) - 1) / (np - p);
            a.fuse[k][p] = f;
            left -= f;
        }
        phases_of[k] = np;
    }
    return a.phases;
}
cudaError_t launch_blur_multi(rss_ctx* c, cudaStream_t st, BlurMultiArgs a, int G, unsigned int* barrier, unsigned int barrier_base,
                              const BlurShape& shape) {
    const int grid = shape.grid, block = shape.block;
    void* args[] = {&a, &G, &barrier, &barrier_base};
    cudaEvent_t ea = nullptr, eb = nullptr;
    if (c->profile) { ea = c->prof_event(); eb = c->prof_event(); cudaEventRecord(ea, st); }
    const cudaError_t e = cudaLaunchCooperativeKernel((const void*)blur_multi_coop_kernel, dim3(grid), dim3(block), args, 0, st);
    c->launches++;
    if (c->profile) { cudaEventRecord(eb, st); c->prof_pending.push_back(rss_ctx::Pending{"blur_multi_coop_kernel", ea, eb}); }
    return e;
}

void launch_splat_ones_runs(rss_ctx* c, cudaStream_t st, const int* offsets, const float* bary, int N, int d1,
                            const uint32_t* counts, float* values) {
    const int Pc = 4;  // points per thread: short runs, but 4x the threads of a 16-point chunk (the kernel is latency-bound)
    const int grid = rss_div_up(rss_div_up(N, Pc), 256);
#define RSS_ONES(D) RSS_LAUNCH(c, splat_ones_runs_kernel<D>, grid, 256, 0, st, offsets, bary, N, Pc, counts, values)
    switch (d1) {
        case 2: RSS_ONES(2); break;
        case 3: RSS_ONES(3); break;
        case 4: RSS_ONES(4); break;
        case 5: RSS_ONES(5); break;
        case 6: RSS_ONES(6); break;
        case 7: RSS_ONES(7); break;
        default: RSS_ONES(8); break;
    }
#undef RSS_ONES
}

}  // namespace rss
