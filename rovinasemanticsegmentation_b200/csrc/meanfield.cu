// Fused mean-field iteration for point sets whose order is spatially coherent (image raster order: the keyframe path,
// DenseCRF2D; any point set whose consecutive points share lattice vertices).  Per iteration there are two launches:
//   meanfield_tile_kernel   slice of the previous iteration's blurred value tables + Potts + unary + per-layer soft-max
//                           (+ gated argmax / Q store on the last pass) AND the splat of the new marginals;
//   blur_multi_coop_kernel  every axis of every lattice of the CRF, one cooperative launch with grid barriers.
// Reference: third-party/densecrf/src/densecrf.cpp:98-131 (expAndNormalize, inference), pairwise.cpp:63-80
// (DenseKernel::filter), labelcompatibility.cpp:46-48 (Potts), permutohedral.cpp:529-589 (sseCompute).
//
// A tile (one CTA of the point kernel) is a block of ~512 points: a 32 x TH pixel block for image CRFs, else a run of
// consecutive points.  Phase 1 keeps the tile's marginals in shared memory; phase 2 gathers them per (lattice vertex
// touched by the tile) through a tile-local CSR built once per lattice (tile_csr_build_kernel) and issues one
// red.global.add.v4.f32 per (segment of <= 32 pairs, channel group).  So Q is never re-read from L2, there is no
// separate splat launch, and the number of global atomics is (distinct vertices per tile), not (nonzeros) - which is
// what makes noisy images cheap: +-8 colour noise sends 64 % of consecutive pixels to a different bilateral vertex.
// Value tables rotate through three buffers per lattice: `res` (blurred result being sliced), `tgt` (all zero, receives
// the splat) and `spare`; the blur ping-pongs between tgt and spare and clears the old `res`, the next tgt.
//
// Compile-time tuning knobs (build.py: RSS_NVCC_DEFS="-DNAME=value"), defaults measured on B200:
#include <algorithm>

#include "kernels.hpp"
#include "lattice.cuh"
#include "meanfield.cuh"

#ifndef RSS_TILE_MINB
#define RSS_TILE_MINB 4  // resident CTAs per SM the point kernel is compiled for (register budget)
#endif
#ifndef RSS_TILE_REGS  // explicit register cap instead of the occupancy-derived one
#define RSS_TILE_BOUNDS __launch_bounds__(256, RSS_TILE_MINB)
#else
#define RSS_TILE_BOUNDS __maxnreg__(RSS_TILE_REGS)
#endif
#ifndef RSS_BLUR_U
#define RSS_BLUR_U 2      // independent (vertex, channel group) items a blur thread keeps in flight
#endif
#ifndef RSS_BLUR_MAXT
#define RSS_BLUR_MAXT 128  // CTA size of the cooperative blur: small register / thread footprint on purpose, so that
                           // kernels of OTHER keyframes in flight run on the SMs while its CTAs wait at the grid barriers
#endif
#ifndef RSS_TILE_SINGLE_WAVE
#define RSS_TILE_SINGLE_WAVE 0  // grow the tiles until all CTAs of the point kernel are resident at once
#endif
#ifndef RSS_TILE_IU
#define RSS_TILE_IU 1  // splat segments a thread walks at once
#endif
#ifndef RSS_TILE_POINTS
#define RSS_TILE_POINTS 288  // target points per tile (<= 512: the tile CSR build kernel's hash capacity)
#endif

namespace rss {

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
__device__ __forceinline__ void red_add_v4(float* dst, const float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// ---------------------------------------------------------------------------------------------------------------
// Tile-local CSR of the splat matrix.  A tile is TP consecutive points (the points one CTA of the mean-field kernel
// owns).  For every tile: the distinct lattice vertices its points touch ("entries") and, per entry, the list of
// (local point, weight) pairs.  Built once per lattice by one CTA per tile with a shared-memory hash table (native
// 32-bit CAS / integer adds only): insert keys -> count -> compact + scan -> fill.  weight = bary * norm_i when the
// kernel is pre-normalised (DenseKernel::filter, pairwise.cpp:65-66), so the gather needs no norm lookup.
// Entry meta: x = (start of the segment in the tile's pair array) | (length << 16), y = vertex id.
// ---------------------------------------------------------------------------------------------------------------
#ifndef RSS_TILE_SEG
#define RSS_TILE_SEG 32
#endif
constexpr int TILE_SEG = RSS_TILE_SEG;  // pairs per splat segment (one thread walks one segment serially)
constexpr int TILE_CHUNK = 8;           // pairs requested at once by the gather

__device__ __forceinline__ int2 block_excl_scan2(int a, int b, int2* total) {  // 256 threads
    __shared__ int2 wsum[8];
    __shared__ int2 btot;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int ia = a, ib = b;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int ta = __shfl_up_sync(0xffffffffu, ia, o), tb = __shfl_up_sync(0xffffffffu, ib, o);
        if (lane >= o) { ia += ta; ib += tb; }
    }
    if (lane == 31) wsum[w] = make_int2(ia, ib);
    __syncthreads();
    if (threadIdx.x == 0) {
        int sa = 0, sb = 0;
        for (int k = 0; k < 8; k++) { const int2 v = wsum[k]; wsum[k] = make_int2(sa, sb); sa += v.x; sb += v.y; }
        btot = make_int2(sa, sb);
    }
    __syncthreads();
    const int2 r = make_int2(ia - a + wsum[w].x, ib - b + wsum[w].y);
    *total = btot;
    __syncthreads();
    return r;
}

template <int D1>
__global__ void __launch_bounds__(256) tile_csr_build_kernel(const int* __restrict__ offsets, const float* __restrict__ bary,
                                                             const float* __restrict__ norm, const TileMap tm, int row_bytes,
                                                             int HC, const uint32_t* __restrict__ counts,
                                                             uint2* __restrict__ pairs, int2* __restrict__ ent_meta,
                                                             int* __restrict__ tile_nent) {
    __shared__ int hist[TILE_SEG + 1], binstart[TILE_SEG + 1];
    extern __shared__ int tile_smem[];  // TILE_SMEM_BYTES, above the 48 KB static limit
    int* hkeys = tile_smem;  // HC slots (power of two, > pairs per tile)
    int* hcnt = tile_smem + HC;
    unsigned short* pslot = reinterpret_cast<unsigned short*>(tile_smem + 2 * HC);
    const int hshift = 32 - (31 - __clz(HC));
    const int tile = blockIdx.x;
    const int TP = tm.TP;
    const TileOrigin org = tile_origin(tm, tile);
    const int npairs = TP * D1;  // slots; points outside the image / beyond N are skipped
    const size_t tb = (size_t)tile * TP * D1;
    if (counts[1]) {
        if (threadIdx.x == 0) tile_nent[tile] = 0;
        return;
    }
    for (int i = threadIdx.x; i < HC; i += 256) { hkeys[i] = -1; hcnt[i] = 0; }
    if (threadIdx.x <= TILE_SEG) hist[threadIdx.x] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < npairs; i += 256) {
        const int lp = i / D1, p = tile_point(tm, org, lp);
        if (p < 0) continue;
        const int key = offsets[(size_t)p * D1 + (i - lp * D1)];
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
        atomicAdd(&hcnt[h], 1);
    }
    __syncthreads();
    // Lists are cut into segments of at most TILE_SEG pairs and the segments are ordered by length (longest first), so
    // that the lanes of a warp of the gather walk lists of (nearly) equal length.  Thread t owns slots [32t, 32t+32).
    int nseg = 0, ncnt = 0;
    const int spt = HC / 256, s0 = threadIdx.x * spt;  // slots per thread
#pragma unroll 4
    for (int k = 0; k < spt; k++) {
        const int c = hcnt[s0 + k];
        if (c > 0) {
            const int full = c / TILE_SEG, rem = c - full * TILE_SEG;
            if (full) atomicAdd(&hist[TILE_SEG], full);
            if (rem) atomicAdd(&hist[rem], 1);
            nseg += full + (rem ? 1 : 0);
            ncnt += c;
        }
    }
    int2 tot;
    int2 pre = block_excl_scan2(nseg, ncnt, &tot);  // ends with a barrier: hist is complete
    if (threadIdx.x == 0) {
        int pos = 0;
        for (int len = TILE_SEG; len >= 1; len--) { binstart[len] = pos; pos += hist[len]; hist[len] = 0; }
        tile_nent[tile] = tot.x;
    }
    __syncthreads();
    for (int k = 0; k < spt; k++) {
        const int c = hcnt[s0 + k];
        if (c > 0) {
            const int key = hkeys[s0 + k];
            for (int o = 0; o < c; o += TILE_SEG) {
                const int len = min(TILE_SEG, c - o);
                const int idx = binstart[len] + atomicAdd(&hist[len], 1);
                ent_meta[tb + idx] = make_int2((pre.y + o) | (len << 16), key);
            }
            hcnt[s0 + k] = pre.y;  // becomes the fill cursor of the slot's pair list
            pre.y += c;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < npairs; i += 256) {
        const int lp = i / D1, p = tile_point(tm, org, lp);
        if (p < 0) continue;
        const int pos = atomicAdd(&hcnt[pslot[i]], 1);
        float w = bary[(size_t)p * D1 + (i - lp * D1)];
        if (norm) w = __fmul_rn(w, norm[p]);
        pairs[tb + pos] = make_uint2((unsigned)lp * (unsigned)row_bytes, __float_as_uint(w));
    }
}

// ---------------------------------------------------------------------------------------------------------------
// The point kernel.  One CTA = one tile of TP consecutive points.
//  phase 1 (lane = (point, channel group), G lanes per point, 32/G points per warp-step):
//     t = -unary + sum over lattices of w * norm_i * sum_j (bary_ij * alpha) * blurred[vertex_ij]; soft-max per label
//     layer across the G lanes; marginals go to the shared-memory tile (and to global Q / label maps on request).
//     A value-table row (Mp floats) is read by G adjacent lanes as one contiguous 16*G-byte access.
//  phase 2 (item = (tile entry, channel group)): sum w * Q[point] over the entry's pair list out of shared memory and
//     issue ONE red.global.add.v4.f32 per item.  Global atomics per iteration = (distinct vertices per tile) * G
//     instead of (nonzeros) * G, Q is never re-read from L2, and nothing depends on how noisy the point order is.
// `mode` bit 0: slice (not the first pass), bit 1: splat (not the last pass), bit 2: store Q.
// ---------------------------------------------------------------------------------------------------------------
// streaming inputs of one point for one lattice, loaded one step ahead of their use
template <int D1>
struct PointIn {
    int key[D1 > 0 ? D1 : 1];
    float w[D1 > 0 ? D1 : 1];
    float nrm;
    __device__ __forceinline__ void load(const FusedLat& L, int p) {
        if constexpr (D1 > 0) {
            load_row_i<D1>(L.offsets + (size_t)p * D1, key);
            load_row_f<D1>(L.bary + (size_t)p * D1, w);
            nrm = __ldg(L.norm + p);
        }
    }
};
template <int D1>
__device__ __forceinline__ void slice_lattice(const FusedLat& L, const PointIn<D1>& in, int Mp, int g, float4& t) {
    float4 row[D1];
#pragma unroll
    for (int j = 0; j < D1; j++) row[j] = __ldg(reinterpret_cast<const float4*>(L.vin + (size_t)in.key[j] * Mp) + g);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < D1; j++) {
        acc.x = fmaf(in.w[j], row[j].x, acc.x); acc.y = fmaf(in.w[j], row[j].y, acc.y);
        acc.z = fmaf(in.w[j], row[j].z, acc.z); acc.w = fmaf(in.w[j], row[j].w, acc.w);
    }
    // tmp = -unary - (-w * (alpha * filtered) * norm)   (densecrf.cpp:126, pairwise.cpp:78-79, permutohedral.cpp:571)
    const float c = L.potts * L.alpha * (L.post ? in.nrm : 1.f);
    t.x = fmaf(c, acc.x, t.x); t.y = fmaf(c, acc.y, t.y); t.z = fmaf(c, acc.z, t.z); t.w = fmaf(c, acc.w, t.w);
}

// exp(x) for x <= 0 as one multiply and one ex2.approx.ftz (|rel err| < 2e-6; the CRF tolerance is 1e-4 abs; results
// below 2^-126 flush to zero, which the normalisation cannot tell from the true value)
__device__ __forceinline__ float fast_exp(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x * 1.4426950408889634f));
    return y;
}
template <int G>
__device__ __forceinline__ float group_gather_max(float v, int gbase) {
    float m = v;
#pragma unroll
    for (int i = 0; i < G; i++) m = fmaxf(m, __shfl_sync(0xffffffffu, v, (gbase + i) & 31));
    return m;
}
template <int G>
__device__ __forceinline__ float group_gather_sum(float v, int gbase) {  // same order on every lane of the group
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < G; i++) s += __shfl_sync(0xffffffffu, v, (gbase + i) & 31);
    return s;
}

// pairs[].x holds the BYTE offset of the local point's row inside the shared tile (lp * G * 16).
// Entry meta of the tile was staged into shared memory at kernel start (`cap` entries; the rest comes from global).
// Every thread walks IU segments at once (independent load streams); segments are sorted by length, so the lanes of
// a warp finish together.
template <int G, int IU>
__device__ __forceinline__ void gather_entries(const uint2* pr, const int2* meta, int cap,
                                               const int2* __restrict__ meta_g, int ne, float* __restrict__ vout,
                                               const float4* qtile) {
    constexpr int Mp = 4 * G;
    const char* qbytes = reinterpret_cast<const char*>(qtile);
    const int items = ne * G;
    for (int it0 = threadIdx.x; it0 < items; it0 += 256 * IU) {
        const uint2* pp[IU];
        int len[IU], vertex[IU];
        const char* qb[IU];
        float4 acc[IU];
        int longest = 0;
#pragma unroll
        for (int u = 0; u < IU; u++) {
            const int it = it0 + u * 256;
            len[u] = 0; vertex[u] = -1; pp[u] = pr; qb[u] = qbytes;
            acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (it < items) {
                const int e = it / G;
                const int2 m = e < cap ? meta[e] : __ldg(meta_g + e);
                pp[u] = pr + (m.x & 0xffff);
                len[u] = m.x >> 16;
                vertex[u] = m.y;
                qb[u] = qbytes + 16 * (it - e * G);
                longest = max(longest, len[u]);
            }
        }
        // pairs are requested TILE_CHUNK at a time (independent loads, one L2 round trip per chunk), then the
        // multiply-adds run out of shared memory
        for (int c0 = 0; c0 < longest; c0 += TILE_CHUNK) {
            uint2 pw[IU][TILE_CHUNK];
#pragma unroll
            for (int u = 0; u < IU; u++)
#pragma unroll
                for (int i = 0; i < TILE_CHUNK; i++)
                    if (c0 + i < len[u]) pw[u][i] = pp[u][c0 + i];
#pragma unroll
            for (int i = 0; i < TILE_CHUNK; i++) {
#pragma unroll
                for (int u = 0; u < IU; u++) {
                    if (c0 + i < len[u]) {
                        const float w = __uint_as_float(pw[u][i].y);
                        const float4 q = *reinterpret_cast<const float4*>(qb[u] + pw[u][i].x);
                        acc[u].x = fmaf(w, q.x, acc[u].x); acc[u].y = fmaf(w, q.y, acc[u].y);
                        acc[u].z = fmaf(w, q.z, acc[u].z); acc[u].w = fmaf(w, q.w, acc[u].w);
                    }
                }
            }
        }
#pragma unroll
        for (int u = 0; u < IU; u++)
            if (vertex[u] >= 0) red_add_v4(vout + (size_t)vertex[u] * Mp + (qb[u] - qbytes) / 4, acc[u]);
    }
}

// MODE is a compile-time constant (2 = first pass, 3 = middle passes, 5 = last pass): the slice / splat / store branches
// disappear from the instruction stream of each variant.
template <int G, int D1A, int D1B, int MODE>
__global__ void RSS_TILE_BOUNDS meanfield_tile_kernel(const __grid_constant__ FusedArgs a,
                                                             const float* __restrict__ unary, float* __restrict__ Q,
                                                             uint8_t* __restrict__ labels,
                                                             const __grid_constant__ TileMap tm, int steps,
                                                             const __grid_constant__ FusedLayers ls) {
    extern __shared__ float4 qtile[];  // [TP][G]
    if (a.counts[0][1]) return;  // lattice overflow: the host rebuilds with a larger table and runs again
    if constexpr (D1B > 0) { if (a.counts[1][1]) return; }
    constexpr int Mp = 4 * G, cpw = 32 / G;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int sub = lane / G, g = lane - sub * G, gbase = sub * G;
    const bool lane_on = sub < cpw;
    const int tile = blockIdx.x, TP = tm.TP, N = tm.N;
    const TileOrigin org = tile_origin(tm, tile);
    constexpr bool do_slice = MODE & 1, do_splat = MODE & 2, store_q = MODE & 4;
    const int c0 = 4 * g;
    // The tile's splat inputs - segment metadata (start | length, vertex) and the pair lists - go to shared memory with
    // TMA bulk copies (cp.async.bulk, completion on an mbarrier): ONE thread issues four copies, they run during
    // phase 1, and phase 2 never waits for L2.  The metadata region holds 2 * TP segments shared by the lattices;
    // segments beyond the staged ones (never seen in practice) are read from global memory.
    __shared__ alignas(8) unsigned long long stage_bar;
    int2* metaA = reinterpret_cast<int2*>(qtile + (size_t)TP * G);
    int2* metaB = metaA;
    uint2* spairsA = reinterpret_cast<uint2*>(metaA + 2 * TP);
    uint2* spairsB = spairsA + TP * D1A;
    int neA = 0, neB = 0, capA = 0, capB = 0;
    if (do_splat) {
        neA = __ldg(a.tile_nent[0] + tile);
        capA = min(neA, 2 * TP);
        if constexpr (D1B > 0) {
            neB = __ldg(a.tile_nent[1] + tile);
            metaB = metaA + ((capA + 1) & ~1);  // keep 16-byte alignment for the bulk copy
            capB = min(neB, 2 * TP - ((capA + 1) & ~1));
        }
        const unsigned bar = (unsigned)__cvta_generic_to_shared(&stage_bar);
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            auto bulk = [&](void* dst, const void* src, unsigned bytes) {
                if (bytes == 0) return;
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 (unsigned)__cvta_generic_to_shared(dst)),
                             "l"(src), "r"(bytes), "r"(bar)
                             : "memory");
            };
            // sizes rounded up to 16 bytes: the arrays have TP * D1 (even) slots per tile, so the extra 8 bytes exist
            const unsigned szMA = ((unsigned)capA * 8 + 15) & ~15u, szMB = ((unsigned)capB * 8 + 15) & ~15u;
            const unsigned szPA = (unsigned)TP * D1A * 8, szPB = (unsigned)TP * D1B * 8;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(szMA + szMB + szPA + szPB) : "memory");
            bulk(metaA, a.ent_meta[0] + (size_t)tile * TP * D1A, szMA);
            bulk(spairsA, a.pairs[0] + (size_t)tile * TP * D1A, szPA);
            if constexpr (D1B > 0) {
                bulk(metaB, a.ent_meta[1] + (size_t)tile * TP * D1B, szMB);
                bulk(spairsB, a.pairs[1] + (size_t)tile * TP * D1B, szPB);
            }
        }
    }
    // per channel group (host-precomputed, FusedLayers): which of my four channels belong to which layer (one nibble per
    // layer), and for aligned layers my layer, my valid channels and which lanes of the group share the layer
    const unsigned lmask = ls.group_lmask[g];
    const unsigned vm = ls.group_valid[g];
    const int my_l = ls.group_layer[g];
    // bit i: lane i of my group holds channels of MY layer (a pad-only lane keeps itself as its only peer so that its
    // discarded result stays finite)
    unsigned peer_mask = 0;
#pragma unroll
    for (int i = 0; i < G; i++)
        if (my_l >= 0 ? ls.group_layer[i] == my_l : i == g) peer_mask |= 1u << i;

    // one warp-step = 32 / G points per warp: load the streaming inputs, then gather / soft-max / store
    struct StepIn {
        float4 u;
        PointIn<D1A> A;
        PointIn<D1B> B;
        int p;
        bool valid;
    };
    auto load_step = [&](int s, StepIn& in) {
        const int lp = (s * 8 + wib) * cpw + sub;
        const int p = lane_on ? tile_point(tm, org, lp) : -1;
        in.valid = p >= 0;
        in.p = p;
        in.u = make_float4(0.f, 0.f, 0.f, 0.f);
        if (in.valid) {
            in.u = __ldg(reinterpret_cast<const float4*>(unary + (size_t)p * Mp) + g);
            if (do_slice) { in.A.load(a.lat[0], p); in.B.load(a.lat[1], p); }
        }
    };
    auto run_step = [&](int s, const StepIn& in) {
        const int lp = (s * 8 + wib) * cpw + sub;
        const int p = in.p;
        const bool valid = in.valid;
        float4 t = make_float4(-in.u.x, -in.u.y, -in.u.z, -in.u.w);
        if (valid && do_slice) {
            slice_lattice<D1A>(a.lat[0], in.A, Mp, g, t);
            if constexpr (D1B > 0) slice_lattice<D1B>(a.lat[1], in.B, Mp, g, t);
        }
        // expAndNormalize per label layer across the G lanes of the point (densecrf.cpp:98-106)
        const float tv[4] = {t.x, t.y, t.z, t.w};
        float qv[4] = {0.f, 0.f, 0.f, 0.f};
        if (ls.aligned) {
            // every lane's four channels lie in ONE layer: a single pass, the lanes of the other layers are masked out
            // of the all-gathers by a -inf bias (max) and a 0/1 weight (sum); same summation order on every lane
            float mx = -INFINITY;
#pragma unroll
            for (int k = 0; k < 4; k++) mx = fmaxf(mx, (vm >> k) & 1u ? tv[k] : -INFINITY);
            float m = -INFINITY;
#pragma unroll
            for (int i = 0; i < G; i++) {
                const float o = __shfl_sync(0xffffffffu, mx, (gbase + i) & 31);
                m = fmaxf(m, (peer_mask >> i) & 1u ? o : -INFINITY);
            }
            float e[4], sl = 0.f;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                e[k] = (vm >> k) & 1u ? fast_exp(tv[k] - m) : 0.f;
                sl += e[k];
            }
            float sum = 0.f;
#pragma unroll
            for (int i = 0; i < G; i++) {
                const float o = __shfl_sync(0xffffffffu, sl, (gbase + i) & 31);
                sum += (peer_mask >> i) & 1u ? o : 0.f;
            }
            const float rs = __fdividef(1.0f, sum);
#pragma unroll
            for (int k = 0; k < 4; k++) qv[k] = e[k] * rs;
        } else {
            for (int l = 0; l < ls.n_layers; l++) {
                const unsigned m4 = (lmask >> (4 * l)) & 15u;
                float mx = -INFINITY;
#pragma unroll
                for (int k = 0; k < 4; k++) mx = fmaxf(mx, (m4 >> k) & 1u ? tv[k] : -INFINITY);
                const float m = group_gather_max<G>(mx, gbase);
                float e[4], sl = 0.f;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    e[k] = (m4 >> k) & 1u ? fast_exp(tv[k] - m) : 0.f;
                    sl += e[k];
                }
                const float rs = __fdividef(1.0f, group_gather_sum<G>(sl, gbase));
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if ((m4 >> k) & 1u) qv[k] = e[k] * rs;
            }
        }
        const float4 q = make_float4(qv[0], qv[1], qv[2], qv[3]);
        if (valid) {
            qtile[lp * G + g] = q;
            if (store_q) reinterpret_cast<float4*>(Q + (size_t)p * Mp)[g] = q;
        }
        if (labels) {
            // gated argmax (segmenter.cpp:645-657) / plain argmax (densecrf.cpp:200-208); ties -> lower label
            for (int l = 0; l < ls.n_layers; l++) {
                const unsigned m4 = (lmask >> (4 * l)) & 15u;
                const float gate = ls.gate[l];
                float bv = gate;
                int best = 1 << 20;
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if (((m4 >> k) & 1u) && qv[k] > bv) { bv = qv[k]; best = c0 + k - ls.off[l]; }
                float fv = gate;
                int fb = 1 << 20;
#pragma unroll
                for (int i = 0; i < G; i++) {  // lanes in channel order: strict '>' keeps the first maximum
                    const float ov = __shfl_sync(0xffffffffu, bv, (gbase + i) & 31);
                    const int ob = __shfl_sync(0xffffffffu, best, (gbase + i) & 31);
                    if (ob < (1 << 20) && ov > fv) { fv = ov; fb = ob; }
                }
                if (valid && g == 0)
                    labels[(size_t)l * N + p] = (uint8_t)(fb < (1 << 20) ? fb : (ls.unknown[l] >= 0 ? ls.unknown[l] : 0));
            }
        }
    };
    // (a register double-buffer that requests step s + 1 before step s gathers its rows was measured slower: it costs
    // half of the resident warps, and phase 1 is issue-bound, not latency-bound)
    for (int s = 0; s < steps; s++) {
        StepIn in0;
        load_step(s, in0);
        run_step(s, in0);
    }
    if (!do_splat) return;
    __syncthreads();  // the tile's marginals are complete (and thread 0's mbarrier.init is visible)
    {                 // wait for the bulk copies (phase 0 of the barrier)
        const unsigned bar = (unsigned)__cvta_generic_to_shared(&stage_bar);
        unsigned done = 0;
        while (!done)
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}"
                         : "=r"(done) : "r"(bar) : "memory");
    }
    // phase 2: tile-local gather splat out of shared memory
    {
        const size_t tbA = (size_t)tile * TP * D1A;
        gather_entries<G, RSS_TILE_IU>(spairsA, metaA, capA, a.ent_meta[0] + tbA, neA, a.lat[0].vout, qtile);
    }
    if constexpr (D1B > 0) {
        const size_t tbB = (size_t)tile * TP * D1B;
        gather_entries<G, RSS_TILE_IU>(spairsB, metaB, capB, a.ent_meta[1] + tbB, neB, a.lat[1].vout, qtile);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Blur of every lattice of the CRF in ONE cooperative launch (permutohedral.cpp:555-569).  Phase j blurs axis j of
// each lattice that has one; phases are separated by a grid barrier on an L2 counter (lattice.cuh).  Prefetching the
// neighbour pairs ahead of the rows was measured and does not help: the rows' L2 round trip dominates a round.
// Phase 0 additionally clears `zero` (the table the
// point kernel sliced from one iteration ago), which becomes the next splat target - so no phase follows the last axis.
// ---------------------------------------------------------------------------------------------------------------
// one axis of one lattice: items are (vertex, channel group); U independent items per thread are in flight at once
template <int U>
__device__ __forceinline__ void blur_axis(const float4* __restrict__ src, float4* __restrict__ dst, const int2* __restrict__ nb_j,
                                          uint32_t items, int G, uint32_t tid, uint32_t nthr) {
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
            if (it < items) {
                o[u] = __ldcg(src + it);
                x[u] = __ldcg(src + (size_t)nb[u].x * G + gq[u]);
                y[u] = __ldcg(src + (size_t)nb[u].y * G + gq[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint32_t it = base + u * nthr;
            if (it < items) __stcg(dst + it, blur_item(o[u], x[u], y[u]));
        }
    }
}
__global__ void __launch_bounds__(RSS_BLUR_MAXT) blur_multi_coop_kernel(const __grid_constant__ BlurMultiArgs a, int G,
                                                              unsigned int* barrier, unsigned int barrier_base) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, nthr = gridDim.x * blockDim.x;
    uint32_t V[FUSED_MAX_LAT];
    int maxd1 = 0;
    for (int k = 0; k < a.K; k++) {
        V[k] = a.counts[k][1] ? 0u : a.counts[k][0];
        maxd1 = max(maxd1, a.d1[k]);
    }
    for (int j = 0; j < maxd1; j++) {
        for (int k = 0; k < a.K; k++) {
            if (j >= a.d1[k]) continue;
            const float4* src = (j & 1) ? a.pong[k] : a.ping[k];
            float4* dst = (j & 1) ? a.ping[k] : a.pong[k];
            const uint32_t items = V[k] * (uint32_t)G;
            blur_axis<RSS_BLUR_U>(src, dst, a.nbr[k] + (size_t)j * a.vcap[k], items, G, tid, nthr);
            if (j == 0 && a.zero[k])
                for (uint32_t it = tid; it < items; it += nthr) __stcg(a.zero[k] + it, make_float4(0.f, 0.f, 0.f, 0.f));
        }
        if (j + 1 < maxd1) grid_barrier(barrier, barrier_base + (unsigned int)(j + 1) * gridDim.x);
    }
}

// splat of the all-ones vector for the normalisation (pairwise.cpp:44): values[v][0] += sum of barycentric weights.
// Same run-accumulation as the point kernel, one thread per chunk, rows of 4 floats with channel 0 live.
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
static int fused_tile_steps(int G, int TP) {  // warp-steps a CTA needs for TP points: 8 warps x (32 / G) points per step
    const int per_step = 8 * (32 / G);
    return (TP + per_step - 1) / per_step;
}
// The point kernel is latency-bound per CTA, so ONE full wave of resident CTAs is the sweet spot: when the default
// tile size would need slightly more tiles than fit at once (sm_count * RSS_TILE_MINB), the tiles grow (up to
// TILE_MAX_POINTS) until they fit.
#ifndef RSS_TILE_MAX_POINTS
#define RSS_TILE_MAX_POINTS 576
#endif
constexpr int TILE_MAX_POINTS = RSS_TILE_MAX_POINTS;
TileMap fused_tile_map(int G, int N, int W, int H, int sm_count) {
    TileMap m;
    m.N = N;
    const int slots = std::max(1, sm_count * RSS_TILE_MINB);
    if (W > 0 && H > 0 && (long long)W * H == N) {
        m.W = W; m.H = H; m.TW = 32; m.TH = RSS_TILE_POINTS / 32;
        m.tiles_x = (W + m.TW - 1) / m.TW;
        if (RSS_TILE_SINGLE_WAVE && m.tiles_x <= slots) {
            const int rows_fit = slots / m.tiles_x;                 // tile rows of one wave
            const int th = (H + rows_fit - 1) / rows_fit;           // tile height that makes the image fit in one wave
            if (th > m.TH && th * m.TW <= TILE_MAX_POINTS) m.TH = th;
        }
        m.TP = m.TW * m.TH;
        m.ntiles = m.tiles_x * ((H + m.TH - 1) / m.TH);
    } else {
        const int per_step = 8 * (32 / G);
        m.W = m.H = 0; m.TW = m.TH = 0; m.tiles_x = 0;
        m.TP = per_step * std::max(1, (RSS_TILE_POINTS + per_step / 2) / per_step);
        const long long fit = ((long long)N + slots - 1) / slots;  // points per tile for one wave
        const int tp_fit = (int)((fit + per_step - 1) / per_step) * per_step;
        if (RSS_TILE_SINGLE_WAVE && tp_fit > m.TP && tp_fit <= TILE_MAX_POINTS) m.TP = tp_fit;
        m.ntiles = (int)(((long long)N + m.TP - 1) / m.TP);
    }
    return m;
}

bool fused_group_supported(int G) { return G == 1 || G == 2 || G == 3 || G == 5 || G == 6; }
bool fused_signature_supported(int G, int d1a, int d1b) {
    if (!fused_group_supported(G)) return false;
    switch (d1a * 16 + d1b) {
        case 0x46: case 0x36: case 0x70: case 0x60: case 0x40: case 0x30: return true;
        default: return false;
    }
}

template <int G>
static void launch_tile_g(rss_ctx* c, cudaStream_t st, const FusedArgs& a, int d1a, int d1b, const float* unary, float* Q,
                          uint8_t* labels, const TileMap& tm, const FusedLayers& ls, int mode) {
    const int TP = tm.TP, steps = fused_tile_steps(G, TP);
    const int grid = tm.ntiles;
#define RSS_TILE_M(A, B, M)                                                                                             \
    do {                                                                                                                \
        auto kfn = meanfield_tile_kernel<G, A, B, M>;                                                                   \
        const size_t smem = (size_t)TP * G * sizeof(float4) + (size_t)2 * TP * sizeof(int2) +                           \
                            (size_t)TP * (A + B) * sizeof(uint2);                                                       \
        if (c->smem_attr_done.insert((const void*)kfn).second) { /* once per context (= per device) and instantiation */ \
            cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);                        \
            cudaFuncSetAttribute(kfn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);  \
        }                                                                                                               \
        RSS_LAUNCH_NAMED(c, "meanfield_tile_kernel", kfn, grid, 256, smem, st, a, unary, Q, labels, tm, steps, ls);     \
    } while (0)
#define RSS_TILE(A, B)                                                                                                  \
    do {                                                                                                                \
        if (mode == 2) RSS_TILE_M(A, B, 2);                                                                             \
        else if (mode == 3) RSS_TILE_M(A, B, 3);                                                                        \
        else RSS_TILE_M(A, B, 5);                                                                                       \
    } while (0)
    switch (d1a * 16 + d1b) {
        case 0x46: RSS_TILE(4, 6); break;
        case 0x36: RSS_TILE(3, 6); break;
        case 0x70: RSS_TILE(7, 0); break;
        case 0x60: RSS_TILE(6, 0); break;
        case 0x40: RSS_TILE(4, 0); break;
        case 0x30: RSS_TILE(3, 0); break;
        default: break;
    }
#undef RSS_TILE
#undef RSS_TILE_M
}
void launch_meanfield_fused(rss_ctx* c, cudaStream_t st, const FusedArgs& a, int d1a, int d1b, const float* unary, float* Q,
                            uint8_t* labels, const TileMap& tm, int G, const FusedLayers& ls, int mode) {
    switch (G) {
        case 1: launch_tile_g<1>(c, st, a, d1a, d1b, unary, Q, labels, tm, ls, mode); break;
        case 2: launch_tile_g<2>(c, st, a, d1a, d1b, unary, Q, labels, tm, ls, mode); break;
        case 3: launch_tile_g<3>(c, st, a, d1a, d1b, unary, Q, labels, tm, ls, mode); break;
        case 5: launch_tile_g<5>(c, st, a, d1a, d1b, unary, Q, labels, tm, ls, mode); break;
        case 6: launch_tile_g<6>(c, st, a, d1a, d1b, unary, Q, labels, tm, ls, mode); break;
        default: break;
    }
}

void launch_tile_csr_build(rss_ctx* c, cudaStream_t st, const int* offsets, const float* bary, const float* norm,
                           const TileMap& tm, int d1, int row_bytes, const uint32_t* counts, uint2* pairs, int2* ent_meta,
                           int* tile_nent) {
    const int grid = tm.ntiles, TP = tm.TP;
    int HC = 1024;  // power of two with load factor <= 0.8 even when every pair of the tile hits a different vertex
    while (HC * 4 < TP * d1 * 5) HC *= 2;
    const size_t tsm = (size_t)2 * HC * 4 + (size_t)TP * d1 * 2;
#define RSS_TCB(D)                                                                                                      \
    do {                                                                                                                \
        if (c->smem_attr_done.insert((const void*)tile_csr_build_kernel<D>).second)  /* once per context (= per device) */ \
            cudaFuncSetAttribute(tile_csr_build_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);     \
        RSS_LAUNCH(c, tile_csr_build_kernel<D>, grid, 256, tsm, st, offsets, bary, norm, tm, row_bytes, HC, counts, pairs,    \
                   ent_meta, tile_nent);                                                                                \
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

int blur_multi_grid(const rss_ctx* c) { return c->sm_count; }  // one CTA per SM
void launch_blur_multi(rss_ctx* c, cudaStream_t st, BlurMultiArgs a, int G, unsigned int* barrier, unsigned int barrier_base) {
    const int grid = blur_multi_grid(c), block = RSS_BLUR_MAXT;
    void* args[] = {&a, &G, &barrier, &barrier_base};
    cudaEvent_t ea = nullptr, eb = nullptr;
    if (c->profile) { ea = c->prof_event(); eb = c->prof_event(); cudaEventRecord(ea, st); }
    cudaLaunchCooperativeKernel((const void*)blur_multi_coop_kernel, dim3(grid), dim3(block), args, 0, st);
    c->launches++;
    if (c->profile) { cudaEventRecord(eb, st); c->prof_pending.push_back(rss_ctx::Pending{"blur_multi_coop_kernel", ea, eb}); }
}

void launch_splat_ones_runs(rss_ctx* c, cudaStream_t st, const int* offsets, const float* bary, int N, int d1,
                            const uint32_t* counts, float* values) {
    const int Pc = 16;
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
