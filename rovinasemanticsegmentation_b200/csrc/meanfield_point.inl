// The point kernel of the fused mean-field iteration and its launch dispatch (see meanfield.cu for the design).  This file is
// compiled TWICE, into two namespaces with two register budgets:
//   point_alone   (meanfield.cu)         __launch_bounds__(256, RSS_TILE_MINB = 3): up to 85 registers, the fastest single launch;
//   point_shared  (meanfield_shared.cu)  __maxnreg__(64): three point CTAs (3 x 256 x 64 registers) and one 256-thread CTA of
//                 the cooperative blur (256 x 64) fill the register file of an SM exactly.  When several keyframes are in
//                 flight on one GPU a blur is resident most of the time, and the 80-register build then runs with two
//                 instead of three CTAs per SM: measured 1072 -> 1100 keyframes/s with three keyframes in flight, although
//                 the kernel alone is 4 % slower (50.0 vs 47.8 us).
// launch_meanfield_fused (meanfield.cu) picks the variant by the number of live contexts on the device.
// Required before inclusion: RSS_POINT_NS, RSS_POINT_BOUNDS, RSS_POINT_ENTRY (name of the exported launch function).
namespace rss {
namespace RSS_POINT_NS {

__device__ __forceinline__ void red_add_v4(float* dst, const float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// ---------------------------------------------------------------------------------------------------------------
// The point kernel.  `MODE` bit 0: slice (not the first pass), bit 1: splat (not the last pass), bit 2: store Q.
// ---------------------------------------------------------------------------------------------------------------
// exp(x) for x <= 0 as one multiply and one ex2.approx.ftz (|rel err| < 2e-6; the CRF tolerance is 1e-4 abs; results
// below 2^-126 flush to zero, which the normalisation cannot tell from the true value)
__device__ __forceinline__ float fast_exp(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x * 1.4426950408889634f));
    return y;
}
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    unsigned done = 0;
    while (!done)
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     (unsigned)__cvta_generic_to_shared(dst)),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

// a point's D1 slice weights and row slots, fetched with vector loads (pt_index order) BEFORE the staging barrier is
// waited on, so that their latency overlaps the TMA / cp.async traffic
template <int D1>
struct PointIn {
    float w[D1 > 0 ? D1 : 1];
    int s[D1 > 0 ? D1 : 1];
    __device__ __forceinline__ void load(const FusedLat& L, size_t tb, int lp) {
        if constexpr (D1 > 0) {
            constexpr int TP = TILE_POINTS, n4 = D1 / 4, n2 = (D1 % 4) / 2, n1 = D1 % 2;
            const float* wp = L.pt_w + tb;
            const uint16_t* sp = L.pt_slot + tb;
#pragma unroll
            for (int k = 0; k < n4; k++) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(wp) + k * TP + lp);
                const uint2 u = __ldg(reinterpret_cast<const uint2*>(sp) + k * TP + lp);
                w[4 * k] = v.x; w[4 * k + 1] = v.y; w[4 * k + 2] = v.z; w[4 * k + 3] = v.w;
                s[4 * k] = u.x & 0xffff; s[4 * k + 1] = u.x >> 16; s[4 * k + 2] = u.y & 0xffff; s[4 * k + 3] = u.y >> 16;
            }
            if constexpr (n2 > 0) {
                const float2 v = __ldg(reinterpret_cast<const float2*>(wp + 4 * n4 * TP) + lp);
                const unsigned u = __ldg(reinterpret_cast<const unsigned*>(sp + 4 * n4 * TP) + lp);
                w[4 * n4] = v.x; w[4 * n4 + 1] = v.y;
                s[4 * n4] = u & 0xffff; s[4 * n4 + 1] = u >> 16;
            }
            if constexpr (n1 > 0) {
                w[D1 - 1] = __ldg(wp + (4 * n4 + 2 * n2) * TP + lp);
                s[D1 - 1] = __ldg(sp + (4 * n4 + 2 * n2) * TP + lp);
            }
        }
    }
};
// t[c] += w * row[c] for the D1 corners of one lattice; rows of slots < TILE_ROW_CAP are in shared memory
template <int G, int D1>
__device__ __forceinline__ void slice_lattice(const FusedLat& L, size_t tb, const PointIn<D1>& in, const float4* rows,
                                              float (&t)[4 * G]) {
    constexpr int MP = 4 * G;
#pragma unroll
    for (int j = 0; j < D1; j++) {
        const float w = in.w[j];
        if (in.s[j] < TILE_ROW_CAP) {
            const float4* row = rows + in.s[j] * G;
#pragma unroll
            for (int g = 0; g < G; g++) {
                const float4 v = row[g];
                t[4 * g] = fmaf(w, v.x, t[4 * g]); t[4 * g + 1] = fmaf(w, v.y, t[4 * g + 1]);
                t[4 * g + 2] = fmaf(w, v.z, t[4 * g + 2]); t[4 * g + 3] = fmaf(w, v.w, t[4 * g + 3]);
            }
        } else {  // more distinct vertices in this tile than staged rows (rare): straight from the value table
            const float4* row = reinterpret_cast<const float4*>(L.vin + (size_t)__ldg(L.tile_vert + tb + in.s[j]) * MP);
#pragma unroll
            for (int g = 0; g < G; g++) {
                const float4 v = __ldg(row + g);
                t[4 * g] = fmaf(w, v.x, t[4 * g]); t[4 * g + 1] = fmaf(w, v.y, t[4 * g + 1]);
                t[4 * g + 2] = fmaf(w, v.z, t[4 * g + 2]); t[4 * g + 3] = fmaf(w, v.w, t[4 * g + 3]);
            }
        }
    }
}
// the tile's distinct value rows -> shared memory, 16 bytes per cp.async, all threads
template <int G>
__device__ __forceinline__ void stage_rows(const FusedLat& L, size_t tb, int nrows, float4* rows) {
    constexpr int MP = 4 * G;
    for (int i = threadIdx.x; i < nrows * G; i += TILE_POINTS) {
        const int r = i / G, g = i - r * G;
        const float* src = L.vin + (size_t)__ldg(L.tile_vert + tb + r) * MP + 4 * g;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(rows + i)), "l"(src)
                     : "memory");
    }
}

// one thread = one splat segment, all channels in registers; pairs[].x = byte offset of the point's row in the Q tile
// A warp walks a GROUP of 32 consecutive segments, lane <-> segment; the group's pairs are stored column-major
// (tile_csr_build_kernel): at step k the lanes that still have a pair - a prefix of the group, the segments are ordered
// longest first - read consecutive 8-byte pairs.  The column starts follow from ballots, no per-column metadata.
// REV: the lanes AND the warps take the segments in reverse order (thread TP-1 the first = longest one).  When the second
// lattice of a tile is walked in reverse, the threads that had the long segments of the first lattice get the short ones
// of the second: the serial chain per thread is ~(longest + shortest) instead of 2 x longest.  The bank-conflict ordering
// of the pairs only needs the eight lanes of a quarter-warp to hold eight consecutive segments, in either direction.
template <int G, bool REV>
__device__ __forceinline__ void gather_entries(const uint2* pr, const int2* meta, int cap, const int2* __restrict__ meta_g,
                                               int ne, float* __restrict__ vout, const float4* qtile) {
    constexpr int MP = 4 * G;
    const char* qbytes = reinterpret_cast<const char*>(qtile);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int li = REV ? 31 - lane : lane;  // position of this lane's segment in its group
    for (int e0 = 32 * (REV ? TILE_POINTS / 32 - 1 - w : w); e0 < ne; e0 += TILE_POINTS) {
        const int e = e0 + li;
        int2 m = make_int2(0, 0);
        if (e < ne) m = e < cap ? meta[e] : __ldg(meta_g + e);
        const int len = m.x >> 16;
        const int2 m0 = make_int2(__shfl_sync(0xffffffffu, m.x, REV ? 31 : 0), 0);  // the group's first (longest) segment
        const int len0 = m0.x >> 16;
        const uint2* col = pr + (m0.x & 0xffff) + li;
        float4 acc[G];
#pragma unroll
        for (int g = 0; g < G; g++) acc[g] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int k = 0; k < len0; k += 2) {  // two pairs per trip: the second pair's loads overlap the first pair's multiply-adds
            const bool a0 = k < len, a1 = k + 1 < len;
            const int n0 = __popc(__ballot_sync(0xffffffffu, a0)), n1 = __popc(__ballot_sync(0xffffffffu, a1));
            if (a0) {
                const uint2 p0 = col[0];
                const uint2 p1 = a1 ? col[n0] : make_uint2(p0.x, 0u);  // no second pair: weight 0 on the same row
                const float w0 = __uint_as_float(p0.y), w1 = __uint_as_float(p1.y);
                const float4* q0 = reinterpret_cast<const float4*>(qbytes + p0.x);
                const float4* q1 = reinterpret_cast<const float4*>(qbytes + p1.x);
#pragma unroll
                for (int g = 0; g < G; g++) {
                    const float4 a = q0[g], b = q1[g];
                    acc[g].x = fmaf(w0, a.x, acc[g].x); acc[g].y = fmaf(w0, a.y, acc[g].y);
                    acc[g].z = fmaf(w0, a.z, acc[g].z); acc[g].w = fmaf(w0, a.w, acc[g].w);
                    acc[g].x = fmaf(w1, b.x, acc[g].x); acc[g].y = fmaf(w1, b.y, acc[g].y);
                    acc[g].z = fmaf(w1, b.z, acc[g].z); acc[g].w = fmaf(w1, b.w, acc[g].w);
                }
            }
            col += n0 + n1;
        }
        if (len > 0) {
            float* dst = vout + (size_t)m.y * MP;
#pragma unroll
            for (int g = 0; g < G; g++) red_add_v4(dst + 4 * g, acc[g]);
        }
    }
}

template <int G, int D1A, int D1B, int MODE>
__global__ void RSS_POINT_BOUNDS
    meanfield_point_kernel(const __grid_constant__ FusedArgs a, const float* __restrict__ unary, float* __restrict__ Q,
                           uint8_t* __restrict__ labels, const __grid_constant__ TileMap tm,
                           const __grid_constant__ FusedLayers ls) {
    constexpr int MP = 4 * G, TP = TILE_POINTS, RC = TILE_ROW_CAP;
    constexpr bool do_slice = MODE & 1, do_splat = MODE & 2, store_q = MODE & 4;
    extern __shared__ float4 smem_f4[];
    if (a.lat[0].counts[1]) return;  // lattice overflow: the host rebuilds with a larger table and runs again
    if constexpr (D1B > 0) { if (a.lat[1].counts[1]) return; }
    float4* qtile = smem_f4;            // [TP][G]  unary rows on arrival, marginals after phase 1
    float4* rowsA = qtile + TP * G;     // [RC][G]  staged value rows of lattice A ...
    float4* rowsB = rowsA + RC * G;     //          ... and B
#if RSS_POINT_ALIAS
    // the splat lists share the shared memory of the staged value rows (phase 2 follows phase 1): ~46 instead of ~70 KB per
    // CTA, i.e. 4 instead of 3 CTAs per SM; their bulk copy is then issued after phase 1, not at the start
    constexpr bool alias_lists = do_slice && do_splat;
    int2* metaA = reinterpret_cast<int2*>(alias_lists ? rowsA : rowsB + (D1B > 0 ? RC * G : 0));
#else
    constexpr bool alias_lists = false;
    int2* metaA = reinterpret_cast<int2*>(rowsB + (D1B > 0 ? RC * G : 0));  // 2 * TP segment slots shared by the lattices
#endif
    int2* metaB = metaA;
    uint2* spairsA = reinterpret_cast<uint2*>(metaA + 2 * TP);
    uint2* spairsB = spairsA + TP * D1A;
    __shared__ alignas(8) unsigned long long stage_bar[2];  // [0] unary + value rows (phase 1), [1] splat lists (phase 2)
    const unsigned bar0 = (unsigned)__cvta_generic_to_shared(&stage_bar[0]);
    const unsigned bar1 = (unsigned)__cvta_generic_to_shared(&stage_bar[1]);
    const int tile = blockIdx.x, N = tm.N, lp = threadIdx.x;
    const TileOrigin org = tile_origin(tm, tile);
    const size_t tbA = (size_t)tile * TP * D1A, tbB = (size_t)tile * TP * D1B;
    const int2 infoA = __ldg(a.lat[0].tile_info + tile);
    int2 infoB = make_int2(0, 0);
    if constexpr (D1B > 0) infoB = __ldg(a.lat[1].tile_info + tile);
    const int nrA = min(infoA.y, RC), nrB = min(infoB.y, RC);
    const int neA = infoA.x, neB = infoB.x;
    const int capA = min(neA, 2 * TP);
    int capB = 0;
    if constexpr (D1B > 0) {
        metaB = metaA + ((capA + 1) & ~1);  // keep 16-byte alignment for the bulk copy
        capB = min(neB, 2 * TP - ((capA + 1) & ~1));
    }
    if (threadIdx.x == 0) {
        mbar_init(bar0, 1 + (do_slice ? TP : 0));  // thread 0's expect_tx + (slice passes) every thread's cp.async arrival
        mbar_init(bar1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();  // the barriers exist before anybody's copy can complete on them
    // ---- staging: everything below is issued up front and lands while the threads fetch their own inputs
    if (threadIdx.x == 0) {
        // unary rows of the tile -> Q tile
        unsigned ubytes = 0;
        if (tm.W == 0) {
            const long long n = min((long long)TP, (long long)N - org.base);
            ubytes = (unsigned)n * MP * 4;
            mbar_expect_tx(bar0, ubytes);
            bulk_g2s(qtile, unary + (size_t)org.base * MP, ubytes, bar0);
        } else {
            const int w = min(tm.TW, tm.W - org.x0), h = min(tm.TH, tm.H - org.y0);
            ubytes = (unsigned)(w * h) * MP * 4;
            mbar_expect_tx(bar0, ubytes);
            for (int ly = 0; ly < h; ly++)
                bulk_g2s(qtile + (size_t)ly * tm.TW * G, unary + ((size_t)(org.y0 + ly) * tm.W + org.x0) * MP,
                         (unsigned)w * MP * 4, bar0);
        }
        if (do_splat && !alias_lists) {
            // sizes rounded up to 16 bytes: the arrays have TP * D1 (even) slots per tile, so the extra 8 bytes exist
            const unsigned szMA = ((unsigned)capA * 8 + 15) & ~15u, szMB = ((unsigned)capB * 8 + 15) & ~15u;
            const unsigned szPA = (unsigned)TP * D1A * 8, szPB = (unsigned)TP * D1B * 8;
            mbar_expect_tx(bar1, szMA + szMB + szPA + szPB);
            if (szMA) bulk_g2s(metaA, a.lat[0].ent_meta + tbA, szMA, bar1);
            bulk_g2s(spairsA, a.lat[0].pairs + tbA, szPA, bar1);
            if constexpr (D1B > 0) {
                if (szMB) bulk_g2s(metaB, a.lat[1].ent_meta + tbB, szMB, bar1);
                bulk_g2s(spairsB, a.lat[1].pairs + tbB, szPB, bar1);
            }
        }
    }
    PointIn<D1A> inA;
    PointIn<D1B> inB;
    if constexpr (do_slice) {
        // the DISTINCT value rows the tile's points reference: ~100 rows per lattice instead of (d+1) * 256 gathers
        stage_rows<G>(a.lat[0], tbA, nrA, rowsA);
        if constexpr (D1B > 0) stage_rows<G>(a.lat[1], tbB, nrB, rowsB);
        // this thread's cp.asyncs arrive on bar0 when they have landed (the arrival is pre-counted in the init)
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar0) : "memory");
        inA.load(a.lat[0], tbA, lp);
        inB.load(a.lat[1], tbB, lp);
    }
    // ---- phase 1: one thread = one point
    const int p = tile_point(tm, org, lp);
    const unsigned live_lanes = __ballot_sync(0xffffffffu, p >= 0);  // the lanes that take part in phase 1
    mbar_wait(bar0, 0);
    if (p >= 0) {
        float t[MP];
#pragma unroll
        for (int g = 0; g < G; g++) {
            const float4 u = qtile[lp * G + g];
            t[4 * g] = -u.x; t[4 * g + 1] = -u.y; t[4 * g + 2] = -u.z; t[4 * g + 3] = -u.w;
        }
        if constexpr (do_slice) {
            // tmp = -unary - (-w * (alpha * filtered) * norm)   (densecrf.cpp:126, pairwise.cpp:78-79, permutohedral.cpp:571)
            slice_lattice<G, D1A>(a.lat[0], tbA, inA, rowsA, t);
            if constexpr (D1B > 0) slice_lattice<G, D1B>(a.lat[1], tbB, inB, rowsB, t);
        }
        // expAndNormalize per label layer (densecrf.cpp:98-106).  Layers are whole float4 groups and their padding
        // channels carry t = -inf (unary = +inf), so everything below is per GROUP: no per-channel predicates.
        constexpr float L2E = 1.4426950408889634f;
        float gm[G], gs[G];
#pragma unroll
        for (int g = 0; g < G; g++) gs[g] = 0.f;
#pragma unroll
        for (int g = 0; g < G; g++) gm[g] = fmaxf(fmaxf(t[4 * g], t[4 * g + 1]), fmaxf(t[4 * g + 2], t[4 * g + 3]));
        for (int l = 0; l < ls.n_layers; l++) {  // every group takes the maximum of its layer
            const unsigned m = ls.gmask[l];
            float mx = -INFINITY;
#pragma unroll
            for (int g = 0; g < G; g++)
                if ((m >> g) & 1u) mx = fmaxf(mx, gm[g]);
#pragma unroll
            for (int g = 0; g < G; g++)
                if ((m >> g) & 1u) gs[g] = mx;
        }
#pragma unroll
        for (int g = 0; g < G; g++) {
            const float nm = -gs[g] * L2E;  // exp(t - m) = ex2(t * log2(e) - m * log2(e)): one FFMA + one MUFU per channel
            float sum = 0.f;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                float e;
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fmaf(t[4 * g + k], L2E, nm)));
                t[4 * g + k] = e;
                sum += e;
            }
            gm[g] = sum;
        }
        for (int l = 0; l < ls.n_layers; l++) {
            const unsigned m = ls.gmask[l];
            float sum = 0.f;
#pragma unroll
            for (int g = 0; g < G; g++)
                if ((m >> g) & 1u) sum += gm[g];
            const float rs = __fdividef(1.0f, sum);
#pragma unroll
            for (int g = 0; g < G; g++)
                if ((m >> g) & 1u) gs[g] = rs;
        }
        // the marginals go to the ROTATED row (q_row): another lane's unary row, which that lane read at the top of phase 1 -
        // the rotation stays inside the warp's 32 rows, so a warp-level barrier orders its reads before these writes
        const int qr = q_row(lp);
        if (do_splat) __syncwarp(live_lanes);
#pragma unroll
        for (int g = 0; g < G; g++) {
            const float4 v = make_float4(t[4 * g] * gs[g], t[4 * g + 1] * gs[g], t[4 * g + 2] * gs[g], t[4 * g + 3] * gs[g]);
            t[4 * g] = v.x; t[4 * g + 1] = v.y; t[4 * g + 2] = v.z; t[4 * g + 3] = v.w;
            if (do_splat) qtile[qr * G + g] = v;
            if (store_q && Q) reinterpret_cast<float4*>(Q + (size_t)p * MP)[g] = v;
        }
        if (labels) {
            // gated argmax (segmenter.cpp:645-657) / plain argmax (densecrf.cpp:200-208); strict '>' keeps the first maximum
            for (int l = 0; l < ls.n_layers; l++) {
                const int ca = ls.off[l], cb = ca + ls.count[l];
                float bv = ls.gate[l];
                int best = ls.unknown[l] >= 0 ? ls.unknown[l] : 0;
#pragma unroll
                for (int c = 0; c < MP; c++)
                    if (c >= ca && c < cb && t[c] > bv) { bv = t[c]; best = c - ca; }
                labels[(size_t)l * N + p] = (uint8_t)best;
            }
        }
    }
    if (!do_splat) return;
    __syncthreads();  // the tile's marginals are complete (and nobody reads the staged value rows any more)
    if (alias_lists && threadIdx.x == 0) {
        // generic-proxy reads of the rows region are ordered before the async-proxy writes of the bulk copies
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            // sizes rounded up to 16 bytes: the arrays have TP * D1 (even) slots per tile, so the extra 8 bytes exist
            const unsigned szMA = ((unsigned)capA * 8 + 15) & ~15u, szMB = ((unsigned)capB * 8 + 15) & ~15u;
            const unsigned szPA = (unsigned)TP * D1A * 8, szPB = (unsigned)TP * D1B * 8;
            mbar_expect_tx(bar1, szMA + szMB + szPA + szPB);
            if (szMA) bulk_g2s(metaA, a.lat[0].ent_meta + tbA, szMA, bar1);
            bulk_g2s(spairsA, a.lat[0].pairs + tbA, szPA, bar1);
            if constexpr (D1B > 0) {
                if (szMB) bulk_g2s(metaB, a.lat[1].ent_meta + tbB, szMB, bar1);
                bulk_g2s(spairsB, a.lat[1].pairs + tbB, szPB, bar1);
            }
    }
    mbar_wait(bar1, 0);
    // ---- phase 2: tile-local gather splat out of shared memory
    gather_entries<G, false>(spairsA, metaA, capA, a.lat[0].ent_meta + tbA, neA, a.lat[0].vout, qtile);
    if constexpr (D1B > 0)
        gather_entries<G, RSS_SPLAT_REV != 0>(spairsB, metaB, capB, a.lat[1].ent_meta + tbB, neB, a.lat[1].vout, qtile);
}

template <int G, int A, int B, int M>
static cudaError_t launch_point_m(rss_ctx* c, cudaStream_t st, const FusedArgs& a, const float* unary, float* Q, uint8_t* labels,
                                  const TileMap& tm, const FusedLayers& ls) {
    auto kfn = meanfield_point_kernel<G, A, B, M>;
    const size_t rows_b = (size_t)TILE_ROW_CAP * G * sizeof(float4) * (B > 0 ? 2 : 1);
    const size_t lists_b = (size_t)2 * TILE_POINTS * sizeof(int2) + (size_t)TILE_POINTS * (A + B) * sizeof(uint2);
#if RSS_POINT_ALIAS
    const size_t smem = (size_t)TILE_POINTS * G * sizeof(float4) + ((M & 3) == 3 ? std::max(rows_b, lists_b) : rows_b + lists_b);
#else
    const size_t smem = (size_t)TILE_POINTS * G * sizeof(float4) + rows_b + lists_b;
#endif
    if (c->smem_attr_done.insert((const void*)kfn).second) { /* once per context (= per device) and instantiation */
        cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(kfn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
    }
    RSS_LAUNCH_NAMED(c, M == 3 ? "meanfield_point_kernel" : (M == 2 ? "meanfield_point_kernel<first>" : "meanfield_point_kernel<last>"),
                     kfn, tm.ntiles, TILE_POINTS, smem, st, a, unary, Q, labels, tm, ls);
    return cudaPeekAtLastError();
}
template <int G, int A, int B>
static cudaError_t launch_point_ab(rss_ctx* c, cudaStream_t st, const FusedArgs& a, const float* unary, float* Q, uint8_t* labels,
                                   const TileMap& tm, const FusedLayers& ls, int mode) {
    if (mode == 2) return launch_point_m<G, A, B, 2>(c, st, a, unary, Q, labels, tm, ls);
    if (mode == 3) return launch_point_m<G, A, B, 3>(c, st, a, unary, Q, labels, tm, ls);
    return launch_point_m<G, A, B, 5>(c, st, a, unary, Q, labels, tm, ls);
}
template <int G>
static cudaError_t launch_point_g(rss_ctx* c, cudaStream_t st, const FusedArgs& a, int d1a, int d1b, const float* unary, float* Q,
                                  uint8_t* labels, const TileMap& tm, const FusedLayers& ls, int mode) {
    switch (d1a * 16 + d1b) {
        case 0x46: return launch_point_ab<G, 4, 6>(c, st, a, unary, Q, labels, tm, ls, mode);
        case 0x36: return launch_point_ab<G, 3, 6>(c, st, a, unary, Q, labels, tm, ls, mode);
        case 0x70: return launch_point_ab<G, 7, 0>(c, st, a, unary, Q, labels, tm, ls, mode);
        case 0x60: return launch_point_ab<G, 6, 0>(c, st, a, unary, Q, labels, tm, ls, mode);
        case 0x40: return launch_point_ab<G, 4, 0>(c, st, a, unary, Q, labels, tm, ls, mode);
        case 0x30: return launch_point_ab<G, 3, 0>(c, st, a, unary, Q, labels, tm, ls, mode);
        default: return cudaErrorInvalidValue;
    }
}
cudaError_t RSS_POINT_ENTRY(rss_ctx* c, cudaStream_t st, const FusedArgs& a, int d1a, int d1b, const float* unary, float* Q,
                                   uint8_t* labels, const TileMap& tm, int G, const FusedLayers& ls, int mode) {
    switch (G) {
        case 1: return launch_point_g<1>(c, st, a, d1a, d1b, unary, Q, labels, tm, ls, mode);
        case 2: return launch_point_g<2>(c, st, a, d1a, d1b, unary, Q, labels, tm, ls, mode);
        case 3: return launch_point_g<3>(c, st, a, d1a, d1b, unary, Q, labels, tm, ls, mode);
        case 4: return launch_point_g<4>(c, st, a, d1a, d1b, unary, Q, labels, tm, ls, mode);
        case 5: return launch_point_g<5>(c, st, a, d1a, d1b, unary, Q, labels, tm, ls, mode);
        case 6: return launch_point_g<6>(c, st, a, d1a, d1b, unary, Q, labels, tm, ls, mode);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace RSS_POINT_NS
}  // namespace rss
