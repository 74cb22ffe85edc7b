// Device data structures of the permutohedral lattice (reference third-party/densecrf/src/permutohedral.cpp).
#pragma once
#include "common.cuh"

namespace rss {

constexpr int LAT_MAX_D = 7;   // feature dimensions; keys are d int16 packed into 128 bits (one spare lane)
constexpr int SPLAT_SEG = 32;  // nonzeros per splat work item (one warp)
constexpr int BLUR_COOP_MAX_ITEMS = 1 << 20;  // float4 items (vcap * Mp/4) up to which the one-launch blur is used

struct __align__(16) Key128 {
    unsigned long long lo, hi;
};

// One pairwise term = one lattice.  Vertex-count dependent sizes are allocated for vcap (= hash capacity / 2)
// and the true count lives on the device (counts[0]) so that no host synchronisation is needed to launch.
struct Lattice {
    int d = 0, N = 0;
    float potts_w = 0.f;
    int norm_type = RSS_NORMALIZE_SYMMETRIC;
    uint32_t hcap = 0;   // hash capacity (power of two)
    uint32_t vcap = 0;   // vertex capacity = hcap / 2
    uint32_t want_hcap = 0;  // capacity requested for the next build after an overflow
    int V_host = -1;     // vertex count once read back (diagnostics)
    DevBuf table;        // Key128[hcap]
    DevBuf slot_id;      // uint32[hcap]   slot -> vertex id
    DevBuf first_ref;    // uint32[hcap]   smallest (point, corner) pair index touching the slot
    DevBuf rank;         // uint32[N*(d+1)] first-appearance flags, then their exclusive scan
    DevBuf vkeys;        // Key128[vcap]   vertex id -> key
    DevBuf offsets;      // int  [N][d+1]  vertex id of each enclosing-simplex corner
    DevBuf bary;         // float[N][d+1]
    DevBuf nbr;          // int2 [d+1][vcap]  blur neighbours (n1, n2); missing -> zero row (index = vcap)
    DevBuf norm;         // float[N]
    DevBuf counts;       // uint32[16]: [0] V, [1] overflow flag, [2] segments, [4] seg total, [8] barrier
    DevBuf deg;          // uint32[vcap+1]  row degree, then row_start (in-place scan)
    DevBuf cursor;       // uint32[vcap]
    DevBuf nseg;         // uint32[vcap+1]  segments per row, then segment offsets
    DevBuf csr_pt;       // int  [N*(d+1)]
    DevBuf csr_w;        // float[N*(d+1)]
    DevBuf seg_v;        // int  [maxseg]   vertex of each splat segment
    DevBuf seg_begin;    // uint32[maxseg]
    DevBuf seg_end;      // uint32[maxseg]
    DevBuf val_a, val_b; // float[(vcap+1)][Mp] ping-pong value tables (row vcap stays zero)
    DevBuf val_c;        // third table of the fused mean-field path (meanfield.cu)
    DevBuf tile_pairs, tile_ent_meta, tile_vert, tile_info, tile_pt_w, tile_pt_slot;  // per-tile data of the fused point kernel (meanfield.cuh, FusedLat)
    int tile_TP = 0;     // points per tile the tile CSR was built for (0 = not built)
    int tile_W = 0;      // image width of the 2-D tiling the tile CSR was built for (0 = 1-D tiles)
    bool have_csr = false;  // vertex-major CSR + segments (generic splat path) built
    bool ordered = false;  // consecutive points share lattice vertices (raster order): the fused path applies
    long long runs = -1;   // number of (point, corner) pairs whose vertex differs from the previous point's (diagnostics)
    DevBuf scan_tmp;
    uint32_t maxseg = 0;
    int splat_target = 0;  // which value table is all-zero and receives the next splat (0 = val_a)
    unsigned int barrier_base = 0;  // grid-barrier arrivals issued so far (counts[8] is the barrier word)
    void release() {
        DevBuf* b[] = {&table, &slot_id, &first_ref, &rank, &vkeys, &offsets, &bary, &nbr, &norm, &counts, &deg, &cursor, &nseg,
                       &csr_pt, &csr_w, &seg_v, &seg_begin, &seg_end, &val_a, &val_b, &val_c, &scan_tmp, &tile_pairs, &tile_ent_meta, &tile_vert, &tile_info, &tile_pt_w, &tile_pt_slot};
        for (DevBuf* p : b) p->release();
    }
};

#ifdef __CUDACC__
// blur along one axis (permutohedral.cpp:555-569): new[v] = old[v] + 0.5 * (old[n1] + old[n2]), rounded like the reference
__device__ __forceinline__ float4 blur_item(const float4 o, const float4 a, const float4 b) {
    float4 r;
    r.x = __fadd_rn(o.x, __fmul_rn(0.5f, __fadd_rn(a.x, b.x)));
    r.y = __fadd_rn(o.y, __fmul_rn(0.5f, __fadd_rn(a.y, b.y)));
    r.z = __fadd_rn(o.z, __fmul_rn(0.5f, __fadd_rn(a.z, b.z)));
    r.w = __fadd_rn(o.w, __fmul_rn(0.5f, __fadd_rn(a.w, b.w)));
    return r;
}
// Grid barrier for cooperative launches (all CTAs co-resident): arrivals on one L2 counter with a monotonic target, so
// the counter never needs resetting.  1.2 us per barrier on B200 (tools/micro/barrier_bench.cu).
__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        unsigned int v;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
        } while ((int)(v - target) < 0);
    }
    __syncthreads();
}
#endif

}  // namespace rss

struct rss_crf {
    rss_ctx* ctx = nullptr;
    bool skip_q_store = false;     // fused path: the last point kernel writes only the label maps (keyframes without Q output)
    int N = 0, n_layers = 0;
    int M[RSS_MAX_LAYERS] = {0};
    // Device channel layout: layer l owns channels [moff[l], moff[l] + M[l]) of a point's row; every layer starts at a
    // multiple of 4 channels (float4 group) and its padding channels hold unary = +inf / Q = 0.  hoff[l] = sum of the
    // label counts of the layers before l: the layer's offset in host-side concatenations.
    int moff[RSS_MAX_LAYERS + 1] = {0};
    int hoff[RSS_MAX_LAYERS + 1] = {0};
    int Mtot = 0, Mp = 0;          // total labels; channels per row = sum over layers of M_l rounded up to 4
    rss::DevBuf unary;             // float[N][Mp]  energies
    rss::DevBuf Q;                 // float[N][Mp]
    rss::DevBuf scratch;           // float[N][Mp] staging for host-layout conversions / filter tests
    rss::DevBuf labels;            // uint8[n_layers][N]
    rss::DevBuf feat_stage;        // float[N][d] staging for feature upload
    rss::DevBuf cloud_xyz, cloud_rgb;  // float[N][3] each: the local map's points, resident (rss_crf_set_cloud)
    rss::DevBuf zbuf, index_dev;   // u64[H*W] z-buffer keys, int32[H*W] index image of the device projector
    bool have_cloud = false;
    int grid_w = 0, grid_h = 0;    // the points are the pixels of a grid_w x grid_h image in raster order (0 = unknown)
    std::vector<rss::Lattice*> kernels;
    std::vector<rss::Lattice*> pool;  // released lattices whose device buffers are reused by the next build
    cudaStream_t side[4] = {nullptr, nullptr, nullptr, nullptr};  // per-lattice streams
    cudaEvent_t ev_fork = nullptr, ev_join[4] = {nullptr, nullptr, nullptr, nullptr};
    bool unary_set = false;
    // Point sets whose own order is not coherent (local maps): points sorted by their first two lattice vertices; the fused
    // mean-field path runs over this order (meanfield.cuh, TileMap::perm).  Everything else stays in the caller's order.
    rss::DevBuf perm;              // int[N]   sorted position -> point
    rss::DevBuf unary_sorted;      // float[N][Mp] copy of the unary rows in sorted order (made at the start of an inference)
    rss::DevBuf sort_keys, sort_keys2, sort_vals, sort_tmp;  // scratch of the sort
    bool sorted = false;
};
