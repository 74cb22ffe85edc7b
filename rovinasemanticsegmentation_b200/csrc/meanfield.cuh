// Argument blocks of the fused mean-field kernels (meanfield.cu).
#pragma once
#include "common.cuh"

namespace rss {

constexpr int FUSED_MAX_LAT = 2;
constexpr int TILE_POINTS = 256;  // points per tile = threads per CTA of the point kernel (one thread per point)
constexpr int TILE_W = 32, TILE_H = 8;  // pixel block of an image tile
#ifndef RSS_TILE_ROW_CAP
#define RSS_TILE_ROW_CAP 160
#endif
constexpr int TILE_ROW_CAP = RSS_TILE_ROW_CAP;  // value rows per lattice a tile stages in shared memory (more: read from L2)

// How a tile's local point index maps to the point index.  W == 0: tile t owns the TP consecutive points starting at
// t * TP (generic point sets).  W > 0: the points are the pixels of a W x H image in raster order and a tile is a
// TW x TH pixel block - neighbouring pixels in BOTH directions share lattice vertices, so a block touches several
// times fewer distinct vertices than a strip of the same size (fewer splat atomics, fewer value rows to stage).
struct TileMap {
    int N, TP, W, H, TW, TH, tiles_x, ntiles;  // TW is a power of two
    // 1-D tiles only: tile t owns the SORTED positions [t * TP, (t + 1) * TP) of a point set whose own order is not
    // coherent, and position q is the point perm[q] (crf.cu: points sorted by their first two lattice vertices); NULL =
    // the points' own order.  Per-point lattice data is read, and Q / labels are written, through the indirection; the
    // unary rows arrive as a copy in sorted order, so their staging stays one contiguous bulk copy per tile.
    const int* perm;
};
// per-CTA constants of the mapping (the divisions happen once per CTA, not once per point)
struct TileOrigin {
    int x0, y0;        // first pixel of a 2-D tile
    long long base;    // first point of a 1-D tile
    int tw_shift;      // log2(TW)
};
__device__ __forceinline__ TileOrigin tile_origin(const TileMap& m, int tile) {
    TileOrigin o;
    o.x0 = o.y0 = 0;
    o.base = 0;
    o.tw_shift = 0;
    if (m.W == 0) {
        o.base = (long long)tile * m.TP;
    } else {
        const int ty = tile / m.tiles_x, tx = tile - ty * m.tiles_x;
        o.x0 = tx * m.TW;
        o.y0 = ty * m.TH;
        o.tw_shift = 31 - __clz(m.TW);
    }
    return o;
}
__device__ __forceinline__ int tile_point(const TileMap& m, const TileOrigin& o, int lp) {
    if (lp >= m.TP) return -1;
    if (m.W == 0) {
        const long long p = o.base + lp;
        return p < m.N ? (m.perm ? __ldg(m.perm + p) : (int)p) : -1;
    }
    const int ly = lp >> o.tw_shift, lx = lp & (m.TW - 1);
    const int x = o.x0 + lx, y = o.y0 + ly;
    return (x < m.W && y < m.H) ? y * m.W + x : -1;
}

// Per-tile data of one lattice, built once per lattice by tile_csr_build_kernel (tile t owns the slice
// [t * TP * d1, (t + 1) * TP * d1) of every array):
//   point-major, for the slice: pt_w(j, lp) = barycentric weight * post-normalisation * Potts weight * alpha, and
//     pt_slot(j, lp) = the tile-local row slot of the vertex (index into the value rows staged in shared memory);
//   vertex-major, for the splat: the distinct vertices the tile touches, cut into segments of <= TILE_SEG
//     (local point, weight * pre-normalisation) pairs; tile_vert[slot] = vertex id of a row slot.
struct FusedLat {
    const float* pt_w;         // D1 * TP floats per tile, pt_index() order
    const uint16_t* pt_slot;   // D1 * TP row slots per tile, pt_index() order
    const uint2* pairs;        // (byte offset of the local point's row in the shared Q tile - q_row(lp) - and weight bits); the 32
                               // consecutive segments one warp walks form a group, stored column-major (pair k of every
                               // segment of the group that has one, then pair k + 1, ...)
    const int2* ent_meta;      // per segment: x = first pair of the segment's GROUP | length << 16, y = vertex id
    const int* tile_vert;      // row slot -> vertex id
    const int2* tile_info;     // per tile: x = segments, y = distinct vertices (row slots)
    const float* vin;          // blurred value table of the previous iteration (slice source)
    float* vout;               // all-zero table receiving this iteration's splat
    const uint32_t* counts;    // [0] V, [1] overflow flag
};
struct FusedArgs {
    FusedLat lat[FUSED_MAX_LAT];
};
// Label layers of the CRF in the device channel layout: every layer starts at a multiple of 4 channels (a float4
// group) and its padding channels carry unary = +inf, so that exp(-unary - max) = 0 without any per-channel predicate.
struct FusedLayers {
    int n_layers;
    int off[RSS_MAX_LAYERS];      // first channel of the layer (multiple of 4)
    int count[RSS_MAX_LAYERS];    // labels of the layer
    unsigned gmask[RSS_MAX_LAYERS];  // bit g: float4 group g belongs to the layer
    int unknown[RSS_MAX_LAYERS];  // < 0: plain argmax
    float gate[RSS_MAX_LAYERS];   // 2 / M_l when gated (segmenter.cpp:647), else -inf
};
// position of (corner j, local point lp) inside a tile's point-major block of D1 * TILE_POINTS elements: corners are
// packed in vectors of 4, then 2, then 1, each vector array indexed by lp - so a thread fetches its D1 weights (floats)
// with ceil-ish(D1 / 4) fully coalesced vector loads, and its D1 row slots (uint16) likewise
__host__ __device__ constexpr int pt_index(int D1, int j, int lp) {
    const int n4 = D1 / 4, n2 = (D1 % 4) / 2;
    if (j < 4 * n4) return ((j / 4) * TILE_POINTS + lp) * 4 + j % 4;
    if (j < 4 * n4 + 2 * n2) return 4 * n4 * TILE_POINTS + lp * 2 + (j - 4 * n4);
    return (4 * n4 + 2 * n2) * TILE_POINTS + lp;
}
constexpr int BLUR_MAX_PHASES = 8;
struct BlurMultiArgs {
    int K;
    int phases;                   // grid phases of the launch (blur_multi_plan)
    int fuse[FUSED_MAX_LAT][BLUR_MAX_PHASES];  // axes lattice k blurs in phase p (0 = it has finished)
    float4* ping[FUSED_MAX_LAT];  // holds the splat on entry
    float4* pong[FUSED_MAX_LAT];
    float4* zero[FUSED_MAX_LAT];  // table to clear (next splat target) or NULL
    const int2* nbr[FUSED_MAX_LAT];
    const uint32_t* counts[FUSED_MAX_LAT];
    int d1[FUSED_MAX_LAT];
    uint32_t vcap[FUSED_MAX_LAT];
};
// what the build kernel writes for one lattice
struct TileCsrOut {
    uint2* pairs;
    int2* ent_meta;
    int* tile_vert;
    int2* tile_info;
    float* pt_w;
    uint16_t* pt_slot;
};

#ifndef RSS_TILE_SEG
#define RSS_TILE_SEG 32  // pairs per splat segment (one thread walks one segment serially)
#endif
#ifndef RSS_SPLAT_REV
#define RSS_SPLAT_REV 1  // the splat walks the second lattice's segments in reverse thread order (gather_entries)
#endif
#ifndef RSS_TILE_MINB
#define RSS_TILE_MINB 3  // resident CTAs per SM the "alone" build of the point kernel is compiled for (register budget)
#endif
#ifndef RSS_POINT_ALIAS
#define RSS_POINT_ALIAS 0  // the point kernel's splat lists reuse the shared memory of the staged value rows (4 CTAs per SM)
#endif
#ifndef RSS_QROW_SKEW
#define RSS_QROW_SKEW 3
#endif
// Row of local point lp in the point kernel's shared Q tile: the 32 rows of every warp (= one image row of a 32-pixel-wide
// tile) are rotated by RSS_QROW_SKEW * warp index.  The bank group of a row is (G * row + g) mod 8, i.e. - G odd - the row
// mod 8; unrotated that is x mod 8 for every image row, so a vertex whose pixel blob is narrower than 8 pixels only offers
// the splat a few of the eight classes however tall it is (the eight lanes of a quarter-warp read eight different rows
// conflict-free only if these differ mod 8).  With the skew the class is (x + 3 y) mod 8: a 2 x 8 blob covers all eight.
__host__ __device__ __forceinline__ int q_row(int lp) { return (lp & ~31) | ((lp + RSS_QROW_SKEW * (lp >> 5)) & 31); }
constexpr int TILE_SEG = RSS_TILE_SEG;
// the 64-register build of the point kernel (meanfield_shared.cu), for SMs shared with the cooperative blur
cudaError_t launch_meanfield_fused_shared(rss_ctx* c, cudaStream_t st, const FusedArgs& a, int d1a, int d1b, const float* unary,
                                          float* Q, uint8_t* labels, const TileMap& tm, int G, const FusedLayers& ls, int mode);

bool fused_group_supported(int G);  // channel-group counts the point kernel is instantiated for
bool fused_signature_supported(int G, int d1a, int d1b);
TileMap fused_tile_map(int N, int W, int H);  // W = H = 0 for point sets without an image grid
cudaError_t launch_meanfield_fused(rss_ctx* c, cudaStream_t st, const FusedArgs& a, int d1a, int d1b, const float* unary, float* Q,
                                   uint8_t* labels, const TileMap& tm, int G, const FusedLayers& ls, int mode);
// pre / post: fold norm[] into the splat / slice weights (NormalizationType); slice_scale = Potts weight * alpha
void launch_tile_csr_build(rss_ctx* c, cudaStream_t st, const int* offsets, const float* bary, const float* norm, bool pre,
                           bool post, float slice_scale, const TileMap& tm, int d1, int row_bytes, const uint32_t* counts,
                           const TileCsrOut& out);
struct BlurShape { int grid, block; };  // CTAs (one barrier arrival each) and threads per CTA of the cooperative blur
BlurShape blur_multi_shape(const rss_ctx* c);  // read ONCE per inference: the barrier targets depend on the grid
int blur_multi_plan(BlurMultiArgs& a, int G, int* phases_of);  // fills a.phases / a.fuse; phases per lattice -> phases_of
cudaError_t launch_blur_multi(rss_ctx* c, cudaStream_t st, BlurMultiArgs a, int G, unsigned int* barrier, unsigned int barrier_base,
                              const BlurShape& shape);
void launch_splat_ones_runs(rss_ctx* c, cudaStream_t st, const int* offsets, const float* bary, int N, int d1,
                            const uint32_t* counts, float* values);

}  // namespace rss
