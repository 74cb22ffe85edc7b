// Argument blocks of the fused mean-field kernels (meanfield.cu).
#pragma once
#include "common.cuh"

namespace rss {

constexpr int FUSED_MAX_LAT = 2;

// How a tile's local point index maps to the point index.  W == 0: tile t owns the TP consecutive points starting at
// t * TP (generic point sets).  W > 0: the points are the pixels of a W x H image in raster order and a tile is a
// TW x TH pixel block - neighbouring pixels in BOTH directions share lattice vertices, so a block touches several
// times fewer distinct vertices than a strip of the same size (fewer splat atomics, better L1 reuse of value rows).
struct TileMap {
    int N, TP, W, H, TW, TH, tiles_x, ntiles;  // TW is a power of two
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
        return p < m.N ? (int)p : -1;
    }
    const int ly = lp >> o.tw_shift, lx = lp & (m.TW - 1);
    const int x = o.x0 + lx, y = o.y0 + ly;
    return (x < m.W && y < m.H) ? y * m.W + x : -1;
}

struct FusedLat {
    const int* offsets;   // [N][d+1] vertex ids
    const float* bary;    // [N][d+1]
    const float* norm;    // [N]
    const float* vin;     // blurred value table of the previous iteration (slice source)
    float* vout;          // all-zero table receiving this iteration's splat
    float potts, alpha;   // Potts weight w, slice scale 1/(1+2^-d)
    int pre, post;        // normalisation applied before the splat / after the slice (NormalizationType)
};
struct FusedArgs {
    FusedLat lat[FUSED_MAX_LAT];
    const uint32_t* counts[FUSED_MAX_LAT];  // [0] V, [1] overflow flag
    // tile-local CSR of the splat matrix (tile t owns the slice [t * TP * d1, (t + 1) * TP * d1) of each array)
    const uint2* pairs[FUSED_MAX_LAT];      // (byte offset of the local point's row in the shared tile, weight bits), by entry
    const int2* ent_meta[FUSED_MAX_LAT];    // per segment: x = start in the tile's pair array | length << 16, y = vertex id
    const int* tile_nent[FUSED_MAX_LAT];    // segments per tile
};
struct FusedLayers {
    int n_layers;
    int off[RSS_MAX_LAYERS + 1];
    int unknown[RSS_MAX_LAYERS];  // < 0: plain argmax
    float gate[RSS_MAX_LAYERS];   // 2 / M_l when gated (segmenter.cpp:647), else -inf
    int aligned;                  // every layer boundary is a multiple of 4 channels
    // per channel group g (channels 4g..4g+3): layer membership nibbles (4 bits per layer), and for aligned layers the
    // group's layer (-1: padding only) and its valid-channel nibble
    unsigned group_lmask[8];
    unsigned group_valid[8];
    int group_layer[8];
};
struct BlurMultiArgs {
    int K;
    float4* ping[FUSED_MAX_LAT];  // holds the splat on entry
    float4* pong[FUSED_MAX_LAT];
    float4* zero[FUSED_MAX_LAT];  // table to clear (next splat target) or NULL
    const int2* nbr[FUSED_MAX_LAT];
    const uint32_t* counts[FUSED_MAX_LAT];
    int d1[FUSED_MAX_LAT];
    uint32_t vcap[FUSED_MAX_LAT];
};

bool fused_group_supported(int G);  // channel-group counts the tile kernel is instantiated for
bool fused_signature_supported(int G, int d1a, int d1b);
TileMap fused_tile_map(int G, int N, int W, int H, int sm_count);  // W = H = 0 for point sets without an image grid
void launch_meanfield_fused(rss_ctx* c, cudaStream_t st, const FusedArgs& a, int d1a, int d1b, const float* unary, float* Q,
                            uint8_t* labels, const TileMap& tm, int G, const FusedLayers& ls, int mode);
void launch_tile_csr_build(rss_ctx* c, cudaStream_t st, const int* offsets, const float* bary, const float* norm,
                           const TileMap& tm, int d1, int row_bytes, const uint32_t* counts, uint2* pairs, int2* ent_meta,
                           int* tile_nent);
int blur_multi_grid(const rss_ctx* c);  // CTAs of the cooperative blur (one barrier arrival each)
void launch_blur_multi(rss_ctx* c, cudaStream_t st, BlurMultiArgs a, int G, unsigned int* barrier, unsigned int barrier_base);
void launch_splat_ones_runs(rss_ctx* c, cudaStream_t st, const int* offsets, const float* bary, int N, int d1,
                            const uint32_t* counts, float* values);

}  // namespace rss
