// libforest multi-label forest inference + the frame worker's glue (low-res scatter, 2x bilinear upsample),
// sm_100a.  Reference: third-party/libforest/src/classifier.cpp:97-131,187-208; src/segmenter.cpp:355-431.
//
// Trees are flattened once (rss_create) into one array of 16-byte nodes {feature, threshold, left child,
// leaf row}: one LDG.128 per level, the whole forest (a few hundred KB) stays L1/L2 resident.  Leaf
// histograms live in a dense [leaf][sumC] table.  One thread per (sample, tree); the per-tree leaf rows are
// then summed in tree order t = 0,1,..,T-1 exactly like RandomForest::multiClassLogPosterior, so the summed
// log-posteriors are bit-identical to the reference's.
#include "kernels.hpp"

namespace rss {

__device__ __forceinline__ Node load_node(const Node* p) {
    const int4 v = __ldg(reinterpret_cast<const int4*>(p));
    Node n;
    n.feat = v.x;
    n.thr = __int_as_float(v.y);
    n.left = v.z;
    n.leaf = v.w;
    return n;
}

// DecisionTree::findLeafNode (classifier.cpp:97-117): strict fp32 '<', right child = left + 1
__global__ void __launch_bounds__(256) forest_traverse_kernel(const Node* __restrict__ nodes,
                                                              const int* __restrict__ tree_off, int T,
                                                              const float* __restrict__ feats, int D, int n, int ld,
                                                              int* __restrict__ leaf_ids) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)n * T) return;
    const int t = (int)(gid / n), s = (int)(gid - (long long)t * n);  // tree-major: a warp walks one tree
    const Node* tree = nodes + tree_off[t];
    const float* x = feats + (size_t)s * D;
    int node = 0;
    Node nd = load_node(tree);
    while (nd.left != 0) {
        node = __ldg(x + nd.feat) < nd.thr ? nd.left : nd.left + 1;
        nd = load_node(tree + node);
    }
    leaf_ids[(size_t)t * ld + s] = node;
}
void launch_forest_traverse(rss_ctx* c, cudaStream_t st, const Node* nodes, const int* tree_off, int T,
                            const float* feats, int D, int n, int ld, int* leaf_ids) {
    if (n <= 0) return;
    RSS_LAUNCH(c, forest_traverse_kernel, rss_div_up((long long)n * T, 256), 256, 0, st, nodes, tree_off, T, feats,
               D, n, ld, leaf_ids);
}

// RandomForest::multiClassLogPosterior (classifier.cpp:187-208): tree 0's leaf row, then += trees 1..T-1
__global__ void __launch_bounds__(256) forest_posterior_kernel(const Node* __restrict__ nodes,
                                                               const int* __restrict__ tree_off, int T,
                                                               const float* __restrict__ leaves, int sumC,
                                                               const int* __restrict__ leaf_ids, int n, int ld,
                                                               float* __restrict__ post) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)n * sumC) return;
    const int s = (int)(gid / sumC), k = (int)(gid - (long long)s * sumC);
    float acc = 0.f;
    for (int t = 0; t < T; t++) {
        const int node = leaf_ids[(size_t)t * ld + s];
        const int row = __ldg(&nodes[tree_off[t] + node].leaf);
        const float h = __ldg(leaves + (size_t)row * sumC + k);
        acc = t == 0 ? h : __fadd_rn(acc, h);
    }
    post[(size_t)s * sumC + k] = acc;
}
void launch_forest_posterior(rss_ctx* c, cudaStream_t st, const Node* nodes, const int* tree_off, int T,
                             const float* leaves, int sumC, const int* leaf_ids, int n, int ld, float* post) {
    if (n <= 0) return;
    RSS_LAUNCH(c, forest_posterior_kernel, rss_div_up((long long)n * sumC, 256), 256, 0, st, nodes, tree_off, T,
               leaves, sumC, leaf_ids, n, ld, post);
}

// ------------------------------------------------------------------------------------------------
// segmenter.cpp:355-376: per layer a low-res image [gh][gw][C_l], pre-filled, samples scattered at
// (y/stride, C_l*x/stride).  Layers are stored back to back.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fill_kernel(float* __restrict__ p, size_t n, float v) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t step = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += step) p[i] = v;
}
void launch_lowres_fill(rss_ctx* c, cudaStream_t st, float* lowres, size_t n, float fill) {
    if (n == 0) return;
    const int blocks = (int)min((size_t)c->sm_count * 8, (n + 255) / 256);
    RSS_LAUNCH(c, fill_kernel, blocks, 256, 0, st, lowres, n, fill);
}
struct LayerDims {
    int L;
    int C[RSS_MAX_LAYERS];
    int coff[RSS_MAX_LAYERS];  // class offset of the layer inside a sumC row
    int uoff[RSS_MAX_LAYERS];  // channel offset of the layer inside a CRF unary row (layers padded to 4 channels)
};
__global__ void __launch_bounds__(256) lowres_scatter_kernel(const float* __restrict__ post, int sumC,
                                                             const int* __restrict__ xs, const int* __restrict__ ys,
                                                             int n, int stride, int gw, int gh, LayerDims ld,
                                                             float* __restrict__ lowres) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)n * sumC) return;
    const int s = (int)(gid / sumC), k = (int)(gid - (long long)s * sumC);
    int l = 0;
    while (l + 1 < ld.L && k >= ld.coff[l + 1]) l++;
    const int cl = k - ld.coff[l], C = ld.C[l];
    const size_t layer_base = (size_t)gw * gh * ld.coff[l];
    const int gx = xs[s] / stride, gy = ys[s] / stride;
    lowres[layer_base + ((size_t)gy * gw + gx) * C + cl] = post[gid];
}
static LayerDims make_dims(int L, const int* C) {
    LayerDims d;
    d.L = L;
    int off = 0, uoff = 0;
    for (int l = 0; l < RSS_MAX_LAYERS; l++) {
        d.C[l] = l < L ? C[l] : 0;
        d.coff[l] = off;
        d.uoff[l] = uoff;
        off += d.C[l];
        uoff += (d.C[l] + 3) / 4 * 4;  // the CRF's device rows start every layer on a float4 group (lattice.cuh, rss_crf)
    }
    return d;
}
void launch_lowres_scatter(rss_ctx* c, cudaStream_t st, const float* post, int sumC, const int* xs, const int* ys,
                           int n, int stride, int gw, int gh, int L, const int* C, float* lowres) {
    if (n <= 0) return;
    RSS_LAUNCH(c, lowres_scatter_kernel, rss_div_up((long long)n * sumC, 256), 256, 0, st, post, sumC, xs, ys, n,
               stride, gw, gh, make_dims(L, C), lowres);
}

// ------------------------------------------------------------------------------------------------
// cv::resize(INTER_LINEAR) on 32FC(C_l), gw x gh -> W x H (segmenter.cpp:380-382), written straight into the
// flattened [layer][y][x][class] vector (:413-431).  OpenCV semantics: source coordinate
// f = (float)((d+0.5)*scale-0.5) with scale = 1/(dst/src) in double; horizontally the fraction is zeroed when
// the index is clamped, vertically the two row indices are clipped and the fraction kept; horizontal pass
// first (S0*a0 + S1*a1), then vertical (H0*b0 + H1*b1), each product and sum rounded to float.
// ------------------------------------------------------------------------------------------------
// One thread per OUTPUT PIXEL: the source coordinates, weights and row pointers are computed once and reused for all
// classes of all layers (a thread per output element would redo the double-precision coordinate arithmetic sumC times).
// UNARY = false: out is the flattened [layer][y][x][class] vector.  UNARY = true (keyframe path): the value is negated
// (energy = -log-posterior, src/segmenter.cpp:642) and written straight into the CRF's [pixel][Mp] unary layout, which
// saves the separate posterior -> unary pass; -x is exact, so the energies are bit-identical to the two-pass route.
template <bool UNARY>
__global__ void __launch_bounds__(256) upsample_kernel(const float* __restrict__ lowres, int gw, int gh, int W,
                                                       int H, LayerDims ld, int sumC, int Mp, double scale_x,
                                                       double scale_y, float* __restrict__ out) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= W * H) return;
    const int y = p / W, x = p - y * W;
    float fx = (float)(((double)x + 0.5) * scale_x - 0.5);
    int sx = (int)floorf(fx);
    fx = __fsub_rn(fx, (float)sx);
    if (sx < 0) { sx = 0; fx = 0.f; }
    if (sx >= gw - 1) { sx = gw - 1; fx = 0.f; }
    const int sx1 = min(sx + 1, gw - 1);
    float fy = (float)(((double)y + 0.5) * scale_y - 0.5);
    const int sy = (int)floorf(fy);
    fy = __fsub_rn(fy, (float)sy);
    const int y0 = min(max(sy, 0), gh - 1), y1 = min(max(sy + 1, 0), gh - 1);
    const float a0 = __fsub_rn(1.f, fx), a1 = fx, b0 = __fsub_rn(1.f, fy), b1 = fy;
    for (int l = 0; l < ld.L; l++) {
        const int C = ld.C[l];
        const float* src = lowres + (size_t)gw * gh * ld.coff[l];
        const float* r00 = src + ((size_t)y0 * gw + sx) * C;
        const float* r01 = src + ((size_t)y0 * gw + sx1) * C;
        const float* r10 = src + ((size_t)y1 * gw + sx) * C;
        const float* r11 = src + ((size_t)y1 * gw + sx1) * C;
        float* o = UNARY ? out + (size_t)p * Mp + ld.uoff[l] : out + (size_t)W * H * ld.coff[l] + (size_t)p * C;
        for (int cl = 0; cl < C; cl++) {
            const float h0 = __fadd_rn(__fmul_rn(__ldg(r00 + cl), a0), __fmul_rn(__ldg(r01 + cl), a1));
            const float h1 = __fadd_rn(__fmul_rn(__ldg(r10 + cl), a0), __fmul_rn(__ldg(r11 + cl), a1));
            const float v = __fadd_rn(__fmul_rn(h0, b0), __fmul_rn(h1, b1));
            o[cl] = UNARY ? -v : v;
        }
    }
}
// The keyframe path's variant: one thread per (pixel, float4 channel group) of the CRF's [pixel][Mp] unary matrix, so a warp
// stores 512 contiguous bytes (the per-pixel kernel above stores 17 scalars per thread at an 80-byte stride: 6.9e6 L2
// sectors for 26 MB).  Same arithmetic per value; the padding channels of a layer get +inf like unary_init_kernel.
__global__ void __launch_bounds__(256) upsample_unary_groups_kernel(const float* __restrict__ lowres, int gw, int gh, int W, int H,
                                                                    LayerDims ld, int G, double scale_x, double scale_y,
                                                                    float4* __restrict__ out) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)W * H * G) return;
    const int p = (int)(gid / G), g = (int)(gid - (long long)p * G);
    const int y = p / W, x = p - y * W;
    float fx = (float)(((double)x + 0.5) * scale_x - 0.5);
    int sx = (int)floorf(fx);
    fx = __fsub_rn(fx, (float)sx);
    if (sx < 0) { sx = 0; fx = 0.f; }
    if (sx >= gw - 1) { sx = gw - 1; fx = 0.f; }
    const int sx1 = min(sx + 1, gw - 1);
    float fy = (float)(((double)y + 0.5) * scale_y - 0.5);
    const int sy = (int)floorf(fy);
    fy = __fsub_rn(fy, (float)sy);
    const int y0 = min(max(sy, 0), gh - 1), y1 = min(max(sy + 1, 0), gh - 1);
    const float a0 = __fsub_rn(1.f, fx), a1 = fx, b0 = __fsub_rn(1.f, fy), b1 = fy;
    int l = 0;
    while (l + 1 < ld.L && 4 * g >= ld.uoff[l + 1]) l++;
    const int C = ld.C[l], c0 = 4 * g - ld.uoff[l];
    const float* src = lowres + (size_t)gw * gh * ld.coff[l];
    const float* r00 = src + ((size_t)y0 * gw + sx) * C;
    const float* r01 = src + ((size_t)y0 * gw + sx1) * C;
    const float* r10 = src + ((size_t)y1 * gw + sx) * C;
    const float* r11 = src + ((size_t)y1 * gw + sx1) * C;
    float v[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int cl = c0 + k;
        v[k] = INFINITY;
        if (cl < C) {
            const float h0 = __fadd_rn(__fmul_rn(__ldg(r00 + cl), a0), __fmul_rn(__ldg(r01 + cl), a1));
            const float h1 = __fadd_rn(__fmul_rn(__ldg(r10 + cl), a0), __fmul_rn(__ldg(r11 + cl), a1));
            v[k] = -__fadd_rn(__fmul_rn(h0, b0), __fmul_rn(h1, b1));
        }
    }
    out[gid] = make_float4(v[0], v[1], v[2], v[3]);
}
void launch_upsample(rss_ctx* c, cudaStream_t st, const float* lowres, int gw, int gh, int W, int H, int L,
                     const int* C, float* posteriors, int unary_stride) {
    LayerDims d = make_dims(L, C);
    int sumC = 0;
    for (int l = 0; l < L; l++) sumC += C[l];
    const double scale_x = 1.0 / ((double)W / (double)gw), scale_y = 1.0 / ((double)H / (double)gh);
    const long long total = (long long)W * H;  // one thread per output pixel
    if (unary_stride > 0 && unary_stride % 4 == 0 && d.uoff[L - 1] + ((C[L - 1] + 3) & ~3) == unary_stride)
        RSS_LAUNCH(c, upsample_unary_groups_kernel, rss_div_up(total * (unary_stride / 4), 256), 256, 0, st, lowres, gw, gh, W, H, d,
                   unary_stride / 4, scale_x, scale_y, reinterpret_cast<float4*>(posteriors));
    else if (unary_stride > 0)
        RSS_LAUNCH(c, upsample_kernel<true>, rss_div_up(total, 256), 256, 0, st, lowres, gw, gh, W, H, d, sumC, unary_stride,
                   scale_x, scale_y, posteriors);
    else
        RSS_LAUNCH(c, upsample_kernel<false>, rss_div_up(total, 256), 256, 0, st, lowres, gw, gh, W, H, d, sumC, 0, scale_x,
                   scale_y, posteriors);
}

}  // namespace rss
