// Per-pixel feature extraction kernels (reference include/feature_extractor.h:41-291), sm_100a.
// Integer / byte work, HBM- and L2-bound: coalesced vectorised reads, the Lab image is stored as uchar4
// so that every bilinear tap is one aligned 32-bit load.  Compiled with -fmad=false: every float
// product and sum below rounds separately, like the reference's SSE2 build.
#include "kernels.hpp"
#include "normals.cuh"

namespace rss {

// ------------------------------------------------------------------------------------------------
// F1: cv::cvtColor(CV_BGR2Lab, 8U) + cv::copyMakeBorder(BORDER_REFLECT, P)   (:129-130)
// One thread per bordered pixel; the source pixel is found by reflection (fedcba|abcdefgh|hgfedcb).
// gamma (256 x u16) and cube-root (3072 x u16) LUTs are read through the read-only path (L1-resident).
// ------------------------------------------------------------------------------------------------
__constant__ int c_lab_coef[9] = {778, 1541, 1777, 296, 2929, 871, 3575, 448, 73};

__device__ __forceinline__ int reflect_index(int p, int n) {
    if (n == 1) return 0;
    while (p < 0 || p >= n) p = p < 0 ? -p - 1 : 2 * n - 1 - p;
    return p;
}
__device__ __forceinline__ int sat_u8(int v) { return min(max(v, 0), 255); }

__global__ void __launch_bounds__(256) lab_border_kernel(const uint8_t* __restrict__ rgb, int W, int H, int P,
                                                         const uint16_t* __restrict__ gamma,
                                                         const uint16_t* __restrict__ cbrt_tab,
                                                         uchar4* __restrict__ lab) {
    const int Wb = W + 2 * P, Hb = H + 2 * P;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= Wb || y >= Hb) return;
    const int sx = reflect_index(x - P, W), sy = reflect_index(y - P, H);
    const uint8_t* px = rgb + ((size_t)sy * W + sx) * 3;
    const int c0 = __ldg(gamma + px[0]), c1 = __ldg(gamma + px[1]), c2 = __ldg(gamma + px[2]);
    const int fX = __ldg(cbrt_tab + ((c0 * c_lab_coef[0] + c1 * c_lab_coef[1] + c2 * c_lab_coef[2] + 2048) >> 12));
    const int fY = __ldg(cbrt_tab + ((c0 * c_lab_coef[3] + c1 * c_lab_coef[4] + c2 * c_lab_coef[5] + 2048) >> 12));
    const int fZ = __ldg(cbrt_tab + ((c0 * c_lab_coef[6] + c1 * c_lab_coef[7] + c2 * c_lab_coef[8] + 2048) >> 12));
    const int L = (296 * fY - 1336934 + 16384) >> 15;
    const int a = (500 * (fX - fY) + 128 * 32768 + 16384) >> 15;
    const int b = (200 * (fY - fZ) + 128 * 32768 + 16384) >> 15;
    lab[(size_t)y * Wb + x] = make_uchar4((unsigned char)sat_u8(L), (unsigned char)sat_u8(a), (unsigned char)sat_u8(b), 0);
}

void launch_lab_border(rss_ctx* c, cudaStream_t st, const uint8_t* rgb, int W, int H, int P, uchar4* lab) {
    dim3 grid(rss_div_up(W + 2 * P, 256), H + 2 * P);
    RSS_LAUNCH(c, lab_border_kernel, grid, 256, 0, st, rgb, W, H, P, c->lab_gamma.as<uint16_t>(),
               c->lab_cbrt.as<uint16_t>(), lab);
}

// ------------------------------------------------------------------------------------------------
// F3: point cloud (:200-232).  rect = ((m0*v0 + m1*v1) + m2*v2) + t with v = [d*x, d*y, d], NaN when the
// depth (in metres, float) is outside [dmin, dmax].  One thread per pixel, float4 store.
// ------------------------------------------------------------------------------------------------
// The pose (M = R * Kinv, t) is read from device memory (rss_ctx::pose_dev), not passed by value: the keyframe's CUDA
// graph is captured once and replayed with a new pose every keyframe.
__global__ void __launch_bounds__(256) cloud_kernel(const uint16_t* __restrict__ depth, int W, int H,
                                                    const PoseParams* __restrict__ pose, float dmin, float dmax,
                                                    float4* __restrict__ xyz) {
    const PoseParams p = *pose;
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const size_t i = (size_t)y * W + x;
    const float d = __fdiv_rn((float)depth[i], 1000.0f);
    float v0, v1, v2;
    if (d < dmin || d > dmax) {
        v0 = v1 = v2 = __int_as_float(0x7fc00000);
    } else {
        v0 = __fmul_rn(d, (float)x);
        v1 = __fmul_rn(d, (float)y);
        v2 = d;
    }
    float o[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        float s = __fadd_rn(__fmul_rn(p.M[3 * k], v0), __fmul_rn(p.M[3 * k + 1], v1));
        s = __fadd_rn(s, __fmul_rn(p.M[3 * k + 2], v2));
        o[k] = __fadd_rn(s, p.t[k]);
    }
    xyz[i] = make_float4(o[0], o[1], o[2], 0.f);
}
void launch_cloud(rss_ctx* c, cudaStream_t st, const uint16_t* depth, int W, int H, const PoseParams* pose_dev, float dmin,
                  float dmax, float4* xyz) {
    dim3 grid(rss_div_up(W, 256), H);
    RSS_LAUNCH(c, cloud_kernel, grid, 256, 0, st, depth, W, H, pose_dev, dmin, dmax, xyz);
}

// ------------------------------------------------------------------------------------------------
// F0: sample selection (:56-121) -> one flag per stride-grid position; compaction keeps raster order.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) select_kernel(const uint16_t* __restrict__ depth,
                                                     const int8_t* __restrict__ labels, int n_layers,
                                                     int extract_type, int W, int H, int stride, int gw, int gh,
                                                     float dmin_mm, float dmax_mm, uint32_t* __restrict__ flags) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= gw * gh) return;
    const int x = (g % gw) * stride, y = (g / gw) * stride;
    const size_t i = (size_t)y * W + x;
    const float d = (float)depth[i];
    bool ok = d >= dmin_mm && d <= dmax_mm;
    if (ok && extract_type == RSS_WITH_POSITIVE_LABEL)
        for (int l = 0; l < n_layers; l++) ok = ok && labels[(size_t)l * W * H + i] >= 0;
    flags[g] = ok ? 1u : 0u;
}
void launch_select(rss_ctx* c, cudaStream_t st, const uint16_t* depth, const int8_t* labels, int n_label_layers,
                   int extract_type, int W, int H, int stride, float dmin_mm, float dmax_mm, uint32_t* flags) {
    const int gw = rss_div_up(W, stride), gh = rss_div_up(H, stride);
    RSS_LAUNCH(c, select_kernel, rss_div_up((long long)gw * gh, 256), 256, 0, st, depth, labels, n_label_layers,
               extract_type, W, H, stride, gw, gh, dmin_mm, dmax_mm, flags);
}
__global__ void __launch_bounds__(256) compact_kernel(const uint32_t* __restrict__ flags,
                                                      const uint32_t* __restrict__ sidx, int W, int H, int stride,
                                                      int gw, int gh, const int8_t* __restrict__ labels, int n_layers,
                                                      int* __restrict__ xs, int* __restrict__ ys,
                                                      int* __restrict__ slabels) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= gw * gh || !flags[g]) return;
    const int x = (g % gw) * stride, y = (g / gw) * stride;
    const uint32_t s = sidx[g];
    xs[s] = x;
    ys[s] = y;
    if (labels && slabels)
        for (int l = 0; l < n_layers; l++) slabels[(size_t)s * n_layers + l] = labels[(size_t)l * W * H + (size_t)y * W + x];
}
void launch_compact(rss_ctx* c, cudaStream_t st, const uint32_t* flags, const uint32_t* sidx, int W, int H,
                    int stride, const int8_t* labels, int n_label_layers, int* xs, int* ys, int* slabels) {
    const int gw = rss_div_up(W, stride), gh = rss_div_up(H, stride);
    RSS_LAUNCH(c, compact_kernel, rss_div_up((long long)gw * gh, 256), 256, 0, st, flags, sidx, W, H, stride, gw, gh,
               labels, n_label_layers, xs, ys, slabels);
}

// ------------------------------------------------------------------------------------------------
// F2: colour patch (:134-173).  For a sample at depth d the reference crops a (2h+1)^2 ROI,
// h = int(P / (2.0*d)), from the bordered Lab image and cv::resize()s it to r x r (INTER_LINEAR, 8U).
// Every output pixel of that resize is a fixed-point blend of exactly four source pixels, whatever
// the ROI size, so the work per sample is r*r*4 taps - not (2h+1)^2 pixels.  The tap table
// (ResizeTap[h][r], built on the host with OpenCV's float arithmetic) holds indices and 11-bit weights.
// One warp per sample: lanes walk the r*r output pixels, stage the 3r^2 floats in shared memory and
// write the feature row with coalesced stores.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int patch_half(int P, float depth_m) {
    // int current_size_half = _patch_size/(2.0*depth);  (int / double, truncated)  feature_extractor.h:140
    const double q = (double)P / (2.0 * (double)depth_m);
    int h = (q >= (double)P) ? P : (int)q;  // the reference assumes depth >= 0.5 m (:37); clamp instead of reading out of bounds
    return h < 0 ? 0 : h;
}
__device__ __forceinline__ int resize_blend(int p00, int p01, int p10, int p11, int a0, int a1, int b0, int b1) {
    const int h0 = p00 * a0 + p01 * a1;
    const int h1 = p10 * a0 + p11 * a1;
    return (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
}

constexpr int PATCH_WARPS = 8;
__global__ void __launch_bounds__(PATCH_WARPS * 32) patch_features_kernel(
    const uchar4* __restrict__ lab, const uint16_t* __restrict__ depth, int W, int P, int r,
    const ResizeTap* __restrict__ tapx, const ResizeTap* __restrict__ tapy, const int* __restrict__ xs,
    const int* __restrict__ ys, int n, float* __restrict__ feats, int D) {
    extern __shared__ float stage[];  // PATCH_WARPS * 3r^2
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int s = blockIdx.x * PATCH_WARPS + warp;
    if (s >= n) return;
    const int nf = 3 * r * r;
    float* my = stage + warp * nf;
    const int x = xs[s], y = ys[s];
    const int Wb = W + 2 * P;
    const float d = __fdiv_rn((float)depth[(size_t)y * W + x], 1000.0f);
    const int half = patch_half(P, d);
    const uchar4* roi = lab + (size_t)(y + P - half) * Wb + (x + P - half);
    for (int k = lane; k < r * r; k += 32) {
        const int dy = k / r, dx = k - dy * r;
        const ResizeTap tx = tapx[half * r + dx], ty = tapy[half * r + dy];
        const uchar4 p00 = __ldg(roi + (size_t)ty.i0 * Wb + tx.i0), p01 = __ldg(roi + (size_t)ty.i0 * Wb + tx.i1);
        const uchar4 p10 = __ldg(roi + (size_t)ty.i1 * Wb + tx.i0), p11 = __ldg(roi + (size_t)ty.i1 * Wb + tx.i1);
        my[3 * k + 0] = (float)sat_u8(resize_blend(p00.x, p01.x, p10.x, p11.x, tx.w0, tx.w1, ty.w0, ty.w1));
        my[3 * k + 1] = (float)sat_u8(resize_blend(p00.y, p01.y, p10.y, p11.y, tx.w0, tx.w1, ty.w0, ty.w1));
        my[3 * k + 2] = (float)sat_u8(resize_blend(p00.z, p01.z, p10.z, p11.z, tx.w0, tx.w1, ty.w0, ty.w1));
    }
    __syncwarp();
    float* out = feats + (size_t)s * D;
    for (int k = lane; k < nf; k += 32) out[k] = my[k];
}
void launch_patch_features(rss_ctx* c, cudaStream_t st, const uchar4* lab, const uint16_t* depth, int W, int H,
                           int P, int r, const ResizeTap* tapx, const ResizeTap* tapy, const int* xs,
                           const int* ys, int n, float* feats, int D) {
    (void)H;
    if (n <= 0) return;
    const size_t smem = (size_t)PATCH_WARPS * 3 * r * r * sizeof(float);
    RSS_LAUNCH(c, patch_features_kernel, rss_div_up(n, PATCH_WARPS), PATCH_WARPS * 32, smem, st, lab, depth, W, P, r,
               tapx, tapy, xs, ys, n, feats, D);
}

// ------------------------------------------------------------------------------------------------
// depth (:180-197), height (:236-251), normal angle (:265-291) features for the compacted samples.
// The normal itself is PCL's AVERAGE_3D_GRADIENT estimate evaluated only where a sample needs it
// (see normals.cu for the integral images and the distance map).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) scalar_features_kernel(const uint16_t* __restrict__ depth,
                                                              const float4* __restrict__ xyz,
                                                              const float* __restrict__ dist,
                                                              const double* __restrict__ integ,
                                                              const int* __restrict__ cnt, int W, int H,
                                                              const int* __restrict__ xs, const int* __restrict__ ys,
                                                              int n, float* __restrict__ feats, int D, int pos_depth,
                                                              int pos_height, int pos_normal) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const int x = xs[s], y = ys[s];
    const size_t i = (size_t)y * W + x;
    float* f = feats + (size_t)s * D;
    if (pos_depth >= 0) f[pos_depth] = __fdiv_rn((float)depth[i], 1000.0f);
    if (pos_height >= 0) f[pos_height] = xyz[i].z;
    if (pos_normal >= 0) {
        const float3 nrm = pcl_normal_at(xyz, dist, integ, cnt, W, H, x, y);
        // acos(fabs(float)) in the reference binds to the double overloads of <math.h>
        f[pos_normal] = isnan(nrm.x) ? -2.0f : (float)acos(fabs((double)nrm.z));
    }
}
// ------------------------------------------------------------------------------------------------
// Forest traversal straight from the frame (segment_frame / keyframe path): DecisionTree::findLeafNode
// (classifier.cpp:97-117) where x[f] is EVALUATED ON DEMAND from the Lab image, the depth image, the cloud and the
// integral images instead of being read from a materialised [n][366] matrix.  A tree touches at most 31 of the 366
// features of a sample, so this skips ~90 % of the resize arithmetic and the 112 MB feature matrix never exists
// ("~6 MB if fused with F2").  The value of a feature is computed by exactly the same integer / float
// operations as patch_features_kernel and scalar_features_kernel, so leaf ids stay bit-exact.
// One thread per (tree, sample); feat_xy[k] = dx | dy << 8 of patch pixel k.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ Node load_node_ro(const Node* p) {
    const int4 v = __ldg(reinterpret_cast<const int4*>(p));
    Node n;
    n.feat = v.x; n.thr = __int_as_float(v.y); n.left = v.z; n.leaf = v.w;
    return n;
}
#ifndef RSS_FOREST_MINB
#define RSS_FOREST_MINB 8  // resident CTAs per SM (64 registers): the traversal is a chain of dependent loads, more warps
                           // hide it (measured: 120 us at 4, 102 us at 8)
#endif
struct FrameFeat {  // everything the on-demand feature evaluation reads
    const uchar4* lab;
    const uint16_t* depth;
    const float4* xyz;
    const float* dist;
    const double* integ;
    const int* cnt;
    const ResizeTap* tapx;
    const ResizeTap* tapy;
    const uint16_t* feat_xy;
    int W, H, P, r, ncolor, pos_depth, pos_height, pos_normal;
};
// findLeafNode for the sample at pixel (x, y); returns the leaf's row in the dense leaf table
__device__ __forceinline__ int traverse_frame(const Node* __restrict__ tree, const FrameFeat& F, int x, int y) {
    const size_t i = (size_t)y * F.W + x;
    const int Wb = F.W + 2 * F.P;
    const float dm = __fdiv_rn((float)F.depth[i], 1000.0f);
    const int half = patch_half(F.P, dm);
    const uchar4* roi = F.lab + (size_t)(y + F.P - half) * Wb + (x + F.P - half);
    const ResizeTap* tx_h = F.tapx + half * F.r;
    const ResizeTap* ty_h = F.tapy + half * F.r;
    bool have_normal = false;
    float normal_feature = 0.f;
    Node nd = load_node_ro(tree);
    while (nd.left != 0) {
        // both children are fetched BEFORE the feature value is known: the node load leaves the dependent chain (feature
        // offset -> taps -> four pixels -> compare -> node).  Siblings are adjacent, and the loader places every root at an
        // odd node index, so the pair (left, left + 1) - left is odd inside a libforest tree - is one aligned 32-byte sector.
        const int4* ch = reinterpret_cast<const int4*>(tree + nd.left);
        const int4 cl = __ldg(ch), cr = __ldg(ch + 1);
        float v;
        const int f = nd.feat;
        if (f < F.ncolor) {
            const int k = f / 3, ch = f - 3 * k;
            const int xy = __ldg(F.feat_xy + k);
            const ResizeTap tx = tx_h[xy & 255], ty = ty_h[xy >> 8];
            const uchar4 p00 = __ldg(roi + (size_t)ty.i0 * Wb + tx.i0), p01 = __ldg(roi + (size_t)ty.i0 * Wb + tx.i1);
            const uchar4 p10 = __ldg(roi + (size_t)ty.i1 * Wb + tx.i0), p11 = __ldg(roi + (size_t)ty.i1 * Wb + tx.i1);
            const int a = ch == 0 ? p00.x : (ch == 1 ? p00.y : p00.z), b = ch == 0 ? p01.x : (ch == 1 ? p01.y : p01.z);
            const int c = ch == 0 ? p10.x : (ch == 1 ? p10.y : p10.z), d = ch == 0 ? p11.x : (ch == 1 ? p11.y : p11.z);
            v = (float)sat_u8(resize_blend(a, b, c, d, tx.w0, tx.w1, ty.w0, ty.w1));
        } else if (f == F.pos_depth) {
            v = dm;
        } else if (f == F.pos_height) {
            v = F.xyz[i].z;
        } else {
            if (!have_normal) {
                const float3 nrm = pcl_normal_at(F.xyz, F.dist, F.integ, F.cnt, F.W, F.H, x, y);
                normal_feature = isnan(nrm.x) ? -2.0f : (float)acos(fabs((double)nrm.z));
                have_normal = true;
            }
            v = normal_feature;
        }
        const int4 nx = v < nd.thr ? cl : cr;
        nd.feat = nx.x; nd.thr = __int_as_float(nx.y); nd.left = nx.z; nd.leaf = nx.w;
    }
    return nd.leaf;
}

// The frame worker's forest stage in ONE kernel (segmenter.cpp:349-376): for the 32 consecutive stride-grid positions
// of a CTA, warp w walks tree t0 + w for all 32 positions (tree-major: the lanes of a warp visit the same nodes near the
// root and read neighbouring pixels), the leaf rows meet in shared memory, and the CTA sums them IN TREE ORDER
// (RandomForest::multiClassLogPosterior, classifier.cpp:187-208: tree 0's row, then += trees 1..T-1, so the sums are
// bit-identical) and writes the per-layer low-resolution images [gh][gw][C_l] directly - `fill` where the depth is out
// of range (the reference's pre-fill).  No sample list, no leaf-id buffer, no count on the host.
constexpr int FOREST_WARPS = 4;
struct ForestLayers {
    int L, sumC;
    int C[RSS_MAX_LAYERS];
    int coff[RSS_MAX_LAYERS];
};
__global__ void __launch_bounds__(32 * FOREST_WARPS, RSS_FOREST_MINB)
    forest_frame_lowres_kernel(const Node* __restrict__ nodes, const int* __restrict__ tree_off, int T,
                               const float* __restrict__ leaves, const ForestLayers fl, const FrameFeat F, int stride, int gw,
                               int gh, float dmin_mm, float dmax_mm, float fill, float* __restrict__ lowres) {
    __shared__ int srow[FOREST_WARPS][32];
    __shared__ unsigned svalid;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g0 = blockIdx.x * 32, g = g0 + lane, ncell = gw * gh;
    const int x = (g % gw) * stride, y = (g / gw) * stride;
    bool ok = false;
    if (g < ncell) {
        const float d = (float)F.depth[(size_t)y * F.W + x];  // sample selection, feature_extractor.h:43-44,62
        ok = d >= dmin_mm && d <= dmax_mm;
    }
    const unsigned valid = __ballot_sync(0xffffffffu, ok);
    if (threadIdx.x == 0) svalid = valid;
    // output element o of the CTA: layers back to back, inside a layer (position, class) - the order of the low-res image
    constexpr int NT = 32 * FOREST_WARPS;
    const int npos = min(32, ncell - g0), nout = npos * fl.sumC;
    float acc[8];  // ceil(32 * sumC / NT) <= 8 for sumC <= 32
    for (int t0 = 0; t0 < T; t0 += FOREST_WARPS) {
        const int t = t0 + w;
        if (t < T && ok) srow[w][lane] = traverse_frame(nodes + tree_off[t], F, x, y);
        __syncthreads();
        const int nt = min(FOREST_WARPS, T - t0);
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int o = threadIdx.x + k * NT;
            if (o >= nout) break;
            int l = 0;
            while (l + 1 < fl.L && o >= npos * fl.coff[l + 1]) l++;
            const int q = o - npos * fl.coff[l], pos = q / fl.C[l], cl = q - pos * fl.C[l];
            if (!((svalid >> pos) & 1u)) continue;
            float a = t0 == 0 ? 0.f : acc[k];
            for (int i = 0; i < nt; i++) {
                const float h = __ldg(leaves + (size_t)srow[i][pos] * fl.sumC + fl.coff[l] + cl);
                a = (t0 + i == 0) ? h : __fadd_rn(a, h);
            }
            acc[k] = a;
        }
        __syncthreads();
    }
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int o = threadIdx.x + k * NT;
        if (o >= nout) break;
        int l = 0;
        while (l + 1 < fl.L && o >= npos * fl.coff[l + 1]) l++;
        const int q = o - npos * fl.coff[l], pos = q / fl.C[l];
        lowres[(size_t)ncell * fl.coff[l] + (size_t)g0 * fl.C[l] + q] = ((svalid >> pos) & 1u) ? acc[k] : fill;
    }
}
void launch_forest_frame_lowres(rss_ctx* c, cudaStream_t st, const Node* nodes, const int* tree_off, int T, const float* leaves,
                                int L, const int* C, const uchar4* lab, const uint16_t* depth, const float4* xyz,
                                const float* dist, const double* integ, const int* cnt, const ResizeTap* tapx,
                                const ResizeTap* tapy, const uint16_t* feat_xy, int W, int H, int P, int r, int ncolor,
                                int pos_depth, int pos_height, int pos_normal, int stride, float dmin_mm, float dmax_mm,
                                float fill, float* lowres) {
    const int gw = rss_div_up(W, stride), gh = rss_div_up(H, stride);
    ForestLayers fl;
    fl.L = L;
    int off = 0;
    for (int l = 0; l < RSS_MAX_LAYERS; l++) {
        fl.C[l] = l < L ? C[l] : 0;
        fl.coff[l] = off;
        off += fl.C[l];
    }
    fl.sumC = off;
    const FrameFeat F{lab, depth, xyz, dist, integ, cnt, tapx, tapy, feat_xy, W, H, P, r, ncolor, pos_depth, pos_height, pos_normal};
    RSS_LAUNCH(c, forest_frame_lowres_kernel, rss_div_up((long long)gw * gh, 32), 32 * FOREST_WARPS, 0, st, nodes, tree_off, T,
               leaves, fl, F, stride, gw, gh, dmin_mm, dmax_mm, fill, lowres);
}

void launch_scalar_features(rss_ctx* c, cudaStream_t st, const uint16_t* depth, const float4* xyz,
                            const float* dist, const double* integ, const int* integ_cnt, int W, int H,
                            const int* xs, const int* ys, int n, float* feats, int D, int pos_depth,
                            int pos_height, int pos_normal) {
    if (n <= 0) return;
    RSS_LAUNCH(c, scalar_features_kernel, rss_div_up(n, 256), 256, 0, st, depth, xyz, dist, integ, integ_cnt, W, H,
               xs, ys, n, feats, D, pos_depth, pos_height, pos_normal);
}

}  // namespace rss
