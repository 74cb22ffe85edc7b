// DenseCRF mean-field inference on the device and the C ABI for it (include/rss.h, part 2).
// Reference: third-party/densecrf/src/densecrf.cpp:98-131 (expAndNormalize, inference),
// pairwise.cpp:40-80,173-178 (DenseKernel::initLattice/filter, PairwisePotential::apply),
// labelcompatibility.cpp:46-48 (Potts), src/segmenter.cpp:561-657 (unary accumulation, gated argmax).
//
// All label layers of a CRF share every lattice: the channels of the layers are concatenated per point
// ([N][Mtot]) and filtered together (the filter is channel-independent); the soft-max runs per layer.
// One mean-field iteration = per lattice {zero, splat, d+1 blurs} on its own stream, then ONE fused
// point-parallel kernel: slice of every lattice + post-scale by norm + Potts + unary + per-layer soft-max
// (+ the gated argmax on the last iteration).  Q never leaves the device between iterations.
#include <cmath>
#include <cstdlib>
#include <cstring>


#include "kernels.hpp"
#include "lattice.cuh"
#include "meanfield.cuh"

namespace rss {
float lattice_alpha(int d);
rss_status lattice_build(rss_ctx* ctx, cudaStream_t st, Lattice& L, const float* feat, int N, int d, uint32_t hcap, int Mp,
                         bool want_csr);
rss_status lattice_build_tile_csr(rss_ctx* ctx, cudaStream_t st, Lattice& L, int G, int grid_w, int grid_h, const int* perm = nullptr);
float* lattice_splat_blur(rss_ctx* ctx, cudaStream_t st, Lattice& L, const float* in, int in_stride, const float* norm,
                          int Mp, bool reverse = false);
void lattice_slice(rss_ctx* ctx, cudaStream_t st, Lattice& L, const float* values, int M, int Mp, int seq, float* out,
                   int out_stride);
rss_status lattice_normalization(rss_ctx* ctx, cudaStream_t st, Lattice& L);
void launch_run_count(rss_ctx* c, cudaStream_t st, Lattice& L);

constexpr int CRF_MAX_KERNELS = 4;
constexpr int CRF_MAX_CH = 32;

struct LayerSpec {
    int n_layers;
    int off[RSS_MAX_LAYERS + 1];  // first channel of the layer in the device row (a multiple of 4)
    int end[RSS_MAX_LAYERS];      // one past its last label (channels between end[l] and off[l + 1] are padding)
    int unknown[RSS_MAX_LAYERS];  // < 0: plain argmax
};
struct SliceArgs {
    int K;
    const int* offsets[CRF_MAX_KERNELS];
    const float* bary[CRF_MAX_KERNELS];
    const float* values[CRF_MAX_KERNELS];
    const float* norm[CRF_MAX_KERNELS];
    const uint32_t* counts[CRF_MAX_KERNELS];
    int d1[CRF_MAX_KERNELS];
    float alpha[CRF_MAX_KERNELS];
    float potts[CRF_MAX_KERNELS];
};

// ---------------------------------------------------------------------------------------------------------------
// Point-parallel kernels.  Q and the unary energies are stored [N][Mp] (Mp = labels padded to a multiple of 4, pad
// channels = 0) so that every row is 16-byte aligned.  A point is handled by a GROUP OF 8 LANES; lane g of the group
// owns the float4 channel group g (channels 4g..4g+3).  All row accesses of a group are one contiguous 16*G-byte
// read, and the per-layer soft-max / argmax reductions are three xor-shuffles inside the group.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float group_max(float v) {
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
    return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 4));
}
__device__ __forceinline__ float group_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    return v + __shfl_xor_sync(0xffffffffu, v, 4);
}
// expAndNormalize (densecrf.cpp:98-106) per layer; t = this lane's four channels, c0 = 4g.  Pad channels become 0.
__device__ __forceinline__ void softmax_layers(float (&t)[4], int c0, const LayerSpec& ls) {
    float out[4] = {0.f, 0.f, 0.f, 0.f};
    for (int l = 0; l < ls.n_layers; l++) {
        const int a = ls.off[l], b = ls.end[l];
        float mx = -INFINITY;
#pragma unroll
        for (int q = 0; q < 4; q++)
            if (c0 + q >= a && c0 + q < b) mx = fmaxf(mx, t[q]);
        mx = group_max(mx);
        float e[4], s = 0.f;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const bool in = c0 + q >= a && c0 + q < b;
            e[q] = in ? __expf(__fsub_rn(t[q], mx)) : 0.f;  // ex2.approx: |rel err| < 2e-6, tolerance is 1e-4 abs
            s += e[q];
        }
        s = group_sum(s);
        const float rs = __fdividef(1.0f, s);
#pragma unroll
        for (int q = 0; q < 4; q++)
            if (c0 + q >= a && c0 + q < b) out[q] = e[q] * rs;
    }
#pragma unroll
    for (int q = 0; q < 4; q++) t[q] = out[q];
}
// gated argmax (segmenter.cpp:645-657: first label whose Q exceeds 2/M and every earlier candidate, else the layer's
// "Unknown") or plain argmax (DenseCRF::currentMap, densecrf.cpp:200-208).  Ties resolve to the lower label, like the
// reference's strict '>' scan.
__device__ __forceinline__ void write_labels(const float (&q)[4], int c0, int g, const LayerSpec& ls, int i, int N,
                                             uint8_t* __restrict__ labels) {
    for (int l = 0; l < ls.n_layers; l++) {
        const int a = ls.off[l], b = ls.end[l];
        const bool gated = ls.unknown[l] >= 0;
        float bv = gated ? (float)(2.0 / (double)(b - a)) : -INFINITY;
        int best = 1 << 20;  // "no label passed the gate"
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (c0 + k >= a && c0 + k < b && q[k] > bv) {
                bv = q[k];
                best = c0 + k - a;
            }
#pragma unroll
        for (int o = 1; o <= 4; o <<= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int ob = __shfl_xor_sync(0xffffffffu, best, o);
            if (ob < (1 << 20) && (best >= (1 << 20) || ov > bv || (ov == bv && ob < best))) {
                bv = ov;
                best = ob;
            }
        }
        if (g == 0) labels[(size_t)l * N + i] = (uint8_t)(best < (1 << 20) ? best : (gated ? ls.unknown[l] : 0));
    }
}

// Q0 = expAndNormalize(-unary)   (densecrf.cpp:120)
__global__ void __launch_bounds__(256) softmax_init_kernel(const float* __restrict__ unary, int N, int G, int Mp,
                                                           LayerSpec ls, float* __restrict__ Q) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int i = (int)(gid >> 3), g = (int)(gid & 7);
    const bool live = i < N && g < G;
    float t[4] = {0.f, 0.f, 0.f, 0.f};
    if (live) {
        const float4 u = __ldg(reinterpret_cast<const float4*>(unary + (size_t)i * Mp) + g);
        t[0] = -u.x; t[1] = -u.y; t[2] = -u.z; t[3] = -u.w;
    }
    softmax_layers(t, 4 * g, ls);
    if (live) reinterpret_cast<float4*>(Q + (size_t)i * Mp)[g] = make_float4(t[0], t[1], t[2], t[3]);
}

// One mean-field update (densecrf.cpp:123-128), ONE THREAD PER POINT:
//   tmp = -unary;  for each kernel: tmp -= -w * (slice(values) * norm);  Q = expAndNormalize(tmp)
// With first-appearance vertex numbering neighbouring pixels reference the same few value rows, so the 32 row
// gathers of a warp hit a handful of cache lines (L1 resident).  All label channels of the point live in registers
// (MP = padded channel count, compile time), rows are read as float4.
template <int MP>
__device__ __forceinline__ void softmax_regs(float (&t)[MP], const LayerSpec& ls) {
    unsigned live = 0;  // channels that belong to a layer; the others (padding) come out as 0
    for (int l = 0; l < ls.n_layers; l++) live |= (ls.end[l] >= 32 ? 0xffffffffu : ((1u << ls.end[l]) - 1u)) & ~((1u << ls.off[l]) - 1u);
#pragma unroll
    for (int c = 0; c < MP; c++)
        if (!((live >> c) & 1u)) t[c] = 0.f;
    for (int l = 0; l < ls.n_layers; l++) {
        const int a = ls.off[l], b = ls.end[l];
        float mx = -INFINITY;
#pragma unroll
        for (int c = 0; c < MP; c++)
            if (c >= a && c < b) mx = fmaxf(mx, t[c]);
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < MP; c++)
            if (c >= a && c < b) {
                t[c] = __expf(__fsub_rn(t[c], mx));  // ex2.approx: |rel err| < 2e-6, the tolerance is 1e-4 abs
                s += t[c];
            }
        const float rs = __fdividef(1.0f, s);
#pragma unroll
        for (int c = 0; c < MP; c++)
            if (c >= a && c < b) t[c] *= rs;
    }
}
template <int MP, int D1>
__device__ __forceinline__ void slice_one(const SliceArgs& sa, int k, int i, int Mp, float (&t)[MP]) {
    float acc[MP];
#pragma unroll
    for (int c = 0; c < MP; c++) acc[c] = 0.f;
    const int* offp = sa.offsets[k] + (size_t)i * D1;
    const float* bp = sa.bary[k] + (size_t)i * D1;
    const float alpha = sa.alpha[k];
    int v[D1];
    float w[D1];
#pragma unroll
    for (int j = 0; j < D1; j++) {
        v[j] = __ldg(offp + j);
        w[j] = __fmul_rn(__ldg(bp + j), alpha);
    }
#pragma unroll
    for (int j = 0; j < D1; j++) {
        const float4* row = reinterpret_cast<const float4*>(sa.values[k] + (size_t)v[j] * Mp);
#pragma unroll
        for (int g = 0; g < MP / 4; g++) {
            const float4 x = __ldg(row + g);
            acc[4 * g + 0] = __fadd_rn(acc[4 * g + 0], __fmul_rn(w[j], x.x));
            acc[4 * g + 1] = __fadd_rn(acc[4 * g + 1], __fmul_rn(w[j], x.y));
            acc[4 * g + 2] = __fadd_rn(acc[4 * g + 2], __fmul_rn(w[j], x.z));
            acc[4 * g + 3] = __fadd_rn(acc[4 * g + 3], __fmul_rn(w[j], x.w));
        }
    }
    const float nv = sa.norm[k] ? __ldg(sa.norm[k] + i) : 1.f, negw = -sa.potts[k];
#pragma unroll
    for (int c = 0; c < MP; c++) {
        const float o = __fmul_rn(negw, __fmul_rn(acc[c], nv));  // pairwise.cpp:78-79, labelcompatibility.cpp:46-48
        t[c] = __fsub_rn(t[c], o);                               // densecrf.cpp:126
    }
}
template <int MP>
__global__ void __launch_bounds__(128) slice_softmax_kernel(SliceArgs sa, const float* __restrict__ unary, int N, int M,
                                                            LayerSpec ls, float* __restrict__ Q,
                                                            uint8_t* __restrict__ labels) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    float t[MP];
    const float4* up = reinterpret_cast<const float4*>(unary + (size_t)i * MP);
#pragma unroll
    for (int g = 0; g < MP / 4; g++) {
        const float4 u = __ldg(up + g);
        t[4 * g] = -u.x; t[4 * g + 1] = -u.y; t[4 * g + 2] = -u.z; t[4 * g + 3] = -u.w;
    }
    for (int k = 0; k < sa.K; k++) {
        if (sa.counts[k][1]) return;  // lattice overflow: the host re-runs with a larger table
        switch (sa.d1[k]) {
            case 2: slice_one<MP, 2>(sa, k, i, MP, t); break;
            case 3: slice_one<MP, 3>(sa, k, i, MP, t); break;
            case 4: slice_one<MP, 4>(sa, k, i, MP, t); break;
            case 5: slice_one<MP, 5>(sa, k, i, MP, t); break;
            case 6: slice_one<MP, 6>(sa, k, i, MP, t); break;
            case 7: slice_one<MP, 7>(sa, k, i, MP, t); break;
            default: slice_one<MP, 8>(sa, k, i, MP, t); break;
        }
    }
    softmax_regs<MP>(t, ls);
    float4* qp = reinterpret_cast<float4*>(Q + (size_t)i * MP);
#pragma unroll
    for (int g = 0; g < MP / 4; g++)
        qp[g] = make_float4(t[4 * g], t[4 * g + 1], t[4 * g + 2], t[4 * g + 3]);
    if (labels) {
        for (int l = 0; l < ls.n_layers; l++) {
            const int a = ls.off[l], b = ls.end[l];
            const bool gated = ls.unknown[l] >= 0;
            float bv = gated ? (float)(2.0 / (double)(b - a)) : -INFINITY;  // segmenter.cpp:647
            int best = gated ? ls.unknown[l] : 0;
#pragma unroll
            for (int c = 0; c < MP; c++)
                if (c >= a && c < b && t[c] > bv) {
                    bv = t[c];
                    best = c - a;
                }
            labels[(size_t)l * N + i] = (uint8_t)best;
        }
    }
}
template <int MP>
static void launch_slice_softmax(rss_ctx* c, cudaStream_t st, const SliceArgs& sa, const float* unary, int N, int M,
                                 const LayerSpec& ls, float* Q, uint8_t* labels) {
    RSS_LAUNCH(c, slice_softmax_kernel<MP>, rss_div_up(N, 128), 128, 0, st, sa, unary, N, M, ls, Q, labels);
}

__global__ void __launch_bounds__(256) argmax_kernel(const float* __restrict__ Q, int N, int G, int Mp, LayerSpec ls,
                                                     uint8_t* __restrict__ labels) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int i = (int)(gid >> 3), g = (int)(gid & 7);
    float q[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    if (i < N && g < G) {
        const float4 u = __ldg(reinterpret_cast<const float4*>(Q + (size_t)i * Mp) + g);
        q[0] = u.x; q[1] = u.y; q[2] = u.z; q[3] = u.w;
    }
    if (i < N) write_labels(q, 4 * g, g, ls, i, N, labels);
}

// layout helpers between the reference's per-layer matrices ([N][M_l]) and the padded interleaved device layout
// host: rows of hstride floats, the layer's Ml labels at column hoff; device: rows of Mp floats, the layer at column off
__global__ void __launch_bounds__(256) interleave_kernel(const float* __restrict__ src, int N, int Ml, int hstride, int hoff,
                                                         int Mp, int off, float scale, float* __restrict__ dst) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)N * Ml) return;
    const int i = (int)(gid / Ml), c = (int)(gid - (long long)i * Ml);
    dst[(size_t)i * Mp + off + c] = scale * src[(size_t)i * hstride + hoff + c];
}
__global__ void __launch_bounds__(256) deinterleave_kernel(const float* __restrict__ src, int N, int Ml, int Mp, int off,
                                                           int hstride, int hoff, float* __restrict__ dst) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)N * Ml) return;
    const int i = (int)(gid / Ml), c = (int)(gid - (long long)i * Ml);
    dst[(size_t)i * hstride + hoff + c] = src[(size_t)i * Mp + off + c];
}
// unary rows start as energy 0 on the label channels and +inf on the padding channels of every layer (bit c of live:
// channel c is a label), so that exp(-unary - max) of a padding channel is 0 in every soft-max
__global__ void __launch_bounds__(256) unary_init_kernel(float* __restrict__ unary, size_t n, int Mp, unsigned live) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i < n; i += (size_t)gridDim.x * blockDim.x) unary[i] = ((live >> (unsigned)(i % Mp)) & 1u) ? 0.f : INFINITY;
}
// segmenter.cpp:597-616: unaries(c, idx) += posterior[pixel*C + c] for idx >= 0.  The CRF stores ENERGIES
// (= -unaries, segmenter.cpp:642), so the log-posterior is subtracted.
__global__ void __launch_bounds__(256) unary_accumulate_kernel(const int* __restrict__ index_image, int npix,
                                                               const float* __restrict__ post, int Ml, int Mp, int off,
                                                               int N, float* __restrict__ unary) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)npix * Ml) return;
    const int p = (int)(gid / Ml), c = (int)(gid - (long long)p * Ml);
    const int idx = index_image[p];
    if (idx < 0 || idx >= N) return;
    atomicAdd(unary + (size_t)idx * Mp + off + c, -post[gid]);
}
// The projector of src/segmenter.cpp:581 (fps_mapper's pinhole projector, un-vendored) as a z-buffer kernel: every
// point of the resident cloud goes to the camera frame, p_cam = R^T (p - t), is projected with K to the nearest pixel
// and competes for it with key = (float bits of z) << 32 | index through one 64-bit atomicMin (z > 0, so the bit pattern
// orders like the value): the nearest point wins, equal depths resolve to the lower index - the rule of the oracle's
// serial loop (orc_project_zbuffer), with the same float operations in the same order.
struct ProjParams {
    float R[9], t[3], fx, fy, cx, cy, zmin, zmax;
};
__global__ void __launch_bounds__(256) project_zbuffer_kernel(const float* __restrict__ xyz, int N, ProjParams P, int W, int H,
                                                              unsigned long long* __restrict__ zbuf) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const float dx = __fsub_rn(xyz[3 * (size_t)i], P.t[0]), dy = __fsub_rn(xyz[3 * (size_t)i + 1], P.t[1]),
                dz = __fsub_rn(xyz[3 * (size_t)i + 2], P.t[2]);
    const float x = __fadd_rn(__fadd_rn(__fmul_rn(P.R[0], dx), __fmul_rn(P.R[3], dy)), __fmul_rn(P.R[6], dz));
    const float y = __fadd_rn(__fadd_rn(__fmul_rn(P.R[1], dx), __fmul_rn(P.R[4], dy)), __fmul_rn(P.R[7], dz));
    const float z = __fadd_rn(__fadd_rn(__fmul_rn(P.R[2], dx), __fmul_rn(P.R[5], dy)), __fmul_rn(P.R[8], dz));
    if (!(z >= P.zmin && z <= P.zmax)) return;
    const float iz = __fdiv_rn(1.0f, z);
    const float u = __fadd_rn(__fmul_rn(__fmul_rn(P.fx, x), iz), P.cx), v = __fadd_rn(__fmul_rn(__fmul_rn(P.fy, y), iz), P.cy);
    const float fu = floorf(__fadd_rn(u, 0.5f)), fv = floorf(__fadd_rn(v, 0.5f));
    if (!(fu >= 0.f && fu < (float)W && fv >= 0.f && fv < (float)H)) return;
    const unsigned long long key = ((unsigned long long)__float_as_uint(z) << 32) | (unsigned)i;
    atomicMin(zbuf + (size_t)(int)fv * W + (int)fu, key);
}
__global__ void __launch_bounds__(256) zbuffer_to_index_kernel(const unsigned long long* __restrict__ zbuf, int npix,
                                                               int* __restrict__ index_image) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npix) return;
    const unsigned long long k = zbuf[p];
    index_image[p] = k == ~0ull ? -1 : (int)(unsigned)(k & 0xffffffffull);
}
// Sorting an incoherent point set for the fused path: key = (vertex of simplex corner 0, vertex of corner 1), so that
// consecutive sorted points share most of their lattice vertices
__global__ void __launch_bounds__(256) sort_key_kernel(const int* __restrict__ offsets, int N, int d1, uint64_t* __restrict__ keys,
                                                       uint32_t* __restrict__ vals) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= N) return;
    const uint32_t v0 = (uint32_t)offsets[(size_t)p * d1], v1 = d1 > 1 ? (uint32_t)offsets[(size_t)p * d1 + 1] : 0u;
    keys[p] = ((uint64_t)v0 << 32) | v1;
    vals[p] = (uint32_t)p;
}
// the run statistic of run_count_kernel (lattice.cu) over the sorted order: counts[6] += (position, corner) pairs whose
// vertex differs from the same corner of the previous sorted point
__global__ void __launch_bounds__(256) run_count_perm_kernel(const int* __restrict__ offsets, const int* __restrict__ perm, int N,
                                                             int d1, uint32_t* __restrict__ counts) {
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool differs = false;
    if (k < (size_t)N * d1) {
        const size_t q = k / d1, j = k - q * d1;
        differs = q == 0 || offsets[(size_t)perm[q] * d1 + j] != offsets[(size_t)perm[q - 1] * d1 + j];
    }
    const unsigned m = __ballot_sync(0xffffffffu, differs);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(counts + 6, (uint32_t)__popc(m));
}
// rows of Mp floats in sorted order: dst[q] = src[perm[q]]
__global__ void __launch_bounds__(256) gather_rows_kernel(const float4* __restrict__ src, const int* __restrict__ perm, int N, int G,
                                                          float4* __restrict__ dst) {
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= (size_t)N * G) return;
    const size_t q = k / G, g = k - q * G;
    dst[k] = src[(size_t)perm[q] * G + g];
}
// feature builders: DenseCRF2D::addPairwiseGaussian / Bilateral (densecrf.cpp:61-81), segmenter.cpp:629-637
__global__ void __launch_bounds__(256) feat_gaussian_kernel(int W, int H, float sx, float sy, float* __restrict__ f) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= W * H) return;
    const int i = p % W, j = p / W;
    f[2 * (size_t)p] = __fdiv_rn((float)i, sx);
    f[2 * (size_t)p + 1] = __fdiv_rn((float)j, sy);
}
__global__ void __launch_bounds__(256) feat_bilateral_kernel(int W, int H, float sx, float sy, float sr, float sg, float sb,
                                                             const uint8_t* __restrict__ im, float* __restrict__ f) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= W * H) return;
    const int i = p % W, j = p / W;
    f[5 * (size_t)p] = __fdiv_rn((float)i, sx);
    f[5 * (size_t)p + 1] = __fdiv_rn((float)j, sy);
    f[5 * (size_t)p + 2] = __fdiv_rn((float)im[3 * (size_t)p], sr);
    f[5 * (size_t)p + 3] = __fdiv_rn((float)im[3 * (size_t)p + 1], sg);
    f[5 * (size_t)p + 4] = __fdiv_rn((float)im[3 * (size_t)p + 2], sb);
}
__global__ void __launch_bounds__(256) feat_xyzrgb_kernel(int N, const float* __restrict__ xyz, const float* __restrict__ rgb,
                                                          float wxyz, float wrgb, float* __restrict__ f) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= N) return;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        f[6 * (size_t)p + k] = __fmul_rn(xyz[3 * (size_t)p + k], wxyz);
        f[6 * (size_t)p + 3 + k] = __fmul_rn(rgb[3 * (size_t)p + k], wrgb);
    }
}
// keyframe features from the frame's own buffers: back-projected points / sigma (invalid depth -> NaN is replaced
// by the camera centre, i.e. depth 0), and (x, y, r, g, b) / sigma
__global__ void __launch_bounds__(256) feat_frame_xyz_kernel(int N, const float4* __restrict__ xyz, float inv_sigma,
                                                             const PoseParams* __restrict__ pose, float* __restrict__ f) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= N) return;
    float4 v = xyz[p];
    if (isnan(v.x)) { v.x = pose->t[0]; v.y = pose->t[1]; v.z = pose->t[2]; }
    f[3 * (size_t)p] = __fmul_rn(v.x, inv_sigma);
    f[3 * (size_t)p + 1] = __fmul_rn(v.y, inv_sigma);
    f[3 * (size_t)p + 2] = __fmul_rn(v.z, inv_sigma);
}
static LayerSpec make_layers(const rss_crf* crf, const int* unknown) {
    LayerSpec ls;
    ls.n_layers = crf->n_layers;
    for (int l = 0; l <= RSS_MAX_LAYERS; l++) ls.off[l] = l <= crf->n_layers ? crf->moff[l] : crf->Mp;
    for (int l = 0; l < RSS_MAX_LAYERS; l++) ls.end[l] = l < crf->n_layers ? crf->moff[l] + crf->M[l] : crf->Mp;
    for (int l = 0; l < RSS_MAX_LAYERS; l++) ls.unknown[l] = (unknown && l < crf->n_layers) ? unknown[l] : -1;
    return ls;
}

// ---------------------------------------------------------------------------------------------------------------
// Fused path (meanfield.cu): K <= 2 lattices over spatially coherent points.  n iterations = n + 1 point kernels
// (Q0 + splat | slice + soft-max + splat | ... | slice + soft-max + labels) and n cooperative blur launches.
// ---------------------------------------------------------------------------------------------------------------
static bool crf_fused_order(const rss_crf* crf, int* first, int* second) {
    const int K = (int)crf->kernels.size();
    if (K < 1 || K > FUSED_MAX_LAT) return false;
    const int G = crf->Mp / 4;
    const TileMap tm = fused_tile_map(crf->N, crf->grid_w, crf->grid_h);
    for (const Lattice* L : crf->kernels)
        if (!L->ordered || L->tile_TP != tm.TP || L->tile_W != tm.W) return false;
    if (K == 1) {
        *first = 0; *second = -1;
        return fused_signature_supported(G, crf->kernels[0]->d + 1, 0);
    }
    const int a = crf->kernels[0]->d + 1, b = crf->kernels[1]->d + 1;
    if (fused_signature_supported(G, a, b)) { *first = 0; *second = 1; return true; }
    if (fused_signature_supported(G, b, a)) { *first = 1; *second = 0; return true; }
    return false;
}
static rss_status crf_run_fused(rss_crf* crf, int iters, const int* unknown, uint8_t* labels_dev, int i0, int i1) {
    rss_ctx* ctx = crf->ctx;
    cudaStream_t s0 = ctx->s0;
    const int N = crf->N, Mp = crf->Mp, G = Mp / 4;
    Lattice* Ls[FUSED_MAX_LAT] = {crf->kernels[i0], i1 >= 0 ? crf->kernels[i1] : nullptr};
    const int K = i1 >= 0 ? 2 : 1;
    const LayerSpec ls0 = make_layers(crf, unknown);
    FusedLayers ls;
    ls.n_layers = ls0.n_layers;
    for (int l = 0; l < RSS_MAX_LAYERS; l++) {
        const int Ml = l < crf->n_layers ? crf->M[l] : 0;
        ls.off[l] = ls0.off[l];
        ls.count[l] = Ml;
        ls.gmask[l] = 0;
        for (int g = ls0.off[l] / 4; l < crf->n_layers && g < (ls0.off[l] + Ml + 3) / 4; g++) ls.gmask[l] |= 1u << g;
        ls.unknown[l] = ls0.unknown[l];
        ls.gate[l] = (ls0.unknown[l] >= 0 && Ml > 0) ? (float)(2.0 / (double)Ml) : -INFINITY;  // segmenter.cpp:647
    }
    DevBuf* tgt[FUSED_MAX_LAT];
    DevBuf* spare[FUSED_MAX_LAT];
    DevBuf* res[FUSED_MAX_LAT];
    FusedArgs fa;
    BlurMultiArgs ba;
    ba.K = K;
    for (int k = 0; k < FUSED_MAX_LAT; k++) {
        if (k >= K) {
            fa.lat[k] = FusedLat{};
            continue;
        }
        Lattice& L = *Ls[k];
        tgt[k] = L.splat_target ? &L.val_b : &L.val_a;
        spare[k] = L.splat_target ? &L.val_a : &L.val_b;
        res[k] = &L.val_c;
        FusedLat& f = fa.lat[k];
        f.pt_w = L.tile_pt_w.as<float>(); f.pt_slot = L.tile_pt_slot.as<uint16_t>();
        f.pairs = L.tile_pairs.as<uint2>(); f.ent_meta = L.tile_ent_meta.as<int2>();
        f.tile_vert = L.tile_vert.as<int>(); f.tile_info = L.tile_info.as<int2>();
        f.counts = L.counts.as<uint32_t>();
        ba.nbr[k] = L.nbr.as<int2>(); ba.counts[k] = L.counts.as<uint32_t>(); ba.d1[k] = L.d + 1; ba.vcap[k] = L.vcap;
    }
    const int d1a = Ls[0]->d + 1, d1b = K > 1 ? Ls[1]->d + 1 : 0;
    int phases_of[FUSED_MAX_LAT] = {0};
    const int blur_phases = blur_multi_plan(ba, G, phases_of);
    const BlurShape blur_shape = blur_multi_shape(ctx);  // once per inference: the barrier targets depend on the grid
    const float* U = crf->unary.as<float>();  // (replaced by the sorted copy for incoherent point sets, below)
    float* Q = crf->Q.as<float>();
    Lattice& L0 = *Ls[0];
    TileMap tm = fused_tile_map(N, crf->grid_w, crf->grid_h);
    if (crf->sorted && tm.W == 0) {
        // incoherent point set: tiles run over the sorted order; the unary rows are copied into that order once per
        // inference (one gather), Q and the labels are written through the permutation by the last point kernel
        tm.perm = crf->perm.as<int>();
        RSS_CU(ctx, crf->unary_sorted.reserve((size_t)N * Mp * 4));
        RSS_LAUNCH(ctx, gather_rows_kernel, rss_div_up((long long)N * G, 256), 256, 0, s0, crf->unary.as<float4>(), tm.perm, N, G,
                   crf->unary_sorted.as<float4>());
        U = crf->unary_sorted.as<float>();
    }
    // pass 0: Q0 = expAndNormalize(-unary) and its splat
    for (int k = 0; k < K; k++) { fa.lat[k].vin = nullptr; fa.lat[k].vout = tgt[k]->as<float>(); }
    RSS_CU(ctx, launch_meanfield_fused(ctx, s0, fa, d1a, d1b, U, Q, nullptr, tm, G, ls, 2));
    for (int it = 0; it < iters; it++) {
        for (int k = 0; k < K; k++) {
            ba.ping[k] = tgt[k]->as<float4>(); ba.pong[k] = spare[k]->as<float4>(); ba.zero[k] = res[k]->as<float4>();
        }
        RSS_CU(ctx, launch_blur_multi(ctx, s0, ba, G, L0.counts.as<unsigned int>() + 8, L0.barrier_base, blur_shape));
        L0.barrier_base += (unsigned int)(blur_phases - 1) * (unsigned int)blur_shape.grid;
        for (int k = 0; k < K; k++) {
            DevBuf* X = (phases_of[k] % 2 == 0) ? tgt[k] : spare[k];  // blurred result
            DevBuf* Y = (phases_of[k] % 2 == 0) ? spare[k] : tgt[k];
            DevBuf* Z = res[k];                                    // cleared by the blur kernel: next splat target
            res[k] = X; spare[k] = Y; tgt[k] = Z;
            fa.lat[k].vin = res[k]->as<float>();
            fa.lat[k].vout = tgt[k]->as<float>();
        }
        const bool last = it == iters - 1;
        // the last pass writes the marginals unless the caller only wants the label maps (rss_segment_keyframe with Q = NULL)
        float* Qlast = (crf->skip_q_store && labels_dev) ? nullptr : Q;
        RSS_CU(ctx, launch_meanfield_fused(ctx, s0, fa, d1a, d1b, U, last ? Qlast : Q, last ? labels_dev : nullptr, tm, G, ls,
                                           last ? (1 | 4) : (1 | 2)));
    }
    // restore the two-table convention of the generic path: val_a = the all-zero table, splat_target = 0
    for (int k = 0; k < K; k++) {
        Lattice& L = *Ls[k];
        const DevBuf t = *tgt[k], r = *res[k], sp = *spare[k];
        L.val_a = t; L.val_b = r; L.val_c = sp;
        L.splat_target = 0;
    }
    RSS_CU(ctx, cudaGetLastError());
    return RSS_OK;
}

// the mean-field loop, fully enqueued (no host synchronisation).  labels_dev may be NULL.
// init: start from Q0 = expAndNormalize(-unary) (DenseCRF::inference / startInference); otherwise continue from the
// resident Q (DenseCRF::stepInference).
rss_status crf_run(rss_crf* crf, int iters, const int* unknown, uint8_t* labels_dev, bool init = true) {
    rss_ctx* ctx = crf->ctx;
    cudaStream_t s0 = ctx->s0;
    const int N = crf->N, Mp = crf->Mp, K = (int)crf->kernels.size();
    const LayerSpec ls = make_layers(crf, unknown);
    float* Q = crf->Q.as<float>();
    const float* U = crf->unary.as<float>();
    const int G = Mp / 4;
    int fi0 = 0, fi1 = -1;
    // the fused path computes Q0 = expAndNormalize(-unary) inside its first point kernel
    if (init && iters > 0 && K > 0 && crf_fused_order(crf, &fi0, &fi1))
        return crf_run_fused(crf, iters, unknown, labels_dev, fi0, fi1);
    if (init) RSS_LAUNCH(ctx, softmax_init_kernel, rss_div_up((long long)N * 8, 256), 256, 0, s0, U, N, G, Mp, ls, Q);
    if (iters <= 0 || K == 0) {
        // no pairwise terms: every iteration reproduces expAndNormalize(-unary), i.e. Q0
        if (!init && K == 0 && iters > 0)
            RSS_LAUNCH(ctx, softmax_init_kernel, rss_div_up((long long)N * 8, 256), 256, 0, s0, U, N, G, Mp, ls, Q);
        if (labels_dev)
            RSS_LAUNCH(ctx, argmax_kernel, rss_div_up((long long)N * 8, 256), 256, 0, s0, (const float*)Q, N, G, Mp, ls, labels_dev);
        RSS_CU(ctx, cudaGetLastError());
        return RSS_OK;
    }
    for (int it = 0; it < iters; it++) {
        SliceArgs sa;
        sa.K = K;
        RSS_CU(ctx, cudaEventRecord(crf->ev_fork, s0));
        for (int k = 0; k < K; k++) {
            Lattice& L = *crf->kernels[k];
            cudaStream_t sk = K > 1 ? crf->side[k] : s0;
            if (K > 1) RSS_CU(ctx, cudaStreamWaitEvent(sk, crf->ev_fork, 0));
            const bool pre = L.norm_type == RSS_NORMALIZE_SYMMETRIC || L.norm_type == RSS_NORMALIZE_BEFORE;
            const bool post = L.norm_type == RSS_NORMALIZE_SYMMETRIC || L.norm_type == RSS_NORMALIZE_AFTER;
            float* vals = lattice_splat_blur(ctx, sk, L, Q, Mp, pre ? L.norm.as<float>() : nullptr, Mp);
            if (!vals) return ctx->fail(RSS_ERR_CUDA, std::string("cooperative blur launch: ") + cudaGetErrorString(cudaGetLastError()));
            if (K > 1) RSS_CU(ctx, cudaEventRecord(crf->ev_join[k], sk));
            sa.offsets[k] = L.offsets.as<int>();
            sa.bary[k] = L.bary.as<float>();
            sa.values[k] = vals;
            sa.norm[k] = post ? L.norm.as<float>() : nullptr;
            sa.counts[k] = L.counts.as<uint32_t>();
            sa.d1[k] = L.d + 1;
            sa.alpha[k] = lattice_alpha(L.d);
            sa.potts[k] = L.potts_w;
        }
        if (K > 1)
            for (int k = 0; k < K; k++) RSS_CU(ctx, cudaStreamWaitEvent(s0, crf->ev_join[k], 0));
        uint8_t* lab = (it == iters - 1) ? labels_dev : (uint8_t*)nullptr;
        switch (Mp) {
            case 4: launch_slice_softmax<4>(ctx, s0, sa, U, N, crf->Mtot, ls, Q, lab); break;
            case 8: launch_slice_softmax<8>(ctx, s0, sa, U, N, crf->Mtot, ls, Q, lab); break;
            case 12: launch_slice_softmax<12>(ctx, s0, sa, U, N, crf->Mtot, ls, Q, lab); break;
            case 16: launch_slice_softmax<16>(ctx, s0, sa, U, N, crf->Mtot, ls, Q, lab); break;
            case 20: launch_slice_softmax<20>(ctx, s0, sa, U, N, crf->Mtot, ls, Q, lab); break;
            case 24: launch_slice_softmax<24>(ctx, s0, sa, U, N, crf->Mtot, ls, Q, lab); break;
            case 28: launch_slice_softmax<28>(ctx, s0, sa, U, N, crf->Mtot, ls, Q, lab); break;
            default: launch_slice_softmax<32>(ctx, s0, sa, U, N, crf->Mtot, ls, Q, lab); break;
        }
    }
    RSS_CU(ctx, cudaGetLastError());
    return RSS_OK;
}

static uint32_t next_pow2(uint64_t v) {
    uint32_t p = 1;
    while ((uint64_t)p < v && p < (1u << 30)) p <<= 1;
    return p;
}

// sorted position -> point of an incoherent point set (crf->perm): one 64-bit radix sort of (vertex 0, vertex 1) keys
static rss_status crf_sort_points(rss_crf* crf, cudaStream_t st, Lattice& L) {
    rss_ctx* ctx = crf->ctx;
    const int N = crf->N, d1 = L.d + 1;
    RSS_CU(ctx, crf->sort_keys.reserve((size_t)N * 8));
    RSS_CU(ctx, crf->sort_keys2.reserve((size_t)N * 8));
    RSS_CU(ctx, crf->sort_vals.reserve((size_t)N * 4));
    RSS_CU(ctx, crf->perm.reserve((size_t)N * 4));
    RSS_LAUNCH(ctx, sort_key_kernel, rss_div_up(N, 256), 256, 0, st, L.offsets.as<int>(), N, d1, crf->sort_keys.as<uint64_t>(),
               crf->sort_vals.as<uint32_t>());
    int vbits = 1;
    while (vbits < 32 && (1ull << vbits) <= (uint64_t)L.vcap) vbits++;
    RSS_CU(ctx, sort_pairs_u64(st, crf->sort_tmp, crf->sort_keys.as<uint64_t>(), crf->sort_keys2.as<uint64_t>(),
                               crf->sort_vals.as<uint32_t>(), crf->perm.as<uint32_t>(), (size_t)N, 32 + vbits));
    return RSS_OK;
}
// is the lattice coherent over the sorted order?  (same rule as for the points' own order: on average a vertex run
// spans more than two points)
static rss_status crf_sorted_run_stat(rss_crf* crf, cudaStream_t st, Lattice& L, uint64_t maxv, bool* ordered) {
    rss_ctx* ctx = crf->ctx;
    uint32_t* counts = L.counts.as<uint32_t>();
    RSS_CU(ctx, cudaMemsetAsync(counts + 6, 0, 4, st));
    RSS_LAUNCH(ctx, run_count_perm_kernel, rss_div_up((long long)maxv, 256), 256, 0, st, L.offsets.as<int>(), crf->perm.as<int>(),
               crf->N, L.d + 1, counts);
    uint32_t runs = 0;
    RSS_CU(ctx, cudaMemcpyAsync(&runs, counts + 6, 4, cudaMemcpyDeviceToHost, st));
    RSS_CU(ctx, cudaStreamSynchronize(st));
    L.runs = (long long)runs;
    *ordered = 2 * (uint64_t)runs <= maxv;
    return RSS_OK;
}

// Builds one more lattice from device-resident features on stream `st`.
// sync = true : waits, checks the overflow flag and regrows the hash table until it fits (reference-shaped API).
// sync = false: only enqueues; the caller checks crf_lattice_overflow() later (keyframe path, no host sync).
// raster: the caller guarantees image raster order (keyframe path) -> the fused mean-field path is used without
// measuring the run statistic.
rss_status crf_add_kernel_dev(rss_crf* crf, cudaStream_t st, const float* feat_dev, int d, float potts_w, int norm_type,
                              bool sync, bool raster = false) {
    rss_ctx* ctx = crf->ctx;
    if ((int)crf->kernels.size() >= CRF_MAX_KERNELS) return ctx->fail(RSS_ERR_INVALID, "too many pairwise terms (max 4)");
    if (norm_type < RSS_NO_NORMALIZATION || norm_type > RSS_NORMALIZE_SYMMETRIC)
        return ctx->fail(RSS_ERR_INVALID, "unsupported normalization type");
    Lattice* L;
    if (!crf->pool.empty()) {  // rebuild in place: buffers (and the capacity that worked last time) are reused
        L = crf->pool.back();
        crf->pool.pop_back();
    } else {
        L = new Lattice();
    }
    L->potts_w = potts_w;
    L->norm_type = norm_type;
    crf->kernels.push_back(L);
    // a failed build (allocation, capacity) must not leave a half-built lattice registered: later inference / filter calls
    // would launch on null or uninitialised buffers.  The buffers go back to the pool.
    auto fail_build = [&](rss_status rc) {
        crf->kernels.pop_back();
        L->tile_TP = 0; L->have_csr = false; L->ordered = false;
        crf->pool.push_back(L);
        return rc;
    };
    const uint64_t maxv = (uint64_t)crf->N * (d + 1);
    uint32_t hcap = next_pow2(std::min<uint64_t>(2 * maxv, 1u << 17));
    if (L->hcap > hcap && L->d == d) hcap = L->hcap;
    if (L->want_hcap > hcap) hcap = L->want_hcap;
    // test knob: start from a deliberately undersized table so that the overflow -> regrow -> re-run path executes
    if (const char* dbg = getenv("RSS_DEBUG_HCAP")) {
        const long v = atol(dbg);
        if (v >= 64 && L->want_hcap == 0) hcap = next_pow2((uint64_t)v);
    }
    for (;;) {
        // raster order: the tile kernel does the splat, the vertex-major CSR of the generic path is not needed
        const bool tile_ok = fused_group_supported(crf->Mp / 4);
        rss_status rc = lattice_build(ctx, st, *L, feat_dev, crf->N, d, hcap, crf->Mp, !(raster && tile_ok));
        if (rc != RSS_OK) return fail_build(rc);
        rc = lattice_normalization(ctx, st, *L);
        if (rc != RSS_OK) return fail_build(rc);
        L->ordered = raster;
        if (raster && tile_ok) {
            rc = lattice_build_tile_csr(ctx, st, *L, crf->Mp / 4, crf->grid_w, crf->grid_h);
            if (rc != RSS_OK) return fail_build(rc);
        }
        if (!sync) return RSS_OK;
        launch_run_count(ctx, st, *L);
        uint32_t h[8];
        RSS_CU(ctx, cudaMemcpyAsync(h, L->counts.ptr, sizeof(h), cudaMemcpyDeviceToHost, st));
        RSS_CU(ctx, cudaStreamSynchronize(st));
        if (!h[1]) {
            L->V_host = (int)h[0];
            L->runs = (long long)h[5];
            // coherent point order: on average every vertex run spans more than two points
            L->ordered = raster || 2 * L->runs <= (long long)maxv;
            const bool grid = crf->grid_w > 0 && (long long)crf->grid_w * crf->grid_h == crf->N;
            if (crf->sorted && !grid) {
                // the CRF's points already run in sorted order on the fused path: this lattice joins if it is coherent there
                L->ordered = false;
                rc = crf_sorted_run_stat(crf, st, *L, maxv, &L->ordered);
                if (rc != RSS_OK) return fail_build(rc);
                if (L->ordered && tile_ok) return lattice_build_tile_csr(ctx, st, *L, crf->Mp / 4, 0, 0, crf->perm.as<int>());
                return RSS_OK;
            }
            // (not for NO_NORMALIZATION: un-normalised messages saturate the marginals and amplify the run-to-run float noise of
            // the fused path's atomic splat beyond the 1e-4 bar; that type is never used on the reference's path)
            if (!L->ordered && tile_ok && !grid && crf->kernels.size() == 1 && norm_type != RSS_NO_NORMALIZATION &&
                !getenv("RSS_NO_POINT_SORT")) {
                // an incoherent point set (a local map): sort the points by their first two lattice vertices and take the
                // fused path over the sorted order when that makes consecutive points share their vertices
                rc = crf_sort_points(crf, st, *L);
                if (rc != RSS_OK) return fail_build(rc);
                rc = crf_sorted_run_stat(crf, st, *L, maxv, &L->ordered);
                if (rc != RSS_OK) return fail_build(rc);
                crf->sorted = L->ordered;
                if (L->ordered) return lattice_build_tile_csr(ctx, st, *L, crf->Mp / 4, 0, 0, crf->perm.as<int>());
                return RSS_OK;
            }
            if (L->ordered && tile_ok && L->tile_TP == 0) return lattice_build_tile_csr(ctx, st, *L, crf->Mp / 4, crf->grid_w, crf->grid_h);
            return RSS_OK;
        }
        if ((uint64_t)hcap >= 2 * next_pow2(2 * maxv) || hcap >= (1u << 30))
            return fail_build(ctx->fail(RSS_ERR_CAPACITY, "lattice hash table overflow"));
        hcap = (uint32_t)std::min<uint64_t>((uint64_t)hcap * 8, 1u << 30);
    }
}
// after a synchronisation: true when some lattice overflowed; its next build will use a 4x larger table
rss_status crf_lattice_overflow(rss_crf* crf, bool* any) {
    rss_ctx* ctx = crf->ctx;
    *any = false;
    for (Lattice* L : crf->kernels) {
        uint32_t h[2];
        RSS_CU(ctx, cudaMemcpy(h, L->counts.ptr, sizeof(h), cudaMemcpyDeviceToHost));
        L->V_host = (int)h[0];
        if (h[1]) {
            const uint64_t maxv = (uint64_t)crf->N * (L->d + 1);
            if ((uint64_t)L->hcap >= 2 * next_pow2(2 * maxv) || L->hcap >= (1u << 30))
                return ctx->fail(RSS_ERR_CAPACITY, "lattice hash table overflow");
            L->want_hcap = L->hcap * 4;
            *any = true;
        }
    }
    return RSS_OK;
}

void crf_free(rss_crf* crf) {
    if (!crf) return;
    for (Lattice* L : crf->kernels) { L->release(); delete L; }
    for (Lattice* L : crf->pool) { L->release(); delete L; }
    crf->kernels.clear();
    crf->pool.clear();
    crf->unary.release(); crf->Q.release(); crf->scratch.release(); crf->labels.release(); crf->feat_stage.release();
    crf->perm.release(); crf->unary_sorted.release(); crf->sort_keys.release(); crf->sort_keys2.release();
    crf->sort_vals.release(); crf->sort_tmp.release();
    crf->cloud_xyz.release(); crf->cloud_rgb.release(); crf->zbuf.release(); crf->index_dev.release();
    for (int k = 0; k < 4; k++) {
        if (crf->side[k]) cudaStreamDestroy(crf->side[k]);
        if (crf->ev_join[k]) cudaEventDestroy(crf->ev_join[k]);
    }
    if (crf->ev_fork) cudaEventDestroy(crf->ev_fork);
    delete crf;
}
void crf_release_cached(rss_ctx* ctx) {
    if (ctx->keyframe_crf) crf_free(ctx->keyframe_crf);
    ctx->keyframe_crf = nullptr;
}

static void unary_clear(rss_crf* crf) {
    unsigned live = 0;
    for (int l = 0; l < crf->n_layers; l++)
        for (int c = crf->moff[l]; c < crf->moff[l] + crf->M[l]; c++) live |= 1u << c;
    const size_t n = (size_t)crf->N * crf->Mp;
    rss_ctx* ctx = crf->ctx;
    RSS_LAUNCH(ctx, unary_init_kernel, (int)std::min<size_t>((n + 255) / 256, 4096), 256, 0, ctx->s0, crf->unary.as<float>(), n,
               crf->Mp, live);
}
rss_status crf_new(rss_ctx* ctx, int N, int n_layers, const int* M, rss_crf** out) {
    if (!ctx || !out) return RSS_ERR_INVALID;
    if (N < 1 || n_layers < 1 || n_layers > RSS_MAX_LAYERS || !M) return ctx->fail(RSS_ERR_INVALID, "bad CRF dimensions");
    rss_crf* crf = new rss_crf();
    crf->ctx = ctx; crf->N = N; crf->n_layers = n_layers;
    int off = 0, hoff = 0;
    for (int l = 0; l < n_layers; l++) {
        if (M[l] < 1) { delete crf; return ctx->fail(RSS_ERR_INVALID, "layer with no labels"); }
        crf->M[l] = M[l];
        crf->moff[l] = off;
        crf->hoff[l] = hoff;
        off += (M[l] + 3) / 4 * 4;  // every layer starts on a float4 group
        hoff += M[l];
    }
    for (int l = n_layers; l <= RSS_MAX_LAYERS; l++) { crf->moff[l] = off; crf->hoff[l] = hoff; }
    crf->Mtot = hoff;
    crf->Mp = off;
    if (crf->Mp > CRF_MAX_CH) {
        delete crf;
        return ctx->fail(RSS_ERR_INVALID, "more than 32 label channels (every layer rounded up to 4) are not supported");
    }
    cudaError_t e = crf->unary.reserve((size_t)N * crf->Mp * 4);
    if (e == cudaSuccess) e = crf->Q.reserve((size_t)N * crf->Mp * 4);
    if (e == cudaSuccess) e = crf->labels.reserve((size_t)N * n_layers);
    if (e == cudaSuccess) e = crf->scratch.reserve((size_t)N * (crf->Mp + 4) * 4);
    if (e == cudaSuccess) {
        unary_clear(crf);
        e = cudaGetLastError();
    }
    for (int k = 0; k < 4 && e == cudaSuccess; k++) {
        // lattice construction runs beside the frame path and is on the keyframe's critical path: high priority
        int prio_lo = 0, prio_hi = 0;
        cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
        e = cudaStreamCreateWithPriority(&crf->side[k], cudaStreamNonBlocking, prio_hi);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&crf->ev_join[k], cudaEventDisableTiming);
    }
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&crf->ev_fork, cudaEventDisableTiming);
    if (e != cudaSuccess) {
        crf_free(crf);
        return ctx->fail(RSS_ERR_CUDA, std::string("CRF allocation: ") + cudaGetErrorString(e));
    }
    *out = crf;
    return RSS_OK;
}

}  // namespace rss

using namespace rss;

extern "C" rss_status rss_crf_create_layers(rss_ctx* ctx, int N, int n_layers, const int* M, rss_crf** out) {
    if (!ctx) return RSS_ERR_INVALID;
    RSS_CU(ctx, cudaSetDevice(ctx->device));
    return crf_new(ctx, N, n_layers, M, out);
}
extern "C" rss_status rss_crf_create(rss_ctx* ctx, int N, int M, rss_crf** out) { return rss_crf_create_layers(ctx, N, 1, &M, out); }

extern "C" rss_status rss_crf_destroy(rss_crf* crf) {
    if (!crf) return RSS_ERR_INVALID;
    cudaSetDevice(crf->ctx->device);
    cudaStreamSynchronize(crf->ctx->s0);
    crf_free(crf);
    return RSS_OK;
}

extern "C" rss_status rss_crf_set_unary(rss_crf* crf, int layer, const float* U) {
    if (!crf) return RSS_ERR_INVALID;
    rss_ctx* ctx = crf->ctx;
    if (layer < 0 || layer >= crf->n_layers || !U) return ctx->fail(RSS_ERR_INVALID, "bad layer or null unary");
    RSS_CU(ctx, cudaSetDevice(ctx->device));
    const int Ml = crf->M[layer];
    RSS_CU(ctx, cudaMemcpyAsync(crf->scratch.ptr, U, (size_t)crf->N * Ml * 4, cudaMemcpyHostToDevice, ctx->s0));
    RSS_LAUNCH(ctx, interleave_kernel, rss_div_up((long long)crf->N * Ml, 256), 256, 0, ctx->s0, crf->scratch.as<float>(),
               crf->N, Ml, Ml, 0, crf->Mp, crf->moff[layer], 1.0f, crf->unary.as<float>());
    RSS_CU(ctx, cudaStreamSynchronize(ctx->s0));
    crf->unary_set = true;
    return RSS_OK;
}

extern "C" rss_status rss_crf_add_pairwise(rss_crf* crf, const float* feats, int d, float potts_w, int norm_type) {
    if (!crf) return RSS_ERR_INVALID;
    rss_ctx* ctx = crf->ctx;
    if (!feats || d < 1 || d > LAT_MAX_D) return ctx->fail(RSS_ERR_INVALID, "null features or dimension outside [1, 7]");
    RSS_CU(ctx, cudaSetDevice(ctx->device));
    RSS_CU(ctx, crf->feat_stage.reserve((size_t)crf->N * d * 4));
    RSS_CU(ctx, cudaMemcpyAsync(crf->feat_stage.ptr, feats, (size_t)crf->N * d * 4, cudaMemcpyHostToDevice, ctx->s0));
    return crf_add_kernel_dev(crf, ctx->s0, crf->feat_stage.as<float>(), d, potts_w, norm_type, true);
}

extern "C" rss_status rss_crf_add_pairwise_gaussian(rss_crf* crf, int W, int H, float sx, float sy, float potts_w) {
    if (!crf) return RSS_ERR_INVALID;
    rss_ctx* ctx = crf->ctx;
    if ((long long)W * H != crf->N) return ctx->fail(RSS_ERR_INVALID, "W*H must equal the CRF's point count");
    if (crf->kernels.empty()) { crf->grid_w = W; crf->grid_h = H; }
    RSS_CU(ctx, cudaSetDevice(ctx->device));
    RSS_CU(ctx, crf->feat_stage.reserve((size_t)crf->N * 2 * 4));
    RSS_LAUNCH(ctx, feat_gaussian_kernel, rss_div_up(crf->N, 256), 256, 0, ctx->s0, W, H, sx, sy, crf->feat_stage.as<float>());
    return crf_add_kernel_dev(crf, ctx->s0, crf->feat_stage.as<float>(), 2, potts_w, RSS_NORMALIZE_SYMMETRIC, true);
}

extern "C" rss_status rss_crf_add_pairwise_bilateral(rss_crf* crf, int W, int H, float sx, float sy, float sr, float sg,
                                                     float sb, const uint8_t* im, float potts_w) {
    if (!crf) return RSS_ERR_INVALID;
    rss_ctx* ctx = crf->ctx;
    if ((long long)W * H != crf->N || !im) return ctx->fail(RSS_ERR_INVALID, "W*H must equal the CRF's point count; image required");
    if (crf->kernels.empty()) { crf->grid_w = W; crf->grid_h = H; }
    RSS_CU(ctx, cudaSetDevice(ctx->device));
    RSS_CU(ctx, crf->feat_stage.reserve((size_t)crf->N * 5 * 4));
    uint8_t* im_dev = reinterpret_cast<uint8_t*>(crf->scratch.ptr);  // N*Mp*4 >= N*3 bytes
    RSS_CU(ctx, cudaMemcpyAsync(im_dev, im, (size_t)crf->N * 3, cudaMemcpyHostToDevice, ctx->s0));
    RSS_LAUNCH(ctx, feat_bilateral_kernel, rss_div_up(crf->N, 256), 256, 0, ctx->s0, W, H, sx, sy, sr, sg, sb, im_dev,
               crf->feat_stage.as<float>());
    return crf_add_kernel_dev(crf, ctx->s0, crf->feat_stage.as<float>(), 5, potts_w, RSS_NORMALIZE_SYMMETRIC, true);
}

extern "C" rss_status rss_crf_add_pairwise_xyzrgb(rss_crf* crf, const float* xyz, const float* rgb, float wxyz, float wrgb,
                                                  float potts_w) {
    if (!crf) return RSS_ERR_INVALID;
    rss_ctx* ctx = crf->ctx;
    if (!xyz || !rgb) return ctx->fail(RSS_ERR_INVALID, "null point cloud");
    RSS_CU(ctx, cudaSetDevice(ctx->device));
    const size_t N = crf->N;
    RSS_CU(ctx, crf->feat_stage.reserve(N * 12 * 4));
    float* stage = crf->feat_stage.as<float>();  // [0,6N): features, [6N,9N): xyz, [9N,12N): rgb
    RSS_CU(ctx, cudaMemcpyAsync(stage + 6 * N, xyz, N * 12, cudaMemcpyHostToDevice, ctx->s0));
    RSS_CU(ctx, cudaMemcpyAsync(stage + 9 * N, rgb, N * 12, cudaMemcpyHostToDevice, ctx->s0));
    RSS_LAUNCH(ctx, feat_xyzrgb_kernel, rss_div_up((long long)N, 256), 256, 0, ctx->s0, (int)N, stage + 6 * N, stage + 9 * N,
               wxyz, wrgb, stage);
    return crf_add_kernel_dev(crf, ctx->s0, stage, 6, potts_w, RSS_NORMALIZE_SYMMETRIC, true);
}

// which mean-field path the next rss_crf_inference takes: *fused = 1 the fused point kernel (meanfield.cu), 0 the generic
// splat / blur / slice kernels; *sorted = 1 when the fused path runs over the sorted order of an incoherent point set
extern "C" rss_status rss_crf_path(rss_crf* crf, int* fused, int* sorted) {
    if (!crf) return RSS_ERR_INVALID;
    int a = 0, b = -1;
    const bool f = !crf->kernels.empty() && crf_fused_order(crf, &a, &b);
    if (fused) *fused = f ? 1 : 0;
    if (sorted) *sorted = (f && crf->sorted && !(crf->grid_w > 0 && (long long)crf->grid_w * crf->grid_h == crf->N)) ? 1 : 0;
    return RSS_OK;
}

extern "C" rss_status rss_crf_lattice_size(rss_crf* crf, int k, int* vertices) {
    if (!crf || !vertices) return RSS_ERR_INVALID;
    if (k < 0 || k >= (int)crf->kernels.size()) return crf->ctx->fail(RSS_ERR_INVALID, "no such pairwise term");
    *vertices = crf->kernels[k]->V_host;
    return RSS_OK;
}

extern "C" rss_status rss_crf_filter(rss_crf* crf, int k, const float* in, float* out) {
    if (!crf) return RSS_ERR_INVALID;
    rss_ctx* ctx = crf->ctx;
    if (k < 0 || k >= (int)crf->kernels.size() || !in || !out) return ctx->fail(RSS_ERR_INVALID, "bad filter arguments");
    RSS_CU(ctx, cudaSetDevice(ctx->device));
    Lattice& L = *crf->kernels[k];
    const int N = crf->N, M = crf->Mtot, Mp = crf->Mp;
    // scratch: [N][Mp] input in the device channel layout; feat_stage: the [N][M] host-layout matrix, then the [N][Mp]
    // filtered rows
    float* padded = crf->scratch.as<float>();
    RSS_CU(ctx, crf->feat_stage.reserve((size_t)N * (M + Mp) * 4));
    float* stage = crf->feat_stage.as<float>();
    float* sliced = stage + (size_t)N * M;
    RSS_CU(ctx, cudaMemsetAsync(padded, 0, (size_t)N * Mp * 4, ctx->s0));
    RSS_CU(ctx, cudaMemcpyAsync(stage, in, (size_t)N * M * 4, cudaMemcpyHostToDevice, ctx->s0));
    for (int l = 0; l < crf->n_layers; l++)
        RSS_LAUNCH(ctx, interleave_kernel, rss_div_up((long long)N * crf->M[l], 256), 256, 0, ctx->s0, stage, N, crf->M[l], M,
                   crf->hoff[l], Mp, crf->moff[l], 1.0f, padded);
    float* vals = lattice_splat_blur(ctx, ctx->s0, L, padded, Mp, nullptr, Mp);
    if (!vals) return ctx->fail(RSS_ERR_CUDA, std::string("cooperative blur launch: ") + cudaGetErrorString(cudaGetLastError()));
    lattice_slice(ctx, ctx->s0, L, vals, Mp, Mp, M <= 2 ? 1 : 0, sliced, Mp);
    for (int l = 0; l < crf->n_layers; l++)
        RSS_LAUNCH(ctx, deinterleave_kernel, rss_div_up((long long)N * crf->M[l], 256), 256, 0, ctx->s0, sliced, N, crf->M[l], Mp,
                   crf->moff[l], M, crf->hoff[l], stage);
    RSS_CU(ctx, cudaMemcpyAsync(out, stage, (size_t)N * M * 4, cudaMemcpyDeviceToHost, ctx->s0));
    RSS_CU(ctx, cudaStreamSynchronize(ctx->s0));
    return RSS_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// DenseCRF::gradient (densecrf.cpp:238-297) for the label-compatibility parameters of Potts terms, with the
// log-likelihood objective (objective.cpp:36-52).  All matrices are [N][Mp] device rows; only the channels of the
// requested layer carry non-zero "b", so the other layers drop out of every sum.
// ---------------------------------------------------------------------------------------------------------------
// LogLikelihood::evaluate: r += log(QQ) / N, d_mul_Q(gt, i) = Q / QQ / N with QQ = max(Q(gt, i) + robust, 1e-20)
__global__ void __launch_bounds__(256) grad_loglik_kernel(const float* __restrict__ Q, const int* __restrict__ gt, int N, int Mp,
                                                          int off, int Ml, float robust, float* __restrict__ b,
                                                          double* __restrict__ objective) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double r = 0.0;
    if (i < N) {
        const int g = gt[i];
        if (g >= 0 && g < Ml) {
            const float q = Q[(size_t)i * Mp + off + g];
            const float QQ = fmaxf(__fadd_rn(q, robust), 1e-20f);
            r = (double)__fdiv_rn(logf(QQ), (float)N);
            b[(size_t)i * Mp + off + g] = __fdiv_rn(__fdiv_rn(q, QQ), (float)N);
        }
    }
    for (int o = 16; o > 0; o >>= 1) r += __shfl_down_sync(0xffffffffu, r, o);
    if ((threadIdx.x & 31) == 0 && r != 0.0) atomicAdd(objective, r);
}
// sumAndNormalize (densecrf.cpp:107-114) of in (.* mul when given): out = sum(b) * q - b per point, over one layer
__global__ void __launch_bounds__(256) grad_sum_normalize_kernel(const float* __restrict__ in, const float* __restrict__ mul,
                                                                 const float* __restrict__ Q, int N, int Mp, int off, int Ml,
                                                                 float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const size_t o = (size_t)i * Mp + off;
    float v[32];
    float s = 0.f;
    for (int c = 0; c < Ml; c++) {
        v[c] = mul ? __fmul_rn(in[o + c], mul[o + c]) : in[o + c];
        s = __fadd_rn(s, v[c]);
    }
    for (int c = 0; c < Ml; c++) out[o + c] = __fsub_rn(__fmul_rn(s, Q[o + c]), v[c]);
}
// out[i][c] (op)= scale * in[i][c] * (norm ? norm[i] : 1) on one layer's channels
__global__ void __launch_bounds__(256) grad_axpy_kernel(const float* __restrict__ in, const float* __restrict__ norm, float scale,
                                                        int N, int Mp, int off, int Ml, int accumulate, float* __restrict__ out) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)N * Ml) return;
    const int i = (int)(gid / Ml), c = (int)(gid - (long long)i * Ml);
    const size_t o = (size_t)i * Mp + off + c;
    float v = in[o];
    if (norm) v = __fmul_rn(v, norm[i]);
    v = __fmul_rn(scale, v);
    out[o] = accumulate ? __fadd_rn(out[o], v) : v;
}
// sum over the layer's channels of a .* b (PottsCompatibility::gradient = -(b .* filtered Q).sum(), labelcompatibility.cpp:57-61)
__global__ void __launch_bounds__(256) grad_dot_kernel(const float* __restrict__ a, const float* __restrict__ bb,
                                                       const float* __restrict__ norm, int N, int Mp, int off, int Ml,
                                                       double* __restrict__ acc) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double r = 0.0;
    if (i < N) {
        const size_t o = (size_t)i * Mp + off;
        const float nv = norm ? norm[i] : 1.f;
        for (int c = 0; c < Ml; c++) r += (double)__fmul_rn(__fmul_rn(a[o + c], nv), bb[o + c]);
    }
    for (int o = 16; o > 0; o >>= 1) r += __shfl_down_sync(0xffffffffu, r, o);
    if ((threadIdx.x & 31) == 0 && r != 0.0) atomicAdd(acc, r);
}
// DenseKernel::filter (pairwise.cpp:40-62) without the trailing normalisation: pre-normalisation + lattice filter
// (reversed axis order when transposed) + slice.  Returns whether the caller still has to multiply by norm (post).
static rss_status grad_filter(rss_crf* crf, Lattice& L, const float* in, bool transpose, float* out, bool* post) {
    rss_ctx* ctx = crf->ctx;
    const int nt = L.norm_type;
    const bool pre = nt == RSS_NORMALIZE_SYMMETRIC || (nt == RSS_NORMALIZE_BEFORE && !transpose) || (nt == RSS_NORMALIZE_AFTER && transpose);
    *post = nt == RSS_NORMALIZE_SYMMETRIC || (nt == RSS_NORMALIZE_BEFORE && transpose) || (nt == RSS_NORMALIZE_AFTER && !transpose);
    float* vals = lattice_splat_blur(ctx, ctx->s0, L, in, crf->Mp, pre ? L.norm.as<float>() : nullptr, crf->Mp, transpose);
    if (!vals) return ctx->fail(RSS_ERR_CUDA, std::string("cooperative blur launch: ") + cudaGetErrorString(cudaGetLastError()));
    lattice_slice(ctx, ctx->s0, L, vals, crf->Mp, crf->Mp, crf->Mtot <= 2 ? 1 : 0, out, crf->Mp);
    return RSS_OK;
}
extern "C" rss_status rss_crf_gradient(rss_crf* crf, int layer, int iters, const int32_t* gt, float robust, double* objective,
                                       float* potts_grad) {
    if (!crf) return RSS_ERR_INVALID;
    rss_ctx* ctx = crf->ctx;
    if (layer < 0 || layer >= crf->n_layers || iters < 0 || !gt || !objective)
        return ctx->fail(RSS_ERR_INVALID, "gradient: bad layer, iteration count or null argument");
    for (Lattice* L : crf->kernels)
        if (!L->have_csr) return ctx->fail(RSS_ERR_STATE, "gradient: needs pairwise terms added through rss_crf_add_pairwise*");
    RSS_CU(ctx, cudaSetDevice(ctx->device));
    cudaStream_t s0 = ctx->s0;
    const int N = crf->N, Mp = crf->Mp, K = (int)crf->kernels.size();
    const int off = crf->moff[layer], Ml = crf->M[layer];
    const size_t rows = (size_t)N * Mp;
    DevBuf hist, bbuf, t1, t2, gt_dev, acc;
    struct Free { DevBuf* b[6]; ~Free() { for (DevBuf* p : b) p->release(); } } guard{{&hist, &bbuf, &t1, &t2, &gt_dev, &acc}};
    RSS_CU(ctx, hist.reserve(rows * 4 * (size_t)(iters + 1)));
    RSS_CU(ctx, bbuf.reserve(rows * 4));
    RSS_CU(ctx, t1.reserve(rows * 4));
    RSS_CU(ctx, t2.reserve(rows * 4));
    RSS_CU(ctx, gt_dev.reserve((size_t)N * 4));
    RSS_CU(ctx, acc.reserve(8 * (size_t)(K + 1)));
    RSS_CU(ctx, cudaMemcpyAsync(gt_dev.ptr, gt, (size_t)N * 4, cudaMemcpyHostToDevice, s0));
    RSS_CU(ctx, cudaMemsetAsync(acc.ptr, 0, 8 * (size_t)(K + 1), s0));
    // forward: Q[0] = expAndNormalize(-unary), Q[it + 1] = one mean-field step (generic kernels), every Q kept
    rss_status st = crf_run(crf, 0, nullptr, nullptr, true);
    if (st != RSS_OK) return st;
    RSS_CU(ctx, cudaMemcpyAsync(hist.ptr, crf->Q.ptr, rows * 4, cudaMemcpyDeviceToDevice, s0));
    for (int it = 0; it < iters; it++) {
        st = crf_run(crf, 1, nullptr, nullptr, false);
        if (st != RSS_OK) return st;
        RSS_CU(ctx, cudaMemcpyAsync(hist.as<float>() + rows * (size_t)(it + 1), crf->Q.ptr, rows * 4, cudaMemcpyDeviceToDevice, s0));
    }
    float* b = bbuf.as<float>();
    float* tmp1 = t1.as<float>();
    float* tmp2 = t2.as<float>();
    double* accd = acc.as<double>();
    const float* Qn = hist.as<float>() + rows * (size_t)iters;
    const int gN = rss_div_up(N, 256), gNM = rss_div_up((long long)N * Ml, 256);
    // objective and b = sumAndNormalize(d_mul_Q, Q[n])
    RSS_CU(ctx, cudaMemsetAsync(tmp1, 0, rows * 4, s0));
    RSS_LAUNCH(ctx, grad_loglik_kernel, gN, 256, 0, s0, Qn, gt_dev.as<int>(), N, Mp, off, Ml, robust, tmp1, accd);
    RSS_CU(ctx, cudaMemsetAsync(b, 0, rows * 4, s0));
    RSS_LAUNCH(ctx, grad_sum_normalize_kernel, gN, 256, 0, s0, (const float*)tmp1, (const float*)nullptr, Qn, N, Mp, off, Ml, b);
    for (int it = iters - 1; it >= 0; it--) {
        const float* Qi = hist.as<float>() + rows * (size_t)it;
        RSS_CU(ctx, cudaMemsetAsync(tmp1, 0, rows * 4, s0));
        for (int k = 0; k < K; k++) {
            Lattice& L = *crf->kernels[k];
            bool post = false;
            if (potts_grad) {  // PairwisePotential::gradient (pairwise.cpp:190-195): -(b .* kernel(Q[it])).sum()
                st = grad_filter(crf, L, Qi, false, tmp2, &post);
                if (st != RSS_OK) return st;
                RSS_LAUNCH(ctx, grad_dot_kernel, gN, 256, 0, s0, (const float*)tmp2, (const float*)b,
                           post ? L.norm.as<float>() : (const float*)nullptr, N, Mp, off, Ml, accd + 1 + k);
            }
            // PairwisePotential::applyTranspose (pairwise.cpp:179-183): kernel^T(b), then Potts: out = -w * out
            st = grad_filter(crf, L, b, true, tmp2, &post);
            if (st != RSS_OK) return st;
            RSS_LAUNCH(ctx, grad_axpy_kernel, gNM, 256, 0, s0, (const float*)tmp2, post ? L.norm.as<float>() : (const float*)nullptr,
                       -L.potts_w, N, Mp, off, Ml, 1, tmp1);
        }
        // b = sumAndNormalize(tmp1 .* Q[it], Q[it])
        RSS_LAUNCH(ctx, grad_sum_normalize_kernel, gN, 256, 0, s0, (const float*)tmp1, Qi, Qi, N, Mp, off, Ml, b);
    }
    std::vector<double> h(K + 1);
    RSS_CU(ctx, cudaMemcpyAsync(h.data(), accd, 8 * (size_t)(K + 1), cudaMemcpyDeviceToHost, s0));
    RSS_CU(ctx, cudaStreamSynchronize(s0));
    RSS_CU(ctx, cudaGetLastError());
    *objective = h[0];
    if (potts_grad)
        for (int k = 0; k < K; k++) potts_grad[k] = (float)(-h[1 + k]);
    for (Lattice* L : crf->kernels) {
        uint32_t hh[2];
        RSS_CU(ctx, cudaMemcpy(hh, L->counts.ptr, sizeof(hh), cudaMemcpyDeviceToHost));
        if (hh[1]) return ctx->fail(RSS_ERR_CAPACITY, "lattice hash table overflow");
    }
    return RSS_OK;
}

// device -> host copies of the current Q (per-layer matrices, layers concatenated for layer = -1) and label maps
static rss_status crf_fetch(rss_crf* crf, int layer, float* Q, uint8_t* labels) {
    rss_ctx* ctx = crf->ctx;
    const int N = crf->N;
    if (Q) {
        if (layer < 0 && crf->n_layers == 1) layer = 0;
        if (layer < 0) {
            float* dstQ = Q;
            for (int l = 0; l < crf->n_layers; l++) {
                RSS_LAUNCH(ctx, deinterleave_kernel, rss_div_up((long long)N * crf->M[l], 256), 256, 0, ctx->s0,
                           crf->Q.as<float>(), N, crf->M[l], crf->Mp, crf->moff[l], crf->M[l], 0, crf->scratch.as<float>());
                RSS_CU(ctx, cudaMemcpyAsync(dstQ, crf->scratch.ptr, (size_t)N * crf->M[l] * 4, cudaMemcpyDeviceToHost, ctx->s0));
                RSS_CU(ctx, cudaStreamSynchronize(ctx->s0));
                dstQ += (size_t)N * crf->M[l];
            }
        } else {
            RSS_LAUNCH(ctx, deinterleave_kernel, rss_div_up((long long)N * crf->M[layer], 256), 256, 0, ctx->s0,
                       crf->Q.as<float>(), N, crf->M[layer], crf->Mp, crf->moff[layer], crf->M[layer], 0, crf->scratch.as<float>());
            RSS_CU(ctx, cudaMemcpyAsync(Q, crf->scratch.ptr, (size_t)N * crf->M[layer] * 4, cudaMemcpyDeviceToHost, ctx->s0));
        }
    }
    if (labels) {
        if (layer < 0) RSS_CU(ctx, cudaMemcpyAsync(labels, crf->labels.ptr, (size_t)N * crf->n_layers, cudaMemcpyDeviceToHost, ctx->s0));
        else RSS_CU(ctx, cudaMemcpyAsync(labels, crf->labels.as<uint8_t>() + (size_t)layer * N, (size_t)N, cudaMemcpyDeviceToHost, ctx->s0));
    }
    RSS_CU(ctx, cudaStreamSynchronize(ctx->s0));
    for (Lattice* L : crf->kernels) {
        uint32_t h[2];
        RSS_CU(ctx, cudaMemcpy(h, L->counts.ptr, sizeof(h), cudaMemcpyDeviceToHost));
        if (h[1]) return ctx->fail(RSS_ERR_CAPACITY, "lattice hash table overflow");
    }
    return RSS_OK;
}
static void fill_unknown(const rss_crf* crf, int layer, const int* unknown_label, int* unk) {
    for (int l = 0; l < RSS_MAX_LAYERS; l++) unk[l] = -1;
    if (unknown_label) {
        if (layer < 0) for (int l = 0; l < crf->n_layers; l++) unk[l] = unknown_label[l];
        else unk[layer] = unknown_label[0];
    }
}

extern "C" rss_status rss_crf_inference(rss_crf* crf, int layer, int iters, float* Q, uint8_t* labels,
                                        const int* unknown_label) {
    if (!crf) return RSS_ERR_INVALID;
    rss_ctx* ctx = crf->ctx;
    if (layer < -1 || layer >= crf->n_layers || iters < 0) return ctx->fail(RSS_ERR_INVALID, "bad layer or iteration count");
    RSS_CU(ctx, cudaSetDevice(ctx->device));
    int unk[RSS_MAX_LAYERS];
    fill_unknown(crf, layer, unknown_label, unk);
    cudaEventRecord(ctx->ev[6], ctx->s0);
    rss_status st = crf_run(crf, iters, unk, labels ? crf->labels.as<uint8_t>() : nullptr);
    if (st != RSS_OK) return st;
    cudaEventRecord(ctx->ev[7], ctx->s0);
    st = crf_fetch(crf, layer, Q, labels);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ctx->ev[6], ctx->ev[7]);
    ctx->tim.meanfield_ms = ms;
    return st;
}

// DenseCRF::startInference / stepInference / currentMap (densecrf.cpp:178-211): Q stays on the device between steps
extern "C" rss_status rss_crf_start_inference(rss_crf* crf) {
    if (!crf) return RSS_ERR_INVALID;
    rss_ctx* ctx = crf->ctx;
    RSS_CU(ctx, cudaSetDevice(ctx->device));
    return crf_run(crf, 0, nullptr, nullptr, true);
}
extern "C" rss_status rss_crf_step_inference(rss_crf* crf, int steps) {
    if (!crf) return RSS_ERR_INVALID;
    rss_ctx* ctx = crf->ctx;
    if (steps < 1) return ctx->fail(RSS_ERR_INVALID, "steps must be >= 1");
    RSS_CU(ctx, cudaSetDevice(ctx->device));
    return crf_run(crf, steps, nullptr, nullptr, false);
}
extern "C" rss_status rss_crf_current(rss_crf* crf, int layer, float* Q, uint8_t* labels, const int* unknown_label) {
    if (!crf) return RSS_ERR_INVALID;
    rss_ctx* ctx = crf->ctx;
    if (layer < -1 || layer >= crf->n_layers) return ctx->fail(RSS_ERR_INVALID, "bad layer");
    RSS_CU(ctx, cudaSetDevice(ctx->device));
    if (labels) {
        int unk[RSS_MAX_LAYERS];
        fill_unknown(crf, layer, unknown_label, unk);
        const LayerSpec ls = make_layers(crf, unk);
        RSS_LAUNCH(ctx, argmax_kernel, rss_div_up((long long)crf->N * 8, 256), 256, 0, ctx->s0, (const float*)crf->Q.as<float>(),
                   crf->N, crf->Mp / 4, crf->Mp, ls, crf->labels.as<uint8_t>());
    }
    return crf_fetch(crf, layer, Q, labels);
}

extern "C" rss_status rss_crf_unary_reset(rss_crf* crf) {
    if (!crf) return RSS_ERR_INVALID;
    rss_ctx* ctx = crf->ctx;
    RSS_CU(ctx, cudaSetDevice(ctx->device));
    unary_clear(crf);
    RSS_CU(ctx, cudaStreamSynchronize(ctx->s0));
    return RSS_OK;
}

extern "C" rss_status rss_crf_clear_pairwise(rss_crf* crf) {
    if (!crf) return RSS_ERR_INVALID;
    rss_ctx* ctx = crf->ctx;
    RSS_CU(ctx, cudaSetDevice(ctx->device));
    RSS_CU(ctx, cudaStreamSynchronize(ctx->s0));
    while (!crf->kernels.empty()) { crf->pool.push_back(crf->kernels.back()); crf->kernels.pop_back(); }
    crf->sorted = false;
    return RSS_OK;
}

extern "C" rss_status rss_crf_unary_accumulate(rss_crf* crf, rss_ctx* frame_ctx, const int32_t* index_image, int npix,
                                               const float* posteriors) {
    if (!crf) return RSS_ERR_INVALID;
    rss_ctx* ctx = crf->ctx;
    if (!index_image || npix < 1) return ctx->fail(RSS_ERR_INVALID, "null index image");
    RSS_CU(ctx, cudaSetDevice(ctx->device));
    const float* post_dev = nullptr;
    RSS_CU(ctx, crf->feat_stage.reserve((size_t)npix * 4 + (posteriors ? (size_t)npix * crf->Mtot * 4 : 0)));
    int* idx_dev = crf->feat_stage.as<int>();
    RSS_CU(ctx, cudaMemcpyAsync(idx_dev, index_image, (size_t)npix * 4, cudaMemcpyHostToDevice, ctx->s0));
    if (posteriors) {
        float* p = crf->feat_stage.as<float>() + npix;
        RSS_CU(ctx, cudaMemcpyAsync(p, posteriors, (size_t)npix * crf->Mtot * 4, cudaMemcpyHostToDevice, ctx->s0));
        post_dev = p;
    } else {
        if (!frame_ctx || frame_ctx != ctx) return ctx->fail(RSS_ERR_INVALID, "frame context must be the CRF's context");
        if (!ctx->fr.have_post || ctx->fr.W * ctx->fr.H != npix || ctx->forest.sumC != crf->Mtot)
            return ctx->fail(RSS_ERR_STATE, "no resident posteriors matching this index image");
        post_dev = ctx->fr.posteriors.as<float>();
    }
    for (int l = 0; l < crf->n_layers; l++) {
        RSS_LAUNCH(ctx, unary_accumulate_kernel, rss_div_up((long long)npix * crf->M[l], 256), 256, 0, ctx->s0, idx_dev, npix,
                   post_dev + (size_t)npix * crf->hoff[l], crf->M[l], crf->Mp, crf->moff[l], crf->N, crf->unary.as<float>());
    }
    RSS_CU(ctx, cudaStreamSynchronize(ctx->s0));
    crf->unary_set = true;
    return RSS_OK;
}

// --------------------------------------------------------------------------------------------------------------------
// map worker with the cloud, the posteriors and the index images on the device (src/segmenter.cpp:559-637)
// --------------------------------------------------------------------------------------------------------------------
extern "C" rss_status rss_crf_set_cloud(rss_crf* crf, const float* xyz, const float* rgb) {
    if (!crf) return RSS_ERR_INVALID;
    rss_ctx* ctx = crf->ctx;
    if (!xyz || !rgb) return ctx->fail(RSS_ERR_INVALID, "null point cloud");
    RSS_CU(ctx, cudaSetDevice(ctx->device));
    const size_t bytes = (size_t)crf->N * 12;
    RSS_CU(ctx, crf->cloud_xyz.reserve(bytes));
    RSS_CU(ctx, crf->cloud_rgb.reserve(bytes));
    RSS_CU(ctx, cudaMemcpyAsync(crf->cloud_xyz.ptr, xyz, bytes, cudaMemcpyHostToDevice, ctx->s0));
    RSS_CU(ctx, cudaMemcpyAsync(crf->cloud_rgb.ptr, rgb, bytes, cudaMemcpyHostToDevice, ctx->s0));
    RSS_CU(ctx, cudaStreamSynchronize(ctx->s0));  // the host buffers may be reused by the caller
    crf->have_cloud = true;
    return RSS_OK;
}

extern "C" rss_status rss_crf_project_accumulate(rss_crf* crf, rss_ctx* frame_ctx, int slot, int W, int H, const float K[9],
                                                 const float R[9], const float t[3], float zmin, float zmax,
                                                 int32_t* index_image_out) {
    if (!crf) return RSS_ERR_INVALID;
    rss_ctx* ctx = crf->ctx;
    if (!K || !R || !t || W < 1 || H < 1) return ctx->fail(RSS_ERR_INVALID, "null calibration or bad image size");
    if (!crf->have_cloud) return ctx->fail(RSS_ERR_STATE, "no resident cloud (call rss_crf_set_cloud first)");
    if (!frame_ctx || frame_ctx != ctx) return ctx->fail(RSS_ERR_INVALID, "frame context must be the CRF's context");
    RSS_CU(ctx, cudaSetDevice(ctx->device));
    const int npix = W * H;
    const FrameState& f = ctx->fr;
    const float* post_dev = nullptr;
    if (slot < 0) {
        if (!f.have_post || f.W * f.H != npix) return ctx->fail(RSS_ERR_STATE, "no resident posteriors matching this image size");
        post_dev = f.posteriors.as<float>();
    } else {
        if (slot >= (int)f.kept.size() || f.kept_npix[slot] != npix)
            return ctx->fail(RSS_ERR_STATE, "no kept posteriors of this image size in this slot");
        post_dev = f.kept[slot].as<float>();
    }
    if (ctx->forest.sumC != crf->Mtot) return ctx->fail(RSS_ERR_STATE, "the CRF's layers do not match the forest's");
    RSS_CU(ctx, crf->zbuf.reserve((size_t)npix * 8));
    RSS_CU(ctx, crf->index_dev.reserve((size_t)npix * 4));
    ProjParams P;
    for (int i = 0; i < 9; i++) P.R[i] = R[i];
    for (int i = 0; i < 3; i++) P.t[i] = t[i];
    P.fx = K[0]; P.cx = K[2]; P.fy = K[4]; P.cy = K[5]; P.zmin = zmin; P.zmax = zmax;
    RSS_CU(ctx, cudaMemsetAsync(crf->zbuf.ptr, 0xFF, (size_t)npix * 8, ctx->s0));
    RSS_LAUNCH(ctx, project_zbuffer_kernel, rss_div_up(crf->N, 256), 256, 0, ctx->s0, crf->cloud_xyz.as<float>(), crf->N, P, W, H,
               crf->zbuf.as<unsigned long long>());
    RSS_LAUNCH(ctx, zbuffer_to_index_kernel, rss_div_up(npix, 256), 256, 0, ctx->s0, crf->zbuf.as<unsigned long long>(), npix,
               crf->index_dev.as<int>());
    // a point projects to one pixel, so every index occurs at most once per key frame: the adds below never collide
    for (int l = 0; l < crf->n_layers; l++) {
        RSS_LAUNCH(ctx, unary_accumulate_kernel, rss_div_up((long long)npix * crf->M[l], 256), 256, 0, ctx->s0,
                   crf->index_dev.as<int>(), npix, post_dev + (size_t)npix * crf->hoff[l], crf->M[l], crf->Mp, crf->moff[l],
                   crf->N, crf->unary.as<float>());
    }
    if (index_image_out)
        RSS_CU(ctx, cudaMemcpyAsync(index_image_out, crf->index_dev.ptr, (size_t)npix * 4, cudaMemcpyDeviceToHost, ctx->s0));
    RSS_CU(ctx, cudaStreamSynchronize(ctx->s0));
    crf->unary_set = true;
    return RSS_OK;
}

extern "C" rss_status rss_crf_add_pairwise_cloud(rss_crf* crf, float wxyz, float wrgb, float potts_w) {
    if (!crf) return RSS_ERR_INVALID;
    rss_ctx* ctx = crf->ctx;
    if (!crf->have_cloud) return ctx->fail(RSS_ERR_STATE, "no resident cloud (call rss_crf_set_cloud first)");
    RSS_CU(ctx, cudaSetDevice(ctx->device));
    const size_t N = crf->N;
    RSS_CU(ctx, crf->feat_stage.reserve(N * 6 * 4));
    RSS_LAUNCH(ctx, feat_xyzrgb_kernel, rss_div_up((long long)N, 256), 256, 0, ctx->s0, (int)N, crf->cloud_xyz.as<float>(),
               crf->cloud_rgb.as<float>(), wxyz, wrgb, crf->feat_stage.as<float>());
    return crf_add_kernel_dev(crf, ctx->s0, crf->feat_stage.as<float>(), 6, potts_w, RSS_NORMALIZE_SYMMETRIC, true);
}

// --------------------------------------------------------------------------------------------------------------------
// fused keyframe (configs[1]/[2] of BASELINE.json)
// --------------------------------------------------------------------------------------------------------------------
namespace rss {
rss_status frame_upload(rss_ctx* ctx, const uint8_t* rgb, const uint16_t* depth, int W, int H);
rss_status frame_segment_resident(rss_ctx* ctx, const float* Kinv, const float* R, const float* t, float fill);
rss_status frame_segment_begin(rss_ctx* ctx, const float* Kinv, const float* R, const float* t);
rss_status frame_segment_finish(rss_ctx* ctx, float fill, float* unary_out = nullptr, int unary_stride = 0);
rss_status frame_set_pose(rss_ctx* ctx, const float* Kinv, const float* R, const float* t);
void frame_collect_timings(rss_ctx* ctx, bool with_d2h);
}

// The device part of a keyframe, captured once into a CUDA graph and replayed: ~60 launches on four streams, no host
// synchronisation inside (sample selection, vertex counts and overflow flags all stay on the device).  The graph holds
// raw pointers and launch parameters, so it is only replayed while (W, H, parameters) are those of the capture and no
// device buffer has been reallocated since; the pose travels through rss_ctx::pose_dev, so it may change every frame.
struct KeyframeGraph {
    cudaGraphExec_t exec = nullptr;
    uint64_t alloc_gen = 0;      // device_alloc_events() at capture time
    int W = 0, H = 0;
    rss_keyframe_params prm{};
    uint64_t launches = 0;       // kernel launches inside the graph
    bool shared_gpu = false;     // captured while other contexts were alive on the device (smaller blur CTAs)
    bool want_q = false;         // captured with the marginals stored by the last point kernel
    // stability detection: the previous eager call's signature and the allocation counter after it
    int prev_W = 0, prev_H = 0;
    rss_keyframe_params prev_prm{};
    uint64_t prev_alloc_gen = ~0ull;
    bool broken = false;         // a capture failed on this context: stay eager
};
namespace rss {
void keyframe_graph_release(rss_ctx* ctx) {
    if (!ctx->kf_graph) return;
    if (ctx->kf_graph->exec) cudaGraphExecDestroy(ctx->kf_graph->exec);
    delete ctx->kf_graph;
    ctx->kf_graph = nullptr;
}
}  // namespace rss

static bool same_params(const rss_keyframe_params& a, const rss_keyframe_params& b) { return memcmp(&a, &b, sizeof(a)) == 0; }

// everything between the frame upload and the label download; the frame is resident, the pose is in pin_pose
static rss_status keyframe_enqueue(rss_ctx* ctx, rss_crf* crf, int W, int H, const rss_keyframe_params* prm) {
    const int N = W * H;
    float* f3 = crf->feat_stage.as<float>();
    float* f5 = f3 + (size_t)N * 3;
    cudaStream_t sA = crf->side[0], sB = crf->side[1];
    rss_status st;
    // lattices of the previous keyframe are rebuilt in place (pooled buffers keep their capacity)
    while (!crf->kernels.empty()) { crf->pool.push_back(crf->kernels.back()); crf->kernels.pop_back(); }
    RSS_CU(ctx, cudaEventRecord(crf->ev_fork, ctx->s0));
    RSS_CU(ctx, cudaStreamWaitEvent(sB, crf->ev_fork, 0));
    RSS_LAUNCH(ctx, feat_bilateral_kernel, rss_div_up(N, 256), 256, 0, sB, W, H, prm->sigma_px, prm->sigma_px,
               prm->sigma_rgb, prm->sigma_rgb, prm->sigma_rgb, ctx->fr.rgb.as<uint8_t>(), f5);
    // Enqueue order = start order: the bilateral lattice needs only the colour image, so its whole construction is
    // queued on sB first; then the first half of the frame path (Lab on s0, cloud + normals on s1); then the Gaussian
    // lattice on sA behind the cloud; then the forest and the up-sample.  kernels[0] = Gaussian, kernels[1] = bilateral.
    Lattice* pooled[2] = {nullptr, nullptr};  // [0] 3-D, [1] 5-D buffers from the previous keyframe
    for (Lattice* L : crf->pool) pooled[L->d == 3 ? 0 : 1] = L;
    crf->pool.clear();
    if (pooled[1]) crf->pool.push_back(pooled[1]);
    st = crf_add_kernel_dev(crf, sB, f5, 5, prm->w_bilateral, RSS_NORMALIZE_SYMMETRIC, false, true);
    if (st != RSS_OK) return st;
    st = frame_segment_begin(ctx, nullptr, nullptr, nullptr);
    if (st != RSS_OK) return st;
    RSS_CU(ctx, cudaStreamWaitEvent(sA, ctx->ev_cloud, 0));
    RSS_LAUNCH(ctx, feat_frame_xyz_kernel, rss_div_up(N, 256), 256, 0, sA, N, ctx->fr.xyz.as<float4>(), 1.0f / prm->sigma_xyz,
               ctx->pose_dev.as<PoseParams>(), f3);
    if (pooled[0]) crf->pool.push_back(pooled[0]);
    st = crf_add_kernel_dev(crf, sA, f3, 3, prm->w_gauss, RSS_NORMALIZE_SYMMETRIC, false, true);
    if (st != RSS_OK) return st;
    std::swap(crf->kernels[0], crf->kernels[1]);
    // the up-sample writes the energies (-log-posteriors, src/segmenter.cpp:642) straight into the CRF's unary matrix
    st = frame_segment_finish(ctx, prm->fill, crf->unary.as<float>(), crf->Mp);
    if (st != RSS_OK) return st;
    RSS_CU(ctx, cudaEventRecord(crf->ev_join[0], sA));
    RSS_CU(ctx, cudaEventRecord(crf->ev_join[1], sB));
    // ---- the mean-field loop once both lattices are ready
    ctx->mark(8);
    RSS_CU(ctx, cudaStreamWaitEvent(ctx->s0, crf->ev_join[0], 0));
    RSS_CU(ctx, cudaStreamWaitEvent(ctx->s0, crf->ev_join[1], 0));
    ctx->mark(9);
    int unk[RSS_MAX_LAYERS];
    for (int l = 0; l < RSS_MAX_LAYERS; l++) unk[l] = l < ctx->cfg.layer_count ? ctx->cfg.unknown_label[l] : 0;
    st = crf_run(crf, prm->iters, unk, crf->labels.as<uint8_t>());
    if (st != RSS_OK) return st;
    ctx->mark(10);
    // vertex counts and overflow flags of both lattices -> pinned memory (read after the final synchronisation)
    for (size_t k = 0; k < crf->kernels.size(); k++)
        RSS_CU(ctx, cudaMemcpyAsync(ctx->pin_small.as<uint32_t>() + 2 * k, crf->kernels[k]->counts.ptr, 8, cudaMemcpyDeviceToHost,
                                    ctx->s0));
    return RSS_OK;
}

extern "C" rss_status rss_keyframe_graph(rss_ctx* ctx, int enable) {
    if (!ctx) return RSS_ERR_INVALID;
    ctx->graph_enabled = enable != 0;
    return RSS_OK;
}

extern "C" rss_status rss_segment_keyframe(rss_ctx* ctx, const uint8_t* rgb, const uint16_t* depth_mm, int W, int H,
                                           const float Kinv[9], const float R[9], const float t[3],
                                           const rss_keyframe_params* prm, uint8_t* labels, float* Qout) {
    if (!ctx) return RSS_ERR_INVALID;
    if (!Kinv || !R || !t || !prm) return ctx->fail(RSS_ERR_INVALID, "null calibration or parameters");
    if (!(ctx->cfg.use_height || ctx->cfg.use_normal))
        return ctx->fail(RSS_ERR_STATE, "the keyframe CRF needs the point cloud (feature_height or feature_normal)");
    RSS_CU(ctx, cudaSetDevice(ctx->device));
    const ForestDev& F = ctx->forest;
    if (!F.loaded) return ctx->fail(RSS_ERR_STATE, "no forest loaded");
    // the gated argmax takes its "Unknown" labels from the config's layers, the CRF its label counts from the forest:
    // the two must describe the same layers (segmenter.cpp:81-87 builds both from one config)
    if (ctx->cfg.layer_count > 0) {
        if (ctx->cfg.layer_count != F.L) return ctx->fail(RSS_ERR_STATE, "config and forest disagree on the number of label layers");
        for (int l = 0; l < F.L; l++)
            if (ctx->cfg.class_counts[l] != F.C[l])
                return ctx->fail(RSS_ERR_STATE, "config and forest disagree on the class count of a label layer");
    }
    if (W % ctx->cfg.rf_stride || H % ctx->cfg.rf_stride)
        return ctx->fail(RSS_ERR_INVALID, "image size must be a multiple of rf_prediction_stride");
    const int N = W * H;
    rss_crf* crf = ctx->keyframe_crf;
    rss_status st;
    if (!crf || crf->N != N || crf->n_layers != F.L || crf->Mtot != F.sumC) {
        crf_release_cached(ctx);
        st = crf_new(ctx, N, F.L, F.C, &crf);
        if (st != RSS_OK) return st;
        ctx->keyframe_crf = crf;
    }
    crf->grid_w = W; crf->grid_h = H;
    RSS_CU(ctx, crf->feat_stage.reserve((size_t)N * 8 * 4));
    RSS_CU(ctx, ctx->pin_small.reserve(64));
    if (!ctx->kf_graph) ctx->kf_graph = new KeyframeGraph();
    KeyframeGraph& G = *ctx->kf_graph;
    // ---- upload (outside the graph: the host pointers are the caller's) and the pose staging block
    cudaEventRecord(ctx->ev[0], ctx->s0);
    st = frame_upload(ctx, rgb, depth_mm, W, H);
    if (st != RSS_OK) return st;
    cudaEventRecord(ctx->ev[1], ctx->s0);
    st = frame_set_pose(ctx, Kinv, R, t);
    if (st != RSS_OK) return st;
    const bool want_graph = ctx->graph_enabled && !ctx->profile && !G.broken && !getenv("RSS_NO_GRAPH");
    bool staged = false;  // per-stage events recorded (eager runs only)
    crf->skip_q_store = Qout == nullptr;
    for (int attempt = 0;; attempt++) {
        const uint64_t gen = device_alloc_events().load();
        bool ran = false;
        const bool shared_gpu = live_contexts(ctx->device).load() > 1;
        if (want_graph && attempt == 0 && G.exec && G.W == W && G.H == H && same_params(G.prm, *prm) && G.alloc_gen == gen &&
            G.shared_gpu == shared_gpu && G.want_q == (Qout != nullptr)) {
            // steady state: replay.  The frame-state flags the eager path maintains:
            ctx->fr.have_cloud = ctx->cfg.use_height || ctx->cfg.use_normal;
            ctx->fr.have_lab = ctx->cfg.use_color;
            ctx->fr.have_integral = ctx->cfg.use_normal;
            ctx->fr.have_post = false;
            RSS_CU(ctx, cudaGraphLaunch(G.exec, ctx->s0));
            ctx->launches += G.launches;
            ran = true;
        } else if (want_graph && attempt == 0 && G.prev_alloc_gen == gen && G.prev_W == W && G.prev_H == H &&
                   same_params(G.prev_prm, *prm)) {
            // the previous (eager) keyframe had this signature and nothing has been allocated since: capture
            if (G.exec) { cudaGraphExecDestroy(G.exec); G.exec = nullptr; }
            const uint64_t l0 = ctx->launches;
            ctx->capturing = true;
            cudaError_t e = cudaStreamBeginCapture(ctx->s0, cudaStreamCaptureModeThreadLocal);
            rss_status cs = RSS_ERR_CUDA;
            cudaGraph_t graph = nullptr;
            if (e == cudaSuccess) {
                cs = keyframe_enqueue(ctx, crf, W, H, prm);
                e = cudaStreamEndCapture(ctx->s0, &graph);
            }
            ctx->capturing = false;
            if (e == cudaSuccess && cs == RSS_OK && device_alloc_events().load() == gen)
                e = cudaGraphInstantiate(&G.exec, graph, 0);
            else if (e == cudaSuccess)
                e = cudaErrorUnknown;
            if (graph) cudaGraphDestroy(graph);
            if (e == cudaSuccess) {
                G.W = W; G.H = H; G.prm = *prm; G.alloc_gen = gen; G.shared_gpu = shared_gpu; G.want_q = Qout != nullptr;
                G.launches = ctx->launches - l0;
                RSS_CU(ctx, cudaGraphLaunch(G.exec, ctx->s0));
                ran = true;
            } else {  // not capturable here (driver, or an allocation slipped in): stay eager on this context
                cudaGetLastError();
                G.exec = nullptr;
                G.broken = true;
                ctx->launches = l0;
            }
        }
        if (!ran) {
            st = keyframe_enqueue(ctx, crf, W, H, prm);
            if (st != RSS_OK) return st;
            staged = true;
        }
        G.prev_W = W; G.prev_H = H; G.prev_prm = *prm;
        if (labels) RSS_CU(ctx, cudaMemcpyAsync(labels, crf->labels.ptr, (size_t)N * F.L, cudaMemcpyDeviceToHost, ctx->s0));
        if (Qout) {
            float* dstQ = Qout;
            for (int l = 0; l < crf->n_layers; l++) {
                RSS_LAUNCH(ctx, deinterleave_kernel, rss_div_up((long long)N * crf->M[l], 256), 256, 0, ctx->s0,
                           crf->Q.as<float>(), N, crf->M[l], crf->Mp, crf->moff[l], crf->M[l], 0, crf->scratch.as<float>());
                RSS_CU(ctx, cudaMemcpyAsync(dstQ, crf->scratch.ptr, (size_t)N * crf->M[l] * 4, cudaMemcpyDeviceToHost, ctx->s0));
                RSS_CU(ctx, cudaStreamSynchronize(ctx->s0));
                dstQ += (size_t)N * crf->M[l];
            }
        }
        cudaEventRecord(ctx->ev[5], ctx->s0);
        RSS_CU(ctx, cudaStreamSynchronize(ctx->s0));
        G.prev_alloc_gen = device_alloc_events().load();
        // overflow flags (pinned copies made at the end of the device part)
        bool overflow = false;
        for (size_t k = 0; k < crf->kernels.size(); k++) {
            Lattice* L = crf->kernels[k];
            const uint32_t* h = ctx->pin_small.as<uint32_t>() + 2 * k;
            L->V_host = (int)h[0];
            if (h[1]) {
                const uint64_t maxv = (uint64_t)crf->N * (L->d + 1);
                if ((uint64_t)L->hcap >= 2 * next_pow2(2 * maxv) || L->hcap >= (1u << 30))
                    return ctx->fail(RSS_ERR_CAPACITY, "lattice hash table overflow");
                L->want_hcap = L->hcap * 4;
                overflow = true;
            }
        }
        if (!overflow) break;
        // rebuild with larger hash tables and run again, eagerly (rare); the captured graph no longer fits
        if (G.exec) { cudaGraphExecDestroy(G.exec); G.exec = nullptr; }
        G.prev_alloc_gen = ~0ull;
    }
    float ms = 0.f;
    ctx->tim = rss_timings{0, 0, 0, 0, 0, 0, 0, 0};
    if (staged) {
        frame_collect_timings(ctx, false);
        cudaEventElapsedTime(&ms, ctx->ev[8], ctx->ev[9]); ctx->tim.lattice_ms = ms;  // lattice time NOT hidden behind the frame path
        cudaEventElapsedTime(&ms, ctx->ev[9], ctx->ev[10]); ctx->tim.meanfield_ms = ms;
        cudaEventElapsedTime(&ms, ctx->ev[10], ctx->ev[5]); ctx->tim.d2h_ms = ms;
    } else {
        cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]); ctx->tim.h2d_ms = ms;
    }
    cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[5]); ctx->tim.total_ms = ms;
    return RSS_OK;
}

extern "C" rss_status rss_keyframe_lattice_info(rss_ctx* ctx, int k, int* d, int* vertices) {
    if (!ctx) return RSS_ERR_INVALID;
    rss_crf* crf = ctx->keyframe_crf;
    if (!crf || k < 0 || k >= (int)crf->kernels.size()) return ctx->fail(RSS_ERR_STATE, "no keyframe lattice with this index");
    if (d) *d = crf->kernels[k]->d;
    if (vertices) *vertices = crf->kernels[k]->V_host;
    return RSS_OK;
}
