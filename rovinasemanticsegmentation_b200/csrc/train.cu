// GPU forest training: libf::DecisionTreeLearner::learn + updateMultiHistograms and
// RandomForestLearner::learn of the reference (third-party/libforest/src/learning.cpp:410-916, 963-1012, 1031-1073),
// set up like src/train.cpp:225-249 (bootstrap, sqrt(D) features per node, multi-label layers, no class-frequency
// weighting of the split objective, inverse-class-frequency weighted, Laplace-smoothed, logged leaf histograms).
//
// The reference grows a tree node by node (depth first, one std::sort per node and candidate feature).  Here a tree is
// grown LEVEL BY LEVEL: all open nodes of a level are processed together, the samples of the level are kept grouped by
// node, and for every candidate-feature slot ONE radix sort of (node, feature value) keys orders every node's samples at
// once.  The split objective is the reference's (learning.cpp:27-340, EfficientEntropyHistogram):
//     E(side) = M * fastlog2(M) - sum_c n_c * fastlog2(n_c)        (ENTROPY(p) = -p * fastlog2(p), learning.cpp:15)
// evaluated from EXACT integer class counts (a prefix count over the sorted order) instead of the reference's incremental
// float updates, at every boundary the reference tests (value gap >= 1e-6, :572-579), with its tie rules (first position
// in scan order, first feature slot: strict '<', :585).  Threshold = (left + right) * 0.5f (:588, :600).
// What cannot be identical: the reference draws its randomness from std::random_device (unseeded: bootstrap, the layer of
// a node :490-491, the feature shuffle :540) - this learner is SEEDED and deterministic - and the low bits of the
// reference's incrementally accumulated entropies depend on std::sort's order among equal feature values.
// Leaf histograms ARE bit-identical to updateMultiHistograms for a given tree: the device counts (leaf, layer, class)
// occurrences exactly, the host then replays the reference's float accumulation (k additions of freq[c]), sum,
// smoothing and std::log (:984-1008) with the same libm.
// The radix sort is CUB's (cub::DeviceRadixSort, ships with the CUDA toolkit): library code, like cuBLAS would be.
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <fstream>
#include <numeric>
#include <stdexcept>

#include "common.cuh"
#include "kernels.hpp"

namespace rss {

constexpr int TR_CM = 16;        // classes per layer the trainer supports
constexpr int TR_BLOCK = 1024;   // sorted positions per block of the class-prefix scan (256 threads x 4)

// ---- deterministic randomness: splitmix64 of (seed, stream, counter)
__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__host__ __device__ __forceinline__ uint64_t rnd(uint64_t seed, uint64_t a, uint64_t b) {
    return splitmix64(splitmix64(seed ^ (a * 0xD1B54A32D192ED03ull)) ^ (b * 0x8CB92BA72F3D8DD7ull));
}
// order-preserving map float -> uint32
__device__ __forceinline__ uint32_t flipf(float v) {
    const uint32_t u = __float_as_uint(v);
    return u ^ ((u >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
__device__ __forceinline__ float unflipf(uint32_t u) {
    return __uint_as_float(u ^ ((u >> 31) ? 0x80000000u : 0xFFFFFFFFu));
}
// Paul Mineiro's fastlog2 (fastapprox), the approximation the reference's ENTROPY macro uses (libforest src/fastlog.h)
__device__ __forceinline__ float fastlog2_dev(float x) {
    const uint32_t i = __float_as_uint(x);
    const float m = __uint_as_float((i & 0x007FFFFFu) | 0x3f000000u);
    float y = __uint2float_rn(i);
    y = __fmul_rn(y, 1.1920928955078125e-7f);
    return __fsub_rn(__fsub_rn(__fsub_rn(y, 124.22551499f), __fmul_rn(1.498030302f, m)),
                     __fdiv_rn(1.72587999f, __fadd_rn(0.3520887068f, m)));
}
// M * fastlog2(M) - sum n_c * fastlog2(n_c): initEntropies (learning.cpp:276-289) on exact counts
__device__ __forceinline__ float side_entropy(const int (&n)[TR_CM], int C, int M) {
    if (M <= 0) return 0.f;
    float e = __fmul_rn((float)M, fastlog2_dev((float)M));  // -ENTROPY(mass)
    for (int c = 0; c < C; c++)
        if (n[c] > 0) e = __fsub_rn(e, __fmul_rn((float)n[c], fastlog2_dev((float)n[c])));  // += ENTROPY(n_c)
    return e;
}

// sample g * nb + i = the i-th bootstrap draw of the g-th tree of the pass (DataStorage::bootstrapmulti, data.cpp:325-349)
__global__ void __launch_bounds__(256) tr_bootstrap_kernel(int* __restrict__ boot, int nb, int G, int n, uint64_t seed, int tree0,
                                                           int use_bootstrap) {
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= (size_t)nb * G) return;
    const int g = (int)(j / nb), i = (int)(j - (size_t)g * nb);
    boot[j] = use_bootstrap ? (int)(rnd(seed, 0x1000 + tree0 + g, i) % (uint64_t)n) : i;
}
__global__ void __launch_bounds__(256) tr_iota_kernel(uint32_t* __restrict__ order, uint32_t* __restrict__ rank, int nb, int G) {
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < (size_t)nb * G) { order[j] = (uint32_t)j; rank[j] = (uint32_t)(j / nb); }
}
// class histogram of every open node for the label layer drawn for it (learning.cpp:493-518)
__global__ void __launch_bounds__(256) tr_node_hist_kernel(const uint32_t* __restrict__ order, const uint32_t* __restrict__ rank,
                                                           int n_active, const int* __restrict__ boot,
                                                           const int* __restrict__ labels, int L,
                                                           const int* __restrict__ node_layer, int* __restrict__ hist) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = p < n_active;
    unsigned slot = 0xffffffffu;
    if (live) {
        const uint32_t k = rank[p];
        const int c = labels[(size_t)boot[order[p]] * L + node_layer[k]];
        slot = k * TR_CM + (unsigned)c;
    }
    const unsigned peers = __match_any_sync(0xffffffffu, slot);  // warp-aggregated: the root has 10^5 samples on <= 16 counters
    if (live && (threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(hist + slot, __popc(peers));
}
// Sort keys of ALL candidate-feature slots of a level at once: position (f, p) gets the key (f * (S + 1) + s, value of
// the f-th candidate feature of node s), s = rank of p's node among the S splittable nodes; samples of nodes that became
// leaves get node id S and sort to the end of their slot's block.  One radix sort then orders every (slot, node) list.
__global__ void __launch_bounds__(256) tr_make_keys_kernel(const uint32_t* __restrict__ order, const uint32_t* __restrict__ rank,
                                                           int n_active, const int* __restrict__ boot,
                                                           const float* __restrict__ feats, int D,
                                                           const int* __restrict__ srank, const int* __restrict__ featsel,
                                                           int F, int S, uint64_t* __restrict__ keys,
                                                           uint32_t* __restrict__ vals) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x, f = blockIdx.y;
    if (p >= n_active) return;
    const int s = srank[rank[p]];
    const uint32_t id = order[p];
    const size_t q = (size_t)f * n_active + p;
    vals[q] = id;
    const uint64_t hi = (uint64_t)f * (uint32_t)(S + 1);
    if (s < 0) { keys[q] = (hi + (uint32_t)S) << 32; return; }
    const float v = feats[(size_t)boot[id] * D + featsel[s * F + f]];
    keys[q] = ((hi + (uint32_t)s) << 32) | flipf(v);
}
// class of every sorted position (for the layer of its node); 255 = not part of any list
__global__ void __launch_bounds__(256) tr_gather_cls_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                                                            size_t Q, int S, const int* __restrict__ boot,
                                                            const int* __restrict__ labels, int L,
                                                            const int* __restrict__ s_layer, uint8_t* __restrict__ cls) {
    const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= Q) return;
    const int s = (int)((uint32_t)(keys[q] >> 32) % (uint32_t)(S + 1));
    cls[q] = s == S ? (uint8_t)255 : (uint8_t)labels[(size_t)boot[vals[q]] * L + s_layer[s]];
}
__global__ void __launch_bounds__(256) tr_block_hist_kernel(const uint8_t* __restrict__ cls, size_t n, int* __restrict__ blockhist) {
    __shared__ int h[TR_CM];
    if (threadIdx.x < TR_CM) h[threadIdx.x] = 0;
    __syncthreads();
    const size_t base = (size_t)blockIdx.x * TR_BLOCK;
    for (int i = threadIdx.x; i < TR_BLOCK && base + i < n; i += 256) {
        const int c = cls[base + i];
        if (c < TR_CM) atomicAdd(&h[c], 1);
    }
    __syncthreads();
    if (threadIdx.x < TR_CM) blockhist[blockIdx.x * TR_CM + threadIdx.x] = h[threadIdx.x];
}
// exclusive scan over the blocks: one CTA per class, a thread owns a run of consecutive blocks
__global__ void __launch_bounds__(1024) tr_block_scan_kernel(int* __restrict__ blockhist, int nblocks) {
    __shared__ int wsum[32];
    const int c = blockIdx.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int per = (nblocks + 1023) / 1024, b0 = min(nblocks, (int)threadIdx.x * per), b1 = min(nblocks, b0 + per);
    int sum = 0;
    for (int b = b0; b < b1; b++) sum += blockhist[(size_t)b * TR_CM + c];
    int inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += u;
    }
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    if (w == 0) {
        const int v = wsum[lane];
        int iv = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, iv, o);
            if (lane >= o) iv += u;
        }
        wsum[lane] = iv - v;
    }
    __syncthreads();
    int run = wsum[w] + inc - sum;
    for (int b = b0; b < b1; b++) {
        const int v = blockhist[(size_t)b * TR_CM + c];
        blockhist[(size_t)b * TR_CM + c] = run;
        run += v;
    }
}
// class counts in front of every (slot, node) list: one warp per list; list (f, s) starts at f * n_active + sstart[s]
__global__ void __launch_bounds__(256) tr_node_base_kernel(const uint8_t* __restrict__ cls, const int* __restrict__ blockbase,
                                                           const int* __restrict__ sstart, int S, int F, int n_active,
                                                           int* __restrict__ nodebase) {
    const int id = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (id >= S * F) return;
    const int f = id / S, s = id - f * S;
    const size_t q = (size_t)f * n_active + sstart[s], b = q / TR_BLOCK;
    int cnt[TR_CM];
#pragma unroll
    for (int c = 0; c < TR_CM; c++) cnt[c] = 0;
    for (size_t i = b * TR_BLOCK + lane; i < q; i += 32) {
        const int cl = cls[i];
#pragma unroll
        for (int c = 0; c < TR_CM; c++) cnt[c] += (cl == c);
    }
#pragma unroll
    for (int c = 0; c < TR_CM; c++) {
        const int t = __reduce_add_sync(0xffffffffu, cnt[c]);
        if (lane == 0) nodebase[(size_t)id * TR_CM + c] = blockbase[b * TR_CM + c] + t;
    }
}
// objective at every boundary of the sorted order; per-node minimum (earliest position on ties) by a 64-bit atomicMin
__global__ void __launch_bounds__(256) tr_objective_kernel(const uint64_t* __restrict__ keys, const uint8_t* __restrict__ cls,
                                                           size_t n_split, int S, int n_active,
                                                           const int* __restrict__ blockbase, const int* __restrict__ nodebase,
                                                           const int* __restrict__ sstart, const int* __restrict__ s_hist,
                                                           const int* __restrict__ s_C, unsigned long long* __restrict__ best64) {
    __shared__ int pre[TR_CM][257];  // inclusive prefix of the per-thread class counts, per class
    const size_t base = (size_t)blockIdx.x * TR_BLOCK;
    const int t = threadIdx.x;
    int mine[4];
    int cnt[TR_CM];
#pragma unroll
    for (int c = 0; c < TR_CM; c++) cnt[c] = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const size_t p = base + 4 * t + j;
        mine[j] = p < n_split ? (int)cls[p] : -1;
#pragma unroll
        for (int c = 0; c < TR_CM; c++) cnt[c] += (mine[j] == c);
    }
    // block scan of the 16 counters: warp shuffles, then the warp totals through shared memory
    const int lane = t & 31, w = t >> 5;
    __shared__ int wtot[TR_CM][8];
#pragma unroll
    for (int c = 0; c < TR_CM; c++) {
        int v = cnt[c];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += u;
        }
        pre[c][t] = v;
        if (lane == 31) wtot[c][w] = v;
    }
    __syncthreads();
    if (t < TR_CM) {
        int run = 0;
        for (int k = 0; k < 8; k++) { const int v = wtot[t][k]; wtot[t][k] = run; run += v; }
    }
    __syncthreads();
    int run[TR_CM];  // class counts of positions [0, first position of this thread) of the whole sorted array
#pragma unroll
    for (int c = 0; c < TR_CM; c++) run[c] = blockbase[blockIdx.x * TR_CM + c] + wtot[c][w] + pre[c][t] - cnt[c];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const size_t p = base + 4 * t + j;
        if (p >= n_split) break;
#pragma unroll
        for (int c = 0; c < TR_CM; c++) run[c] += (mine[j] == c);  // inclusive of position p
        if (p + 1 >= n_split || mine[j] >= TR_CM) continue;
        const uint64_t k0 = keys[p], k1 = keys[p + 1];
        const uint32_t id = (uint32_t)(k0 >> 32);
        if ((uint32_t)(k1 >> 32) != id) continue;  // last sample of its list
        const int f = (int)(id / (uint32_t)(S + 1)), s = (int)(id - (uint32_t)f * (uint32_t)(S + 1));
        const float lv = unflipf((uint32_t)k0), rv = unflipf((uint32_t)k1);
        if (__fsub_rn(rv, lv) < 1e-6f) continue;  // learning.cpp:572-579 (a NaN gap passes, like in the reference)
        const int C = s_C[s];
        const size_t q = (size_t)f * n_active + sstart[s];
        const int list = f * S + s;
        int nl[TR_CM], nr[TR_CM];
#pragma unroll
        for (int c = 0; c < TR_CM; c++) {
            nl[c] = run[c] - nodebase[(size_t)list * TR_CM + c];
            nr[c] = s_hist[s * TR_CM + c] - nl[c];
        }
        const int ML = (int)(p - q) + 1;
        int MR = 0;
        for (int c = 0; c < C; c++) MR += nr[c];
        const float obj = __fadd_rn(side_entropy(nl, C, ML), side_entropy(nr, C, MR));
        atomicMin(best64 + list, ((unsigned long long)flipf(obj) << 32) | (unsigned)(p - q));
    }
}
struct TrBest {
    float obj, thr;
    int feat, left;
};
// per node: the winner over the candidate slots in slot order (strict '<', learning.cpp:585: the earlier slot wins ties)
__global__ void __launch_bounds__(256) tr_merge_best_kernel(const unsigned long long* __restrict__ best64,
                                                            const uint64_t* __restrict__ keys, const int* __restrict__ sstart,
                                                            const int* __restrict__ featsel, int F, int S, int n_active,
                                                            TrBest* __restrict__ best) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    TrBest r{1e35f, 0.f, -1, 0};  // learning.cpp:531-536
    for (int f = 0; f < F; f++) {
        const unsigned long long b = best64[(size_t)f * S + s];
        if (b == ~0ull) continue;
        const float obj = unflipf((uint32_t)(b >> 32));
        if (!(obj < r.obj)) continue;
        const int pos = (int)(b & 0xffffffffu);
        const size_t p = (size_t)f * n_active + sstart[s] + pos;
        const float lv = unflipf((uint32_t)keys[p]), rv = unflipf((uint32_t)keys[p + 1]);
        r.obj = obj;
        r.thr = __fmul_rn(__fadd_rn(lv, rv), 0.5f);
        r.feat = featsel[s * F + f];
        r.left = pos + 1;
    }
    best[s] = r;
}
// new node rank of every active sample after the level's splits (featureValue < threshold goes left, :620-631)
__global__ void __launch_bounds__(256) tr_route_kernel(const uint32_t* __restrict__ order, const uint32_t* __restrict__ rank,
                                                       int n_active, const int* __restrict__ boot, const float* __restrict__ feats,
                                                       int D, const int* __restrict__ split_feat, const float* __restrict__ split_thr,
                                                       const int* __restrict__ child_rank, unsigned drop,
                                                       uint32_t* __restrict__ keys32, uint32_t* __restrict__ vals,
                                                       int* __restrict__ child_count) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = p < n_active;
    unsigned key = drop;  // = number of nodes of the next level: samples of leaves sort last
    if (live) {
        const uint32_t k = rank[p], id = order[p];
        vals[p] = id;
        const int cr = child_rank[k];  // rank of the LEFT child on the next level, or -1 when the node stays a leaf
        if (cr >= 0) {
            const float v = feats[(size_t)boot[id] * D + split_feat[k]];
            key = (unsigned)cr + (v < split_thr[k] ? 0u : 1u);
        }
        keys32[p] = key;
    }
    const unsigned peers = __match_any_sync(0xffffffffu, key);
    if (live && key != drop && (threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(child_count + key, __popc(peers));
}
// leaf of every training example (DecisionTree::findLeafNode, classifier.cpp:97-117) and exact (leaf, layer, class) counts
__global__ void __launch_bounds__(256) tr_leaf_count_kernel(const float* __restrict__ feats, int n, int D, const int* __restrict__ labels,
                                                            int L, const int* __restrict__ nfeat, const float* __restrict__ nthr,
                                                            const int* __restrict__ nleft, const int* __restrict__ leaf_row,
                                                            int* __restrict__ counts) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int node = 0;
    while (nleft[node] > 0) node = feats[(size_t)i * D + nfeat[node]] < nthr[node] ? nleft[node] : nleft[node] + 1;
    const int row = leaf_row[node];
    for (int l = 0; l < L; l++) atomicAdd(counts + ((size_t)row * L + l) * TR_CM + labels[(size_t)i * L + l], 1);
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
namespace {
struct HostTree {
    std::vector<int> feat, left, depth;
    std::vector<float> thr;
    std::vector<std::vector<std::vector<float>>> multi;  // per node: per layer histogram (leaves only)
    int add(int d) {
        feat.push_back(0); thr.push_back(0.f); left.push_back(0); depth.push_back(d);
        return (int)feat.size() - 1;
    }
};
template <class T>
void wr(std::ostream& os, const T& v) { os.write(reinterpret_cast<const char*>(&v), sizeof(T)); }
// DecisionTree::write / RandomForest::write (classifier.cpp:144-152, 210-220; writeBinary of io.h:43-108)
void write_forest(std::ostream& os, const std::vector<HostTree>& trees) {
    wr<int>(os, (int)trees.size());
    for (const HostTree& t : trees) {
        const int n = (int)t.feat.size();
        wr<int>(os, n); os.write((const char*)t.feat.data(), (size_t)n * 4);
        wr<int>(os, n); os.write((const char*)t.thr.data(), (size_t)n * 4);
        wr<int>(os, n); os.write((const char*)t.left.data(), (size_t)n * 4);
        wr<int>(os, n);
        for (int i = 0; i < n; i++) wr<int>(os, 0);  // plain histograms stay empty in multi-label mode (:529, :611)
        wr<int>(os, n);
        for (int i = 0; i < n; i++) {
            wr<int>(os, (int)t.multi[i].size());
            for (const std::vector<float>& h : t.multi[i]) {
                wr<int>(os, (int)h.size());
                os.write((const char*)h.data(), h.size() * 4);
            }
        }
    }
}
#define TR_CU(call)                                                                              \
    do {                                                                                         \
        const cudaError_t e__ = (call);                                                          \
        if (e__ != cudaSuccess) throw std::runtime_error(std::string("CUDA: ") + cudaGetErrorString(e__)); \
    } while (0)
template <class T>
void up(DevBuf& b, const std::vector<T>& v, cudaStream_t st) {
    TR_CU(b.reserve(std::max<size_t>(v.size(), 1) * sizeof(T)));
    if (!v.empty()) TR_CU(cudaMemcpyAsync(b.ptr, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, st));
}
template <class T>
void down(std::vector<T>& v, const DevBuf& b, size_t n, cudaStream_t st) {
    v.resize(n);
    if (n) TR_CU(cudaMemcpyAsync(v.data(), b.ptr, n * sizeof(T), cudaMemcpyDeviceToHost, st));
    TR_CU(cudaStreamSynchronize(st));
}
inline int bits_for(unsigned v) { int b = 1; while ((1u << b) <= v && b < 31) b++; return b; }
}  // namespace

rss_status forest_train(rss_ctx* ctx, const float* feats_h, int n, int D, const int32_t* labels_h, int L, const int* class_counts,
                        const rss_train_params& prm, const char* out_path, rss_train_stats* stats) {
    cudaStream_t st = ctx->s0;
    try {
        for (int l = 0; l < L; l++)
            if (class_counts[l] < 1 || class_counts[l] > TR_CM) return ctx->fail(RSS_ERR_INVALID, "train: 1..16 classes per layer");
        for (size_t i = 0; i < (size_t)n * L; i++)
            if (labels_h[i] < 0 || labels_h[i] >= class_counts[i % L]) return ctx->fail(RSS_ERR_INVALID, "train: label out of range");
        const int F = prm.num_features > 0 ? std::min(prm.num_features, D) : (int)std::ceil(std::sqrt((double)D));  // autoconf, :363-368
        const int nb = prm.use_bootstrap ? (prm.num_bootstrap_examples > 0 ? prm.num_bootstrap_examples : n) : n;
        const int NT = prm.num_trees;
        // Trees are independent (RandomForestLearner::learn, :1031-1073, one OpenMP task per tree): all of them are grown
        // together, level by level - the open nodes of every tree form one list, a sample of tree t has the id t * nb + i.
        // Memory bounds how many trees share a pass: the sort buffers hold F keys per active sample.
        const size_t per_tree = (size_t)nb * F * (8 + 8 + 4 + 4 + 1) + (size_t)nb * 24;
        size_t free_b = 0, total_b = 0;
        TR_CU(cudaMemGetInfo(&free_b, &total_b));
        const size_t budget = free_b / 2 > (size_t)n * D * 4 ? free_b / 2 - (size_t)n * D * 4 : 0;
        const int group = (int)std::max<size_t>(1, std::min<size_t>((size_t)NT, budget / std::max<size_t>(per_tree, 1)));
        const size_t NS = (size_t)group * nb;  // samples of a pass
        if (NS * F >= (1ull << 31)) return ctx->fail(RSS_ERR_CAPACITY, "train: too many (sample, feature) pairs per pass");
        DevBuf d_feats, d_labels, d_boot, d_order, d_rank, d_rank2, d_keys, d_keys2, d_vals, d_vals2, d_cls, d_hist, d_layer,
            d_srank, d_featsel, d_slayer, d_sstart, d_shist, d_sC, d_blockhist, d_nodebase, d_best, d_best64, d_sfeat, d_sthr,
            d_child, d_ccount, d_tmp, d_nfeat, d_nthr, d_nleft, d_leafrow, d_counts;
        DevBuf* all[] = {&d_feats, &d_labels, &d_boot, &d_order, &d_rank, &d_rank2, &d_keys, &d_keys2, &d_vals, &d_vals2, &d_cls,
                         &d_hist, &d_layer, &d_srank, &d_featsel, &d_slayer, &d_sstart, &d_shist, &d_sC, &d_blockhist,
                         &d_nodebase, &d_best, &d_best64, &d_sfeat, &d_sthr, &d_child, &d_ccount, &d_tmp, &d_nfeat, &d_nthr,
                         &d_nleft, &d_leafrow, &d_counts};
        struct Free { DevBuf** b; size_t n; ~Free() { for (size_t i = 0; i < n; i++) b[i]->release(); } } guard{all, sizeof(all) / sizeof(all[0])};
        TR_CU(d_feats.reserve((size_t)n * D * 4));
        TR_CU(d_labels.reserve((size_t)n * L * 4));
        TR_CU(cudaMemcpyAsync(d_feats.ptr, feats_h, (size_t)n * D * 4, cudaMemcpyHostToDevice, st));
        TR_CU(cudaMemcpyAsync(d_labels.ptr, labels_h, (size_t)n * L * 4, cudaMemcpyHostToDevice, st));
        for (DevBuf* b : {&d_boot, &d_order, &d_rank, &d_rank2}) TR_CU(b->reserve(NS * 4));
        const size_t QMAX = NS * F;
        TR_CU(d_keys.reserve(QMAX * 8));
        TR_CU(d_keys2.reserve(QMAX * 8));
        TR_CU(d_vals.reserve(QMAX * 4));
        TR_CU(d_vals2.reserve(QMAX * 4));
        TR_CU(d_cls.reserve(QMAX));
        TR_CU(d_blockhist.reserve((QMAX / TR_BLOCK + 2) * TR_CM * 4));
        size_t tmp_bytes = 0, tmp2 = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr, (const uint32_t*)nullptr,
                                        (uint32_t*)nullptr, (int)QMAX, 0, 64, st);
        cub::DeviceRadixSort::SortPairs(nullptr, tmp2, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const uint32_t*)nullptr,
                                        (uint32_t*)nullptr, (int)NS, 0, 32, st);
        TR_CU(d_tmp.reserve(std::max(tmp_bytes, tmp2)));
        // inverse class frequencies over ALL training examples (data.h:359-370): freq[c] = N / count(c), int / float
        std::vector<std::vector<float>> freq(L);
        for (int l = 0; l < L; l++) {
            freq[l].assign(class_counts[l], 0.f);
            for (int i = 0; i < n; i++) freq[l][labels_h[(size_t)i * L + l]]++;
            for (int c = 0; c < class_counts[l]; c++) freq[l][c] = n / freq[l][c];
        }
        std::vector<HostTree> trees(NT);
        long long total_nodes = 0, total_levels = 0;
        for (int t0 = 0; t0 < NT; t0 += group) {
            const int G = std::min(group, NT - t0);
            const size_t ns = (size_t)G * nb;
            tr_bootstrap_kernel<<<rss_div_up((long long)ns, 256), 256, 0, st>>>(d_boot.as<int>(), nb, G, n, prm.seed, t0, prm.use_bootstrap);
            tr_iota_kernel<<<rss_div_up((long long)ns, 256), 256, 0, st>>>(d_order.as<uint32_t>(), d_rank.as<uint32_t>(), nb, G);
            struct Open { int tree, node; };
            std::vector<Open> open;  // open nodes of the level, in rank order
            std::vector<int> seg_len;
            for (int g = 0; g < G; g++) { trees[t0 + g].add(0); open.push_back(Open{t0 + g, 0}); seg_len.push_back(nb); }
            int n_active = (int)ns;
            for (int level = 0; !open.empty(); level++, total_levels++) {
                const int K = (int)open.size();
                // (1) label layer per node (:490-491) and class histograms
                std::vector<int> layer(K);
                for (int k = 0; k < K; k++)
                    layer[k] = L > 1 ? (int)(rnd(prm.seed, 0x2000 + open[k].tree, open[k].node) % (uint64_t)L) : 0;
                up(d_layer, layer, st);
                TR_CU(d_hist.reserve((size_t)K * TR_CM * 4));
                TR_CU(cudaMemsetAsync(d_hist.ptr, 0, (size_t)K * TR_CM * 4, st));
                tr_node_hist_kernel<<<rss_div_up(n_active, 256), 256, 0, st>>>(d_order.as<uint32_t>(), d_rank.as<uint32_t>(), n_active,
                                                                              d_boot.as<int>(), d_labels.as<int>(), L, d_layer.as<int>(),
                                                                              d_hist.as<int>());
                std::vector<int> hist;
                down(hist, d_hist, (size_t)K * TR_CM, st);
                // (2) stop rules (:520-530): too few examples, pure, too deep
                std::vector<int> srank(K, -1), s_node, s_layer, s_C, s_hist, sstart;
                int S = 0, n_split = 0;
                for (int k = 0; k < K; k++) {
                    const int C = class_counts[layer[k]];
                    int mass = 0, nonzero = 0;
                    for (int c = 0; c < C; c++) { mass += hist[k * TR_CM + c]; nonzero += hist[k * TR_CM + c] > 0; }
                    if (mass < prm.min_split_examples || nonzero <= 1 || trees[open[k].tree].depth[open[k].node] > prm.max_depth) continue;
                    srank[k] = S++;
                    s_node.push_back(k); s_layer.push_back(layer[k]); s_C.push_back(C);
                    s_hist.insert(s_hist.end(), hist.begin() + (size_t)k * TR_CM, hist.begin() + (size_t)(k + 1) * TR_CM);
                    sstart.push_back(n_split);
                    n_split += seg_len[k];
                }
                if (S == 0) break;
                // (3) candidate features: numFeatures draws without replacement per node (:540: shuffle, take the first F)
                std::vector<int> featsel((size_t)S * F), perm(D);
                for (int s = 0; s < S; s++) {
                    std::iota(perm.begin(), perm.end(), 0);
                    const Open& o = open[s_node[s]];
                    for (int j = 0; j < F; j++) {  // partial Fisher-Yates
                        const int r = j + (int)(rnd(prm.seed, 0x3000 + o.tree, (uint64_t)o.node * 4096 + j) % (uint64_t)(D - j));
                        std::swap(perm[j], perm[r]);
                        featsel[(size_t)s * F + j] = perm[j];
                    }
                }
                up(d_srank, srank, st); up(d_featsel, featsel, st); up(d_slayer, s_layer, st); up(d_sstart, sstart, st);
                up(d_shist, s_hist, st); up(d_sC, s_C, st);
                TR_CU(d_best.reserve((size_t)S * sizeof(TrBest)));
                TR_CU(d_best64.reserve((size_t)S * F * 8));
                TR_CU(d_nodebase.reserve((size_t)S * F * TR_CM * 4));
                TR_CU(cudaMemsetAsync(d_best64.ptr, 0xFF, (size_t)S * F * 8, st));
                // (3a) one sort for every (slot, node) list of the level; positions beyond Qs are samples of finished nodes
                const size_t Q = (size_t)F * n_active;
                const size_t Qs = (size_t)(F - 1) * n_active + n_split;  // the last list ends here
                const int key_bits = 32 + bits_for((unsigned)((S + 1) * F));
                tr_make_keys_kernel<<<dim3(rss_div_up(n_active, 256), F), 256, 0, st>>>(
                    d_order.as<uint32_t>(), d_rank.as<uint32_t>(), n_active, d_boot.as<int>(), d_feats.as<float>(), D,
                    d_srank.as<int>(), d_featsel.as<int>(), F, S, d_keys.as<uint64_t>(), d_vals.as<uint32_t>());
                size_t tb = d_tmp.cap;
                TR_CU(cub::DeviceRadixSort::SortPairs(d_tmp.ptr, tb, d_keys.as<uint64_t>(), d_keys2.as<uint64_t>(), d_vals.as<uint32_t>(),
                                                      d_vals2.as<uint32_t>(), (int)Q, 0, key_bits, st));
                const int nblk = (int)((Qs + TR_BLOCK - 1) / TR_BLOCK);
                tr_gather_cls_kernel<<<rss_div_up((long long)Qs, 256), 256, 0, st>>>(d_keys2.as<uint64_t>(), d_vals2.as<uint32_t>(), Qs, S,
                                                                                      d_boot.as<int>(), d_labels.as<int>(), L,
                                                                                      d_slayer.as<int>(), d_cls.as<uint8_t>());
                tr_block_hist_kernel<<<nblk, 256, 0, st>>>(d_cls.as<uint8_t>(), Qs, d_blockhist.as<int>());
                tr_block_scan_kernel<<<TR_CM, 1024, 0, st>>>(d_blockhist.as<int>(), nblk);
                tr_node_base_kernel<<<rss_div_up((long long)S * F * 32, 256), 256, 0, st>>>(d_cls.as<uint8_t>(), d_blockhist.as<int>(),
                                                                                            d_sstart.as<int>(), S, F, n_active,
                                                                                            d_nodebase.as<int>());
                tr_objective_kernel<<<nblk, 256, 0, st>>>(d_keys2.as<uint64_t>(), d_cls.as<uint8_t>(), Qs, S, n_active,
                                                          d_blockhist.as<int>(), d_nodebase.as<int>(), d_sstart.as<int>(),
                                                          d_shist.as<int>(), d_sC.as<int>(), d_best64.as<unsigned long long>());
                tr_merge_best_kernel<<<rss_div_up(S, 256), 256, 0, st>>>(d_best64.as<unsigned long long>(), d_keys2.as<uint64_t>(),
                                                                         d_sstart.as<int>(), d_featsel.as<int>(), F, S, n_active,
                                                                         d_best.as<TrBest>());
                std::vector<TrBest> best;
                down(best, d_best, (size_t)S, st);
                // (4) split (:603-645): children are appended in rank order, left child first
                std::vector<int> split_feat(K, 0), child_rank(K, -1);
                std::vector<float> split_thr(K, 0.f);
                std::vector<Open> next_open;
                for (int s = 0; s < S; s++) {
                    const int k = s_node[s];
                    HostTree& T = trees[open[k].tree];
                    const int node = open[k].node;
                    const int leftm = best[s].left, rightm = seg_len[k] - best[s].left;
                    if (best[s].feat < 0 || leftm < prm.min_child_split_examples || rightm < prm.min_child_split_examples) continue;
                    split_feat[k] = best[s].feat; split_thr[k] = best[s].thr;
                    child_rank[k] = (int)next_open.size();
                    const int lc = T.add(T.depth[node] + 1);
                    T.add(T.depth[node] + 1);
                    T.feat[node] = best[s].feat; T.thr[node] = best[s].thr; T.left[node] = lc;
                    next_open.push_back(Open{open[k].tree, lc}); next_open.push_back(Open{open[k].tree, lc + 1});
                }
                if (next_open.empty()) break;
                up(d_sfeat, split_feat, st); up(d_sthr, split_thr, st); up(d_child, child_rank, st);
                const int K2 = (int)next_open.size();
                TR_CU(d_ccount.reserve((size_t)K2 * 4));
                TR_CU(cudaMemsetAsync(d_ccount.ptr, 0, (size_t)K2 * 4, st));
                tr_route_kernel<<<rss_div_up(n_active, 256), 256, 0, st>>>(d_order.as<uint32_t>(), d_rank.as<uint32_t>(), n_active,
                                                                          d_boot.as<int>(), d_feats.as<float>(), D, d_sfeat.as<int>(),
                                                                          d_sthr.as<float>(), d_child.as<int>(), (unsigned)K2,
                                                                          d_rank2.as<uint32_t>(), d_vals.as<uint32_t>(), d_ccount.as<int>());
                tb = d_tmp.cap;
                TR_CU(cub::DeviceRadixSort::SortPairs(d_tmp.ptr, tb, d_rank2.as<uint32_t>(), d_rank.as<uint32_t>(), d_vals.as<uint32_t>(),
                                                      d_order.as<uint32_t>(), n_active, 0, bits_for((unsigned)K2), st));
                std::vector<int> cc;
                down(cc, d_ccount, (size_t)K2, st);
                seg_len = cc;
                n_active = std::accumulate(cc.begin(), cc.end(), 0);
                open.swap(next_open);
            }
        }
        // (5) leaf histograms from ALL training examples (updateMultiHistograms, :963-1012)
        for (int tr = 0; tr < NT; tr++) {
            HostTree& T = trees[tr];
            const int nn = (int)T.feat.size();
            std::vector<int> leaf_row(nn, -1);
            int leaves = 0;
            for (int i = 0; i < nn; i++)
                if (T.left[i] == 0) leaf_row[i] = leaves++;
            up(d_nfeat, T.feat, st); up(d_nthr, T.thr, st); up(d_nleft, T.left, st); up(d_leafrow, leaf_row, st);
            TR_CU(d_counts.reserve((size_t)leaves * L * TR_CM * 4));
            TR_CU(cudaMemsetAsync(d_counts.ptr, 0, (size_t)leaves * L * TR_CM * 4, st));
            tr_leaf_count_kernel<<<rss_div_up(n, 256), 256, 0, st>>>(d_feats.as<float>(), n, D, d_labels.as<int>(), L, d_nfeat.as<int>(),
                                                                    d_nthr.as<float>(), d_nleft.as<int>(), d_leafrow.as<int>(),
                                                                    d_counts.as<int>());
            std::vector<int> counts;
            down(counts, d_counts, (size_t)leaves * L * TR_CM, st);
            T.multi.assign(nn, {});
            for (int i = 0; i < nn; i++) {
                if (T.left[i] != 0) continue;
                T.multi[i].resize(L);
                for (int l = 0; l < L; l++) {
                    const int C = class_counts[l];
                    std::vector<float>& h = T.multi[i][l];
                    h.assign(C, 0.f);
                    for (int c = 0; c < C; c++) {
                        const int k = counts[((size_t)leaf_row[i] * L + l) * TR_CM + c];
                        float acc = 0.f;
                        const float fr = freq[l][c];
                        for (int r = 0; r < k; r++) acc += fr;  // the reference adds freq[c] once per example (:984-990)
                        h[c] = acc;
                    }
                    float total = 0;
                    for (int c = 0; c < C; c++) total += h[c];
                    for (int c = 0; c < C; c++) h[c] = std::log((h[c] + prm.smoothing) / (total + C * prm.smoothing));  // :1001-1004
                }
            }
            total_nodes += nn;
        }
        TR_CU(cudaGetLastError());
        if (out_path) {
            std::ofstream os(out_path, std::ios::binary);
            if (!os) return ctx->fail(RSS_ERR_IO, std::string("cannot write ") + out_path);
            write_forest(os, trees);
        }
        if (stats) {
            stats->trees = NT;
            stats->nodes = total_nodes;
            stats->levels = total_levels;
            stats->features_per_node = F;
            stats->bootstrap_examples = nb;
        }
        return RSS_OK;
    } catch (const std::exception& e) {
        return ctx->fail(RSS_ERR_CUDA, std::string("train: ") + e.what());
    }
}

}  // namespace rss
