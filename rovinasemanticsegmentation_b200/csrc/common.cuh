// Shared declarations of librss (host side): context, device buffers, launch accounting.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <cstdio>
#include <map>
#include <set>
#include <string>
#include <utility>
#include <vector>

#include "../../include/rss.h"

namespace rss {

// number of device (re)allocations so far in this process: a captured CUDA graph holds raw pointers, so it is only
// replayed while this counter has not moved since the capture
inline std::atomic<uint64_t>& device_alloc_events() {
    static std::atomic<uint64_t> n{0};
    return n;
}

// live contexts per device: with several keyframes in flight on one GPU the cooperative blur runs with a smaller CTA, so
// that the other keyframes' kernels fit beside it on the SMs (measured: +10 % keyframes/s at 3 in flight, -6 % alone)
inline std::atomic<int>& live_contexts(int device) {
    static std::atomic<int> n[64];
    return n[device & 63];
}

// grow-only device / pinned-host buffers: no cudaMalloc on the steady-state path
struct DevBuf {
    void* ptr = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        device_alloc_events()++;
        cudaError_t e = cudaMalloc(&ptr, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        cap = 0;
    }
    template <class T> T* as() const { return reinterpret_cast<T*>(ptr); }
};
struct PinBuf {
    void* ptr = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (ptr) cudaFreeHost(ptr);
        ptr = nullptr;
        cap = 0;
        cudaError_t e = cudaMallocHost(&ptr, bytes + 256);
        if (e == cudaSuccess) cap = bytes + 256;
        return e;
    }
    void release() {
        if (ptr) cudaFreeHost(ptr);
        ptr = nullptr;
        cap = 0;
    }
    template <class T> T* as() const { return reinterpret_cast<T*>(ptr); }
};

// one flattened tree node: 16 bytes -> one LDG.128 per traversal level
struct __align__(16) Node {
    int feat;    // splitFeatures[node]
    float thr;   // thresholds[node]
    int left;    // leftChild[node] (0 = leaf), right child = left + 1 (libforest classifiers.h:169-172)
    int leaf;    // row of the dense leaf table (valid when left == 0)
};

// fixed-point coefficients of cv::resize(INTER_LINEAR, 8U) for one destination index of one ROI size
struct __align__(8) ResizeTap {
    short i0, i1;  // the two source indices (already clipped to the ROI)
    short w0, w1;  // 11-bit weights
};

// camera pose of the current frame as the kernels read it (device copy: rss_ctx::pose_dev)
struct PoseParams {
    float M[9];  // R * Kinv
    float t[3];  // camera centre
};

struct HostConfig {
    bool use_color = true, use_depth = true, use_height = true, use_normal = true;
    int patch_size = 77, patch_size_reduce = 11;
    int layer_count = 0;
    int class_counts[RSS_MAX_LAYERS] = {0};
    int unknown_label[RSS_MAX_LAYERS] = {0};
    bool use_dense_crf = false;
    float dcrf_xyz = 0.5f, dcrf_rgb = 4.0f, dcrf_w = 10.0f;
    int dcrf_iters = 10, rf_stride = 2;
    float depth_min = 0.5f, depth_max = 15.0f;
    int feature_length() const {
        int D = 0;
        if (use_color) D += patch_size_reduce * patch_size_reduce * 3;
        if (use_depth) D += 1;
        if (use_height) D += 1;
        if (use_normal) D += 1;
        return D;
    }
};

struct ForestDev {
    int T = 0, L = 0, sumC = 0, total_nodes = 0, total_leaves = 0;
    int C[RSS_MAX_LAYERS] = {0};
    std::vector<int> tree_off;  // node offset of each tree (host copy)
    DevBuf nodes;               // Node[total_nodes]
    DevBuf tree_off_dev;        // int[T+1]
    DevBuf leaves;              // float[total_leaves][sumC]
    bool loaded = false;
};

struct FrameState {
    int W = 0, H = 0, stride = 0, gw = 0, gh = 0;
    int n_samples = -1;       // compacted sample count of the last extract (-1 = none)
    bool have_feats = false;  // materialised feature matrix valid
    bool have_post = false;   // full-resolution posteriors valid
    bool have_cloud = false, have_lab = false, have_integral = false;
    DevBuf rgb, depth, labels;
    DevBuf lab;                // uchar4 (H+2P) x (W+2P)
    DevBuf xyz;                // float4 H x W
    DevBuf dist_a, dist_b;     // float H x W (forward / final chamfer map)
    DevBuf integ;              // double [2][(H+1)*(W+1)][3]
    DevBuf integ_cnt;          // int    [2][(H+1)*(W+1)]
    DevBuf normals;            // float H*W*3 (debug / parity only)
    DevBuf grad, fin;          // float [6][H*W] gradient planes, u8 [2][H*W] finite flags
    DevBuf flags, sidx;        // uint32 per grid position
    DevBuf scan_tmp;
    DevBuf xs, ys, slabels;    // compacted sample list
    DevBuf n_dev;              // uint32 sample count
    DevBuf feats;              // float [cap][D]
    DevBuf leaf_ids;           // int [T][cap]
    DevBuf post;               // float [cap][sumC]
    DevBuf lowres;             // float, per layer [gh][gw][C_l]
    DevBuf posteriors;         // float [layer][H][W][C_l]
    std::vector<DevBuf> kept;  // posteriors of earlier key frames kept for the map worker (rss_posteriors_keep)
    std::vector<int> kept_npix; // pixels of each kept slot (0 = empty)
};

}  // namespace rss

struct rss_ctx {
    int device = 0, sm_count = 0;
    cudaStream_t s0 = nullptr, s1 = nullptr;
    cudaEvent_t ev[16] = {nullptr};
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_cloud = nullptr;
    rss::HostConfig cfg;
    rss::ForestDev forest;
    rss::FrameState fr;
    rss::DevBuf lab_gamma, lab_cbrt;  // u16 LUTs of cvtColor(BGR2Lab)
    rss::DevBuf tapx, tapy;           // ResizeTap[(P+1)][r]
    rss::DevBuf feat_xy;              // u16[r*r]: patch pixel k -> dx | dy << 8
    rss::PinBuf pin_in, pin_out, pin_small;
    rss::PinBuf pin_pose;             // PoseParams staging: rewritten before every frame, copied by a (capturable) H2D
    rss::DevBuf pose_dev;             // PoseParams
    struct KeyframeGraph* kf_graph = nullptr;  // captured device part of rss_segment_keyframe (crf.cu)
    bool graph_enabled = true;
    std::vector<uint16_t> service_depth;  // raw depth recovered from a service request's rectified cloud (api.cu)
    bool counted_live = false;        // this context is counted in live_contexts(device)
    bool capturing = false;           // the device part of a keyframe is being captured: no per-stage timing events
    void mark(int i) {                // per-stage timing event on s0 (eager runs only)
        if (!capturing) cudaEventRecord(ev[i], s0);
    }
    rss_timings tim = {0, 0, 0, 0, 0, 0, 0, 0};
    uint64_t launches = 0;
    std::set<const void*> smem_attr_done;  // kernels whose dynamic shared memory limit was raised on this device
    std::string err;
    rss_crf* keyframe_crf = nullptr;  // cached CRF of rss_segment_keyframe
    // optional per-kernel timing (rss_profile_enable)
    bool profile = false;
    struct Pending { const char* name; cudaEvent_t a, b; };
    std::vector<Pending> prof_pending;
    std::vector<cudaEvent_t> prof_pool;
    std::map<std::string, std::pair<double, uint64_t>> prof_acc;
    cudaEvent_t prof_event() {
        if (!prof_pool.empty()) { cudaEvent_t e = prof_pool.back(); prof_pool.pop_back(); return e; }
        cudaEvent_t e = nullptr;
        cudaEventCreate(&e);
        return e;
    }
    void prof_collect() {  // call after the streams have been synchronised
        for (Pending& p : prof_pending) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) {
                auto& acc = prof_acc[p.name];
                acc.first += ms;
                acc.second += 1;
            }
            prof_pool.push_back(p.a);
            prof_pool.push_back(p.b);
        }
        prof_pending.clear();
    }
    rss_status fail(rss_status s, const std::string& m) {
        err = m;
        return s;
    }
};

#define RSS_CU(ctx, call)                                                                             \
    do {                                                                                              \
        cudaError_t e__ = (call);                                                                     \
        if (e__ != cudaSuccess)                                                                       \
            return (ctx)->fail(RSS_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));    \
    } while (0)

// every kernel launch of the library goes through this so that rss_kernel_launches() is a real count
#define RSS_LAUNCH(ctx, kernel, grid, block, smem, stream, ...)                  \
    do {                                                                         \
        cudaEvent_t pa__ = nullptr, pb__ = nullptr;                              \
        if ((ctx)->profile) {                                                    \
            pa__ = (ctx)->prof_event();                                          \
            pb__ = (ctx)->prof_event();                                          \
            cudaEventRecord(pa__, (stream));                                     \
        }                                                                        \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);              \
        (ctx)->launches++;                                                       \
        if ((ctx)->profile) {                                                    \
            cudaEventRecord(pb__, (stream));                                     \
            (ctx)->prof_pending.push_back(rss_ctx::Pending{#kernel, pa__, pb__}); \
        }                                                                        \
    } while (0)

// the same for kernels whose name cannot be stringified (template instantiations with several arguments)
#define RSS_LAUNCH_NAMED(ctx, name, kernel, grid, block, smem, stream, ...)     \
    do {                                                                         \
        cudaEvent_t pa__ = nullptr, pb__ = nullptr;                              \
        if ((ctx)->profile) {                                                    \
            pa__ = (ctx)->prof_event();                                          \
            pb__ = (ctx)->prof_event();                                          \
            cudaEventRecord(pa__, (stream));                                     \
        }                                                                        \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);              \
        (ctx)->launches++;                                                       \
        if ((ctx)->profile) {                                                    \
            cudaEventRecord(pb__, (stream));                                     \
            (ctx)->prof_pending.push_back(rss_ctx::Pending{name, pa__, pb__});   \
        }                                                                        \
    } while (0)

static inline int rss_div_up(long long a, long long b) { return (int)((a + b - 1) / b); }
