// C ABI of librss.so, part 1: context (config + libforest model), feature extraction, forest prediction and
// the frame worker body.  See include/rss.h for the reference interface each entry point replaces.
#include <cmath>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>

#include "json_min.hpp"
#include "kernels.hpp"
#include "scan.cuh"

using namespace rss;

namespace rss {
void crf_release_cached(rss_ctx* ctx);  // crf.cu
void keyframe_graph_release(rss_ctx* ctx);  // crf.cu
rss_status frame_segment_resident(rss_ctx* ctx, const float* Kinv, const float* R, const float* t, float fill);
rss_status frame_segment_begin(rss_ctx* ctx, const float* Kinv, const float* R, const float* t);
rss_status frame_segment_finish(rss_ctx* ctx, float fill, float* unary_out = nullptr, int unary_stride = 0);
}

static thread_local std::string g_create_error;

extern "C" const char* rss_status_string(rss_status s) {
    switch (s) {
        case RSS_OK: return "ok";
        case RSS_ERR_INVALID: return "invalid argument";
        case RSS_ERR_IO: return "i/o error";
        case RSS_ERR_CONFIG: return "config key not found";
        case RSS_ERR_MODEL: return "malformed forest model";
        case RSS_ERR_CUDA: return "CUDA error";
        case RSS_ERR_CAPACITY: return "lattice hash table overflow";
        case RSS_ERR_STATE: return "call order violated";
    }
    return "unknown status";
}
extern "C" const char* rss_last_error(const rss_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

// ------------------------------------------------------------------------------------------------
// config (reference: Utils::Config, src/config.cpp; keys of SURVEY 8b)
// ------------------------------------------------------------------------------------------------
static rss_status load_config(const char* path, HostConfig& cfg, std::string& err) {
    std::ifstream f(path);
    if (!f) { err = std::string("cannot open config ") + path; return RSS_ERR_IO; }
    std::stringstream ss;
    ss << f.rdbuf();
    JsonValue root;
    try {
        root = JsonParser::parse(ss.str());
    } catch (const std::exception& e) {
        err = std::string("config parse error: ") + e.what();
        return RSS_ERR_IO;
    }
    auto need = [&](const char* k) -> const JsonValue* {
        const JsonValue* v = root.find(k);
        if (!v) err = std::string("The key: '") + k + "' was not found in the config file.";  // config.h:13-24
        return v;
    };
    const JsonValue* v;
    // Features::FeatureExtractor(Config), feature_extractor.h:29-39 - all six keys are mandatory there
    if (!(v = need("feature_color_patch"))) return RSS_ERR_CONFIG; cfg.use_color = v->asBool();
    if (!(v = need("feature_depth"))) return RSS_ERR_CONFIG; cfg.use_depth = v->asBool();
    if (!(v = need("feature_height"))) return RSS_ERR_CONFIG; cfg.use_height = v->asBool();
    if (!(v = need("feature_normal"))) return RSS_ERR_CONFIG; cfg.use_normal = v->asBool();
    if (!(v = need("patch_size"))) return RSS_ERR_CONFIG; cfg.patch_size = v->asInt();
    if (!(v = need("patch_size_reduce"))) return RSS_ERR_CONFIG; cfg.patch_size_reduce = v->asInt();
    if (cfg.patch_size < 1 || cfg.patch_size > 16000 || cfg.patch_size_reduce < 1 || cfg.patch_size_reduce > 32) {
        err = "patch_size / patch_size_reduce out of range";
        return RSS_ERR_INVALID;
    }
    // Segmenter ctor, segmenter.cpp:73-98: layers, class counts (labels >= 0), "Unknown" default label
    if ((v = root.find("color_codings")) && v->kind == JsonValue::Array) {
        for (const JsonValue& layer : v->arr) {
            if (cfg.layer_count >= RSS_MAX_LAYERS) { err = "too many label layers"; return RSS_ERR_INVALID; }
            const JsonValue* coding = layer.find("coding");
            int names = 0, unknown = -1;
            if (coding && coding->kind == JsonValue::Array)
                for (const JsonValue& cls : coding->arr) {
                    const JsonValue* lab = cls.find("label");
                    const JsonValue* nm = cls.find("name");
                    if (lab && lab->asInt() >= 0) names++;
                    if (nm && nm->str == "Unknown" && unknown < 0) unknown = names - 1;
                }
            cfg.class_counts[cfg.layer_count] = names;
            cfg.unknown_label[cfg.layer_count] = unknown < 0 ? 0 : unknown;
            cfg.layer_count++;
        }
    }
    // segmenter.cpp:120-127.  Optional here (a FeatureExtractor-only config, as used by train.cpp, lacks them);
    // defaults are the values of the reference's resources/config.json.
    if ((v = root.find("use_dense_crf"))) cfg.use_dense_crf = v->asBool();
    if ((v = root.find("dcrf_xyz_kernel"))) cfg.dcrf_xyz = (float)v->asDouble();
    if ((v = root.find("dcrf_rgb_kernel"))) cfg.dcrf_rgb = (float)v->asDouble();
    if ((v = root.find("dcrf_kernel_weight"))) cfg.dcrf_w = (float)v->asDouble();
    if ((v = root.find("dcrf_iterations"))) cfg.dcrf_iters = (int)v->asDouble();
    if ((v = root.find("rf_prediction_stride"))) cfg.rf_stride = (int)v->asDouble();
    if ((v = root.find("depth_min"))) cfg.depth_min = (float)v->asDouble();
    if ((v = root.find("depth_max"))) cfg.depth_max = (float)v->asDouble();
    if (cfg.rf_stride < 1) { err = "rf_prediction_stride must be >= 1"; return RSS_ERR_INVALID; }
    return RSS_OK;
}

// ------------------------------------------------------------------------------------------------
// libforest binary model (RandomForest::read classifier.cpp:222-235, DecisionTree::read :134-142,
// readBinary io.h:43-108): little-endian, no header.  int32 T; per tree five length-prefixed vectors:
// splitFeatures, thresholds, leftChild, histograms (n x vector<float>), multi_histograms (n x vector<vector<float>>).
// ------------------------------------------------------------------------------------------------
namespace {
struct Reader {
    const unsigned char* p;
    const unsigned char* e;
    bool ok = true;
    int32_t i32() {
        if (e - p < 4) { ok = false; return 0; }
        int32_t v;
        memcpy(&v, p, 4);
        p += 4;
        return v;
    }
    bool raw(void* dst, size_t bytes) {
        if ((size_t)(e - p) < bytes) { ok = false; return false; }
        memcpy(dst, p, bytes);
        p += bytes;
        return true;
    }
};
}  // namespace

static rss_status load_forest_bytes(rss_ctx* ctx, const unsigned char* data, size_t size);
static rss_status load_forest(rss_ctx* ctx, const char* path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) return ctx->fail(RSS_ERR_IO, std::string("cannot open forest ") + path);
    std::vector<unsigned char> bytes((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    return load_forest_bytes(ctx, bytes.data(), bytes.size());
}
static rss_status load_forest_bytes_impl(rss_ctx* ctx, const unsigned char* data, size_t size);
// A truncated or corrupt model must come back as RSS_ERR_MODEL, never as an exception through the extern "C" boundary.
static rss_status load_forest_bytes(rss_ctx* ctx, const unsigned char* data, size_t size) {
    try {
        return load_forest_bytes_impl(ctx, data, size);
    } catch (const std::exception& e) {
        ctx->forest.loaded = false;
        return ctx->fail(RSS_ERR_MODEL, std::string("forest: ") + e.what());
    }
}
static rss_status load_forest_bytes_impl(rss_ctx* ctx, const unsigned char* data, size_t size) {
    Reader rd{data, data + size};
    ForestDev& F = ctx->forest;
    const int T = rd.i32();
    if (!rd.ok || T <= 0 || T > 4096) return ctx->fail(RSS_ERR_MODEL, "forest: bad tree count");
    std::vector<Node> nodes;
    std::vector<float> leaves;
    std::vector<int> tree_off(T + 1, 0);
    size_t true_nodes = 0;  // without the alignment padding
    int L = -1, sumC = 0;
    int C[RSS_MAX_LAYERS] = {0};
    for (int t = 0; t < T; t++) {
        const int n = rd.i32();
        // every node costs at least 12 bytes of split data that must still be in the file: bound n BEFORE allocating
        if (!rd.ok || n <= 0 || (size_t)n > (size_t)(rd.e - rd.p) / 12) return ctx->fail(RSS_ERR_MODEL, "forest: bad node count");
        const size_t nb = (size_t)n * 4;
        std::vector<int32_t> feat(n), left(n);
        std::vector<float> thr(n);
        if (!rd.raw(feat.data(), nb)) return ctx->fail(RSS_ERR_MODEL, "forest: truncated splitFeatures");
        if (rd.i32() != n || !rd.raw(thr.data(), nb)) return ctx->fail(RSS_ERR_MODEL, "forest: truncated thresholds");
        if (rd.i32() != n || !rd.raw(left.data(), nb)) return ctx->fail(RSS_ERR_MODEL, "forest: truncated leftChild");
        // roots at ODD node indices: a libforest tree's sibling pairs start at odd local indices (root 0, children 1|2, 3|4,
        // ...), so they then start at even global ones and the two 16-byte nodes share one 32-byte sector (features.cu)
        if (nodes.size() % 2 == 0) nodes.push_back(Node{0, 0.f, 0, -1});
        const size_t base = nodes.size();
        true_nodes += (size_t)n;
        tree_off[t] = (int)base;
        nodes.resize(base + n);
        for (int i = 0; i < n; i++) {
            if (left[i] < 0 || left[i] + 1 >= n + (left[i] == 0 ? 2 : 0) || (left[i] != 0 && left[i] <= i))
                return ctx->fail(RSS_ERR_MODEL, "forest: child index out of range");
            nodes[base + i] = Node{feat[i], thr[i], left[i], -1};
        }
        // plain histograms (single-label forests): a leaf with c > 0 floats
        if (rd.i32() != n) return ctx->fail(RSS_ERR_MODEL, "forest: histograms length mismatch");
        std::vector<std::vector<float>> single(n);
        for (int i = 0; i < n; i++) {
            const int c = rd.i32();
            if (!rd.ok || c < 0 || c > 4096) return ctx->fail(RSS_ERR_MODEL, "forest: bad histogram size");
            single[i].resize(c);
            if (c && !rd.raw(single[i].data(), (size_t)4 * c)) return ctx->fail(RSS_ERR_MODEL, "forest: truncated histogram");
        }
        if (rd.i32() != n) return ctx->fail(RSS_ERR_MODEL, "forest: multi_histograms length mismatch");
        for (int i = 0; i < n; i++) {
            const int l = rd.i32();
            if (!rd.ok || l < 0 || l > RSS_MAX_LAYERS) return ctx->fail(RSS_ERR_MODEL, "forest: bad layer count");
            std::vector<float> row;
            int lc[RSS_MAX_LAYERS] = {0};
            for (int k = 0; k < l; k++) {
                const int c = rd.i32();
                if (!rd.ok || c < 0 || c > 4096) return ctx->fail(RSS_ERR_MODEL, "forest: bad histogram size");
                lc[k] = c;
                const size_t o = row.size();
                row.resize(o + c);
                if (c && !rd.raw(row.data() + o, (size_t)4 * c)) return ctx->fail(RSS_ERR_MODEL, "forest: truncated multi histogram");
            }
            int ll = l;
            if (l == 0 && !single[i].empty()) {  // single-label forest: one layer
                row = single[i];
                ll = 1;
                lc[0] = (int)row.size();
            }
            if (nodes[base + i].left == 0) {
                if (ll == 0 || row.empty()) return ctx->fail(RSS_ERR_MODEL, "forest: leaf without histogram");
                if (L < 0) {
                    L = ll;
                    for (int k = 0; k < ll; k++) { C[k] = lc[k]; sumC += lc[k]; }
                    if (sumC <= 0) return ctx->fail(RSS_ERR_MODEL, "forest: leaf histograms without classes");
                } else {
                    if (ll != L) return ctx->fail(RSS_ERR_MODEL, "forest: inconsistent layer count");
                    for (int k = 0; k < ll; k++)
                        if (lc[k] != C[k]) return ctx->fail(RSS_ERR_MODEL, "forest: inconsistent class count");
                }
                nodes[base + i].leaf = (int)(leaves.size() / (size_t)sumC);
                leaves.insert(leaves.end(), row.begin(), row.end());
            }
        }
    }
    tree_off[T] = (int)nodes.size();
    if (L <= 0 || sumC <= 0) return ctx->fail(RSS_ERR_MODEL, "forest: no leaf histograms");
    const int D = ctx->cfg.feature_length();
    for (const Node& nd : nodes)
        if (nd.left != 0 && (nd.feat < 0 || nd.feat >= D))
            return ctx->fail(RSS_ERR_MODEL, "forest: split feature index exceeds the configured feature length");
    F.T = T; F.L = L; F.sumC = sumC;
    for (int k = 0; k < RSS_MAX_LAYERS; k++) F.C[k] = C[k];
    F.total_nodes = (int)true_nodes;
    F.total_leaves = (int)(leaves.size() / (size_t)sumC);
    F.tree_off = tree_off;
    RSS_CU(ctx, F.nodes.reserve(nodes.size() * sizeof(Node)));
    RSS_CU(ctx, F.leaves.reserve(leaves.size() * sizeof(float)));
    RSS_CU(ctx, F.tree_off_dev.reserve(tree_off.size() * sizeof(int)));
    RSS_CU(ctx, cudaMemcpy(F.nodes.ptr, nodes.data(), nodes.size() * sizeof(Node), cudaMemcpyHostToDevice));
    RSS_CU(ctx, cudaMemcpy(F.leaves.ptr, leaves.data(), leaves.size() * sizeof(float), cudaMemcpyHostToDevice));
    RSS_CU(ctx, cudaMemcpy(F.tree_off_dev.ptr, tree_off.data(), tree_off.size() * sizeof(int), cudaMemcpyHostToDevice));
    F.loaded = true;
    return RSS_OK;
}

// ------------------------------------------------------------------------------------------------
// look-up tables: cvtColor(BGR2Lab, 8U) and the cv::resize(INTER_LINEAR, 8U) tap table
// ------------------------------------------------------------------------------------------------
static rss_status upload_tables(rss_ctx* ctx) {
    // OpenCV RGB2Lab_b: sRGBGammaTab_b (<<3) and LabCbrtTab_b (<<15).  OpenCV builds the cube-root table in
    // 32-bit softfloat; relative to a double evaluation three entries round the other way.
    uint16_t gamma[256], cb[3072];
    for (int i = 0; i < 256; i++) {
        const double x = i / 255.0;
        gamma[i] = (uint16_t)lrint(255.0 * 8.0 * (x <= 0.04045 ? x / 12.92 : pow((x + 0.055) / 1.055, 2.4)));
    }
    for (int i = 0; i < 3072; i++) {
        const double x = i / (255.0 * 8.0);
        cb[i] = (uint16_t)lrint(32768.0 * (x < 216.0 / 24389.0 ? x * (841.0 / 108.0) + 16.0 / 116.0 : cbrt(x)));
    }
    cb[49] = 9454; cb[324] = 17745; cb[628] = 22126;
    RSS_CU(ctx, ctx->lab_gamma.reserve(sizeof(gamma)));
    RSS_CU(ctx, ctx->lab_cbrt.reserve(sizeof(cb)));
    RSS_CU(ctx, cudaMemcpy(ctx->lab_gamma.ptr, gamma, sizeof(gamma), cudaMemcpyHostToDevice));
    RSS_CU(ctx, cudaMemcpy(ctx->lab_cbrt.ptr, cb, sizeof(cb), cudaMemcpyHostToDevice));
    // tap table: for every ROI half-size h (ROI side S = 2h+1) and destination index d in [0, r)
    const int P = ctx->cfg.patch_size, r = ctx->cfg.patch_size_reduce;
    std::vector<ResizeTap> tx((size_t)(P + 1) * r), ty((size_t)(P + 1) * r);
    for (int h = 0; h <= P; h++) {
        const int S = 2 * h + 1;
        const double scale = 1.0 / ((double)r / (double)S);
        for (int d = 0; d < r; d++) {
            float f = (float)((d + 0.5) * scale - 0.5);
            int s = (int)floorf(f);
            f -= (float)s;
            // horizontal: a clamped index zeroes the fraction
            float fx = f;
            int sx = s;
            if (sx < 0) { sx = 0; fx = 0.f; }
            if (sx >= S - 1) { sx = S - 1; fx = 0.f; }
            ResizeTap a;
            a.i0 = (short)sx;
            a.i1 = (short)(sx + 1 > S - 1 ? S - 1 : sx + 1);
            a.w0 = (short)lrintf((1.f - fx) * 2048.f);
            a.w1 = (short)lrintf(fx * 2048.f);
            tx[(size_t)h * r + d] = a;
            // vertical: the two rows are clipped, the fraction is kept
            ResizeTap b;
            b.i0 = (short)(s < 0 ? 0 : (s > S - 1 ? S - 1 : s));
            b.i1 = (short)(s + 1 < 0 ? 0 : (s + 1 > S - 1 ? S - 1 : s + 1));
            b.w0 = (short)lrintf((1.f - f) * 2048.f);
            b.w1 = (short)lrintf(f * 2048.f);
            ty[(size_t)h * r + d] = b;
        }
    }
    // patch pixel k -> (dx, dy) for the on-demand feature evaluation of the frame path
    std::vector<uint16_t> fxy((size_t)r * r);
    for (int k = 0; k < r * r; k++) fxy[k] = (uint16_t)((k % r) | ((k / r) << 8));
    RSS_CU(ctx, ctx->feat_xy.reserve(fxy.size() * sizeof(uint16_t)));
    RSS_CU(ctx, cudaMemcpy(ctx->feat_xy.ptr, fxy.data(), fxy.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
    RSS_CU(ctx, ctx->tapx.reserve(tx.size() * sizeof(ResizeTap)));
    RSS_CU(ctx, ctx->tapy.reserve(ty.size() * sizeof(ResizeTap)));
    RSS_CU(ctx, cudaMemcpy(ctx->tapx.ptr, tx.data(), tx.size() * sizeof(ResizeTap), cudaMemcpyHostToDevice));
    RSS_CU(ctx, cudaMemcpy(ctx->tapy.ptr, ty.data(), ty.size() * sizeof(ResizeTap), cudaMemcpyHostToDevice));
    return RSS_OK;
}

static void free_ctx(rss_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->counted_live) live_contexts(ctx->device)--;
    keyframe_graph_release(ctx);
    crf_release_cached(ctx);
    FrameState& f = ctx->fr;
    DevBuf* bufs[] = {&f.rgb, &f.depth, &f.labels, &f.lab, &f.xyz, &f.dist_a, &f.dist_b, &f.integ, &f.integ_cnt,
                      &f.normals, &f.grad, &f.fin, &f.flags, &f.sidx, &f.scan_tmp, &f.xs, &f.ys, &f.slabels,
                      &f.n_dev, &f.feats, &f.leaf_ids, &f.post, &f.lowres, &f.posteriors, &ctx->forest.nodes,
                      &ctx->forest.tree_off_dev, &ctx->forest.leaves, &ctx->lab_gamma, &ctx->lab_cbrt, &ctx->tapx,
                      &ctx->tapy, &ctx->feat_xy, &ctx->pose_dev};
    for (DevBuf* b : bufs) b->release();
    for (DevBuf& b : f.kept) b.release();
    ctx->pin_in.release();
    ctx->pin_out.release();
    ctx->pin_small.release();
    ctx->pin_pose.release();
    for (cudaEvent_t& e : ctx->ev)
        if (e) cudaEventDestroy(e);
    ctx->prof_collect();
    for (cudaEvent_t e : ctx->prof_pool) cudaEventDestroy(e);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->ev_cloud) cudaEventDestroy(ctx->ev_cloud);
    if (ctx->s0) cudaStreamDestroy(ctx->s0);
    if (ctx->s1) cudaStreamDestroy(ctx->s1);
    delete ctx;
}

extern "C" rss_status rss_create(const char* config_json_path, const char* forest_dat_path, int cuda_device,
                                 rss_ctx** out) {
    g_create_error.clear();
    if (!out || !config_json_path) { g_create_error = "null argument"; return RSS_ERR_INVALID; }
    *out = nullptr;
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0) {
        g_create_error = std::string("no CUDA device: ") + (ce != cudaSuccess ? cudaGetErrorString(ce) : "device count is 0") +
                         " (librss has no CPU fallback)";
        return RSS_ERR_CUDA;
    }
    if (cuda_device < 0 || cuda_device >= ndev) { g_create_error = "cuda_device out of range"; return RSS_ERR_INVALID; }
    rss_ctx* ctx = new rss_ctx();
    ctx->device = cuda_device;
    rss_status st = load_config(config_json_path, ctx->cfg, ctx->err);
    auto bail = [&](rss_status s) {
        g_create_error = ctx->err;
        free_ctx(ctx);
        return s;
    };
    if (st != RSS_OK) return bail(st);
    if (cudaSetDevice(cuda_device) != cudaSuccess) { ctx->err = "cudaSetDevice failed"; return bail(RSS_ERR_CUDA); }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, cuda_device) != cudaSuccess) { ctx->err = "cudaGetDeviceProperties failed"; return bail(RSS_ERR_CUDA); }
    ctx->sm_count = prop.multiProcessorCount;
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    // s1 carries the cloud -> normals chain (the longest dependency chain of the feature stage): high priority
    if (cudaStreamCreateWithFlags(&ctx->s0, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithPriority(&ctx->s1, cudaStreamNonBlocking, prio_hi) != cudaSuccess) {
        ctx->err = "cudaStreamCreate failed";
        return bail(RSS_ERR_CUDA);
    }
    for (cudaEvent_t& e : ctx->ev)
        if (cudaEventCreate(&e) != cudaSuccess) { ctx->err = "cudaEventCreate failed"; return bail(RSS_ERR_CUDA); }
    if (cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_cloud, cudaEventDisableTiming) != cudaSuccess) {
        ctx->err = "cudaEventCreate failed";
        return bail(RSS_ERR_CUDA);
    }
    st = upload_tables(ctx);
    if (st != RSS_OK) return bail(st);
    if (forest_dat_path) {
        st = load_forest(ctx, forest_dat_path);
        if (st != RSS_OK) return bail(st);
    }
    live_contexts(cuda_device)++;
    ctx->counted_live = true;
    *out = ctx;
    return RSS_OK;
}

extern "C" rss_status rss_load_forest(rss_ctx* ctx, const char* forest_dat_path) {
    if (!ctx || !forest_dat_path) return RSS_ERR_INVALID;
    RSS_CU(ctx, cudaSetDevice(ctx->device));
    RSS_CU(ctx, cudaStreamSynchronize(ctx->s0));
    ctx->forest.loaded = false;
    return load_forest(ctx, forest_dat_path);
}
extern "C" rss_status rss_load_forest_memory(rss_ctx* ctx, const void* bytes, size_t size) {
    if (!ctx || !bytes) return RSS_ERR_INVALID;
    RSS_CU(ctx, cudaSetDevice(ctx->device));
    RSS_CU(ctx, cudaStreamSynchronize(ctx->s0));
    ctx->forest.loaded = false;
    return load_forest_bytes(ctx, static_cast<const unsigned char*>(bytes), size);
}

extern "C" void* rss_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) return nullptr;
    return p;
}
extern "C" void rss_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

extern "C" rss_status rss_destroy(rss_ctx* ctx) {
    if (!ctx) return RSS_ERR_INVALID;
    free_ctx(ctx);
    return RSS_OK;
}

static void info_from_config(const HostConfig& c, rss_info* o) {
    o->feature_color_patch = c.use_color; o->feature_depth = c.use_depth;
    o->feature_height = c.use_height; o->feature_normal = c.use_normal;
    o->patch_size = c.patch_size; o->patch_size_reduce = c.patch_size_reduce; o->feature_length = c.feature_length();
    o->layer_count = c.layer_count;
    for (int l = 0; l < c.layer_count; l++) { o->class_counts[l] = c.class_counts[l]; o->total_classes += c.class_counts[l]; }
    for (int l = 0; l < RSS_MAX_LAYERS; l++) o->unknown_label[l] = l < c.layer_count ? c.unknown_label[l] : 0;
    o->use_dense_crf = c.use_dense_crf; o->dcrf_iterations = c.dcrf_iters; o->rf_prediction_stride = c.rf_stride;
    o->dcrf_xyz_kernel = c.dcrf_xyz; o->dcrf_rgb_kernel = c.dcrf_rgb; o->dcrf_kernel_weight = c.dcrf_w;
    o->depth_min = c.depth_min; o->depth_max = c.depth_max;
    o->cuda_device = -1;
}
extern "C" rss_status rss_parse_config(const char* config_json_path, rss_info* out) {
    g_create_error.clear();
    if (!config_json_path || !out) { g_create_error = "null argument"; return RSS_ERR_INVALID; }
    memset(out, 0, sizeof(*out));
    HostConfig cfg;
    std::string err;
    rss_status st;
    try {
        st = load_config(config_json_path, cfg, err);
    } catch (const std::exception& e) {
        st = RSS_ERR_IO;
        err = std::string("config: ") + e.what();
    }
    if (st != RSS_OK) { g_create_error = err; return st; }
    info_from_config(cfg, out);
    return RSS_OK;
}

extern "C" rss_status rss_get_info(const rss_ctx* ctx, rss_info* o) {
    if (!ctx || !o) return RSS_ERR_INVALID;
    memset(o, 0, sizeof(*o));
    const HostConfig& c = ctx->cfg;
    o->feature_color_patch = c.use_color; o->feature_depth = c.use_depth;
    o->feature_height = c.use_height; o->feature_normal = c.use_normal;
    o->patch_size = c.patch_size; o->patch_size_reduce = c.patch_size_reduce; o->feature_length = c.feature_length();
    const ForestDev& F = ctx->forest;
    o->num_trees = F.T; o->total_nodes = F.total_nodes; o->total_leaves = F.total_leaves;
    if (F.loaded) {
        o->layer_count = F.L; o->total_classes = F.sumC;
        for (int l = 0; l < F.L; l++) o->class_counts[l] = F.C[l];
    } else {
        o->layer_count = c.layer_count;
        for (int l = 0; l < c.layer_count; l++) { o->class_counts[l] = c.class_counts[l]; o->total_classes += c.class_counts[l]; }
    }
    for (int l = 0; l < RSS_MAX_LAYERS; l++) o->unknown_label[l] = l < c.layer_count ? c.unknown_label[l] : 0;
    o->use_dense_crf = c.use_dense_crf; o->dcrf_iterations = c.dcrf_iters; o->rf_prediction_stride = c.rf_stride;
    o->dcrf_xyz_kernel = c.dcrf_xyz; o->dcrf_rgb_kernel = c.dcrf_rgb; o->dcrf_kernel_weight = c.dcrf_w;
    o->depth_min = c.depth_min; o->depth_max = c.depth_max;
    o->cuda_device = ctx->device; o->sm_count = ctx->sm_count;
    return RSS_OK;
}
extern "C" rss_status rss_get_timings(const rss_ctx* ctx, rss_timings* out) {
    if (!ctx || !out) return RSS_ERR_INVALID;
    *out = ctx->tim;
    return RSS_OK;
}
extern "C" uint64_t rss_kernel_launches(const rss_ctx* ctx) { return ctx ? ctx->launches : 0; }

extern "C" rss_status rss_profile_enable(rss_ctx* ctx, int enable) {
    if (!ctx) return RSS_ERR_INVALID;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    ctx->prof_collect();
    ctx->prof_acc.clear();
    ctx->profile = enable != 0;
    return RSS_OK;
}
extern "C" rss_status rss_profile_get(rss_ctx* ctx, int index, char* name, int name_cap, double* total_ms,
                                      uint64_t* launches) {
    if (!ctx) return RSS_ERR_INVALID;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    ctx->prof_collect();
    if (index < 0 || index >= (int)ctx->prof_acc.size()) return RSS_ERR_INVALID;
    auto it = ctx->prof_acc.begin();
    std::advance(it, index);
    if (name && name_cap > 0) {
        strncpy(name, it->first.c_str(), name_cap - 1);
        name[name_cap - 1] = 0;
    }
    if (total_ms) *total_ms = it->second.first;
    if (launches) *launches = it->second.second;
    return RSS_OK;
}

// ------------------------------------------------------------------------------------------------
// frame pipeline
// ------------------------------------------------------------------------------------------------
namespace rss {

rss_status frame_upload(rss_ctx* ctx, const uint8_t* rgb, const uint16_t* depth, int W, int H) {
    FrameState& f = ctx->fr;
    if (W < 4 || H < 4 || W > 16384 || H > 16384) return ctx->fail(RSS_ERR_INVALID, "image size out of range");
    const size_t NP = (size_t)W * H;
    RSS_CU(ctx, f.rgb.reserve(NP * 3));
    RSS_CU(ctx, f.depth.reserve(NP * 2));
    if (rgb || depth) {
        if (!rgb || !depth) return ctx->fail(RSS_ERR_INVALID, "rgb and depth must both be given (or both NULL to reuse the resident frame)");
        RSS_CU(ctx, cudaMemcpyAsync(f.rgb.ptr, rgb, NP * 3, cudaMemcpyHostToDevice, ctx->s0));
        RSS_CU(ctx, cudaMemcpyAsync(f.depth.ptr, depth, NP * 2, cudaMemcpyHostToDevice, ctx->s0));
        f.W = W; f.H = H;
    } else if (f.W != W || f.H != H) {
        return ctx->fail(RSS_ERR_STATE, "no resident frame of this size");
    }
    f.have_feats = f.have_post = f.have_cloud = f.have_lab = f.have_integral = false;
    f.n_samples = -1;
    return RSS_OK;
}

// Lab+border on s0; cloud and the normals preparation on s1, joined back into s0.
// host part of the pose: M = R * Kinv, row by column, (a0*b0 + a1*b1) + a2*b2 in float (Eigen 3x3 product,
// feature_extractor.h:223), into the pinned staging block the device copy is refreshed from
rss_status frame_set_pose(rss_ctx* ctx, const float* Kinv, const float* R, const float* t) {
    RSS_CU(ctx, ctx->pin_pose.reserve(sizeof(PoseParams)));
    RSS_CU(ctx, ctx->pose_dev.reserve(sizeof(PoseParams)));
    PoseParams* p = ctx->pin_pose.as<PoseParams>();
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            const float p0 = R[3 * i] * Kinv[j], p1 = R[3 * i + 1] * Kinv[3 + j], p2 = R[3 * i + 2] * Kinv[6 + j];
            const float s = p0 + p1;
            p->M[3 * i + j] = s + p2;
        }
    for (int i = 0; i < 3; i++) p->t[i] = t[i];
    return RSS_OK;
}
rss_status frame_prepare(rss_ctx* ctx, float dmin, float dmax) {
    FrameState& f = ctx->fr;
    const HostConfig& cfg = ctx->cfg;
    const int W = f.W, H = f.H, P = cfg.patch_size;
    const size_t NP = (size_t)W * H;
    const bool need_cloud = cfg.use_height || cfg.use_normal;
    if (cfg.use_normal && H > 1024) return ctx->fail(RSS_ERR_INVALID, "feature_normal supports image heights up to 1024");
    if (need_cloud) {
        RSS_CU(ctx, f.xyz.reserve(NP * sizeof(float4)));
        RSS_CU(ctx, cudaEventRecord(ctx->ev_fork, ctx->s0));
        RSS_CU(ctx, cudaStreamWaitEvent(ctx->s1, ctx->ev_fork, 0));
        RSS_CU(ctx, cudaMemcpyAsync(ctx->pose_dev.ptr, ctx->pin_pose.ptr, sizeof(PoseParams), cudaMemcpyHostToDevice, ctx->s1));
        launch_cloud(ctx, ctx->s1, f.depth.as<uint16_t>(), W, H, ctx->pose_dev.as<PoseParams>(), dmin, dmax, f.xyz.as<float4>());
        RSS_CU(ctx, cudaEventRecord(ctx->ev_cloud, ctx->s1));
        f.have_cloud = true;
        if (cfg.use_normal) {
            RSS_CU(ctx, f.dist_a.reserve(NP * 4));
            RSS_CU(ctx, f.dist_b.reserve(NP * 4));
            RSS_CU(ctx, f.grad.reserve(integral_elems(W, H) * 6 * 4));
            RSS_CU(ctx, f.fin.reserve(integral_elems(W, H) * 2));
            RSS_CU(ctx, f.integ.reserve(integral_elems(W, H) * 6 * sizeof(double)));
            RSS_CU(ctx, f.integ_cnt.reserve(integral_elems(W, H) * 2 * sizeof(int)));
            launch_normals_prepare(ctx, ctx->s1, f.xyz.as<float4>(), W, H, f.dist_a.as<float>(), f.dist_b.as<float>(),
                                   f.integ.as<double>(), f.integ_cnt.as<int>(), f.grad.as<float>(), f.fin.as<uint8_t>());
            f.have_integral = true;
        }
        RSS_CU(ctx, cudaEventRecord(ctx->ev_join, ctx->s1));
    }
    if (cfg.use_color) {
        RSS_CU(ctx, f.lab.reserve((size_t)(W + 2 * P) * (H + 2 * P) * sizeof(uchar4)));
        launch_lab_border(ctx, ctx->s0, f.rgb.as<uint8_t>(), W, H, P, f.lab.as<uchar4>());
        f.have_lab = true;
    }
    if (need_cloud) RSS_CU(ctx, cudaStreamWaitEvent(ctx->s0, ctx->ev_join, 0));
    RSS_CU(ctx, cudaGetLastError());
    return RSS_OK;
}

// sample selection + compaction + materialised features for the compacted list
// materialize = false (frame path): only the sample list; the forest evaluates features on demand
rss_status frame_extract(rss_ctx* ctx, int stride, float dmin, float dmax, int extract_type, const int8_t* labels_dev,
                         int n_label_layers, bool materialize = true) {
    FrameState& f = ctx->fr;
    const HostConfig& cfg = ctx->cfg;
    const int W = f.W, H = f.H, D = cfg.feature_length();
    const int gw = rss_div_up(W, stride), gh = rss_div_up(H, stride);
    const size_t cap = (size_t)gw * gh;
    f.stride = stride; f.gw = gw; f.gh = gh;
    RSS_CU(ctx, f.flags.reserve(cap * 4));
    RSS_CU(ctx, f.sidx.reserve(cap * 4));
    RSS_CU(ctx, f.scan_tmp.reserve(scan_tmp_elems(cap) * 4));
    RSS_CU(ctx, f.n_dev.reserve(16));
    RSS_CU(ctx, f.xs.reserve(cap * 4));
    RSS_CU(ctx, f.ys.reserve(cap * 4));
    if (materialize) RSS_CU(ctx, f.feats.reserve(cap * D * sizeof(float)));
    RSS_CU(ctx, ctx->pin_small.reserve(64));
    if (n_label_layers > 0) RSS_CU(ctx, f.slabels.reserve(cap * n_label_layers * 4));
    const float dmin_mm = (float)(dmin * 1000.0), dmax_mm = (float)(dmax * 1000.0);  // feature_extractor.h:43-44
    launch_select(ctx, ctx->s0, f.depth.as<uint16_t>(), labels_dev, n_label_layers, extract_type, W, H, stride, dmin_mm,
                  dmax_mm, f.flags.as<uint32_t>());
    exclusive_scan_u32(f.flags.as<uint32_t>(), f.sidx.as<uint32_t>(), cap, f.scan_tmp.as<uint32_t>(),
                       f.n_dev.as<uint32_t>(), ctx->s0, &ctx->launches);
    launch_compact(ctx, ctx->s0, f.flags.as<uint32_t>(), f.sidx.as<uint32_t>(), W, H, stride, labels_dev, n_label_layers,
                   f.xs.as<int>(), f.ys.as<int>(), n_label_layers > 0 ? f.slabels.as<int>() : nullptr);
    RSS_CU(ctx, cudaMemcpyAsync(ctx->pin_small.ptr, f.n_dev.ptr, 4, cudaMemcpyDeviceToHost, ctx->s0));
    RSS_CU(ctx, cudaStreamSynchronize(ctx->s0));
    const int n = (int)*ctx->pin_small.as<uint32_t>();
    f.n_samples = n;
    f.have_feats = false;
    if (!materialize) return RSS_OK;
    int pos = 0, pos_depth = -1, pos_height = -1, pos_normal = -1;
    if (cfg.use_color) {
        launch_patch_features(ctx, ctx->s0, f.lab.as<uchar4>(), f.depth.as<uint16_t>(), W, H, cfg.patch_size,
                              cfg.patch_size_reduce, ctx->tapx.as<ResizeTap>(), ctx->tapy.as<ResizeTap>(), f.xs.as<int>(),
                              f.ys.as<int>(), n, f.feats.as<float>(), D);
        pos += 3 * cfg.patch_size_reduce * cfg.patch_size_reduce;
    }
    if (cfg.use_depth) pos_depth = pos++;
    if (cfg.use_height) pos_height = pos++;
    if (cfg.use_normal) pos_normal = pos++;
    if (pos_depth >= 0 || pos_height >= 0 || pos_normal >= 0)
        launch_scalar_features(ctx, ctx->s0, f.depth.as<uint16_t>(), f.xyz.as<float4>(), f.dist_b.as<float>(),
                               f.integ.as<double>(), f.integ_cnt.as<int>(), W, H, f.xs.as<int>(), f.ys.as<int>(), n,
                               f.feats.as<float>(), D, pos_depth, pos_height, pos_normal);
    RSS_CU(ctx, cudaGetLastError());
    f.have_feats = true;
    return RSS_OK;
}

// forest over the device-resident feature matrix
rss_status frame_predict(rss_ctx* ctx, const float* feats_dev, int n) {
    FrameState& f = ctx->fr;
    const ForestDev& F = ctx->forest;
    if (!F.loaded) return ctx->fail(RSS_ERR_STATE, "no forest loaded");
    const int D = ctx->cfg.feature_length();
    RSS_CU(ctx, f.leaf_ids.reserve((size_t)F.T * (n > 0 ? n : 1) * 4));
    RSS_CU(ctx, f.post.reserve((size_t)(n > 0 ? n : 1) * F.sumC * 4));
    launch_forest_traverse(ctx, ctx->s0, F.nodes.as<Node>(), F.tree_off_dev.as<int>(), F.T, feats_dev, D, n, n,
                           f.leaf_ids.as<int>());
    launch_forest_posterior(ctx, ctx->s0, F.nodes.as<Node>(), F.tree_off_dev.as<int>(), F.T, F.leaves.as<float>(),
                            F.sumC, f.leaf_ids.as<int>(), n, n, f.post.as<float>());
    RSS_CU(ctx, cudaGetLastError());
    return RSS_OK;
}

static float ev_ms(cudaEvent_t a, cudaEvent_t b) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
}

// segmenter.cpp:349-434 on the device; leaves posteriors [layer][y][x][class] resident
rss_status frame_segment(rss_ctx* ctx, const uint8_t* rgb, const uint16_t* depth, int W, int H, const float* Kinv,
                         const float* R, const float* t, float fill) {
    const HostConfig& cfg = ctx->cfg;
    const ForestDev& F = ctx->forest;
    if (!F.loaded) return ctx->fail(RSS_ERR_STATE, "no forest loaded");
    const int stride = cfg.rf_stride;
    if (W % stride || H % stride) return ctx->fail(RSS_ERR_INVALID, "image size must be a multiple of rf_prediction_stride");
    cudaEventRecord(ctx->ev[0], ctx->s0);
    rss_status st = frame_upload(ctx, rgb, depth, W, H);
    if (st != RSS_OK) return st;
    cudaEventRecord(ctx->ev[1], ctx->s0);
    return frame_segment_resident(ctx, Kinv, R, t, fill);
}

// the same on the frame that frame_upload() made resident (events 0/1 recorded by the caller)
rss_status frame_segment_resident(rss_ctx* ctx, const float* Kinv, const float* R, const float* t, float fill) {
    rss_status st = frame_segment_begin(ctx, Kinv, R, t);
    if (st != RSS_OK) return st;
    return frame_segment_finish(ctx, fill);
}
// first half: Lab + border on s0, cloud and the normals preparation on s1 (ev_cloud marks the point cloud)
rss_status frame_segment_begin(rss_ctx* ctx, const float* Kinv, const float* R, const float* t) {
    FrameState& f = ctx->fr;
    const HostConfig& cfg = ctx->cfg;
    if (!ctx->forest.loaded) return ctx->fail(RSS_ERR_STATE, "no forest loaded");
    const int stride = cfg.rf_stride, W = f.W, H = f.H;
    if (W % stride || H % stride) return ctx->fail(RSS_ERR_INVALID, "image size must be a multiple of rf_prediction_stride");
    if (Kinv) {  // NULL: the caller has already set the pose (keyframe path: before the graph capture / replay)
        rss_status st = frame_set_pose(ctx, Kinv, R, t);
        if (st != RSS_OK) return st;
    }
    return frame_prepare(ctx, cfg.depth_min, cfg.depth_max);
}
// second half: samples, features, forest, low-res scatter, upsample
// unary_out != NULL: the up-sampled values go, negated, straight into that [pixel][unary_stride] energy matrix and the
// [layer][y][x][class] posterior vector is not produced (keyframe path)
rss_status frame_segment_finish(rss_ctx* ctx, float fill, float* unary_out, int unary_stride) {
    FrameState& f = ctx->fr;
    const HostConfig& cfg = ctx->cfg;
    const ForestDev& F = ctx->forest;
    const int stride = cfg.rf_stride, W = f.W, H = f.H;
    f.stride = stride; f.gw = rss_div_up(W, stride); f.gh = rss_div_up(H, stride);
    f.n_samples = -1;  // no sample list on this path: the forest kernel walks the stride grid itself
    f.have_feats = false;
    const size_t low_elems = (size_t)f.gw * f.gh * F.sumC;
    RSS_CU(ctx, f.lowres.reserve(low_elems * 4));
    if (!unary_out) RSS_CU(ctx, f.posteriors.reserve((size_t)W * H * F.sumC * 4));
    ctx->mark(2);
    const int ncolor = cfg.use_color ? 3 * cfg.patch_size_reduce * cfg.patch_size_reduce : 0;
    int pos = ncolor, pos_depth = -1, pos_height = -1, pos_normal = -1;
    if (cfg.use_depth) pos_depth = pos++;
    if (cfg.use_height) pos_height = pos++;
    if (cfg.use_normal) pos_normal = pos++;
    const float dmin_mm = (float)(cfg.depth_min * 1000.0), dmax_mm = (float)(cfg.depth_max * 1000.0);  // feature_extractor.h:43-44
    launch_forest_frame_lowres(ctx, ctx->s0, F.nodes.as<Node>(), F.tree_off_dev.as<int>(), F.T, F.leaves.as<float>(), F.L, F.C,
                               f.lab.as<uchar4>(), f.depth.as<uint16_t>(), f.xyz.as<float4>(), f.dist_b.as<float>(),
                               f.integ.as<double>(), f.integ_cnt.as<int>(), ctx->tapx.as<ResizeTap>(), ctx->tapy.as<ResizeTap>(),
                               ctx->feat_xy.as<uint16_t>(), W, H, cfg.patch_size, cfg.patch_size_reduce, ncolor, pos_depth,
                               pos_height, pos_normal, stride, dmin_mm, dmax_mm, fill, f.lowres.as<float>());
    ctx->mark(3);
    if (unary_out) launch_upsample(ctx, ctx->s0, f.lowres.as<float>(), f.gw, f.gh, W, H, F.L, F.C, unary_out, unary_stride);
    else launch_upsample(ctx, ctx->s0, f.lowres.as<float>(), f.gw, f.gh, W, H, F.L, F.C, f.posteriors.as<float>());
    ctx->mark(4);
    RSS_CU(ctx, cudaGetLastError());
    f.have_post = unary_out == nullptr;
    return RSS_OK;
}

void frame_collect_timings(rss_ctx* ctx, bool with_d2h) {
    ctx->tim.h2d_ms = ev_ms(ctx->ev[0], ctx->ev[1]);
    ctx->tim.features_ms = ev_ms(ctx->ev[1], ctx->ev[2]);
    ctx->tim.forest_ms = ev_ms(ctx->ev[2], ctx->ev[3]);
    ctx->tim.upsample_ms = ev_ms(ctx->ev[3], ctx->ev[4]);
    if (with_d2h) {
        ctx->tim.d2h_ms = ev_ms(ctx->ev[4], ctx->ev[5]);
        ctx->tim.total_ms = ev_ms(ctx->ev[0], ctx->ev[5]);
    }
}

}  // namespace rss

extern "C" rss_status rss_extract_features(rss_ctx* ctx, const uint8_t* rgb, const uint16_t* depth_mm, int W, int H,
                                           int stride, const float Kinv[9], const float R[9], const float t[3],
                                           float dmin, float dmax, int extract_type, const int8_t* labels,
                                           int n_label_layers, float* feats, int* xs, int* ys, int* out_labels,
                                           int* n_samples) {
    if (!ctx) return RSS_ERR_INVALID;
    if (!Kinv || !R || !t || stride < 1) return ctx->fail(RSS_ERR_INVALID, "null calibration or bad stride");
    if (extract_type < RSS_WITH_ANY_LABEL || extract_type > RSS_NO_LABEL) return ctx->fail(RSS_ERR_INVALID, "bad extract_type");
    if (extract_type != RSS_NO_LABEL && (!labels || n_label_layers < 1 || n_label_layers > RSS_MAX_LAYERS))
        return ctx->fail(RSS_ERR_INVALID, "labelled extraction needs label planes");
    if (extract_type == RSS_NO_LABEL) { labels = nullptr; n_label_layers = 0; }
    RSS_CU(ctx, cudaSetDevice(ctx->device));
    rss_status st = frame_upload(ctx, rgb, depth_mm, W, H);
    if (st != RSS_OK) return st;
    FrameState& f = ctx->fr;
    const size_t NP = (size_t)W * H;
    if (labels) {
        RSS_CU(ctx, f.labels.reserve(NP * n_label_layers));
        RSS_CU(ctx, cudaMemcpyAsync(f.labels.ptr, labels, NP * n_label_layers, cudaMemcpyHostToDevice, ctx->s0));
    }
    st = frame_set_pose(ctx, Kinv, R, t);
    if (st != RSS_OK) return st;
    st = frame_prepare(ctx, dmin, dmax);
    if (st != RSS_OK) return st;
    st = frame_extract(ctx, stride, dmin, dmax, extract_type, labels ? f.labels.as<int8_t>() : nullptr, n_label_layers);
    if (st != RSS_OK) return st;
    const int n = f.n_samples, D = ctx->cfg.feature_length();
    if (n_samples) *n_samples = n;
    if (n > 0) {
        if (feats) RSS_CU(ctx, cudaMemcpyAsync(feats, f.feats.ptr, (size_t)n * D * 4, cudaMemcpyDeviceToHost, ctx->s0));
        if (xs) RSS_CU(ctx, cudaMemcpyAsync(xs, f.xs.ptr, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->s0));
        if (ys) RSS_CU(ctx, cudaMemcpyAsync(ys, f.ys.ptr, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->s0));
        if (out_labels && labels)
            RSS_CU(ctx, cudaMemcpyAsync(out_labels, f.slabels.ptr, (size_t)n * n_label_layers * 4, cudaMemcpyDeviceToHost, ctx->s0));
    }
    RSS_CU(ctx, cudaStreamSynchronize(ctx->s0));
    return RSS_OK;
}

extern "C" rss_status rss_upload_frame(rss_ctx* ctx, const uint8_t* rgb, const uint16_t* depth_mm, int W, int H) {
    if (!ctx) return RSS_ERR_INVALID;
    if (!rgb || !depth_mm) return ctx->fail(RSS_ERR_INVALID, "null frame");
    RSS_CU(ctx, cudaSetDevice(ctx->device));
    rss_status st = frame_upload(ctx, rgb, depth_mm, W, H);
    if (st != RSS_OK) return st;
    RSS_CU(ctx, cudaStreamSynchronize(ctx->s0));
    return RSS_OK;
}

extern "C" rss_status rss_frame_intermediates(rss_ctx* ctx, uint8_t* lab, float* xyz, float* normals) {
    if (!ctx) return RSS_ERR_INVALID;
    FrameState& f = ctx->fr;
    RSS_CU(ctx, cudaSetDevice(ctx->device));
    const int W = f.W, H = f.H, P = ctx->cfg.patch_size;
    const size_t NP = (size_t)W * H;
    if (lab) {
        if (!f.have_lab) return ctx->fail(RSS_ERR_STATE, "no Lab image resident");
        const size_t nb = (size_t)(W + 2 * P) * (H + 2 * P);
        std::vector<uchar4> tmp(nb);
        RSS_CU(ctx, cudaMemcpyAsync(tmp.data(), f.lab.ptr, nb * sizeof(uchar4), cudaMemcpyDeviceToHost, ctx->s0));
        RSS_CU(ctx, cudaStreamSynchronize(ctx->s0));
        for (size_t i = 0; i < nb; i++) { lab[3 * i] = tmp[i].x; lab[3 * i + 1] = tmp[i].y; lab[3 * i + 2] = tmp[i].z; }
    }
    if (xyz) {
        if (!f.have_cloud) return ctx->fail(RSS_ERR_STATE, "no point cloud resident");
        std::vector<float4> tmp(NP);
        RSS_CU(ctx, cudaMemcpyAsync(tmp.data(), f.xyz.ptr, NP * sizeof(float4), cudaMemcpyDeviceToHost, ctx->s0));
        RSS_CU(ctx, cudaStreamSynchronize(ctx->s0));
        for (size_t i = 0; i < NP; i++) { xyz[3 * i] = tmp[i].x; xyz[3 * i + 1] = tmp[i].y; xyz[3 * i + 2] = tmp[i].z; }
    }
    if (normals) {
        if (!f.have_integral) return ctx->fail(RSS_ERR_STATE, "no integral images resident");
        RSS_CU(ctx, f.normals.reserve(NP * 3 * 4));
        launch_normals_full(ctx, ctx->s0, f.xyz.as<float4>(), f.dist_b.as<float>(), f.integ.as<double>(),
                            f.integ_cnt.as<int>(), W, H, f.normals.as<float>());
        RSS_CU(ctx, cudaMemcpyAsync(normals, f.normals.ptr, NP * 3 * 4, cudaMemcpyDeviceToHost, ctx->s0));
        RSS_CU(ctx, cudaStreamSynchronize(ctx->s0));
    }
    return RSS_OK;
}

extern "C" rss_status rss_forest_predict(rss_ctx* ctx, const float* feats, int n, int32_t* leaf_ids, float* log_post) {
    if (!ctx) return RSS_ERR_INVALID;
    if (n < 0) return ctx->fail(RSS_ERR_INVALID, "negative sample count");
    RSS_CU(ctx, cudaSetDevice(ctx->device));
    FrameState& f = ctx->fr;
    const ForestDev& F = ctx->forest;
    if (!F.loaded) return ctx->fail(RSS_ERR_STATE, "no forest loaded");
    const int D = ctx->cfg.feature_length();
    if (feats) {
        RSS_CU(ctx, f.feats.reserve((size_t)(n > 0 ? n : 1) * D * 4));
        RSS_CU(ctx, cudaMemcpyAsync(f.feats.ptr, feats, (size_t)n * D * 4, cudaMemcpyHostToDevice, ctx->s0));
        f.have_feats = true;
        f.n_samples = -1;  // the compacted sample list no longer matches
    } else {
        if (!f.have_feats || f.n_samples != n) return ctx->fail(RSS_ERR_STATE, "no resident features for this sample count");
    }
    if (n == 0) return RSS_OK;
    rss_status st = frame_predict(ctx, f.feats.as<float>(), n);
    if (st != RSS_OK) return st;
    if (leaf_ids) RSS_CU(ctx, cudaMemcpyAsync(leaf_ids, f.leaf_ids.ptr, (size_t)F.T * n * 4, cudaMemcpyDeviceToHost, ctx->s0));
    if (log_post) RSS_CU(ctx, cudaMemcpyAsync(log_post, f.post.ptr, (size_t)n * F.sumC * 4, cudaMemcpyDeviceToHost, ctx->s0));
    RSS_CU(ctx, cudaStreamSynchronize(ctx->s0));
    return RSS_OK;
}

extern "C" rss_status rss_posteriors_keep(rss_ctx* ctx, int slot) {
    if (!ctx) return RSS_ERR_INVALID;
    if (slot < 0 || slot > 4096) return ctx->fail(RSS_ERR_INVALID, "bad slot");
    FrameState& f = ctx->fr;
    if (!f.have_post) return ctx->fail(RSS_ERR_STATE, "no resident posteriors (call rss_segment_frame first)");
    RSS_CU(ctx, cudaSetDevice(ctx->device));
    if ((int)f.kept.size() <= slot) { f.kept.resize(slot + 1); f.kept_npix.resize(slot + 1, 0); }
    const size_t bytes = (size_t)f.W * f.H * ctx->forest.sumC * 4;
    RSS_CU(ctx, f.kept[slot].reserve(bytes));
    RSS_CU(ctx, cudaMemcpyAsync(f.kept[slot].ptr, f.posteriors.ptr, bytes, cudaMemcpyDeviceToDevice, ctx->s0));
    f.kept_npix[slot] = f.W * f.H;
    return RSS_OK;
}

extern "C" rss_status rss_segment_frame(rss_ctx* ctx, const uint8_t* rgb, const uint16_t* depth_mm, int W, int H,
                                        const float Kinv[9], const float R[9], const float t[3], float fill,
                                        float* posteriors) {
    if (!ctx) return RSS_ERR_INVALID;
    if (!Kinv || !R || !t) return ctx->fail(RSS_ERR_INVALID, "null calibration");
    RSS_CU(ctx, cudaSetDevice(ctx->device));
    rss_status st = frame_segment(ctx, rgb, depth_mm, W, H, Kinv, R, t, fill);
    if (st != RSS_OK) return st;
    if (posteriors)
        RSS_CU(ctx, cudaMemcpyAsync(posteriors, ctx->fr.posteriors.ptr, (size_t)W * H * ctx->forest.sumC * 4,
                                    cudaMemcpyDeviceToHost, ctx->s0));
    cudaEventRecord(ctx->ev[5], ctx->s0);
    RSS_CU(ctx, cudaStreamSynchronize(ctx->s0));
    frame_collect_timings(ctx, true);
    return RSS_OK;
}

// ---------------------------------------------------------------------------------------------------
// srv/SingleFrameSegmentation.srv served by this library: the node's external-semantics client
// (src/segmenter.cpp:446-514) sends the RGB8 image and, as "depth", the RECTIFIED CLOUD it computes at :463-488
//     p = (extrinsic.linear() * intrinsic_inverse) * (d x, d y, d)^T + t,  d = depth_mm / 1000.f,  NaN where d < 0.5 or > 15
// (TYPE_32FC3, world frame), and expects float32[] label_distribution = [layer][y][x][class] (:497-500, the layout
// scripts/single_frame_segmentation_server.py:47 builds).  The raw depth is recovered exactly: the camera-frame z of p is
// d itself (the last row of the inverse intrinsics is (0, 0, 1)), and 1000 d is within 0.02 of the integer it came from.
// ---------------------------------------------------------------------------------------------------
extern "C" rss_status rss_service_single_frame(rss_ctx* ctx, const uint8_t* rgb, const float* depth3d, int W, int H,
                                               const float Kinv[9], const float R[9], const float t[3],
                                               float* label_distribution) {
    if (!ctx) return RSS_ERR_INVALID;
    if (!rgb || !depth3d || !Kinv || !R || !t) return ctx->fail(RSS_ERR_INVALID, "null request field");
    if (W <= 0 || H <= 0) return ctx->fail(RSS_ERR_INVALID, "bad image size");
    try {
        std::vector<uint16_t>& depth = ctx->service_depth;
        depth.resize((size_t)W * H);
        const double r20 = R[2], r21 = R[5], r22 = R[8];  // third row of R^T = third column of R (row-major)
        for (size_t i = 0, n = (size_t)W * H; i < n; i++) {
            const float* p = depth3d + 3 * i;
            uint16_t d = 0;  // NaN (the node's invalid marker) -> depth 0 = invalid for the feature extractor
            if (p[0] == p[0] && p[1] == p[1] && p[2] == p[2]) {
                const double z = r20 * ((double)p[0] - (double)t[0]) + r21 * ((double)p[1] - (double)t[1]) +
                                 r22 * ((double)p[2] - (double)t[2]);
                const double mm = std::nearbyint(z * 1000.0);
                d = mm <= 0.0 ? 0 : (mm >= 65535.0 ? 65535 : (uint16_t)mm);
            }
            depth[i] = d;
        }
        return rss_segment_frame(ctx, rgb, depth.data(), W, H, Kinv, R, t, 0.0f, label_distribution);  // node: fill 0 (:357-362)
    } catch (const std::exception& e) {
        return ctx->fail(RSS_ERR_INVALID, e.what());
    }
}

// ---------------------------------------------------------------------------------------------------
// GPU forest training (train.cu)
// ---------------------------------------------------------------------------------------------------
extern "C" void rss_train_params_default(rss_train_params* p) {
    if (!p) return;
    p->num_trees = 4; p->max_depth = 30; p->min_split_examples = 50; p->min_child_split_examples = 1;
    p->num_features = 0; p->use_bootstrap = 1; p->num_bootstrap_examples = 0; p->smoothing = 1.0f; p->seed = 1;
}
extern "C" rss_status rss_forest_train(rss_ctx* ctx, const float* feats, int n, int D, const int32_t* labels, int n_layers,
                                       const int* class_counts, const rss_train_params* params, const char* out_dat_path,
                                       rss_train_stats* stats) {
    if (!ctx) return RSS_ERR_INVALID;
    if (!feats || !labels || !class_counts || !params) return ctx->fail(RSS_ERR_INVALID, "train: null argument");
    if (n <= 0 || D <= 0 || n_layers <= 0 || n_layers > RSS_MAX_LAYERS) return ctx->fail(RSS_ERR_INVALID, "train: bad sizes");
    if (params->num_trees <= 0 || params->num_trees > 4096) return ctx->fail(RSS_ERR_INVALID, "train: bad tree count");
    RSS_CU(ctx, cudaSetDevice(ctx->device));
    cudaEvent_t e0, e1;
    RSS_CU(ctx, cudaEventCreate(&e0));
    RSS_CU(ctx, cudaEventCreate(&e1));
    cudaEventRecord(e0, ctx->s0);
    const rss_status st = forest_train(ctx, feats, n, D, labels, n_layers, class_counts, *params, out_dat_path, stats);
    cudaEventRecord(e1, ctx->s0);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (stats && st == RSS_OK) stats->train_ms = ms;
    return st;
}
