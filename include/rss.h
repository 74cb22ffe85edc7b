/* rss.h - C ABI of librss.so: the B200-native per-keyframe inference path of RovinaSemanticSegmentation.
 *
 * The reference has no FFI layer; its "boundary" for this path is three in-process C++ class surfaces
 * (SURVEY.md section 8b).  Each entry point below names the reference interface it replaces
 * (file:line relative to the reference tree).  Thin C++ adapters with the reference's own class names
 * (Features::FeatureExtractor, libf::RandomForest, DenseCRF, ...) live in
 * rovinasemanticsegmentation_b200/host/ and call only these functions.
 *
 * Conventions
 *  - plain pointers and sizes, no C++/torch types; integer status codes, no exceptions across the ABI;
 *  - the caller owns every host buffer; "host" pointers are ordinary (pageable or pinned) memory;
 *  - a context is bound to one GPU and is NOT thread-safe: one context per GPU per host thread, which
 *    matches the reference's one-worker-per-stage model (src/segmenter.cpp:227-232);
 *  - every call is synchronous from the caller's view unless its comment says otherwise;
 *  - there is no CPU fallback: without a CUDA device rss_create fails with RSS_ERR_CUDA.
 *  - matrices follow the reference's Eigen column-major convention: an "M x N" matrix stores the M
 *    values of point i contiguously at [i*M, (i+1)*M).
 */
#ifndef RSS_H
#define RSS_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rss_ctx rss_ctx;
typedef struct rss_crf rss_crf;
typedef int rss_status;

enum {
    RSS_OK = 0,
    RSS_ERR_INVALID = 1,  /* bad argument */
    RSS_ERR_IO = 2,       /* file missing / unreadable */
    RSS_ERR_CONFIG = 3,   /* config key missing (reference: Utils::KeyNotFoundException, include/config.h:13-24) */
    RSS_ERR_MODEL = 4,    /* malformed forest file */
    RSS_ERR_CUDA = 5,     /* CUDA runtime error or no device */
    RSS_ERR_CAPACITY = 6, /* lattice hash table overflow that could not be resolved */
    RSS_ERR_STATE = 7     /* call order violated (e.g. predict before extract) */
};

/* reference: enum class ExtractType, include/feature_extractor.h:21 */
enum { RSS_WITH_ANY_LABEL = 0, RSS_WITH_POSITIVE_LABEL = 1, RSS_NO_LABEL = 2 };
/* reference: enum NormalizationType, third-party/densecrf/include/pairwise.h */
enum { RSS_NO_NORMALIZATION = 0, RSS_NORMALIZE_BEFORE = 1, RSS_NORMALIZE_AFTER = 2, RSS_NORMALIZE_SYMMETRIC = 3 };

#define RSS_MAX_LAYERS 8

typedef struct {
    /* Features::FeatureExtractor state, include/feature_extractor.h:29-39 */
    int feature_color_patch, feature_depth, feature_height, feature_normal;
    int patch_size, patch_size_reduce, feature_length;
    /* forest (third-party/libforest): trees, nodes, layers, classes per layer */
    int num_trees, total_nodes, total_leaves;
    int layer_count, class_counts[RSS_MAX_LAYERS], total_classes;
    /* Segmenter state parsed from the config, src/segmenter.cpp:73-98,120-127 */
    int unknown_label[RSS_MAX_LAYERS];
    int use_dense_crf, dcrf_iterations, rf_prediction_stride;
    float dcrf_xyz_kernel, dcrf_rgb_kernel, dcrf_kernel_weight, depth_min, depth_max;
    int cuda_device, sm_count;
} rss_info;

/* ---------------------------------------------------------------------------------------------------
 * Context = Segmenter constructor state (src/segmenter.cpp:70-127): parses the config JSON
 * (resources/config.json keys) and the libforest binary model
 * (RandomForest::read, third-party/libforest/src/classifier.cpp:222-235; io.h:43-108), flattens the
 * trees to a structure-of-arrays node table + dense leaf table and uploads them.
 * forest_dat_path may be NULL (feature extraction / CRF only).
 * ------------------------------------------------------------------------------------------------- */
rss_status rss_create(const char* config_json_path, const char* forest_dat_path, int cuda_device, rss_ctx** out);
rss_status rss_destroy(rss_ctx* ctx);
/* RandomForest::read(std::istream&) (classifier.cpp:222-235) for a context created without a model, or to swap
 * models: from a file, or from a buffer holding the bytes of the libforest stream. */
rss_status rss_load_forest(rss_ctx* ctx, const char* forest_dat_path);
rss_status rss_load_forest_memory(rss_ctx* ctx, const void* bytes, size_t size);
rss_status rss_get_info(const rss_ctx* ctx, rss_info* out);
/* Host-only parse of a config JSON (Utils::Config + the Segmenter / FeatureExtractor constructors, src/config.cpp,
 * src/segmenter.cpp:70-127, include/feature_extractor.h:29-39): fills the config-derived fields of rss_info without
 * touching a CUDA device, so a config can be validated before a context exists.  Same status codes and messages as
 * rss_create (message via rss_last_error(NULL)). */
rss_status rss_parse_config(const char* config_json_path, rss_info* out);
/* message of the last failing call on this context ("" if none); ctx may be NULL for create failures */
const char* rss_last_error(const rss_ctx* ctx);
const char* rss_status_string(rss_status s);

/* ---------------------------------------------------------------------------------------------------
 * FeatureExtractor::extract (include/feature_extractor.h:41-291).
 *   rgb      H*W*3 u8, in the channel order the reference receives (it feeds RGB to CV_BGR2Lab; kept)
 *   depth_mm H*W u16, millimetres
 *   Kinv, R  row-major 3x3 (Calibration::_intrinsic_inverse, _extrinsic.linear()), t = translation
 *   labels   n_label_layers planes of H*W int8 (reference label_type = char) or NULL for RSS_NO_LABEL
 * Outputs (host, each may be NULL): feats [n][D] float, xs/ys [n], out_labels [n][n_label_layers].
 * Capacity needed: ceil(W/stride)*ceil(H/stride) samples.  Samples come in raster order like the
 * reference's loop (:58-71).  The features stay resident on the device for rss_forest_predict.
 * ------------------------------------------------------------------------------------------------- */
rss_status rss_extract_features(rss_ctx* ctx, const uint8_t* rgb, const uint16_t* depth_mm, int W, int H, int stride,
                                const float Kinv[9], const float R[9], const float t[3], float dmin, float dmax,
                                int extract_type, const int8_t* labels, int n_label_layers, float* feats, int* xs,
                                int* ys, int* out_labels, int* n_samples);
/* Intermediate products of the last extract / segment call, for parity tests (each may be NULL):
 * lab  (H+2P)*(W+2P)*3 u8  = cvtColor + copyMakeBorder (:129-130);  xyz H*W*3 (:200-232);
 * normals H*W*3 (PCL IntegralImageNormalEstimation, :256-261), computed for every pixel on request. */
rss_status rss_frame_intermediates(rss_ctx* ctx, uint8_t* lab, float* xyz, float* normals);

/* ---------------------------------------------------------------------------------------------------
 * RandomForest::multiClassLogPosterior over a batch (classifier.cpp:187-208) + DecisionTree::findLeafNode
 * (:97-117).  feats: host [n][D] or NULL = use the device-resident features of the last
 * rss_extract_features (then n must equal its sample count).  leaf_ids [T][n] and log_post [n][sumC]
 * (layers concatenated) are host buffers, each may be NULL.
 * ------------------------------------------------------------------------------------------------- */
rss_status rss_forest_predict(rss_ctx* ctx, const float* feats, int n, int32_t* leaf_ids, float* log_post);

/* ---------------------------------------------------------------------------------------------------
 * Frame worker body, Segmenter::processFramesFromQueueInternalRF (src/segmenter.cpp:349-434) and
 * test_multi.cpp:166-199: extract(NO_LABEL, stride = rf_prediction_stride) -> forest -> low-res scatter
 * (unsampled pixels = fill: 0 in the node, -1000 in the test tool) -> cv::resize to W x H -> flatten.
 * posteriors: host [layer][y][x][class] (= srv/SingleFrameSegmentation.srv label_distribution), may be
 * NULL to keep the result on the device for rss_crf_* / rss_unary_accumulate.
 * ------------------------------------------------------------------------------------------------- */
rss_status rss_segment_frame(rss_ctx* ctx, const uint8_t* rgb, const uint16_t* depth_mm, int W, int H,
                             const float Kinv[9], const float R[9], const float t[3], float fill, float* posteriors);

/* ---------------------------------------------------------------------------------------------------
 * Forest training on the GPU: libf::DecisionTreeLearner::learn + updateMultiHistograms and RandomForestLearner::learn
 * (third-party/libforest/src/learning.cpp:410-916, 963-1012, 1031-1073) with the learner set-up of src/train.cpp:225-249.
 * feats [n][D] and labels [n][n_layers] are host buffers (the DataStorage of train.cpp: rss_extract_features with
 * RSS_WITH_POSITIVE_LABEL produces both); class_counts[l] = classes of layer l.  The model is written to out_dat_path in
 * the libforest binary format (RandomForest::write, classifier.cpp:210-220) - the file rss_load_forest and the
 * reference's RandomForest::read accept.  Defaults of rss_train_params_default = train.cpp + DecisionTreeLearner():
 * bootstrap of n examples, ceil(sqrt(D)) features per node, min_child_split_examples 1, smoothing 1.
 * The learner is seeded and deterministic (the reference draws from std::random_device).
 * ------------------------------------------------------------------------------------------------- */
typedef struct rss_train_params {
    int num_trees;                 /* config num_trees */
    int max_depth;                 /* config max_depth (a node splits while depth <= max_depth, learning.cpp:525) */
    int min_split_examples;        /* config min_split_sample */
    int min_child_split_examples;  /* 1 */
    int num_features;              /* 0 = ceil(sqrt(D)) (autoconf, learning.cpp:363-368) */
    int use_bootstrap;             /* 1 */
    int num_bootstrap_examples;    /* 0 = n */
    float smoothing;               /* 1 */
    uint64_t seed;
} rss_train_params;
typedef struct rss_train_stats {
    int trees, features_per_node, bootstrap_examples;
    long long nodes, levels;
    double train_ms;               /* device + host time of the whole call */
} rss_train_stats;
void rss_train_params_default(rss_train_params* p);
rss_status rss_forest_train(rss_ctx* ctx, const float* feats, int n, int D, const int32_t* labels, int n_layers,
                            const int* class_counts, const rss_train_params* params, const char* out_dat_path,
                            rss_train_stats* stats /* may be NULL */);

/* ---------------------------------------------------------------------------------------------------
 * srv/SingleFrameSegmentation.srv (the node's external-semantics service, client side src/segmenter.cpp:446-514,
 * stub server scripts/single_frame_segmentation_server.py:12-52): request = the RGB8 image and, as "depth", the
 * rectified world-frame cloud the node computes at :463-488 (TYPE_32FC3, NaN where the raw depth is outside
 * [0.5, 15] m); response = float32[] label_distribution in [layer][y][x][class] order.  depth3d: H*W*3 float.
 * Kinv / R / t are the calibration of the camera the frame came from (the node rectifies with the same one).
 * Equal to rss_segment_frame(fill = 0) on the raw depth image the cloud was made from.
 * ------------------------------------------------------------------------------------------------- */
rss_status rss_service_single_frame(rss_ctx* ctx, const uint8_t* rgb, const float* depth3d, int W, int H,
                                    const float Kinv[9], const float R[9], const float t[3], float* label_distribution);

/* ---------------------------------------------------------------------------------------------------
 * DenseCRF (third-party/densecrf/include/densecrf.h:36-121).  N points; n_layers label layers with
 * M[l] labels each share every lattice (the reference rebuilds the same lattice per layer,
 * src/segmenter.cpp:639-643; results are identical).  rss_crf_create(ctx,N,M) = one layer.
 * ------------------------------------------------------------------------------------------------- */
rss_status rss_crf_create(rss_ctx* ctx, int N, int M, rss_crf** out);
rss_status rss_crf_create_layers(rss_ctx* ctx, int N, int n_layers, const int* M, rss_crf** out);
/* DenseCRF::setUnaryEnergy(MatrixXf) (densecrf.cpp:88-90): U = M[layer] x N column-major energies, host */
rss_status rss_crf_set_unary(rss_crf* crf, int layer, const float* U);
/* DenseCRF::addPairwiseEnergy(features d x N, new PottsCompatibility(w), DIAG_KERNEL, norm_type)
 * (densecrf.cpp:54-60, pairwise.cpp:40-62,170-172): builds the permutohedral lattice on the device. */
rss_status rss_crf_add_pairwise(rss_crf* crf, const float* feats, int d, float potts_w, int norm_type);
/* DenseCRF2D::addPairwiseGaussian / addPairwiseBilateral (densecrf.cpp:61-81); N must equal W*H */
rss_status rss_crf_add_pairwise_gaussian(rss_crf* crf, int W, int H, float sx, float sy, float potts_w);
rss_status rss_crf_add_pairwise_bilateral(rss_crf* crf, int W, int H, float sx, float sy, float sr, float sg,
                                          float sb, const uint8_t* im, float potts_w);
/* DenseCRF::inference(n) (densecrf.cpp:115-131) for one layer (or all with layer = -1).
 * Q: host M[layer] x N (layers concatenated for -1), may be NULL.
 * labels: host N bytes per layer, gated argmax of src/segmenter.cpp:645-657 (label = argmax if Q > 2/M
 * else unknown_label[layer]); pass unknown_label < 0 for the plain DenseCRF::map argmax (densecrf.cpp:200-208). */
rss_status rss_crf_inference(rss_crf* crf, int layer, int iters, float* Q, uint8_t* labels, const int* unknown_label);
/* DenseCRF::startInference / stepInference / currentMap (densecrf.cpp:178-211): Q stays on the device between
 * calls; rss_crf_current copies the current marginals and/or their (gated) argmax back, arguments as above. */
rss_status rss_crf_start_inference(rss_crf* crf);
rss_status rss_crf_step_inference(rss_crf* crf, int steps);
rss_status rss_crf_current(rss_crf* crf, int layer, float* Q, uint8_t* labels, const int* unknown_label);
/* number of lattice vertices of pairwise term k (diagnostics) */
/* DenseCRF::gradient (third-party/densecrf/src/densecrf.cpp:238-297) with the LogLikelihood objective
 * (src/objective.cpp:36-52): runs `iters` mean-field iterations of `layer`, evaluates
 * objective = sum_i log(max(Q(gt_i, i) + robust, 1e-20)) / N over the points with 0 <= gt_i < M (gt: host [N]),
 * back-propagates through the iterations and returns the gradient w.r.t. the label-compatibility parameter (the Potts
 * weight, labelcompatibility.cpp:41-61) of every pairwise term: potts_grad[K] (may be NULL).  Kernel-parameter and unary
 * gradients (DIAG_KERNEL features, LogisticUnaryEnergy) are not part of the reference's inference path and not built. */
rss_status rss_crf_gradient(rss_crf* crf, int layer, int iters, const int32_t* gt, float robust, double* objective,
                            float* potts_grad);
rss_status rss_crf_lattice_size(rss_crf* crf, int k, int* vertices);
/* Diagnostics: the mean-field path of the next inference - fused point kernel or generic kernels; sorted = the fused path
 * runs over the sorted order of a point set whose own order is not coherent (local maps).  No reference counterpart. */
rss_status rss_crf_path(rss_crf* crf, int* fused, int* sorted);
/* Permutohedral::compute on pairwise term k without normalisation (permutohedral.cpp:596-604):
 * in/out host M x N with M = total labels of the CRF; for parity tests of splat/blur/slice. */
rss_status rss_crf_filter(rss_crf* crf, int k, const float* in, float* out);
rss_status rss_crf_destroy(rss_crf* crf);

/* ---------------------------------------------------------------------------------------------------
 * Map worker pieces, Segmenter::processMapFromQueue (src/segmenter.cpp:561-657).
 * rss_unary_accumulate: unaries[l](c, idx) += posterior[...] for every pixel with index >= 0 (:597-616).
 *   index_image host npix int32 (the fps_mapper projector's output), posteriors: host
 *   [layer][pixel][class] or NULL = the device-resident result of the last rss_segment_frame.
 *   The accumulated unaries live in the CRF (log-posteriors summed; the energy is their negation, :642).
 * rss_crf_add_pairwise_xyzrgb: the 6-D feature matrix of :629-637 (points xyz * wxyz, rgb * wrgb).
 * ------------------------------------------------------------------------------------------------- */
rss_status rss_crf_unary_reset(rss_crf* crf);
/* Drops every pairwise term (their device buffers are pooled and reused by the next ones): with rss_crf_unary_reset this
 * lets one CRF object serve a sequence of local maps of the same size without reallocating. */
rss_status rss_crf_clear_pairwise(rss_crf* crf);
rss_status rss_crf_unary_accumulate(rss_crf* crf, rss_ctx* frame_ctx, const int32_t* index_image, int npix,
                                    const float* posteriors);
rss_status rss_crf_add_pairwise_xyzrgb(rss_crf* crf, const float* xyz, const float* rgb, float wxyz, float wrgb,
                                       float potts_w);

/* The same map worker with the cloud, the key frames' posteriors and the index images all ON THE DEVICE
 * (src/segmenter.cpp:559-637): nothing but the cloud goes up and nothing but the labels comes down.
 *   rss_crf_set_cloud        uploads the local map's points (xyz, rgb: host [N][3] floats, N = the CRF's point count) once.
 *   rss_posteriors_keep      copies the posteriors left resident by the last rss_segment_frame into device slot `slot`
 *                            (slots grow on demand): the frame worker runs ahead of the map worker (:349-434 vs :518-719),
 *                            so a local map's key frames are segmented before the map is fused.
 *   rss_crf_project_accumulate  the projector of :581 as a z-buffer kernel (pinhole K row-major 3x3, key-frame pose R, t:
 *                            camera -> map, nearest pixel, z in [zmin, zmax], the nearest point wins a pixel, equal z: the
 *                            lower index), then unaries[l](c, idx) += posterior of that pixel (:597-616) straight from
 *                            slot `slot` (-1: the posteriors of the last rss_segment_frame).  index_image_out: optional
 *                            host [H][W] int32 copy of the index image (diagnostics / parity).
 *   rss_crf_add_pairwise_cloud  the 6-D kernel of :629-637 from the resident cloud.
 * fps_mapper's projector is not vendored in the reference; the projection rule above is this library's definition
 * (oracle: orc_project_zbuffer). */
rss_status rss_crf_set_cloud(rss_crf* crf, const float* xyz, const float* rgb);
rss_status rss_posteriors_keep(rss_ctx* ctx, int slot);
rss_status rss_crf_project_accumulate(rss_crf* crf, rss_ctx* frame_ctx, int slot, int W, int H, const float K[9],
                                      const float R[9], const float t[3], float zmin, float zmax, int32_t* index_image_out);
rss_status rss_crf_add_pairwise_cloud(rss_crf* crf, float wxyz, float wrgb, float potts_w);

/* ---------------------------------------------------------------------------------------------------
 * Fused keyframe: rss_segment_frame, then per layer a DenseCRF over the frame's W*H pixels with
 * unary = -posteriors, a Gaussian kernel on the frame's 3-D points (xyz / sigma_xyz, weight w_gauss) and
 * a bilateral kernel (x/sigma_px, y/sigma_px, rgb/sigma_rgb, weight w_bilateral), `iters` mean-field
 * iterations and the gated argmax.  Everything stays on the device between the H2D copy of
 * rgb/depth and the D2H copy of the labels.  labels: host [layer][H*W]; Q: host, layers concatenated, or NULL.
 * This is BASELINE.json configs[1]/[2] ("single-frame RF + DenseCRF, Gaussian 3-D + bilateral 5-D").
 * ------------------------------------------------------------------------------------------------- */
typedef struct {
    float sigma_xyz, w_gauss;             /* 3-D Gaussian kernel on the back-projected points */
    float sigma_px, sigma_rgb, w_bilateral; /* 5-D bilateral kernel */
    int iters;
    float fill;                           /* low-res fill value, see rss_segment_frame */
} rss_keyframe_params;
rss_status rss_segment_keyframe(rss_ctx* ctx, const uint8_t* rgb, const uint16_t* depth_mm, int W, int H,
                                const float Kinv[9], const float R[9], const float t[3],
                                const rss_keyframe_params* params, uint8_t* labels, float* Q);

/* Makes a frame resident on the device (H2D copy only).  Every call that takes rgb/depth_mm accepts NULL for
 * both to work on the resident frame instead: that is how the device-resident throughput is measured. */
/* The device part of rss_segment_keyframe is captured into a CUDA graph after the first calls with a given size and
 * parameter set and replayed afterwards (the pose may change from call to call; per-stage timings are only measured on
 * eager calls).  enable = 0 turns the replay off for this context (on by default; RSS_NO_GRAPH=1 disables it globally). */
rss_status rss_keyframe_graph(rss_ctx* ctx, int enable);
rss_status rss_upload_frame(rss_ctx* ctx, const uint8_t* rgb, const uint16_t* depth_mm, int W, int H);
/* Page-locked host memory for frame / label staging buffers (so that a C++ host needs no CUDA headers): copies
 * from and to such buffers run at full PCIe rate and asynchronously.  Not tied to a context. */
void* rss_host_alloc(size_t bytes);
void rss_host_free(void* p);
/* Diagnostics of the last rss_segment_keyframe: feature dimension and vertex count of its lattice k
 * (0 = Gaussian 3-D, 1 = bilateral 5-D); RSS_ERR_STATE before the first keyframe. */
rss_status rss_keyframe_lattice_info(rss_ctx* ctx, int k, int* d, int* vertices);

/* ---------------------------------------------------------------------------------------------------
 * Instrumentation: device time (CUDA events on the context's stream) of the stages of the last call,
 * in milliseconds, and the number of kernels this library launched since the context was created.
 * ------------------------------------------------------------------------------------------------- */
typedef struct {
    float h2d_ms, features_ms, forest_ms, upsample_ms, lattice_ms, meanfield_ms, d2h_ms, total_ms;
} rss_timings;
rss_status rss_get_timings(const rss_ctx* ctx, rss_timings* out);
uint64_t rss_kernel_launches(const rss_ctx* ctx);
/* Per-kernel device times: when enabled, every launch is bracketed by CUDA events on its own stream.
 * rss_profile_get returns entry `index` (0 .. count-1): kernel name, accumulated milliseconds, launch count;
 * RSS_ERR_INVALID past the end.  rss_profile_enable(ctx, x) also clears the accumulated numbers. */
rss_status rss_profile_enable(rss_ctx* ctx, int enable);
rss_status rss_profile_get(rss_ctx* ctx, int index, char* name, int name_cap, double* total_ms, uint64_t* launches);

#ifdef __cplusplus
}
#endif
#endif /* RSS_H */
