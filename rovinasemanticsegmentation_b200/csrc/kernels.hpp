// Launchers of the hand-written sm_100a kernels (defined in features.cu, normals.cu, forest.cu).
// All take the context (for launch accounting), a stream and raw device pointers.
#pragma once
#include "common.cuh"

namespace rss {

// ---- features.cu -----------------------------------------------------------------------------
// cvtColor(BGR2Lab) + copyMakeBorder(REFLECT, P) fused; lab is uchar4 per bordered pixel
void launch_lab_border(rss_ctx* c, cudaStream_t st, const uint8_t* rgb, int W, int H, int P, uchar4* lab);
// per-pixel back-projection (feature_extractor.h:200-232); pose_dev: M = R*Kinv (host, float) and t, xyz as float4
void launch_cloud(rss_ctx* c, cudaStream_t st, const uint16_t* depth, int W, int H, const PoseParams* pose_dev, float dmin,
                  float dmax, float4* xyz);
// sample selection (:56-121): flags per grid position
void launch_select(rss_ctx* c, cudaStream_t st, const uint16_t* depth, const int8_t* labels, int n_label_layers,
                   int extract_type, int W, int H, int stride, float dmin_mm, float dmax_mm, uint32_t* flags);
void launch_compact(rss_ctx* c, cudaStream_t st, const uint32_t* flags, const uint32_t* sidx, int W, int H,
                    int stride, const int8_t* labels, int n_label_layers, int* xs, int* ys, int* slabels);
// colour patch resample (:134-173) for the compacted samples -> feats[:, 0:3r^2)
void launch_patch_features(rss_ctx* c, cudaStream_t st, const uchar4* lab, const uint16_t* depth, int W, int H,
                           int P, int r, const ResizeTap* tapx, const ResizeTap* tapy, const int* xs,
                           const int* ys, int n, float* feats, int D);
// depth / height / normal-angle features (:180-197, :236-251, :265-291)
void launch_scalar_features(rss_ctx* c, cudaStream_t st, const uint16_t* depth, const float4* xyz,
                            const float* dist, const double* integ, const int* integ_cnt, int W, int H,
                            const int* xs, const int* ys, int n, float* feats, int D, int pos_depth,
                            int pos_height, int pos_normal);

// the frame worker's forest stage in one kernel: on-demand features, traversal, leaf rows summed in tree order, the
// per-layer low-resolution images written directly (`fill` where the depth is out of range)
void launch_forest_frame_lowres(rss_ctx* c, cudaStream_t st, const Node* nodes, const int* tree_off, int T, const float* leaves,
                                int L, const int* C, const uchar4* lab, const uint16_t* depth, const float4* xyz,
                                const float* dist, const double* integ, const int* cnt, const ResizeTap* tapx,
                                const ResizeTap* tapy, const uint16_t* feat_xy, int W, int H, int P, int r, int ncolor,
                                int pos_depth, int pos_height, int pos_normal, int stride, float dmin_mm, float dmax_mm,
                                float fill, float* lowres);

// ---- normals.cu (PCL IntegralImageNormalEstimation, AVERAGE_3D_GRADIENT) ---------------------------
// dist_b receives the final distance map; grad: float[6][H*W], fin: u8[2][H*W] scratch
void launch_normals_prepare(rss_ctx* c, cudaStream_t st, const float4* xyz, int W, int H, float* dist_a,
                            float* dist_b, double* integ, int* integ_cnt, float* grad, uint8_t* fin);
void launch_normals_full(rss_ctx* c, cudaStream_t st, const float4* xyz, const float* dist, const double* integ,
                         const int* integ_cnt, int W, int H, float* normals);
size_t integral_elems(int W, int H);

// ---- forest.cu ---------------------------------------------------------------------------------
void launch_forest_traverse(rss_ctx* c, cudaStream_t st, const Node* nodes, const int* tree_off, int T,
                            const float* feats, int D, int n, int ld, int* leaf_ids);
void launch_forest_posterior(rss_ctx* c, cudaStream_t st, const Node* nodes, const int* tree_off, int T,
                             const float* leaves, int sumC, const int* leaf_ids, int n, int ld, float* post);
// low-res scatter (segmenter.cpp:366-376) + fill
void launch_lowres_fill(rss_ctx* c, cudaStream_t st, float* lowres, size_t n, float fill);
void launch_lowres_scatter(rss_ctx* c, cudaStream_t st, const float* post, int sumC, const int* xs, const int* ys,
                           int n, int stride, int gw, int gh, int L, const int* C, float* lowres);
// cv::resize(INTER_LINEAR) 32FC(C) to W x H + flatten to [layer][y][x][class] (segmenter.cpp:380-431)
// unary_stride > 0: write -value into a [pixel][unary_stride] energy matrix instead (keyframe path)
void launch_upsample(rss_ctx* c, cudaStream_t st, const float* lowres, int gw, int gh, int W, int H, int L,
                     const int* C, float* posteriors, int unary_stride = 0);

// sort.cu: radix sort of (u64 key, u32 value) pairs (CUB); tmp = caller-owned scratch
cudaError_t sort_pairs_u64(cudaStream_t st, DevBuf& tmp, const uint64_t* keys_in, uint64_t* keys_out, const uint32_t* vals_in,
                           uint32_t* vals_out, size_t n, int bits);

// train.cu: GPU forest training (learning.cpp:410-916, 963-1012, 1031-1073)
rss_status forest_train(rss_ctx* ctx, const float* feats, int n, int D, const int32_t* labels, int L, const int* class_counts,
                        const rss_train_params& prm, const char* out_path, rss_train_stats* stats);

}  // namespace rss
