/* TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C CPU restatement of the reference's per-keyframe inference path
 * (RovinaSemanticSegmentation: feature_extractor.h -> libforest -> segmenter glue -> densecrf).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; the product (librss.so) never links, loads or calls it.
 *
 * Parity pinning (see DESIGN.md "Oracle"):
 *   - forest read/traverse/posterior and the permutohedral lattice are validated against the
 *     UNMODIFIED reference sources compiled into oracle/_ref/libref_oracle.so.
 *   - the OpenCV pieces (BGR2Lab 8U, copyMakeBorder REFLECT, resize INTER_LINEAR 8UC3 / 32FC)
 *     are validated against cv2 4.13.0 (fixtures in tests/golden/, generator committed).
 *   - PCL IntegralImageNormalEstimation and the Eigen 3x3 product order are restated from the
 *     published algorithm: "parity unpinned" for feature 365 (normal angle) and the cloud.
 */
#ifndef ORACLE_H
#define ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* number of OpenMP threads the oracle may use (1 = how the reference actually runs inference) */
void orc_set_threads(int n);
int orc_get_threads(void);

/* ---- OpenCV restatements ------------------------------------------------------------------ */
void orc_bgr2lab_u8(const uint8_t* src, int64_t npix, uint8_t* dst);
void orc_border_reflect_u8c3(const uint8_t* src, int W, int H, int b, uint8_t* dst);
/* src: S x S window inside an image with `sstep` bytes per row; dst: r x r x 3 */
void orc_resize_linear_u8c3(const uint8_t* src, int sstep, int S, uint8_t* dst, int r);
/* generic cv::resize(INTER_LINEAR) on 32FC(C) */
void orc_resize_linear_f32(const float* src, int sw, int sh, int C, float* dst, int dw, int dh);

/* ---- feature extractor (include/feature_extractor.h:29-291) -------------------------------- */
typedef struct {
    int use_color_patch, use_depth, use_height, use_normal; /* feature_extractor.h:30-33 */
    int patch_size, patch_size_reduce;                      /* :35-36 */
} orc_fe_config;
enum { ORC_WITH_ANY_LABEL = 0, ORC_WITH_POSITIVE_LABEL = 1, ORC_NO_LABEL = 2 };
int orc_feature_length(const orc_fe_config* cfg);
/* rgb: H*W*3 u8, depth: H*W u16 (mm); Kinv,R row-major 3x3; labels: L planes of H*W int8 or NULL.
 * feats: [cap][D] (row-major), xs/ys: [cap], out_labels: [cap][L] or NULL.  Returns the sample count. */
int orc_extract(const orc_fe_config* cfg, int stride, const uint8_t* rgb, const uint16_t* depth, int W, int H,
                const float* Kinv, const float* R, const float* t, float dmin, float dmax, int extract_type,
                const int8_t* labels, int L, float* feats, int* xs, int* ys, int* out_labels);
void orc_cloud(const uint16_t* depth, int W, int H, const float* Kinv, const float* R, const float* t, float dmin,
               float dmax, float* xyz);
/* PCL IntegralImageNormalEstimation<AVERAGE_3D_GRADIENT>, MaxDepthChangeFactor 0.02, NormalSmoothingSize 10.
 * normals: H*W*3, NaN where PCL yields NaN.  dist (optional, may be NULL): the chamfer distance map. */
void orc_normals(const float* xyz, int W, int H, float* normals, float* dist);

/* ---- libforest (third-party/libforest/src/classifier.cpp, include/libforest/io.h) ----------- */
typedef struct orc_forest orc_forest;
orc_forest* orc_forest_read(const char* path);
void orc_forest_free(orc_forest*);
int orc_forest_trees(const orc_forest*);
int orc_forest_nodes(const orc_forest*, int tree);
int orc_forest_layers(const orc_forest*);
int orc_forest_classes(const orc_forest*, int layer);
/* leaf_ids: [T][n] or NULL; logpost: [n][sumC] or NULL */
void orc_forest_predict(const orc_forest*, const float* feats, int n, int D, int* leaf_ids, float* logpost);

/* ---- segmenter frame worker body (src/segmenter.cpp:349-434) ------------------------------- */
/* posteriors: [layer][y][x][class]; fill = value of unsampled low-res pixels (0 node, -1000 test tool) */
void orc_segment_frame(const orc_fe_config* cfg, const orc_forest* f, int stride, const uint8_t* rgb,
                       const uint16_t* depth, int W, int H, const float* Kinv, const float* R, const float* t,
                       float dmin, float dmax, float fill, float* posteriors);

/* ---- permutohedral lattice + mean field (third-party/densecrf/src) ------------------------- */
typedef struct orc_lattice orc_lattice;
orc_lattice* orc_lattice_init(const float* feats /* d x N col-major */, int d, int N);
void orc_lattice_free(orc_lattice*);
int orc_lattice_vertices(const orc_lattice*);
void orc_lattice_get(const orc_lattice*, int* offsets, float* bary);
void orc_lattice_compute(const orc_lattice*, const float* in /* M x N */, int M, float* out);

typedef struct {
    const float* feats; /* d x N col-major */
    int d;
    float potts_w;
} orc_pairwise;
/* unary: M x N col-major energies; Q out: M x N.  NORMALIZE_SYMMETRIC, Potts. (densecrf.cpp:115-131) */
void orc_crf_inference(int N, int M, const float* unary, const orc_pairwise* kernels, int K, int iters, float* Q);
/* the same with an explicit NormalizationType (pairwise.h: 0 NO, 1 BEFORE, 2 AFTER, 3 SYMMETRIC) */
void orc_crf_inference_ex(int N, int M, const float* unary, const orc_pairwise* kernels, int K, int iters, int norm_type,
                          float* Q);
/* segmenter.cpp:597-616: unary(c,idx) += posterior[px*C + c] for idx>=0 */
void orc_project_zbuffer(const float* xyz, int N, const float* K, const float* R, const float* t, int W, int H, float zmin,
                         float zmax, int* index_image);
void orc_unary_accumulate(const int* index_image, int npix, const float* posterior, int C, float* unary);
/* segmenter.cpp:645-657 */
void orc_gated_argmax(const float* Q, int M, int N, int unknown_label, uint8_t* labels);
/* densecrf.cpp:61-81 (DenseCRF2D::addPairwiseGaussian / addPairwiseBilateral feature builders) */
void orc_features_gaussian2d(int W, int H, float sx, float sy, float* feats /* 2 x N */);
void orc_features_bilateral2d(int W, int H, float sx, float sy, float sr, float sg, float sb, const uint8_t* im,
                              float* feats /* 5 x N */);
/* segmenter.cpp:629-637 */
void orc_features_xyzrgb(int N, const float* xyz, const float* rgb, float wxyz, float wrgb, float* feats /* 6 x N */);

#ifdef __cplusplus
}
#endif
#endif
