// TEST INFRASTRUCTURE ONLY - never linked into or called from the product path.
//
// C wrappers around the UNMODIFIED reference sources, compiled from where they lie under
// /root/reference by oracle/Makefile into oracle/_ref/libref_oracle.so:
//   * third-party/libforest/src/{classifier,data,learning,tools}.cpp  (learner = model producer,
//     RandomForest::read / multiClassLogPosterior / DecisionTree::findLeafNode = a9-a12 oracle)
//   * third-party/densecrf/src/permutohedral.cpp (Permutohedral::init / compute = a16-a18 oracle)
// No reference source is copied into this repository; this file only *calls* the reference.
#include <cstdio>
#include <cstring>
#include <fstream>
#include <vector>

#include "libforest/libforest.h"
#include "permutohedral.h"
#include "densecrf.h"

extern "C" {

// ---------------------------------------------------------------------------------------------
// libforest: training (mirrors the learner set-up of reference src/train.cpp:225-249)
// ---------------------------------------------------------------------------------------------
// feats: [n][D] row-major; labels: [n][L].  Returns 0 on success.
int ref_forest_train(const float* feats, int n, int D, const int* labels, int L, int num_trees,
                     int max_depth, int min_split, int num_threads, const char* out_path) {
    libf::DataStorage storage(L);
    for (int i = 0; i < n; i++) {
        libf::DataPoint* p = new libf::DataPoint(D);
        for (int k = 0; k < D; k++) p->at(k) = feats[(size_t)i * D + k];
        std::vector<int> lab(labels + (size_t)i * L, labels + (size_t)(i + 1) * L);
        storage.addDataPointMulti(p, lab);
    }
    libf::DecisionTreeLearner treeLearner;
    treeLearner.autoconf(&storage);
    treeLearner.setUseBootstrap(true);
    treeLearner.setMaxDepth(max_depth);
    treeLearner.setMinSplitExamples(min_split);
    treeLearner.setUseClassFrequency(false);
    treeLearner.useMultiLabelLayers(true);
    libf::RandomForestLearner forestLearner;
    forestLearner.setTreeLearner(&treeLearner);
    forestLearner.setNumTrees(num_trees);
    forestLearner.setNumThreads(num_threads);
    libf::RandomForest* forest = forestLearner.learn(&storage);
    std::filebuf fb;
    if (!fb.open(out_path, std::ios::out | std::ios::binary)) return 1;
    std::ostream os(&fb);
    forest->write(os);
    fb.close();
    delete forest;
    return 0;
}

// The same learner with its options exposed (src/train.cpp:225-249 sets bootstrap on, class frequency off, multi-label
// layers on).  With use_bootstrap = 0, num_features = D and one label layer nothing random is left in
// DecisionTreeLearner::learn, so the tree is a deterministic function of the data: the GPU learner is compared with it.
int ref_forest_train_opts(const float* feats, int n, int D, const int* labels, int L, int num_trees, int max_depth,
                          int min_split, int min_child_split, int num_features, int use_bootstrap, float smoothing,
                          int num_threads, const char* out_path) {
    libf::DataStorage storage(L);
    for (int i = 0; i < n; i++) {
        libf::DataPoint* p = new libf::DataPoint(D);
        for (int k = 0; k < D; k++) p->at(k) = feats[(size_t)i * D + k];
        std::vector<int> lab(labels + (size_t)i * L, labels + (size_t)(i + 1) * L);
        storage.addDataPointMulti(p, lab);
    }
    libf::DecisionTreeLearner treeLearner;
    treeLearner.autoconf(&storage);
    treeLearner.setUseBootstrap(use_bootstrap != 0);
    if (num_features > 0) treeLearner.setNumFeatures(num_features);
    treeLearner.setMaxDepth(max_depth);
    treeLearner.setMinSplitExamples(min_split);
    treeLearner.setMinChildSplitExamples(min_child_split);
    treeLearner.setSmoothingParameter(smoothing);
    treeLearner.setUseClassFrequency(false);
    treeLearner.useMultiLabelLayers(true);
    libf::RandomForestLearner forestLearner;
    forestLearner.setTreeLearner(&treeLearner);
    forestLearner.setNumTrees(num_trees);
    forestLearner.setNumThreads(num_threads);
    libf::RandomForest* forest = forestLearner.learn(&storage);
    std::filebuf fb;
    if (!fb.open(out_path, std::ios::out | std::ios::binary)) return 1;
    std::ostream os(&fb);
    forest->write(os);
    fb.close();
    delete forest;
    return 0;
}
// DecisionTreeLearner::updateMultiHistograms (learning.cpp:963-1012) applied to every tree of an existing forest file:
// the leaf histograms the reference computes for THAT tree structure from the given training set.
int ref_forest_update_histograms(const char* in_path, const float* feats, int n, int D, const int* labels, int L,
                                 float smoothing, const char* out_path) {
    std::filebuf fi;
    if (!fi.open(in_path, std::ios::in | std::ios::binary)) return 1;
    std::istream is(&fi);
    libf::RandomForest forest;
    forest.read(is);
    fi.close();
    libf::DataStorage storage(L);
    for (int i = 0; i < n; i++) {
        libf::DataPoint* p = new libf::DataPoint(D);
        for (int k = 0; k < D; k++) p->at(k) = feats[(size_t)i * D + k];
        std::vector<int> lab(labels + (size_t)i * L, labels + (size_t)(i + 1) * L);
        storage.addDataPointMulti(p, lab);
    }
    libf::DecisionTreeLearner treeLearner;
    treeLearner.setSmoothingParameter(smoothing);
    for (int t = 0; t < forest.getSize(); t++) treeLearner.updateMultiHistograms(dynamic_cast<libf::DecisionTree*>(forest.getTree(t)), &storage);
    std::filebuf fb;
    if (!fb.open(out_path, std::ios::out | std::ios::binary)) return 2;
    std::ostream os(&fb);
    forest.write(os);
    fb.close();
    return 0;
}

void* ref_forest_load(const char* path) {
    std::filebuf fb;
    if (!fb.open(path, std::ios::in | std::ios::binary)) return nullptr;
    std::istream is(&fb);
    libf::RandomForest* f = new libf::RandomForest();
    f->read(is);
    fb.close();
    return f;
}
void ref_forest_free(void* f) { delete (libf::RandomForest*)f; }
int ref_forest_num_trees(void* f) { return ((libf::RandomForest*)f)->getSize(); }

// leaf_ids: [T][n] (may be NULL), logpost: [n][sumC] with the layers concatenated (may be NULL).
// Returns sumC (floats per sample).
int ref_forest_predict(void* fv, const float* feats, int n, int D, int* leaf_ids, float* logpost) {
    libf::RandomForest* f = (libf::RandomForest*)fv;
    const int T = f->getSize();
    int sumC = 0;
    for (int i = 0; i < n; i++) {
        libf::DataPoint p(const_cast<float*>(feats + (size_t)i * D), D, false);
        if (leaf_ids)
            for (int t = 0; t < T; t++)
                leaf_ids[(size_t)t * n + i] = ((libf::DecisionTree*)f->getTree(t))->findLeafNode(&p);
        std::vector<std::vector<float> > post;
        f->multiClassLogPosterior(&p, post);
        if (i == 0)
            for (size_t l = 0; l < post.size(); l++) sumC += (int)post[l].size();
        if (logpost) {
            float* o = logpost + (size_t)i * sumC;
            for (size_t l = 0; l < post.size(); l++)
                for (size_t c = 0; c < post[l].size(); c++) *o++ = post[l][c];
        }
    }
    return sumC;
}

// ---------------------------------------------------------------------------------------------
// permutohedral lattice, verbatim
// ---------------------------------------------------------------------------------------------
struct RefLattice : public Permutohedral {
    int V() const { return M_; }
    const int* offsets() const { return offset_.data(); }
    const float* bary() const { return barycentric_.data(); }
    const Neighbors* nbr() const { return blur_neighbors_.data(); }
};

// feats: d x N column-major (Eigen layout, reference segmenter.cpp:629-637)
void* ref_lattice_init(const float* feats, int d, int N) {
    Eigen::MatrixXf f(d, N);
    memcpy(f.data(), feats, sizeof(float) * (size_t)d * N);
    RefLattice* l = new RefLattice();
    l->init(f);
    return l;
}
int ref_lattice_vertices(void* l) { return ((RefLattice*)l)->V(); }
// in/out: M x N column-major.  Dispatches like Permutohedral::compute (scalar path for M<=2).
void ref_lattice_compute(void* l, const float* in, int M, int N, float* out) {
    Eigen::MatrixXf i(M, N), o(M, N);
    memcpy(i.data(), in, sizeof(float) * (size_t)M * N);
    ((RefLattice*)l)->compute(o, i, false);
    memcpy(out, o.data(), sizeof(float) * (size_t)M * N);
}
// offsets/bary: [(d+1)*N] in the reference's point-major order (permutohedral.cpp:268-275)
void ref_lattice_get(void* l, int d, int N, int* offsets, float* bary) {
    memcpy(offsets, ((RefLattice*)l)->offsets(), sizeof(int) * (size_t)(d + 1) * N);
    memcpy(bary, ((RefLattice*)l)->bary(), sizeof(float) * (size_t)(d + 1) * N);
}
// neighbours: [(d+1)][V][2]
void ref_lattice_neighbors(void* l, int d, int* nb) {
    RefLattice* L = (RefLattice*)l;
    const int V = L->V();
    for (int j = 0; j <= d; j++)
        for (int i = 0; i < V; i++) {
            nb[((size_t)j * V + i) * 2 + 0] = L->nbr()[(size_t)j * V + i].n1;
            nb[((size_t)j * V + i) * 2 + 1] = L->nbr()[(size_t)j * V + i].n2;
        }
}
void ref_lattice_free(void* l) { delete (RefLattice*)l; }

// ---------------------------------------------------------------------------------------------
// DenseCRF glue, verbatim: densecrf.cpp (inference, expAndNormalize, currentMap, DenseCRF2D feature builders),
// pairwise.cpp (DenseKernel::initLattice / filter, all four NormalizationTypes), labelcompatibility.cpp (Potts),
// unary.cpp - compiled against the Eigen stand-in of oracle/shim.
// ---------------------------------------------------------------------------------------------
// unary, Q: M x N column-major (= the (N, M) row-major arrays of the Python side); feats[k]: d[k] x N column-major.
// The call sequence is the map worker's (src/segmenter.cpp:639-643): setUnaryEnergy, addPairwiseEnergy(feature,
// new PottsCompatibility(w)) with the default DIAG_KERNEL, then inference(iters); map (optional) = currentMap(Q).
void ref_crf_inference(int N, int M, const float* unary, const float* const* feats, const int* d, const float* w, int K,
                       int norm_type, int iters, float* Q, short* map) {
    DenseCRF crf(N, M);
    Eigen::MatrixXf U(M, N);
    memcpy(U.data(), unary, sizeof(float) * (size_t)M * N);
    crf.setUnaryEnergy(U);
    for (int k = 0; k < K; k++) {
        Eigen::MatrixXf f(d[k], N);
        memcpy(f.data(), feats[k], sizeof(float) * (size_t)d[k] * N);
        crf.addPairwiseEnergy(f, new PottsCompatibility(w[k]), DIAG_KERNEL, (NormalizationType)norm_type);
    }
    Eigen::MatrixXf R = crf.inference(iters);
    memcpy(Q, R.data(), sizeof(float) * (size_t)M * N);
    if (map) {
        VectorXs m = crf.currentMap(R);
        for (int i = 0; i < N; i++) map[i] = m[i];
    }
}
// DenseCRF::gradient (densecrf.cpp:238-297) with the reference's LogLikelihood objective (objective.cpp:36-52, compiled in
// place): objective value and the gradient w.r.t. the label-compatibility parameters (one Potts weight per pairwise term).
double ref_crf_gradient(int N, int M, const float* unary, const float* const* feats, const int* d, const float* w, int K,
                        int norm_type, int iters, const short* gt, float robust, float* potts_grad) {
    DenseCRF crf(N, M);
    Eigen::MatrixXf U(M, N);
    memcpy(U.data(), unary, sizeof(float) * (size_t)M * N);
    crf.setUnaryEnergy(U);
    for (int k = 0; k < K; k++) {
        Eigen::MatrixXf f(d[k], N);
        memcpy(f.data(), feats[k], sizeof(float) * (size_t)d[k] * N);
        crf.addPairwiseEnergy(f, new PottsCompatibility(w[k]), DIAG_KERNEL, (NormalizationType)norm_type);
    }
    VectorXs gtv(N);
    for (int i = 0; i < N; i++) gtv[i] = gt[i];
    LogLikelihood obj(gtv, robust);
    VectorXf g;
    const double r = crf.gradient(iters, obj, nullptr, &g, nullptr);
    for (int k = 0; k < K; k++) potts_grad[k] = g[k];
    return r;
}
// examples/dense_inference.cpp's model: DenseCRF2D with addPairwiseGaussian + addPairwiseBilateral (densecrf.cpp:61-81)
void ref_crf2d_inference(int W, int H, int M, const float* unary, float gsx, float gsy, float gw, float bsx, float bsy,
                         float bsr, float bsg, float bsb, const unsigned char* im, float bw, int iters, float* Q, short* map) {
    DenseCRF2D crf(W, H, M);
    Eigen::MatrixXf U(M, W * H);
    memcpy(U.data(), unary, sizeof(float) * (size_t)M * W * H);
    crf.setUnaryEnergy(U);
    crf.addPairwiseGaussian(gsx, gsy, new PottsCompatibility(gw));
    crf.addPairwiseBilateral(bsx, bsy, bsr, bsg, bsb, im, new PottsCompatibility(bw));
    Eigen::MatrixXf R = crf.inference(iters);
    memcpy(Q, R.data(), sizeof(float) * (size_t)M * W * H);
    if (map) {
        VectorXs m = crf.currentMap(R);
        for (int i = 0; i < W * H; i++) map[i] = m[i];
    }
}
// startInference / stepInference (densecrf.cpp:178-199)
void ref_crf_step_inference(int N, int M, const float* unary, const float* feats, int d, float w, int steps, float* Q) {
    DenseCRF crf(N, M);
    Eigen::MatrixXf U(M, N), f(d, N);
    memcpy(U.data(), unary, sizeof(float) * (size_t)M * N);
    memcpy(f.data(), feats, sizeof(float) * (size_t)d * N);
    crf.setUnaryEnergy(U);
    crf.addPairwiseEnergy(f, new PottsCompatibility(w));
    Eigen::MatrixXf R = crf.startInference(), t1, t2;
    for (int s = 0; s < steps; s++) crf.stepInference(R, t1, t2);
    memcpy(Q, R.data(), sizeof(float) * (size_t)M * N);
}

}  // extern "C"
