// C++ host adapters over the C ABI of librss.so (include/rss.h).
//
// They keep the reference's class names, method names, argument meaning and ownership rules for the per-keyframe
// path, so that a caller written against the reference (src/segmenter.cpp, src/test_multi.cpp,
// third-party/densecrf/examples/dense_inference.cpp) compiles against this header with its includes swapped:
//
//   Features::FeatureExtractor   include/feature_extractor.h:24-41
//   libf::DataPoint / DataStorage third-party/libforest/include/libforest/data.h:27-126,210-460 (read side only)
//   libf::RandomForest           third-party/libforest/include/libforest/classifiers.h:274-344
//   DenseCRF / DenseCRF2D        third-party/densecrf/include/densecrf.h:36-121
//   PottsCompatibility           third-party/densecrf/include/labelcompatibility.h:50-61
//
// What differs, and why:
//   * OpenCV / Eigen / PCL are not dependencies.  Images are passed as rss::Image views (pointer + rows + cols; for a
//     cv::Mat m that is {m.data, m.rows, m.cols}), matrices as the column-major rss::MatrixXf below (same storage
//     order as Eigen::MatrixXf, so `Eigen::Map<Eigen::MatrixXf>(m.data(), m.rows(), m.cols())` is a zero-copy view).
//   * libf::DataStorage is ONE dense [n][D] float array (the reference allocates one heap float[D] per sample,
//     data.h:39); DataPoint is a view into it.  The forest consumes the device-resident copy when the storage was
//     filled by FeatureExtractor::extract on the same session, so features never travel back to the GPU.
//   * All work happens on the GPU behind rss::Session (one per GPU per host thread).  Errors become rss::Error
//     (the reference asserts or segfaults); there is no CPU fallback.
// Header-only; link with -lrss.
#pragma once
#include <cstdint>
#include <cstring>
#include <istream>
#include <iterator>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/rss.h"

namespace rss {

struct Error : std::runtime_error {
    rss_status status;
    Error(rss_status s, const std::string& m) : std::runtime_error("rss status " + std::to_string(s) + ": " + m), status(s) {}
};

// One GPU context = the Segmenter constructor's state (config + forest), src/segmenter.cpp:70-127.
class Session {
   public:
    Session(const std::string& config_json, const std::string& forest_dat = std::string(), int cuda_device = 0) {
        const rss_status st = rss_create(config_json.c_str(), forest_dat.empty() ? nullptr : forest_dat.c_str(), cuda_device, &ctx_);
        if (st != RSS_OK) throw Error(st, rss_last_error(nullptr));
        check(rss_get_info(ctx_, &info_));
    }
    ~Session() { if (ctx_) rss_destroy(ctx_); }
    Session(const Session&) = delete;
    Session& operator=(const Session&) = delete;
    rss_ctx* handle() const { return ctx_; }
    const rss_info& info() { check(rss_get_info(ctx_, &info_)); return info_; }
    void check(rss_status st) const { if (st != RSS_OK) throw Error(st, rss_last_error(ctx_)); }

   private:
    rss_ctx* ctx_ = nullptr;
    rss_info info_{};
};

// pointer + shape views of caller-owned images (cv::Mat stand-ins; rows are contiguous)
template <class T, int CH>
struct Image {
    T* data = nullptr;
    int rows = 0, cols = 0;
    Image() = default;
    Image(T* d, int r, int c) : data(d), rows(r), cols(c) {}
};
using Image8UC3 = Image<const uint8_t, 3>;   // CV_8UC3
using Image16UC1 = Image<const uint16_t, 1>; // CV_16UC1, millimetres
using Image8SC1 = Image<const int8_t, 1>;    // label_type = char (feature_extractor.h:23)

// column-major float matrix, the storage order of Eigen::MatrixXf
class MatrixXf {
   public:
    MatrixXf() = default;
    MatrixXf(int rows, int cols) : r_(rows), c_(cols), v_((size_t)rows * cols, 0.f) {}
    int rows() const { return r_; }
    int cols() const { return c_; }
    float& operator()(int i, int j) { return v_[(size_t)j * r_ + i]; }
    const float& operator()(int i, int j) const { return v_[(size_t)j * r_ + i]; }
    float* data() { return v_.data(); }
    const float* data() const { return v_.data(); }
    void resize(int rows, int cols) { r_ = rows; c_ = cols; v_.assign((size_t)rows * cols, 0.f); }

   private:
    int r_ = 0, c_ = 0;
    std::vector<float> v_;
};
using VectorXs = std::vector<short>;

struct Calibration {  // include/calibration.h:20-22, row-major 3x3 + translation
    float intrinsic_inverse[9];
    float extrinsic_linear[9];
    float extrinsic_translation[3];
};

}  // namespace rss

// ---------------------------------------------------------------------------------------------------------------------
// libforest read side
// ---------------------------------------------------------------------------------------------------------------------
namespace libf {

class DataStorage;

// A view of one row of a DataStorage (data.h:27-103 keeps an owning float* per point).
class DataPoint {
   public:
    DataPoint() = default;
    DataPoint(const float* row, int D) : row_(row), D_(D) {}
    const float& at(int i) const { return row_[i]; }
    int getDimensionality() const { return D_; }
    const float* data() const { return row_; }

   private:
    const float* row_ = nullptr;
    int D_ = 0;
};

// Dense sample container (data.h:210-460, read side): samples in insertion order, optional per-layer labels.
class DataStorage {
   public:
    int getSize() const { return n_; }
    int getDimensionality() const { return D_; }
    DataPoint getDataPoint(int i) const { return DataPoint(feats_.data() + (size_t)i * D_, D_); }
    int getClassLabelsMulti(int i, int layer) const { return labels_[(size_t)i * layers_ + layer]; }
    std::vector<int> getClassLabelsMulti(int i) const {
        return std::vector<int>(labels_.begin() + (size_t)i * layers_, labels_.begin() + (size_t)(i + 1) * layers_);
    }
    void clear() { n_ = 0; feats_.clear(); labels_.clear(); resident_on_ = nullptr; }
    const float* data() const { return feats_.data(); }
    // append n rows (used by FeatureExtractor::extract); returns the write position
    float* append(int n, int D, int label_layers) {
        if (n_ > 0 && D != D_) throw std::invalid_argument("DataStorage: dimensionality mismatch");
        D_ = D; layers_ = label_layers;
        feats_.resize((size_t)(n_ + n) * D);
        labels_.resize((size_t)(n_ + n) * layers_);
        float* p = feats_.data() + (size_t)n_ * D;
        n_ += n;
        return p;
    }
    int* label_rows(int first) { return labels_.data() + (size_t)first * layers_; }
    // the session whose device memory currently holds exactly these rows (set by extract into an EMPTY storage)
    const rss::Session* resident_on_ = nullptr;

   private:
    int n_ = 0, D_ = 0, layers_ = 0;
    std::vector<float> feats_;
    std::vector<int> labels_;
};

// RandomForest, prediction side (classifiers.h:274-344).  The model lives in the session (SoA nodes + dense leaf table).
class RandomForest {
   public:
    explicit RandomForest(rss::Session& s) : s_(s) {}
    // RandomForest::read (classifier.cpp:222-235): the whole libforest binary stream
    void read(std::istream& stream) {
        std::vector<char> bytes((std::istreambuf_iterator<char>(stream)), std::istreambuf_iterator<char>());
        s_.check(rss_load_forest_memory(s_.handle(), bytes.data(), bytes.size()));
    }
    int getSize() { return s_.info().num_trees; }
    // classifier.cpp:187-208 for one sample: post[layer][class] = sum over trees of the leaf's log-histogram
    void multiClassLogPosterior(const DataPoint* x, std::vector<std::vector<float>>& probabilities) {
        const rss_info& i = s_.info();
        std::vector<float> flat(i.total_classes);
        s_.check(rss_forest_predict(s_.handle(), x->data(), 1, nullptr, flat.data()));
        split(i, flat.data(), probabilities);
    }
    // classifier.cpp:166-184 (single-label forests have one layer)
    void classLogPosterior(const DataPoint* x, std::vector<float>& probabilities) {
        const rss_info& i = s_.info();
        probabilities.assign(i.total_classes, 0.f);
        s_.check(rss_forest_predict(s_.handle(), x->data(), 1, nullptr, probabilities.data()));
    }
    // The batched form the GPU wants: every sample of the storage in one call.  log_post: [n][sumC] (layers
    // concatenated per sample); leaf_ids (optional): [T][n] = DecisionTree::findLeafNode per tree.
    void multiClassLogPosterior(const DataStorage& storage, std::vector<float>& log_post, std::vector<int32_t>* leaf_ids = nullptr) {
        const rss_info& i = s_.info();
        const int n = storage.getSize();
        log_post.resize((size_t)n * i.total_classes);
        if (leaf_ids) leaf_ids->resize((size_t)n * i.num_trees);
        if (n == 0) return;
        const float* feats = storage.resident_on_ == &s_ ? nullptr : storage.data();
        s_.check(rss_forest_predict(s_.handle(), feats, n, leaf_ids ? leaf_ids->data() : nullptr, log_post.data()));
    }

   private:
    static void split(const rss_info& i, const float* flat, std::vector<std::vector<float>>& out) {
        out.resize(i.layer_count);
        int off = 0;
        for (int l = 0; l < i.layer_count; l++) {
            out[l].assign(flat + off, flat + off + i.class_counts[l]);
            off += i.class_counts[l];
        }
    }
    rss::Session& s_;
};

}  // namespace libf

// ---------------------------------------------------------------------------------------------------------------------
// Features::FeatureExtractor (include/feature_extractor.h:24-291)
// ---------------------------------------------------------------------------------------------------------------------
enum class ExtractType { WITH_ANY_LABEL, WITH_POSITIVE_LABEL, NO_LABEL };  // feature_extractor.h:21

namespace Features {

class FeatureExtractor {
   public:
    // the reference constructs from Utils::Config; here the session already parsed the same keys (:30-36)
    explicit FeatureExtractor(rss::Session& s) : s_(s) {}

    // Appends the selected samples (raster order) to storage / x_v / y_v, like the reference (:41-291).
    void extract(int stride, const rss::Image8UC3& color, const rss::Image16UC1& depth_im, const rss::Calibration& c,
                 libf::DataStorage& storage, std::vector<int>& x_v, std::vector<int>& y_v, ExtractType label_extraction,
                 float d_min, float d_max, const std::vector<rss::Image8SC1>& label = std::vector<rss::Image8SC1>()) {
        const int W = depth_im.cols, H = depth_im.rows;
        if (color.rows != H || color.cols != W) throw std::invalid_argument("extract: colour and depth sizes differ");
        const int cap = ((W + stride - 1) / stride) * ((H + stride - 1) / stride);
        const int D = s_.info().feature_length;
        const int layers = label_extraction == ExtractType::NO_LABEL ? 0 : (int)label.size();
        std::vector<int8_t> planes;
        for (int l = 0; l < layers; l++) {
            if (label[l].rows != H || label[l].cols != W) throw std::invalid_argument("extract: label image size");
            planes.insert(planes.end(), label[l].data, label[l].data + (size_t)W * H);
        }
        const bool was_empty = storage.getSize() == 0;
        const int first = storage.getSize();
        float* rows = storage.append(cap, D, layers);
        const size_t x0 = x_v.size();
        x_v.resize(x0 + cap);
        y_v.resize(x0 + cap);
        int n = 0;
        const int et = label_extraction == ExtractType::WITH_ANY_LABEL ? RSS_WITH_ANY_LABEL
                       : label_extraction == ExtractType::WITH_POSITIVE_LABEL ? RSS_WITH_POSITIVE_LABEL : RSS_NO_LABEL;
        s_.check(rss_extract_features(s_.handle(), color.data, depth_im.data, W, H, stride, c.intrinsic_inverse,
                                      c.extrinsic_linear, c.extrinsic_translation, d_min, d_max, et,
                                      layers ? planes.data() : nullptr, layers, rows, x_v.data() + x0, y_v.data() + x0,
                                      layers ? storage.label_rows(first) : nullptr, &n));
        storage.append(n - cap, D, layers);  // shrink to the true sample count
        x_v.resize(x0 + n);
        y_v.resize(x0 + n);
        storage.resident_on_ = was_empty ? &s_ : nullptr;
    }

   private:
    rss::Session& s_;
};

}  // namespace Features

// ---------------------------------------------------------------------------------------------------------------------
// DenseCRF (third-party/densecrf/include/densecrf.h:36-121)
// ---------------------------------------------------------------------------------------------------------------------
enum KernelType { CONST_KERNEL, DIAG_KERNEL, FULL_KERNEL };  // pairwise.h; only DIAG_KERNEL is on the reference's path
enum NormalizationType { NO_NORMALIZATION, NORMALIZE_BEFORE, NORMALIZE_AFTER, NORMALIZE_SYMMETRIC };

class LabelCompatibility {
   public:
    virtual ~LabelCompatibility() {}
    virtual float pottsWeight() const = 0;
};
class PottsCompatibility : public LabelCompatibility {  // labelcompatibility.cpp:41-48: out = -w * Q
   public:
    explicit PottsCompatibility(float weight = 1.0f) : w_(weight) {}
    float pottsWeight() const override { return w_; }

   private:
    float w_;
};

class DenseCRF {
   public:
    DenseCRF(rss::Session& s, int N, int M) : s_(s), N_(N), M_(M) { s_.check(rss_crf_create(s_.handle(), N, M, &crf_)); }
    virtual ~DenseCRF() { if (crf_) rss_crf_destroy(crf_); }
    DenseCRF(const DenseCRF&) = delete;

    // setUnaryEnergy(const MatrixXf&) (densecrf.cpp:88-90): M x N energies
    void setUnaryEnergy(const rss::MatrixXf& unary) {
        if (unary.rows() != M_ || unary.cols() != N_) throw std::invalid_argument("setUnaryEnergy: expected M x N");
        s_.check(rss_crf_set_unary(crf_, 0, unary.data()));
    }
    // addPairwiseEnergy (densecrf.cpp:54-60); ownership of `function` is transferred like in the reference
    void addPairwiseEnergy(const rss::MatrixXf& features, LabelCompatibility* function, KernelType kernel_type = DIAG_KERNEL,
                           NormalizationType normalization_type = NORMALIZE_SYMMETRIC) {
        std::unique_ptr<LabelCompatibility> own(function);
        if (kernel_type != DIAG_KERNEL) throw std::invalid_argument("only DIAG_KERNEL is supported (the reference path uses no other)");
        if (features.cols() != N_) throw std::invalid_argument("addPairwiseEnergy: expected d x N features");
        s_.check(rss_crf_add_pairwise(crf_, features.data(), features.rows(), function->pottsWeight(), (int)normalization_type));
    }
    // inference (densecrf.cpp:115-131): M x N marginals
    rss::MatrixXf inference(int n_iterations) const {
        rss::MatrixXf Q(M_, N_);
        s_.check(rss_crf_inference(crf_, 0, n_iterations, Q.data(), nullptr, nullptr));
        return Q;
    }
    // map (densecrf.cpp:132-137): plain argmax per variable
    rss::VectorXs map(int n_iterations) const {
        std::vector<uint8_t> lab(N_);
        s_.check(rss_crf_inference(crf_, 0, n_iterations, nullptr, lab.data(), nullptr));
        return rss::VectorXs(lab.begin(), lab.end());
    }
    // gated argmax of src/segmenter.cpp:645-657 (label = argmax if Q > 2/M else `unknown`)
    std::vector<unsigned char> gatedMap(int n_iterations, int unknown) const {
        std::vector<unsigned char> lab(N_);
        s_.check(rss_crf_inference(crf_, 0, n_iterations, nullptr, lab.data(), &unknown));
        return lab;
    }
    // step-by-step inference (densecrf.cpp:178-211).  Q lives on the device; the tmp arguments of the reference
    // (scratch matrices) are not needed.
    rss::MatrixXf startInference() const {
        s_.check(rss_crf_start_inference(crf_));
        rss::MatrixXf Q(M_, N_);
        s_.check(rss_crf_current(crf_, 0, Q.data(), nullptr, nullptr));
        return Q;
    }
    void stepInference(rss::MatrixXf& Q) const {
        s_.check(rss_crf_step_inference(crf_, 1));
        s_.check(rss_crf_current(crf_, 0, Q.data(), nullptr, nullptr));
    }
    rss::VectorXs currentMap() const {
        std::vector<uint8_t> lab(N_);
        s_.check(rss_crf_current(crf_, 0, nullptr, lab.data(), nullptr));
        return rss::VectorXs(lab.begin(), lab.end());
    }
    rss_crf* handle() const { return crf_; }

   protected:
    rss::Session& s_;
    int N_, M_;
    rss_crf* crf_ = nullptr;
};

class DenseCRF2D : public DenseCRF {
   public:
    DenseCRF2D(rss::Session& s, int W, int H, int M) : DenseCRF(s, W * H, M), W_(W), H_(H) {}
    // densecrf.cpp:61-69
    void addPairwiseGaussian(float sx, float sy, LabelCompatibility* function = nullptr) {
        std::unique_ptr<LabelCompatibility> own(function);
        s_.check(rss_crf_add_pairwise_gaussian(crf_, W_, H_, sx, sy, function ? function->pottsWeight() : 1.0f));
    }
    // densecrf.cpp:70-81
    void addPairwiseBilateral(float sx, float sy, float sr, float sg, float sb, const unsigned char* im,
                              LabelCompatibility* function = nullptr) {
        std::unique_ptr<LabelCompatibility> own(function);
        s_.check(rss_crf_add_pairwise_bilateral(crf_, W_, H_, sx, sy, sr, sg, sb, im, function ? function->pottsWeight() : 1.0f));
    }

   protected:
    int W_, H_;
};

// ---------------------------------------------------------------------------------------------------------------------
// semantic_segmentation::SingleFrameSegmentation (srv/SingleFrameSegmentation.srv)
// ---------------------------------------------------------------------------------------------------------------------
namespace semantic_segmentation {
// srv/SingleFrameSegmentation.srv with plain structs in place of sensor_msgs/Image (same fields the node fills at
// src/segmenter.cpp:490-497: height, width, encoding, step, data).  In a ROS build, `call` is the body of the callback
// registered with advertiseService("/semantic_segmentation/SingleFrameSegmentation", ...): INTEGRATION.md.
struct ImageMsg {
    uint32_t height = 0, width = 0;
    std::string encoding;  // "rgb8" for Request::rgb, "32FC3" for Request::depth
    uint32_t step = 0;     // bytes per row
    std::vector<uint8_t> data;
};
struct SingleFrameSegmentationRequest { ImageMsg rgb, depth; };
struct SingleFrameSegmentationResponse { std::vector<float> label_distribution; };
class SingleFrameSegmentationService {
   public:
    SingleFrameSegmentationService(rss::Session& s, const float Kinv[9], const float R[9], const float t[3]) : s_(s) {
        std::memcpy(Kinv_, Kinv, sizeof Kinv_); std::memcpy(R_, R, sizeof R_); std::memcpy(t_, t, sizeof t_);
    }
    // returns false like a ROS service callback that fails (the node then throws, src/segmenter.cpp:502-504)
    bool call(const SingleFrameSegmentationRequest& req, SingleFrameSegmentationResponse& resp) {
        const uint32_t W = req.rgb.width, H = req.rgb.height;
        if (W == 0 || H == 0 || req.depth.width != W || req.depth.height != H) return false;
        if (req.rgb.encoding != "rgb8" || req.depth.encoding != "32FC3") return false;
        if (req.rgb.step < 3 * W || req.depth.step < 12 * W) return false;
        if (req.rgb.data.size() < (size_t)req.rgb.step * H || req.depth.data.size() < (size_t)req.depth.step * H) return false;
        const uint8_t* rgb = req.rgb.data.data();
        const float* cloud = reinterpret_cast<const float*>(req.depth.data.data());
        std::vector<uint8_t> rgb_packed;
        std::vector<float> cloud_packed;
        if (req.rgb.step != 3 * W) {  // padded rows: pack
            rgb_packed.resize((size_t)3 * W * H);
            for (uint32_t y = 0; y < H; y++) std::memcpy(&rgb_packed[(size_t)3 * W * y], rgb + (size_t)req.rgb.step * y, 3 * W);
            rgb = rgb_packed.data();
        }
        if (req.depth.step != 12 * W) {
            cloud_packed.resize((size_t)3 * W * H);
            for (uint32_t y = 0; y < H; y++)
                std::memcpy(&cloud_packed[(size_t)3 * W * y], req.depth.data.data() + (size_t)req.depth.step * y, 12 * W);
            cloud = cloud_packed.data();
        }
        resp.label_distribution.resize((size_t)s_.info().total_classes * W * H);
        return rss_service_single_frame(s_.handle(), rgb, cloud, (int)W, (int)H, Kinv_, R_, t_,
                                        resp.label_distribution.data()) == RSS_OK;
    }

   private:
    rss::Session& s_;
    float Kinv_[9], R_[9], t_[3];
};

}  // namespace semantic_segmentation
