// Exercises the reference-named C++ adapters (rovinasemanticsegmentation_b200/host/rss_adapters.hpp) the way
// src/test_multi.cpp:150-199 and third-party/densecrf/examples/dense_inference.cpp:60-110 use the reference classes,
// and dumps every result as raw binary for tests/test_gpu_host_cpp.py to compare with the oracle.
//   adapters_check <config> <forest> <frame.raw> <calib.raw> <W> <H> <outdir>
#include <cstdio>
#include <fstream>
#include <limits>

#include "../../rovinasemanticsegmentation_b200/host/rss_adapters.hpp"

template <class T>
static void dump(const std::string& path, const T* p, size_t n) {
    std::ofstream(path, std::ios::binary).write((const char*)p, n * sizeof(T));
}

int main(int argc, char** argv) {
    if (argc != 8) { fprintf(stderr, "usage\n"); return 2; }
    const std::string config = argv[1], forest_path = argv[2], frame = argv[3], calib = argv[4], out = argv[7];
    const int W = atoi(argv[5]), H = atoi(argv[6]);
    try {
        std::vector<uint8_t> rgb((size_t)W * H * 3);
        std::vector<uint16_t> depth((size_t)W * H);
        std::ifstream in(frame, std::ios::binary);
        in.read((char*)rgb.data(), rgb.size());
        in.read((char*)depth.data(), depth.size() * 2);
        rss::Calibration cal;
        std::ifstream(calib, std::ios::binary).read((char*)&cal, sizeof(cal));

        rss::Session session(config);  // no model yet: RandomForest::read loads it, like segmenter.cpp:106-115
        libf::RandomForest forest(session);
        std::ifstream model(forest_path, std::ios::binary);
        forest.read(model);
        if (forest.getSize() < 1) return 1;

        // FeatureExtractor::extract -> RandomForest::multiClassLogPosterior (test_multi.cpp:160-175)
        Features::FeatureExtractor fe(session);
        libf::DataStorage storage;
        std::vector<int> x_v, y_v;
        fe.extract(2, rss::Image8UC3(rgb.data(), H, W), rss::Image16UC1(depth.data(), H, W), cal, storage, x_v, y_v,
                   ExtractType::NO_LABEL, 0.5f, 15.0f);
        const int n = storage.getSize(), D = storage.getDimensionality();
        dump(out + "/feats.bin", storage.data(), (size_t)n * D);
        dump(out + "/xs.bin", x_v.data(), x_v.size());
        dump(out + "/ys.bin", y_v.data(), y_v.size());
        std::vector<float> post;
        std::vector<int32_t> leaves;
        forest.multiClassLogPosterior(storage, post, &leaves);  // batched, device-resident features
        dump(out + "/post.bin", post.data(), post.size());
        dump(out + "/leaves.bin", leaves.data(), leaves.size());
        // the reference's per-sample call shape on a few samples
        std::vector<float> single;
        for (int i = 0; i < n; i += n / 7 + 1) {
            std::vector<std::vector<float>> p;
            libf::DataPoint x = storage.getDataPoint(i);
            forest.multiClassLogPosterior(&x, p);
            for (auto& layer : p) single.insert(single.end(), layer.begin(), layer.end());
        }
        dump(out + "/post_single.bin", single.data(), single.size());

        // DenseCRF2D as in dense_inference.cpp:86-101: unary from the first layer's posteriors of a labelled grid
        const int M = 5, N = W * H;
        rss::MatrixXf unary(M, N);
        for (int i = 0; i < N; i++)
            for (int c = 0; c < M; c++) unary(c, i) = ((i / 700) % M == c) ? 0.3f : 1.9f;
        DenseCRF2D crf(session, W, H, M);
        crf.setUnaryEnergy(unary);
        crf.addPairwiseGaussian(3, 3, new PottsCompatibility(3));
        crf.addPairwiseBilateral(80, 80, 13, 13, 13, rgb.data(), new PottsCompatibility(10));
        rss::MatrixXf Q = crf.inference(5);
        dump(out + "/Q.bin", Q.data(), (size_t)M * N);
        rss::VectorXs map = crf.map(5);
        dump(out + "/map.bin", map.data(), map.size());
        // step-wise inference must reproduce inference(5)
        rss::MatrixXf Qs = crf.startInference();
        for (int it = 0; it < 5; it++) crf.stepInference(Qs);
        dump(out + "/Qstep.bin", Qs.data(), (size_t)M * N);
        rss::VectorXs cm = crf.currentMap();
        dump(out + "/mapstep.bin", cm.data(), cm.size());

        // generic DenseCRF with a feature matrix (segmenter.cpp:629-643 shape: d x N, Potts w)
        rss::MatrixXf feats(3, N);
        for (int i = 0; i < N; i++) { feats(0, i) = (i % W) / 6.0f; feats(1, i) = (i / W) / 6.0f; feats(2, i) = rgb[3 * (size_t)i] / 20.0f; }
        DenseCRF crf2(session, N, M);
        crf2.setUnaryEnergy(unary);
        crf2.addPairwiseEnergy(feats, new PottsCompatibility(10.0f));
        std::vector<unsigned char> gated = crf2.gatedMap(10, M - 1);
        dump(out + "/gated.bin", gated.data(), gated.size());
        dump(out + "/feats3.bin", feats.data(), (size_t)3 * N);
        dump(out + "/unary.bin", unary.data(), (size_t)M * N);
        // semantic_segmentation::SingleFrameSegmentation: the request the node builds at src/segmenter.cpp:463-497
        // (rectified cloud = (R * Kinv) * (d x, d y, d) + t, NaN outside [0.5, 15] m), rows padded by 16 bytes
        {
            using namespace semantic_segmentation;
            float M3[9];
            for (int i = 0; i < 3; i++)
                for (int j = 0; j < 3; j++) {
                    float acc = 0.f;
                    for (int k = 0; k < 3; k++) acc += cal.extrinsic_linear[3 * i + k] * cal.intrinsic_inverse[3 * k + j];
                    M3[3 * i + j] = acc;
                }
            SingleFrameSegmentationRequest req;
            req.rgb.height = req.depth.height = H; req.rgb.width = req.depth.width = W;
            req.rgb.encoding = "rgb8"; req.depth.encoding = "32FC3";
            req.rgb.step = 3 * W + 16; req.depth.step = 12 * W + 16;
            req.rgb.data.assign((size_t)req.rgb.step * H, 0);
            req.depth.data.assign((size_t)req.depth.step * H, 0);
            for (int y = 0; y < H; y++) {
                memcpy(&req.rgb.data[(size_t)req.rgb.step * y], &rgb[(size_t)3 * W * y], 3 * W);
                float* row = reinterpret_cast<float*>(&req.depth.data[(size_t)req.depth.step * y]);
                for (int x = 0; x < W; x++) {
                    const float d = static_cast<float>(depth[(size_t)y * W + x]) / 1000.0f;
                    for (int i = 0; i < 3; i++) {
                        const float v[3] = {d * x, d * y, d};
                        row[3 * x + i] = (d < 0.5 || d > 15.0) ? std::numeric_limits<float>::quiet_NaN()
                                         : M3[3 * i] * v[0] + M3[3 * i + 1] * v[1] + M3[3 * i + 2] * v[2] + cal.extrinsic_translation[i];
                    }
                }
            }
            SingleFrameSegmentationService service(session, cal.intrinsic_inverse, cal.extrinsic_linear, cal.extrinsic_translation);
            SingleFrameSegmentationResponse resp;
            if (!service.call(req, resp)) throw std::runtime_error("SingleFrameSegmentation service call failed");
            dump(out + "/service.bin", resp.label_distribution.data(), resp.label_distribution.size());
            req.depth.encoding = "mono16";
            if (service.call(req, resp)) throw std::runtime_error("a wrong encoding must make the service call fail");
        }
        printf("adapters ok: %d samples, D=%d\n", n, D);
    } catch (const std::exception& e) {
        fprintf(stderr, "adapters_check: %s\n", e.what());
        return 1;
    }
    return 0;
}
