// Multi-GPU keyframe worker: the C++ host side of the per-keyframe path.
//
// The reference runs one worker thread per pipeline stage on one CPU (src/segmenter.cpp:227-232: frame worker
// processFramesFromQueueInternalRF, map worker processMapFromQueue).  Keyframes are independent units, so here the
// unit of parallelism is the keyframe: every GPU gets `inflight` workers, each with its own rss::Session (streams,
// device buffers, pinned staging) and all workers pull keyframe indices from one shared counter.  There is no
// collective and no peer traffic: inputs go host -> device over PCIe, label maps come back.
//
//   keyframe_worker --config resources/keyframe_config.json --forest tests/golden/forest_shared.dat
//                   [--gpus N] [--inflight C] [--frames K] [--width 640 --height 480]
//                   [--frames-file raw] [--calib raw21floats] [--out labels.bin] [--iters 10]
//
// --frames-file: K' frames stored back to back as rgb (H*W*3 u8) + depth (H*W u16); keyframe k uses frame k % K'.
// Without it a small set of synthetic frames is generated.  --out: label maps [frame][layer][H*W] u8 of the first
// min(K, K') keyframes.  Prints one JSON line with the measured keyframes/s (wall clock over all workers).
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>
#include <thread>

#include "rss_adapters.hpp"

namespace {

struct Args {
    std::string config, forest, frames_file, out, calib;
    int gpus = 1, inflight = 2, frames = 64, W = 640, H = 480, iters = 10;
};

struct HostFrames {  // pinned staging: all distinct frames
    int count = 0;
    size_t rgb_bytes = 0, depth_bytes = 0;
    uint8_t* rgb = nullptr;
    uint16_t* depth = nullptr;
    ~HostFrames() { rss_host_free(rgb); rss_host_free(depth); }
};

void synth_frames(HostFrames& f, int count, int W, int H) {
    f.count = count;
    f.rgb_bytes = (size_t)W * H * 3;
    f.depth_bytes = (size_t)W * H * 2;
    f.rgb = (uint8_t*)rss_host_alloc(f.rgb_bytes * count);
    f.depth = (uint16_t*)rss_host_alloc(f.depth_bytes * count);
    if (!f.rgb || !f.depth) throw std::runtime_error("pinned allocation failed");
    uint32_t lcg = 12345u;
    for (int k = 0; k < count; k++)
        for (int y = 0; y < H; y++)
            for (int x = 0; x < W; x++) {
                const size_t i = (size_t)k * W * H + (size_t)y * W + x;
                for (int c = 0; c < 3; c++) {
                    lcg = lcg * 1664525u + 1013904223u;
                    const double v = 128 + 70 * std::sin(0.01 * (c + 1) * x + k) * std::cos(0.013 * y + c) + ((lcg >> 24) & 15) - 8;
                    f.rgb[3 * i + c] = (uint8_t)std::fmin(255.0, std::fmax(0.0, v));
                }
                const double d = 1000.0 * (1 + 3 * (0.5 + 0.5 * std::sin(0.004 * x + k) * std::cos(0.006 * y)));
                lcg = lcg * 1664525u + 1013904223u;
                f.depth[i] = (lcg >> 20) % 500 == 0 ? 0 : (uint16_t)d;
            }
}

void load_frames(HostFrames& f, const std::string& path, int W, int H) {
    std::ifstream in(path, std::ios::binary | std::ios::ate);
    if (!in) throw std::runtime_error("cannot open " + path);
    const size_t total = (size_t)in.tellg();
    f.rgb_bytes = (size_t)W * H * 3;
    f.depth_bytes = (size_t)W * H * 2;
    f.count = (int)(total / (f.rgb_bytes + f.depth_bytes));
    if (f.count < 1) throw std::runtime_error("frames file too small");
    f.rgb = (uint8_t*)rss_host_alloc(f.rgb_bytes * f.count);
    f.depth = (uint16_t*)rss_host_alloc(f.depth_bytes * f.count);
    in.seekg(0);
    for (int k = 0; k < f.count; k++) {
        in.read((char*)f.rgb + f.rgb_bytes * k, f.rgb_bytes);
        in.read((char*)f.depth + f.depth_bytes * k, f.depth_bytes);
    }
}

}  // namespace

int main(int argc, char** argv) {
    Args a;
    for (int i = 1; i < argc; i++) {
        const std::string k = argv[i];
        auto val = [&]() -> std::string { if (i + 1 >= argc) { fprintf(stderr, "missing value for %s\n", k.c_str()); exit(2); } return argv[++i]; };
        if (k == "--config") a.config = val();
        else if (k == "--forest") a.forest = val();
        else if (k == "--frames-file") a.frames_file = val();
        else if (k == "--out") a.out = val();
        else if (k == "--calib") a.calib = val();
        else if (k == "--gpus") a.gpus = atoi(val().c_str());
        else if (k == "--inflight") a.inflight = atoi(val().c_str());
        else if (k == "--frames") a.frames = atoi(val().c_str());
        else if (k == "--width") a.W = atoi(val().c_str());
        else if (k == "--height") a.H = atoi(val().c_str());
        else if (k == "--iters") a.iters = atoi(val().c_str());
        else { fprintf(stderr, "unknown argument %s\n", k.c_str()); return 2; }
    }
    if (a.config.empty() || a.forest.empty()) { fprintf(stderr, "--config and --forest are required\n"); return 2; }
    try {
        HostFrames frames;
        if (a.frames_file.empty()) synth_frames(frames, 8, a.W, a.H);
        else load_frames(frames, a.frames_file, a.W, a.H);
        // calibration of the synthetic workload: fx = fy = 525 * W / 640, principal point at the centre, camera pitched 12 degrees
        const double fpx = 525.0 * a.W / 640.0, ang = 12.0 * M_PI / 180.0;
        rss::Calibration cal = {{(float)(1 / fpx), 0, (float)(-a.W / 2.0 / fpx), 0, (float)(1 / fpx), (float)(-a.H / 2.0 / fpx), 0, 0, 1},
                                {(float)-0.0, (float)-std::sin(ang), (float)std::cos(ang), -1, 0, 0, 0, (float)-std::cos(ang), (float)-std::sin(ang)},
                                {0, 0, 1.2f}};
        if (!a.calib.empty()) {  // 21 raw floats: Kinv (9), R (9), t (3)
            std::ifstream in(a.calib, std::ios::binary);
            if (!in.read((char*)&cal, sizeof(cal))) throw std::runtime_error("cannot read " + a.calib);
        }
        rss_keyframe_params prm = {0.05f, 3.0f, 80.0f, 13.0f, 10.0f, a.iters, 0.0f};
        const int workers = a.gpus * a.inflight;
        const size_t NP = (size_t)a.W * a.H;
        std::vector<std::unique_ptr<rss::Session>> sessions(workers);
        for (int w = 0; w < workers; w++) sessions[w].reset(new rss::Session(a.config, a.forest, w % a.gpus));
        const int layers = sessions[0]->info().layer_count;
        const int keep = a.out.empty() ? 0 : std::min(a.frames, frames.count);
        std::vector<uint8_t> kept((size_t)keep * layers * NP);
        std::vector<uint8_t*> staging(workers);
        for (int w = 0; w < workers; w++) staging[w] = (uint8_t*)rss_host_alloc(layers * NP);

        auto run = [&](int total, bool record) {
            std::atomic<int> next(0);
            std::atomic<int> failed(0);
            std::vector<std::thread> th;
            for (int w = 0; w < workers; w++)
                th.emplace_back([&, w]() {
                    rss::Session& s = *sessions[w];
                    for (;;) {
                        const int k = next.fetch_add(1);
                        if (k >= total) break;
                        const int f = k % frames.count;
                        const rss_status st = rss_segment_keyframe(s.handle(), frames.rgb + frames.rgb_bytes * f,
                                                                   frames.depth + NP * f, a.W, a.H, cal.intrinsic_inverse,
                                                                   cal.extrinsic_linear, cal.extrinsic_translation, &prm,
                                                                   staging[w], nullptr);
                        if (st != RSS_OK) { fprintf(stderr, "worker %d: %s\n", w, rss_last_error(s.handle())); failed++; break; }
                        if (record && k < keep) memcpy(kept.data() + (size_t)k * layers * NP, staging[w], layers * NP);
                    }
                });
            for (auto& t : th) t.join();
            return failed.load();
        };
        if (run(2 * workers, false)) return 1;  // warm-up: allocations, lattice capacities
        const auto t0 = std::chrono::steady_clock::now();
        if (run(a.frames, true)) return 1;
        const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        for (uint8_t* p : staging) rss_host_free(p);
        if (!a.out.empty()) std::ofstream(a.out, std::ios::binary).write((const char*)kept.data(), kept.size());
        printf("{\"metric\": \"keyframes/sec RF+DenseCRF %dx%d\", \"value\": %.3f, \"unit\": \"keyframes/s\", \"n_gpus\": %d, "
               "\"inflight_per_gpu\": %d, \"keyframes\": %d, \"seconds\": %.4f, \"host\": \"c++ keyframe_worker\"}\n",
               a.W, a.H, a.frames / dt, a.gpus, a.inflight, a.frames, dt);
    } catch (const std::exception& e) {
        fprintf(stderr, "keyframe_worker: %s\n", e.what());
        return 1;
    }
    return 0;
}
