#!/usr/bin/env python
"""Benchmark of the per-keyframe inference path (BASELINE.json metric: keyframes/sec, RF + DenseCRF, 640x480).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one synthetic 640x480 RGB-D keyframe through the whole hot path: Lab + patch features + depth/height/
normal features -> multi-label forest (4 trees, 17 classes in 2 layers) -> 2x upsample -> DenseCRF over the 307 200
pixels with a 3-D Gaussian kernel on the back-projected points and a 5-D bilateral kernel, 10 mean-field iterations,
both label layers, gated argmax (BASELINE.json configs[1], with configs[2]'s second layer included).

  value : keyframes/s with the frame already resident in HBM (device time, CUDA events on the library's stream)
  e2e   : keyframes/s through the C ABI with HOST buffers: pinned rgb/depth in, label maps out, copies timed
N > 1 (torchrun): every rank owns one GPU and its own keyframes (weak scaling, no collective on the data path);
the timed region is bracketed by a barrier + synchronize and the slowest rank's time is used.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
FOREST = os.path.join(ROOT, "tests", "golden", "forest_shared.dat")
W, H = 640, 480
METRIC = "keyframes/sec RF+DenseCRF 640x480"
WORKLOAD = ("single-frame RF + DenseCRF: 640x480 RGB-D, stride-2 forest (4 trees, 366 features, 8+9 classes), "
            "Gaussian 3-D (5 cm, w=3) + bilateral 5-D (80 px / 13, w=10) kernels, 10 mean-field iterations, "
            "both label layers")
KF = dict(sigma_xyz=0.05, w_gauss=3.0, sigma_px=80.0, sigma_rgb=13.0, w_bilateral=10.0, iters=10, fill=0.0)
N_FRAMES = 8  # distinct synthetic frames cycled through


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.stop_flag = False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [x.strip() for x in out.strip().split(",")]
                if len(parts) >= 6:
                    self.rows.append(parts)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[2 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows)}


# ------------------------------------------------------------------------------------------------ CPU reference arm
def cpu_keyframe(seed):
    """The oracle's single-thread restatement of one keyframe (how the reference runs inference)."""
    import oracle
    from rovinasemanticsegmentation_b200 import synth
    oracle.set_threads(1)
    rgb, depth = synth.frame(seed)
    Kinv, R, t = synth.calibration()
    forest = oracle.Forest(FOREST)
    t0 = time.perf_counter()
    post = oracle.segment_frame(oracle.default_config(), forest, 2, rgb, depth, Kinv, R, t, 0.5, 15.0, KF["fill"])
    xyz = oracle.cloud(depth, Kinv, R, t, 0.5, 15.0).reshape(-1, 3)
    xyz[np.isnan(xyz[:, 0])] = t
    f3 = (xyz * np.float32(1.0 / KF["sigma_xyz"])).astype(np.float32)
    f5 = oracle.features_bilateral2d(W, H, KF["sigma_px"], KF["sigma_px"], KF["sigma_rgb"], KF["sigma_rgb"],
                                     KF["sigma_rgb"], rgb)
    off, N = 0, W * H
    for M, unk in ((8, 7), (9, 8)):
        Q = oracle.crf_inference(-post[off:off + N * M].reshape(N, M), [(f3, KF["w_gauss"]), (f5, KF["w_bilateral"])],
                                 KF["iters"])
        oracle.gated_argmax(Q, unk)
        off += N * M
    return time.perf_counter() - t0


def _cpu_worker(seed):
    return cpu_keyframe(seed)


def cpu_throughput(procs, rounds):
    """`procs` independent single-thread pipelines side by side (keyframes are independent), `rounds` times."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    with ctx.Pool(procs) as pool:
        pool.map(_cpu_worker, range(procs))  # warm-up: page in libraries, build LUTs
        t0 = time.perf_counter()
        for r in range(rounds):
            pool.map(_cpu_worker, range(100 + r * procs, 100 + (r + 1) * procs))
        dt = time.perf_counter() - t0
    return procs * rounds / dt, dt


def run_reference(args, rank, world):
    if rank != 0:
        return
    import oracle
    oracle.build(ref=False)
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, 32))
    steps = max(1, min(args.steps, 3))
    for _ in range(0):
        pass
    val, dt = cpu_throughput(procs, steps)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "keyframes/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": 1, "ms_per_step": 1000.0 * dt / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "l2": "CPU run"},
        "cpu_baseline": {"value": val, "unit": "keyframes/s", "cores": procs, "kind": "port",
                         "sample": "%d rounds of %d full 640x480 keyframes, one single-thread oracle pipeline per core "
                                   "(oracle/oracle.c, validated bit-exact against the compiled reference)" % (steps, procs)},
        "e2e": {"value": val, "unit": "keyframes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
def algo_bytes(name, env):
    """Algorithmic (compulsory) bytes of ONE launch of kernel `name`: every unique input byte read once, every output
    byte written once (DESIGN.md "Kernels").  Kernels launched once per lattice get the mean over the
    lattices.  env: N = W*H CRF points, M = labels over both layers, Ns = forest samples, D = features per sample,
    T trees, nodes/leaves of the forest, lat = [(d, V)] of the keyframe's lattices, P = patch_size (border)."""
    N, M, Ns, D, T, P = env["N"], env["M"], env["Ns"], env["D"], env["T"], env["P"]
    Wb, Hb = W + 2 * P, H + 2 * P
    lat = env["lat"]
    mean = lambda f: sum(f(d, V) for d, V in lat) / max(len(lat), 1)
    table = {
        # F1..F4 (feature_extractor.h)
        "lab_border_kernel": 3 * N + 3 * Wb * Hb,
        "patch_features_kernel": 3 * Wb * Hb + 2 * Ns + 8 * Ns + 4 * (D - 3) * Ns,
        "cloud_kernel": 2 * N + 12 * N,
        "select_kernel": 2 * Ns + 4 * Ns, "compact_kernel": 8 * Ns + 8 * Ns,
        "gradient_mask_kernel": 12 * N + 24 * N + 2 * N + 4 * N,
        "dist_forward_kernel": 8 * N, "dist_backward_kernel": 8 * N,
        "integral_wavefront_kernel<512>": 24 * N + 2 * N + 48 * N + 8 * N,
        "scalar_features_kernel": 8 * Ns + 2 * Ns + 12 * Ns + 12 * Ns,
        # R1, R2 (classifier.cpp, segmenter.cpp:355-431)
        "forest_traverse_kernel": 4 * D * Ns + 16 * env["nodes"] + 4 * T * Ns,
        "forest_posterior_kernel": 4 * T * Ns + 4 * M * env["leaves"] + 4 * M * Ns,
        "fill_kernel": 4 * M * (N // 4), "lowres_scatter_kernel": 8 * M * Ns + 8 * Ns,
        "upsample_kernel": 4 * M * (N // 4) + 4 * M * N,
        "unary_from_posteriors_kernel": 8 * M * N,
        # C0 lattice construction (permutohedral.cpp:140-321)
        "lattice_embed_kernel<D>": mean(lambda d, V: 4 * d * N + 8 * (d + 1) * N + 2 * d * V),
        "remap_offsets_kernel": mean(lambda d, V: 8 * (d + 1) * N), "csr_fill_kernel": mean(lambda d, V: 16 * (d + 1) * N),
        "first_flags_kernel": mean(lambda d, V: 8 * (d + 1) * N), "assign_ids_kernel": mean(lambda d, V: 8 * (d + 1) * N),
        # C1 mean-field iteration (permutohedral.cpp:529-589, densecrf.cpp:98-131)
        "softmax_init_kernel": 8 * M * N,
        "splat_kernel": mean(lambda d, V: 4 * M * N + 8 * (d + 1) * N + 4 * N + 4 * M * V),
        "blur_coop_kernel": mean(lambda d, V: (d + 1) * (8 * M * V + 8 * V) + 4 * M * V),
        "blur_kernel": mean(lambda d, V: 8 * M * V + 8 * V),
        "slice_softmax_kernel<MP>": sum(8 * (d + 1) * N + 4 * N + 4 * M * V for d, V in lat) + 8 * M * N,
    }
    return table.get(name)


def run_gpu(args, rank, world, local_rank):
    import torch
    import rovinasemanticsegmentation_b200 as rss
    from rovinasemanticsegmentation_b200 import synth

    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    ctx = rss.Context(rss.DEFAULT_CONFIG, FOREST, local_rank)
    Kinv, R, t = synth.calibration()
    frames = []
    for k in range(N_FRAMES):
        rgb, depth = synth.frame(10_000 + rank * N_FRAMES + k)
        prgb = torch.empty((H, W, 3), dtype=torch.uint8, pin_memory=True)
        pdep = torch.empty((H, W), dtype=torch.int16, pin_memory=True)
        prgb.numpy()[...] = rgb
        pdep.numpy().view(np.uint16)[...] = depth
        frames.append((prgb.numpy(), pdep.numpy().view(np.uint16)))
    plabels = torch.empty((2, H * W), dtype=torch.uint8, pin_memory=True)
    labels_np = plabels.numpy()
    prm = rss.KeyframeParams(KF["sigma_xyz"], KF["w_gauss"], KF["sigma_px"], KF["sigma_rgb"], KF["w_bilateral"],
                             KF["iters"], KF["fill"])
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    import ctypes as C
    lib = ctx._lib
    Kp, Rp, tp = (np.ascontiguousarray(a, np.float32).reshape(-1) for a in (Kinv, R, t))

    def call(rgb, depth, labels):
        st = lib.rss_segment_keyframe(ctx.h, rss._ptr(rgb, C.c_uint8), rss._ptr(depth, C.c_uint16), W, H,
                                      rss._ptr(Kp, C.c_float), rss._ptr(Rp, C.c_float), rss._ptr(tp, C.c_float),
                                      C.byref(prm), rss._ptr(labels, C.c_uint8), None)
        ctx._check(st)

    # ---- warm-up (untimed)
    for k in range(args.warmup):
        call(frames[k % N_FRAMES][0], frames[k % N_FRAMES][1], labels_np)
    barrier()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()

    # ---- device-resident throughput: frame k resident before the step, L2 flushed, device time from CUDA events
    # (the library brackets every call with events on its own stream).  Pass 1 is the timed region; pass 2 repeats the
    # same K steps with per-kernel event bracketing switched on (costs a few microseconds per launch, so it is kept out
    # of the headline number) and yields the per-kernel table and the roofline of the dominant kernel.
    def resident_pass(profile):
        ctx.profile_enable(profile)
        l0 = ctx.kernel_launches
        ms, stg = 0.0, {}
        barrier()
        for k in range(args.steps):
            ctx.upload_frame(*frames[k % N_FRAMES])
            flush.zero_()
            torch.cuda.synchronize()
            call(None, None, None)
            tm = ctx.timings()
            ms += tm["total_ms"]
            for n, v in tm.items():
                stg[n] = stg.get(n, 0.0) + v
        barrier()
        rep = ctx.profile_report() if profile else {}
        ctx.profile_enable(False)
        return ms, stg, ctx.kernel_launches - l0, rep

    dev_ms, stage, launches, _ = resident_pass(False)
    lat_info = [ctx.keyframe_lattice_info(k) for k in range(2)]
    n_samples = int(((frames[0][1][::2, ::2] >= 500) & (frames[0][1][::2, ::2] <= 15000)).sum())
    dev_ms_prof, _, _, prof = resident_pass(True)

    # ---- several keyframes in flight: C contexts on this GPU, one host thread each (keyframes are independent, the
    # reference runs one worker per stage; here one worker per context).  Wall clock over the whole batch between
    # device-wide synchronisations; every context works on its own frames.
    def run_inflight(C_, steps_total, host_io):
        import threading as th
        ctxs = [ctx] + [rss.Context(rss.DEFAULT_CONFIG, FOREST, local_rank) for _ in range(C_ - 1)]
        outs = [np.empty((2, H * W), np.uint8) if not host_io else
                torch.empty((2, H * W), dtype=torch.uint8, pin_memory=True).numpy() for _ in range(C_)]
        per = [steps_total // C_ + (1 if i < steps_total % C_ else 0) for i in range(C_)]

        def worker(ci, n, warm):
            c = ctxs[ci]
            for k in range(n):
                fr = frames[(ci * 3 + k) % N_FRAMES]
                if host_io:
                    stc = lib.rss_segment_keyframe(c.h, rss._ptr(fr[0], C.c_uint8), rss._ptr(fr[1], C.c_uint16), W, H,
                                                   rss._ptr(Kp, C.c_float), rss._ptr(Rp, C.c_float), rss._ptr(tp, C.c_float),
                                                   C.byref(prm), rss._ptr(outs[ci], C.c_uint8), None)
                else:
                    if warm or k == 0:
                        c.upload_frame(*fr)
                    stc = lib.rss_segment_keyframe(c.h, None, None, W, H, rss._ptr(Kp, C.c_float), rss._ptr(Rp, C.c_float),
                                                   rss._ptr(tp, C.c_float), C.byref(prm), None, None)
                c._check(stc)

        for phase in ("warm", "timed"):
            ths = [th.Thread(target=worker, args=(i, 2 if phase == "warm" else per[i], phase == "warm")) for i in range(C_)]
            torch.cuda.synchronize()
            t0_ = time.perf_counter()
            for t_ in ths:
                t_.start()
            for t_ in ths:
                t_.join()
            torch.cuda.synchronize()
            dt_ = time.perf_counter() - t0_
        for c in ctxs[1:]:
            c.close()
        return dt_

    sweep = [] if args.quick else ([args.inflight] if args.inflight > 0 else [1, 2, 3, 4])
    inflight = {}
    for C_ in sweep:
        barrier()
        dt_res = run_inflight(C_, args.steps, False)
        dt_e2e = run_inflight(C_, args.steps, True)
        inflight[C_] = (dt_res, dt_e2e)
    barrier()

    # ---- end to end through the C ABI with host buffers (pinned), wall clock over K synchronous calls
    barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        call(frames[k % N_FRAMES][0], frames[k % N_FRAMES][1], labels_np)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()

    if rank == 0:
        sampler.stop_flag = True
        sampler.join(timeout=2)

    # slowest rank decides
    flat = [dev_ms, e2e_s * 1000.0]
    for C_ in sweep:
        flat += [inflight[C_][0] * 1000.0, inflight[C_][1] * 1000.0]
    times = torch.tensor(flat, dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    flat = times.tolist()
    dev_ms_max, e2e_ms_max = flat[0], flat[1]
    inflight_max = {C_: (flat[2 + 2 * i], flat[3 + 2 * i]) for i, C_ in enumerate(sweep)}

    if rank == 0:
        total_kf = args.steps * world
        value = total_kf / (dev_ms_max / 1000.0)
        e2e = total_kf / (e2e_ms_max / 1000.0)
        peak, peak_src = peaks()
        # dominant kernel = largest accumulated device time in the timed region
        top = sorted(prof.items(), key=lambda kv: -kv[1][0])
        kernel_table = [{"kernel": n, "ms_per_step": ms / args.steps, "launches_per_step": cnt / args.steps,
                         "us_per_launch": 1000.0 * ms / max(cnt, 1)} for n, (ms, cnt) in top[:16]]
        info = ctx.info
        env = {"N": W * H, "M": int(info.total_classes), "Ns": int(n_samples), "D": int(info.feature_length),
               "T": int(info.num_trees), "nodes": int(info.total_nodes), "leaves": int(info.total_leaves),
               "P": int(info.patch_size), "lat": lat_info}
        for row, (n, (ms, cnt)) in zip(kernel_table, top[:16]):
            ab = algo_bytes(n, env)
            if ab is not None:
                row["algorithmic_bytes_per_launch"] = int(ab)
                row["achieved_gbs"] = ab / (ms / max(cnt, 1) / 1000.0) / 1e9
                row["frac_of_hbm_peak"] = row["achieved_gbs"] / peak
        dom_name, (dom_ms, dom_cnt) = top[0]
        ab = algo_bytes(dom_name, env)
        ach = ab / (dom_ms / dom_cnt / 1000.0) / 1e9 if ab is not None else None
        traffic = None
        tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")  # dram bytes per launch from `ncu --set full` captures
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get(dom_name, {}).get("dram_bytes_per_launch")
        roof = {"bound": "hbm", "kernel": dom_name, "achieved": ach, "peak": peak, "unit": "GB/s",
                "frac": ach / peak if ach is not None else None, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": int(ab) if ab is not None else None,
                "us_per_launch": 1000.0 * dom_ms / dom_cnt,
                "lattices": [{"d": d, "vertices": V} for d, V in lat_info]}
        # whole-step roofline: algorithmic bytes of every launch of the step / the step's device time
        step_bytes = sum((algo_bytes(n, env) or 0) * cnt / args.steps for n, (ms, cnt) in top)
        roof["step_algorithmic_bytes"] = int(step_bytes)
        roof["step_achieved"] = step_bytes / (dev_ms_max / args.steps / 1000.0) / 1e9
        roof["step_frac"] = roof["step_achieved"] / peak
        sweep_out = {str(C_): {"resident_kf_s": total_kf / (a / 1000.0), "e2e_kf_s": total_kf / (b / 1000.0)}
                     for C_, (a, b) in inflight_max.items()}
        line = {
            "inflight_sweep": sweep_out,
            "metric": METRIC, "value": value, "unit": "keyframes/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "l2": "flushed with a 256 MiB write before every timed step",
                       "frames": "%d distinct synthetic frames per rank" % N_FRAMES},
            "clocks": sampler.summary(),
            "e2e": {"value": e2e, "unit": "keyframes/s", "h2d_bytes_per_step": W * H * 3 + W * H * 2,
                    "d2h_bytes_per_step": 2 * W * H, "ms_per_step": e2e_ms_max / args.steps},
            "gpu_launches": int(launches),
            "roofline": roof,
            "stages_ms_per_step": {n: v / args.steps for n, v in stage.items()},
            "kernels": kernel_table,
            "kernels_note": "per-kernel times from a second pass over the same steps with event bracketing on "
                            "(%.3f ms/step vs %.3f ms/step in the timed pass)" % (dev_ms_prof / args.steps, dev_ms / args.steps),
            "ms_per_meanfield_iter": stage.get("meanfield_ms", 0.0) / args.steps / KF["iters"],
        }
        if world == 1 and not args.no_cpu and not args.quick:
            import oracle
            oracle.build(ref=False)
            procs = max(1, min(os.cpu_count() or 1, 32))
            cv, cdt = cpu_throughput(procs, 1)
            line["cpu_baseline"] = {"value": cv, "unit": "keyframes/s", "cores": procs, "kind": "port",
                                    "sample": "%d full 640x480 keyframes, one single-thread oracle pipeline per core "
                                              "(%.1f s)" % (procs, cdt)}
        print(json.dumps(line), flush=True)
    ctx.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--quick", action="store_true", help="skip the keyframes-in-flight sweep and the cpu_baseline leg (profiling runs)")
    ap.add_argument("--inflight", type=int, default=0,
                    help="keyframes in flight per GPU (one context + host thread each); 0 = sweep 1,2,3,4 and report the best")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world == 1 and args.gpus > 1:
        # plain `python bench.py --gpus N`: re-launch under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__), "--gpus", str(args.gpus),
               "--steps", str(args.steps), "--warmup", str(args.warmup)] + (["--no-cpu"] if args.no_cpu else [])
        sys.exit(subprocess.call(cmd))
    run_gpu(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
