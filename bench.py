#!/usr/bin/env python
"""Benchmark of the per-keyframe inference path (BASELINE.json metric: keyframes/sec, RF + DenseCRF, 640x480).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one synthetic 640x480 RGB-D keyframe through the whole hot path: Lab + patch features + depth/height/
normal features -> multi-label forest (4 trees, 17 classes in 2 layers) -> 2x upsample -> DenseCRF over the 307 200
pixels with a 3-D Gaussian kernel on the back-projected points and a 5-D bilateral kernel, 10 mean-field iterations,
both label layers, gated argmax (BASELINE.json configs[1], with configs[2]'s second layer included).

  value : keyframes/s with the frames already resident in HBM, --inflight keyframes in flight per GPU (CUDA events around
          the K steps, device-wide synchronisation on both sides)
  e2e   : the same through the C ABI with HOST buffers: pinned rgb/depth in, label maps out, copies inside the timed region
  latency: one keyframe at a time (library CUDA events), with the per-stage and per-kernel break-down
  cpp_worker: the C++ host (host/keyframe_worker.cpp, one process driving all N GPUs, one thread per keyframe in flight) on
          the same workload, run by rank 0 after the timed region while the ranks are idle: wall clock, end to end
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


def frame_seeds(rank, n=None):
    """Seeds of the synthetic frames rank `rank` owns: every rank works on its own keyframes (weak scaling; keyframes are
    independent units, so the shards share nothing)."""
    n = N_FRAMES if n is None else n
    return [10_000 + rank * n + k for k in range(n)]


def max_over_ranks(values, dist=None, device=None):
    """Slowest rank decides: element-wise MAX of per-rank timings over the process group (identity without one)."""
    import torch
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


def job_throughput(steps_per_rank, world, ms):
    """Whole-job keyframes/s: every rank ran `steps_per_rank` keyframes inside the same (max over ranks) `ms`."""
    return steps_per_rank * world / (ms / 1000.0)


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
_CPU_FOREST = None
_CPU_BARRIER = None


def _cpu_init(barrier=None):
    """Pool initialiser: load the checker library and the forest once per worker (outside every timed region)."""
    global _CPU_FOREST, _CPU_BARRIER
    import oracle
    oracle.set_threads(1)
    _CPU_FOREST = oracle.Forest(FOREST)
    _CPU_BARRIER = barrier


def cpu_keyframe(seed, want_labels=False):
    """The oracle's single-thread restatement of one keyframe (how the reference runs inference).  Only the compute is
    timed: frame synthesis and the forest load are outside the clock."""
    import oracle
    from rovinasemanticsegmentation_b200 import synth
    if _CPU_FOREST is None:
        _cpu_init()
    rgb, depth = synth.frame(seed)
    Kinv, R, t = synth.calibration()
    if _CPU_BARRIER is not None:
        # all pipelines of a round start together (a worker waiting here cannot take a second task, so the tasks of a round
        # land on distinct workers and the round's time is the slowest pipeline's compute time)
        _CPU_BARRIER.wait(timeout=300)
    t0 = time.perf_counter()
    labels, _, _ = oracle.keyframe(_CPU_FOREST, rgb, depth, Kinv, R, t, **KF)
    dt = time.perf_counter() - t0
    return (dt, labels) if want_labels else (dt, None)


def _cpu_worker(arg):
    seed, want = arg
    return cpu_keyframe(seed, want)


def cpu_throughput(procs, rounds, warmup=1, first_seed=None):
    """`procs` independent single-thread pipelines side by side (keyframes are independent), `rounds` times.  A round's
    time is its slowest pipeline's compute time.  first_seed: worker 0 of the first timed round runs that frame and its
    label maps are returned (parity check of the GPU arm against the very keyframes the CPU arm computed)."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    with ctx.Pool(procs, initializer=_cpu_init, initargs=(ctx.Barrier(procs),)) as pool:
        for w in range(max(1, warmup)):  # page in libraries, build LUTs
            pool.map(_cpu_worker, [(50 + w * procs + i, False) for i in range(procs)], chunksize=1)
        total, labels = 0.0, None
        for r in range(rounds):
            seeds = [(100 + r * procs + i, False) for i in range(procs)]
            if r == 0 and first_seed is not None:
                seeds[0] = (first_seed, True)
            res = pool.map(_cpu_worker, seeds, chunksize=1)
            total += max(dt for dt, _ in res)
            if r == 0 and first_seed is not None:
                labels = res[0][1]
    return procs * rounds / total, total, labels


def run_reference(args, rank, world):
    if rank != 0:
        return
    import oracle
    oracle.build(ref=False)
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, 32))
    steps, warmup = max(1, args.steps), max(1, args.warmup)
    val, dt, _ = cpu_throughput(procs, steps, warmup)
    sample = ("one step = %d full 640x480 keyframes side by side, one single-thread oracle pipeline per host core "
              "(oracle/oracle.c, pinned bit-exact / 1e-6 against the compiled reference); %d steps after %d warm-up steps; "
              "a step's time is its slowest pipeline's compute time (frame synthesis and forest load excluded)"
              % (procs, steps, warmup))
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "keyframes/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": 1000.0 * dt / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "l2": "CPU run", "step": "%d keyframes (one per host core)" % procs},
        "cpu_baseline": {"value": val, "unit": "keyframes/s", "cores": procs, "kind": "port", "sample": sample},
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
        # frame path: features evaluated on demand - the Lab image, the depth image and the cloud in, leaf ids out
        "forest_traverse_frame_kernel": 3 * Wb * Hb + 2 * N + 16 * N + 8 * Ns + 16 * env["nodes"] + 4 * T * Ns,
        "forest_posterior_kernel": 4 * T * Ns + 4 * M * env["leaves"] + 4 * M * Ns,
        "fill_kernel": 4 * M * (N // 4), "lowres_scatter_kernel": 8 * M * Ns + 8 * Ns,
        "upsample_kernel": 4 * M * (N // 4) + 4 * M * N, "upsample_kernel<true>": 4 * M * (N // 4) + 4 * M * N,
        "upsample_kernel<false>": 4 * M * (N // 4) + 4 * M * N,
        "upsample_unary_groups_kernel": 4 * M * (N // 4) + 4 * M * N,
        # fused frame kernel: Lab image, depth + cloud + distance / integral images at the samples,
        # the model, the low-res posterior image out
        "forest_frame_lowres_kernel": 3 * Wb * Hb + 18 * N + 16 * env["nodes"] + 4 * M * env["leaves"] + 4 * M * (N // 4),
        "unary_from_posteriors_kernel": 8 * M * N,
        # C0 lattice construction (permutohedral.cpp:140-321)
        "lattice_embed_kernel<D>": mean(lambda d, V: 4 * d * N + 8 * (d + 1) * N + 2 * d * V),
        "remap_offsets_kernel": mean(lambda d, V: 8 * (d + 1) * N), "csr_fill_kernel": mean(lambda d, V: 16 * (d + 1) * N),
        "first_flags_kernel": mean(lambda d, V: 8 * (d + 1) * N), "assign_ids_kernel": mean(lambda d, V: 8 * (d + 1) * N),
        # bitmap numbering (lattice.cu): hash table of ~4 V slots; one bit per (point, corner) pair; one prefix word per 32 bits
        "first_bitmap_kernel": mean(lambda d, V: 16 * V + (d + 1) * N / 8),
        "bitmap_chunk_sums_kernel": mean(lambda d, V: (d + 1) * N / 8),
        "bitmap_prefix_kernel": mean(lambda d, V: (d + 1) * N / 8 + (d + 1) * N / 8),
        "assign_ids_bitmap_kernel": mean(lambda d, V: 32 * V + 8 * V),
        "unary_init_kernel": 4 * (M + 3) * N,
        # C1 mean-field iteration (permutohedral.cpp:529-589, densecrf.cpp:98-131)
        "softmax_init_kernel": 8 * M * N,
        "splat_kernel": mean(lambda d, V: 4 * M * N + 8 * (d + 1) * N + 4 * N + 4 * M * V),
        "blur_coop_kernel": mean(lambda d, V: (d + 1) * (8 * M * V + 8 * V) + 4 * M * V),
        "blur_kernel": mean(lambda d, V: 8 * M * V + 8 * V),
        "slice_softmax_kernel<MP>": sum(8 * (d + 1) * N + 4 * N + 4 * M * V for d, V in lat) + 8 * M * N,
        # fused path (meanfield.cu): one point kernel per iteration = slice (vertex ids + weights, norms, unique rows) +
        # unary + tile-local splat (pair list, unique rows out); one cooperative blur for all lattices
        # one point kernel per iteration: unary rows + per (point, corner) slice weight (4 B) and row slot (2 B) + the
        # distinct value rows in; splat pairs (8 B per (point, corner)) in, the distinct value rows out
        "meanfield_point_kernel": 4 * M * N + sum(14 * (d + 1) * N + 8 * M * V for d, V in lat),
        "meanfield_point_kernel<first>": 4 * M * N + sum(8 * (d + 1) * N + 4 * M * V for d, V in lat),
        # the bench's keyframes ask for label maps only (Q = NULL): the last pass reads the unary rows and writes 2 label bytes
        "meanfield_point_kernel<last>": 4 * M * N + 2 * N + sum(6 * (d + 1) * N + 4 * M * V for d, V in lat),
        "blur_multi_coop_kernel": sum((d + 1) * (8 * M * V + 8 * V) + 4 * M * V for d, V in lat),
        "tile_csr_build_kernel<D>": mean(lambda d, V: 8 * (d + 1) * N + 4 * N + (8 + 6) * (d + 1) * N + 12 * V),
        "splat_ones_runs_kernel<D>": mean(lambda d, V: 8 * (d + 1) * N + 4 * V),
        "slice_kernel": mean(lambda d, V: 8 * (d + 1) * N + 4 * V + 4 * N),
        # two hash look-ups per (vertex, axis): the vertex key once, the int2 neighbour pair out, the two probed slots in
        "neighbors_kernel<D>": mean(lambda d, V: 16 * V + (d + 1) * V * (8 + 2 * 16 + 2 * 4)),
        "zero_rows_kernel": mean(lambda d, V: 16 * V), "norm_kernel": 8 * N,
        "feat_bilateral_kernel": 3 * N + 20 * N, "feat_frame_xyz_kernel": 16 * N + 12 * N,
        "lowres_fill_kernel": 4 * M * (N // 4),
    }
    if name in table:
        return table[name]
    import re
    return table.get(re.sub(r"<[0-9, ]+>", "<D>", name))


# ------------------------------------------------------------------------------------------------ local-map sweep
# BASELINE configs[4]: independent local maps sharded over the GPUs (src/segmenter.cpp:518-719, the map worker's unit of
# work).  One map = MAP_KF key frames (640x480, the frame worker's forest pass each) fused into one cloud of MAP_POINTS points:
# projector -> unary accumulation -> one 6-D Potts kernel (xyz * 0.5, rgb * 4, w = 10) -> 10 mean-field iterations for
# both label layers -> gated argmax.  Host buffers in (cloud 48 MB, frames) and label maps out, all inside the timed region.
MAP_POINTS, MAP_KF, MAP_DISTINCT = 2_000_000, 4, 2
MAP_METRIC = "local maps/sec (2M points, 4 keyframes 640x480, RF + 6-D DenseCRF, 10 iters, 2 layers)"
MAP_Z = (0.3, 8.0)


def map_inputs(seed):
    from rovinasemanticsegmentation_b200 import synth
    xyz, col = synth.local_map(seed=seed, n_points=MAP_POINTS)
    frames = [synth.frame(seed * 16 + k) for k in range(MAP_KF)]
    poses = [synth.map_keyframe_pose(k, MAP_KF) for k in range(MAP_KF)]
    return xyz, col, frames, poses


def cpu_local_map(seed, want_labels=False):
    """One local map through the oracle, single thread (the reference's map worker is one thread)."""
    import oracle
    from rovinasemanticsegmentation_b200 import synth
    if _CPU_FOREST is None:
        _cpu_init()
    xyz, col, frames, poses = map_inputs(seed)
    Kinv, Rc, tc = synth.calibration()
    K = synth.intrinsics()
    if _CPU_BARRIER is not None:
        _CPU_BARRIER.wait(timeout=600)
    t0 = time.perf_counter()
    un = [np.zeros((MAP_POINTS, m), np.float32) for m in (8, 9)]
    for (rgb, depth), (R, t) in zip(frames, poses):
        post = oracle.segment_frame(oracle.default_config(), _CPU_FOREST, 2, rgb, depth, Kinv, Rc, tc, 0.5, 15.0, 0.0)
        idx = oracle.project_zbuffer(xyz, K, R, t, W, H, *MAP_Z)
        off = 0
        for l, m in enumerate((8, 9)):
            oracle.unary_accumulate(idx, post[off:off + W * H * m].reshape(W * H, m), un[l])
            off += W * H * m
    f6 = oracle.features_xyzrgb(xyz, col, 0.5, 4.0)
    labels = np.empty((2, MAP_POINTS), np.uint8)
    for l, unk in enumerate((7, 8)):
        labels[l] = oracle.gated_argmax(oracle.crf_inference(-un[l], [(f6, 10.0)], 10), unk)
    dt = time.perf_counter() - t0
    return (dt, labels) if want_labels else (dt, None)


def _cpu_map_worker(arg):
    seed, want = arg
    return cpu_local_map(seed, want)


def cpu_map_throughput(procs, rounds, first_seed=None):
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    with ctx.Pool(procs, initializer=_cpu_init, initargs=(ctx.Barrier(procs),)) as pool:
        total, labels = 0.0, None
        for r in range(rounds):
            seeds = [(5000 + r * procs + i, False) for i in range(procs)]
            if r == 0 and first_seed is not None:
                seeds[0] = (first_seed, True)
            res = pool.map(_cpu_map_worker, seeds, chunksize=1)
            total += max(dt for dt, _ in res)
            if r == 0 and first_seed is not None:
                labels = res[0][1]
    return procs * rounds / total, total, labels


def run_local_maps_reference(args, rank, world):
    if rank != 0:
        return
    import oracle
    oracle.build(ref=False)
    procs = max(1, min(os.cpu_count() or 1, 16))  # ~1.5 GB per pipeline
    rounds = max(1, min(args.steps, 2))
    val, dt, _ = cpu_map_throughput(procs, rounds)
    print(json.dumps({
        "impl": "reference", "metric": MAP_METRIC, "value": val, "unit": "maps/s", "n_gpus": args.gpus, "steps": rounds,
        "warmup": 0, "ms_per_step": 1000.0 * dt / rounds, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "local_maps", "step": "%d maps side by side, one single-thread oracle pipeline per core" % procs},
        "cpu_baseline": {"value": val, "unit": "maps/s", "cores": procs, "kind": "port",
                         "sample": "%d rounds of %d maps (oracle/oracle.c)" % (rounds, procs)},
        "e2e": {"value": val, "unit": "maps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)


def run_local_maps(args, rank, world, local_rank):
    import threading as th

    import torch
    import rovinasemanticsegmentation_b200 as rss
    from rovinasemanticsegmentation_b200 import synth

    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if dist is not None:
        dist.all_reduce(torch.zeros(1, device=dev))
        torch.cuda.synchronize()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    NC = max(1, min(2, (os.cpu_count() or 1) // max(world, 1)))  # maps in flight per GPU
    ctxs = [rss.Context(rss.DEFAULT_CONFIG, FOREST, local_rank) for _ in range(NC)]
    crfs = [c.crf(MAP_POINTS, [8, 9]) for c in ctxs]
    Kinv, Rc, tc = synth.calibration()
    K = synth.intrinsics()

    def pinned(a):
        t = torch.empty(a.shape, dtype=torch.from_numpy(a[:0]).dtype, pin_memory=True)
        t.numpy()[...] = a
        return t.numpy()

    maps = []
    seeds = [7000 + rank * MAP_DISTINCT + i for i in range(MAP_DISTINCT)]
    for seed in seeds:
        xyz, col, frames, poses = map_inputs(seed)
        pf = []
        for rgb, depth in frames:
            pd = torch.empty(depth.shape, dtype=torch.int16, pin_memory=True)
            pd.numpy().view(np.uint16)[...] = depth
            pf.append((pinned(rgb), pd.numpy().view(np.uint16)))
        maps.append((pinned(xyz), pinned(col), pf, poses))
    outs = [None] * NC

    def one_map(i, m):
        c, crf = ctxs[i], crfs[i]
        xyz, col, frames, poses = maps[m]
        for k, (rgb, depth) in enumerate(frames):  # frame worker: forest pass, posteriors stay on the device
            c.segment_frame(rgb, depth, Kinv, Rc, tc, 0.0, want_host=False)
            c.posteriors_keep(k)
        crf.unary_reset()
        crf.clear_pairwise()
        crf.set_cloud(xyz, col)
        for k, (R, t) in enumerate(poses):       # map worker: projector + accumulation on the device
            crf.project_accumulate(W, H, K, R, t, MAP_Z[0], MAP_Z[1], slot=k)
        crf.add_pairwise_cloud(0.5, 4.0, 10.0)
        outs[i] = crf.inference(10, unknown=[7, 8], want_Q=False, want_labels=True)

    def run(steps):
        per = [steps // NC + (1 if i < steps % NC else 0) for i in range(NC)]
        errs = []

        def worker(i):
            try:
                for k in range(per[i]):
                    one_map(i, (i + k) % MAP_DISTINCT)
            except Exception as e:  # noqa: BLE001
                errs.append(e)
        ths = [th.Thread(target=worker, args=(i,)) for i in range(NC)]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for t_ in ths:
            t_.start()
        for t_ in ths:
            t_.join()
        torch.cuda.synchronize()
        e1.record()
        e1.synchronize()
        if errs:
            raise errs[0]
        return e0.elapsed_time(e1)

    l0 = sum(c.kernel_launches for c in ctxs)
    run(max(args.warmup, NC))
    launches_warm = sum(c.kernel_launches for c in ctxs) - l0
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = sum(c.kernel_launches for c in ctxs)
    ms = run(args.steps)
    launches = sum(c.kernel_launches for c in ctxs) - l0
    one_map(0, 0)
    labels0 = outs[0].copy()
    barrier()
    if rank == 0:
        sampler.stop_flag = True
        sampler.join(timeout=2)
    ms_max, = max_over_ranks([ms], dist, dev)
    if rank == 0:
        value = job_throughput(args.steps, world, ms_max)
        bytes_in = MAP_POINTS * 24 + MAP_KF * (W * H * 5)
        line = {"metric": MAP_METRIC, "value": value, "unit": "maps/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, NC), "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "local_maps: %d points, %d keyframes 640x480 per map, device projector, 6-D kernel "
                                       "(xyz*0.5, rgb*4, w=10), 10 iterations, 2 layers; %d distinct maps per rank cycled; "
                                       "%d maps in flight per GPU" % (MAP_POINTS, MAP_KF, MAP_DISTINCT, NC),
                           "l2": "inputs (48 MB cloud per map) exceed nothing by themselves; the working set of a map's mean "
                                 "field (2 M x 20 floats x 3 + lattice) is ~0.6 GB >> L2"},
                "clocks": sampler.summary(),
                "e2e": {"value": value, "unit": "maps/s", "h2d_bytes_per_step": bytes_in, "d2h_bytes_per_step": 2 * MAP_POINTS,
                        "ms_per_step": ms_max / args.steps,
                        "note": "this workload is end to end by construction: every map's cloud and frames come from pinned host "
                                "buffers and its label maps go back, inside the timed region"},
                "gpu_launches": int(launches)}
        if world == 1 and not args.no_cpu:
            import oracle
            oracle.build(ref=False)
            procs = max(1, min(os.cpu_count() or 1, 16))
            cv, cdt, cpu_labels = cpu_map_throughput(procs, 1, first_seed=seeds[0])
            line["cpu_baseline"] = {"value": cv, "unit": "maps/s", "cores": procs, "kind": "port",
                                    "sample": "%d maps, one single-thread oracle pipeline per core (%.1f s compute)" % (procs, cdt)}
            agree = [float((labels0[l] == cpu_labels[l]).mean()) for l in range(2)]
            line["parity_check"] = {"label_agreement": min(agree), "per_layer": agree, "bar": 0.999, "ok": bool(min(agree) >= 0.999),
                                    "what": "labels of local map 0 (device projector, resident posteriors) vs the oracle's map worker"}
        print(json.dumps(line), flush=True)
    for crf in crfs:
        crf.close()
    for c in ctxs:
        c.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def local_map_bench(ctx, synth, NM, wxyz, wrgb, peak, regime):
    """BASELINE configs[3] side measurement: one local map of NM points, both label layers through one 6-D lattice
    (src/segmenter.cpp:629-643), 10 mean-field iterations; device time from the library's events."""
    xyz, col = synth.local_map(seed=5, n_points=NM)
    rng = np.random.default_rng(0)
    U = [rng.random((NM, m), dtype=np.float32) for m in (8, 9)]
    builds = []
    crf = None
    for rep in range(2):  # the first build also pays for every device allocation of a map of this size
        if crf is not None:
            crf.close()
        crf = ctx.crf(NM, [8, 9])
        for l in range(2):
            crf.set_unary(U[l], l)
        t0 = time.perf_counter()
        crf.add_pairwise_xyzrgb(xyz, col, wxyz, wrgb, 10.0)
        builds.append(time.perf_counter() - t0)
    crf.inference(10, unknown=[7, 8], want_Q=False, want_labels=True)
    ms_iter = ctx.timings()["meanfield_ms"] / 10
    V, M, Mp, d = crf.lattice_size(0), 17, 20, 6
    fused, sorted_ = crf.path()
    if fused:
        # fused path over the sorted point order: per iteration one point kernel (unary rows in sorted order + per (point,
        # corner) slice weight, row slot and splat pair + the distinct value rows in and out) and the blur
        ab = (4 * Mp * NM + 14 * (d + 1) * NM + 8 * M * V) + (d + 1) * (8 * M * V + 8 * V) + 4 * M * V
    else:
        # generic path: splat (Q rows + vertex-major CSR in, value rows out), d+1 blur axes (rows in/out + neighbour pairs),
        # slice + soft-max (offsets + weights + norm + unique rows + unary in, Q out)
        ab = (4 * M * NM + 8 * (d + 1) * NM + 4 * NM + 4 * M * V) + (d + 1) * (8 * M * V + 8 * V) + \
             (8 * (d + 1) * NM + 4 * NM + 4 * M * V + 8 * M * NM)
    out = {"points": NM, "labels": 17, "regime": regime, "vertices": V, "ms_per_meanfield_iter": ms_iter,
           "path": ("fused point kernel over the sorted point order" if sorted_ else "fused point kernel") if fused
                   else "generic splat / blur / slice kernels",
           "lattice_build_ms_incl_h2d": 1000.0 * builds[-1], "first_build_ms_incl_allocation": 1000.0 * builds[0],
           "algorithmic_bytes_per_iter": int(ab), "achieved_gbs": ab / (ms_iter / 1000.0) / 1e9,
           "frac_of_hbm_peak": ab / (ms_iter / 1000.0) / 1e9 / peak}
    crf.close()
    return out


def run_gpu(args, rank, world, local_rank):
    import ctypes as C
    import threading as th

    import torch
    import rovinasemanticsegmentation_b200 as rss
    from rovinasemanticsegmentation_b200 import synth

    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if dist is not None:  # NCCL sets itself up lazily: pay for that here, not inside a timed region
        warm = torch.zeros(1, device=dev)
        dist.all_reduce(warm)
        dist.barrier()
        torch.cuda.synchronize()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # keyframes in flight on this GPU: one context + one host thread each (the threads spin in cudaStreamSynchronize, so
    # never more of them than host cores over all ranks)
    NC = max(1, min(args.inflight, (os.cpu_count() or 1) // max(world, 1)))
    # ONE context first: the latency and one-keyframe-in-flight numbers are those of a single synchronous caller (the library
    # sizes its cooperative kernels by the number of live contexts on the device); the other contexts are created afterwards
    ctxs = [rss.Context(rss.DEFAULT_CONFIG, FOREST, local_rank)]
    ctx = ctxs[0]
    lib = ctx._lib
    Kinv, R, t = synth.calibration()
    Kp, Rp, tp = (np.ascontiguousarray(a, np.float32).reshape(-1) for a in (Kinv, R, t))
    frames = []
    for seed in frame_seeds(rank):
        rgb, depth = synth.frame(seed)
        prgb = torch.empty((H, W, 3), dtype=torch.uint8, pin_memory=True)
        pdep = torch.empty((H, W), dtype=torch.int16, pin_memory=True)
        prgb.numpy()[...] = rgb
        pdep.numpy().view(np.uint16)[...] = depth
        frames.append((prgb.numpy(), pdep.numpy().view(np.uint16)))
    outs = [torch.empty((2, H * W), dtype=torch.uint8, pin_memory=True).numpy() for _ in range(NC)]
    prm = rss.KeyframeParams(KF["sigma_xyz"], KF["w_gauss"], KF["sigma_px"], KF["sigma_rgb"], KF["w_bilateral"],
                             KF["iters"], KF["fill"])
    FLUSH_MIB = 160  # > the 126 MB L2 of a B200
    flushes = [torch.empty(FLUSH_MIB << 20, dtype=torch.uint8, device=dev) for _ in range(NC)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(NC)]

    def call(c, rgb, depth, labels):
        st = lib.rss_segment_keyframe(c.h, rss._ptr(rgb, C.c_uint8), rss._ptr(depth, C.c_uint16), W, H,
                                      rss._ptr(Kp, C.c_float), rss._ptr(Rp, C.c_float), rss._ptr(tp, C.c_float),
                                      C.byref(prm), rss._ptr(labels, C.c_uint8), None)
        c._check(st)

    def flush_l2(i):
        if os.environ.get("RSS_BENCH_NOFLUSH"):  # experiments only: the reported numbers always flush
            return
        with torch.cuda.stream(streams[i]):
            flushes[i].zero_()
        streams[i].synchronize()

    # ---- warm-up (untimed): every context, host buffers (allocations, lattice capacities, graph capture)
    def warm():
        for k in range(max(args.warmup, 4)):
            for i, c in enumerate(ctxs):
                call(c, frames[(i + k) % N_FRAMES][0], frames[(i + k) % N_FRAMES][1], outs[i])
        barrier()

    warm()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()

    # ---- (1) latency: ONE keyframe at a time on one context, frame resident, L2 flushed, device time from the library's
    # CUDA events on its own stream; (2) the same steps again with per-kernel event bracketing on (costs a few
    # microseconds per launch, so it is kept out of every headline number) -> per-kernel table and roofline.
    def latency_pass(profile):
        ctx.profile_enable(profile)
        ms, stg = 0.0, {}
        for k in range(args.steps):
            ctx.upload_frame(*frames[k % N_FRAMES])
            flush_l2(0)
            torch.cuda.synchronize()
            call(ctx, None, None, None)
            tm = ctx.timings()
            ms += tm["total_ms"]
            for n, v in tm.items():
                stg[n] = stg.get(n, 0.0) + v
        rep = ctx.profile_report() if profile else {}
        ctx.profile_enable(False)
        return ms, stg, rep

    lat_ms, _, _ = latency_pass(False)       # steady state: the keyframe's device part replays as a CUDA graph
    ctx.keyframe_graph(False)
    lat_ms_eager, stage, _ = latency_pass(False)  # eager launches: the per-stage events only exist here
    ctx.keyframe_graph(True)
    lat_info = [ctx.keyframe_lattice_info(k) for k in range(2)]
    n_samples = int(((frames[0][1][::2, ::2] >= 500) & (frames[0][1][::2, ::2] <= 15000)).sum())
    lat_ms_prof, _, prof = latency_pass(True)

    # ---- (3) throughput: K steps shared by the NC contexts (keyframes are independent units; the reference runs one
    # worker per stage, here one worker per context).  Every step is preceded by a 160 MiB write on the worker's side
    # stream (L2 flush).  resident: the context's frame is already in HBM; e2e: pinned host rgb/depth in, labels out.
    def throughput_pass(host_io, steps=None, nc=None):
        steps = args.steps if steps is None else steps
        nc = NC if nc is None else nc
        per = [steps // nc + (1 if i < steps % nc else 0) for i in range(nc)]
        if not host_io:
            for i, c in enumerate(ctxs[:nc]):
                c.upload_frame(*frames[i % N_FRAMES])
        errs = []

        def worker(i):
            try:
                c = ctxs[i]
                for k in range(per[i]):
                    flush_l2(i)
                    if host_io:
                        fr = frames[(i * 3 + k) % N_FRAMES]
                        call(c, fr[0], fr[1], outs[i])
                    else:
                        call(c, None, None, None)
            except Exception as e:  # noqa: BLE001
                errs.append(e)

        l0 = sum(c.kernel_launches for c in ctxs)
        ths = [th.Thread(target=worker, args=(i,)) for i in range(nc)]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for t_ in ths:
            t_.start()
        for t_ in ths:
            t_.join()
        torch.cuda.synchronize()  # device-wide: every stream of every context
        e1.record()
        e1.synchronize()
        if errs:
            raise errs[0]
        return e0.elapsed_time(e1), sum(c.kernel_launches for c in ctxs) - l0

    R = max(1, args.repeats)
    # one keyframe in flight per GPU (no overlap between keyframes: what a single synchronous caller gets)
    throughput_pass(False, steps=2, nc=1)
    res1_all = [throughput_pass(False, nc=1)[0] for _ in range(R)]
    e2e1_all = [throughput_pass(True, nc=1)[0] for _ in range(R)]
    # ---- now the other contexts: NC keyframes in flight
    ctxs.extend(rss.Context(rss.DEFAULT_CONFIG, FOREST, local_rank) for _ in range(NC - 1))
    warm()
    throughput_pass(False, steps=2 * NC)  # untimed: worker threads, side streams and flush buffers warm
    # The timed region of K keyframes is short (tens of ms), so it is REPEATED `repeats` times back to back; every repeat
    # times exactly K steps, the slowest rank decides per repeat, and the MEDIAN repeat is reported (min / max beside it).
    res_all, e2e_all, launches = [], [], 0
    for _ in range(R):
        ms, launches = throughput_pass(False)
        res_all.append(ms)
    for _ in range(R):
        e2e_all.append(throughput_pass(True)[0])
    # parity of the timed path: the label maps the e2e entry point writes for this rank's first frame (checked on rank 0
    # against the CPU arm's label maps of the same keyframe, below)
    call(ctx, frames[0][0], frames[0][1], outs[0])
    gpu_labels0 = outs[0].copy()
    barrier()
    if rank == 0:
        sampler.stop_flag = True
        sampler.join(timeout=2)

    # slowest rank decides (per repeat)
    allv = max_over_ranks(res_all + e2e_all + res1_all + e2e1_all + [lat_ms], dist, dev)
    res_all, e2e_all, res1_all, e2e1_all = (allv[k * R:(k + 1) * R] for k in range(4))
    lat_ms_max = allv[4 * R]
    med = lambda v: sorted(v)[len(v) // 2]
    res_ms_max, e2e_ms_max = med(res_all), med(e2e_all)

    if rank == 0:
        value = job_throughput(args.steps, world, res_ms_max)
        e2e = job_throughput(args.steps, world, e2e_ms_max)
        spread = lambda v: {"median": job_throughput(args.steps, world, med(v)), "min": job_throughput(args.steps, world, max(v)),
                            "max": job_throughput(args.steps, world, min(v)), "repeats": len(v)}
        peak, peak_src = peaks()
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
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")  # dram bytes per launch from `ncu --set full` captures
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            # exact name first; otherwise the instantiation of that kernel that was captured most often, i.e. the variant
            # that runs once per mean-field iteration (not the first / last pass variants)
            cand = [(k, v) for k, v in tj.items() if k == dom_name] or \
                   [(k, v) for k, v in tj.items() if k.split("<")[0] == dom_name.split("<")[0]]
            if cand:
                tk, tv = max(cand, key=lambda kv: kv[1].get("launches_profiled", 0))
                traffic, traffic_src = tv.get("dram_bytes_per_launch"), tk
        roof = {"bound": "hbm", "kernel": dom_name, "achieved": ach, "peak": peak, "unit": "GB/s",
                "frac": ach / peak if ach is not None else None, "traffic": traffic,
                "traffic_kernel": traffic_src if traffic is not None else None,
                "traffic_ratio": (traffic / ab) if (traffic is not None and ab) else None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": int(ab) if ab is not None else None,
                "us_per_launch": 1000.0 * dom_ms / dom_cnt,
                "lattices": [{"d": d, "vertices": V} for d, V in lat_info]}
        step_bytes = sum((algo_bytes(n, env) or 0) * cnt / args.steps for n, (ms, cnt) in top)
        roof["step_algorithmic_bytes"] = int(step_bytes)
        roof["step_achieved"] = step_bytes / (res_ms_max / args.steps / 1000.0) / 1e9
        roof["step_frac"] = roof["step_achieved"] / peak
        line = {
            "metric": METRIC, "value": value, "unit": "keyframes/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": res_ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "l2": "every step is preceded by a %d MiB device write (L2 flush; L2 = 126 MB)" % FLUSH_MIB,
                       "frames": "%d distinct synthetic frames per rank" % N_FRAMES,
                       "inflight": "%d keyframes in flight per GPU (one context + host thread each)" % NC,
                       "repeats": "the K-step timed region is repeated %d times; value / e2e are the median repeat" % R},
            "clocks": sampler.summary(),
            "e2e": {"value": e2e, "unit": "keyframes/s", "h2d_bytes_per_step": W * H * 3 + W * H * 2,
                    "d2h_bytes_per_step": 2 * W * H, "ms_per_step": e2e_ms_max / args.steps},
            "spread": {"value": spread(res_all), "e2e": spread(e2e_all)},
            "inflight_1": {"value": spread(res1_all), "e2e": spread(e2e1_all),
                           "note": "one keyframe in flight per GPU: a single synchronous caller, no overlap between keyframes"},
            "gpu_launches": int(launches),
            "roofline": roof,
            "latency": {"ms_per_keyframe": lat_ms_max / args.steps,
                        "ms_per_keyframe_eager": lat_ms_eager / args.steps,
                        "note": "one keyframe at a time on one context, frame resident, L2 flushed, library CUDA events; "
                                "ms_per_keyframe = CUDA-graph replay (steady state), stages_ms from an eager pass",
                        "stages_ms": {n: v / args.steps for n, v in stage.items()},
                        "ms_per_meanfield_iter": stage.get("meanfield_ms", 0.0) / args.steps / KF["iters"]},
            "kernels": kernel_table,
            "kernels_note": "per-kernel device times from a second single-context pass with event bracketing on "
                            "(%.3f ms/keyframe vs %.3f ms/keyframe without)" % (lat_ms_prof / args.steps, lat_ms / args.steps),
        }
        if world == 1 and not args.quick:
            line["local_map"] = local_map_bench(ctx, synth, 2_000_000, 0.5, 4.0, peak, "node scales (xyz*0.5, rgb*4)")
        if world == 1 and not args.quick and not args.no_15m:
            # BASELINE configs[3]: 50 keyframes fused into one ~15 M-point map, both lattice regimes
            line["local_map_15m"] = [
                local_map_bench(ctx, synth, 15_000_000, 0.5, 4.0, peak, "node scales (xyz*0.5, rgb*4): few vertices"),
                local_map_bench(ctx, synth, 15_000_000, 20.0, 40.0, peak, "fine scales (xyz*20, rgb*40): millions of vertices")]
        if world == 1 and not args.quick:
            line["forest_train"] = forest_train_bench(ctx, synth, not args.no_cpu)
        if world == 1 and not args.no_cpu and not args.quick:
            import oracle
            oracle.build(ref=False)
            procs = max(1, min(os.cpu_count() or 1, 32))
            cv, cdt, cpu_labels0 = cpu_throughput(procs, 1, 1, first_seed=frame_seeds(0)[0])
            line["cpu_baseline"] = {"value": cv, "unit": "keyframes/s", "cores": procs, "kind": "port",
                                    "sample": "%d full 640x480 keyframes, one single-thread oracle pipeline per core "
                                              "(%.1f s compute)" % (procs, cdt)}
            # the label maps the timed e2e entry point wrote for frame 0 vs the CPU arm's label maps of the same keyframe
            agree = [float((gpu_labels0[l] == cpu_labels0[l]).mean()) for l in range(2)]
            line["parity_check"] = {"label_agreement": min(agree), "per_layer": agree, "frame_seed": frame_seeds(0)[0],
                                    "bar": 0.999, "ok": bool(min(agree) >= 0.999),
                                    "what": "rss_segment_keyframe (host buffers, the e2e path) vs oracle.keyframe on the same 640x480 frame"}
    # the C++ multi-GPU host (one process, all N GPUs) runs while the ranks are idle: every rank waits here, rank 0 runs it
    for c in ctxs:
        c.close()
    if dist is not None:
        dist.barrier()
    if rank == 0:
        if not args.quick or world > 1:
            line["cpp_worker"] = cpp_worker_bench(world, NC, 1500 * world)
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def cpp_worker_bench(n_gpus, inflight, frames):
    """The C++ host path over the same C ABI (host/keyframe_worker.cpp: one thread per (GPU, keyframe in flight), pinned host
    buffers in, label maps out, no Python): its own throughput figure next to the torchrun ranks, at every N."""
    from rovinasemanticsegmentation_b200 import build as rss_build
    import rovinasemanticsegmentation_b200 as rss
    try:
        exe = rss_build.build_host()
        r = subprocess.run([exe, "--config", rss.DEFAULT_CONFIG, "--forest", FOREST, "--gpus", str(n_gpus), "--inflight",
                            str(inflight), "--frames", str(frames)], capture_output=True, text=True, timeout=300)
        if r.returncode != 0:
            return {"error": (r.stderr or "rc=%d" % r.returncode).strip()[-200:]}
        out = json.loads(r.stdout.strip().splitlines()[-1])
        out["note"] = "wall clock around %d keyframes, end to end with host buffers, no L2 flush between keyframes" % frames
        return out
    except Exception as e:  # a missing g++ or binary is not a reason to lose the bench line
        return {"error": str(e)[-200:]}


def forest_train_bench(ctx, synth, with_cpu):
    """Side measurement of SURVEY 8(f) rank 2: training the 4-tree forest of the benchmark (src/train.cpp:225-249 set-up) on the
    features of synthetic frames at training_sample_stride 5 - rss_forest_train on the GPU next to the unmodified reference
    learner (oracle/_ref, all host cores) on the same DataStorage."""
    import tempfile
    Kinv, R, t = synth.calibration(W, H)
    fs = []
    for seed in range(900, 908):
        rgb, depth = synth.frame(seed, W, H)
        fs.append(ctx.extract_features(rgb, depth, Kinv, R, t, 5, 0.5, 15.0)[0].copy())
    feats = np.concatenate(fs)
    labels = synth.labels_from_features(feats, synth.label_thresholds(feats))
    cc = [int(labels[:, 0].max()) + 1, int(labels[:, 1].max()) + 1]
    out = {"samples": int(feats.shape[0]), "features": int(feats.shape[1]), "trees": 4, "max_depth": 30, "min_split": 50}
    with tempfile.TemporaryDirectory() as tmp:
        best = None
        for _ in range(3):
            st = ctx.forest_train(feats, labels, cc, os.path.join(tmp, "gpu.dat"), num_trees=4, max_depth=30,
                                  min_split_examples=50, seed=1)
            best = st.train_ms if best is None else min(best, st.train_ms)
        out.update({"gpu_ms": best, "nodes": int(st.nodes), "levels": int(st.levels), "features_per_node": int(st.features_per_node),
                    "note": "rss_forest_train incl. the H2D copy of the feature matrix and the host-side leaf histograms, best of 3"})
        if with_cpu:
            import oracle
            oracle.build(ref=True)
            if oracle.ref_available():
                threads = max(1, min(os.cpu_count() or 1, 4))  # RandomForestLearner::learn: one OpenMP task per tree
                t0 = time.perf_counter()
                oracle.ref_forest_train(feats, labels, os.path.join(tmp, "ref.dat"), num_trees=4, max_depth=30, min_split=50,
                                        threads=threads)
                out["cpu_reference_ms"] = 1000.0 * (time.perf_counter() - t0)
                out["cpu_threads"] = threads
                out["cpu_kind"] = "reference (third-party/libforest/src/learning.cpp compiled in place, oracle/_ref)"
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="keyframes", choices=["keyframes", "local_maps"],
                    help="keyframes: the headline metric (BASELINE configs[1]+[2]); local_maps: the configs[4] sweep")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--quick", action="store_true", help="skip the keyframes-in-flight sweep and the cpu_baseline leg (profiling runs)")
    ap.add_argument("--no-15m", dest="no_15m", action="store_true", help="skip the 15 M-point local-map side measurement")
    ap.add_argument("--repeats", type=int, default=5, help="repeats of the K-step timed region (median reported)")
    ap.add_argument("--inflight", type=int, default=4,
                    help="keyframes in flight per GPU (one context + host thread each)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if args.workload == "local_maps":
            run_local_maps_reference(args, rank, world)
        else:
            run_reference(args, rank, world)
        return
    if world == 1 and args.gpus > 1:
        # plain `python bench.py --gpus N`: re-launch under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__), "--gpus", str(args.gpus),
               "--steps", str(args.steps), "--warmup", str(args.warmup), "--workload", args.workload] + \
              (["--no-cpu"] if args.no_cpu else [])
        sys.exit(subprocess.call(cmd))
    if args.workload == "local_maps":
        run_local_maps(args, rank, world, local_rank)
        return
    run_gpu(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
