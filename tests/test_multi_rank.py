"""CPU tests of the N > 1 host logic of bench.py with a world-size-2 gloo process group: disjoint per-rank keyframe
shards, max-over-ranks timing reduction, whole-job throughput, and the reference arm's rank-0-only behaviour."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

WORKER = r'''
import os, sys, json
sys.path.insert(0, %r)
import torch, torch.distributed as dist
import bench
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
seeds = bench.frame_seeds(rank)
gathered = [None] * world
dist.all_gather_object(gathered, seeds)
t = bench.max_over_ranks([10.0 + rank, 5.0 - rank, 1.0], dist)
dist.barrier()
if rank == 0:
    print(json.dumps({"seeds": gathered, "max": t, "value": bench.job_throughput(20, world, t[0])}))
dist.destroy_process_group()
'''


def test_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
    out = subprocess.check_output([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                                   "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)], env=env, text=True,
                                  stderr=subprocess.DEVNULL, timeout=300)
    line = json.loads([l for l in out.splitlines() if l.startswith("{")][-1])
    s0, s1 = line["seeds"]
    assert len(s0) == len(s1) == 8 and not set(s0) & set(s1)  # disjoint shards
    assert line["max"] == [11.0, 5.0, 1.0]                    # slowest rank decides, element-wise
    assert abs(line["value"] - 2 * 20 / 0.011) < 1e-6         # whole-job aggregate over both ranks


def test_reference_arm_runs_on_rank0_only():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                                   "--steps", "1"], env=env, text=True, timeout=120)
    assert out.strip() == ""  # ranks > 0 exit 0 without work


def test_algorithmic_bytes_table():
    sys.path.insert(0, ROOT)
    import bench
    env = {"N": 640 * 480, "M": 17, "Ns": 70000, "D": 366, "T": 4, "nodes": 16000, "leaves": 8000, "P": 77,
           "lat": [(3, 40000), (5, 11000)]}
    b = bench.algo_bytes("meanfield_point_kernel", env)
    # one mean-field iteration of this CRF moves on the order of 80 MB
    assert 60e6 < b < 110e6
    assert bench.algo_bytes("patch_features_kernel", env) > 4 * 363 * 70000
    assert bench.algo_bytes("no_such_kernel", env) is None


def test_every_profiled_kernel_has_algorithmic_bytes():
    """Every kernel of the library in the committed ncu launch list of the bench (profiles/r02z_launches.csv) has an entry in
    bench.algo_bytes - a kernel without one reads 0 % in the per-kernel roofline table."""
    import csv
    import re
    sys.path.insert(0, ROOT)
    import bench
    env = {"N": 640 * 480, "M": 17, "Ns": 70000, "D": 366, "T": 4, "nodes": 16000, "leaves": 8000, "P": 77,
           "lat": [(3, 40000), (5, 11000)]}
    lines = [l for l in open(os.path.join(ROOT, "profiles", "r02z_launches.csv")) if not l.startswith("==")]
    names = set()
    for row in csv.DictReader(lines):
        if "rss::" not in row["Kernel Name"]:
            continue  # the bench's own L2-flush fill kernel (torch)
        n = re.sub(r"^(?:\w+::)+", "", row["Kernel Name"].split("(")[0].replace("void ", "").strip())
        # the library's profile table names the three variants of the point kernel by their role
        m = re.match(r"meanfield_point_kernel<\d+, \d+, \d+, (\d+)>", n)
        if m:
            n = {"2": "meanfield_point_kernel<first>", "3": "meanfield_point_kernel", "5": "meanfield_point_kernel<last>"}[m.group(1)]
        names.add(n)
    assert len(names) > 25
    missing = [n for n in sorted(names) if not bench.algo_bytes(n, env)]
    assert not missing, missing
