"""Builds librss.so (hand-written CUDA for sm_100a + the C ABI) in-tree with nvcc.

    python -m rovinasemanticsegmentation_b200.build [--force]

nvcc cross-compiles without a GPU.  -fmad=false: the parity-critical kernels must round every float
multiply and add separately (threshold compares in the forest, lattice embedding, OpenCV bilinear).
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "librss.so")
SOURCES = ["api.cu", "features.cu", "normals.cu", "forest.cu", "lattice.cu", "meanfield.cu", "meanfield_shared.cu", "crf.cu", "train.cu", "sort.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-fmad=false",
         "-Xcompiler", "-fPIC,-ffp-contract=off,-Wall,-Wno-unused-function", "--expt-relaxed-constexpr"]


def _newer(a, b):
    return not os.path.exists(b) or os.path.getmtime(a) > os.path.getmtime(b)


def build(force=False, verbose=False):
    extra = os.environ.get("RSS_NVCC_DEFS", "").split()  # e.g. "-DRSS_TILE_MINB=2" for tuning sweeps
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "rss.h")]
    objs = []
    jobs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(CSRC, src.replace(".cu", ".o"))
        objs.append(o)
        if force or any(_newer(d, o) for d in deps):
            jobs.append([NVCC] + FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o])
    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0 or verbose:
            sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for " + cmd[-3])
    with ThreadPoolExecutor(max_workers=6) as ex:
        list(ex.map(run, jobs))
    if jobs or not os.path.exists(OUT):
        # (the arch on the link line too: nvcc's device-link stub otherwise targets its default, sm_52)
        run([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", OUT] + objs + ["-lcudart"])
    return OUT


HOST_BIN = os.path.join(HERE, "host", "keyframe_worker")


def build_host(force=False):
    """C++ host programs over the C ABI (g++, no CUDA headers needed): the multi-GPU keyframe worker."""
    src = os.path.join(HERE, "host", "keyframe_worker.cpp")
    hdr = os.path.join(HERE, "host", "rss_adapters.hpp")
    if force or _newer(src, HOST_BIN) or _newer(hdr, HOST_BIN) or _newer(OUT, HOST_BIN):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-o", HOST_BIN, src, "-L" + HERE, "-lrss",
                               "-Wl,-rpath,$ORIGIN/..", "-lpthread"])
    return HOST_BIN


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_host(force="--force" in sys.argv))
