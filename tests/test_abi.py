"""CPU tests of the drop-in boundary: librss.so builds for sm_100a, loads without a GPU and exports every symbol that
include/rss.h declares; without a CUDA device the product fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest

import rovinasemanticsegmentation_b200 as rss
from conftest import CONFIG, FOREST


@pytest.fixture(scope="module")
def lib():
    from rovinasemanticsegmentation_b200 import build
    build.build()
    return rss.load_library()


def test_every_declared_symbol_is_exported(lib):
    text = open(rss.HEADER_PATH).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = set(re.findall(r"\b(rss_[a-z0-9_]+)\s*\(", text))
    assert len(names) >= 20
    for n in sorted(names):
        assert hasattr(lib, n), "librss.so does not export " + n


def test_no_cpu_fallback_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(rss.RssError) as e:
        rss.Context(CONFIG, FOREST, 0)
    assert e.value.status == 5  # RSS_ERR_CUDA
    assert "no CPU fallback" in str(e.value)


def test_product_never_imports_the_oracle():
    root = os.path.dirname(rss.__file__)
    for dp, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp")):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "liboracle" not in src and "oracle.h" not in src, f


def test_status_strings(lib):
    lib.rss_status_string.restype = ctypes.c_char_p
    assert lib.rss_status_string(0) == b"ok"
    assert b"CUDA" in lib.rss_status_string(5)


def test_cpp_host_side_builds_without_cuda_headers(lib, tmp_path):
    """The C++ adapters and the multi-GPU worker compile with plain g++ against include/rss.h (no CUDA headers) and
    link against librss.so; without a GPU the worker fails loudly instead of falling back."""
    import subprocess
    from rovinasemanticsegmentation_b200 import build
    root = os.path.dirname(os.path.dirname(rss.__file__))
    exe = build.build_host(force=True)
    assert os.path.exists(exe)
    chk = str(tmp_path / "adapters_check")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-o", chk,
                           os.path.join(root, "tests", "cpp", "adapters_check.cpp"), "-L" + os.path.dirname(rss.LIB_PATH),
                           "-lrss", "-Wl,-rpath," + os.path.dirname(rss.LIB_PATH)])
    import torch
    if not torch.cuda.is_available():
        r = subprocess.run([exe, "--config", CONFIG, "--forest", FOREST, "--frames", "2"], capture_output=True, text=True)
        assert r.returncode != 0 and "keyframe_worker:" in r.stderr
