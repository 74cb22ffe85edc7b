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


def test_parse_config_host_only(lib, tmp_path):
    """rss_parse_config needs no device: the repo's own config, a missing key (reference message, config.h:13-24), and -
    where the reference tree is mounted - the reference's resources/config.json fed through UNCHANGED
    (src/segmenter.cpp:70-127: two layers, 8 + 9 non-negative labels, "Unknown" = 7 / 8)."""
    import json
    info = rss.parse_config(CONFIG)
    assert (info.patch_size, info.patch_size_reduce, info.feature_length) == (77, 11, 366)
    assert info.layer_count == 2 and list(info.class_counts)[:2] == [8, 9] and list(info.unknown_label)[:2] == [7, 8]
    cfg = json.load(open(CONFIG))
    del cfg["feature_normal"]
    bad = tmp_path / "bad.json"
    bad.write_text(json.dumps(cfg))
    with pytest.raises(rss.RssError) as e:
        rss.parse_config(str(bad))
    assert e.value.status == 3 and "The key: 'feature_normal' was not found in the config file." in str(e.value)
    ref = "/root/reference/resources/config.json"
    if os.path.exists(ref):
        r = rss.parse_config(ref)
        assert (r.feature_color_patch, r.feature_depth, r.feature_height, r.feature_normal) == (1, 1, 1, 1)
        assert (r.patch_size, r.patch_size_reduce, r.feature_length) == (77, 11, 366)
        assert r.layer_count == 2 and list(r.class_counts)[:2] == [8, 9] and r.total_classes == 17
        assert list(r.unknown_label)[:2] == [7, 8]
        assert (r.use_dense_crf, r.dcrf_iterations, r.rf_prediction_stride) == (0, 10, 2)
        assert (r.dcrf_xyz_kernel, r.dcrf_rgb_kernel, r.dcrf_kernel_weight) == (0.5, 4.0, 10.0)
        assert (r.depth_min, r.depth_max) == (0.5, 15.0)


def test_corrupt_forest_is_an_error_not_a_crash(lib):
    """node counts are bounded by the bytes left in the file BEFORE anything is allocated, and no exception
    crosses the C boundary.  (Loading needs a context, hence a GPU; on the CPU box only the export is checked.)"""
    assert hasattr(lib, "rss_load_forest_memory")
