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
