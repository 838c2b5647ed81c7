"""The C-ABI library loads on a CPU-only box and exports every symbol include/vsm.h declares.
No compute call is made here (that needs a B200)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
import vsm_b200
from vsm_b200 import matcher


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "vsm.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vsm_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    lib = ctypes.CDLL(vsm_b200.lib_path())
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/vsm.h but not exported by libvsm.so"


def test_binding_covers_header():
    assert sorted(matcher.SYMBOLS) == declared_symbols()


def test_dmatch_layout_is_cv_dmatch():
    # cv::DMatch {int queryIdx; int trainIdx; int imgIdx; float distance;} = 16 bytes
    assert vsm_b200.DMATCH.itemsize == 16
    assert [vsm_b200.DMATCH.fields[k][1] for k in ("queryIdx", "trainIdx", "imgIdx", "distance")] == [0, 4, 8, 12]


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(vsm_b200.VsmError):
        vsm_b200.Matcher()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "visual-slam-pipeline_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower(), f"{f} mentions the oracle"
