"""include/vsm_cv.hpp -- the reference-shaped C++ call sites (cv::Mat in, std::vector<cv::DMatch>
out) -- compiled against libvsm.so.  The compile is checked everywhere; the run needs a B200."""
import os
import subprocess

import pytest

from conftest import ROOT
import vsm_b200
from oracle import oracle


def build(tmp):
    exe = os.path.join(tmp, "test_adaptor")
    libdir = os.path.dirname(vsm_b200.lib_path())
    odir = os.path.dirname(oracle.build())
    cmd = ["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "test_adaptor.cpp"),
           "-o", exe, "-L", libdir, "-lvsm", "-L", odir, "-lvsm_oracle", f"-Wl,-rpath,{libdir}", f"-Wl,-rpath,{odir}",
           "-lpthread", "-ldl", "-lm"]
    subprocess.check_call(cmd)
    return exe


def test_adaptor_compiles_and_links(tmp_path):
    assert os.path.exists(build(str(tmp_path)))


@pytest.mark.gpu
def test_adaptor_matches_oracle(tmp_path):
    exe = build(str(tmp_path))
    import torch
    n = torch.cuda.device_count()
    devices = [str(d) for d in range(min(n, 8))] if n >= 2 else ["0", "0", "0"]     # one GPU: three contexts on it
    out = subprocess.run([exe] + devices, capture_output=True, text=True, timeout=180)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "adaptor test: OK" in out.stdout and "multi-device matcher" in out.stdout
