"""The C++ mirror of the reference's plugin interface (include/ofps_b200.hpp) compiles against the C ABI;
without a GPU the context constructor throws (no CPU fallback), with one the Decoder -> Detector ->
Estimator chain runs (tests/cpp_host_check.cpp)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp):
    from ofps_b200 import capi
    capi.lib()
    exe = os.path.join(tmp, "cpp_host_check")
    libdir = os.path.join(ROOT, "ofps_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp_host_check.cpp"), "-o", exe, "-L", libdir, "-lofps_b200",
                           f"-Wl,-rpath,{libdir}"])
    return exe


def test_cpp_host_compiles_and_refuses_cpu(tmp_path):
    import torch
    exe = _build(str(tmp_path))
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by test_cpp_host_runs")
    p = subprocess.run([exe], capture_output=True, text=True)
    assert p.returncode == 3 and "NO_DEVICE" in p.stdout and "no CPU fallback" in p.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["lsq", "ransac"])
def test_cpp_host_runs(tmp_path, mode):
    exe = _build(str(tmp_path))
    p = subprocess.run([exe, mode], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "OK frames=2 detections=2" in p.stdout and "DENSE OK" in p.stdout
