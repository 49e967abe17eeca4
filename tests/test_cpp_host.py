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


def _checksum(frames):
    s = 0
    for f in frames:
        for v in f.ravel().tolist():
            s = (s * 31 + v) & 0xFFFFFFFFFFFFFFFF
    return s


def test_luma_file_source(tmp_path):
    """LumaFileSource: raw luma ('WxH@FPS:path', the Rust shim's input string) and YUV4MPEG2 with 4:2:0 / 4:4:4 / mono."""
    import numpy as np
    exe = str(tmp_path / "luma_source_check")
    libdir = os.path.join(ROOT, "ofps_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "luma_source_check.cpp"), "-o", exe, "-L", libdir, "-lofps_b200",
                           f"-Wl,-rpath,{libdir}"])
    rng = np.random.default_rng(1)
    w, h = 37, 22
    frames = [rng.integers(0, 256, (h, w), dtype=np.uint8) for _ in range(3)]
    raw = tmp_path / "f.y"
    raw.write_bytes(b"".join(f.tobytes() for f in frames))
    out = subprocess.run([exe, f"{w}x{h}@29.97:{raw}"], capture_output=True, text=True).stdout.split()
    assert out[:2] == [str(w), str(h)] and abs(float(out[2]) - 29.97) < 1e-6 and int(out[3]) == 3 and int(out[4]) == _checksum(frames)
    for tag, cw, ch in (("C420jpeg", 2, 2), ("C444", 1, 1), ("Cmono", 0, 0), ("", 2, 2)):
        p = tmp_path / f"clip_{tag or 'default'}.y4m"
        with open(p, "wb") as f:
            f.write(f"YUV4MPEG2 W{w} H{h} F30000:1001 Ip A1:1 {tag}".rstrip().encode() + b"\n")
            for fr in frames:
                f.write(b"FRAME\n" + fr.tobytes())
                if cw:
                    f.write(bytes(2 * ((w + cw - 1) // cw) * ((h + ch - 1) // ch)))
        out = subprocess.run([exe, str(p)], capture_output=True, text=True).stdout.split()
        assert out[:2] == [str(w), str(h)] and abs(float(out[2]) - 30000 / 1001) < 1e-3, (tag, out)
        assert int(out[3]) == 3 and int(out[4]) == _checksum(frames), (tag, out)
    bad = subprocess.run([exe, str(tmp_path / "missing.y4m")], capture_output=True, text=True)
    assert bad.returncode == 1 and "ERROR -5" in bad.stdout


def _build_motion_extract(tmp):
    exe = os.path.join(tmp, "motion_extract")
    libdir = os.path.join(ROOT, "ofps_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tools", "motion_extract.cpp"), "-o", exe, "-L", libdir, "-lofps_b200",
                           f"-Wl,-rpath,{libdir}"])
    return exe


def test_motion_extract_compiles_and_refuses_cpu(tmp_path):
    import numpy as np
    import torch
    exe = _build_motion_extract(str(tmp_path))
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by test_motion_extract_matches_oracle")
    raw = tmp_path / "f.y"
    raw.write_bytes(np.zeros((2, 32, 48), np.uint8).tobytes())
    p = subprocess.run([exe, f"48x32@30:{raw}", str(tmp_path / "o.mvec")], capture_output=True, text=True)
    assert p.returncode == 3 and "no CPU fallback" in p.stderr


@pytest.mark.gpu
def test_motion_extract_matches_oracle(tmp_path, oracle):
    """motion_extract (the reference's motion-extract for the B200 decoder): raw luma stream -> .mvec, every frame's
    entries equal to the oracle's block matcher on the same pair."""
    import numpy as np
    from ofps_b200 import capi, synth
    exe = _build_motion_extract(str(tmp_path))
    frames = synth.make_stream(4, 640, 360, 8)
    raw = tmp_path / "f.y"
    raw.write_bytes(frames.tobytes())
    out = str(tmp_path / "o.mvec")
    p = subprocess.run([exe, f"640x360@30:{raw}", out, "16", "8"], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "3 frames" in p.stdout
    for i in range(3):
        want = oracle.block_match(frames[i], frames[i + 1], 16, 8, 0)[2]
        assert capi.mvec_read(out, i).tobytes() == want.tobytes(), i


def test_flow_extract_compiles_and_refuses_cpu(tmp_path):
    """tools/flow_extract.cpp (the reference's flow-extract on the C ABI) builds; without a GPU it stops at the context."""
    import torch
    exe = str(tmp_path / "flow_extract")
    libdir = os.path.join(ROOT, "ofps_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tools", "flow_extract.cpp"), "-o", exe, "-L", libdir, "-lofps_b200",
                           f"-Wl,-rpath,{libdir}"])
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    p = subprocess.run([exe, str(tmp_path / "x.mvec"), str(tmp_path / "out"), "64", "48"], capture_output=True, text=True)
    assert p.returncode == 3 and "no CPU fallback" in p.stderr
