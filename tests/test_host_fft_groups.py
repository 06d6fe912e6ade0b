"""Host logic of the frame-resident FFT scan (csrc/fft_frame.cuh): the group arithmetic that makes the block sums
independent of how a stream is split into calls is plain integer code shared by host and device, so it is checked on the
CPU - `tests/host/fft_groups_check.cu` is compiled with nvcc (host code only runs) and walks ~7e5 groups."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fft_frame_group_arithmetic(tmp_path):
    nvcc = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "fft_groups_check")
    subprocess.check_call([nvcc, "-std=c++17", "-O1", "-gencode", "arch=compute_100a,code=sm_100a",
                           "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "radiocapture_rf_b200", "csrc"),
                           "-o", exe, os.path.join(ROOT, "tests", "host", "fft_groups_check.cu")])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.startswith("ok ")
