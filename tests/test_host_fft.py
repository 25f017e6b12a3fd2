"""The FFT-400 butterflies the STFT / iSTFT kernels use (nhans_b200/csrc/fft400.cuh) are __host__ __device__;
this compiles them for the host with nvcc and checks them against a naive double-precision DFT."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc not available")
def test_fft400_host_check(tmp_path):
    exe = str(tmp_path / "fft_check")
    subprocess.check_call(["nvcc", "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-o", exe,
                           os.path.join(ROOT, "tests", "host", "fft_check.cu")])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
