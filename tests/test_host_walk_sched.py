"""The issue schedule of the row-walk kernel (nhans_b200/csrc/walk_sched.h) is plain C++ shared between the
device code and this host simulation: a software model of the TMEM slot ring replays every step and checks the
products each output row receives, the accumulate flags, ring wrap handling and deadlock freedom."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")
def test_walk_schedule_simulation(tmp_path):
    exe = str(tmp_path / "walk_sched_check")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-Wall", "-Werror", "-o", exe,
                           os.path.join(ROOT, "tests", "host", "walk_sched_check.cc")])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "walk schedule ok" in out.stdout
