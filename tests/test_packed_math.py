"""The two-samples-per-register arithmetic of csrc/ilf_packed.cuh, compiled as host C++ and checked against scalar code."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_packed_helpers_match_scalar(tmp_path):
    exe = str(tmp_path / "packed_host_test")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "vvcsoftware_vtm_b200", "csrc"),
                           os.path.join(ROOT, "tests", "host", "packed_host_test.cpp"), "-o", exe])
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:]
