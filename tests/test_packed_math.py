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


def test_alf_dot_product_layout_matches_direct_sum(tmp_path):
    """csrc/ilf_alf_tab.cuh: slot numbering, coefficient bytes, high parts and the 5x5-in-7x7 embedding (host model of IDP.2A)."""
    exe = str(tmp_path / "alf_tab_host_test")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "vvcsoftware_vtm_b200", "csrc"),
                           os.path.join(ROOT, "tests", "host", "alf_tab_host_test.cpp"), "-o", exe])
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:]
