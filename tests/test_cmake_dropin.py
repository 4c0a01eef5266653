"""The CMake option INTEGRATION.md documents, exercised: a scratch copy of the reference gets the ILF_B200 option and the five
`#if !ILF_B200` guards (tools/apply_dropin_patch.py), is configured with -DILF_B200=ON and built (DecoderApp, EncoderApp); the
binaries must be linked against libilf_b200.so and reach it.  Build container only (needs /root/reference and cmake)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not (os.path.isdir("/root/reference/source") and shutil.which("cmake") and os.path.exists(os.path.join(ROOT, "vvcsoftware_vtm_b200", "libilf_b200.so"))),
                    reason="needs /root/reference, cmake and the built library")
def test_reference_builds_with_the_documented_cmake_option(tmp_path):
    r = subprocess.run([os.path.join(ROOT, "tools", "cmake_dropin_build.sh"), str(tmp_path / "w")], capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
    assert "cmake drop-in build OK" in r.stdout
