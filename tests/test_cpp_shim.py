"""include/gsfm_rotation_estimator.hpp: the C++ mirror of theia::RotationEstimator / GSfMNonlinearRotationEstimator
(rotation_estimator.h:50-66, GSfM_nonlinear_rotation_estimator.hpp:22-59) over the C ABI, compiled with g++ against
stand-in container types shaped like Theia's (tests/cpp/test_shim.cc) and run as a separate process."""
import os
import subprocess

import pytest

from globalsfmpy_b200 import _capi as capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "globalsfmpy_b200", "csrc")


def _build(tmp_path):
    exe = str(tmp_path / "test_shim")
    cmd = ["/usr/bin/g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-o", exe, os.path.join(ROOT, "tests", "cpp", "test_shim.cc"),
           "-L" + CSRC, "-lgsfm_ra", "-Wl,-rpath," + CSRC]
    subprocess.check_call(cmd)
    return exe


def test_shim_compiles_and_fails_loudly_without_device(tmp_path):
    capi.lib()  # built by conftest
    exe = _build(tmp_path)
    if capi.lib().gsfm_ra_device_count() > 0:
        pytest.skip("a GPU is visible: the full run is test_shim_on_gpu")
    out = subprocess.run([exe, "--expect-no-device"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "FAIL" not in out.stdout


@pytest.mark.gpu
def test_shim_on_gpu(tmp_path):
    capi.lib()
    exe = _build(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    print(out.stdout)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "ALL OK" in out.stdout
