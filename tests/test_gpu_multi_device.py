"""One process driving several GPUs through nothing but the C ABI (tests/cpp/test_multi_device.cpp): cqb_init_multi,
cqb_bases_register_sharded, host-pointer and resident-scalar MSMs over a point-range-sharded SRS, against the CPU oracle.
On a one-GPU box the same program runs with one device slot (the sharded calls degrade to the plain ones)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "sha2-on-cq-halo2_b200")


def _build(tmp):
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s", "all"])
    exe = os.path.join(tmp, "test_multi_device")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "cpp", "test_multi_device.cpp"),
                           "-L" + PKG, "-lcqb200", "-L" + os.path.join(ROOT, "oracle", "_build"), "-loracle",
                           "-Wl,-rpath," + PKG, "-Wl,-rpath," + os.path.join(ROOT, "oracle", "_build")])
    return exe


def test_multi_device_program_links(tmp_path):
    """CPU: the program compiles against include/cqb200.h and links (every multi-device symbol is exported)"""
    if not os.path.exists(os.path.join(PKG, "libcqb200.so")):
        pytest.skip("libcqb200.so not built")
    assert os.path.exists(_build(str(tmp_path)))


@pytest.mark.gpu
@pytest.mark.parametrize("devices,log_n", [(1, 16), (2, 18), (8, 21)])
def test_multi_device_msm(tmp_path, devices, log_n):
    exe = _build(str(tmp_path))
    out = subprocess.run([exe, str(devices), str(log_n)], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "ALL OK" in out.stdout, out.stdout + out.stderr
