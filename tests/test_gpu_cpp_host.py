"""Builds and runs the C++ host-API tests (tests/cpp/test_host_api.cpp over csrc/host/halo2_b200.hpp): the C++ mirror of
best_multiexp / best_fft / EvaluationDomain / ParamsKZG, checked against the oracle, on a GPU box."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "sha2-on-cq-halo2_b200")


def _build(tmp):
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s", "all"])
    exe = os.path.join(tmp, "test_host_api")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "cpp", "test_host_api.cpp"),
                           "-L" + PKG, "-lcqb200", "-L" + os.path.join(ROOT, "oracle", "_build"), "-loracle",
                           "-Wl,-rpath," + PKG, "-Wl,-rpath," + os.path.join(ROOT, "oracle", "_build")])
    return exe


def test_cpp_host_header_compiles(tmp_path):
    """CPU: the header + test program compile and link against the built libraries (no GPU needed to link)"""
    if not os.path.exists(os.path.join(PKG, "libcqb200.so")):
        pytest.skip("libcqb200.so not built")
    assert os.path.exists(_build(str(tmp_path)))


@pytest.mark.gpu
def test_cpp_host_api(tmp_path):
    exe = _build(str(tmp_path))
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "ALL OK" in out.stdout, out.stdout + out.stderr
