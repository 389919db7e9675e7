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


def _build_prover(tmp):
    exe = os.path.join(tmp, "test_create_proof")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "cpp", "test_create_proof.cpp"),
                           "-L" + PKG, "-lcqb200", "-Wl,-rpath," + PKG])
    return exe


def test_cpp_create_proof_program_compiles_and_hashes(tmp_path):
    """CPU: the C++ create_proof mirror compiles, links, and its Blake2b (test infrastructure of the program) equals hashlib's"""
    import hashlib

    if not os.path.exists(os.path.join(PKG, "libcqb200.so")):
        pytest.skip("libcqb200.so not built")
    exe = _build_prover(str(tmp_path))
    for n in (0, 1, 127, 128, 129, 1000):
        msg = bytes((i * 7 + 3) & 255 for i in range(n))
        got = subprocess.run([exe, "--blake2b", str(n)], capture_output=True, text=True, check=True).stdout.strip()
        assert got == hashlib.blake2b(msg, digest_size=64, person=b"Halo2-Transcript").hexdigest()


@pytest.mark.gpu
@pytest.mark.parametrize("k,log_table,A", [(6, 5, 2), (8, 7, 5), (10, 12, 3)])
def test_cpp_create_proof_bytes_equal_python_mirror(tmp_path, k, log_table, A):
    """csrc/host/halo2_b200_prover.hpp vs sha2_on_cq_halo2_b200/prover.py on the same circuit, witness, rng and Blake2b transcript: identical
    proof bytes (the Python mirror's are tied to the oracle prover and verified in tests/test_gpu_full_proof.py)"""
    import sys

    import numpy as np

    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import cqb200
    import prove_real
    from sha2_on_cq_halo2_b200 import prover as PR
    from sha2_on_cq_halo2_b200.fields import fr_to_limbs

    c = prove_real.circuit_arrays(k, log_table, A, seed=7 + k)
    path = str(tmp_path / "circuit.bin")
    with open(path, "wb") as f:
        f.write(np.array([k, c["N"], A, c["bf"], c["cs_degree"], len(c["idx"]), 0, 0], np.uint64).tobytes())
        f.write(np.ascontiguousarray(c["s"], np.uint64).tobytes())
        f.write(fr_to_limbs(c["vk_repr"]).tobytes())
        for group in (c["tables"], c["advice"], c["sigma"], c["permutation_blinds"]):
            for arr in group:
                f.write(np.ascontiguousarray(arr, np.uint64).tobytes())
        f.write(np.ascontiguousarray(c["random_poly"], np.uint64).tobytes())
        idx = np.ascontiguousarray(c["idx"], np.uint32)
        f.write(idx.tobytes())
        if len(idx) & 1:
            f.write(b"\0\0\0\0")
        f.write(np.ascontiguousarray(c["mult"], np.uint64).tobytes())
    exe = _build_prover(str(tmp_path))
    out_path = str(tmp_path / "proof.out")
    res = subprocess.run([exe, path, out_path], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "ALL OK" in res.stdout, res.stdout + res.stderr
    cpp_proof = open(out_path, "rb").read()
    cqb200._lib.init(0)
    pk, witness, m_sparse, rnd, keep = prove_real.build_circuit(cqb200, k, log_table, A, seed=7 + k)
    t = prove_real.Blake2bTranscript()
    info = PR.create_proof(pk, witness, [m_sparse], rnd, t)
    assert PR.expected_h_eval(pk, info) == info["h_eval"]
    pk.free()
    params, tsrs, big, bound, tables = keep
    for tb in tables:
        tb.free()
    bound.free()
    if big is not tsrs:
        big.free()
    tsrs.free()
    params.free()
    assert len(cpp_proof) == len(t.proof) and cpp_proof == bytes(t.proof)
