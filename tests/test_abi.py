"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/cqb200.h declares,
and fails loudly (no CPU fallback) when no CUDA device is present. No compute calls here."""
import os
import re

import pytest


def _declared_symbols():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "include", "cqb200.h")).read()
    return sorted(set(re.findall(r"CQB_API[^;(]*?\b(cqb_\w+)\s*\(", text)))


def test_header_symbols_all_exported_and_bound():
    import cqb200

    lib = cqb200._lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 30
    bound = {name for name, _, _ in cqb200._lib.SYMBOLS}
    assert set(declared) == bound, (set(declared) ^ bound)
    for name in declared:
        assert getattr(lib, name) is not None


def test_no_cpu_fallback_without_device():
    import cqb200

    lib = cqb200._lib.load()
    if lib.cqb_device_count() > 0:
        pytest.skip("a CUDA device is visible; this checks the no-device failure mode")
    import ctypes
    import numpy as np

    assert lib.cqb_init(0) == cqb200._lib.CQB_E_NO_DEVICE
    out = np.zeros(8, np.uint64)
    inf = ctypes.c_int(0)
    a = np.zeros((4, 4), np.uint64)
    b = np.zeros((4, 8), np.uint64)
    rc = lib.cqb_msm_bn254_g1_host(cqb200._lib.p64(b), cqb200._lib.p64(a), 4, cqb200._lib.p64(out), ctypes.byref(inf))
    assert rc == cqb200._lib.CQB_E_NO_DEVICE
    assert b"no CPU fallback" in lib.cqb_last_error()
    rc = lib.cqb_ntt_bn254_fr(cqb200._lib.p64(a), cqb200._lib.p64(a[0]), 2)
    assert rc == cqb200._lib.CQB_E_NO_DEVICE
    with pytest.raises(cqb200._lib.CqbError):
        cqb200.best_multiexp(a, b)


def test_host_domain_constants_match_oracle(oracle):
    """EvaluationDomain::new is host logic (poly/domain.rs:39-142): the package's constants equal the oracle's"""
    import numpy as np

    import cqb200

    for j, k in ((1, 3), (3, 3), (4, 10), (5, 16), (3, 24)):
        d = cqb200.EvaluationDomain(j, k)
        od = oracle.domain_new(j, k)
        assert d.extended_k == od.extended_k
        for name in ("omega", "omega_inv", "extended_omega", "extended_omega_inv", "g_coset", "g_coset_inv",
                     "ifft_divisor", "extended_ifft_divisor"):
            assert np.array_equal(getattr(d, name), od.f(name)), (j, k, name)
        assert np.array_equal(d.t_evaluations, od.t_evals())
