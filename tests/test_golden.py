"""Committed golden fixtures (tests/golden/vectors.json, made by tests/golden/make_golden.py): the oracle must reproduce
them on CPU, the CUDA path must reproduce them on the GPU."""
import json
import os

import numpy as np
import pytest

from oracle import pyref as P

G = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vectors.json")))


def hx(a):
    return [f"{int(v):016x}" for v in np.asarray(a, dtype=np.uint64).reshape(-1)]


def _msm_inputs(O, row):
    n = row["n"]
    sc = O.synth_scalars(row["scalar_seed"], n)
    bs = O.synth_bases(row["base_seed"], n, 2)
    if row["zeroed_scalar"] is not None:
        sc[row["zeroed_scalar"]] = 0
        bs[row["identity_base"]] = 0
    return sc, bs


def test_oracle_reproduces_golden(oracle):
    O = oracle
    for row in G["msm"]:
        sc, bs = _msm_inputs(O, row)
        _, aff = O.best_multiexp(sc, bs, 1)
        assert hx(aff) == row["affine"] and O.g1_to_bytes(aff).hex() == row["compressed"]
    for row in G["ntt"]:
        k = row["log_n"]
        a = O.synth_scalars(row["seed"], 1 << k)
        res = O.best_fft(a, P.int_to_limbs(P.to_mont(P.omega_for(k), P.R_MOD)), k, 1)
        assert hx(res[0]) == row["first"] and hx(res[-1]) == row["last"]
        assert hx(np.bitwise_xor.reduce(res, axis=0)) == row["xor_of_all_limbs"]
    s = O.synth_scalars(G["kzg"][0]["toxic_seed"], 1)[0]
    g, gl = O.params_setup(5, s)
    assert hx(g[-1]) == G["kzg"][0]["g_last"] and hx(gl[-1]) == G["kzg"][0]["g_lagrange_last"]


def _product_inputs(O, row):
    n = 1 << row["k"]
    beta, gamma, last_z = (O.synth_scalars(row["challenge_seed"] + j, 1)[0] for j in range(3))
    if row["kind"] == "permutation":
        cols = [O.synth_scalars(row["col_seed"] + j, n) for j in range(row["ncols"])]
        perms = [O.synth_scalars(row["perm_seed"] + j, n) for j in range(row["ncols"])]
        return cols, perms, beta, gamma, last_z
    return [O.synth_scalars(row["seed"] + j, n) for j in range(4)], None, beta, gamma, last_z


def test_oracle_reproduces_golden_products(oracle):
    O = oracle
    one = P.fr_array_from_ints([1])[0]
    for row in G["products"]:
        vecs, perms, beta, gamma, last_z = _product_inputs(O, row)
        if row["kind"] == "permutation":
            z, dw = O.permutation_product(vecs, perms, beta, gamma, P.int_to_limbs(P.to_mont(P.omega_for(row["k"]), P.R_MOD)), one, last_z)
            assert hx(dw) == row["deltaomega_out"]
        else:
            z = O.lookup_product(*vecs, beta, gamma)
        assert hx(z[-1]) == row["z_last"] and hx(np.bitwise_xor.reduce(z, axis=0)) == row["z_xor"]


@pytest.mark.gpu
def test_cuda_reproduces_golden_products(oracle):
    import ctypes

    import cqb200

    cqb200._lib.init(0)
    L, lib = cqb200._lib, cqb200._lib.lib()
    O = oracle

    def dev(arr):
        d = ctypes.c_void_p()
        L.check(lib.cqb_dev_alloc(max(arr.nbytes, 64), ctypes.byref(d)))
        L.check(lib.cqb_memcpy_h2d(d, np.ascontiguousarray(arr).ctypes.data_as(ctypes.c_void_p), arr.nbytes))
        return d

    for row in G["products"]:
        n = 1 << row["k"]
        vecs, perms, beta, gamma, last_z = _product_inputs(O, row)
        d_z = dev(np.zeros((n, 4), np.uint64))
        if row["kind"] == "permutation":
            dc, dp = [dev(c) for c in vecs], [dev(p_) for p_ in perms]
            dw = P.fr_array_from_ints([1])[0].copy()
            arr_c = (ctypes.c_void_p * len(dc))(*dc)
            arr_p = (ctypes.c_void_p * len(dp))(*dp)
            L.check(lib.cqb_permutation_product_dev(arr_c, arr_p, len(dc), row["k"], L.p64(beta), L.p64(gamma),
                                                    L.p64(P.int_to_limbs(P.to_mont(P.omega_for(row["k"]), P.R_MOD))),
                                                    L.p64(P.int_to_limbs(P.to_mont(cqb200.permutation.FR_DELTA, P.R_MOD))), L.p64(dw), L.p64(last_z), d_z))
            assert hx(dw) == row["deltaomega_out"]
            bufs = dc + dp
        else:
            bufs = [dev(v) for v in vecs]
            L.check(lib.cqb_lookup_product_dev(bufs[0], bufs[1], bufs[2], bufs[3], row["k"], L.p64(beta), L.p64(gamma), d_z))
        z = np.zeros((n, 4), np.uint64)
        L.check(lib.cqb_memcpy_d2h(z.ctypes.data_as(ctypes.c_void_p), d_z, n * 32))
        L.check(lib.cqb_sync())
        assert hx(z[-1]) == row["z_last"] and hx(np.bitwise_xor.reduce(z, axis=0)) == row["z_xor"]
        for b in bufs + [d_z]:
            L.check(lib.cqb_dev_free(b))


@pytest.mark.gpu
def test_cuda_reproduces_golden(oracle):
    import cqb200

    cqb200._lib.init(0)
    O = oracle
    for row in G["msm"]:
        sc, bs = _msm_inputs(O, row)
        assert hx(cqb200.best_multiexp(sc, bs).to_affine()) == row["affine"]
    for row in G["ntt"]:
        k = row["log_n"]
        a = O.synth_scalars(row["seed"], 1 << k)
        b = a.copy()
        cqb200.best_fft(b, P.int_to_limbs(P.to_mont(P.omega_for(k), P.R_MOD)), k)
        assert hx(b[0]) == row["first"] and hx(b[-1]) == row["last"] and hx(np.bitwise_xor.reduce(b, axis=0)) == row["xor_of_all_limbs"]
        if k >= 1:
            d = cqb200.EvaluationDomain(3, k)
            assert d.extended_k == row["coset_extended_k"]
            ext = d.coeff_to_extended(a)
            assert hx(np.bitwise_xor.reduce(ext.values, axis=0)) == row["coset_xor"]
            q = d.extended_to_coeff(d.divide_by_vanishing_poly(ext))
            assert hx(np.bitwise_xor.reduce(q, axis=0)) == row["quotient_xor"]
    kz = G["kzg"][0]
    s = O.synth_scalars(kz["toxic_seed"], 1)[0]
    params = cqb200.ParamsKZG.setup_from_toxic_waste(kz["k"], s, precompute=False)
    assert hx(params.g.to_host()[-1]) == kz["g_last"] and hx(params.g_lagrange.to_host()[-1]) == kz["g_lagrange_last"]
    assert hx(params.commit_lagrange(P.fr_array_from_ints(list(range(32)))).to_affine()) == kz["commit_lagrange_0_to_31"]
    params.free()
    cqr = G["cq"][0]
    t = cqb200.TableSRS.setup_from_toxic_waste(cqr["N"] - 1, s, precompute=False)
    assert hx(t.g_lagrange_opening_at_0.to_host()[-1]) == cqr["opening_at_0_last"]
    tv = cqb200.cq.StaticTableValues(O.synth_scalars(cqr["value_seed"], cqr["N"]), t.g1)
    qs = tv.qs.to_host()
    assert hx(qs[0]) == cqr["qs_first"] and hx(qs[-1]) == cqr["qs_last"]
    tv.free()
    t.free()
