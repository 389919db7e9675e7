"""SRS generation on the device and the element-wise helpers it uses, checked against the CPU
oracle's restatement of ParamsKZG::setup_from_toxic_waste (poly/kzg/commitment.rs:209-276)."""
import ctypes
import io

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import pyref as P  # noqa: E402


@pytest.fixture(scope="module")
def cq():
    import cqb200

    cqb200._lib.init(0)
    return cqb200


def _dev(cq, arr):
    lib = cq._lib.lib()
    d = ctypes.c_void_p()
    cq._lib.check(lib.cqb_dev_alloc(max(arr.nbytes, 64), ctypes.byref(d)))
    if arr.nbytes:
        cq._lib.check(lib.cqb_memcpy_h2d(d, arr.ctypes.data_as(ctypes.c_void_p), arr.nbytes))
    return d


def _host(cq, d, shape):
    out = np.zeros(shape, np.uint64)
    cq._lib.check(cq._lib.lib().cqb_memcpy_d2h(out.ctypes.data_as(ctypes.c_void_p), d, out.nbytes))
    return out


@pytest.mark.parametrize("k", [0, 1, 3, 6, 10])
def test_srs_setup_matches_reference_formulas(cq, oracle, k):
    s = oracle.synth_scalars(0xC0 + k, 1)[0]
    g_exp, gl_exp = oracle.params_setup(k, s)
    params = cq.ParamsKZG.setup_from_toxic_waste(k, s, precompute=False)
    assert np.array_equal(params.g.to_host(), g_exp)
    assert np.array_equal(params.g_lagrange.to_host(), gl_exp)
    # and it is usable as an SRS: kzg/commitment.rs:570-593 commit(ifft(a)) == commit_lagrange(a)
    if k >= 1:
        d = cq.EvaluationDomain(1, k)
        a = oracle.synth_scalars(5, 1 << k)
        assert params.commit(d.lagrange_to_coeff(a)) == params.commit_lagrange(a)
    params.free()


def test_srs_large_k_identities(cq, oracle):
    """k = 16: too slow for the CPU oracle's 2^17 scalar multiplications; checked through identities instead:
    sum_i g_lagrange[i] = G (the Lagrange basis sums to 1), g[0] = G, g[1] = [s]G, commit identity with the precomputed layout"""
    k = 16
    s = oracle.synth_scalars(0xD00D, 1)[0]
    params = cq.ParamsKZG.setup_from_toxic_waste(k, s)
    n = 1 << k
    ones = np.tile(P.int_to_limbs(P.MONT % P.R_MOD), (n, 1))
    assert np.array_equal(params.commit_lagrange(ones).to_affine(), oracle.g1_generator())
    e0 = np.zeros((n, 4), np.uint64)
    e0[0] = ones[0]
    assert np.array_equal(params.commit(e0).to_affine(), oracle.g1_generator())
    e1 = np.zeros((n, 4), np.uint64)
    e1[1] = ones[0]
    assert np.array_equal(params.commit(e1).to_affine(), oracle.g1_to_affine(oracle.g1_mul_a(oracle.g1_generator(), s)))
    d = cq.EvaluationDomain(1, k)
    a = oracle.synth_scalars(6, n)
    assert params.commit(d.lagrange_to_coeff(a)) == params.commit_lagrange(a)
    params.free()


def test_params_raw_format_roundtrip(cq, oracle):
    """commitment.rs:366-459 write_custom / read_custom (RawBytesUnchecked) and test_parameter_serialisation_roundtrip :595-621"""
    k = 4
    s = oracle.synth_scalars(0xC4, 1)[0]
    p0 = cq.ParamsKZG.setup_from_toxic_waste(k, s, precompute=False)
    g2, s_g2 = bytes(range(128)), bytes(reversed(range(128)))
    buf = io.BytesIO()
    p0.write(buf, g2=g2, s_g2=s_g2)
    raw = buf.getvalue()
    assert len(raw) == 4 + 2 * 16 * 64 + 256 and raw[:4] == (4).to_bytes(4, "little")
    g_exp, gl_exp = oracle.params_setup(k, s)
    assert raw[4:4 + 16 * 64] == g_exp.tobytes() and raw[4 + 16 * 64:4 + 32 * 64] == gl_exp.tobytes()
    p1 = cq.ParamsKZG.read(io.BytesIO(raw), precompute=False)
    assert p1.k == 4 and p1.g2 == g2 and p1.s_g2 == s_g2
    a = oracle.synth_scalars(8, 16)
    assert p0.commit(a) == p1.commit(a) and p0.commit_lagrange(a) == p1.commit_lagrange(a)
    p0.free()
    p1.free()


def test_generator_mul_batch(cq, oracle):
    lib = cq._lib.lib()
    n = 300
    sc = oracle.synth_scalars(0xF00, n)
    sc[0] = 0
    sc[1] = P.int_to_limbs(P.to_mont(P.R_MOD - 1, P.R_MOD))
    sc[2] = P.int_to_limbs(P.to_mont(1, P.R_MOD))
    sc[3] = P.int_to_limbs(P.to_mont(0x8000, P.R_MOD))
    sc[4] = P.int_to_limbs(P.to_mont(0xFFFF_FFFF, P.R_MOD))
    d_s = _dev(cq, sc)
    d_o = _dev(cq, np.zeros((n, 8), np.uint64))
    cq._lib.check(lib.cqb_g1_generator_mul_dev(d_s, n, d_o))
    got = _host(cq, d_o, (n, 8))
    g = oracle.g1_generator()
    for i in list(range(8)) + [57, 131, 299]:
        assert np.array_equal(got[i], oracle.g1_to_affine(oracle.g1_mul_a(g, sc[i]))), i
    cq._lib.check(lib.cqb_dev_free(d_s))
    cq._lib.check(lib.cqb_dev_free(d_o))


@pytest.mark.parametrize("n", [1, 31, 32, 33, 1000])
def test_batch_invert_and_powers(cq, oracle, n):
    lib = cq._lib.lib()
    a = oracle.synth_scalars(0xB1 + n, n)
    if n > 5:
        a[3] = 0  # zeros are skipped, as ff::BatchInvert does
    d = _dev(cq, a)
    cq._lib.check(lib.cqb_fr_batch_invert_dev(d, n))
    got = _host(cq, d, (n, 4))
    for i in range(n):
        assert np.array_equal(got[i], oracle.fr_op("invert", a[i])), i
    base = oracle.synth_scalars(0xB2, 1)[0]
    cq._lib.check(lib.cqb_fr_powers_dev(cq._lib.p64(base), n, d))
    pw = _host(cq, d, (n, 4))
    acc = oracle.fr_const("ONE")
    for i in range(n):
        assert np.array_equal(pw[i], acc), i
        acc = oracle.fr_op("mul", acc, base)
    cq._lib.check(lib.cqb_dev_free(d))


@pytest.mark.parametrize("k", [0, 1, 2, 5, 8])
def test_g_to_lagrange_parity(cq, oracle, k):
    """arithmetic.rs:277-301 (G1 EC-FFT) vs the oracle's restatement, on arbitrary curve points"""
    lib = cq._lib.lib()
    n = 1 << k
    g = oracle.synth_bases(0xEC + k, n, 2)
    if n >= 8:
        g[3] = 0                 # identity input
        g[5] = g[4]              # repeated point
    exp = oracle.g_to_lagrange(g, k)
    d_g = _dev(cq, g)
    d_o = _dev(cq, np.zeros((n, 8), np.uint64))
    cq._lib.check(lib.cqb_g_to_lagrange_dev(d_g, k, d_o))
    assert np.array_equal(_host(cq, d_o, (n, 8)), exp)
    cq._lib.check(lib.cqb_dev_free(d_g))
    cq._lib.check(lib.cqb_dev_free(d_o))


def test_downsize_matches_fresh_setup(cq, oracle):
    """poly/kzg/commitment.rs:482-490: downsize(k') of a k-params equals setup(k') with the same toxic waste"""
    s = oracle.synth_scalars(0xD5, 1)[0]
    big = cq.ParamsKZG.setup_from_toxic_waste(10, s, precompute=False)
    small = cq.ParamsKZG.setup_from_toxic_waste(7, s, precompute=False)
    big.downsize(7)
    assert big.k == 7 and big.n == 128
    assert np.array_equal(big.g.to_host(), small.g.to_host())
    assert np.array_equal(big.g_lagrange.to_host(), small.g_lagrange.to_host())
    big.free()
    small.free()


@pytest.mark.parametrize("N", [2, 16, 64])
def test_table_srs_setup_matches_reference(cq, oracle, N):
    """poly/kzg/commitment.rs:73-178 TableSRS::setup_from_toxic_waste, G1 parts incl. g_lagrange_opening_at_0"""
    s = oracle.synth_scalars(0xC7 + N, 1)[0]
    g1, gl, op0 = oracle.table_srs_setup(N, s)
    t = cq.TableSRS.setup_from_toxic_waste(N - 1, s, precompute=False)
    assert np.array_equal(t.g1.to_host(), g1)
    assert np.array_equal(t.g1_lagrange.to_host(), gl)
    assert np.array_equal(t.g_lagrange_opening_at_0.to_host(), op0)
    t.free()


@pytest.mark.parametrize("N", [2, 4, 16, 64])
def test_cq_table_qs_fk_matches_reference(cq, oracle, N):
    """StaticTableValues::new's qs (plonk/static_lookup.rs:77-126) — FK on the device vs the oracle's
    restatement of the reference's O(N^2) loop (N kate_divisions + N MSMs)"""
    s = oracle.synth_scalars(0xF4 + N, 1)[0]
    srs = cq.TableSRS.setup_from_toxic_waste(N - 1, s, precompute=False)
    rng = np.random.default_rng(N)
    vals = P.fr_array_from_ints([int(v) for v in rng.choice(1 << 30, N, replace=False)])
    exp = oracle.cq_table_qs(vals, srs.g1.to_host(), 4)
    tv = cq.cq.StaticTableValues(vals, srs.g1)
    assert np.array_equal(tv.qs.to_host(), exp)
    tv.free()
    srs.free()


def test_cq_table_qs_large_n_spot_check(cq, oracle):
    """N = 2^12: the oracle's O(N^2) loop would take minutes; spot-check a few rows against one kate_division + MSM each"""
    N, k = 1 << 12, 12
    s = oracle.synth_scalars(0xF5, 1)[0]
    srs = cq.TableSRS.setup_from_toxic_waste(N - 1, s, precompute=False)
    vals = oracle.synth_scalars(0xF6, N)
    tv = cq.cq.StaticTableValues(vals, srs.g1)
    qs = tv.qs.to_host()
    g1 = srs.g1.to_host()
    od = oracle.domain_new(2, k)
    coeffs = oracle.lagrange_to_coeff(od, vals)
    w = P.omega_for(k)
    n_inv = pow(N, -1, P.R_MOD)
    for i in (0, 1, 77, N - 1):
        gi = pow(w, i, P.R_MOD)
        q = oracle.kate_division(coeffs, P.int_to_limbs(P.to_mont(gi, P.R_MOD)))
        sc = P.int_to_limbs(P.to_mont(gi * n_inv % P.R_MOD, P.R_MOD))
        q = np.stack([oracle.fr_op("mul", row, sc) for row in q])
        _, exp = oracle.best_multiexp(q, g1[: N - 1], 8)
        assert np.array_equal(qs[i], exp), i
    tv.free()
    srs.free()


@pytest.mark.parametrize("n", [1, 2, 63, 64, 65, 4096, 4097, 300000])
def test_eval_polynomial_and_kate_division(cq, oracle, n):
    """arithmetic.rs:304-329 and :351-387 on the device vs the oracle's serial restatements"""
    a = oracle.synth_scalars(0xE0 + n, n)
    x = oracle.synth_scalars(0xE1, 1)[0]
    assert np.array_equal(cq.eval_polynomial(a, x), oracle.eval_polynomial(a, x))
    zero = np.zeros(4, np.uint64)
    assert np.array_equal(cq.eval_polynomial(a, zero), a[0])
    q = cq.kate_division(a, x)
    assert q.shape == (n - 1, 4)
    if n > 1:
        assert np.array_equal(q, oracle.kate_division(a, x))


def test_g2_powers_and_table_commit_parity(cq, oracle):
    """The G2 half of the SRS and the CQ table commitment: [s^i]G2 (poly/kzg/commitment.rs:94-104, 265-266) and StaticTableValues::commit
    (plonk/static_lookup.rs:127-160: zv, t = G2 multiexp over the SORTED table values' coefficients, x_b0_bound) vs the oracle"""
    N, k_circ = 64, 5
    s = oracle.synth_scalars(0x62, 1)[0]
    srs = cq.TableSRS.setup_from_toxic_waste(N - 1, s, precompute=False, max_g2_power=N)
    exp_g2 = oracle.g2_powers(s, N + 1)
    assert np.array_equal(srs.g2, exp_g2)
    assert np.array_equal(srs.g2[0], oracle.g2_generator())
    # G2 multiexp with edge scalars / an identity base
    sc = oracle.synth_scalars(0x63, N)
    sc[0] = 0
    sc[1] = P.int_to_limbs(P.to_mont(P.R_MOD - 1, P.R_MOD))
    bases = exp_g2[:N].copy()
    bases[2] = 0
    out = np.zeros(16, np.uint64)
    inf = ctypes.c_int(0)
    L = cq._lib
    L.check(L.lib().cqb_msm_bn254_g2(L.p64(bases), L.p64(sc), N, L.p64(out), ctypes.byref(inf)))
    assert np.array_equal(out, oracle.g2_msm(bases, sc)) and inf.value == 0
    L.check(L.lib().cqb_msm_bn254_g2(L.p64(bases), L.p64(np.zeros((N, 4), np.uint64)), N, L.p64(out), ctypes.byref(inf)))
    assert inf.value == 1 and not out.any()
    # table commitment
    rng = np.random.default_rng(4)
    vals = P.fr_array_from_ints([int(v) for v in rng.choice(1 << 40, N, replace=False)])
    table = cq.cq.StaticTableValues(vals, srs.g1)
    ct = table.commit(N, srs.g2, 1 << k_circ)
    assert np.array_equal(ct["zv"], oracle.g2_add_aa(exp_g2[N], oracle.g2_neg_a(exp_g2[0])))
    order = sorted(range(N), key=lambda i: P.fr_array_to_ints(vals[i:i + 1])[0])
    od = oracle.domain_new(2, 6)
    coeffs = oracle.lagrange_to_coeff(od, np.ascontiguousarray(vals[order]))
    assert np.array_equal(ct["t"], oracle.g2_msm(exp_g2[:N], coeffs))
    assert np.array_equal(ct["x_b0_bound"], exp_g2[N - 1 - ((1 << k_circ) - 2)]) and ct["size"] == N
    table.free()
    srs.free()
    # ParamsKZG: g2 / s_g2 come with the setup now, so the params can be written and read back without external blobs
    import io

    params = cq.ParamsKZG.setup_from_toxic_waste(5, s, precompute=False)
    assert params.g2 == exp_g2[0].tobytes() and params.s_g2 == exp_g2[1].tobytes()
    buf = io.BytesIO()
    params.write(buf)
    back = cq.ParamsKZG.read(io.BytesIO(buf.getvalue()), precompute=False)
    assert back.k == 5 and back.s_g2 == params.s_g2
    buf2 = io.BytesIO()
    back.write(buf2)                     # read -> write round trip of host-built params
    assert buf2.getvalue() == buf.getvalue()
    back.downsize(4)                     # and downsize on params that came from a file
    fresh = cq.ParamsKZG.setup_from_toxic_waste(4, s, precompute=False)
    assert np.array_equal(back.g_lagrange.to_host(), fresh.g_lagrange.to_host())
    for p_ in (params, back, fresh):
        p_.free()
