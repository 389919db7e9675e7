"""oracle_lib.py — TEST INFRASTRUCTURE. ctypes binding of oracle/_build/liboracle.so (the C restatement of the reference).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
Arrays are numpy uint64: Fr/Fq = (...,4), G1 affine = (...,8), G1 Jacobian = (...,12); Montgomery form throughout.
"""
import ctypes
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None

u64p = ctypes.POINTER(ctypes.c_uint64)
u32p = ctypes.POINTER(ctypes.c_uint32)
u8p = ctypes.POINTER(ctypes.c_uint8)


def build(force=False):
    if force or not os.path.exists(_SO) or any(
        os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_SO) for f in ("bn254_oracle.c", "field_impl.inc", "Makefile")
    ):
        subprocess.check_call(["make", "-C", _HERE, "-s", "all"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
        _lib.oracle_g1_is_on_curve.restype = ctypes.c_int
        _lib.oracle_extended_to_coeff.restype = ctypes.c_size_t
        _lib.oracle_domain_sizeof.restype = ctypes.c_size_t
        _lib.oracle_hw_threads.restype = ctypes.c_int
        _lib.oracle_domain_new.restype = ctypes.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(u64p)


def _c(a, shape_last):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    assert a.shape[-1] == shape_last, a.shape
    return a


FR_OPS = {"add": 0, "sub": 1, "mul": 2, "square": 3, "neg": 4, "invert": 5, "double": 6, "from_mont": 7, "from_raw": 8}


def _field_op(fn, op, a, b=None):
    a = _c(a, 4)
    b = _c(b, 4) if b is not None else a
    out = np.zeros(4, np.uint64)
    fn(FR_OPS[op], _p(a), _p(b), _p(out))
    return out


def fr_op(op, a, b=None):
    return _field_op(lib().oracle_fr_op, op, a, b)


def fq_op(op, a, b=None):
    return _field_op(lib().oracle_fq_op, op, a, b)


def fr_from_u512(limbs8):
    a = np.ascontiguousarray(limbs8, dtype=np.uint64)
    out = np.zeros(4, np.uint64)
    lib().oracle_fr_from_u512(_p(a), _p(out))
    return out


def fq_from_u512(limbs8):
    a = np.ascontiguousarray(limbs8, dtype=np.uint64)
    out = np.zeros(4, np.uint64)
    lib().oracle_fq_from_u512(_p(a), _p(out))
    return out


def fr_const(name):
    which = {"ROOT_OF_UNITY": 0, "ROOT_OF_UNITY_INV": 1, "TWO_INV": 2, "DELTA": 3, "ZETA": 4, "ONE": 5, "GENERATOR": 6}[name]
    out = np.zeros(4, np.uint64)
    lib().oracle_fr_const(which, _p(out))
    return out


def fr_pow(a, e):
    a = _c(a, 4)
    ee = np.array([(e >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)
    out = np.zeros(4, np.uint64)
    lib().oracle_fr_pow(_p(a), _p(ee), _p(out))
    return out


def fr_to_repr(a):
    a = _c(a, 4)
    out = np.zeros(32, np.uint8)
    lib().oracle_fr_to_repr(_p(a), out.ctypes.data_as(u8p))
    return bytes(out)


def g1_generator():
    out = np.zeros(8, np.uint64)
    lib().oracle_g1_generator(_p(out))
    return out


def g1_is_on_curve(a):
    return bool(lib().oracle_g1_is_on_curve(_p(_c(a, 8))))


def _g1(fn, a, la, b, lb, lo):
    a = _c(a, la)
    out = np.zeros(lo, np.uint64)
    if b is None:
        fn(_p(a), _p(out))
    else:
        fn(_p(a), _p(_c(b, lb)), _p(out))
    return out


def g1_add_jj(a, b):
    return _g1(lib().oracle_g1_add_jj, a, 12, b, 12, 12)


def g1_add_ja(a, b):
    return _g1(lib().oracle_g1_add_ja, a, 12, b, 8, 12)


def g1_add_aa(a, b):
    return _g1(lib().oracle_g1_add_aa, a, 8, b, 8, 12)


def g1_double(a):
    return _g1(lib().oracle_g1_double, a, 12, None, 0, 12)


def g1_neg_a(a):
    return _g1(lib().oracle_g1_neg_a, a, 8, None, 0, 8)


def g1_to_affine(a):
    return _g1(lib().oracle_g1_to_affine, a, 12, None, 0, 8)


def g1_to_curve(a):
    return _g1(lib().oracle_g1_to_curve, a, 8, None, 0, 12)


def g1_mul_a(a, s):
    return _g1(lib().oracle_g1_mul_a, a, 8, s, 4, 12)


def g1_mul_j(a, s):
    return _g1(lib().oracle_g1_mul_j, a, 12, s, 4, 12)


def g1_batch_normalize(p):
    p = _c(p, 12)
    out = np.zeros((p.shape[0], 8), np.uint64)
    lib().oracle_g1_batch_normalize(_p(p), _p(out), ctypes.c_size_t(p.shape[0]))
    return out


def g1_to_bytes(a):
    out = np.zeros(32, np.uint8)
    lib().oracle_g1_to_bytes(_p(_c(a, 8)), out.ctypes.data_as(u8p))
    return bytes(out)


def hw_threads():
    return int(lib().oracle_hw_threads())


def best_multiexp(coeffs, bases, num_threads=1):
    """reference arithmetic.rs:132 — returns (jacobian(12,), affine(8,))"""
    coeffs = _c(coeffs, 4)
    bases = _c(bases, 8)
    assert coeffs.shape[0] == bases.shape[0], "assert_eq!(coeffs.len(), bases.len())"  # arithmetic.rs:133
    jac = np.zeros(12, np.uint64)
    aff = np.zeros(8, np.uint64)
    lib().oracle_best_multiexp(_p(coeffs), _p(bases), ctypes.c_size_t(coeffs.shape[0]), ctypes.c_size_t(num_threads), _p(jac), _p(aff))
    return jac, aff


def best_fft(a, omega, log_n, threads=1):
    """reference arithmetic.rs:171 — returns a new array"""
    a = np.array(_c(a, 4), copy=True)
    assert a.shape[0] == 1 << log_n, "assert_eq!(n, 1 << log_n)"  # arithmetic.rs:184
    lib().oracle_best_fft(_p(a), _p(_c(omega, 4)), ctypes.c_uint32(log_n), ctypes.c_size_t(threads))
    return a


def ifft(a, omega_inv, log_n, divisor, threads=1):
    a = np.array(_c(a, 4), copy=True)
    lib().oracle_ifft(_p(a), _p(_c(omega_inv, 4)), ctypes.c_uint32(log_n), _p(_c(divisor, 4)), ctypes.c_size_t(threads))
    return a


class Domain(ctypes.Structure):
    _fields_ = [
        ("n", ctypes.c_uint64), ("k", ctypes.c_uint32), ("extended_k", ctypes.c_uint32),
        ("quotient_poly_degree", ctypes.c_uint64), ("t_len", ctypes.c_uint32), ("_pad", ctypes.c_uint32),
        ("omega", ctypes.c_uint64 * 4), ("omega_inv", ctypes.c_uint64 * 4), ("extended_omega", ctypes.c_uint64 * 4),
        ("extended_omega_inv", ctypes.c_uint64 * 4), ("g_coset", ctypes.c_uint64 * 4), ("g_coset_inv", ctypes.c_uint64 * 4),
        ("ifft_divisor", ctypes.c_uint64 * 4), ("extended_ifft_divisor", ctypes.c_uint64 * 4),
        ("barycentric_weight", ctypes.c_uint64 * 4), ("t_evaluations", (ctypes.c_uint64 * 4) * 64),
    ]

    def f(self, name):
        return np.array(list(getattr(self, name)), dtype=np.uint64)

    def t_evals(self):
        return np.array([list(self.t_evaluations[i]) for i in range(self.t_len)], dtype=np.uint64)


def domain_new(j, k):
    """reference poly/domain.rs:39 EvaluationDomain::new(j, k)"""
    d = Domain()
    assert ctypes.sizeof(Domain) == lib().oracle_domain_sizeof(), (ctypes.sizeof(Domain), lib().oracle_domain_sizeof())
    rc = lib().oracle_domain_new(ctypes.c_uint32(j), ctypes.c_uint32(k), ctypes.byref(d))
    assert rc == 0, rc
    return d


def lagrange_to_coeff(d, a, threads=1):
    a = np.array(_c(a, 4), copy=True)
    assert a.shape[0] == 1 << d.k
    lib().oracle_lagrange_to_coeff(ctypes.byref(d), _p(a), ctypes.c_size_t(threads))
    return a


def coeff_to_extended(d, a, threads=1):
    a = _c(a, 4)
    assert a.shape[0] == 1 << d.k
    out = np.zeros((1 << d.extended_k, 4), np.uint64)
    lib().oracle_coeff_to_extended(ctypes.byref(d), _p(a), _p(out), ctypes.c_size_t(threads))
    return out


def divide_by_vanishing_poly(d, a):
    a = np.array(_c(a, 4), copy=True)
    assert a.shape[0] == 1 << d.extended_k
    lib().oracle_divide_by_vanishing_poly(ctypes.byref(d), _p(a))
    return a


def extended_to_coeff(d, a, threads=1):
    a = np.array(_c(a, 4), copy=True)
    assert a.shape[0] == 1 << d.extended_k
    keep = lib().oracle_extended_to_coeff(ctypes.byref(d), _p(a), ctypes.c_size_t(threads))
    return a[:keep]


def params_setup(k, s):
    """reference kzg/commitment.rs:209 setup_from_toxic_waste -> (g, g_lagrange)"""
    n = 1 << k
    g = np.zeros((n, 8), np.uint64)
    gl = np.zeros((n, 8), np.uint64)
    lib().oracle_params_setup(ctypes.c_uint32(k), _p(_c(s, 4)), _p(g), _p(gl))
    return g, gl


def table_srs_setup(g1_len, s):
    """reference kzg/commitment.rs:73 TableSRS::setup_from_toxic_waste (G1 parts)"""
    g1 = np.zeros((g1_len, 8), np.uint64)
    gl = np.zeros((g1_len, 8), np.uint64)
    op0 = np.zeros((g1_len, 8), np.uint64)
    lib().oracle_table_srs_setup(ctypes.c_size_t(g1_len), _p(_c(s, 4)), _p(g1), _p(gl), _p(op0))
    return g1, gl, op0


def sparse_commit(bases, idx, scalars):
    """reference static_lookup/prover.rs:167-170 / :245-257 serial scalar-mul loop -> affine"""
    bases = _c(bases, 8)
    idx = np.ascontiguousarray(idx, dtype=np.uint32)
    scalars = _c(scalars, 4)
    out = np.zeros(8, np.uint64)
    lib().oracle_sparse_commit(_p(bases), idx.ctypes.data_as(u32p), _p(scalars), ctypes.c_size_t(idx.shape[0]), _p(out))
    return out


def synth_scalars(seed, n, start=0):
    out = np.zeros((n, 4), np.uint64)
    lib().oracle_synth_scalars(ctypes.c_uint64(seed), ctypes.c_size_t(start), ctypes.c_size_t(n), _p(out))
    return out


def synth_bases(seed, n, threads=None):
    out = np.zeros((n, 8), np.uint64)
    lib().oracle_synth_bases(ctypes.c_uint64(seed), ctypes.c_size_t(n), ctypes.c_size_t(threads or hw_threads()), _p(out))
    return out


def g_to_lagrange(g_affine, k):
    """reference arithmetic.rs:277-301 (G1 EC-FFT), as used by ParamsKZG::downsize (commitment.rs:482-490)"""
    g_affine = _c(g_affine, 8)
    assert g_affine.shape[0] == 1 << k
    out = np.zeros((1 << k, 8), np.uint64)
    lib().oracle_g_to_lagrange(_p(g_affine), ctypes.c_uint32(k), _p(out))
    return out


def eval_polynomial(poly, point):
    """reference arithmetic.rs:304-329"""
    poly = _c(poly, 4)
    out = np.zeros(4, np.uint64)
    lib().oracle_eval_polynomial(_p(poly), ctypes.c_size_t(poly.shape[0]), _p(_c(point, 4)), _p(out))
    return out


def kate_division(a, b):
    """reference arithmetic.rs:351-387: (a(X) - a(b)) / (X - b), n-1 coefficients"""
    a = _c(a, 4)
    q = np.zeros((a.shape[0] - 1, 4), np.uint64)
    lib().oracle_kate_division(_p(a), ctypes.c_size_t(a.shape[0]), _p(_c(b, 4)), _p(q))
    return q


def cq_table_qs(values, srs_g1, threads=None):
    """reference plonk/static_lookup.rs:77-126 StaticTableValues::new: the N cached quotient commitments (affine)"""
    values = _c(values, 4)
    srs_g1 = _c(srs_g1, 8)
    n = values.shape[0]
    assert n & (n - 1) == 0 and srs_g1.shape[0] >= max(n - 1, 1)
    out = np.zeros((n, 8), np.uint64)
    lib().oracle_cq_table_qs(_p(values), ctypes.c_size_t(n), _p(srs_g1), ctypes.c_size_t(threads or hw_threads()), _p(out))
    return out


def _ptr_array(arrs):
    keep = [np.ascontiguousarray(a, dtype=np.uint64) for a in arrs]
    arr = (ctypes.c_void_p * max(len(keep), 1))(*[a.ctypes.data_as(ctypes.c_void_p) for a in keep])
    return arr, keep


def graph_evaluate(constants, rotations, code, n_calc, num_intermediates, fixed, advice, instance, challenges, beta, gamma, theta, y,
                   values, rot_scale):
    """reference plonk/evaluation.rs:718-775 GraphEvaluator::evaluate over all rows, on the serialised graph"""
    values = np.array(_c(values, 4), copy=True)
    constants = _c(constants, 4)
    rotations = np.ascontiguousarray(rotations, dtype=np.int32)
    code = np.ascontiguousarray(code, dtype=np.uint32)
    fa, k1 = _ptr_array(fixed)
    aa, k2 = _ptr_array(advice)
    ia, k3 = _ptr_array(instance)
    ch = _c(challenges, 4) if len(challenges) else np.zeros((1, 4), np.uint64)
    bgty = np.concatenate([_c(beta, 4), _c(gamma, 4), _c(theta, 4), _c(y, 4)])
    lib().oracle_graph_evaluate(_p(constants), rotations.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), ctypes.c_uint32(rotations.shape[0]),
                                code.ctypes.data_as(u32p), ctypes.c_uint32(n_calc), ctypes.c_uint32(num_intermediates), fa, aa, ia, _p(ch),
                                _p(bgty), _p(values), ctypes.c_size_t(values.shape[0]), ctypes.c_int32(rot_scale))
    return values


def cq_lookup_h(values, b_coset, f_coset, l_active_row, beta, y):
    """reference plonk/evaluation.rs:533-548"""
    values = np.array(_c(values, 4), copy=True)
    lib().oracle_cq_lookup_h(_p(values), _p(_c(b_coset, 4)), _p(_c(f_coset, 4)), _p(_c(l_active_row, 4)), _p(_c(beta, 4)), _p(_c(y, 4)),
                             ctypes.c_size_t(values.shape[0]))
    return values


def permutation_h(values, rot_scale, last_rotation, chunk_len, sets, columns, perm_cosets, l0, l_last, l_active, beta, gamma, y, extended_omega):
    """reference plonk/evaluation.rs:376-452"""
    values = np.array(_c(values, 4), copy=True)
    sa, k1 = _ptr_array(sets)
    ca, k2 = _ptr_array(columns)
    pa, k3 = _ptr_array(perm_cosets)
    lib().oracle_permutation_h(_p(values), ctypes.c_size_t(values.shape[0]), ctypes.c_int32(rot_scale), ctypes.c_int32(last_rotation),
                               ctypes.c_uint32(chunk_len), sa, ctypes.c_uint32(len(sets)), ca, pa, ctypes.c_uint32(len(columns)),
                               _p(_c(l0, 4)), _p(_c(l_last, 4)), _p(_c(l_active, 4)), _p(_c(beta, 4)), _p(_c(gamma, 4)), _p(_c(y, 4)),
                               _p(_c(extended_omega, 4)))
    return values


def permutation_product(columns, perms, beta, gamma, omega, deltaomega, last_z):
    """one column set of permutation::Argument::commit (permutation/prover.rs:82-166): returns (z, next deltaomega);
    z is the Lagrange grand-product vector before the blinding rows are overwritten"""
    cols = [_c(a, 4) for a in columns]
    prm = [_c(a, 4) for a in perms]
    n = cols[0].shape[0]
    z = np.zeros((n, 4), np.uint64)
    dw = np.array(deltaomega, dtype=np.uint64).copy()
    pc, _k1 = _ptr_array(cols)
    pp, _k2 = _ptr_array(prm)
    lib().oracle_permutation_product(pc, pp, len(cols), ctypes.c_size_t(n), _p(_c(beta, 4)), _p(_c(gamma, 4)), _p(_c(omega, 4)), _p(dw),
                                     _p(_c(last_z, 4)), _p(z))
    return z, dw


def lookup_product(compressed_input, compressed_table, permuted_input, permuted_table, beta, gamma):
    """lookup::prover::Permuted::commit_product (plonk/lookup/prover.rs:173-262): z before the blinding rows"""
    a, s, ap, sp = (_c(x, 4) for x in (compressed_input, compressed_table, permuted_input, permuted_table))
    n = a.shape[0]
    z = np.zeros((n, 4), np.uint64)
    lib().oracle_lookup_product(_p(a), _p(s), _p(ap), _p(sp), ctypes.c_size_t(n), _p(_c(beta, 4)), _p(_c(gamma, 4)), _p(z))
    return z


def lookup_h(values, rot_scale, table_value, product, permuted_input, permuted_table, l0, l_last, l_active, beta, gamma, y):
    """evaluate_h, plookup constraints of one lookup (plonk/evaluation.rs:458-531); returns the updated values"""
    v = _c(values, 4).copy()
    arrs = [_c(x, 4) for x in (table_value, product, permuted_input, permuted_table, l0, l_last, l_active)]
    lib().oracle_lookup_h(_p(v), ctypes.c_size_t(v.shape[0]), ctypes.c_int32(rot_scale), *[_p(x) for x in arrs], _p(_c(beta, 4)), _p(_c(gamma, 4)),
                          _p(_c(y, 4)))
    return v


# ---- G2 (affine (16,) uint64: x.c0 x.c1 y.c0 y.c1, Montgomery limbs; identity = zeros) ----------------------------------------
def g2_generator():
    out = np.zeros(16, np.uint64)
    lib().oracle_g2_generator(_p(out))
    return out


def g2_is_on_curve(a):
    lib().oracle_g2_is_on_curve.restype = ctypes.c_int
    return bool(lib().oracle_g2_is_on_curve(_p(_c(a, 16))))


def g2_mul_a(a, s):
    out = np.zeros(16, np.uint64)
    lib().oracle_g2_mul_a(_p(_c(a, 16)), _p(_c(s, 4)), _p(out))
    return out


def g2_add_aa(a, b):
    out = np.zeros(16, np.uint64)
    lib().oracle_g2_add_aa(_p(_c(a, 16)), _p(_c(b, 16)), _p(out))
    return out


def g2_neg_a(a):
    out = np.zeros(16, np.uint64)
    lib().oracle_g2_neg_a(_p(_c(a, 16)), _p(out))
    return out


def g2_powers(s, count):
    out = np.zeros((count, 16), np.uint64)
    lib().oracle_g2_powers(_p(_c(s, 4)), ctypes.c_size_t(count), _p(out))
    return out


def g2_msm(bases, scalars):
    bases, scalars = _c(bases, 16), _c(scalars, 4)
    assert bases.shape[0] == scalars.shape[0]
    out = np.zeros(16, np.uint64)
    lib().oracle_g2_msm(_p(bases), _p(scalars), ctypes.c_size_t(bases.shape[0]), _p(out))
    return out
