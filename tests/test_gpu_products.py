"""grand products on the device — the exclusive running product behind the z vectors and the
permutation argument's product sets (reference plonk/permutation/prover.rs:82-166) — against the oracle's restatement of
the reference's serial loops."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import pyref as P  # noqa: E402


@pytest.fixture(scope="module")
def cq():
    import cqb200

    cqb200._lib.init(0)
    return cqb200


class Dev:
    """device copy of a (n, 4) uint64 array"""

    def __init__(self, cq, arr=None, n=None):
        self.L = cq._lib
        self.lib = cq._lib.lib()
        nbytes = arr.nbytes if arr is not None else n * 32
        self.nbytes = nbytes
        d = ctypes.c_void_p()
        self.L.check(self.lib.cqb_dev_alloc(max(nbytes, 64), ctypes.byref(d)))
        self.ptr = d.value
        if arr is not None:
            arr = np.ascontiguousarray(arr, dtype=np.uint64)
            self.L.check(self.lib.cqb_memcpy_h2d(d, arr.ctypes.data_as(ctypes.c_void_p), arr.nbytes))

    def get(self, n):
        out = np.zeros((n, 4), np.uint64)
        self.L.check(self.lib.cqb_memcpy_d2h(out.ctypes.data_as(ctypes.c_void_p), ctypes.c_void_p(self.ptr), n * 32))
        self.L.check(self.lib.cqb_sync())
        return out

    def free(self):
        self.L.check(self.lib.cqb_dev_free(ctypes.c_void_p(self.ptr)))


@pytest.mark.parametrize("n", [1, 2, 31, 32, 33, 1024, 1025, 40000, (1 << 17) + 5])
def test_prefix_product_matches_serial_loop(cq, oracle, n):
    v = oracle.synth_scalars(0xA11CE + n, n)
    if n > 40:
        v[7] = P.fr_array_from_ints([1])[0]
        v[n // 2] = P.fr_array_from_ints([P.R_MOD - 1])[0]
    init = oracle.synth_scalars(0xB0B, 1)[0]
    exp = np.zeros((n, 4), np.uint64)
    acc = init
    for i in range(n):  # permutation/prover.rs:157-163
        exp[i] = acc
        acc = oracle.fr_op("mul", acc, v[i])
    d_in, d_out = Dev(cq, v), Dev(cq, n=n)
    cq._lib.check(cq._lib.lib().cqb_fr_prefix_product_dev(ctypes.c_void_p(d_in.ptr), n, cq._lib.p64(init), ctypes.c_void_p(d_out.ptr)))
    assert np.array_equal(d_out.get(n), exp)
    # in place
    cq._lib.check(cq._lib.lib().cqb_fr_prefix_product_dev(ctypes.c_void_p(d_in.ptr), n, cq._lib.p64(init), ctypes.c_void_p(d_in.ptr)))
    assert np.array_equal(d_in.get(n), exp)
    # a zero factor zeroes everything after it (the reference's loop does the same)
    if n >= 33:
        v2 = v.copy()
        v2[20] = 0
        d2 = Dev(cq, v2)
        cq._lib.check(cq._lib.lib().cqb_fr_prefix_product_dev(ctypes.c_void_p(d2.ptr), n, cq._lib.p64(init), ctypes.c_void_p(d2.ptr)))
        got = d2.get(n)
        assert np.array_equal(got[:21], exp[:21]) and not got[21:].any()
        d2.free()
    d_in.free()
    d_out.free()


@pytest.mark.parametrize("k,ncols,cs_degree,blinding_factors", [(0, 1, 3, 0), (1, 2, 4, 0), (3, 2, 3, 2), (6, 5, 4, 5), (10, 7, 5, 5), (13, 3, 9, 6)])
def test_permutation_commit_matches_reference_loop(cq, oracle, k, ncols, cs_degree, blinding_factors):
    """permutation::Argument::commit over all column sets: z vectors (with the caller's blinding rows), last_z / deltaomega
    threading between sets, and the commitment of every z (params.commit_lagrange, :166)"""
    n = 1 << k
    rng = np.random.default_rng(100 + k)
    omega = P.omega_for(k)
    beta, gamma = 0x1234567890ABCDEF1234567, 0xFEDCBA987654321
    # columns with small witness-like values; permutation polynomials: sigma_j[i] = delta^j' * omega^i' for a random permutation
    # of the (column, row) cells (keygen's build_pk, permutation/keygen.rs) — any values exercise the same arithmetic
    cols = [P.fr_array_from_ints([int(x) for x in rng.integers(0, 1 << 40, n)]) for _ in range(ncols)]
    cells = [(j, i) for j in range(ncols) for i in range(n)]
    perm = rng.permutation(len(cells))
    delta_pows = [pow(cq.permutation.FR_DELTA, j, P.R_MOD) for j in range(ncols)]
    omega_pows = [1] * n
    for i in range(1, n):
        omega_pows[i] = omega_pows[i - 1] * omega % P.R_MOD
    sig = [[0] * n for _ in range(ncols)]
    for src, dst in zip(range(len(cells)), perm):
        j, i = cells[src]
        jj, ii = cells[int(dst)]
        sig[j][i] = delta_pows[jj] * omega_pows[ii] % P.R_MOD
    perms = [P.fr_array_from_ints(s) for s in sig]
    chunk_len = cs_degree - 2
    nsets = (ncols + chunk_len - 1) // chunk_len
    blind_rows = [oracle.synth_scalars(0xB11D + s, blinding_factors) for s in range(nsets)]
    # oracle: the reference loop, set by set
    L = lambda x: P.fr_array_from_ints([x])[0]  # noqa: E731
    dw, last_z = L(1), L(1)
    exp_z = []
    for s in range(nsets):
        z, dw = oracle.permutation_product(cols[s * chunk_len:(s + 1) * chunk_len], perms[s * chunk_len:(s + 1) * chunk_len], L(beta), L(gamma),
                                           L(omega), dw, last_z)
        z[n - blinding_factors:] = blind_rows[s]
        last_z = z[n - (blinding_factors + 1)].copy()
        exp_z.append(z)
    # device
    d_cols = [Dev(cq, c) for c in cols]
    d_perms = [Dev(cq, p) for p in perms]
    d_z = [Dev(cq, n=n) for _ in range(nsets)]
    got_last = cq.permutation.commit_dev([d.ptr for d in d_cols], [d.ptr for d in d_perms], k, cs_degree, blinding_factors, beta, gamma, omega,
                                         blind_rows, [d.ptr for d in d_z])
    for s in range(nsets):
        assert np.array_equal(d_z[s].get(n), exp_z[s]), f"set {s}"
    assert got_last == P.fr_array_to_ints(last_z[None, :])[0]
    # the product over the unblinded rows telescopes to 1 for a genuine permutation of equal values: use equal columns
    # commitment of each z through the resident g_lagrange
    s_int = P.fr_array_to_ints(oracle.synth_scalars(0xC9, 1))[0]
    if k <= 10:
        g, g_lagrange = oracle.params_setup(k, P.fr_array_from_ints([s_int])[0])
        params = cq.ParamsKZG(k, g, g_lagrange)
        for s in range(nsets):
            got = params.commit_lagrange(exp_z[s])
            assert np.array_equal(got.to_affine(), oracle.best_multiexp(exp_z[s], g_lagrange, 2)[1])
        params.free()
    for d in d_cols + d_perms + d_z:
        d.free()


def test_permutation_product_of_a_true_permutation_closes(cq, oracle):
    """with copy-constrained cells holding equal values the grand product returns to 1 at row n - (blinding_factors + 1) —
    the identity the verifier checks (l_last * (z^2 - z), evaluation.rs:402-406)"""
    k, ncols, bf = 8, 3, 5
    n = 1 << k
    usable = n - (bf + 1)
    rng = np.random.default_rng(5)
    omega = P.omega_for(k)
    delta_pows = [pow(cq.permutation.FR_DELTA, j, P.R_MOD) for j in range(ncols)]
    omega_pows = [pow(omega, i, P.R_MOD) for i in range(n)]
    vals = [[0] * n for _ in range(ncols)]
    sig = [[delta_pows[j] * omega_pows[i] % P.R_MOD for i in range(n)] for j in range(ncols)]
    cells = [(j, i) for j in range(ncols) for i in range(usable)]
    order = rng.permutation(len(cells))
    # cycles of length 3 over random usable cells: equal values, sigma maps each cell to the next of its cycle
    for t in range(0, len(order) - 2, 3):
        cyc = [cells[int(order[t + u])] for u in range(3)]
        v = int(rng.integers(0, 1 << 60))
        for u in range(3):
            j, i = cyc[u]
            jn, i_n = cyc[(u + 1) % 3]
            vals[j][i] = v
            sig[j][i] = delta_pows[jn] * omega_pows[i_n] % P.R_MOD
    for j in range(ncols):
        for i in range(usable, n):
            vals[j][i] = int(rng.integers(0, 1 << 60))
    d_cols = [Dev(cq, P.fr_array_from_ints(v)) for v in vals]
    d_perms = [Dev(cq, P.fr_array_from_ints(s)) for s in sig]
    d_z = Dev(cq, n=n)
    blind = oracle.synth_scalars(1, bf)
    last = cq.permutation.commit_dev([d.ptr for d in d_cols], [d.ptr for d in d_perms], k, ncols + 2, bf, 0xABCDEF, 0x123456, omega, [blind], [d_z.ptr])
    assert last == 1
    for d in d_cols + d_perms + [d_z]:
        d.free()


@pytest.mark.parametrize("k,bf", [(4, 3), (9, 5), (14, 5)])
def test_plookup_commit_product_and_h_terms(cq, oracle, k, bf):
    """lookup/prover.rs:173-262 commit_product and evaluation.rs:458-531 on the device vs the oracle's restatement. The
    permuted columns are built as the reference's permute_expression_pair would (sorted input; table entries aligned to the
    first occurrence of every input value), so the product closes to 1 at row n - bf - 1."""
    n = 1 << k
    usable = n - (bf + 1)
    rng = np.random.default_rng(7 + k)
    table_vals = [int(v) for v in rng.choice(1 << 30, usable, replace=False)]
    inputs = [table_vals[int(j)] for j in rng.integers(0, max(1, usable // 3), usable)]
    perm_in = sorted(inputs)
    used = sorted(set(inputs))
    rest = [t for t in table_vals if t not in set(used)]
    perm_tab, ri = [], 0
    for i, v in enumerate(perm_in):
        if i == 0 or v != perm_in[i - 1]:
            perm_tab.append(v)
        else:
            perm_tab.append(rest[ri]); ri += 1
    assert sorted(perm_tab) == sorted(table_vals)
    pad = lambda xs: xs + [int(v) for v in rng.integers(0, 1 << 60, n - usable)]  # noqa: E731  blinded tail rows
    a, s, ap, sp = (P.fr_array_from_ints(pad(list(x))) for x in (inputs, table_vals, perm_in, perm_tab))
    beta, gamma = 0x1357913579, 0x2468024680
    L1 = lambda x: P.fr_array_from_ints([x])[0]  # noqa: E731
    exp_z = oracle.lookup_product(a, s, ap, sp, L1(beta), L1(gamma))
    assert P.fr_array_to_ints(exp_z[usable:usable + 1])[0] == 1     # the grand product closes on the usable rows
    d = [Dev(cq, x) for x in (a, s, ap, sp)]
    d_z = Dev(cq, n=n)
    cq.lookup.commit_product_dev(d[0].ptr, d[1].ptr, d[2].ptr, d[3].ptr, k, beta, gamma, d_z.ptr)
    assert np.array_equal(d_z.get(n), exp_z)
    # evaluate_h terms on an "extended domain" of 4n rows with arbitrary vectors (the arithmetic is what is checked)
    size, rot_scale = 4 * n, 4
    vecs = [oracle.synth_scalars(0x900 + i, size) for i in range(8)]  # values, table_value, product, pin, ptab, l0, l_last, l_active
    y = oracle.synth_scalars(0x99, 1)[0]
    exp_v = oracle.lookup_h(vecs[0], rot_scale, *vecs[1:], L1(beta), L1(gamma), y)
    dv = [Dev(cq, v) for v in vecs]
    cq.lookup.lookup_h_dev(*[x.ptr for x in dv], L1(beta), L1(gamma), y, size, rot_scale)
    assert np.array_equal(dv[0].get(size), exp_v)
    for x in d + [d_z] + dv:
        x.free()
