"""A COMPLETE create_proof (reference plonk/prover.rs:37-797 + poly/kzg/multiopen/gwc/prover.rs:42-86) for a my_test.rs-shaped
circuit — two advice columns, one static (CQ) lookup over two tables (halo2_proofs/tests/my_test.rs:179-259), a permutation over
both columns — through the real Blake2b transcript (tests/transcript_ref.py), twice:

  * the device prover: sha2_on_cq_halo2_b200.prover.create_proof (libcqb200 kernels, every polynomial resident in HBM),
  * the oracle prover: the same steps restated here over the CPU oracle's primitives (best_multiexp, best_fft wrappers,
    evaluate_h's permutation / CQ terms, eval_polynomial, kate_division).

With the same SRS, witness and rng values the two must emit the SAME PROOF BYTES, commitment phase, evaluations and opening
witnesses alike. The proof is then VERIFIED the way plonk/verifier.rs does, with the toxic s known instead of a pairing:
h(x) recomputed from the evaluations in the proof must be the quotient's evaluation, and every GWC opening must satisfy
C - [eval]G = [s - z]W. (The CQ pairing checks e(A, T) = e(Q_A, Z_V) e(M - beta A, 1) are verifier-side and out of scope.)"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import pyref as P  # noqa: E402
from tests.transcript_ref import Blake2bWrite  # noqa: E402

R = P.R_MOD


def F(vals):
    return P.fr_array_from_ints(vals)


def L1(x):
    return P.fr_array_from_ints([x])[0]


def I(arr):
    return P.fr_array_to_ints(np.ascontiguousarray(arr).reshape(-1, 4))


@pytest.fixture(scope="module")
def cq():
    import cqb200

    cqb200._lib.init(0)
    return cqb200


class Transcript(Blake2bWrite):
    """Blake2bWrite with the point compression bound to the oracle (test infrastructure on both sides)"""

    def __init__(self, oracle):
        super().__init__()
        self._o = oracle
        self.points = []

    def write_point(self, affine_limbs, oracle=None):
        self.points.append(np.array(affine_limbs, dtype=np.uint64, copy=True))
        super().write_point(np.ascontiguousarray(affine_limbs, dtype=np.uint64), self._o)


@pytest.mark.parametrize("k,N,A", [(3, 16, 2), (6, 64, 2), (7, 32, 5), (9, 128, 3), (12, 1024, 2)])
def test_full_proof_bytes_and_verification(cq, oracle, k, N, A):
    """A advice columns: columns 0 and 1 carry the looked-up tuple, the others repeat them (so that the permutation can tie cells of
    different columns together); cs_degree 4 => column sets of two, i.e. ceil(A / 2) permutation product polynomials chained through
    their last rows (permutation/prover.rs:82-166, the omega^last openings of :320-340)"""
    O = oracle
    from sha2_on_cq_halo2_b200 import prover as PR

    n, cs_degree = 1 << k, 4
    bf = 5 if k > 3 else 2                     # blinding factors (my_test.rs at k = 3 leaves few usable rows)
    usable = n - (bf + 1)
    rng = np.random.default_rng(77 + k)
    s_limbs = O.synth_scalars(0xF5 + k, 1)[0]
    s = I(s_limbs)[0]
    Nt = max(N, n)
    g, g_lagrange = O.params_setup(k, s_limbs)
    t_g1_big, _, _ = O.table_srs_setup(Nt, s_limbs)
    t_g1, t_lag, t_op0 = O.table_srs_setup(N, s_limbs)
    tvals = [[int(v) + (j << 40) for v in rng.choice(1 << 30, N, replace=False)] for j in range(2)]
    rows = [int(v) for v in rng.integers(0, N, usable)]
    adv = [[tv[r] for r in rows] + [int(v) for v in rng.integers(0, 1 << 50, n - usable)] for tv in tvals]
    for j in range(2, A):
        adv.append(adv[j % 2][:usable] + [int(v) for v in rng.integers(0, 1 << 50, n - usable)])
    omega = P.omega_for(k)
    delta = cq.permutation.FR_DELTA
    # sigma: identity permutation with a few cycles between equal cells (rows using the same table row hold equal values)
    sig = [[pow(delta, j, R) * pow(omega, i, R) % R for i in range(n)] for j in range(A)]
    seen = {}
    for i, r in enumerate(rows):
        if r in seen and len(seen) % 3 == 0:
            i0 = seen[r]
            for j in range(2):
                sig[j][i0], sig[j][i] = sig[j][i], sig[j][i0]
        seen.setdefault(r, i)
    for j in range(2, A):  # cross-column cycles: cell (j, i) holds the value of cell (j % 2, i) on the usable rows
        for i in range(j, usable, 4):
            sig[j][i], sig[j % 2][i] = sig[j % 2][i], sig[j][i]
    chunk_len = cs_degree - 2
    nsets = (A + chunk_len - 1) // chunk_len
    vk_repr = 0x1234ABCD + k
    rnd_poly = O.synth_scalars(0x4444 + k, n)
    blind_rows = [O.synth_scalars(0xB11D + k + 97 * s_, bf) for s_ in range(nsets)]
    b0_bound = np.ascontiguousarray(t_g1_big[Nt - (n - 1):])
    m = {}
    for r in rows:
        m[r] = m.get(r, 0) + 1
    idx = np.array(sorted(m), dtype=np.uint32)
    mult = F([m[int(i)] for i in idx])
    odom = O.domain_new(cs_degree, k)
    ek = odom.extended_k
    en = 1 << ek
    th = 4
    ext_omega_limbs = odom.f("extended_omega")
    advice_queries = [(j, 0) for j in range(A)]
    x_rot = lambda x, rot: x * pow(omega, rot, R) % R if rot >= 0 else x * pow(pow(omega, -1, R), -rot, R) % R  # noqa: E731

    def lagrange_basis(rows_set):
        v = np.zeros((n, 4), np.uint64)
        for r in rows_set:
            v[r] = L1(1)
        return O.coeff_to_extended(odom, O.lagrange_to_coeff(odom, v, th), th)

    # ------------------------------------------------------------------------------------------------- oracle prover
    def oracle_prover():
        t = Transcript(O)
        info = {}
        t.common_scalar(vk_repr)
        adv_l = [F(c) for c in adv]
        for a in adv_l:
            t.write_point(O.best_multiexp(a, g_lagrange, th)[1])
        theta = info["theta"] = t.squeeze_challenge_scalar()
        f_int = [(a0 * theta + a1) % R for a0, a1 in zip(adv[0], adv[1])]
        t.write_point(O.best_multiexp(F(f_int), g_lagrange, th)[1])
        t.write_point(O.sparse_commit(t_lag, idx, mult))
        beta = info["beta"] = t.squeeze_challenge_scalar()
        gamma = info["gamma"] = t.squeeze_challenge_scalar()
        z_polys, dw, last_z = [], L1(1), L1(1)
        for s_ in range(nsets):                                                    # permutation/prover.rs:82-186
            sl = slice(s_ * chunk_len, (s_ + 1) * chunk_len)
            z, dw = O.permutation_product(adv_l[sl], [F(sg) for sg in sig[sl]], L1(beta), L1(gamma), L1(omega), dw, last_z)
            z[n - bf:] = blind_rows[s_]
            last_z = z[n - bf - 1].copy()
            t.write_point(O.best_multiexp(z, g_lagrange, th)[1])
            z_polys.append(O.lagrange_to_coeff(odom, z, th))
        # commit_log_derivatives (static_lookup/prover.rs:187-342), the reference's per-index loop
        qs_host = [O.cq_table_qs(F(v), t_g1, th) for v in tvals]
        a_acc = qa_acc = a0_acc = None
        add = lambda acc, p: p if acc is None else O.g1_add_jj(acc, p)  # noqa: E731
        for i in idx:
            i = int(i)
            values = (tvals[0][i] * theta + tvals[1][i]) % R
            qs = O.g1_to_affine(O.g1_add_ja(O.g1_mul_a(qs_host[0][i], L1(theta)), qs_host[1][i]))
            a_i = L1(m[i] * pow((values + beta) % R, -1, R) % R)
            a_acc, qa_acc, a0_acc = add(a_acc, O.g1_mul_a(t_lag[i], a_i)), add(qa_acc, O.g1_mul_a(qs, a_i)), add(a0_acc, O.g1_mul_a(t_op0[i], a_i))
        beta_inv = pow(beta, -1, R)
        bs = F([pow((fv + beta) % R, -1, R) for fv in f_int[:usable]] + [beta_inv] * (bf + 1))
        b_poly = O.lagrange_to_coeff(odom, bs, th)
        b0 = np.ascontiguousarray(b_poly[1:])
        p_cm = O.best_multiexp(b0, b0_bound, th)[1]
        b0_poly = np.concatenate([b0, np.zeros((1, 4), np.uint64)])
        b0_cm = O.best_multiexp(b0_poly, g, th)[1]
        for acc in (a_acc, qa_acc, a0_acc):
            t.write_point(O.g1_to_affine(acc))
        t.write_point(b0_cm)
        t.write_point(p_cm)
        b_at_zero = I(b_poly[:1])[0]
        a_at_zero = (b_at_zero * n - (bf + 1) * beta_inv) * pow(N, -1, R) % R
        f_poly = O.lagrange_to_coeff(odom, F(f_int), th)
        t.write_point(O.best_multiexp(rnd_poly, g, th)[1])
        y = info["y"] = t.squeeze_challenge_scalar()
        # h(X): evaluate_h's permutation and CQ terms on the extended domain (evaluation.rs:376-452, 533-548)
        adv_poly = [O.lagrange_to_coeff(odom, a, th) for a in adv_l]
        adv_coset = [O.coeff_to_extended(odom, p, th) for p in adv_poly]
        sig_poly = [O.lagrange_to_coeff(odom, F(sg), th) for sg in sig]
        sig_coset = [O.coeff_to_extended(odom, p, th) for p in sig_poly]
        l0, l_last = lagrange_basis([0]), lagrange_basis([n - bf - 1])
        l_blind = lagrange_basis(range(n - bf, n))
        one = L1(1)
        l_act = np.stack([O.fr_op("sub", one, O.fr_op("add", a, b)) for a, b in zip(l_last, l_blind)])
        h = np.zeros((en, 4), np.uint64)
        h = O.permutation_h(h, 1 << (ek - k), -(bf + 1), chunk_len, [O.coeff_to_extended(odom, zp, th) for zp in z_polys], adv_coset, sig_coset, l0,
                            l_last, l_act, L1(beta), L1(gamma), L1(y), ext_omega_limbs)
        h = O.cq_lookup_h(h, O.coeff_to_extended(odom, b_poly, th), O.coeff_to_extended(odom, f_poly, th), l_act, L1(beta), L1(y))
        h_coeff = O.extended_to_coeff(odom, O.divide_by_vanishing_poly(odom, h), th)
        pieces = [np.ascontiguousarray(h_coeff[i * n:(i + 1) * n]) for i in range(cs_degree - 1)]
        for pc in pieces:
            t.write_point(O.best_multiexp(pc, g, th)[1])
        x = info["x"] = t.squeeze_challenge_scalar()
        xn = pow(x, n, R)
        ev = lambda poly, pt: I(O.eval_polynomial(poly, L1(pt)).reshape(1, 4))[0]  # noqa: E731
        adv_evals = [ev(adv_poly[c], x_rot(x, rot)) for c, rot in advice_queries]
        for e in adv_evals:
            t.write_scalar(e)
        hx_int = [0] * n
        for pc in pieces[::-1]:
            hx_int = [(a * xn + b) % R for a, b in zip(hx_int, I(pc))]
        h_x_poly = F(hx_int)
        random_eval = ev(rnd_poly, x)
        t.write_scalar(random_eval)
        sigma_evals = [ev(p, x) for p in sig_poly]
        for e in sigma_evals:
            t.write_scalar(e)
        x_next, x_last = x_rot(x, 1), x_rot(x, -(bf + 1))
        z_evals = []
        for s_, zp in enumerate(z_polys):                                          # permutation/prover.rs:244-288
            ze = [ev(zp, x), ev(zp, x_next), ev(zp, x_last) if s_ + 1 < nsets else None]
            for e_ in ze:
                if e_ is not None:
                    t.write_scalar(e_)
            z_evals.append(tuple(ze))
        b0_eval, f_eval = ev(b0_poly, x), ev(f_poly, x)
        for e in (b0_eval, f_eval, a_at_zero):
            t.write_scalar(e)
        h_eval = info["h_eval"] = ev(h_x_poly, x)
        queries = [(x_rot(x, rot), adv_poly[c], e) for (c, rot), e in zip(advice_queries, adv_evals)]
        for zp, ze in zip(z_polys, z_evals):                                       # permutation/prover.rs:304-340
            queries += [(x, zp, ze[0]), (x_next, zp, ze[1])]
        for zp, ze in list(zip(z_polys, z_evals))[::-1][1:]:
            queries.append((x_last, zp, ze[2]))
        queries += [(x, b0_poly, b0_eval), (x, f_poly, f_eval)]
        queries += [(x, p, e) for p, e in zip(sig_poly, sigma_evals)]
        queries += [(x, h_x_poly, h_eval), (x, rnd_poly, random_eval)]
        v = info["v"] = t.squeeze_challenge_scalar()
        sets = []
        for q in queries:
            for ps in sets:
                if ps[0] == q[0]:
                    ps[1].append(q)
                    break
            else:
                sets.append((q[0], [q]))
        for zpt, qs in sets:
            batch, eb, pw = [0] * n, 0, 1
            for _, poly, e in qs:
                batch = [(a + pw * b) % R for a, b in zip(batch, I(poly))]
                eb = (eb + pw * e) % R
                pw = pw * v % R
            batch[0] = (batch[0] - eb) % R
            wit = O.kate_division(F(batch), L1(zpt))
            t.write_point(O.best_multiexp(np.ascontiguousarray(wit), g[: n - 1], th)[1])
        info["evals"] = dict(advice=adv_evals, random=random_eval, sigma=sigma_evals, z=z_evals, lk=(b0_eval, f_eval, a_at_zero))
        return t, info

    # ------------------------------------------------------------------------------------------------- device prover
    def device_prover():
        params = cq.ParamsKZG(k, g, g_lagrange)
        tsrs = cq.TableSRS.setup_from_toxic_waste(N - 1, s_limbs, precompute=False)
        tables = [cq.cq.StaticTableValues(F(v), tsrs.g1) for v in tvals]
        bound = cq.DeviceBases(b0_bound)
        lk = PR.StaticLookup([0, 1], tsrs, tables, bound)
        pk = PR.ProvingKey(params, k, cs_degree, bf, list(range(A)), [F(sg) for sg in sig], advice_queries, [lk], vk_transcript_repr=vk_repr)
        t = Transcript(O)
        try:
            info = PR.create_proof(pk, [F(c) for c in adv], [(idx, mult)], {"permutation_blinds": blind_rows, "random_poly": rnd_poly}, t)
            info["expected_h"] = PR.expected_h_eval(pk, info)   # the verifier's side of the quotient identity, from the product module
        finally:
            pk.free()
            for tb in tables:
                tb.free()
            bound.free()
            params.free()
            tsrs.free()
        return t, info

    t_o, info_o = oracle_prover()
    t_d, info_d = device_prover()
    for name in ("theta", "beta", "gamma", "y", "x", "v", "h_eval"):
        assert info_d[name] == info_o[name], name
    npts_open = 2 + (1 if nsets > 1 else 0)                                    # distinct opening points: x, omega x, omega^last x
    npoints = A + 2 + nsets + 5 + 1 + (cs_degree - 1) + npts_open
    nscalars = A + 1 + A + (3 * nsets - 1) + 3
    assert len(t_o.proof) == 32 * (npoints + nscalars)
    assert bytes(t_d.proof) == bytes(t_o.proof), "proof bytes differ"

    # ------------------------------------------------------------------------------------------------- verification
    x, y, v, beta, gamma = (info_d[c] for c in ("x", "y", "v", "beta", "gamma"))
    e = info_o["evals"]
    xn = pow(x, n, R)
    inv = lambda a: pow(a, -1, R)  # noqa: E731
    lag_at = lambda i: (xn - 1) * inv(n) % R * pow(omega, i, R) % R * inv((x - pow(omega, i, R)) % R) % R  # noqa: E731
    l0_x, l_last_x = lag_at(0), lag_at(n - bf - 1)
    l_blind_x = sum(lag_at(i) for i in range(n - bf, n)) % R
    l_act_x = (1 - (l_last_x + l_blind_x)) % R
    a_x, sg_x, zs = e["advice"], e["sigma"], e["z"]
    b0_x, f_x, a_at_zero = e["lk"]
    exp = 0
    exp = (exp * y + l0_x * (1 - zs[0][0])) % R                                  # plonk/permutation/verifier.rs expressions
    exp = (exp * y + l_last_x * (zs[-1][0] * zs[-1][0] - zs[-1][0])) % R
    for i in range(1, nsets):
        exp = (exp * y + l0_x * (zs[i][0] - zs[i - 1][2])) % R
    cur = beta * x % R
    for s_ in range(nsets):
        left, right = zs[s_][1], zs[s_][0]
        for j in range(s_ * chunk_len, min(A, (s_ + 1) * chunk_len)):
            left = left * (a_x[j] + beta * sg_x[j] + gamma) % R
            right = right * (a_x[j] + cur + gamma) % R
            cur = cur * delta % R
        exp = (exp * y + (left - right) * l_act_x) % R
    b_at_zero = (a_at_zero * N + (bf + 1) * inv(beta)) % R * inv(n) % R       # the sumcheck identity n B(0) = N A(0), solved for B(0)
    b_x = (b0_x * x + b_at_zero) % R
    exp = (exp * y + (b_x * (f_x * l_act_x + beta) - 1)) % R
    h_x_expected = exp * inv((xn - 1) % R) % R
    assert h_x_expected == info_d["h_eval"], "the quotient identity fails at x"
    assert info_d["expected_h"] == h_x_expected, "prover.expected_h_eval disagrees with the verifier restated here"
    # GWC openings with the toxic s: sum_i v^i C_i - [sum_i v^i e_i] G == [s - z] W
    pts = [P.g1_affine_to_ints(p.reshape(1, 8))[0] for p in t_d.points]
    adv_cm, f_cm = pts[0:A], pts[A]
    z_cm = pts[A + 2:A + 2 + nsets]
    o = A + 2 + nsets
    b0_cm, rnd_cm = pts[o + 3], pts[o + 5]
    h_cms = pts[o + 6:o + 6 + cs_degree - 1]
    wit = pts[o + 6 + cs_degree - 1:]
    sigma_cm = [P.g1_affine_to_ints(O.best_multiexp(F(sg), g_lagrange, th)[1].reshape(1, 8))[0] for sg in sig]  # the vk's permutation commitments
    h_cm = None
    for hc in h_cms[::-1]:
        h_cm = P.g1_add(P.g1_mul(h_cm, xn) if h_cm is not None else None, hc)
    x_next, x_last = x_rot(x, 1), x_rot(x, -(bf + 1))
    q_at_x = [(adv_cm[j], a_x[j]) for j in range(A)] + [(z_cm[s_], zs[s_][0]) for s_ in range(nsets)]
    q_at_x += [(b0_cm, b0_x), (f_cm, f_x)] + [(sigma_cm[j], sg_x[j]) for j in range(A)] + [(h_cm, h_x_expected), (rnd_cm, e["random"])]
    q_at_next = [(z_cm[s_], zs[s_][1]) for s_ in range(nsets)]
    q_at_last = [(z_cm[s_], zs[s_][2]) for s_ in range(nsets - 2, -1, -1)]
    groups = [(x, q_at_x), (x_next, q_at_next)] + ([(x_last, q_at_last)] if nsets > 1 else [])
    assert len(wit) == len(groups)
    # the order of the queries INSIDE the point set of x follows the prover's chain (advice, then per set z at x, ...): rebuild it
    # from the interleaved chain instead of the grouped lists above
    chain = [(x, adv_cm[j], a_x[j]) for j in range(A)]
    for s_ in range(nsets):
        chain += [(x, z_cm[s_], zs[s_][0]), (x_next, z_cm[s_], zs[s_][1])]
    for s_ in range(nsets - 2, -1, -1):
        chain.append((x_last, z_cm[s_], zs[s_][2]))
    chain += [(x, b0_cm, b0_x), (x, f_cm, f_x)] + [(x, sigma_cm[j], sg_x[j]) for j in range(A)] + [(x, h_cm, h_x_expected), (x, rnd_cm, e["random"])]
    sets_v = []
    for q in chain:
        for ps in sets_v:
            if ps[0] == q[0]:
                ps[1].append(q)
                break
        else:
            sets_v.append((q[0], [q]))
    assert [z_ for z_, _ in sets_v] == [z_ for z_, _ in groups]
    for (zpt, qs), w in zip(sets_v, wit):
        c_acc, e_acc, pw = None, 0, 1
        for _, cm, ev_ in qs:
            c_acc = P.g1_add(c_acc, P.g1_mul(cm, pw))
            e_acc = (e_acc + pw * ev_) % R
            pw = pw * v % R
        lhs = P.g1_add(c_acc, P.g1_neg(P.g1_mul(P.G1_GEN, e_acc)))
        rhs = P.g1_mul(w, (s - zpt) % R)
        assert lhs == rhs, "KZG opening check failed"
