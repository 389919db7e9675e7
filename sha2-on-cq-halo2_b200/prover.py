"""Mirror of halo2_proofs::plonk::create_proof (reference halo2_proofs/src/plonk/prover.rs:37-797) for circuits made of advice
columns, an optional custom-gate program, one permutation argument and static (CQ) lookups, with the KZG / GWC backend
(poly/kzg/multiopen/gwc/prover.rs:42-86) — the call path the sha crate's CQ circuits take. Every polynomial lives in HBM
from the witness upload to the last opening witness; what crosses PCIe is the witness in, and 32-byte commitments and
evaluations out. The transcript object is the caller's (the reference's Blake2bWrite, transcript.rs:199-240, restated in
tests/transcript_ref.py): this module only calls write_point / write_scalar / squeeze_challenge_scalar in the reference's
order, which is what makes the proof bytes identical.

Not here (CPU work outside the hot path, done by the caller exactly as in the reference): witness synthesis, the m_sparse
map of the lookup (static_lookup/prover.rs:123-160), the rng (blinding rows, the vanishing argument's random polynomial).
"""
import ctypes

import numpy as np

from . import _lib, cq, permutation
from .domain import EvaluationDomain
from .evaluation import cq_lookup_h_dev, permutation_h_dev
from .fields import R_MOD, fr_from_limbs, fr_to_limbs
from .kzg import DeviceBases


def _vp(p):
    return ctypes.c_void_p(p)


class _Arena:
    """Device buffers of one object, released together. With a pool (one device allocation kept by the proving key and reused by
    every proof) an allocation is a pointer bump: cudaMalloc / cudaFree cost ~0.4-1 ms each and a proof makes ~40 of them — at
    circuit sizes that was most of the time of a proof (profiles/r02_summary.md)."""

    def __init__(self, pool=None, pool_bytes=0):
        self.ptrs = []
        self.pool, self.pool_bytes, self.used = pool, pool_bytes, 0

    def alloc(self, nbytes):
        nbytes = max(nbytes, 64)
        if self.pool is not None:
            start = (self.used + 255) & ~255
            if start + nbytes <= self.pool_bytes:
                self.used = start + nbytes
                return self.pool + start
        d = ctypes.c_void_p()
        _lib.check(_lib.lib().cqb_dev_alloc(nbytes, ctypes.byref(d)))
        self.ptrs.append(d)
        return d.value

    def upload(self, arr):
        arr = np.ascontiguousarray(arr, dtype=np.uint64)
        d = self.alloc(arr.nbytes)
        if arr.nbytes:
            _lib.check(_lib.lib().cqb_memcpy_h2d(_vp(d), arr.ctypes.data_as(ctypes.c_void_p), arr.nbytes))
        return d

    def free(self):
        if self.pool is not None:
            _lib.check(_lib.lib().cqb_sync())  # the pool is reused by the next proof: everything queued on it must have finished
        for d in self.ptrs:
            _lib.check(_lib.lib().cqb_dev_free(d))
        self.ptrs = []
        self.used = 0


def _commit(bases, d_ptr, count):
    out = np.zeros(8, np.uint64)
    inf = ctypes.c_int(0)
    _lib.check(_lib.lib().cqb_msm_bn254_g1_dev(bases.handle, 0, _vp(d_ptr), count, _lib.p64(out), ctypes.byref(inf)))
    return out


def _commit_batch(bases, d_ptr, count, batch):
    """`batch` commitments over vectors stored back to back (plonk/prover.rs:356-360 commits every advice column in a loop; the h pieces
    likewise, vanishing/prover.rs:101-105): one launch sequence for all of them"""
    outs = []
    done = 0
    while done < batch:
        b = min(64, batch - done)
        out = np.zeros((b, 8), np.uint64)
        inf = (ctypes.c_int * b)()
        _lib.check(_lib.lib().cqb_msm_bn254_g1_batch_dev(bases.handle, 0, _vp(d_ptr + done * count * 32), count, b, _lib.p64(out), inf))
        outs.extend(out[i].copy() for i in range(b))
        done += b
    return outs


def _eval(d_poly, n, point):
    out = np.zeros(4, np.uint64)
    _lib.check(_lib.lib().cqb_eval_polynomial_dev(_vp(d_poly), n, _lib.p64(fr_to_limbs(point)), _lib.p64(out)))
    return fr_from_limbs(out)


def _eval_many(queries, n):
    """[(device polynomial, point)] -> evaluations, one device read-back for all of them"""
    if not queries:
        return []
    cnt = len(queries)
    ptrs = (ctypes.c_void_p * cnt)(*[_vp(p) for p, _ in queries])
    pts = np.ascontiguousarray(np.stack([fr_to_limbs(x) for _, x in queries]))
    out = np.zeros((cnt, 4), np.uint64)
    _lib.check(_lib.lib().cqb_eval_polynomials_dev(ptrs, n, _lib.p64(pts), cnt, _lib.p64(out)))
    return [fr_from_limbs(out[i]) for i in range(cnt)]


class StaticLookup:
    """one lookup_static of the constraint system: the advice columns whose theta-compression is looked up, the tables it is
    looked up in (plonk/static_lookup.rs:69-126) and the table SRS (poly/kzg/commitment.rs:42-47)"""

    def __init__(self, input_columns, table_srs, tables, b0_g1_bound):
        self.input_columns, self.table_srs, self.tables, self.b0_g1_bound = list(input_columns), table_srs, list(tables), b0_g1_bound


class ProvingKey:
    """The device-resident part of plonk::ProvingKey that create_proof reads (plonk/keygen.rs:300-400): the permutation's
    sigma polynomials in Lagrange / coefficient / extended form, l0, l_last, l_active_row on the extended domain, the
    evaluation domain, and the queries of the constraint system."""

    def __init__(self, params, k, cs_degree, blinding_factors, permutation_columns, sigma_lagrange, advice_queries, static_lookups=(),
                 vk_transcript_repr=0):
        lib = _lib.lib()
        self.params, self.k, self.n = params, k, 1 << k
        self.cs_degree, self.blinding_factors = cs_degree, blinding_factors
        self.domain = dom = EvaluationDomain(cs_degree, k)
        self.permutation_columns = list(permutation_columns)  # advice column indices, in cs.permutation.columns order
        self.advice_queries = list(advice_queries)            # (column, rotation) in cs.advice_queries order
        self.static_lookups = list(static_lookups)
        self.vk_transcript_repr = vk_transcript_repr
        self._pool, self._pool_bytes = None, 0
        self._arena = ar = _Arena()
        n, en = self.n, dom.extended_len()

        def coeff_and_coset(lagrange):
            d_c = ar.upload(lagrange)
            _lib.check(lib.cqb_intt_bn254_fr_dev(_vp(d_c), _lib.p64(dom.omega_inv), _lib.p64(dom.ifft_divisor), k))
            d_e = ar.alloc(en * 32)
            _lib.check(lib.cqb_coset_ntt_bn254_fr_dev(_vp(d_c), n, _vp(d_e), _lib.p64(dom.extended_omega), dom.extended_k, _lib.p64(dom.g_coset),
                                                      _lib.p64(dom.g_coset_inv)))
            return d_c, d_e

        # permutation::ProvingKey { permutations, polys, cosets } (plonk/permutation/keygen.rs:180-220)
        self.sigma_lagrange = [ar.upload(s) for s in sigma_lagrange]
        pc = [coeff_and_coset(s) for s in sigma_lagrange]
        self.sigma_polys, self.sigma_cosets = [p[0] for p in pc], [p[1] for p in pc]
        # l0, l_blind, l_last (keygen.rs:344-363), l_active_row = 1 - (l_last + l_blind) on the extended domain (:367-373)
        one = fr_to_limbs(1)
        lag = np.zeros((n, 4), np.uint64)
        lag[0] = one
        _, self.l0 = coeff_and_coset(lag)
        lag[:] = 0
        if blinding_factors:
            lag[n - blinding_factors:] = one
        _, d_lblind = coeff_and_coset(lag)
        lag[:] = 0
        lag[n - blinding_factors - 1] = one
        _, self.l_last = coeff_and_coset(lag)
        self.l_active_row = ar.upload(np.tile(one, (en, 1)))
        # d_lblind <- l_last + l_blind ; l_active_row <- ones ... then (l_last + l_blind) * (-1) + ones
        _lib.check(lib.cqb_fr_axpy_dev(_vp(d_lblind), _lib.p64(one), _vp(self.l_last), en))
        _lib.check(lib.cqb_fr_axpy_dev(_vp(d_lblind), _lib.p64(fr_to_limbs(R_MOD - 1)), _vp(self.l_active_row), en))
        _lib.check(lib.cqb_memcpy_d2d(_vp(self.l_active_row), _vp(d_lblind), en * 32))
        _lib.check(lib.cqb_sync())

    def rotate_omega(self, x, rot):
        """poly/domain.rs:414-424"""
        w = self.domain._omega
        return x * pow(w, rot, R_MOD) % R_MOD if rot >= 0 else x * pow(pow(w, -1, R_MOD), -rot, R_MOD) % R_MOD

    def proof_pool(self, n_advice):
        """the working memory of one create_proof, allocated once and reused: every polynomial a proof holds in HBM at the same time"""
        n, en = self.n, self.domain.extended_len()
        chunk = self.cs_degree - 2
        nsets = (len(self.permutation_columns) + chunk - 1) // chunk if self.permutation_columns else 0
        L = len(self.static_lookups)
        # n-sized: advice (Lagrange + coefficients), f, z (Lagrange + coefficients), the CQ argument's b / b0 / f / a / t / m, random
        # poly, h(X), batch, witness; extended: z, advice, b, f cosets and the quotient
        elems = n * (2 * n_advice + 2 * nsets + 8 * L + 6) + en * (nsets + n_advice + 2 * L + 1)
        need = elems * 32 + (1 << 20)
        if need > self._pool_bytes:
            if self._pool is not None:
                _lib.check(_lib.lib().cqb_dev_free(_vp(self._pool)))
            d = ctypes.c_void_p()
            _lib.check(_lib.lib().cqb_dev_alloc(need, ctypes.byref(d)))
            self._pool, self._pool_bytes = d.value, need
        return self._pool, self._pool_bytes

    def free(self):
        self._arena.free()
        if self._pool is not None:
            _lib.check(_lib.lib().cqb_dev_free(_vp(self._pool)))
            self._pool, self._pool_bytes = None, 0


def create_proof(pk, advice_lagrange, lookups_m_sparse, rng, transcript):
    """plonk/prover.rs:37-797 for ONE circuit instance.

    advice_lagrange   : one (n, 4) uint64 array per advice column (Lagrange values, blinding rows already drawn by the caller)
    lookups_m_sparse  : per static lookup, (idx uint32 array, multiplicities (m, 4) array) — the m_sparse map in key order
    rng               : {"permutation_blinds": [per column set, (blinding_factors, 4)], "random_poly": (n, 4)} — the values the
                        reference draws from its RngCore (permutation/prover.rs:152-155, vanishing/prover.rs:46-55)
    transcript        : write_point(affine limbs) / write_scalar(int) / squeeze_challenge_scalar() / common_scalar(int)
    Returns the challenges and evaluations (for tests); the proof is whatever the transcript wrote."""
    lib = _lib.lib()
    dom, params = pk.domain, pk.params
    k, n, en, bf = pk.k, pk.n, pk.domain.extended_len(), pk.blinding_factors
    ar = _Arena(*pk.proof_pool(len(advice_lagrange)))
    info = {}
    try:
        def to_coeff(d_lagrange):
            d_c = ar.alloc(n * 32)
            _lib.check(lib.cqb_memcpy_d2d(_vp(d_c), _vp(d_lagrange), n * 32))
            _lib.check(lib.cqb_intt_bn254_fr_dev(_vp(d_c), _lib.p64(dom.omega_inv), _lib.p64(dom.ifft_divisor), k))
            return d_c

        def to_extended(d_coeff):
            d_e = ar.alloc(en * 32)
            _lib.check(lib.cqb_coset_ntt_bn254_fr_dev(_vp(d_coeff), n, _vp(d_e), _lib.p64(dom.extended_omega), dom.extended_k, _lib.p64(dom.g_coset),
                                                      _lib.p64(dom.g_coset_inv)))
            return d_e

        transcript.common_scalar(pk.vk_transcript_repr)                                   # prover.rs:85
        d_adv0 = ar.upload(np.concatenate([np.ascontiguousarray(a, dtype=np.uint64).reshape(n, 4) for a in advice_lagrange]))
        d_adv = [d_adv0 + i * n * 32 for i in range(len(advice_lagrange))]
        for pt in _commit_batch(params.g_lagrange, d_adv0, n, len(d_adv)):                 # :356-374 advice commitments
            transcript.write_point(pt)
        theta = info["theta"] = transcript.squeeze_challenge_scalar()                      # :472
        # static lookups, first phase: f and m (static_lookup/prover.rs:51-184)
        d_f = []
        for lk, (idx, mult) in zip(pk.static_lookups, lookups_m_sparse):
            ptrs = (ctypes.c_void_p * len(lk.input_columns))(*[_vp(d_adv[c]) for c in lk.input_columns])
            d = ar.alloc(n * 32)
            _lib.check(lib.cqb_fr_compress_dev(ptrs, len(lk.input_columns), None, n, _lib.p64(fr_to_limbs(theta)), _vp(d)))  # :108-121
            d_f.append(d)
            transcript.write_point(_commit(params.g_lagrange, d, n))                       # f_cm :165, :174
            # m_cm (:167-175): one sparse MSM over the support, handed over as the (index, multiplicity) arrays in key order
            transcript.write_point(lk.table_srs.g1_lagrange.msm_sparse(idx, mult).to_affine())
        beta = info["beta"] = transcript.squeeze_challenge_scalar()                        # :529
        gamma = info["gamma"] = transcript.squeeze_challenge_scalar()                      # :532
        # permutation argument (permutation/prover.rs:46-200)
        chunk_len = pk.cs_degree - 2
        ncols = len(pk.permutation_columns)
        nsets = (ncols + chunk_len - 1) // chunk_len if ncols else 0
        d_z0 = ar.alloc(max(nsets, 1) * n * 32)                                          # the z's back to back: committed in one batch
        d_z = [d_z0 + i * n * 32 for i in range(nsets)]
        z_poly, z_coset = [], []
        if nsets:
            permutation.commit_dev([d_adv[c] for c in pk.permutation_columns], pk.sigma_lagrange, k, pk.cs_degree, bf, beta, gamma, dom._omega,
                                   rng["permutation_blinds"], d_z)
            for pt in _commit_batch(params.g_lagrange, d_z0, n, nsets):                    # :166-186
                transcript.write_point(pt)
            for d in d_z:
                z_poly.append(to_coeff(d))                                                 # :168
            z_coset = [to_extended(d) for d in z_poly]                                     # :171
        # static lookups, second phase (static_lookup/prover.rs:187-342)
        clds = []
        for lk, d, (idx, mult) in zip(pk.static_lookups, d_f, lookups_m_sparse):
            cld = cq.commit_log_derivatives_dev(params, lk.table_srs, lk.tables, lk.b0_g1_bound, k, bf, d, idx, mult, beta, theta, alloc=ar.alloc)
            clds.append(cld)
            for pt in (cld.a_cm, cld.qa_cm, cld.a0_cm, cld.b0_cm, cld.p_cm):               # :301-313
                transcript.write_point(pt.to_affine())
        # vanishing argument: random polynomial (vanishing/prover.rs:37-65)
        d_rnd = ar.upload(rng["random_poly"])
        transcript.write_point(_commit(params.g, d_rnd, n))
        y = info["y"] = transcript.squeeze_challenge_scalar()                              # prover.rs:584
        # advice polys and h(X) (prover.rs:587-624, evaluation.rs:285-551)
        adv_poly = [to_coeff(d) for d in d_adv]
        adv_coset = [to_extended(d) for d in adv_poly]
        d_h = ar.upload(np.zeros((en, 4), np.uint64))
        rot_scale = 1 << (dom.extended_k - k)
        if nsets:
            permutation_h_dev(d_h, en, rot_scale, -(bf + 1), chunk_len, z_coset, [adv_coset[c] for c in pk.permutation_columns], pk.sigma_cosets,
                              pk.l0, pk.l_last, pk.l_active_row, fr_to_limbs(beta), fr_to_limbs(gamma), fr_to_limbs(y), dom.extended_omega)
        for cld in clds:                                                                   # evaluation.rs:533-548
            b_coset, f_coset = to_extended(cld.d_b), to_extended(cld.d_f)
            cq_lookup_h_dev(d_h, b_coset, f_coset, pk.l_active_row, fr_to_limbs(beta), fr_to_limbs(y), en)
        # vanishing construct (vanishing/prover.rs:69-120): divide by t(X), back to coefficients, commit the pieces
        _lib.check(lib.cqb_coset_intt_bn254_fr_dev(_vp(d_h), dom.extended_k, _lib.p64(dom.extended_omega_inv), _lib.p64(dom.extended_ifft_divisor),
                                                   _lib.p64(dom.g_coset), _lib.p64(dom.g_coset_inv), _lib.p64(dom.t_evaluations),
                                                   dom.t_evaluations.shape[0]))
        npieces = dom.quotient_poly_degree
        for pt in _commit_batch(params.g, d_h, n, npieces):
            transcript.write_point(pt)
        x = info["x"] = transcript.squeeze_challenge_scalar()                              # prover.rs:627
        xn = pow(x, n, R_MOD)
        evals = info["evals"] = {}
        # every evaluation the proof carries (and h(x), which it does not) is queued at once and read back together; they are written
        # to the transcript in the reference's order: advice :652-670, random_eval (vanishing/prover.rs:123-157), the sigma polys
        # (permutation/prover.rs:229-241), z at x / omega x / omega^last x (:244-288), b0, f, A(0) (static_lookup/prover.rs:346-375)
        d_hx = ar.alloc(n * 32)                                                            # h(X) = sum_i h_i(X) xn^i, Horner over the pieces
        _lib.check(lib.cqb_memcpy_d2d(_vp(d_hx), _vp(d_h + (npieces - 1) * n * 32), n * 32))
        for i in range(npieces - 2, -1, -1):
            _lib.check(lib.cqb_fr_axpy_dev(_vp(d_hx), _lib.p64(fr_to_limbs(xn)), _vp(d_h + i * n * 32), n))
        x_next, x_last = pk.rotate_omega(x, 1), pk.rotate_omega(x, -(bf + 1))
        ev_q = [(adv_poly[c], pk.rotate_omega(x, rot)) for c, rot in pk.advice_queries]
        ev_q.append((d_rnd, x))
        ev_q += [(p, x) for p in pk.sigma_polys]
        for s_, zp in enumerate(z_poly):
            ev_q += [(zp, x), (zp, x_next)] + ([(zp, x_last)] if s_ + 1 < nsets else [])
        for cld in clds:
            ev_q += [(cld.d_b0, x), (cld.d_f, x)]
        ev_q.append((d_hx, x))
        ev = _eval_many(ev_q, n)
        pos = 0
        adv_evals = ev[pos:pos + len(pk.advice_queries)]
        pos += len(pk.advice_queries)
        for e in adv_evals:
            transcript.write_scalar(e)
        random_eval = ev[pos]
        pos += 1
        transcript.write_scalar(random_eval)
        sigma_evals = ev[pos:pos + len(pk.sigma_polys)]
        pos += len(pk.sigma_polys)
        for e in sigma_evals:
            transcript.write_scalar(e)
        z_evals = []
        for s_ in range(len(z_poly)):
            e_cur, e_next = ev[pos], ev[pos + 1]
            pos += 2
            transcript.write_scalar(e_cur)
            transcript.write_scalar(e_next)
            e_last = None
            if s_ + 1 < nsets:
                e_last = ev[pos]
                pos += 1
                transcript.write_scalar(e_last)
            z_evals.append((e_cur, e_next, e_last))
        lk_evals = []
        for cld in clds:
            b0_eval, f_eval = ev[pos], ev[pos + 1]
            pos += 2
            for e in (b0_eval, f_eval, cld.a_at_zero):
                transcript.write_scalar(e)
            lk_evals.append((b0_eval, f_eval, cld.a_at_zero))
        h_eval_batched = ev[pos]
        evals.update(advice=adv_evals, random=random_eval, sigma=sigma_evals, z=z_evals, static_lookups=lk_evals)
        # the queries, in the order prover.rs:718-774 chains them: (point, polynomial, evaluation)
        queries = [(pk.rotate_omega(x, rot), adv_poly[c], e) for (c, rot), e in zip(pk.advice_queries, adv_evals)]
        for zp, (e_cur, e_next, _) in zip(z_poly, z_evals):                                # permutation open :291-340
            queries += [(x, zp, e_cur), (x_next, zp, e_next)]
        for zp, (_, _, e_last) in list(zip(z_poly, z_evals))[::-1][1:]:
            queries.append((x_last, zp, e_last))
        for cld, (b0_eval, f_eval, _) in zip(clds, lk_evals):                              # static_lookup open :378-400
            queries += [(x, cld.d_b0, b0_eval), (x, cld.d_f, f_eval)]
        queries += [(x, p, e) for p, e in zip(pk.sigma_polys, sigma_evals)]                # pk.permutation.open :220-227
        h_eval = info["h_eval"] = h_eval_batched
        queries += [(x, d_hx, h_eval), (x, d_rnd, random_eval)]                            # vanishing open :160-173
        # GWC multi-open (poly/kzg/multiopen/gwc/prover.rs:42-86)
        v = info["v"] = transcript.squeeze_challenge_scalar()
        point_sets = []                                                                    # construct_intermediate_sets (gwc.rs:36-60)
        for q in queries:
            for ps in point_sets:
                if ps[0] == q[0]:
                    ps[1].append(q)
                    break
            else:
                point_sets.append((q[0], [q]))
        info["point_sets"] = [(z, [(e) for _, _, e in qs]) for z, qs in point_sets]
        d_batch = ar.alloc(n * 32)
        d_wit0 = ar.upload(np.zeros((len(point_sets) * n, 4), np.uint64))                  # witness polynomials, n - 1 coefficients each + a zero
        tmp = np.zeros(4, np.uint64)
        v_l = fr_to_limbs(v)
        for j, (z, qs) in enumerate(point_sets):
            # poly_batch = sum_i v^i p_i by Horner from the last query; eval_batch likewise
            _lib.check(lib.cqb_memcpy_d2d(_vp(d_batch), _vp(qs[-1][1]), n * 32))
            eval_batch = qs[-1][2]
            for _, d_p, e in qs[-2::-1]:
                _lib.check(lib.cqb_fr_axpy_dev(_vp(d_batch), _lib.p64(v_l), _vp(d_p), n))
                eval_batch = (eval_batch * v + e) % R_MOD
            # poly_batch - eval_batch touches the constant coefficient only
            _lib.check(lib.cqb_memcpy_d2h(tmp.ctypes.data_as(ctypes.c_void_p), _vp(d_batch), 32))
            _lib.check(lib.cqb_sync())
            c0 = fr_to_limbs((fr_from_limbs(tmp) - eval_batch) % R_MOD)
            _lib.check(lib.cqb_memcpy_h2d(_vp(d_batch), c0.ctypes.data_as(ctypes.c_void_p), 32))
            _lib.check(lib.cqb_kate_division_dev(_vp(d_batch), n, _lib.p64(fr_to_limbs(z)), _vp(d_wit0 + j * n * 32)))   # arithmetic.rs:351-387
        for pt in _commit_batch(params.g, d_wit0, n, len(point_sets)):                     # gwc/prover.rs:79-84, all witnesses in one batch
            transcript.write_point(pt)
        for cld in clds:
            cld.free()
        info["table_sizes"] = [lk.tables[0].size for lk in pk.static_lookups]
        return info
    finally:
        ar.free()


def expected_h_eval(pk, info):
    """What plonk/verifier.rs computes from the evaluations in the proof: the value h(x) must have for the quotient identity
    h(X) (X^n - 1) = sum of the constraint terms folded with y to hold at x — the permutation terms (plonk/permutation/verifier.rs, the
    same expressions as evaluation.rs:376-452) and the static-lookup terms (evaluation.rs:533-548; B(x) = B_0(x) x + B(0) with B(0) from
    the sumcheck identity n B(0) = N A(0), static_lookup/prover.rs:315-325). Plain integer arithmetic: used by bench.py and the tests to
    check a finished proof without a pairing."""
    n, bf, omega = pk.n, pk.blinding_factors, pk.domain._omega
    x, y, beta, gamma = info["x"], info["y"], info["beta"], info["gamma"]
    e = info["evals"]
    xn = pow(x, n, R_MOD)
    inv = lambda a: pow(a % R_MOD, -1, R_MOD)  # noqa: E731

    def lag_at(i):
        wi = pow(omega, i, R_MOD)
        return (xn - 1) * inv(n) % R_MOD * wi % R_MOD * inv(x - wi) % R_MOD

    l0, l_last = lag_at(0), lag_at(n - bf - 1)
    l_blind = sum(lag_at(i) for i in range(n - bf, n)) % R_MOD
    l_act = (1 - (l_last + l_blind)) % R_MOD
    adv_at = {}
    for (c, rot), v in zip(pk.advice_queries, e["advice"]):
        adv_at[(c, rot)] = v
    exp = 0
    zs = e["z"]
    if zs:
        exp = (exp * y + l0 * (1 - zs[0][0])) % R_MOD
        exp = (exp * y + l_last * (zs[-1][0] * zs[-1][0] - zs[-1][0])) % R_MOD
        for i in range(1, len(zs)):
            exp = (exp * y + l0 * (zs[i][0] - zs[i - 1][2])) % R_MOD
        chunk_len = pk.cs_degree - 2
        cur = beta * x % R_MOD
        for s_, (z_x, z_wx, _) in enumerate(zs):
            cols = pk.permutation_columns[s_ * chunk_len:(s_ + 1) * chunk_len]
            sig = e["sigma"][s_ * chunk_len:(s_ + 1) * chunk_len]
            left, right = z_wx, z_x
            for c, sg in zip(cols, sig):
                a = adv_at[(c, 0)]
                left = left * (a + beta * sg + gamma) % R_MOD
                right = right * (a + cur + gamma) % R_MOD
                cur = cur * permutation.FR_DELTA % R_MOD
            exp = (exp * y + (left - right) * l_act) % R_MOD
    for (b0_x, f_x, a_at_zero), N in zip(e["static_lookups"], info["table_sizes"]):
        b_at_zero = (a_at_zero * N + (bf + 1) * inv(beta)) % R_MOD * inv(n) % R_MOD
        b_x = (b0_x * x + b_at_zero) % R_MOD
        exp = (exp * y + (b_x * (f_x * l_act + beta) - 1)) % R_MOD
    return exp * inv(xn - 1) % R_MOD
