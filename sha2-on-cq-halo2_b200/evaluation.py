"""Mirror of halo2_proofs::plonk::evaluation for the quotient evaluation on the device (reference
halo2_proofs/src/plonk/evaluation.rs): ValueSource / Calculation / GraphEvaluator with the reference's add_expression
optimiser (:571-704), serialised for cqb_graph_evaluate_dev; plus the CQ static-lookup and permutation terms of
evaluate_h (:376-452, :533-548). Expressions use a minimal AST (the reference's Expression enum, plonk/circuit.rs)."""
import ctypes

import numpy as np

from . import _lib
from .fields import R_MOD, fr_to_limbs

# ValueSource kinds in the reference's enum order (evaluation.rs:41-65) — the derived PartialOrd compares this first
CONSTANT, INTERMEDIATE, FIXED, ADVICE, INSTANCE, CHALLENGE, BETA, GAMMA, THETA, Y, PREVIOUS = range(11)
ADD, SUB, MUL, SQUARE, DOUBLE, NEGATE, HORNER, STORE = range(8)


def vs(kind, a=0, b=0):
    return (kind, a, b)


class Expr:
    """Expression<F> (plonk/circuit.rs): ("const", int) | ("fixed"|"advice"|"instance", col, rot) | ("challenge", i) |
    ("neg", e) | ("sum", a, b) | ("prod", a, b) | ("scaled", e, int)"""

    def __init__(self, *node):
        self.node = node

    def __neg__(self):
        return Expr("neg", self)

    def __add__(self, o):
        return Expr("sum", self, o)

    def __sub__(self, o):
        return Expr("sum", self, Expr("neg", o))  # the reference stores a - b as a + (-b)

    def __mul__(self, o):
        return Expr("scaled", self, o % R_MOD) if isinstance(o, int) else Expr("prod", self, o)


class GraphEvaluator:
    """reference evaluation.rs:197-207, 553-775"""

    def __init__(self):
        self.constants = [0, 1, 2]  # :557-562 fixed positions
        self.rotations = []
        self.calculations = []      # (calculation tuple, target)
        self.num_intermediates = 0

    def add_rotation(self, rot):  # :573-583
        if rot in self.rotations:
            return self.rotations.index(rot)
        self.rotations.append(rot)
        return len(self.rotations) - 1

    def add_constant(self, c):  # :586-596
        c %= R_MOD
        if c in self.constants:
            return vs(CONSTANT, self.constants.index(c))
        self.constants.append(c)
        return vs(CONSTANT, len(self.constants) - 1)

    def add_calculation(self, calc):  # :602-621
        for existing, target in self.calculations:
            if existing == calc:
                return vs(INTERMEDIATE, target)
        target = self.num_intermediates
        self.calculations.append((calc, target))
        self.num_intermediates += 1
        return vs(INTERMEDIATE, target)

    def add_expression(self, e):  # :624-704
        n = e.node
        t = n[0]
        if t == "const":
            return self.add_constant(n[1])
        if t in ("fixed", "advice", "instance"):
            kind = {"fixed": FIXED, "advice": ADVICE, "instance": INSTANCE}[t]
            return self.add_calculation((STORE, vs(kind, n[1], self.add_rotation(n[2]))))
        if t == "challenge":
            return self.add_calculation((STORE, vs(CHALLENGE, n[1])))
        if t == "neg":
            if n[1].node[0] == "const":
                return self.add_constant(-n[1].node[1])
            ra = self.add_expression(n[1])
            return ra if ra == vs(CONSTANT, 0) else self.add_calculation((NEGATE, ra))
        if t == "sum":
            a, b = n[1], n[2]
            if b.node[0] == "neg":
                ra, rb = self.add_expression(a), self.add_expression(b.node[1])
                if ra == vs(CONSTANT, 0):
                    return self.add_calculation((NEGATE, rb))
                if rb == vs(CONSTANT, 0):
                    return ra
                return self.add_calculation((SUB, ra, rb))
            ra, rb = self.add_expression(a), self.add_expression(b)
            if ra == vs(CONSTANT, 0):
                return rb
            if rb == vs(CONSTANT, 0):
                return ra
            return self.add_calculation((ADD, ra, rb) if ra <= rb else (ADD, rb, ra))
        if t == "prod":
            ra, rb = self.add_expression(n[1]), self.add_expression(n[2])
            if ra == vs(CONSTANT, 0) or rb == vs(CONSTANT, 0):
                return vs(CONSTANT, 0)
            if ra == vs(CONSTANT, 1):
                return rb
            if rb == vs(CONSTANT, 1):
                return ra
            if ra == vs(CONSTANT, 2):
                return self.add_calculation((DOUBLE, rb))
            if rb == vs(CONSTANT, 2):
                return self.add_calculation((DOUBLE, ra))
            if ra == rb:
                return self.add_calculation((SQUARE, ra))
            return self.add_calculation((MUL, ra, rb) if ra <= rb else (MUL, rb, ra))
        if t == "scaled":
            f = n[2] % R_MOD
            if f == 0:
                return vs(CONSTANT, 0)
            if f == 1:
                return self.add_expression(n[1])
            cst = self.add_constant(f)
            ra = self.add_expression(n[1])
            return self.add_calculation((MUL, ra, cst))
        raise ValueError(t)

    # ---- serialisation for the C ABI (include/cqb200.h) ----
    @staticmethod
    def _vs_words(v):
        kind, a, b = v
        rot = b if kind in (FIXED, ADVICE, INSTANCE) else 0
        return [kind | (rot << 8), a]

    def serialize(self):
        code = []
        for calc, target in self.calculations:
            op = calc[0]
            code += [op, target]
            if op == HORNER:  # (HORNER, start, parts, factor)
                _, start, parts, factor = calc
                code += self._vs_words(start) + self._vs_words(factor) + [len(parts)]
                for p in parts:
                    code += self._vs_words(p)
            else:
                for operand in calc[1:]:
                    code += self._vs_words(operand)
        consts = np.stack([fr_to_limbs(c) for c in self.constants])
        return consts, np.array(self.rotations, dtype=np.int32), np.array(code, dtype=np.uint32)

    def evaluate_dev(self, fixed_ptrs, advice_ptrs, instance_ptrs, challenges, beta, gamma, theta, y, d_values, size, rot_scale):
        """GraphEvaluator::evaluate for all rows (:718-775) on device-resident columns; in place on d_values"""
        consts, rots, code = self.serialize()
        g = CqbGraph(consts.ctypes.data_as(_lib.u64p), consts.shape[0], rots.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
                     rots.shape[0], code.ctypes.data_as(_lib.u32p), code.shape[0], len(self.calculations), self.num_intermediates)

        def ptr_array(ptrs):
            arr = (ctypes.c_void_p * max(len(ptrs), 1))(*[ctypes.c_void_p(p) for p in ptrs])
            return arr, len(ptrs)

        fa, nf = ptr_array(fixed_ptrs)
        aa, na = ptr_array(advice_ptrs)
        ia, ni = ptr_array(instance_ptrs)
        ch = np.ascontiguousarray(challenges, dtype=np.uint64).reshape(-1, 4) if len(challenges) else np.zeros((1, 4), np.uint64)
        _lib.check(_lib.lib().cqb_graph_evaluate_dev(ctypes.byref(g), fa, nf, aa, na, ia, ni, _lib.p64(ch), len(challenges),
                                                     _lib.p64(_lib.fr_limbs(beta)), _lib.p64(_lib.fr_limbs(gamma)),
                                                     _lib.p64(_lib.fr_limbs(theta)), _lib.p64(_lib.fr_limbs(y)),
                                                     ctypes.c_void_p(d_values), size, rot_scale))


class CqbGraph(ctypes.Structure):
    _fields_ = [("constants", _lib.u64p), ("n_constants", ctypes.c_uint32), ("rotations", ctypes.POINTER(ctypes.c_int32)),
                ("n_rotations", ctypes.c_uint32), ("code", _lib.u32p), ("code_words", ctypes.c_uint32),
                ("n_calculations", ctypes.c_uint32), ("num_intermediates", ctypes.c_uint32)]


def custom_gates_evaluator(gate_polys):
    """Evaluator::new, custom-gate part (:228-246): Horner over the gate polynomials with PreviousValue and y"""
    ev = GraphEvaluator()
    parts = [ev.add_expression(p) for p in gate_polys]
    ev.add_calculation((HORNER, vs(PREVIOUS), tuple(parts), vs(Y)))
    return ev


def cq_lookup_h_dev(d_values, d_b_coset, d_f_coset, d_l_active_row, beta, y, size):
    """evaluate_h, static lookups (:533-548)"""
    _lib.check(_lib.lib().cqb_cq_lookup_h_dev(ctypes.c_void_p(d_values), ctypes.c_void_p(d_b_coset), ctypes.c_void_p(d_f_coset),
                                              ctypes.c_void_p(d_l_active_row), _lib.p64(_lib.fr_limbs(beta)), _lib.p64(_lib.fr_limbs(y)), size))


def permutation_h_dev(d_values, size, rot_scale, last_rotation, chunk_len, set_ptrs, column_ptrs, perm_coset_ptrs, d_l0, d_l_last,
                      d_l_active_row, beta, gamma, y, extended_omega):
    """evaluate_h, permutation constraints (:376-452)"""
    def ptr_array(ptrs):
        return (ctypes.c_void_p * max(len(ptrs), 1))(*[ctypes.c_void_p(p) for p in ptrs])

    assert len(column_ptrs) == len(perm_coset_ptrs)
    _lib.check(_lib.lib().cqb_permutation_h_dev(
        ctypes.c_void_p(d_values), size, rot_scale, last_rotation, chunk_len, ptr_array(set_ptrs), len(set_ptrs), ptr_array(column_ptrs),
        ptr_array(perm_coset_ptrs), len(column_ptrs), ctypes.c_void_p(d_l0), ctypes.c_void_p(d_l_last), ctypes.c_void_p(d_l_active_row),
        _lib.p64(_lib.fr_limbs(beta)), _lib.p64(_lib.fr_limbs(gamma)), _lib.p64(_lib.fr_limbs(y)), _lib.p64(_lib.fr_limbs(extended_omega))))
