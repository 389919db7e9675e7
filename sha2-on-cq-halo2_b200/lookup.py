"""Mirror of the plookup argument's data-parallel parts (reference halo2_proofs/src/plonk/lookup/prover.rs:173-262
commit_product; plonk/evaluation.rs:458-531 its evaluate_h constraints) on device-resident vectors. The CQ circuits of the
reference use static lookups instead (cq.py); create_proof runs both arguments, so both are here. The permutation of the
input / table expressions (permute_expression_pair, :395-470: sorting + row matching) is data-dependent CPU work and stays
with the caller."""
import ctypes

from . import _lib
from .fields import fr_to_limbs


def commit_product_dev(d_compressed_input, d_compressed_table, d_permuted_input, d_permuted_table, k, beta, gamma, d_z):
    """prover.rs:206-254: z into d_z (2^k Fr); the caller then overwrites the last blinding_factors rows (:259) and commits
    with params.commit_lagrange (:298). beta, gamma: canonical ints."""
    vp = ctypes.c_void_p
    _lib.check(_lib.lib().cqb_lookup_product_dev(vp(d_compressed_input), vp(d_compressed_table), vp(d_permuted_input), vp(d_permuted_table), k,
                                                 _lib.p64(fr_to_limbs(beta)), _lib.p64(fr_to_limbs(gamma)), vp(d_z)))


def lookup_h_dev(d_values, d_table_value, d_product_coset, d_permuted_input_coset, d_permuted_table_coset, d_l0, d_l_last, d_l_active_row, beta,
                 gamma, y, size, rot_scale):
    """evaluation.rs:458-531 for one lookup; beta / gamma / y: (4,) uint64 Montgomery limbs as in evaluation.py"""
    vp = ctypes.c_void_p
    _lib.check(_lib.lib().cqb_lookup_h_dev(vp(d_values), vp(d_table_value), vp(d_product_coset), vp(d_permuted_input_coset),
                                           vp(d_permuted_table_coset), vp(d_l0), vp(d_l_last), vp(d_l_active_row), _lib.p64(_lib.fr_limbs(beta)),
                                           _lib.p64(_lib.fr_limbs(gamma)), _lib.p64(_lib.fr_limbs(y)), size, rot_scale))
