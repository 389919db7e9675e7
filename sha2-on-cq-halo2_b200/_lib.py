"""ctypes binding of libcqb200.so (include/cqb200.h). The library is the product; there is no CPU fallback: if the
shared object is missing or no CUDA device is visible, every compute call raises."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("CQB200_LIB") or os.path.join(_HERE, "libcqb200.so")  # env override: A/B experiments only

u64p = ctypes.POINTER(ctypes.c_uint64)
u32p = ctypes.POINTER(ctypes.c_uint32)

CQB_OK, CQB_E_NO_DEVICE, CQB_E_CUDA, CQB_E_BAD_ARG, CQB_E_LEN_MISMATCH, CQB_E_BAD_SIZE, CQB_E_OOM = range(7)

# every symbol include/cqb200.h declares: (name, restype, argtypes)
_vp, _sz, _u64, _u32, _int = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int
_ip = ctypes.POINTER(ctypes.c_int)
SYMBOLS = [
    ("cqb_init", _int, [_int]),
    ("cqb_shutdown", None, []),
    ("cqb_last_error", ctypes.c_char_p, []),
    ("cqb_device_count", _int, []),
    ("cqb_set_stream", _int, [_vp]),
    ("cqb_init_multi", _int, [_int]),
    ("cqb_active_devices", _int, []),
    ("cqb_bases_register_sharded", _int, [u64p, _sz, u64p]),
    ("cqb_msm_bn254_g1_multi_dev", _int, [_u64, _sz, ctypes.POINTER(_vp), _sz, u64p, _ip]),
    ("cqb_dev_alloc_on", _int, [_int, _sz, ctypes.POINTER(_vp)]),
    ("cqb_dev_free_on", _int, [_int, _vp]),
    ("cqb_memcpy_h2d_on", _int, [_int, _vp, _vp, _sz]),
    ("cqb_synth_scalars_dev_on", _int, [_int, _u64, _sz, _sz, _vp]),
    ("cqb_sync", _int, []),
    ("cqb_launch_count", ctypes.c_ulonglong, []),
    ("cqb_bases_register", _int, [u64p, _sz, u64p]),
    ("cqb_bases_register_device", _int, [_vp, _sz, u64p]),
    ("cqb_bases_free", _int, [_u64]),
    ("cqb_bases_len", _sz, [_u64]),
    ("cqb_bases_download", _int, [_u64, _sz, _sz, u64p]),
    ("cqb_bases_copy_dev", _int, [_u64, _sz, _sz, _vp]),
    ("cqb_bases_precompute", _int, [_u64, _int]),
    ("cqb_bases_drop_precomputed", _int, [_u64]),
    ("cqb_bases_precomputed_window_bits", _int, [_u64]),
    ("cqb_msm_bn254_g1", _int, [_u64, _sz, u64p, _sz, u64p, _ip]),
    ("cqb_msm_bn254_g1_dev", _int, [_u64, _sz, _vp, _sz, u64p, _ip]),
    ("cqb_msm_bn254_g1_batch", _int, [_u64, _sz, u64p, _sz, _int, u64p, _ip]),
    ("cqb_msm_bn254_g1_batch_dev", _int, [_u64, _sz, _vp, _sz, _int, u64p, _ip]),
    ("cqb_msm_bn254_g1_host", _int, [u64p, u64p, _sz, u64p, _ip]),
    ("cqb_g1_batch_normalize", _int, [u64p, _sz, u64p]),
    ("cqb_msm_bn254_g1_jacobian", _int, [u64p, u64p, _sz, u64p, _ip]),
    ("cqb_set_host_bases_cache", _int, [ctypes.c_longlong]),
    ("cqb_msm_bn254_g1_sparse", _int, [_u64, u32p, u64p, _sz, u64p, _ip]),
    ("cqb_g1_sum_affine", _int, [u64p, _sz, u64p, _ip]),
    ("cqb_g1_sum_affine_dev", _int, [_vp, _sz, u64p, _ip]),
    ("cqb_msm_bn254_g1_dev_to", _int, [_u64, _sz, _vp, _sz, _vp]),
    ("cqb_msm_bn254_g1_to", _int, [_u64, _sz, u64p, _sz, _vp]),
    ("cqb_ntt_bn254_fr", _int, [u64p, u64p, _u32]),
    ("cqb_ntt_bn254_fr_dev", _int, [_vp, u64p, _u32]),
    ("cqb_ntt_bn254_fr_batch_dev", _int, [_vp, u64p, _u32, _u32]),
    ("cqb_ntt_bn254_fr_batch_map_dev", _int, [_vp, _vp, u64p, _u32, _u32, _int, _int, u64p, _u32, _sz, _u32]),
    ("cqb_ntt_bn254_fr_batch_p2p_dev", _int, [_vp, _vp, _vp, _u32, _u32, u64p, _u32, _u32, _int, u64p, _u32, _sz]),
    ("cqb_ipc_export", _int, [_vp, ctypes.c_char_p]),
    ("cqb_ipc_open", _int, [ctypes.c_char_p, ctypes.POINTER(_vp)]),
    ("cqb_ipc_close", _int, [_vp]),
    ("cqb_fr_mul_omega_powers_dev", _int, [_vp, _sz, _sz, _sz, u64p, _u32]),
    ("cqb_fr_transpose_dev", _int, [_vp, _vp, _sz, _sz]),
    ("cqb_intt_bn254_fr", _int, [u64p, u64p, u64p, _u32]),
    ("cqb_intt_bn254_fr_dev", _int, [_vp, u64p, u64p, _u32]),
    ("cqb_coset_ntt_bn254_fr", _int, [u64p, _sz, u64p, u64p, _u32, u64p, u64p]),
    ("cqb_coset_ntt_bn254_fr_dev", _int, [_vp, _sz, _vp, u64p, _u32, u64p, u64p]),
    ("cqb_coset_intt_bn254_fr", _int, [u64p, _u32, u64p, u64p, u64p, u64p, u64p, _u32]),
    ("cqb_coset_intt_bn254_fr_dev", _int, [_vp, _u32, u64p, u64p, u64p, u64p, u64p, _u32]),
    ("cqb_synth_scalars_dev", _int, [_u64, _sz, _sz, _vp]),
    ("cqb_synth_bases_dev", _int, [_u64, _sz, _sz, _vp]),
    ("cqb_srs_setup_dev", _int, [_u32, u64p, _vp, _vp]),
    ("cqb_table_srs_setup_dev", _int, [_u32, u64p, _vp, _vp, _vp]),
    ("cqb_g1_generator_mul_dev", _int, [_vp, _sz, _vp]),
    ("cqb_g_to_lagrange_dev", _int, [_vp, _u32, _vp]),
    ("cqb_cq_table_qs_dev", _int, [_vp, _u32, _vp, _vp]),
    ("cqb_fr_scale_dev", _int, [_vp, _sz, u64p]),
    ("cqb_fr_batch_invert_dev", _int, [_vp, _sz]),
    ("cqb_graph_evaluate_dev", _int, [_vp, _vp, _u32, _vp, _u32, _vp, _u32, u64p, _u32, u64p, u64p, u64p, u64p, _vp, _u64, ctypes.c_int32]),
    ("cqb_cq_lookup_h_dev", _int, [_vp, _vp, _vp, _vp, u64p, u64p, _u64]),
    ("cqb_permutation_h_dev", _int, [_vp, _u64, ctypes.c_int32, ctypes.c_int32, _u32, _vp, _u32, _vp, _vp, _u32, _vp, _vp, _vp, u64p, u64p,
                                     u64p, u64p]),
    ("cqb_eval_polynomial_dev", _int, [_vp, _sz, u64p, u64p]),
    ("cqb_eval_polynomials_dev", _int, [_vp, _sz, u64p, _u32, u64p]),
    ("cqb_kate_division_dev", _int, [_vp, _sz, u64p, _vp]),
    ("cqb_fr_powers_dev", _int, [u64p, _sz, _vp]),
    ("cqb_fr_prefix_product_dev", _int, [_vp, _sz, u64p, _vp]),
    ("cqb_lookup_product_dev", _int, [_vp, _vp, _vp, _vp, _u32, u64p, u64p, _vp]),
    ("cqb_lookup_h_dev", _int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, u64p, u64p, u64p, _u64, ctypes.c_int32]),
    ("cqb_fr_compress_dev", _int, [_vp, _u32, _vp, _sz, u64p, _vp]),
    ("cqb_fr_inv_shifted_dev", _int, [_vp, _sz, _sz, u64p, _vp]),
    ("cqb_fr_mul_dev", _int, [_vp, _vp, _sz, _vp]),
    ("cqb_fr_axpy_dev", _int, [_vp, u64p, _vp, _sz]),
    ("cqb_msm_bn254_g1_sparse_dev", _int, [_u64, _vp, _vp, _sz, u64p, _ip]),
    ("cqb_permutation_product_dev", _int, [_vp, _vp, _u32, _u32, u64p, u64p, u64p, u64p, u64p, u64p, _vp]),
    ("cqb_g2_powers", _int, [u64p, _sz, u64p]),
    ("cqb_msm_bn254_g2", _int, [u64p, u64p, _sz, u64p, _ip]),
    ("cqb_g2_generator_mul_dev", _int, [_vp, _sz, _vp]),
    ("cqb_dev_alloc", _int, [_sz, ctypes.POINTER(_vp)]),
    ("cqb_dev_free", _int, [_vp]),
    ("cqb_memcpy_h2d", _int, [_vp, _vp, _sz]),
    ("cqb_memcpy_d2h", _int, [_vp, _vp, _sz]),
    ("cqb_memcpy_d2d", _int, [_vp, _vp, _sz]),
    ("cqb_host_alloc_pinned", _int, [_sz, ctypes.POINTER(_vp)]),
    ("cqb_host_free_pinned", _int, [_vp]),
    ("cqb_msm_set_window_bits", _int, [_int]),
    ("cqb_msm_set_parts", _int, [_int]),
    ("cqb_msm_set_accumulator", _int, [_int, _int]),
    ("cqb_msm_set_tree_levels", _int, [_int]),
    ("cqb_msm_set_sort_mode", _int, [_int]),
    ("cqb_msm_last_tree_levels", _int, []),
    ("cqb_msm_set_profiling", _int, [_int]),
    ("cqb_msm_phase_ms", _int, [ctypes.POINTER(ctypes.c_float), _int]),
]


class CqbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libcqb200 error {code}: {msg}")
        self.code = code


_lib = None
_inited_device = None


def load():
    """dlopen the in-tree library and bind every declared symbol; raises if the extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise ImportError(
                f"{SO_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback for this path)")
        lib = ctypes.CDLL(SO_PATH)
        for name, res, args in SYMBOLS:
            fn = getattr(lib, name)  # AttributeError here means header and library disagree
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        raise CqbError(rc, load().cqb_last_error().decode("utf-8", "replace"))


def init(device=None):
    """Bind this process to one GPU (one process per GPU). Raises CqbError(CQB_E_NO_DEVICE) without a GPU."""
    global _inited_device
    lib = load()
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    if _inited_device != device:
        check(lib.cqb_init(device))
        _inited_device = device
    return lib


def lib():
    if _inited_device is None:
        return init()
    return _lib


def p64(a):
    return a.ctypes.data_as(u64p)


def fr_limbs(x):
    a = np.ascontiguousarray(x, dtype=np.uint64).reshape(-1)
    assert a.shape[0] == 4
    return a
