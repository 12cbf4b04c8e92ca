"""ctypes binding of the C-ABI in include/clipper_b200.h (libclipper_b200.so).

The library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a) and lives in
``clipper_b200/lib/``.  There is no fallback: if the shared object is missing, or no Blackwell
GPU is usable, loading / handle creation raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libclipper_b200.so")

OK = 0
STORE_F32, STORE_F64 = 0, 1
ROUND_NONZERO, ROUND_DSD, ROUND_DSD_HEU = 0, 1, 2

# every symbol include/clipper_b200.h declares (checked by tests/test_capi_symbols.py)
SYMBOLS = [
    "clp_create", "clp_destroy", "clp_last_error", "clp_default_params", "clp_set_params",
    "clp_get_params", "clp_set_stream", "clp_version",
    "clp_score_euclidean", "clp_score_pointnormal", "clp_score_euclidean_dev", "clp_score_pointnormal_dev",
    "clp_set_dense", "clp_set_sparse_upper", "clp_get_dense", "clp_num_associations",
    "clp_get_associations", "clp_count_nonzeros",
    "clp_solve", "clp_solve_dev", "clp_matvec", "clp_matvec_dev",
    "clp_k2ij", "clp_create_all_to_all", "clp_find_k_largest", "clp_find_above", "clp_dsd_dense",
    "clp_shard_config", "clp_shard_rows", "clp_shard_export", "clp_shard_import", "clp_shard_blob_bytes",
    "clp_set_ctas_per_sm", "clp_set_grid_cap", "clp_set_dense_mode", "clp_get_dense_mode", "clp_sparse_info",
    "clp_batch_create", "clp_batch_destroy", "clp_batch_last_error", "clp_batch_set_params",
    "clp_batch_solve_euclidean", "clp_batch_solve_pointnormal", "clp_batch_info",
]


class ClpParams(C.Structure):
    """POD mirror of clipper::Params (reference include/clipper/clipper.h:27-60)."""
    _fields_ = [
        ("tol_u", C.c_double), ("tol_F", C.c_double), ("tol_Fop", C.c_double),
        ("maxiniters", C.c_int32), ("maxoliters", C.c_int32),
        ("beta", C.c_double), ("maxlsiters", C.c_int32),
        ("eps", C.c_double), ("affinityeps", C.c_double),
        ("rescale_u0", C.c_int32), ("rounding", C.c_int32),
    ]


class ClpSolution(C.Structure):
    """POD mirror of clipper::Solution (reference include/clipper/clipper.h:65-73) + counters."""
    _fields_ = [
        ("t", C.c_double), ("ifinal", C.c_int32), ("n_nodes", C.c_int32),
        ("score", C.c_double), ("d_final", C.c_double),
        ("n_evals", C.c_int64), ("n_matvec", C.c_int64), ("n_inner", C.c_int64),
        ("kernel_ms", C.c_double),
        ("prof_matvec_ms", C.c_double), ("prof_combine_ms", C.c_double), ("prof_exchange_ms", C.c_double),
    ]


class ClipperError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("clipper_b200 error %d: %s" % (code, msg))
        self.code = code


_lib = None


def load():
    """Load libclipper_b200.so and declare prototypes.  Raises if the library is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "clipper_b200: %s is missing -- run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  There is no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    dp, ip, lp = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_int64)
    i32, i64, dbl = C.c_int32, C.c_int64, C.c_double

    L.clp_create.argtypes = [C.c_int, C.c_int, C.POINTER(vp)]
    L.clp_destroy.argtypes = [vp]
    L.clp_last_error.argtypes = [vp]; L.clp_last_error.restype = C.c_char_p
    L.clp_version.restype = C.c_char_p
    L.clp_default_params.argtypes = [C.POINTER(ClpParams)]; L.clp_default_params.restype = None
    L.clp_set_params.argtypes = [vp, C.POINTER(ClpParams)]
    L.clp_get_params.argtypes = [vp, C.POINTER(ClpParams)]
    L.clp_set_stream.argtypes = [vp, vp]
    # host-pointer scoring takes typed pointers; *_dev variants take raw addresses (void*)
    L.clp_score_euclidean.argtypes = [vp, dp, i32, i64, dp, i64, ip, i64, dbl, dbl, dbl]
    L.clp_score_pointnormal.argtypes = [vp, dp, i64, dp, i64, ip, i64, dbl, dbl, dbl, dbl]
    L.clp_score_euclidean_dev.argtypes = [vp, vp, i32, i64, vp, i64, vp, i64, dbl, dbl, dbl]
    L.clp_score_pointnormal_dev.argtypes = [vp, vp, i64, vp, i64, vp, i64, dbl, dbl, dbl, dbl]
    L.clp_set_dense.argtypes = [vp, dp, dp, i64]
    L.clp_set_sparse_upper.argtypes = [vp, i64, lp, ip, dp, lp, ip, dp]
    L.clp_get_dense.argtypes = [vp, C.c_int, dp]
    L.clp_num_associations.argtypes = [vp, lp]
    L.clp_get_associations.argtypes = [vp, ip]
    L.clp_count_nonzeros.argtypes = [vp, lp, lp]
    L.clp_solve.argtypes = [vp, dp, C.POINTER(ClpSolution), dp, ip, dp]
    L.clp_solve_dev.argtypes = [vp, vp, C.POINTER(ClpSolution), vp, ip]
    L.clp_matvec.argtypes = [vp, dp, dbl, dp, dp, dp]
    L.clp_matvec_dev.argtypes = [vp, vp, dbl, vp, vp, vp, C.c_int, dp]
    L.clp_k2ij.argtypes = [C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.clp_k2ij.restype = None
    L.clp_create_all_to_all.argtypes = [i64, i64, ip]; L.clp_create_all_to_all.restype = None
    L.clp_find_k_largest.argtypes = [dp, i64, i32, ip]; L.clp_find_k_largest.restype = i32
    L.clp_find_above.argtypes = [dp, i64, dbl, ip]; L.clp_find_above.restype = i32
    L.clp_dsd_dense.argtypes = [dp, i64, ip, i32, ip]; L.clp_dsd_dense.restype = i32
    L.clp_shard_config.argtypes = [vp, C.c_int, C.c_int]
    L.clp_shard_blob_bytes.restype = i64
    L.clp_shard_export.argtypes = [vp, vp, i64, lp]
    L.clp_shard_import.argtypes = [vp, vp, i64, C.c_int]
    L.clp_shard_rows.argtypes = [i64, C.c_int, C.c_int, lp, lp]; L.clp_shard_rows.restype = None
    L.clp_set_ctas_per_sm.argtypes = [vp, C.c_int]
    L.clp_set_grid_cap.argtypes = [vp, C.c_int]
    L.clp_set_dense_mode.argtypes = [vp, C.c_int]
    L.clp_get_dense_mode.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.clp_sparse_info.argtypes = [vp, lp, lp]
    # batches of small problems: arrays of host pointers (void**) and sizes
    pp = C.POINTER(C.c_void_p)
    L.clp_batch_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.clp_batch_destroy.argtypes = [vp]
    L.clp_batch_last_error.argtypes = [vp]; L.clp_batch_last_error.restype = C.c_char_p
    L.clp_batch_set_params.argtypes = [vp, C.POINTER(ClpParams)]
    L.clp_batch_solve_euclidean.argtypes = [vp, i32, i32, pp, lp, pp, lp, pp, lp, pp, dbl, dbl, dbl,
                                            C.POINTER(ClpSolution), pp, pp]
    L.clp_batch_solve_pointnormal.argtypes = [vp, i32, pp, lp, pp, lp, pp, lp, pp, dbl, dbl, dbl, dbl,
                                              C.POINTER(ClpSolution), pp, pp]
    L.clp_batch_info.argtypes = [vp, ip, lp, lp]
    _lib = L
    return L


def check(h, rc):
    if rc != OK:
        msg = load().clp_last_error(h)
        raise ClipperError(rc, msg.decode() if msg else "")
