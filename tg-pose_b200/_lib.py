"""ctypes binding of libtgpose_b200.so (include/tgpose_b200.h).

The library is built in-tree by tg-pose_b200/build.py.  There is no fallback: if the shared
object is missing or a call fails, this raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libtgpose_b200.so")

c_void_p, c_int, c_long, c_size_t = ctypes.c_void_p, ctypes.c_int, ctypes.c_long, ctypes.c_size_t


class OutSeg(ctypes.Structure):
    """tgp_out_seg"""
    _fields_ = [("col_begin", c_int), ("col_end", c_int), ("mode", c_int), ("slab_width", c_int),
                ("ld", c_long), ("ptr", c_void_p)]


class ConcatSrc(ctypes.Structure):
    """tgp_concat_src"""
    _fields_ = [("ptr", c_void_p), ("C", c_int), ("ld", c_long), ("idx", c_void_p), ("n_src", c_int)]


class GemmArgs(ctypes.Structure):
    """tgp_gemm_args"""
    _fields_ = [("A", c_void_p), ("lda", c_long),
                ("Bmat", c_void_p), ("ldb", c_long), ("b_is_nk", c_int),
                ("M", c_long), ("K", c_int), ("Ncols", c_int),
                ("bias", c_void_p), ("group_bias", c_void_p), ("rows_per_group", c_int),
                ("res1", c_void_p), ("ld_res1", c_long),
                ("res2", c_void_p), ("ld_res2", c_long),
                ("scale", c_void_p), ("shift", c_void_p),
                ("relu", c_int), ("neg_slope", c_void_p), ("nseg", c_int), ("seg", OutSeg * 4),
                ("A_split", c_void_p), ("B_split", c_void_p), ("mixed", c_int), ("a_kp", c_int), ("a_group_cols", c_int),
                ("res1_idx", c_void_p), ("res2_idx", c_void_p)]


class RangerHyper(ctypes.Structure):
    """tgp_ranger_hyper"""
    _fields_ = [("beta1", ctypes.c_float), ("beta2", ctypes.c_float), ("eps", ctypes.c_float),
                ("weight_decay", ctypes.c_float), ("one_minus_beta1", ctypes.c_float),
                ("one_minus_beta2", ctypes.c_float), ("neg_step", ctypes.c_float), ("rectified", c_int),
                ("lookahead", c_int), ("la_alpha", ctypes.c_float), ("max_norm", ctypes.c_float), ("gc_on_update", c_int)]


# name -> (restype, argtypes); must list every symbol include/tgpose_b200.h declares
SIGNATURES = {
    "tgp_version": (c_int, []),
    "tgp_last_error": (ctypes.c_char_p, []),
    "tgp_launch_count": (ctypes.c_ulonglong, []),
    "tgp_knn_xyz": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "tgp_knn_feat_workspace": (c_size_t, [c_int, c_int, c_int, c_int]),
    "tgp_knn_feat": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_size_t,
                             c_void_p]),
    "tgp_nearest": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "tgp_gather_rows": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "tgp_select_rows": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "tgp_direction_norm": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "tgp_gather_max": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p,
                               c_void_p, c_void_p]),
    "tgp_orl_workspace": (c_size_t, [c_int, c_int, c_int]),
    "tgp_orl_global": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                               c_size_t, c_void_p]),
    "tgp_surface_conv_fwd": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                     c_void_p, c_void_p, c_void_p, c_void_p]),
    "tgp_edge_records": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "tgp_layer_conv_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_long, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                   c_void_p, c_void_p, c_void_p, c_void_p]),
    "tgp_gemm": (c_int, [ctypes.POINTER(GemmArgs), c_void_p]),
    "tgp_split_kpad": (c_int, [c_int]),
    "tgp_split_tf32": (c_int, [c_void_p, c_long, c_int, c_long, c_int, c_void_p, c_void_p]),
    "tgp_chamfer_fwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_void_p]),
    "tgp_chamfer_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                c_void_p, c_void_p, c_void_p]),
    "tgp_concat_rows": (c_int, [ctypes.POINTER(ConcatSrc), c_int, c_int, c_int, c_void_p, c_long, c_void_p, c_int,
                                c_int, c_void_p]),
    "tgp_decode_max": (c_int, [c_void_p, c_long, c_void_p, c_void_p]),
    "tgp_mixed_kpad": (c_int, [c_int]),
    "tgp_split_mixed": (c_int, [c_void_p, c_long, c_int, c_long, c_void_p, c_void_p]),
    "tgp_split_mixed_w16": (c_int, [c_void_p, c_long, c_int, c_long, c_void_p, c_void_p]),
    "tgp_split_mixed_t": (c_int, [c_void_p, c_long, c_int, c_long, c_void_p, c_void_p]),
    "tgp_split_mixed_t_bytes": (c_size_t, [c_long, c_int]),
    "tgp_split_tf32_t": (c_int, [c_void_p, c_long, c_int, c_long, c_void_p, c_void_p]),
    "tgp_split_tf32_t_bytes": (c_size_t, [c_long, c_int]),
    "tgp_dcd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, ctypes.c_float, ctypes.c_float,
                        c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "tgp_act_bwd": (c_int, [c_void_p, c_long, c_void_p, c_long, c_void_p, c_int, c_long, c_int, c_void_p, c_long,
                            c_void_p]),
    "tgp_colsum_workspace": (c_size_t, [c_long, c_int, c_long]),
    "tgp_colsum": (c_int, [c_void_p, c_long, c_long, c_int, c_long, c_void_p, c_void_p, c_size_t, c_void_p]),
    "tgp_gather_max_bwd": (c_int, [c_void_p, c_int, ctypes.c_float, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int,
                                   c_int, c_int, c_int, c_void_p, c_void_p]),
    "tgp_scatter_add_rows": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "tgp_layer_conv_bwd_workspace": (c_size_t, [c_int, c_int, c_int]),
    "tgp_layer_conv_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_long, c_int, c_int, c_int, c_int,
                                   c_int, c_void_p, c_long, c_void_p, c_void_p, c_size_t, c_void_p]),
    "tgp_surface_conv_bwd_workspace": (c_size_t, [c_int, c_int, c_int, c_int]),
    "tgp_surface_conv_bwd": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_long, c_int, c_int,
                                     c_int, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "tgp_bn_workspace": (c_size_t, [c_long, c_int]),
    "tgp_colsumsq_dev": (c_int, [c_void_p, c_long, c_long, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "tgp_affine_act": (c_int, [c_void_p, c_long, c_void_p, c_void_p, ctypes.c_float, c_long, c_int, c_void_p, c_long,
                               c_void_p, c_int, c_int, c_void_p]),
    "tgp_bn_bwd": (c_int, [c_void_p, c_long, c_void_p, c_long, c_void_p, c_long, c_void_p, c_void_p, c_void_p,
                           c_void_p, c_void_p, ctypes.c_float, c_long, c_int, c_void_p, c_long, c_void_p, c_void_p, c_void_p, c_void_p,
                           c_size_t, c_void_p]),
    "tgp_gemm_tn_workspace": (c_size_t, [c_long, c_int, c_int]),
    "tgp_gemm_tn": (c_int, [c_void_p, c_long, c_void_p, c_long, c_long, c_int, c_int, c_void_p, c_long, c_void_p,
                            c_size_t, c_void_p]),
    "tgp_gemm_tn_tc_workspace": (c_size_t, [c_long, c_int, c_int]),
    "tgp_ranger_reduce": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "tgp_ranger_update": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p,
                                  c_void_p, c_void_p, ctypes.POINTER(RangerHyper), c_void_p, c_void_p]),
    "tgp_gemm_tn_tc": (c_int, [c_void_p, c_void_p, c_long, c_int, c_int, c_void_p, c_long, c_int, c_void_p, c_size_t,
                               c_void_p]),
    "tgp_gemm_tn_tc_rm": (c_int, [c_void_p, c_int, c_void_p, c_int, c_long, c_int, c_int, c_void_p, c_long, c_void_p, c_size_t,
                          c_void_p]),
}

_lib = None


def load():
    """Load the shared object (raises if it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise RuntimeError(
                f"{SO_PATH} is missing: build it with `python tg-pose_b200/build.py` "
                "(or __graft_entry__.build()); there is no CPU/PyTorch fallback for this path")
        lib = ctypes.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc, name):
    if rc != 0:
        msg = load().tgp_last_error().decode(errors="replace")
        raise RuntimeError(f"{name} failed (rc={rc}): {msg}")


def launch_count():
    return int(load().tgp_launch_count())
