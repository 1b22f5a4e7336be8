"""ctypes binding of libp3d_b200.so -- the C ABI declared in include/p3d_b200.h.

This is the only place the shared library is loaded.  There is no fallback of any kind:
if the library is missing (not built) the import of an operator raises, and every entry
point fails when no CUDA device is present.  The library is built in-tree by
`python -m torch_points3d_b200.build` (or `__graft_entry__.build()`).
"""
import ctypes
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libp3d_b200.so")

c_int, c_i64, c_f32, c_vp, c_sz = ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_void_p, ctypes.c_size_t

# name -> (restype, argtypes); mirrors include/p3d_b200.h one to one (tests check both directions)
SIGNATURES = {
    "p3d_version": (c_int, []),
    "p3d_last_error": (ctypes.c_char_p, []),
    "p3d_launch_count": (ctypes.c_uint64, []),
    "p3d_arch": (ctypes.c_char_p, []),
    "p3d_fps_dense_ws_bytes": (c_sz, [c_int, c_int]),
    "p3d_fps_dense_f32": (c_int, [c_vp, c_int, c_int, c_int, c_vp, c_int, c_vp, c_vp, c_sz, c_vp]),
    "p3d_fps_csr_ws_bytes": (c_sz, [c_i64, c_i64]),
    "p3d_fps_csr_f32": (c_int, [c_vp, c_vp, c_int, c_vp, c_vp, c_vp, c_i64, c_vp, c_sz, c_vp]),
    "p3d_ball_query_dense_f32": (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_f32, c_int, c_int, c_vp, c_vp, c_vp]),
    "p3d_ball_query_dense_multi_f32": (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp, c_int, c_vp, c_vp]),
    "p3d_ball_query_batch_f32": (c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_f32, c_int, c_int, c_i64, c_vp,
                                         c_vp, c_vp]),
    "p3d_knn_batch_f32": (c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_int, c_vp, c_vp, c_vp]),
    "p3d_knn_dense_f32": (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp]),
    "p3d_cellgrid_ws_bytes": (c_sz, [c_i64]),
    "p3d_ball_query_batch_grid_f32": (c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_f32, c_int, c_int, c_i64, c_vp,
                                              c_vp, c_vp, c_sz, c_vp]),
    "p3d_knn_batch_grid_f32": (c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_int, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "p3d_three_nn_f32": (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp]),
    "p3d_group_fwd_f32": (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_vp, c_i64, c_int, c_vp]),
    "p3d_group_bwd_ws_bytes": (c_sz, [c_int, c_int, c_int, c_int, c_int]),
    "p3d_group_bwd_prepare": (c_int, [c_vp, c_int, c_int, c_int, c_int, c_int, c_vp, c_sz, c_vp]),
    "p3d_group_bwd_f32": (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_vp, c_i64, c_int, c_vp, c_sz,
                                  c_int, c_vp]),
    "p3d_group_xyz_f32": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_f32, c_vp, c_i64, c_vp]),
    "p3d_group_xyz_feat_f32": (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_f32, c_vp, c_vp]),
    "p3d_three_interp_fwd_f32": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp]),
    "p3d_three_interp_bwd_ws_bytes": (c_sz, [c_int, c_int, c_int, c_int]),
    "p3d_three_interp_bwd_prepare": (c_int, [c_vp, c_int, c_int, c_int, c_int, c_vp, c_sz, c_vp]),
    "p3d_three_interp_bwd_f32": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp, c_sz, c_int,
                                         c_vp]),
    "p3d_rowgemm_ws_bytes": (c_sz, [c_int, c_int]),
    "p3d_rowgemm_f32": (c_int, [c_vp, c_i64, c_vp, c_i64, c_i64, c_i64, c_int, c_int, c_vp, c_i64, c_vp, c_sz, c_vp]),
    "p3d_rowgemm_dw_ws_bytes": (c_sz, [c_i64, c_int, c_int]),
    "p3d_rowgemm_dw_f32": (c_int, [c_vp, c_vp, c_i64, c_int, c_int, c_vp, c_vp, c_sz, c_vp]),
    "p3d_kpconv_fwd_ws_bytes": (c_sz, [c_i64, c_int, c_int, c_int]),
    "p3d_kpconv_fwd_f32": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_int, c_int, c_int, c_int,
                                   c_f32, c_int, c_int, c_int, c_vp, c_vp, c_sz, c_vp]),
    "p3d_kpconv_bwd_ws_bytes": (c_sz, [c_i64, c_int, c_int, c_int, c_int]),
    "p3d_kpconv_bwd_f32": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_int, c_int, c_int,
                                   c_int, c_f32, c_int, c_int, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "p3d_kpconv_fwd_wf_f32": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_int, c_int, c_int, c_int,
                                      c_f32, c_int, c_int, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "p3d_kpconv_bwd_wf_f32": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_int, c_int, c_int,
                                      c_int, c_f32, c_int, c_int, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "p3d_kpconv_wf_f32": (c_int, [c_vp] * 5 + [c_i64, c_i64, c_int, c_int, c_int, c_f32, c_int, c_int, c_vp, c_vp]),
    "p3d_kpconv_scatter_f32": (c_int, [c_vp] * 5 + [c_i64, c_i64, c_int, c_int, c_int, c_f32, c_int, c_int, c_vp, c_vp]),
    "p3d_kpconv_deform_fwd_f32": (c_int, [c_vp] * 8 + [c_i64, c_i64, c_int, c_int, c_int, c_int, c_f32, c_int] +
                                  [c_vp] * 5 + [c_vp]),
    "p3d_kpconv_deform_bwd_ws_bytes": (c_sz, [c_i64, c_int, c_int]),
    "p3d_kpconv_deform_bwd_f32": (c_int, [c_vp] * 13 + [c_i64, c_i64, c_int, c_int, c_int, c_int, c_f32, c_int] +
                                  [c_vp] * 4 + [c_vp, c_sz, c_vp]),
    "p3d_neighbor_max_f32": (c_int, [c_vp, c_vp, c_i64, c_i64, c_int, c_int, c_vp, c_vp, c_vp]),
    "p3d_neighbor_max_bwd_f32": (c_int, [c_vp, c_vp, c_i64, c_i64, c_int, c_vp, c_vp]),
    "p3d_edge_message_f32": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_int, c_vp, c_vp]),
    "p3d_edge_message_bwd_f32": (c_int, [c_vp, c_vp, c_i64, c_int, c_vp, c_vp]),
    "p3d_edge_encoding_f32": (c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_vp]),
    "p3d_segment_reduce_f32": (c_int, [c_vp, c_vp, c_i64, c_i64, c_int, c_int, c_vp, c_vp, c_vp]),
    "p3d_segment_reduce_bwd_f32": (c_int, [c_vp, c_vp, c_vp, c_i64, c_i64, c_int, c_int, c_vp, c_vp]),
    "p3d_grid_cluster_ws_bytes": (c_sz, [c_i64]),
    "p3d_grid_cluster_f32": (c_int, [c_vp, c_vp, c_i64, c_f32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp,
                                     c_sz, c_vp]),
    "p3d_segment_mean_f32": (c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, c_int, c_vp, c_vp]),
    "p3d_segment_vote_i64": (c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, c_int, c_vp, c_vp, c_vp]),
}

_lib = None
_lock = threading.Lock()


class P3DError(RuntimeError):
    """an entry point of libp3d_b200.so returned a negative status"""


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise ImportError(
                        "libp3d_b200.so is not built (%s): run `python -m torch_points3d_b200.build`; "
                        "there is no CPU or PyTorch fallback for these operators" % LIB_PATH)
                handle = ctypes.CDLL(LIB_PATH)
                for name, (res, args) in SIGNATURES.items():
                    fn = getattr(handle, name)  # AttributeError = header/library mismatch
                    fn.restype = res
                    fn.argtypes = args
                _lib = handle
    return _lib


def call(name: str, *args) -> None:
    """call an int-returning entry point and raise P3DError on a negative status"""
    handle = lib()
    status = getattr(handle, name)(*args)
    if status != 0:
        raise P3DError("%s failed (%d): %s" % (name, status, handle.p3d_last_error().decode()))


def launch_count() -> int:
    return int(lib().p3d_launch_count())
