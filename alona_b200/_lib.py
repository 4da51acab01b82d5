"""ctypes binding of libalona_b200.so (include/alona_b200.h).

There is no CPU fallback: if the shared library cannot be loaded, or no sm_100 GPU is
present, every compute entry point raises.
"""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
# ALONA_B200_LIB: load another build of the same library (A/B timing of kernel variants on one box)
_LIB_PATH = os.environ.get('ALONA_B200_LIB') or os.path.join(_HERE, 'libalona_b200.so')

_lib = None
_lock = threading.Lock()

c_void_p = ctypes.c_void_p
c_i64 = ctypes.c_int64
c_i32 = ctypes.c_int32
c_f64 = ctypes.c_double
c_size = ctypes.c_size_t
c_u64 = ctypes.c_uint64
p_i64 = ctypes.POINTER(ctypes.c_int64)

# name -> (restype, argtypes); mirrors include/alona_b200.h one to one
SIGNATURES = {
    'ab_version': (ctypes.c_int, []),
    'ab_last_error': (ctypes.c_char_p, []),
    'ab_source_digest': (ctypes.c_char_p, []),
    'ab_ctx_create': (ctypes.c_int, [ctypes.POINTER(c_void_p), ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    'ab_ctx_destroy': (ctypes.c_int, [c_void_p]),
    'ab_nccl_unique_id': (ctypes.c_int, [ctypes.c_char_p, c_void_p]),
    'ab_ctx_init_nccl': (ctypes.c_int, [c_void_p, ctypes.c_char_p, c_void_p]),
    'ab_allreduce_f64': (ctypes.c_int, [c_void_p, c_void_p, c_i64, c_void_p]),
    'ab_allreduce_i64': (ctypes.c_int, [c_void_p, c_void_p, c_i64, c_void_p]),
    'ab_allgather_bytes': (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_i64, c_void_p]),
    'ab_peer_handle_bytes': (c_size, []),
    'ab_peer_alloc': (ctypes.c_int, [c_void_p, c_void_p]),
    'ab_peer_open': (ctypes.c_int, [c_void_p, c_void_p]),
    'ab_peer_enabled': (ctypes.c_int, [c_void_p]),
    'ab_norm_moments': (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_i32, c_f64,
                                       c_void_p, c_void_p, c_void_p]),
    'ab_moments_finalize': (ctypes.c_int, [c_void_p, c_void_p, c_i32, c_void_p, c_void_p, c_void_p]),
    'ab_norm_moments_linear': (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_i32, c_f64,
                                              c_void_p, c_void_p, c_void_p]),
    'ab_hvg_seurat': (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_i64, c_i32, c_i32, c_i32,
                                     c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    'ab_hvg_extract_ws_bytes': (c_size, [c_i64]),
    'ab_hvg_extract_count': (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_i64, c_void_p, c_void_p, p_i64,
                                            c_void_p, c_size, c_void_p]),
    'ab_hvg_extract_fill': (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_void_p,
                                           c_f64, c_void_p, c_void_p, c_void_p, c_void_p]),
    'ab_qc_metrics': (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_i32, c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_void_p, c_void_p]),
    'ab_filter_ws_bytes': (c_size, [c_i64]),
    'ab_filter_count': (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_void_p, c_void_p, p_i64, p_i64,
                                       c_void_p, c_size, c_void_p]),
    'ab_filter_fill': (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_void_p, c_void_p, c_void_p, c_void_p,
                                      c_void_p, c_void_p, c_void_p, c_size, c_void_p]),
    'ab_cluster_moments': (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_i32, c_f64, c_void_p,
                                          c_i32, c_void_p, c_void_p, c_void_p, c_void_p]),
    'ab_cluster_median_ws_bytes': (c_size, [c_i32, c_i32, c_i64]),
    'ab_cluster_median_count': (ctypes.c_int, [c_void_p, c_void_p, p_i64, c_i32, c_i32, p_i64, c_void_p, c_size, c_void_p]),
    'ab_cluster_median_fill': (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_i32, c_f64, c_void_p,
                                              c_i32, c_void_p, p_i64, c_i64, c_void_p, c_void_p, c_size, c_void_p]),
    'ab_irlb_ws_bytes': (c_size, [c_i64, c_i32, c_i32]),
    'ab_irlb_ws_bytes_compact': (c_size, [c_i64, c_i32, c_i32, c_i64]),
    'ab_irlb': (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_i64, c_i32, c_i32, c_f64, c_i32,
                               c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, p_i64, c_void_p, c_size,
                               c_void_p]),
    'ab_knn_ws_bytes': (c_size, [c_i64, c_i32, c_i64, c_i32]),
    'ab_knn_schedule': (ctypes.c_int, [c_i64, c_i32, c_i32, c_void_p]),
    'ab_knn': (ctypes.c_int, [c_void_p, c_void_p, c_i64, c_i32, c_i64, c_i64, c_i32, c_i32, c_void_p, c_void_p,
                              c_void_p, c_void_p, c_size, c_void_p]),
    'ab_snn_ws_bytes': (c_size, [c_i64, c_i32, c_i64]),
    'ab_snn_count': (ctypes.c_int, [c_void_p, c_void_p, c_i64, c_i32, c_i64, c_i64, c_i32, c_void_p, p_i64, p_i64,
                                    c_void_p, c_size, c_void_p]),
    'ab_snn_fill': (ctypes.c_int, [c_void_p, c_void_p, c_i64, c_i32, c_i64, c_i64, c_i32, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_void_p, c_size, c_void_p]),
    'ab_edges_csv_ws_bytes': (c_size, [c_i64]),
    'ab_edges_csv_count': (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_i64, c_void_p, p_i64, c_void_p, c_size, c_void_p]),
    'ab_edges_csv_fill': (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_i64, c_void_p, c_void_p, c_void_p]),
    'ab_synth_ws_bytes': (c_size, [c_i64]),
    'ab_synth_nb_count': (ctypes.c_int, [c_void_p, c_void_p, c_i32, c_i32, c_i64, c_i64, c_f64, c_f64, c_f64, c_u64,
                                         c_void_p, p_i64, c_void_p, c_size, c_void_p]),
    'ab_synth_nb_fill': (ctypes.c_int, [c_void_p, c_void_p, c_i32, c_i32, c_i64, c_i64, c_f64, c_f64, c_f64, c_u64,
                                        c_void_p, c_void_p, c_void_p, c_void_p]),
    'ab_launch_count': (c_i64, [c_void_p]),
    'ab_prof_set_level': (ctypes.c_int, [c_void_p, ctypes.c_int]),
    'ab_prof_slot_name': (ctypes.c_char_p, [ctypes.c_int]),
    'ab_prof_read': (ctypes.c_int, [c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_double), p_i64]),
}


def lib_path():
    return _LIB_PATH


def load():
    """Loads the shared library (building it in-tree with nvcc if it is missing)."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.environ.get('ALONA_B200_LIB'):
            # build() is a no-op when the library carries the digest of the sources next to it; a stale library
            # (sources changed since it was built) is rebuilt when nvcc is there, else refused below
            from . import build as _build
            if not _build.is_fresh():
                import shutil
                if shutil.which(os.environ.get('NVCC', 'nvcc')):
                    _build.build()
                elif os.path.exists(_LIB_PATH):
                    raise RuntimeError('alona_b200: %s was built from other sources (digest %s, tree %s) and nvcc is not '
                                       'available to rebuild it' % (_LIB_PATH, _build.lib_digest(), _build._digest()[:16]))
        try:
            lib = ctypes.CDLL(_LIB_PATH)
        except OSError as e:
            raise RuntimeError('alona_b200: cannot load %s (%s). The CUDA extension is required; '
                               'there is no CPU fallback. Build it with `python -m alona_b200.build`.'
                               % (_LIB_PATH, e))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)            # AttributeError here = header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return _lib


def last_error():
    return load().ab_last_error().decode('utf-8', 'replace')


def check(rc, what=''):
    if rc != 0:
        raise RuntimeError('alona_b200 %s failed: %s' % (what, last_error()))
