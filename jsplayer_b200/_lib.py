"""ctypes binding of libjsplayer_cuda.so (include/jsplayer_cuda.h). Fails loudly when the library is missing."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libjsplayer_cuda.so")

JSP_N_KERNELS = 8
KERNEL_NAMES = ["msv1_decode", "frame_copy", "sp_entropy_rc", "sp_entropy_ans", "sp_recon", "signif", "sp_entropy_mixed", "k7"]
JSP_BATCH_SIGNIFICANCE = 1
JSP_BATCH_NUMA_BIND = 2
JSP_BATCH_DISPLAY = 4
JSP_BATCH_DISPLAY_FLIP = 8
JSP_FRAME_CHANGED, JSP_FRAME_SIGNIFICANT, JSP_FRAME_ERROR, JSP_FRAME_DIFFERS = 1, 2, 4, 8
JSP_DISPLAY_FLIP = 1


FRAME_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_uint8)      # jsp_frame_fn


class PFrameResultC(C.Structure):
    _fields_ = [("data_pnt", C.c_void_p), ("significant_changes", C.c_int32)]


class AviInfoC(C.Structure):
    _fields_ = [("codec", C.c_int32), ("width", C.c_int32), ("height", C.c_int32), ("bpp", C.c_int32),
                ("fourcc", C.c_uint32), ("n_frames", C.c_int32), ("n_frames_header", C.c_int32),
                ("palette_bytes", C.c_int32), ("has_index", C.c_int32), ("fps", C.c_double)]


class StreamDescC(C.Structure):
    _fields_ = [("codec", C.c_int32), ("width", C.c_int32), ("height", C.c_int32), ("bpp", C.c_int32),
                ("palette", C.c_void_p), ("palette_bytes", C.c_int32), ("n_frames", C.c_int32),
                ("bytes", C.c_void_p), ("frame_off", C.c_void_p), ("frame_len", C.c_void_p),
                ("frame_key", C.c_void_p), ("sp_version", C.c_int32), ("reserved", C.c_int32)]


# every symbol include/jsplayer_cuda.h declares: name -> (restype, argtypes)
PROTOTYPES = {
    "jsp_device_count": (C.c_int, []),
    "jsp_last_error": (C.c_char_p, []),
    "jsp_version": (C.c_char_p, []),
    "jsp_host_alloc": (C.c_void_p, [C.c_size_t]),
    "jsp_host_free": (None, [C.c_void_p]),
    "jsp_numa_node_of_device": (C.c_int, [C.c_int]),
    "jsp_numa_bind_thread": (C.c_int, [C.c_int]),
    "jsp_host_d2h_gbs": (C.c_double, [C.c_int, C.c_size_t, C.c_int]),
    "jsp_create": (C.c_void_p, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int]),
    "jsp_destroy": (None, [C.c_void_p]),
    "jsp_preinit": (None, [C.c_void_p, C.c_int]),
    "jsp_previous_frame": (C.c_void_p, [C.c_void_p]),
    "jsp_is_key_frame": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "jsp_state_of": (C.c_int, [C.c_void_p]),
    "jsp_decompress_i": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "jsp_continue_i": (C.c_int, [C.c_void_p]),
    "jsp_decompress_p": (PFrameResultC, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "jsp_needs_index": (C.c_int, [C.c_void_p]),
    "jsp_stop_and_clean": (None, [C.c_void_p]),
    "jsp_batch_create": (C.c_void_p, [C.c_int, C.c_int, C.c_int]),
    "jsp_batch_destroy": (None, [C.c_void_p]),
    "jsp_batch_configure": (C.c_int64, [C.c_void_p, C.c_void_p, C.c_int]),
    "jsp_batch_upload": (C.c_int, [C.c_void_p]),
    "jsp_batch_run": (C.c_int, [C.c_void_p]),
    "jsp_batch_sync": (C.c_int, [C.c_void_p]),
    "jsp_batch_download": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "jsp_batch_download_display": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "jsp_batch_results": (C.c_int, [C.c_void_p, C.c_void_p]),
    "jsp_batch_next_significant": (C.c_int64, [C.c_void_p, C.c_int, C.c_int64]),
    "jsp_batch_decode_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "jsp_batch_decode_host_delta": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "jsp_batch_delta_bytes": (C.c_uint64, [C.c_void_p]),
    "jsp_batch_device_frame": (C.c_uint64, [C.c_void_p, C.c_int64]),
    "jsp_batch_time_runs": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "jsp_batch_stats": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "jsp_batch_symbols": (C.c_int64, [C.c_void_p]),
    "jsp_batch_kernel_bytes": (C.c_int, [C.c_void_p, C.c_void_p]),
    "jsp_avi_parse": (C.c_void_p, [C.c_void_p, C.c_uint64]),
    "jsp_avi_free": (None, [C.c_void_p]),
    "jsp_avi_get_info": (C.c_int, [C.c_void_p, C.c_void_p]),
    "jsp_avi_get_palette": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "jsp_avi_frame_table": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "jsp_avi_last_error": (C.c_char_p, []),
    "jsp_segment_stream": (C.c_int, [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "jsp_batch_decode": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
}

_lib = None


def load():
    """Loads libjsplayer_cuda.so; raises (never falls back) if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libjsplayer_cuda.so is not built (run `python jsplayer_b200/build.py`); "
            "jsplayer_b200 has no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error():
    return load().jsp_last_error().decode("utf-8", "replace")


def require_gpu():
    lib = load()
    if lib.jsp_device_count() <= 0:
        raise RuntimeError("no CUDA device visible: jsplayer_b200 decodes only on the GPU (no CPU fallback)")
    return lib
