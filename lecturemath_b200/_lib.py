"""ctypes binding of libaccessmath_b200.so (include/accessmath_b200.h).

There is NO CPU fallback: if the library is missing or no CUDA device is visible the product raises."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AM_B200_LIB") or os.path.join(_HERE, "libaccessmath_b200.so")      # AM_B200_LIB: tuning variants only

c_int, c_void_p, c_double, c_ll, c_ull = ctypes.c_int, ctypes.c_void_p, ctypes.c_double, ctypes.c_longlong, ctypes.c_ulonglong

# name -> (restype, argtypes): one entry per symbol declared in include/accessmath_b200.h
SIGNATURES = {
    "CC_AgeBoundaries": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int] + [c_void_p] * 6),
    "am_cc_age_boundaries_dev": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "adapthisteq": (c_int, [c_void_p, c_int, c_int, c_double, c_int, c_int, c_void_p]),
    "regionCumulativeDistribution": (None, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_double, c_void_p]),
    "combine_results": (c_int, [c_void_p, c_void_p, c_int, c_int, ctypes.c_ubyte, c_void_p]),
    "speaker_detection_handle_frame": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "am_adapthisteq_dev": (c_int, [c_void_p, c_int, c_int, c_double, c_int, c_int, c_void_p, c_void_p]),
    "am_region_cdf_dev": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_double, c_void_p, c_void_p]),
    "am_combine_results_dev": (c_int, [c_void_p, c_void_p, c_int, c_int, ctypes.c_ubyte, c_void_p, c_void_p]),
    "am_speaker_detection_dev": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "am_version": (c_int, []),
    "am_device_count": (c_int, []),
    "am_words_per_row": (c_int, [c_int]),
    "am_pack_mask_u8": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "am_pack_mask_u8_exact": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "am_unpack_mask_u8": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "am_cc_create": (c_void_p, [c_int] * 8),
    "am_cc_destroy": (None, [c_void_p]),
    "am_cc_label_batch": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "am_cc_counts": (c_int, [c_void_p, c_int, c_void_p, c_void_p]),
    "am_cc_read_label_table": (c_int, [c_void_p, c_int, c_int] + [c_void_p] * 6),
    "am_cc_read_kept": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "am_cc_read_crops": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "am_cc_pack_rows": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p]),
    "am_est_create": (c_void_p, [c_int, c_int, c_double, c_double, c_int, c_int, c_int, c_ll]),
    "am_est_destroy": (None, [c_void_p]),
    "am_est_add_frames": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "am_est_state": (c_int, [c_void_p, c_void_p, c_void_p]),
    "am_est_read_uniques": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "am_est_read_unique_crop": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "am_est_export_sizes": (c_int, [c_void_p, c_void_p, c_void_p]),
    "am_est_export": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "am_est_import": (c_int, [c_void_p, c_int, c_int, c_int, c_ull, c_void_p, c_void_p, c_ll, c_void_p]),
    "am_est_export_dev": (c_int, [c_void_p, c_void_p, c_ll, c_void_p]),
    "am_est_import_dev": (c_int, [c_void_p, c_void_p, c_void_p]),
    "am_p2p_alloc": (c_void_p, [c_ll]),
    "am_p2p_free": (c_int, [c_void_p]),
    "am_p2p_export_handle": (c_int, [c_void_p, c_void_p]),
    "am_p2p_open_handle": (c_void_p, [c_void_p]),
    "am_p2p_close_handle": (c_int, [c_void_p]),
    "am_stream_write32": (c_int, [c_void_p, ctypes.c_uint, c_void_p]),
    "am_stream_wait_geq32": (c_int, [c_void_p, ctypes.c_uint, c_void_p]),
    "am_conv_gemm": (c_int, [c_void_p, c_void_p]),
    "am_conv_plan_create": (c_void_p, [c_void_p]),
    "am_conv_plan_destroy": (None, [c_void_p]),
    "am_conv_plan_launch": (c_int, [c_void_p, c_void_p]),
    "am_conv_plan_info": (c_int, [c_void_p, c_void_p]),
    "am_conv_plan_bind": (c_int, [c_void_p, c_int, c_void_p]),
    "am_fcn_prep_input": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_int, c_void_p]),
    "am_fcn_maxpool2": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p]),
    "am_fcn_fill_border": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "am_fcn_heads_post": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "am_fcn_threshold_pack": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "am_fcn_working_size": (c_int, [c_int, c_int, c_void_p, c_void_p]),
    "am_frame_sums_bits": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "am_frame_sums_u8": (c_int, [c_void_p, c_int, c_ll, c_void_p, c_void_p]),
    "am_png1_size": (c_ll, [c_int, c_int]),
    "am_png1_encode": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "am_png1_capacity": (c_ll, [c_int, c_int]),
    "am_png1_encode_deflate": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "am_png8_capacity": (c_ll, [c_int, c_int]),
    "am_png8_encode_deflate": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "am_png1_scanlines_to_bits": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "am_est_unique_view": (c_int, [c_void_p, c_void_p, c_void_p]),
    "am_group_overlaps": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_ll, c_void_p, c_void_p]),
    "am_group_images": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_double, c_void_p, c_void_p, c_void_p]),
    "am_paint_frames": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "am_lanczos_resize_u8": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "am_resize_linear_u8": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "am_bits_resize_nearest": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
}

_lib = None


class AccessMathB200Error(RuntimeError):
    pass


def load():
    """dlopen the library and bind every declared symbol (no device needed)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise AccessMathB200Error(
                "libaccessmath_b200.so is not built (%s); run `python -m lecturemath_b200.build`. "
                "There is no CPU fallback." % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def lib():
    """The bound library, after checking that a CUDA device is present."""
    l = load()
    if l.am_device_count() <= 0:
        raise AccessMathB200Error("no CUDA device visible: lecturemath_b200 has no CPU fallback")
    return l


def check(rc, what):
    if rc != 0:
        raise AccessMathB200Error("%s failed with code %d (1=CUDA error, 2=bad argument, 3=capacity exceeded)" % (what, rc))
