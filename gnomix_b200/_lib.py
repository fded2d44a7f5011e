"""ctypes binding of libgnx.so (include/gnx.h).  There is no CPU fallback: a missing
library or a missing sm_100 device is an error, raised loudly."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgnx.so")
_lib = None

c_i64 = C.c_int64
c_vp = C.c_void_p


class GnxError(RuntimeError):
    pass


_SIGS = {
    "gnx_version": (C.c_int, []),
    "gnx_last_error": (C.c_char_p, []),
    "gnx_device_count": (C.c_int, []),
    "gnx_selftest_math": (C.c_int, [c_i64, C.c_uint64, c_vp]),
    "gnx_lr_model_create": (C.c_int, [C.POINTER(c_vp), C.c_int, c_i64, c_i64, c_i64, c_vp, c_vp, C.c_int]),
    "gnx_lr_model_destroy": (None, [c_vp]),
    "gnx_lr_model_scale": (C.c_int, [c_vp]),
    "gnx_lr_model_windows": (C.c_int, [c_vp]),
    "gnx_lr_set_kernel": (C.c_int, [c_vp, C.c_int]),
    "gnx_lr_predict": (C.c_int, [c_vp, c_vp, c_i64, c_i64, c_vp, c_vp]),
    "gnx_lr_predict_f64": (C.c_int, [c_vp, c_vp, c_i64, c_i64, c_vp, c_vp]),
    "gnx_gbt_model_create": (C.c_int, [C.POINTER(c_vp), C.c_int, C.c_int, C.c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "gnx_gbt_model_destroy": (None, [c_vp]),
    "gnx_gbt_set_kernel": (C.c_int, [c_vp, C.c_int]),
    "gnx_gbt_set_profile": (C.c_int, [c_vp, C.c_int]),
    "gnx_gbt_last_phase_ms": (C.c_int, [c_vp, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "gnx_gbt_smooth": (C.c_int, [c_vp, c_vp, c_i64, C.c_int, c_vp, c_vp, c_vp]),
    "gnx_gbt_rows": (C.c_int, [c_vp, c_vp, c_i64, c_vp, c_vp]),
    "gnx_crf_model_create": (C.c_int, [C.POINTER(c_vp), C.c_int, C.c_int, c_vp, c_vp]),
    "gnx_crf_model_destroy": (None, [c_vp]),
    "gnx_crf_smooth": (C.c_int, [c_vp, c_vp, c_i64, C.c_int, c_vp, c_vp, c_vp]),
    "gnx_svc_model_create": (C.c_int, [C.POINTER(c_vp), C.c_int, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, C.c_int]),
    "gnx_svc_model_destroy": (None, [c_vp]),
    "gnx_svc_predict": (C.c_int, [c_vp, c_vp, c_i64, c_i64, c_vp, c_vp]),
    "gnx_svc_kernel_window": (C.c_int, [c_vp, C.c_int, c_vp, c_i64, c_i64, c_vp, c_vp]),
    "gnx_gnofix": (C.c_int, [c_vp, c_vp, c_i64, c_i64, c_vp, c_i64, C.c_int, C.c_int, c_vp, c_vp, c_vp]),
    "gnx_cal_model_create": (C.c_int, [C.POINTER(c_vp), C.c_int, C.c_int, c_vp, c_vp, c_vp]),
    "gnx_cal_model_destroy": (None, [c_vp]),
    "gnx_calibrate": (C.c_int, [c_vp, c_vp, C.c_int, c_i64, c_vp, c_vp, c_vp]),
    "gnx_gnofix_last_stats": (C.c_int, [c_vp]),
    "gnx_gnofix_crf": (C.c_int, [c_vp, C.c_int, c_vp, c_i64, c_i64, c_vp, c_i64, C.c_int, C.c_int, c_vp, c_vp, c_vp]),
    "gnx_gnofix_crf_last_stats": (C.c_int, [c_vp]),
    "gnx_infer_host": (C.c_int, [c_vp, c_vp, c_vp, c_i64, c_i64, c_vp, c_vp, c_i64]),
    "gnx_infer_host_ex": (C.c_int, [c_vp, c_vp, c_i64, c_i64, c_vp, c_vp, c_vp, c_i64]),
    "gnx_vcf_to_haplotypes_packed": (C.c_int, [c_vp, c_i64, c_i64, c_vp, c_vp, c_vp, c_i64, c_i64, C.c_int, c_vp, c_i64, C.c_int]),
    "gnx_host_alloc_pinned": (C.c_int, [C.POINTER(c_vp), c_i64]),
    "gnx_host_free_pinned": (C.c_int, [c_vp]),
    "gnx_upload_haplotypes": (C.c_int, [c_vp, c_i64, c_i64, c_i64, c_vp, c_i64]),
    "gnx_release_workspace": (C.c_int, []),
    "gnx_pack_rows_host": (C.c_int, [c_vp, c_i64, c_i64, c_i64, c_vp, c_i64, C.c_int, C.POINTER(C.c_int)]),
    "gnx_unpack_dev": (C.c_int, [c_vp, c_i64, c_i64, c_i64, c_vp, c_i64, c_vp]),
    "gnx_host_threads": (C.c_int, []),
    "gnx_vcf_open": (C.c_int, [C.POINTER(c_vp), C.c_char_p, C.c_char_p, C.c_int]),
    "gnx_vcf_close": (None, [c_vp]),
    "gnx_vcf_num_records": (c_i64, [c_vp]),
    "gnx_vcf_num_samples": (c_i64, [c_vp]),
    "gnx_vcf_copy": (C.c_int, [c_vp, c_vp, c_vp, c_vp]),
    "gnx_vcf_strings": (c_i64, [c_vp, C.c_int, c_vp, c_i64]),
    "gnx_vcf_to_haplotypes": (C.c_int, [c_vp, c_i64, c_i64, c_vp, c_vp, c_vp, c_i64, c_i64, C.c_int, c_vp, c_i64, C.c_int]),
    "gnx_write_fb_body": (C.c_int, [C.c_char_p, C.c_int, c_vp, C.c_int, c_i64, c_i64, c_i64, c_vp, C.c_int]),
    "gnx_write_vcf_body": (C.c_int, [C.c_char_p, C.c_int, c_i64, c_i64, c_vp, c_i64, c_vp, C.c_char_p, c_i64, C.c_char_p, c_i64,
                                     C.c_char_p, c_i64, C.c_char_p, c_i64, C.c_char_p, c_i64, C.c_int]),
    "gnx_format_floats": (c_i64, [c_vp, C.c_int, c_i64, c_vp, c_i64]),
    "gnx_infer_host_rates": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "gnx_infer_host_last_transfer": (C.c_int, [C.POINTER(C.c_double), C.POINTER(c_i64), C.POINTER(c_i64)]),
}

EXPORTS = tuple(_SIGS)


class Pipeline(C.Structure):
    """gnx_pipeline_t (include/gnx.h)."""
    _fields_ = [("lr", c_vp), ("svc", c_vp), ("gbt", c_vp), ("crf", c_vp), ("cal", c_vp),
                ("phase", C.c_int), ("max_it", C.c_int), ("x_packed", C.c_int), ("crf_phase_S", C.c_int)]


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GnxError(
                "libgnx.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `python -m gnomix_b200.build`. gnomix_b200 has no CPU fallback." % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().gnx_last_error()
        raise GnxError("%s failed (rc=%d): %s" % (what or "libgnx call", rc, msg.decode() if msg else "?"))


def require_gpu():
    import torch
    if not torch.cuda.is_available() or lib().gnx_device_count() < 1:
        raise GnxError("gnomix_b200 needs a B200 (sm_100) GPU; none is visible and there is no CPU fallback")
