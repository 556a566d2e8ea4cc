"""ctypes binding of the C-ABI engine library (include/segb200.h).

There is NO fallback: if ``libsegb200.so`` is missing or a call fails, a RuntimeError is raised
(the reference's own convention for its native op: ``AT_ERROR("Not implemented on the CPU")``,
segmentron/modules/csrc/criss_cross_attention/ca.h:34).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SEGB200_LIB", os.path.join(_HERE, "libsegb200.so"))   # override: A/B builds only

BF16, F16, F32 = 0, 1, 2
ACT_NONE, ACT_RELU, ACT_RELU6 = 0, 1, 2
ACT = {None: ACT_NONE, "none": ACT_NONE, "relu": ACT_RELU, "relu6": ACT_RELU6}

i32 = C.c_int32
vp = C.c_void_p


class ConvArgs(C.Structure):
    _fields_ = [("x", vp), ("wgt", vp), ("scale", vp), ("shift", vp), ("residual", vp), ("y", vp),
                ("n", i32), ("h", i32), ("w", i32), ("cin", i32), ("x_ld", i32),
                ("ho", i32), ("wo", i32), ("cout", i32), ("y_ld", i32), ("res_ld", i32),
                ("kh", i32), ("kw", i32), ("stride", i32), ("dilation", i32), ("pad_t", i32), ("pad_l", i32),
                ("act", i32), ("dtype", i32), ("max_ctas", i32), ("y_f32", i32)]


class DwArgs(C.Structure):
    _fields_ = [("x", vp), ("wgt", vp), ("shift", vp), ("y", vp),
                ("n", i32), ("h", i32), ("w", i32), ("c", i32), ("x_ld", i32), ("y_ld", i32),
                ("ho", i32), ("wo", i32), ("stride", i32), ("dilation", i32),
                ("pre_relu", i32), ("act", i32), ("dtype", i32)]


class WgradArgs(C.Structure):
    _fields_ = [("x", vp), ("dy", vp), ("dw", vp),
                ("n", i32), ("h", i32), ("w", i32), ("cin", i32), ("x_ld", i32),
                ("ho", i32), ("wo", i32), ("cout", i32), ("dy_ld", i32),
                ("kh", i32), ("kw", i32), ("stride", i32), ("dilation", i32), ("pad_t", i32), ("pad_l", i32),
                ("dtype", i32), ("max_ctas", i32), ("splits", i32)]


ll = C.c_longlong
ci = C.c_int
cf = C.c_float

# name -> (restype, argtypes); kept in one table so tests can check every header symbol is exported
SYMBOLS = {
    "segb200_version": (C.c_int, []),
    "segb200_last_error": (C.c_char_p, []),
    "segb200_set_option": (C.c_int, [C.c_char_p, C.c_int]),
    "segb200_debug_set_counters": (C.c_int, [vp]),
    "segb200_conv_kblock": (C.c_int, [C.c_int]),
    "segb200_conv_gemm": (C.c_int, [C.POINTER(ConvArgs), vp]),
    "segb200_dwconv3x3": (C.c_int, [C.POINTER(DwArgs), vp]),
    "segb200_pack_s2d": (C.c_int, [vp, C.c_int, vp, C.c_int] + [C.c_int] * 5 + [vp]),
    "segb200_global_avgpool": (C.c_int, [vp, vp] + [C.c_int] * 6 + [vp]),
    "segb200_adaptive_avgpool": (C.c_int, [vp, vp] + [C.c_int] * 8 + [vp]),
    "segb200_maxpool3x3s2": (C.c_int, [vp, vp] + [C.c_int] * 7 + [vp]),
    "segb200_upsample_add": (C.c_int, [vp, vp, vp] + [C.c_int] * 10 + [vp]),
    "segb200_bilinear_nhwc": (C.c_int, [vp, vp] + [C.c_int] * 10 + [vp]),
    "segb200_bilinear_nchw_out": (C.c_int, [vp, vp, vp] + [C.c_int] * 10 + [vp]),
    "segb200_pam_attention": (C.c_int, [vp] * 9 + [C.c_int] * 9 + [vp]),
    "segb200_nonlocal_attention": (C.c_int, [vp] * 9 + [C.c_int] * 10 + [vp]),
    "segb200_cam_softmax": (C.c_int, [vp, vp] + [C.c_int] * 5 + [vp]),
    "segb200_cca_weight_softmax": (C.c_int, [vp, vp, vp] + [C.c_int] * 8 + [vp]),
    "segb200_cca_map": (C.c_int, [vp, vp, vp, vp, vp] + [C.c_int] * 9 + [vp]),
    "segb200_nhwc_to_cn": (C.c_int, [vp, vp] + [C.c_int] * 6 + [vp]),
    "segb200_nchw_to_nhwc": (C.c_int, [vp, C.c_int, vp, C.c_int] + [C.c_int] * 5 + [vp]),
    "segb200_nhwc_to_nchw": (C.c_int, [vp, C.c_int, vp, C.c_int] + [C.c_int] * 5 + [vp]),
    # ---- training path ----
    "segb200_conv_wgrad": (ci, [C.POINTER(WgradArgs), vp]),
    "segb200_wgrad_debug_swap": (ci, [ci]),
    "segb200_reduce_slabs": (ci, [ll, ci, ci]),
    "segb200_reduce_partials": (ci, [vp, ci, ci, ci, vp, ll, ll, ci, cf, vp]),
    "segb200_bn_stats": (ci, [vp, ll, ci, ci, ci, vp, ci, vp]),
    "segb200_bn_finalize": (ci, [vp, ci, ci, C.c_double, vp, vp, vp, vp, cf, cf, vp, vp, vp, vp, vp]),
    "segb200_bn_apply": (ci, [vp, vp, vp, vp, vp, vp, ll, ll, ci, ci, ci, ci, ci, ci, vp]),
    "segb200_bn_bwd_reduce": (ci, [vp] * 7 + [ll, ll] + [ci] * 7 + [vp]),
    "segb200_bn_bwd_finalize": (ci, [vp, ci, ci, vp, vp, vp, vp, vp, vp]),
    "segb200_bn_bwd_apply": (ci, [vp] * 8 + [C.c_double, vp, vp, vp, ci, ll, ll] + [ci] * 8 + [vp]),
    "segb200_maxpool3x3s2_bwd": (ci, [vp, vp, vp] + [ci] * 8 + [vp]),
    "segb200_maxpool3x3s2_idx": (ci, [vp, vp, vp] + [ci] * 7 + [vp]),
    "segb200_maxpool3x3s2_bwd_idx": (ci, [vp, vp, vp] + [ci] * 7 + [vp]),
    "segb200_bilinear_nhwc_bwd": (ci, [vp, vp] + [ci] * 10 + [vp, ci, vp]),
    "segb200_upsample_ce_blocks": (ci, [ci, ci, ci]),
    "segb200_upsample_ce": (ci, [vp, vp, vp, vp, vp] + [ci] * 11 + [vp]),
    "segb200_dw_wgrad": (ci, [vp, vp, vp] + [ci] * 10 + [vp]),
    "segb200_nc_broadcast": (ci, [vp, vp, ci, ll, ci, ci, ci, cf, ci, ci, vp]),
    "segb200_stride2_place": (ci, [vp, vp] + [ci] * 8 + [vp]),
    "segb200_gather_cast": (ci, [vp, vp, vp, ll, ci, vp]),
    "segb200_scatter_add": (ci, [vp, vp, vp, ll, vp]),
    "segb200_cca_weight_bwd_blocks": (ci, [ci, ci, ci]),
    "segb200_cca_weight_bwd": (ci, [vp] * 6 + [ci] * 8 + [vp]),
    "segb200_cca_gather": (ci, [vp, vp, vp] + [ci] * 7 + [cf, ci, ci, vp]),
    "segb200_cca_scatter": (ci, [vp, vp, vp] + [ci] * 7 + [cf, vp, ci, ci, vp]),
    "segb200_sgd_step": (ci, [vp, vp, vp, ll, cf, cf, cf, cf, vp]),
    "segb200_adaptive_avgpool_bwd": (ci, [vp, vp] + [ci] * 9 + [vp]),
    "segb200_row_softmax": (ci, [vp, vp] + [ci] * 5 + [vp]),
    "segb200_row_softmax_bwd": (ci, [vp, vp, vp, vp, vp] + [ci] * 6 + [vp]),
    "segb200_cam_softmax_bwd": (ci, [vp, vp, vp, vp, vp] + [ci] * 5 + [vp]),
    "segb200_cam_bwd_pack": (ci, [vp, vp, vp, vp, vp] + [ci] * 5 + [vp]),
    "segb200_upsample_add_bwd": (ci, [vp, vp, vp, vp] + [ci] * 13 + [vp]),
    "segb200_dw_wgrad_v2_slabs": (ci, [ll, ci, ci]),
    "segb200_dw_wgrad_v2": (ci, [vp, vp, vp] + [ci] * 10 + [vp]),
    # ---- SyncBatchNorm exchange over peer memory ----
    "segb200_syncbn_slot_floats": (ci, [ci, ci]),
    "segb200_syncbn_slot_flags": (ci, [ci]),
    "segb200_counter_add": (ci, [vp, ci, vp]),
    "segb200_bn_finalize_sync": (ci, [vp, ci, ci, C.c_double, vp, vp, vp, vp, cf, cf, vp, vp, vp, vp, vp, ci, ci, ci, ll, ll, vp, vp]),
    "segb200_bn_bwd_finalize_sync": (ci, [vp, ci, ci, vp, vp, vp, vp, vp, vp, ci, ci, ci, ll, ll, vp, vp]),
    # ---- segmentron._C (NCHW, the reference's own layout) ----
    "segb200_ca_forward": (ci, [vp, vp, vp] + [ci] * 5 + [vp]),
    "segb200_ca_backward": (ci, [vp] * 5 + [ci] * 5 + [vp]),
    "segb200_ca_map_forward": (ci, [vp, vp, vp] + [ci] * 5 + [vp]),
    "segb200_ca_map_backward": (ci, [vp] * 5 + [ci] * 5 + [vp]),
    # ---- evaluation metric ----
    "segb200_seg_metric": (ci, [vp, ci, vp, ci, ci, ci, ci, vp, vp]),
    "segb200_seg_metric_lowres": (ci, [vp, ci, ci, ci, ci, ci, ci, vp, ci, ci, ci, ci, vp, vp]),
    "segb200_seg_metric_accumulate": (ci, [vp, ci, vp, vp, vp, vp]),
    "segb200_image_normalize": (ci, [vp, vp, ci, ci, ci, ci, vp, vp, vp]),
    # ---- multi-scale + flip evaluation ----
    "segb200_eval_prepare": (ci, [vp, vp] + [ci] * 9 + [vp]),
    "segb200_eval_accumulate": (ci, [vp, vp] + [ci] * 11 + [vp]),
}

_lib = None


def load():
    """Load the engine library; raises RuntimeError (never falls back) if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"segb200: native engine library not found at {LIB_PATH}; build it with "
                "`python -c 'import __graft_entry__ as g; g.build()'` (there is no CPU/PyTorch fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            try:
                fn = getattr(lib, name)
            except AttributeError:
                if "SEGB200_LIB" in os.environ:      # A/B builds of older sources may lack newer diagnostics
                    continue
                raise
            fn.restype, fn.argtypes = res, args
        # A/B measurements only: SEGB200_OPTS="gemm_dual=0,dw_cols2=0" applies segb200_set_option knobs at load time
        for kv in filter(None, os.environ.get("SEGB200_OPTS", "").split(",")):
            k, v = kv.split("=")
            if lib.segb200_set_option(k.strip().encode(), int(v)) != 0:
                raise RuntimeError(f"segb200: unknown option in SEGB200_OPTS: {kv}")
        _lib = lib
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = load().segb200_last_error().decode(errors="replace")
        raise RuntimeError(f"segb200 {what} failed (code {rc}): {msg}")
