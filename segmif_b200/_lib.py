"""ctypes binding of libsegmif_b200.so (the C ABI declared in include/segmif_b200.h).

There is deliberately no fallback: if the shared library is missing or the device is not an sm_100
part, importing the ops fails with a RuntimeError that says how to build it.  Nothing here (or anywhere
in the package) imports `oracle/`.
"""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsegmif_b200.so")

c_void_p, c_int, c_int64, c_float, c_size_t, c_char_p = (
    ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_size_t, ctypes.c_char_p)

F32, BF16 = 0, 1
ACT_NONE, ACT_RELU, ACT_PRELU, ACT_GELU = 0, 1, 2, 3


class ConvParams(ctypes.Structure):
    """Mirror of segmif_conv_params (include/segmif_b200.h)."""
    _fields_ = [
        ("src", c_void_p), ("weight", c_void_p), ("bias", c_void_p), ("prelu_alpha", c_void_p),
        ("residual", c_void_p), ("dst", c_void_p),
        ("B", c_int), ("H", c_int), ("W", c_int), ("Cin", c_int), ("ld_src", c_int), ("src_coff", c_int),
        ("KH", c_int), ("KW", c_int), ("stride", c_int), ("pad", c_int), ("dil", c_int), ("Ho", c_int),
        ("Wo", c_int), ("Cout", c_int),
        ("act", c_int),
        ("res_dtype", c_int), ("ld_res", c_int), ("res_coff", c_int),
        ("dst_dtype", c_int), ("ld_dst", c_int), ("dst_coff", c_int),
        ("pre_add", c_void_p), ("ld_pre", c_int), ("pre_coff", c_int),
    ]


class LinearParams(ctypes.Structure):
    """Mirror of segmif_linear_params."""
    _fields_ = [
        ("src", c_void_p), ("weight", c_void_p), ("bias", c_void_p), ("prelu_alpha", c_void_p),
        ("residual", c_void_p), ("dst", c_void_p),
        ("M", c_int), ("N", c_int), ("K", c_int), ("ld_src", c_int), ("src_coff", c_int),
        ("act", c_int),
        ("res_dtype", c_int), ("ld_res", c_int), ("res_coff", c_int),
        ("dst_dtype", c_int), ("ld_dst", c_int), ("dst_coff", c_int),
        ("weight_kn", c_int),
        ("row_scale", c_void_p), ("rows_per_scale", c_int),
    ]


class DrdbPushGroup(ctypes.Structure):
    _fields_ = [("bias", c_void_p), ("partial_in", c_void_p), ("dst", c_void_p), ("ld_partial_in", c_int),
                ("coff_partial_in", c_int), ("ld_dst", c_int), ("coff_dst", c_int), ("relu", c_int)]


class DrdbPushParams(ctypes.Structure):
    """Mirror of segmif_drdb_push_params."""
    _fields_ = [("src", c_void_p), ("weight", c_void_p), ("B", c_int), ("H", c_int), ("W", c_int), ("ld_src", c_int),
                ("slab_offset", c_int), ("slab_width", c_int), ("n_out", c_int), ("groups", DrdbPushGroup * 4)]


class SplitGemmParams(ctypes.Structure):
    """Mirror of segmif_split_gemm_params."""
    _fields_ = [
        ("a_planes", c_void_p), ("a_plane_stride", c_int64), ("w_planes", c_void_p), ("bias", c_void_p),
        ("prelu_alpha", c_void_p), ("residual", c_void_p), ("dst", c_void_p), ("dst_planes", c_void_p),
        ("dst_plane_stride", c_int64),
        ("M", c_int), ("N", c_int), ("K", c_int), ("ld_a", c_int), ("a_coff", c_int),
        ("act", c_int),
        ("ld_res", c_int), ("res_coff", c_int), ("ld_dst", c_int), ("dst_coff", c_int), ("ld_dp", c_int), ("dp_coff", c_int),
        ("nterms", c_int),
        ("B", c_int), ("H", c_int), ("W", c_int),
        ("ntaps", c_int),
        ("tap_dx", c_int * 9), ("tap_dy", c_int * 9),
    ]


class DrdbDataflowParams(ctypes.Structure):
    """Mirror of segmif_drdb_dataflow_params."""
    _fields_ = [("growth", c_void_p), ("ld", c_int), ("partial", c_void_p), ("ld_partial", c_int),
                ("w_push_a", c_void_p), ("w_push_b", c_void_p), ("w_pull", c_void_p * 4), ("bias", c_void_p * 5),
                ("w_1x1", c_void_p), ("bias_1x1", c_void_p), ("out", c_void_p), ("ld_out", c_int), ("out_coff", c_int),
                ("B", c_int), ("H", c_int), ("W", c_int), ("flags", c_void_p), ("ctas", c_int * 7)]


DP_MAX_OPS = 6


class DpSample(ctypes.Structure):
    """Mirror of segmif_dp_sample (training data path, include/segmif_b200.h)."""
    _fields_ = [("ir", c_void_p), ("vis", c_void_p), ("mask", c_void_p), ("label", c_void_p),
                ("H", c_int), ("W", c_int), ("nh", c_int), ("nw", c_int), ("resized", c_int), ("flip", c_int),
                ("pad_h", c_int), ("pad_w", c_int), ("PH", c_int), ("PW", c_int),
                ("cand_hs", c_int * 10), ("cand_ws", c_int * 10), ("hs", c_int), ("ws", c_int), ("n_ops", c_int),
                ("op_kind", c_int * DP_MAX_OPS), ("op_u8", c_int * DP_MAX_OPS), ("op_delta", c_int * DP_MAX_OPS),
                ("op_alpha", c_float * DP_MAX_OPS), ("op_beta", c_float * DP_MAX_OPS),
                ("ks_x", c_int), ("ks_y", c_int), ("mask_c", c_int), ("reserved", c_int), ("roi_x0", c_int), ("roi_x1", c_int), ("roi_y0", c_int), ("roi_y1", c_int),
                ("src_y0", c_int), ("src_y1", c_int),
                ("tab_off", c_int64), ("tmp_off", c_int64), ("rs_off", c_int64), ("lab_off", c_int64)]


P = c_void_p
# name -> argtypes; every function returns int except where noted in _RESTYPES
SIGNATURES = {
    "segmif_abi_version": [],
    "segmif_last_error": [],
    "segmif_init": [c_int],
    "segmif_layernorm_fwd": [P, c_int, P, P, P, c_int, c_int64, c_int, c_float, P],
    "segmif_conv_fwd": [ctypes.POINTER(ConvParams), P],
    "segmif_linear_tc_fwd": [ctypes.POINTER(LinearParams), P],
    "segmif_conv3x3_tc_fwd": [ctypes.POINTER(ConvParams), P],
    "segmif_drdb_push_tc_fwd": [ctypes.POINTER(DrdbPushParams), P],
    "segmif_patch_embed7_ln_fwd": [P, P, P, P, P, c_float, P, P, P, c_int, c_int, c_int, c_int, P],
    "segmif_sr_attention_fwd": [P, c_int, P, P, c_int, P, c_int, c_int, c_int, c_int, c_int, c_int, c_float, P],
    "segmif_dwconv3x3_gelu_fwd": [P, P, P, P, c_int, c_int, c_int, c_int, P],
    "segmif_bilinear_nhwc_fwd": [P, c_int, c_int, c_int, c_int, c_int, c_int, P, c_int, c_int, c_int, c_int, c_int, P],
    "segmif_upsample_argmax_fwd": [P, c_int, c_int, c_int, c_int, P, c_int, c_int, P],
    "segmif_nhwc_to_nchw": [P, c_int, c_int, c_int, P, c_int, c_int, c_int, P],
    "segmif_nchw_to_nhwc": [P, P, c_int, c_int, c_int, c_int, c_int, c_int, P],
    "segmif_conv3x3_in1_fwd": [P, c_int64, P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, P],
    "segmif_conv3x3_out1_fwd": [P, c_int, P, P, P, P, c_int, c_int, c_int, c_int, P],
    "segmif_ffm_gram_fwd": [P, c_int, c_int, P, c_int, c_int, P, c_int, c_int, P, P, P, c_int, c_int, c_int64, P],
    "segmif_ffm_ctx_fwd": [P, c_int, P, P, P, P, c_int, P],
    "segmif_ffm_apply_fwd": [P, c_int, c_int, P, c_int, c_int, P, c_int, c_int, P, P, P, P, P, P, c_float,
                             P, c_int, c_int, P, c_int, c_int, c_int, c_int64, P],
    "segmif_ffm_gram_lr_fwd": [P, c_int, c_int, P, c_int, c_int, P, c_int, c_int, c_int, c_int, P, P, P, c_int, c_int, P],
    "segmif_ffm_apply_lr_fwd": [P, c_int, c_int, P, c_int, c_int, P, c_int, c_int, c_int, c_int, P, P, P, P, P, P, c_float,
                                P, c_int, c_int, P, c_int, c_int, c_int, P],
    "segmif_rgb2ycrcb": [P, P, c_int, c_int64, P],
    "segmif_ycrcb2rgb": [P, P, c_int, c_int64, P],
    "segmif_recompose_rgb": [P, P, P, c_int, c_int, c_int64, P],
    "segmif_loss_workspace_bytes": [c_int, c_int, c_int],
    "segmif_ssim_fwd": [P, P, c_int, c_int, c_int, c_int, P, P, P],
    "segmif_laploss2_fwd": [P, P, P, c_int, c_int, c_int, P, P, P],
    "segmif_laploss_fwd": [P, P, c_int, c_int, c_int, P, P, P],
    "segmif_entropy_fwd": [P, c_int, c_int, c_int, c_int, P, P, P],
    "segmif_sobel_l1_fwd": [P, P, c_int, c_int, c_int, P, P, P],
    "segmif_mse_l1_fwd": [P, P, c_int64, P, P, P],
    "segmif_upsample_ce_fwd": [P, c_int, c_int, c_int, c_int, P, c_int, c_int, c_int, P, P, P],
    "segmif_mse_l1_bwd": [P, P, c_int64, P, P, c_int, P],
    "segmif_sobel_l1_bwd": [P, P, c_int, c_int, c_int, P, P, c_int, P],
    "segmif_ssim_bwd": [P, P, c_int, c_int, c_int, c_int, P, P, c_int, P],
    "segmif_laploss2_bwd": [P, P, P, c_int, c_int, c_int, P, P, c_int, P],
    "segmif_laploss_bwd": [P, P, c_int, c_int, c_int, P, P, c_int, P],
    "segmif_entropy_bwd": [P, c_int, c_int, c_int, c_int, P, P, c_int, P],
    "segmif_act_bwd": [P, c_int, c_int, P, c_int, c_int, P, c_int, c_int, c_int64, c_int, c_int, P, P, P, P],
    "segmif_prelu_plane_bwd": [P, P, c_int64, P, P, c_int, c_int, P, P, P],
    "segmif_colsum": [P, c_int, c_int, c_int64, c_int, P, P],
    "segmif_wgrad_lin_chunks": [c_int64, c_int, c_int],
    "segmif_wgrad_lin": [P, c_int, c_int, P, c_int, c_int, c_int64, c_int, c_int, P, c_int, P, c_int64, c_int64, c_int, c_int, P, P],
    "segmif_add_bf16": [P, c_int, c_int, P, c_int, c_int, P, c_int, c_int, c_int64, c_int, P],
    "segmif_layernorm_bwd": [P, c_int, P, c_int, c_int, c_int, P, c_float, P, c_int, c_int, c_int, c_int64, c_int, P, P, P, c_int, P],
    "segmif_wgrad_workspace_bytes": [c_int, c_int, c_int, c_int],
    "segmif_wgrad": [P, c_int, c_int, P, c_int, c_int, c_int, c_int, c_int, c_int64, c_int, c_int, c_int, c_int, P, c_int, P,
                     c_int64, c_int64, c_int64, c_int, c_int, P],
    "segmif_ffm_apply_train_fwd": [P, c_int, c_int, P, c_int, c_int, P, c_int, c_int, P, P, P, P, P, P, c_float,
                                   P, c_int, c_int, P, c_int, c_int, c_int, c_int64, P, P, P],
    "segmif_ffm_bwd_gram": [P, c_int, c_int, P, c_int, c_int, P, c_int, c_int, P, P, P, P, P, c_int, c_int, c_int64, P],
    "segmif_ffm_bwd_ctx": [P, c_int, P, c_int, P, P, P, P, P, P, P, c_int, P],
    "segmif_ffm_bwd_apply": [P, c_int, c_int, P, c_int, c_int, P, c_int, c_int, P, P, P, P, P, P, P, P, c_int, c_int64, P],
    "segmif_adamw_step": [P, P, P, P, c_int64, c_float, c_float, c_float, c_float, c_float, c_int, c_float, P],
    "segmif_sr_attention_train_fwd": [P, c_int, P, P, c_int, P, c_int, c_int, c_int, c_int, c_int, c_int, c_float, P, P],
    "segmif_sr_attention_tc_fwd": [P, c_int, P, P, c_int, P, c_int, c_int, c_int, c_int, c_int, c_int, c_float, P, P],
    "segmif_sr_attention_fa_fwd": [P, c_int, P, P, c_int, P, c_int, c_int, c_int, c_int, c_int, c_int, c_float, P, P],
    "segmif_sr_attention_bwd": [P, c_int, P, P, c_int, P, P, c_int, P, P, c_int, P, c_int, c_int, c_int, c_int, c_int, c_int,
                                c_int, c_float, P],
    "segmif_upsample_ce_bwd": [P, c_int, c_int, c_int, c_int, P, c_int, c_int, c_int, P, P, P, P],
    "segmif_bilinear_nhwc_bwd": [P, c_int, c_int, c_int, c_int, P, c_int, c_int, c_int, c_int, P],
    "segmif_bn_train_fwd": [P, c_int64, c_int, P, P, c_float, c_float, P, P, P, P, P, P],
    "segmif_bn_train_bwd": [P, P, P, P, P, c_int64, c_int, P, P, P, P, P],
    "segmif_bn_eval_fwd": [P, c_int64, c_int, P, P, c_float, P, P, P, P, P],
    "segmif_bn_eval_bwd": [P, P, P, P, P, c_int64, c_int, P, P, P, P, P],
    "segmif_channel_scale": [P, P, P, c_int, c_int64, c_int, P],
    "segmif_dwconv3x3": [P, P, P, P, c_int, c_int, c_int, c_int, c_int, P],
    "segmif_dwconv3x3_gelu_bwd": [P, P, P, P, P, c_int, c_int, c_int, c_int, P, P, P, P],
    "segmif_dwconv3x3_gelu_bwd_workspace": [c_int, c_int, c_int, c_int],
    "segmif_confusion_matrix": [P, P, c_int64, c_int, P, P],
    "segmif_dp_label_stage": [P, P, c_int, c_int, c_int, P, P, P, P, P],
    "segmif_dp_image_stage": [P, P, c_int, c_int, P, P, P, P, P, P, P, P, P, P, P],
    "segmif_dp_u8_to_chw_f64": [P, c_int, c_int, c_int, P, P],
    "segmif_fused_to_uint8": [P, P, P, c_int, c_int64, P],
    "segmif_wgrad_chunks": [c_int, c_int, c_int, c_int64, c_int, c_int, c_int, c_int],
    "segmif_col2im": [P, c_int, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P],
    "segmif_channel_affine_nchw": [P, P, P, P, c_int, c_int, c_int64, P],
    "segmif_recompose_rgb_bwd": [P, P, P, c_int, c_int, c_int64, P],
    "segmif_cast": [P, c_int, P, c_int, c_int64, P],
    "segmif_scale_cast_rows": [P, P, P, c_int64, c_int64, c_int, P],
    "segmif_scale_add_rows": [P, P, c_int, P, P, c_int64, c_int64, c_int, P],
    "segmif_sobel_map_fwd": [P, P, c_int, c_int, c_int, P],
    "segmif_sobel_map_bwd": [P, P, P, c_int, c_int, c_int, c_int, P],
    "segmif_ew2": [P, P, c_float, c_float, c_int, P, c_int64, P],
    "segmif_drdb_dataflow_workspace_bytes": [c_int, c_int],
    "segmif_drdb_dataflow_prepare": [c_int],
    "segmif_drdb_dataflow_fwd": [ctypes.POINTER(DrdbDataflowParams), P],
    "segmif_split_gemm_fwd": [ctypes.POINTER(SplitGemmParams), P],
    "segmif_split3": [P, c_int, c_int, c_int64, c_int, c_int, P, c_int, c_int, P, c_int, c_int, c_int64, P],
    "segmif_im2col_split3": [P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P, c_int64, P],
    "segmif_sr_attention_f32_fwd": [P, c_int, P, P, c_int, P, c_int, c_int, c_int, c_int, c_int, c_int, c_float, P],
    "segmif_dwconv3x3_f32_fwd": [P, P, P, P, c_int, c_int, c_int, c_int, c_int, P],
    "segmif_gram64_f64": [P, c_int, c_int, c_int, c_int64, c_int, P, c_int, P],
    "segmif_ffm_ctx_f64_fwd": [P, c_int, P, P, P, P, c_int, P],
    "segmif_xty_f64": [P, c_int, c_int, c_int, P, c_int, c_int, c_int, c_int, c_int64, P, c_int, P],
    "segmif_ctx_blockdiag": [P, c_int, c_int, c_int, c_float, P, P, c_int, P],
    "segmif_sigmoid_gate": [P, P, c_int64, P],
    "segmif_conv3x3_in1_f32_fwd": [P, c_int64, P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, P],
    "segmif_conv3x3_out1_f32_fwd": [P, c_int, P, P, P, P, c_int, c_int, c_int, c_int, P],
}
_RESTYPES = {"segmif_drdb_dataflow_workspace_bytes": c_size_t, "segmif_last_error": c_char_p, "segmif_loss_workspace_bytes": c_size_t, "segmif_wgrad_workspace_bytes": c_size_t,
             "segmif_dwconv3x3_gelu_bwd_workspace": c_int64}

_lib = None
_lock = threading.Lock()
_inited = set()
launch_count = 0          # kernels-launching entry points called so far (bench.py reports the delta)


def load():
    """Returns the loaded CDLL; raises if the library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    f"segmif_b200: {LIB_PATH} is missing. Build it with `python -m segmif_b200.build` "
                    "(needs nvcc; sm_100a only). There is no CPU or PyTorch fallback.")
            lib = ctypes.CDLL(LIB_PATH)
            for name, argtypes in SIGNATURES.items():
                fn = getattr(lib, name)          # AttributeError here == header/library mismatch
                fn.argtypes = argtypes
                fn.restype = _RESTYPES.get(name, c_int)
            if lib.segmif_abi_version() != 1:
                raise RuntimeError("segmif_b200: ABI version mismatch between _lib.py and the shared library")
            _lib = lib
    return _lib


def last_error():
    return load().segmif_last_error().decode()


def check(rc, what=""):
    if rc != 0:
        raise RuntimeError(f"segmif_b200 {what} failed (code {rc}): {last_error()}")


def ensure_init(device_index):
    if device_index not in _inited:
        check(load().segmif_init(int(device_index)), "segmif_init")
        _inited.add(device_index)


def call(name, *args):
    global launch_count
    launch_count += 1
    check(getattr(load(), name)(*args), name)
