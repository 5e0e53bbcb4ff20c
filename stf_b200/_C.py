"""ctypes binding of libstf_b200.so (include/stf_b200.h) -- the only way Python reaches the kernels.

There is no fallback: if the library is missing or fails to load, importing any compute entry
point raises.  Device pointers are passed as integers (tensor.data_ptr()), the CUDA stream as
torch.cuda.current_stream().cuda_stream.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libstf_b200.so")

c_int, c_i64, c_f32, c_vp = ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_void_p
_i32p = ctypes.POINTER(ctypes.c_int32)
_u8p = ctypes.POINTER(ctypes.c_uint8)
_f32p = ctypes.POINTER(ctypes.c_float)

STF_EB_PARAMS = 60
STF_EB_MEDIAN_SLOT = 58
ROWS_DENSE, ROWS_WINDOW, ROWS_MERGE = 0, 1, 2
EPI_STORE, EPI_QKV, EPI_GELU, EPI_RESIDUAL, EPI_WINDOW_RESIDUAL, EPI_PIXEL_SHUFFLE = range(6)
PREC_TF32, PREC_FP32 = 0, 1

_ERRORS = {-1: "invalid argument", -2: "unsupported shape", -3: "pointer not 16-byte aligned",
           -4: "invalid table", -5: "output buffer too small", -6: "corrupt bitstream"}


class LinearArgs(ctypes.Structure):
    """struct stf_linear_args (include/stf_b200.h)."""
    _fields_ = [
        ("M", c_int), ("N", c_int), ("K", c_int),
        ("x", c_vp), ("ldx", c_int),
        ("w_packed", c_vp),
        ("y", c_vp), ("ldy", c_int),
        ("rows", c_int), ("has_ln", c_int), ("x_is_tf32", c_int), ("ln_eps", c_f32),
        ("epilogue", c_int), ("residual", c_vp), ("q_cols", c_int), ("q_scale", c_f32),
        ("batch", c_int), ("H", c_int), ("W", c_int), ("window", c_int), ("shift", c_int), ("precision", c_int),
        ("max_ctas", c_int),
    ]


class ConvArgs(ctypes.Structure):
    """struct stf_conv_args (include/stf_b200.h)."""
    _fields_ = [
        ("batch", c_int), ("H", c_int), ("W", c_int), ("n_src", c_int),
        ("src", c_vp * 3), ("src_channels", c_int * 3), ("src_ld", c_int * 3),
        ("N", c_int), ("ksize", c_int), ("stride", c_int),
        ("w_packed", c_vp), ("y", c_vp), ("ldy", c_int), ("act", c_int), ("residual", c_vp), ("res_ld", c_int),
        ("pixel_shuffle", c_int), ("has_ln", c_int), ("ln_eps", c_f32), ("precision", c_int), ("max_ctas", c_int),
    ]


class SliceArgs(ctypes.Structure):
    """struct stf_slice_args (include/stf_b200.h)."""
    _fields_ = [
        ("y", c_vp), ("y_ld", c_int), ("scales", c_vp), ("scales_ld", c_int), ("means", c_vp), ("means_ld", c_int),
        ("symbols_in", c_vp), ("symbols_in_batch_stride", c_i64),
        ("symbols_out", c_vp), ("indexes_out", c_vp), ("out_batch_stride", c_i64),
        ("y_hat", c_vp), ("y_hat_ld", c_int), ("likelihood", c_vp), ("likelihood_batch_stride", c_i64),
        ("batch", c_int), ("channels", c_int), ("plane", c_i64),
        ("table_host", _f32p), ("levels", c_int), ("scale_bound", c_f32), ("lik_bound", c_f32), ("ste_round", c_int),
        ("narrow", c_int), ("overflow", c_vp),
    ]


class MlpArgs(ctypes.Structure):
    """struct stf_mlp_args (include/stf_b200.h)."""
    _fields_ = [
        ("x", c_vp), ("x_ld", c_int), ("y", c_vp), ("y_ld", c_int), ("M", c_i64), ("C", c_int), ("hidden", c_int),
        ("w1_packed", c_vp), ("w2_packed", c_vp), ("ln_eps", c_f32), ("precision", c_int), ("max_ctas", c_int),
    ]


# name -> (restype, argtypes); mirrors include/stf_b200.h one to one (tests check the export list)
SIGNATURES = {
    "stf_version": (ctypes.c_char_p, []),
    "stf_launch_count": (c_i64, []),
    "stf_build_indexes": (c_int, [c_vp, c_vp, c_i64, _f32p, c_int, c_f32, c_vp]),
    "stf_gaussian_compress_step": (c_int, [c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_int, c_int, c_i64,
                                           _f32p, c_int, c_f32, c_vp]),
    "stf_slice_step_nhwc": (c_int, [ctypes.POINTER(SliceArgs), c_vp]),
    "stf_quantize_symbols": (c_int, [c_vp, c_vp, c_vp, c_i64, c_vp]),
    "stf_quantize_dequantize": (c_int, [c_vp, c_vp, c_vp, c_i64, c_vp]),
    "stf_dequantize": (c_int, [c_vp, c_i64, c_vp, c_vp, c_int, c_int, c_i64, c_vp]),
    "stf_gaussian_likelihood": (c_int, [c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_i64, c_f32, c_f32, c_int, c_vp]),
    "stf_entropy_bottleneck": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_i64, c_f32, c_int, c_vp]),
    "stf_linear_n_tile": (c_int, [c_int]),
    "stf_linear_n_tile_prec": (c_int, [c_int, c_int]),
    "stf_packed_linear_floats": (c_i64, [c_int, c_int, c_int]),
    "stf_pack_linear": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_vp]),
    "stf_linear": (c_int, [ctypes.POINTER(LinearArgs), c_vp]),
    "stf_window_attention": (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_i64, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_vp]),
    "stf_swin_mlp": (c_int, [ctypes.POINTER(MlpArgs), c_vp]),
    "stf_window_attention_tokens": (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_vp]),
    "stf_packed_conv_floats": (c_i64, [ctypes.POINTER(ConvArgs)]),
    "stf_pack_conv": (c_int, [ctypes.POINTER(ConvArgs), c_vp, c_vp, c_vp, c_vp, c_int, c_f32, c_vp, c_vp]),
    "stf_conv2d": (c_int, [ctypes.POINTER(ConvArgs), c_vp]),
    "stf_conv2d_out_hw": (c_int, [c_int, c_int, c_int, c_int, ctypes.POINTER(c_int), ctypes.POINTER(c_int)]),
    "stf_patch_embed": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_f32, c_vp]),
    "stf_bias_act": (c_int, [c_vp, c_vp, c_int, c_i64, c_int, c_vp]),
    "stf_layernorm_fwd": (c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, c_int, c_f32, c_vp]),
    "stf_attention_bwd_ctas": (c_int, [c_i64, c_int, c_int, ctypes.POINTER(c_int)]),
    "stf_attention_bwd_slots": (c_int, [c_i64, c_int, c_int, c_int]),
    "stf_window_attention_bwd": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_int, c_int, c_int, c_int, c_int, c_int, c_f32, c_vp]),
    "stf_layernorm_bwd_ctas": (c_int, [c_i64]),
    "stf_layernorm_bwd": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_int, c_f32, c_vp]),
    "stf_colsum_ctas": (c_int, [c_i64]),
    "stf_colsum": (c_int, [c_vp, c_vp, c_i64, c_int, c_vp]),
    "stf_gelu_bwd": (c_int, [c_vp, c_vp, c_vp, c_i64, c_vp]),
    "stf_gaussian_likelihood_train": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_f32, c_f32, c_vp]),
    "stf_gaussian_likelihood_train_bwd": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_f32, c_f32, c_vp]),
    "stf_rans_table_create": (c_vp, [_i32p, c_int, c_int, _i32p, _i32p]),
    "stf_rans_table_destroy": (None, [c_vp]),
    "stf_rans_encode_bound": (c_i64, [c_i64]),
    "stf_rans_encode": (c_i64, [c_vp, c_vp, c_vp, c_i64, c_vp, c_i64]),
    "stf_rans_encode_batch": (c_int, [c_vp, c_int, ctypes.POINTER(c_vp), ctypes.POINTER(c_vp), ctypes.POINTER(c_i64),
                                      ctypes.POINTER(c_vp), ctypes.POINTER(c_i64), ctypes.POINTER(c_i64), c_int]),
    "stf_rans_encode_batch_narrow": (c_int, [c_vp, c_int, ctypes.POINTER(c_vp), ctypes.POINTER(c_vp), ctypes.POINTER(c_i64),
                                      ctypes.POINTER(c_vp), ctypes.POINTER(c_i64), ctypes.POINTER(c_i64), c_int]),
    "stf_rans_decoder_create": (c_vp, [c_vp, c_i64]),
    "stf_rans_decoder_create_view": (c_vp, [c_vp, c_i64]),
    "stf_rans_decoder_destroy": (None, [c_vp]),
    "stf_rans_decode": (c_int, [c_vp, c_vp, c_vp, c_i64, c_vp]),
    "stf_rans_decode_batch": (c_int, [ctypes.POINTER(c_vp), c_vp, c_int, ctypes.POINTER(c_vp), ctypes.POINTER(c_i64),
                                      ctypes.POINTER(c_vp), c_int]),
    "stf_rans_decode_batch_u8": (c_int, [ctypes.POINTER(c_vp), c_vp, c_int, ctypes.POINTER(c_vp), ctypes.POINTER(c_i64),
                                      ctypes.POINTER(c_vp), c_int]),
    "stf_rans_device_table_bytes": (c_i64, [c_vp]),
    "stf_rans_device_table_pack": (c_int, [c_vp, c_vp]),
    "stf_rans_decode_device": (c_int, [c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_vp, c_i64, c_vp, c_i64, c_int,
                                       c_i64, c_vp]),
    "stf_pmf_to_quantized_cdf": (c_int, [_f32p, c_int, c_int, ctypes.POINTER(ctypes.c_uint32)]),
}

_lib = None


def lib():
    """Load (once) and return the shared library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -m stf_b200.build` (nvcc, sm_100a). "
                "stf_b200 has no CPU or PyTorch fallback.")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(rc, what):
    """Map a C-ABI status to the Python exceptions the reference raises for the same misuse."""
    if rc == 0:
        return
    if rc < 0:
        raise ValueError(f"{what}: {_ERRORS.get(rc, rc)}")
    raise RuntimeError(f"{what}: CUDA error {rc}")


def ptr(t):
    """Device / host address of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream():
    import torch
    return torch.cuda.current_stream().cuda_stream
