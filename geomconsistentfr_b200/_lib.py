"""ctypes binding of libgfr_b200.so (the C ABI declared in include/gfr_b200.h).

There is no fallback: if the library cannot be loaded every op raises."""
import ctypes
import os

from . import build as _build

_c_void_p, _c_int, _c_float = ctypes.c_void_p, ctypes.c_int, ctypes.c_float

# name -> argtypes (restype is always int unless listed in _RESTYPES)
_PROTOS = {
    "gfr_version": [],
    "gfr_error_string": [_c_int],
    "gfr_mask_pack": [_c_void_p, _c_int, _c_int, _c_int, _c_int, _c_void_p, _c_void_p],
    "gfr_shadow_march_fwd": [_c_void_p, _c_void_p, _c_int, _c_void_p, _c_void_p, _c_int, _c_float, _c_void_p,
                             _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_void_p],
    "gfr_march_config": [_c_int, _c_int, _c_int, _c_int],
    "gfr_conv_p16_config": [_c_int],
    "gfr_wgrad_tc_config": [_c_int],
    "gfr_shade_render_fwd": [_c_void_p] * 11 + [_c_int, _c_int, _c_int, _c_int, _c_void_p],
    "gfr_march_shade_fwd": [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_float] + [_c_void_p] * 10
                           + [_c_int] * 4 + [_c_void_p],
    "gfr_shadow_march_bwd": [_c_void_p] * 5 + [_c_int, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_void_p],
    "gfr_shade_render_bwd": [_c_void_p] * 16 + [_c_int, _c_int, _c_int, _c_void_p],
    "gfr_conv2d_fwd": [_c_void_p] * 7 + [_c_int] * 9 + [_c_float, _c_void_p],
    "gfr_maxpool2_fwd": [_c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_void_p],
    "gfr_upsample2_fwd": [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_void_p],
    "gfr_light_head_fwd": [_c_void_p, ctypes.c_longlong, _c_int, _c_int] + [_c_void_p] * 5 + [_c_int, _c_void_p],
    "gfr_ssim_fwd": [_c_void_p] * 4 + [_c_int, _c_int, _c_int, _c_float, _c_void_p],
    "gfr_ssim_bwd": [_c_void_p] * 5 + [_c_int, _c_int, _c_int, _c_void_p],
    "gfr_masked_losses": [_c_void_p] * 12 + [_c_int, _c_int, _c_int, _c_void_p],
    "gfr_adam_step": [_c_void_p] * 4 + [ctypes.c_longlong, _c_void_p, _c_float, _c_float, _c_float, _c_float, _c_float, _c_void_p],
    "gfr_adam_step_segments": [_c_void_p] * 4 + [ctypes.c_longlong, _c_void_p, _c_void_p, _c_int, _c_float, _c_float, _c_float, _c_float,
                               _c_float, _c_void_p],
    # train-mode CNN
    "gfr_conv_tc_pack_weights_dev": [_c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_void_p, _c_void_p],
    "gfr_conv_tc_pack_size_ex": [_c_int] * 5,
    "gfr_conv_tc_pack_weights_dev_ex": [_c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_void_p, _c_void_p],
    "gfr_conv_tc_pack_job_size": [],
    "gfr_conv_tc_pack_job_fill": [_c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_void_p, ctypes.c_longlong],
    "gfr_conv_tc_pack_weights_batch": [_c_void_p, _c_int, ctypes.c_longlong, _c_void_p],
    "gfr_conv_tc_fwd_ex": [_c_void_p] * 6 + [_c_int] * 13 + [_c_float, _c_int, _c_int, _c_void_p],
    "gfr_bn_train_stats": [_c_void_p] * 10 + [_c_int] * 4 + [_c_float, _c_float, _c_void_p],
    "gfr_bn_train_stats_ex": [_c_void_p] * 11 + [_c_int] * 4 + [_c_float, _c_float, _c_void_p],
    "gfr_bn_config": [_c_int],
    "gfr_bn_running_update": [_c_void_p] * 4 + [_c_int] * 4 + [_c_float, _c_void_p],
    "gfr_bn_apply_fwd": [_c_void_p] * 6 + [_c_int] * 6 + [_c_void_p],
    "gfr_bn_apply_bwd": [_c_void_p] * 11 + [_c_int] * 5 + [_c_void_p],
    "gfr_bn_apply_bwd_ex": [_c_void_p] * 14 + [_c_int] * 5 + [_c_void_p],
    "gfr_conv_wgrad_tc_bf16": [_c_void_p] * 3 + [_c_int] * 10 + [_c_void_p],
    "gfr_channel_sum_c4": [_c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_void_p],
    "gfr_conv3x3_wgrad": [_c_void_p] * 4 + [_c_int] * 7 + [_c_void_p],
    "gfr_maxpool2_c4_bwd": [_c_void_p] * 3 + [_c_int] * 3 + [_c_void_p],
    "gfr_sumpool2_c4": [_c_void_p] * 2 + [_c_int] * 3 + [_c_void_p],
    "gfr_avgpool_c4_fwd": [_c_void_p] * 2 + [_c_int] * 5 + [_c_void_p],
    "gfr_avgpool_c4_bwd": [_c_void_p] * 2 + [_c_int] * 5 + [_c_void_p],
    "gfr_pw_conv16_fwd": [_c_void_p] * 4 + [_c_int] * 6 + [_c_float, _c_void_p],
    "gfr_pw_conv16_bwd": [_c_void_p] * 7 + [_c_int] * 6 + [_c_float, _c_void_p],
    "gfr_stem_conv_train_fwd": [_c_void_p] * 4 + [_c_int] * 3 + [_c_void_p],
    "gfr_stem_conv_wgrad": [_c_void_p] * 4 + [_c_int] * 3 + [_c_void_p],
    # PatchGAN support
    "gfr_space_to_depth": [_c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_void_p],
    "gfr_depth_to_space": [_c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_void_p],
    "gfr_space_to_depth_pad": [_c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_void_p],
    "gfr_depth_to_space_pad": [_c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_void_p],
    "gfr_conv2x2_wgrad": [_c_void_p] * 4 + [_c_int] * 6 + [_c_void_p],
    "gfr_lrelu_bwd_c4": [_c_void_p, _c_void_p, _c_void_p, ctypes.c_longlong, _c_void_p],
    "gfr_conv4x4s1_to1_fwd": [_c_void_p] * 4 + [_c_int] * 4 + [_c_void_p],
    "gfr_conv4x4s1_to1_bwd": [_c_void_p] * 6 + [_c_int] * 4 + [_c_void_p],
    # tensor-core path (C4 layout)
    "gfr_nchw_to_c4": [_c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_void_p],
    "gfr_c4_to_nchw": [_c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_void_p],
    "gfr_conv_tc_pack_size": [_c_int, _c_int, _c_int],
    "gfr_conv_tc_pack_weights": [_c_void_p, _c_int, _c_int, _c_int, _c_void_p],
    "gfr_conv3x3_tc_fwd": [_c_void_p] * 6 + [_c_int] * 9 + [_c_float, _c_int, _c_int, _c_float, _c_float, _c_void_p],
    "gfr_conv3x3_tc_head_fwd": [_c_void_p] * 5 + [_c_int] * 7 + [_c_float, _c_int, _c_int, _c_float, _c_float, _c_void_p],
    "gfr_conv_tc_pack_size_f16": [_c_int, _c_int, _c_int],
    "gfr_conv_tc_pack_weights_f16": [_c_void_p, _c_int, _c_int, _c_int, _c_float, _c_void_p],
    "gfr_stem_conv_fwd": [_c_void_p] * 5 + [_c_int] * 3 + [_c_void_p],
    "gfr_head_1x1_fwd": [_c_void_p] * 8 + [_c_int] * 5 + [_c_float, _c_void_p],
    "gfr_light_head_c4_fwd": [_c_void_p, _c_int, _c_int, _c_int] + [_c_void_p] * 5 + [_c_int, _c_void_p],
    "gfr_maxpool2_c4_fwd": [_c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_void_p],
    "gfr_upsample2_c4_fwd": [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_void_p],
    # eval-mode CNN on pre-split fp16-pair activations (P16)
    "gfr_conv_p16_pack_size": [_c_int, _c_int, _c_int, _c_int],
    "gfr_conv_p16_pack_weights": [_c_void_p, _c_int, _c_int, _c_int, _c_int, _c_float, _c_void_p],
    "gfr_conv3x3_p16_fwd": [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_void_p, _c_int, _c_void_p, _c_int, _c_void_p]
                           + [_c_int] * 12 + [_c_float, _c_float, _c_int, _c_void_p],
    "gfr_conv_p16_fwd_ex": [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_void_p, _c_int, _c_void_p, _c_int, _c_void_p, _c_void_p]
                           + [_c_int] * 13 + [_c_float, _c_float, _c_int, _c_void_p],
    "gfr_conv_p16_pack_size_taps": [_c_int] * 5,
    "gfr_conv_p16_pack_weights_taps": [_c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_float, _c_void_p],
    "gfr_stem_unroll_p16": [_c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_void_p],
    "gfr_nchw_to_p16": [_c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_void_p],
    "gfr_p16_to_nchw": [_c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_void_p],
    "gfr_maxpool2_p16_fwd": [_c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_void_p],
    "gfr_stem_conv_p16_fwd": [_c_void_p] * 5 + [_c_int] * 3 + [_c_void_p],
    "gfr_head_1x1_p16_fwd": [_c_void_p] * 8 + [_c_int] * 5 + [_c_float, _c_void_p],
    "gfr_conv3x3_p16_head_fwd": [_c_void_p] * 3 + [_c_int] * 6 + [_c_float, _c_int] + [_c_void_p] * 7 + [_c_int, _c_int, _c_float, _c_void_p],
    "gfr_light_head_p16_fwd": [_c_void_p, _c_int, _c_int, _c_int] + [_c_void_p] * 5 + [_c_int, _c_void_p],
    # output stage of the inference drivers
    "gfr_composite_bgr_u8": [_c_void_p, _c_int, _c_void_p, _c_void_p, _c_int, _c_void_p, _c_int, _c_int, _c_int, _c_void_p],
    "gfr_neg_depth_range": [_c_void_p, ctypes.c_longlong, _c_void_p, _c_void_p],
    "gfr_export_planes_u8": [_c_void_p] * 6 + [_c_int] + [_c_void_p] * 6 + [_c_int, _c_int, _c_int, _c_void_p],
    "gfr_masked_mse_u8": [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_void_p],
    "gfr_masked_ssim_u8": [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_void_p],
    "gfr_lpips_layer_fwd": [_c_void_p] * 4 + [_c_int] * 3 + [_c_void_p],
    "gfr_lpips_layer_bwd": [_c_void_p] * 6 + [_c_int] * 3 + [_c_void_p],
    "gfr_bilinear_up_add": [_c_void_p, _c_void_p] + [_c_int] * 5 + [_c_void_p],
    "gfr_bilinear_up_add_bwd": [_c_void_p, _c_void_p] + [_c_int] * 5 + [_c_void_p],
    "gfr_lpips_masked_sums": [_c_void_p, _c_void_p, _c_int, _c_void_p, _c_int, _c_int, _c_int, _c_void_p],
    "gfr_border_median_fix_u8": [_c_void_p, _c_void_p, _c_int, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_void_p],
}
_RESTYPES = {"gfr_error_string": ctypes.c_char_p, "gfr_conv_tc_pack_size": ctypes.c_longlong,
             "gfr_conv_tc_pack_size_f16": ctypes.c_longlong, "gfr_conv_p16_pack_size": ctypes.c_longlong,
             "gfr_conv_tc_pack_size_ex": ctypes.c_longlong, "gfr_conv_tc_pack_job_fill": ctypes.c_longlong, "gfr_conv_p16_pack_size_taps": ctypes.c_longlong}

_lib = None


def lib_path():
    return _build.LIB


def load():
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    if not os.path.isfile(path):
        try:
            _build.build()
        except Exception as e:  # no nvcc on this machine
            raise RuntimeError("libgfr_b200.so is missing and could not be built (%s). "
                               "Run `python -m geomconsistentfr_b200.build`." % e)
    lib = ctypes.CDLL(path)
    for name, argtypes in _PROTOS.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, _c_int)
    _lib = lib
    return lib


def exported_symbols():
    return sorted(_PROTOS)


def check(code, what):
    if code != 0:
        msg = load().gfr_error_string(code)
        raise RuntimeError("%s failed: %s (code %d)" % (what, msg.decode() if msg else "?", code))
