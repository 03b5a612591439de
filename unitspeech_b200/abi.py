"""ctypes binding of libunitspeech_b200.so (declared in include/unitspeech_b200.h and include/unitspeech_b200_train.h).

There is no fallback: if the shared library is missing or fails to load, importing the decoder raises.
"""

from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int32, c_int64, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libunitspeech_b200.so")


class UsbConfig(ctypes.Structure):
    _fields_ = [
        ("n_feats", c_int32),
        ("dim", c_int32),
        ("n_mults", c_int32),
        ("dim_mults", c_int32 * 8),
        ("groups", c_int32),
        ("spk_emb_dim", c_int32),
        ("pe_scale", c_float),
        ("beta_min", c_float),
        ("beta_max", c_float),
        ("device", c_int32),
    ]


class UsbVocoderConfig(ctypes.Structure):
    _fields_ = [
        ("num_mels", c_int32),
        ("n_upsamples", c_int32),
        ("upsample_rates", c_int32 * 8),
        ("upsample_kernel_sizes", c_int32 * 8),
        ("upsample_initial_channel", c_int32),
        ("resblock_type", c_int32),
        ("n_resblock_kernels", c_int32),
        ("resblock_kernel_sizes", c_int32 * 4),
        ("n_dilations", c_int32),
        ("resblock_dilations", (c_int32 * 4) * 4),
        ("activation", c_int32),
        ("snake_logscale", c_int32),
        ("device", c_int32),
    ]


class UsbError(RuntimeError):
    pass


# name -> (restype, argtypes); the single source of truth checked against include/unitspeech_b200.h by the tests
SIGNATURES = {
    "usb_last_error": (c_char_p, []),
    "usb_version": (c_int32, []),
    "usb_create": (c_int32, [POINTER(UsbConfig), POINTER(c_void_p)]),
    "usb_destroy": (None, [c_void_p]),
    "usb_load_param": (c_int32, [c_void_p, c_char_p, c_void_p, POINTER(c_int64), c_int32]),
    "usb_finalize_params": (c_int32, [c_void_p]),
    "usb_estimator_forward": (c_int32, [c_void_p] + [c_void_p] * 6 + [c_int32, c_int32, c_uint64]),
    "usb_forward_diffusion": (c_int32, [c_void_p] + [c_void_p] * 6 + [c_int32, c_int32, c_uint64]),
    "usb_loss_t": (c_int32, [c_void_p] + [c_void_p] * 8 + [c_int32, c_int32, c_uint64]),
    "usb_reverse_diffusion": (c_int32, [c_void_p] + [c_void_p] * 7 + [c_int32, c_float, c_float, c_void_p, c_void_p,
                                                                    c_int32, c_int32, c_uint64]),
    "usb_reverse_diffusion_host": (c_int32, [c_void_p] + [c_void_p] * 7 + [c_int32, c_float, c_float, c_void_p,
                                                                         c_int32, c_int32, c_uint64]),
    "usb_align_expand": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p,
                                   c_void_p, c_void_p, c_void_p, c_uint64]),
    "usb_set_output_denorm": (c_int32, [c_void_p, c_void_p, c_void_p]),
    "usb_saturation_count": (c_int32, [c_void_p, POINTER(c_int64), c_int32]),
    "usb_set_graph_mode": (c_int32, [c_void_p, c_int32]),
    "usb_graph_steps": (c_int64, [c_void_p]),
    "usb_set_splitk_mode": (c_int32, [c_void_p, c_int32]),
    "usb_workspace_bytes": (c_int64, [c_void_p]),
    "usb_launch_count": (c_int64, [c_void_p]),
    "usb_set_profiling": (c_int32, [c_void_p, c_int32]),
    "usb_get_profile": (c_int32, [c_void_p, POINTER(c_double), POINTER(c_double), POINTER(c_int64)]),
    "usb_op_conv": (c_int32, [c_void_p, c_int32, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32,
                              c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_int32, c_void_p,
                              c_uint64]),
    "usb_dbg_conv_time": (c_int32, [c_void_p] + [c_int32] * 9 + [POINTER(c_float)]),
    "usb_op_gn_apply": (c_int32, [c_void_p] + [c_void_p] * 8 + [c_int32] * 5 + [c_uint64]),
    "usb_op_attn_context": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32,
                                      c_uint64]),
    "usb_vocoder_create": (c_int32, [POINTER(UsbVocoderConfig), POINTER(c_void_p)]),
    "usb_vocoder_destroy": (None, [c_void_p]),
    "usb_vocoder_load_param": (c_int32, [c_void_p, c_char_p, c_void_p, POINTER(c_int64), c_int32]),
    "usb_vocoder_finalize_params": (c_int32, [c_void_p]),
    "usb_vocoder_forward": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_uint64]),
    "usb_vocoder_forward_host": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_void_p]),
    "usb_vocoder_set_input_denorm": (c_int32, [c_void_p, c_void_p, c_void_p]),
    "usb_vocoder_launch_count": (c_int64, [c_void_p]),
    "usb_vocoder_workspace_bytes": (ctypes.c_size_t, [c_void_p]),
    "usb_vocoder_flops_per_call": (c_double, [c_void_p]),
    "usb_vocoder_set_profiling": (c_int32, [c_void_p, c_int32]),
    "usb_vocoder_get_profile": (c_int32, [c_void_p, POINTER(c_double), POINTER(c_double), POINTER(c_int64)]),
    "usb_op_snake_act": (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_uint64]),
    "usb_op_conv1d": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32,
                                c_int32, c_int32, c_uint64]),
    "usb_vocoder_filter": (c_int32, [POINTER(c_float)]),
    # ---- fine-tune step, operator level (include/unitspeech_b200_train.h)
    "usb_t_pack_conv": (c_int32, [c_void_p, c_int32, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_uint64]),
    "usb_t_pack_begin": (c_int32, [c_void_p, c_uint64]),
    "usb_t_pack_flush": (c_int32, [c_void_p, c_uint64]),
    "usb_t_cast": (c_int32, [c_void_p, c_void_p, c_void_p, c_int64, c_uint64]),
    "usb_t_conv": (c_int32, [c_void_p, c_int32, c_void_p, c_int32, c_int32, c_void_p, c_int32, c_int32, c_int32, c_int32,
                             c_int32, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                             c_int32, c_void_p, c_uint64]),
    "usb_t_first_conv": (c_int32, [c_void_p] + [c_void_p] * 11 + [c_int32] * 4 + [c_uint64]),
    "usb_t_gn_apply": (c_int32, [c_void_p] + [c_void_p] * 5 + [c_int64] + [c_void_p] * 3 + [c_int32] * 4 + [c_uint64]),
    "usb_t_embed": (c_int32, [c_void_p] + [c_void_p] * 11 + [c_int32, c_int32, c_uint64]),
    "usb_t_attn_scratch_bytes": (c_int64, [c_int32, c_int32, c_int32]),
    "usb_t_attn_context": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_int32, c_int32, c_int32, c_int32, c_uint64]),
    "usb_t_final": (c_int32, [c_void_p] + [c_void_p] * 8 + [c_int32] * 4 + [c_uint64]),
    "usb_t_loss": (c_int32, [c_void_p] + [c_void_p] * 6 + [c_int32, c_int32, c_uint64]),
    "usb_t_loss_grad": (c_int32, [c_void_p] + [c_void_p] * 4 + [c_float] + [c_void_p] * 2 + [c_int32, c_int32, c_uint64]),
    "usb_t_gn_bwd": (c_int32, [c_void_p] + [c_void_p] * 15 + [c_int64, c_void_p] + [c_int32] * 4 + [c_uint64]),
    "usb_t_colsum": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_int64, c_uint64]),
    "usb_t_add": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_uint64]),
    "usb_t_wgrad": (c_int32, [c_void_p, c_int32, c_void_p, c_int32, c_void_p, c_int32] + [c_int32] * 7 + [c_void_p, c_int32,
                                                                                                        c_uint64]),
    "usb_t_first_conv_wgrad": (c_int32, [c_void_p] + [c_void_p] * 8 + [c_int32] * 4 + [c_uint64]),
    "usb_t_attn_bwd_small": (c_int32, [c_void_p] + [c_void_p] * 11 + [c_int32] * 3 + [c_uint64]),
    "usb_t_attn_bwd_dkv": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_int32, c_int32, c_int32, c_uint64]),
    "usb_t_embed_bwd": (c_int32, [c_void_p] + [c_void_p] * 17 + [c_int32, c_int32, c_uint64]),
    "usb_t_dot": (c_int32, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_uint64]),
    "usb_t_sumsq": (c_int32, [c_void_p, c_void_p, c_int64, c_void_p, c_uint64]),
    "usb_t_adam": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_float, c_float, c_float, c_float,
                             c_int32, c_void_p, c_void_p, c_float, c_float, c_void_p, c_uint64]),
}

_lib = None


def load_library() -> ctypes.CDLL:
    """Loads the CUDA library; raises if it has not been built (python -m unitspeech_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise UsbError(f"{LIB_PATH} not found: build it with `python -m unitspeech_b200.build` "
                       "(there is no CPU or PyTorch fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load_library().usb_last_error()
        raise UsbError(msg.decode() if msg else f"libunitspeech_b200 call failed ({rc})")
