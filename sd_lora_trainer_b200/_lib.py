"""ctypes binding of include/b200_lora.h.  Fails loudly: there is NO CPU / eager fallback behind these calls."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200_lora.so")

c_void_p, c_int32, c_int64, c_float, c_double = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_double


class Operand(C.Structure):
    _fields_ = [("ptr", c_void_p), ("rows", c_int64), ("inner", c_int64), ("row_stride", c_int64),
                ("sb0", c_int64), ("sb1", c_int64), ("mn_major", c_int32), ("batched", c_int32)]


class GemmDesc(C.Structure):
    _fields_ = [("M", c_int32), ("N", c_int32), ("num_seg", c_int32), ("K", c_int32 * 2),
                ("A", Operand * 2), ("B", Operand * 2), ("nb0", c_int32), ("nb1", c_int32),
                ("splits", c_int32), ("block_n", c_int32), ("pair_mode", c_int32),
                ("conv", c_int32), ("conv_N", c_int32), ("conv_H", c_int32), ("conv_W", c_int32), ("conv_C", c_int32),
                ("b_tap_k", c_int32), ("b_tap_n", c_int32),
                ("D", c_void_p), ("d_fp32", c_int32), ("d_atomic", c_int32),
                ("d_sm", c_int64), ("d_sn", c_int64), ("d_sb0", c_int64), ("d_sb1", c_int64),
                ("alpha", c_float), ("bias", c_void_p), ("bias_rows", c_int32), ("bias_sb", c_int64),
                ("R", c_void_p), ("r_sm", c_int64), ("r_sn", c_int64), ("r_sb0", c_int64), ("r_sb1", c_int64),
                ("side", c_int32), ("side_r", c_int32), ("S", Operand), ("B2", Operand), ("side_alpha", c_float),
                ("T_out", c_void_p), ("t_ld", c_int64),
                ("group", c_int32), ("D2", c_void_p), ("d2_sm", c_int64), ("d2_sn", c_int64), ("b_static", c_int32),
                ("geglu_h", c_void_p), ("geglu_h_ld", c_int64), ("geglu_y", c_void_p), ("geglu_y_ld", c_int64)]


class WgradProblem(C.Structure):
    _fields_ = [("X", c_void_p), ("Y", c_void_p), ("out", c_void_p), ("ld_x", c_int64), ("ld_y", c_int64),
                ("out_sn", c_int64), ("out_sj", c_int64), ("M", c_int32), ("Nout", c_int32), ("r", c_int32)]


# name -> argtypes (restype is int unless listed in _RESTYPES); must match include/b200_lora.h exactly.
SIGNATURES = {
    "b200_version": [],
    "b200_last_error": [],
    "b200_launch_count": [],
    "b200_gemm": [C.POINTER(GemmDesc), c_void_p],
    "b200_token_attention_loss": [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p, c_int64, c_int32, c_int32,
                                  c_void_p, c_void_p, c_int32, c_float, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_void_p],
    "b200_token_std_loss": [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_float, c_float,
                            c_float, c_float, c_float, c_void_p, c_void_p],
    "b200_lora_wgrad_batch": [C.POINTER(WgradProblem), c_int32, c_void_p],
    "b200_flash_attn_fwd": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int64,
                            c_int64, c_float, c_void_p],
    "b200_flash_attn_bwd": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                            c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int64, c_int64, c_int64, c_float, c_void_p,
                            c_int64, c_void_p, c_int64, c_int32, c_void_p],
    "b200_softmax_fwd": [c_void_p, c_void_p, c_int64, c_int32, c_int64, c_int64, c_void_p],
    "b200_softmax_bwd": [c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int64, c_int64, c_void_p],
    "b200_groupnorm_stats_floats": [c_int32, c_int64, c_int32, c_int32],
    "b200_token_attention_loss_floats": [c_int32, c_int32, c_int32, c_int32],
    "b200_groupnorm_fwd": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int64, c_int32, c_int32,
                           c_float, c_int32, c_void_p],
    "b200_groupnorm_bwd": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int64,
                           c_int32, c_int32, c_int32, c_void_p],
    "b200_layernorm_fwd": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_float, c_void_p],
    "b200_layernorm_bwd": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_void_p],
    "b200_geglu_fwd": [c_void_p, c_void_p, c_int64, c_int32, c_int32, c_void_p],
    "b200_geglu_bwd": [c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int32, c_void_p],
    "b200_silu_fwd": [c_void_p, c_void_p, c_int64, c_void_p],
    "b200_silu_bwd": [c_void_p, c_void_p, c_void_p, c_int64, c_void_p],
    "b200_norm_param_grad": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64,
                             c_int32, c_int32, c_int32, c_void_p],
    "b200_act_fwd": [c_void_p, c_void_p, c_int64, c_int32, c_void_p],
    "b200_act_bwd": [c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_void_p],
    "b200_add": [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p],
    "b200_head_pad": [c_void_p, c_void_p, c_int64, c_int32, c_int32, c_int32, c_int64, c_int64, c_void_p],
    "b200_upsample2x_fwd": [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p],
    "b200_upsample2x_bwd": [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p],
    "b200_im2col3x3": [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p],
    "b200_col2im3x3": [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p],
    "b200_shift_stack9": [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p],
    "b200_shift_sum9": [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_float, c_void_p],
    "b200_colsum": [c_void_p, c_void_p, c_void_p, c_int32, c_int64, c_int32, c_void_p],
    "b200_lora_transpose_b": [c_void_p, c_void_p, c_void_p, c_int32, c_void_p],
    "b200_lora_pack": [c_void_p, c_void_p, c_void_p, c_int32, c_void_p],
    "b200_bicubic_fwd": [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_int64, c_int64, c_void_p],
    "b200_bicubic_bwd": [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_int64, c_int64, c_void_p],
    "b200_timestep_embedding": [c_void_p, c_void_p, c_int32, c_int32, c_void_p],
    "b200_latent_sample": [c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_int64, c_void_p],
    "b200_noise_prologue": [c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_int32,
                            c_int32, c_int32, c_void_p],
    "b200_snr_weights": [c_void_p, c_void_p, c_float, c_void_p, c_int32, c_void_p],
    "b200_diffusion_loss": [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_int64,
                            c_int32, c_int32, c_int32, c_void_p],
    "b200_abs_sum": [c_void_p, c_int64, c_void_p, c_void_p],
    "b200_adamw": [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_double, c_double, c_double, c_double,
                   c_double, c_double, c_double, c_double, c_int32, c_double, c_int32, c_void_p],
    "b200_adamw_pack_hyper": [c_double, c_double, c_double, c_double, c_double, c_double, c_double, c_double, c_int32,
                              c_double, c_void_p],
    "b200_adamw_dev": [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_int32, c_void_p],
    "b200_prodigy_pack_hyper": [c_double, c_double, c_double, c_double, c_double, c_double, c_double, c_double, c_int32,
                                c_int32, c_double, c_double, c_void_p],
    "b200_prodigy_step": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int32,
                          c_void_p],
}
_RESTYPES = {"b200_last_error": C.c_char_p, "b200_launch_count": C.c_longlong, "b200_groupnorm_stats_floats": C.c_int64,
             "b200_token_attention_loss_floats": C.c_int64}

_lib = None


class B200Error(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load the CUDA library; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B200Error(f"{LIB_PATH} is missing - build it with `python -m sd_lora_trainer_b200.build` "
                        "(there is no CPU fallback for the B200 training step)")
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, C.c_int)
    if lib.b200_version() != 1:
        raise B200Error("libb200_lora.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().b200_last_error()
        raise B200Error(f"{what}: {msg.decode() if msg else 'error'} (code {rc})")


def launch_count() -> int:
    return int(load().b200_launch_count())
