"""ctypes binding of libaedit.so (C ABI: include/aedit.h).  No fallback: if the shared library is missing or
fails to load, importing any device op raises.  The library is built in-tree by audioeditingcode_b200/build.py
(called from __graft_entry__.build())."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# AEDIT_OPERANDS selects the build: fp16 operands (default, libaedit.so) or bf16 operands (libaedit_bf16.so) — see
# include/aedit.h, ae_operand_dtype().  One process uses one of them.
OPERANDS = os.environ.get("AEDIT_OPERANDS", "fp16").lower()
if OPERANDS not in ("fp16", "bf16"):
    raise ValueError(f"AEDIT_OPERANDS must be fp16 or bf16, not {OPERANDS!r}")
LIB_PATH = os.path.join(_HERE, "libaedit.so" if OPERANDS == "fp16" else "libaedit_bf16.so")

vp = C.c_void_p
i32 = C.c_int32
i64 = C.c_int64
f32 = C.c_float


class AeSchedRow(C.Structure):
    _fields_ = [("t", i32), ("prev_t", i32), ("alpha_bar_t", f32), ("alpha_prod_t_prev", f32), ("variance", f32),
                ("sqrt_ab", f32), ("sqrt_1mab", f32), ("sqrt_ap", f32), ("sqrt_var", f32)]


class AeGemmArgs(C.Structure):
    _fields_ = [("A", vp), ("lda", i64), ("W", vp), ("ldw", i64), ("M", i32), ("N", i32), ("K", i32), ("batch", i32),
                ("strideA", i64), ("strideW", i64), ("stride_out", i64), ("stride_res", i64), ("bias", vp),
                ("rowbias", vp), ("ld_rowbias", i64), ("rows_per_group", i32), ("residual", vp), ("ld_res", i64),
                ("out_f32", vp), ("ld_out_f32", i64), ("out_bf16", vp), ("ld_out_bf16", i64), ("act", i32),
                ("alpha", f32), ("conv", i32), ("B", i32), ("H", i32), ("W_", i32), ("C", i32), ("kh", i32),
                ("kw", i32), ("dil_h", i32), ("dil_w", i32), ("force_bn", i32), ("splitk_ws", vp), ("splitk_ws_bytes", i64),
                ("force_split", i32), ("w_dynamic", i32), ("force_stages", i32),
                ("colstats", vp), ("cs_rows_per_sample", i32), ("force_persistent", i32),
                ("sm_L", i32), ("sm_block", i32), ("sm_rows", i32), ("sm_slot", vp),
                ("sm_bias", vp)]


_SIGS = {
    "ae_last_error": (C.c_char_p, []),
    "ae_version": (i32, []),
    "ae_operand_dtype": (i32, []),
    "ae_launch_count": (i64, []),
    "ae_device_ok": (i32, []),
    "ae_set_pdl": (None, [i32]),
    "ae_set_launch_priority": (None, [i32]),
    "ae_greatest_priority": (i32, []),
    "ae_set_shared_sm": (None, [i32]),
    "ae_set_skip_mask": (None, [i32]),
    "ae_set_tile_model_reduce": (None, [i32, i32]),
    "ae_set_persistent_min_tiles": (None, [i32]),
    "ae_set_shallow_kblocks": (None, [i32]),
    "ae_set_pdl_extra": (None, [i32]),
    "ae_set_gn_stream_min_bytes": (None, [i64]),
    "ae_set_splitk_ctas": (None, [i32]),
    "ae_set_fast_epilogue": (None, [i32]),
    "ae_set_tile_model": (None, [i32]),
    "ae_sched_create": (i32, [vp, i32, f32, vp, i32, i32, C.POINTER(vp)]),
    "ae_sched_create_from_rows": (i32, [C.POINTER(AeSchedRow), i32, i32, i32, C.POINTER(vp)]),
    "ae_sched_set_eta": (i32, [vp, vp, vp]),
    "ae_sched_destroy": (None, [vp]),
    "ae_sched_num_steps": (i32, [vp]),
    "ae_sched_row_h": (i32, [vp, i32, C.POINTER(AeSchedRow)]),
    "ae_sched_pos_of_t": (i32, [vp, i64]),
    "ae_sample_xts": (i32, [vp, vp, vp, vp, i64, vp]),
    "ae_cfg_inv_step": (i32, [vp, i32, i32, f32, vp, i64, vp, i64, i32, vp, vp, vp, vp, i32, i64, vp]),
    "ae_cfg_rev_step": (i32, [vp, i32, vp, f32, vp, vp, i32, vp, vp, vp, vp, vp, vp, vp, i64, vp]),
    "ae_ddim_step": (i32, [vp, i32, f32, f32, vp, vp, vp, vp, vp, vp, i64, vp]),
    "ae_pc_workspace_bytes": (i64, [i32, i64]),
    "ae_pc_perturb": (i32, [vp, i64, vp, f32, f32, i32, i32, vp, vp, i64, vp]),
    "ae_pc_subspace_step": (i32, [vp, vp, vp, vp, i32, i64, f32, vp, vp, vp, vp, vp, i64, vp]),
    "ae_pc_apply_drift": (i32, [vp, vp, vp, vp, f32, f32, f32, f32, i32, i32, i32, i64, vp, vp]),
    "ae_gemm": (i32, [C.POINTER(AeGemmArgs), vp]),
    "ae_gemm_conv_supported": (i32, [i32, i32, i32, i32]),
    "ae_im2col": (i32, [vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp, i64, vp]),
    "ae_groupnorm_workspace_bytes": (i64, [i32, i32]),
    "ae_groupnorm": (i32, [vp, i32, vp, i32, i32, i64, i32, f32, vp, vp, i32, vp, vp, vp, vp, vp]),
    "ae_groupnorm_cs": (i32, [vp, i32, vp, vp, i32, vp, i32, i64, i32, f32, vp, vp, i32, vp, vp, vp, vp, vp]),
    "ae_layernorm": (i32, [vp, i64, i32, f32, vp, vp, vp, vp]),
    "ae_geglu": (i32, [vp, i64, i32, vp, vp]),
    "ae_attention": (i32, [vp, i64, i64, vp, i64, i64, vp, i64, i64, vp, vp, i64, i32, i32, i32, i32, i32, f32, vp,
                           i64, i64, vp]),
    "ae_set_attention_tc": (None, [i32]),
    "ae_timestep_embedding": (i32, [vp, i32, i32, vp, vp]),
    "ae_upsample_nearest": (i32, [vp, i32, i32, i32, i32, i32, i32, vp, vp]),
    "ae_nchw_to_nhwc": (i32, [vp, i32, i32, i32, i32, vp, vp, vp]),
    "ae_nhwc_to_nchw": (i32, [vp, i32, i32, i32, i32, vp, vp]),
    "ae_cast_f32_bf16": (i32, [vp, i64, vp, i32, vp]),
    "ae_add_f32": (i32, [vp, vp, f32, i64, vp, vp]),
    "ae_softmax_rows": (i32, [vp, i64, i32, i64, vp, i64, vp]),
    "ae_transpose_bf16": (i32, [vp, i32, i32, i32, vp, vp]),
    "ae_stft_mel": (i32, [vp, i32, i32, i32, vp, vp, i32, i32, vp, vp, vp]),
    "ae_leaky_relu_bf16": (i32, [vp, i64, f32, f32, vp, vp]),
    "ae_tanh_f32": (i32, [vp, i64, vp, vp]),
    "ae_wave_to_int16": (i32, [vp, i64, vp, vp]),
}

EXPORTS = tuple(_SIGS)

_lib = None


class AeditError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load libaedit.so (once).  Raises AeditError if it has not been built — there is no CPU fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise AeditError(f"{LIB_PATH} not found: build it with `python -m audioeditingcode_b200.build` "
                         f"(or __graft_entry__.build()); the hot path has no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.ae_operand_dtype() != (1 if OPERANDS == "fp16" else 0):
        raise AeditError(f"{LIB_PATH} was built for the other operand type (rebuild: python -m audioeditingcode_b200.build)")
    _lib = lib
    return lib


def operand_torch_dtype():
    """torch dtype of the library's 16-bit tensor-core operands."""
    import torch
    return torch.float16 if load().ae_operand_dtype() == 1 else torch.bfloat16


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().ae_last_error()
        raise AeditError(f"{what or 'libaedit'} failed (rc={rc}): {msg.decode() if msg else ''}")
