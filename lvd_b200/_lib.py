"""ctypes binding of the C-ABI library (include/lvd_b200.h).  The library is mandatory: a missing or
unloadable liblvd_b200.so raises -- there is no Python/torch fallback for any op behind it."""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liblvd_b200.so")

LVD_MAX_TAPS = 27


class GemmDesc(ctypes.Structure):
    _fields_ = [
        ("a", ctypes.c_void_p),
        ("a_dims", ctypes.c_int64 * 5),
        ("a_strides", ctypes.c_int64 * 5),
        ("a_box", ctypes.c_int32 * 5),
        ("n_taps", ctypes.c_int32),
        ("k_per_tap", ctypes.c_int32),
        ("tap_off", (ctypes.c_int32 * 5) * LVD_MAX_TAPS),
        ("b", ctypes.c_void_p),
        ("b_mn_major", ctypes.c_int32),
        ("block_n", ctypes.c_int32),
        ("ldb", ctypes.c_int64),
        ("n", ctypes.c_int32),
        ("act", ctypes.c_int32),
        ("out_dims", ctypes.c_int32 * 4),
        ("rv_stride", ctypes.c_int32 * 4),
        ("out_row_stride", ctypes.c_int64 * 4),
        ("bias", ctypes.c_void_p),
        ("rowvec", ctypes.c_void_p),
        ("residual", ctypes.c_void_p),
        ("out", ctypes.c_void_p),
        ("ld_rowvec", ctypes.c_int64),
        ("ld_res", ctypes.c_int64),
        ("ld_out", ctypes.c_int64),
        ("gamma", ctypes.c_float),
        ("b_k_tap_stride", ctypes.c_int32),
        ("b_n_tap_stride", ctypes.c_int32),
        ("cta_pair", ctypes.c_int32),
        ("gn_partial", ctypes.c_void_p),
        ("gn_group_width", ctypes.c_int32),
        ("gn_groups", ctypes.c_int32),
        ("gn_tiles_per_sample", ctypes.c_int32),
        ("gn_subslots", ctypes.c_int32),
        ("ln_stats", ctypes.c_void_p),
        ("ln_colsum", ctypes.c_void_p),
        ("ln_bias", ctypes.c_void_p),
        ("ln_row_stride", ctypes.c_int32 * 4),
        ("lnstat_partial", ctypes.c_void_p),
        ("lnstat_rows", ctypes.c_int64),
        ("lnstat_row_stride", ctypes.c_int32 * 4),
        ("bias_f32", ctypes.c_void_p),
        ("weight_resident", ctypes.c_int32),
    ]


class AttnDesc(ctypes.Structure):
    _fields_ = [
        ("q", ctypes.c_void_p),
        ("kv", ctypes.c_void_p),
        ("out", ctypes.c_void_p),
        ("p_full", ctypes.c_void_p),
        ("p_tok", ctypes.c_void_p),
        ("lse", ctypes.c_void_p),
        ("ld_q", ctypes.c_int64),
        ("ld_kv", ctypes.c_int64),
        ("ld_o", ctypes.c_int64),
        ("q_width", ctypes.c_int64),
        ("kv_width", ctypes.c_int64),
        ("s_q", ctypes.c_int64),
        ("s_k", ctypes.c_int64),
        ("hw", ctypes.c_int64),
        ("frames", ctypes.c_int32),
        ("mode", ctypes.c_int32),
        ("q_col0", ctypes.c_int32),
        ("k_col0", ctypes.c_int32),
        ("v_col0", ctypes.c_int32),
        ("o_col0", ctypes.c_int32),
        ("heads", ctypes.c_int32),
        ("nz", ctypes.c_int32),
        ("kv_z_div", ctypes.c_int32),
        ("n_tok", ctypes.c_int32),
        ("tok_cols", ctypes.c_int32 * 16),
        ("scale", ctypes.c_float),
        ("causal", ctypes.c_int32),
        ("tok_cols_dev", ctypes.c_void_p),
    ]


class EnergyDesc(ctypes.Structure):
    _fields_ = [
        ("p", ctypes.c_void_p),
        ("slot_col", ctypes.c_void_p),
        ("boxes", ctypes.c_void_p),
        ("slot_weight", ctypes.c_void_p),
        ("loss", ctypes.c_void_p),
        ("dp", ctypes.c_void_p),
        ("layout", ctypes.c_int32),
        ("n_text", ctypes.c_int32),
        ("n_slots", ctypes.c_int32),
        ("frames", ctypes.c_int32),
        ("heads", ctypes.c_int32),
        ("H", ctypes.c_int32),
        ("W", ctypes.c_int32),
        ("fg_top_p", ctypes.c_float),
        ("bg_top_p", ctypes.c_float),
        ("fg_weight", ctypes.c_float),
        ("bg_weight", ctypes.c_float),
        ("com_loss_scale", ctypes.c_float),
        ("out_scale", ctypes.c_float),
        ("grad_scale", ctypes.c_float),
        ("loss_kind", ctypes.c_int32),
        ("k_min", ctypes.c_int32),
        ("boxdiff_normed", ctypes.c_int32),
        ("boxdiff_L", ctypes.c_int32),
        ("eps", ctypes.c_float),
        ("attn_sync_weight", ctypes.c_float),
        ("boxdiff_loss_scale", ctypes.c_float),
        ("p_frames", ctypes.c_int32),
        ("slot_frame0", ctypes.c_void_p),
        ("slot_loss", ctypes.c_void_p),
    ]


_lib = None


class LvdError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LvdError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). lvd_b200 has no fallback path."
            )
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.lvd_last_error.restype = ctypes.c_char_p
        _lib.lvd_launch_count.restype = ctypes.c_int64
        _declare(_lib)
    return _lib


def _declare(l):
    vp, i32, i64, f32 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_float
    l.lvd_gemm.argtypes = [ctypes.POINTER(GemmDesc), vp]
    l.lvd_attention.argtypes = [ctypes.POINTER(AttnDesc), vp]
    l.lvd_groupnorm.argtypes = [vp, i64, vp, i64, i32, i64, i32, i32, vp, vp, f32, i32, vp, vp, i64, vp]
    l.lvd_ln_rowstats.argtypes = [vp, i32, i64, i32, f32, vp, vp]
    l.lvd_groupnorm_apply.argtypes = [vp, i64, vp, i64, i32, i64, i32, i32, vp, vp, f32, i32, vp, i32, vp, vp, vp]
    l.lvd_layernorm.argtypes = [vp, i64, vp, i64, i64, i32, vp, vp, f32, vp]
    l.lvd_conv_in.argtypes = [vp, vp, vp, vp, i64, i32, i32, i32, i32, i32, i32, vp]
    l.lvd_conv_out.argtypes = [vp, i64, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp]
    l.lvd_upsample2x.argtypes = [vp, i64, vp, i64, i32, i32, i32, i32, vp]
    l.lvd_timestep_embed.argtypes = [vp, vp, i32, i32, vp]
    l.lvd_silu.argtypes = [vp, vp, i64, vp]
    l.lvd_softmax_rows.argtypes = [vp, i64, i64, i32, ctypes.c_float, vp]
    l.lvd_frames_postprocess.argtypes = [vp, vp, i64, i32, i64, vp]
    l.lvd_im2col_small.argtypes = [vp, vp, i32, i32, i32, i32, i32, vp]
    l.lvd_embed_tokens.argtypes = [vp, vp, vp, vp, i64, i32, i32, i32, vp]
    l.lvd_pointwise_small.argtypes = [vp, vp, vp, vp, i32, i32, i32, i64, vp]
    l.lvd_copy_rows.argtypes = [vp, i64, vp, i64, i64, i32, vp]
    l.lvd_rows_to_ncfhw.argtypes = [vp, i64, vp, i32, i32, i32, i64, vp]
    l.lvd_cfg_dpm_step.argtypes = [vp, vp, vp, vp, vp, vp, i64, f32, f32, f32, f32, f32, f32, f32, i32, vp]
    l.lvd_axpy_h.argtypes = [vp, vp, vp, i64, f32, vp]
    l.lvd_box_energy.argtypes = [ctypes.POINTER(EnergyDesc), vp]
    f9 = ctypes.POINTER(ctypes.c_float)
    l.lvd_attn_smooth.argtypes = [vp, vp, i64, i32, i32, f9, vp]
    l.lvd_attn_smooth_bwd.argtypes = [vp, vp, i32, i64, i32, i32, f9, vp]
    l.lvd_attn_renorm.argtypes = [vp, vp, i64, i32, i32, f32, vp]
    l.lvd_attn_renorm_bwd.argtypes = [vp, vp, vp, i32, i64, i32, i32, f32, vp]
    l.lvd_attn_columns.argtypes = [vp, vp, vp, i32, i64, i32, i32, i32, i32, i32, vp]
    l.lvd_attn_columns_bwd.argtypes = [vp, vp, vp, i32, i32, i64, i32, i32, i32, i32, i32, vp]
    l.lvd_copy_rows_seq.argtypes = [vp, i64, vp, i64, i64, i32, i32, i32, i32, i32, i32, vp]
    l.lvd_gligen_prep.argtypes = [vp, vp, vp, vp, vp, vp, i64, i32, vp]
    l.lvd_groupnorm_bwd.argtypes = [vp, i64, vp, i64, vp, i64, i32, i64, i32, i32, vp, vp, f32, i32, i32, vp, vp, i64,
                                    vp]
    l.lvd_layernorm_bwd.argtypes = [vp, i64, vp, i64, vp, i64, i64, i32, vp, f32, i32, vp]
    l.lvd_geglu.argtypes = [vp, i64, vp, i64, i64, i32, vp]
    l.lvd_geglu_bwd.argtypes = [vp, i64, vp, i64, vp, i64, i64, i32, vp]
    l.lvd_add_rows.argtypes = [vp, i64, vp, i64, vp, i64, i64, i32, vp]
    l.lvd_upsample2x_bwd.argtypes = [vp, i64, vp, i64, i32, i32, i32, i32, vp]
    l.lvd_conv_in_bwd.argtypes = [vp, i64, vp, vp, i32, i32, i32, i32, i32, i32, vp]
    l.lvd_latent_update.argtypes = [vp, vp, vp, vp, i64, f32, f32, vp]
    l.lvd_attn_rowdot.argtypes = [vp, i64, vp, i64, vp, i64, i32, vp]
    l.lvd_self_attn_bwd.argtypes = [vp, i64, i64, i32, i32, i32, vp, i64, vp, vp, vp, i64, i64, i32, i32, f32, vp]
    l.lvd_temporal_attn_bwd.argtypes = [vp, i64, i64, i32, i32, i32, vp, i64, vp, i64, i64, i32, i32, f32, vp]
    l.lvd_cross_attn_bwd.argtypes = [vp, i64, i64, vp, i64, i64, i32, i32, vp, i64, vp, ctypes.POINTER(ctypes.c_int32),
                                     i32, vp, i64, i64, i32, i32, i32, i32, f32, vp, vp]
    for name in EXPORTS:
        fn = getattr(l, name)
        if name not in ("lvd_last_error", "lvd_launch_count"):
            fn.restype = ctypes.c_int


# every symbol include/lvd_b200.h declares (checked by tests/test_abi.py without a GPU)
EXPORTS = [
    "lvd_last_error",
    "lvd_version",
    "lvd_debug_trace_read",
    "lvd_launch_count",
    "lvd_gemm",
    "lvd_attention",
    "lvd_groupnorm",
    "lvd_groupnorm_apply",
    "lvd_ln_rowstats",
    "lvd_layernorm",
    "lvd_conv_in",
    "lvd_conv_out",
    "lvd_upsample2x",
    "lvd_timestep_embed",
    "lvd_silu",
    "lvd_softmax_rows",
    "lvd_pointwise_small",
    "lvd_embed_tokens",
    "lvd_im2col_small",
    "lvd_frames_postprocess",
    "lvd_copy_rows",
    "lvd_rows_to_ncfhw",
    "lvd_cfg_dpm_step",
    "lvd_axpy_h",
    "lvd_box_energy",
    "lvd_attn_smooth",
    "lvd_attn_smooth_bwd",
    "lvd_attn_renorm",
    "lvd_attn_renorm_bwd",
    "lvd_attn_columns",
    "lvd_attn_columns_bwd",
    "lvd_copy_rows_seq",
    "lvd_gligen_prep",
    "lvd_groupnorm_bwd",
    "lvd_layernorm_bwd",
    "lvd_geglu",
    "lvd_geglu_bwd",
    "lvd_add_rows",
    "lvd_upsample2x_bwd",
    "lvd_conv_in_bwd",
    "lvd_latent_update",
    "lvd_attn_rowdot",
    "lvd_self_attn_bwd",
    "lvd_temporal_attn_bwd",
    "lvd_cross_attn_bwd",
]


def check(rc, what=""):
    if rc != 0:
        msg = lib().lvd_last_error().decode("utf-8", "replace")
        if rc == -1:
            raise ValueError(f"lvd_b200 {what}: {msg}")
        raise LvdError(f"lvd_b200 {what} failed (rc={rc}): {msg}")
