"""ctypes binding of librvc_b200.so (declared in include/rvc_b200.h).

There is NO fallback: if the library is missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# RVC_B200_LIB: debugging builds of the same library only (tools/pair_trace.py -> librvc_b200_trace.so); never a fallback
LIB_PATH = os.environ.get("RVC_B200_LIB") or os.path.join(HERE, "librvc_b200.so")
ABI_VERSION = 30
DT_BF16, DT_FP16, DT_FP16_SPLIT = 0, 2, 3

ACT_NONE, ACT_RELU, ACT_TANH = 0, 1, 2

_vp, _i32, _i64, _f32, _f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_double


class Conv1dDesc(C.Structure):
    _fields_ = [
        ("x", _vp), ("x_bs", _i64), ("x_cs", _i64),
        ("w", _vp),
        ("bias", _vp),
        ("bias_b", _vp), ("bias_b_bs", _i64),
        ("mask", _vp),
        ("res", _vp), ("res_bs", _i64), ("res_cs", _i64),
        ("y", _vp), ("y_bs", _i64), ("y_cs", _i64),
        ("B", _i32), ("Cin", _i32), ("Cout", _i32), ("L", _i32), ("K", _i32), ("dil", _i32), ("pad", _i32),
        ("pre_lrelu", _i32), ("pre_slope", _f32),
        ("in_mask", _i32), ("out_mask", _i32),
        ("act", _i32), ("gated", _i32),
        ("alpha", _f32), ("out_scale", _f32), ("beta", _f32),
    ]


class ConvTr1dDesc(C.Structure):
    _fields_ = [
        ("x", _vp), ("x_bs", _i64), ("x_cs", _i64),
        ("w", _vp), ("bias", _vp),
        ("har", _vp), ("har_bs", _i64),
        ("nw", _vp), ("nbias", _vp),
        ("y", _vp), ("y_bs", _i64), ("y_cs", _i64),
        ("B", _i32), ("Cin", _i32), ("Cout", _i32), ("Lin", _i32), ("k", _i32), ("u", _i32), ("kn", _i32), ("sn", _i32), ("pn", _i32),
        ("pre_slope", _f32),
        ("x2", _vp), ("x3", _vp), ("in_div", _f32),
    ]


class TcJob(C.Structure):
    _fields_ = [
        ("a", _vp), ("w", _vp),
        ("bias", _vp),
        ("bias_b", _vp), ("bias_b_bs", _i64),
        ("mask", _vp),
        ("res", _vp), ("res_bs", _i64), ("res_cs", _i64),
        ("y_f32", _vp), ("y_bs", _i64), ("y_cs", _i64),
        ("y_act", _vp),
        ("K", _i32), ("dil", _i32), ("pad", _i32), ("gated", _i32), ("relu", _i32),
        ("alpha", _f32), ("act_slope", _f32),
        ("a_bs", _i64), ("w_bs", _i64), ("y_act_bs", _i64),
        ("w_cs", _i32), ("reserved", _i32),
        ("nz_har", _vp), ("nz_w", _vp), ("nz_b", _vp), ("nz_har_bs", _i64), ("nz_kn", _i32), ("nz_sn", _i32), ("nz_pn", _i32), ("reserved2", _i32),
        ("y_lo", _vp),
        ("res_hi", _vp), ("res_lo", _vp),
    ]


class Conv1dTcDesc(C.Structure):
    _fields_ = [("jobs", TcJob * 6), ("njobs", _i32), ("B", _i32), ("Cin", _i32), ("Cout", _i32), ("L", _i32), ("dtype", _i32),
                ("nt", _i32), ("mt", _i32), ("up", _i32), ("pshift", _i32), ("sum_jobs", _i32), ("f32_blocked", _i32), ("up_cols", _i32), ("reserved", _i32)]


class PairJob(C.Structure):
    """rvc_pair_job (include/rvc_b200.h)."""
    _fields_ = [("a", _vp), ("w1", _vp), ("b1", _vp), ("w2", _vp), ("b2", _vp), ("res", _vp), ("y", _vp), ("y_act", _vp),
                ("K1", _i32), ("dil1", _i32), ("K2", _i32), ("reserved", _i32), ("a_lo", _vp), ("y_lo", _vp)]


class FlowCoupling(C.Structure):
    """rvc_flow_coupling."""
    _fields_ = [("pre_w", _vp), ("in_w", _vp * 3), ("res_w", _vp * 2), ("skip_w", _vp * 3), ("post_w", _vp),
                ("hb", _vp * 3), ("in_b", _vp * 3), ("sb", _vp), ("post_b", _vp), ("gc_off", _i32), ("reserved", _i32)]


class FlowDesc(C.Structure):
    """rvc_flow_desc."""
    _fields_ = [("flows", FlowCoupling * 4), ("zp", _vp), ("z", _vp), ("za", _vp), ("mask", _vp), ("gc", _vp), ("gc_bs", _i64),
                ("B", _i32), ("T", _i32), ("hidden", _i32), ("half", _i32), ("n_flows", _i32), ("n_layers", _i32), ("kernel_size", _i32),
                ("post_rows", _i32)]


class AttnScoresDesc(C.Structure):
    """rvc_attn_scores_desc."""
    _fields_ = [("qk", _vp), ("qk_bs", _i64), ("q_chunk0", _i32), ("k_chunk0", _i32), ("lo_chunks", _i32), ("planes", _i32),
                ("ek", _vp), ("qkv", _vp), ("mask", _vp), ("P", _vp), ("vw", _vp),
                ("B", _i32), ("H", _i32), ("n_heads", _i32), ("T", _i32), ("window", _i32), ("reserved", _i32),
                ("att", _vp), ("att_bs", _i64), ("evT", _vp), ("att_lo_chunks", _i32), ("reserved2", _i32)]


class EncoderTailDesc(C.Structure):
    """rvc_encoder_tail_desc."""
    _fields_ = [("att", _vp), ("att_bs", _i64), ("x", _vp), ("y", _vp), ("ya", _vp), ("ya_bs", _i64), ("mask", _vp),
                ("wo_hi", _vp), ("wo_lo", _vp), ("bo", _vp), ("w1", _vp), ("b1", _vp), ("w2", _vp), ("b2", _vp),
                ("gamma1", _vp), ("beta1", _vp), ("gamma2", _vp), ("beta2", _vp),
                ("B", _i32), ("T", _i32), ("hidden", _i32), ("filter", _i32), ("kernel_size", _i32), ("out_mask", _i32), ("eps", _f32), ("reserved", _i32)]


class ResblockPairDesc(C.Structure):
    """rvc_resblock_pair_desc."""
    _fields_ = [("jobs", PairJob * 3), ("njobs", _i32), ("B", _i32), ("C", _i32), ("L", _i32), ("dtype", _i32), ("slope", _f32),
                ("sum_jobs", _i32), ("y_ncl", _i32), ("alpha", _f32), ("reserved", _i32)]


# name -> (restype, argtypes); kept in one table so tests can check every header symbol is exported
SIGNATURES = {
    "rvc_abi_version": (C.c_int, []),
    "rvc_last_error": (C.c_char_p, []),
    "rvc_device_info": (C.c_int, [C.c_int, C.POINTER(C.c_int)]),
    "rvc_conv1d_f32": (C.c_int, [C.POINTER(Conv1dDesc), _vp]),
    "rvc_convtr1d_f32": (C.c_int, [C.POINTER(ConvTr1dDesc), _vp]),
    "rvc_conv_post_f32": (C.c_int, [_vp, _vp, _vp, _f32, _i64, _i64, _vp, _vp, _i64, _i32, _i32, _i32, _f32, _vp]),
    "rvc_embed_f32": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "rvc_embed_prep": (C.c_int, [_vp, _vp, _vp, _f32, _vp, _vp, _i32, _i32, _i32, _vp]),
    "rvc_pack_rows": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "rvc_layernorm_f32": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _f32, _i32, _vp, _i32, _vp]),
    "rvc_rel_attention_f32": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp]),
    "rvc_cond_gemv_f32": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp]),
    "rvc_sample_zp_f32": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp]),
    "rvc_sine_source_f32": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _f32, _vp]),
    "rvc_noise_conv_f32": (C.c_int, [_vp, _i64, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "rvc_feature_prep_f32": (C.c_int, [_vp, _vp, _vp, _i64, _f32, _vp, _i32, _i32, _i32, _i32, _vp]),
    "rvc_absmax_f32": (C.c_int, [_vp, _i64, _i64, _vp, _i32, _vp]),
    "rvc_rmvpe_decode_f32": (C.c_int, [_vp, _i64, _i32, _f32, _f64, _f64, _vp, _vp]),
    "rvc_f0_coarse_f64": (C.c_int, [_vp, _i64, _vp, _i32, _f64, _f64, _f64, _f64, _vp, _vp, _vp, _vp]),
    "rvc_mel_spectrogram_f32": (C.c_int, [_vp, _i64, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _f32, _vp]),
    "rvc_mel_basis_ranges": (C.c_int, [_vp, _vp, _vp]),
    "rvc_split_points_workspace": (_i64, [_i64, _i64, _i64]),
    "rvc_split_points": (C.c_int, [_vp, _i32, _i64, _i32, _i64, _i64, _vp, _vp, _vp]),
    "rvc_filtfilt_block": (_i32, [_i64]),
    "rvc_filtfilt_workspace": (_i64, [_i64, _i32]),
    "rvc_filtfilt": (C.c_int, [_vp, _i32, _i64, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp]),
    "rvc_retrieval_blend_f32": (C.c_int, [_vp, _vp, _vp, _vp, _f32, _f32, _vp, _i32, _i32, _i32, _vp]),
    "rvc_rmvpe_pitch_f32": (C.c_int, [_vp, _i64, _i32, _f32, _f64, _f64, _vp, _i32, _f64, _f64, _f64, _f64, _vp, _vp, _vp, _vp]),
    "rvc_resblock_pair_tc": (C.c_int, [C.POINTER(ResblockPairDesc), _vp]),
    "rvc_debug_pair_trace": (C.c_int, [_vp, _i32]),
    "rvc_set_grid_limit": (C.c_int, [_i32]),
    "rvc_packed_rows": (_i64, [_i32]),
    "rvc_pack_act": (C.c_int, [_vp, _vp, _vp, _f32, _i64, _i64, _vp, _vp, _i32, _i32, _i32, _i32, _f32, _vp]),
    "rvc_pack_weight": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "rvc_conv1d_tc": (C.c_int, [C.POINTER(Conv1dTcDesc), _vp]),
    "rvc_flow_reverse_tc": (C.c_int, [C.POINTER(FlowDesc), _vp]),
    "rvc_attn_scores_cluster_tc": (C.c_int, [C.POINTER(AttnScoresDesc), _vp]),
    "rvc_encoder_tail_tc": (C.c_int, [C.POINTER(EncoderTailDesc), _vp]),
    "rvc_conv1d_tc_chain_blob_bytes": (C.c_int64, [_i32]),
    "rvc_conv1d_tc_chain_build": (C.c_int, [C.POINTER(Conv1dTcDesc), C.POINTER(_i32), _i32, _vp, _i64]),
    "rvc_conv1d_tc_chain_run": (C.c_int, [_vp, _vp, _vp, _vp]),
    "rvc_attn_softmax": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
}

_lib = None


def load():
    """Load the library once; raise loudly when it is absent or stale."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m rvc_maker_b200.build` (needs nvcc, no GPU required). "
            "rvc_maker_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    v = lib.rvc_abi_version()
    if v != ABI_VERSION:
        raise RuntimeError(f"librvc_b200.so ABI {v} != expected {ABI_VERSION}; rebuild with `python -m rvc_maker_b200.build --force`")
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().rvc_last_error().decode(errors="replace")
        raise RuntimeError(f"librvc_b200 call failed ({what}): rc={rc}: {msg}")
