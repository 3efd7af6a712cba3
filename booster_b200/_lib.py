"""ctypes binding of libbooster_b200.so (include/bridge.h + include/booster_b200.h).

The shared library is the product; this module only declares prototypes. It is built in-tree by
`make` / `__graft_entry__.build()` and must exist: there is no Python or CPU fallback for any compute entry.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# BOOSTER_B200_LIB: A/B builds of the SAME sources with different -D switches (scripts/gpu_variants.sh); never a fallback
LIB_PATH = os.environ.get("BOOSTER_B200_LIB") or os.path.join(_HERE, "libbooster_b200.so")

_lib = None

# every symbol the two headers declare (tests check the .so exports exactly these)
BRIDGE_SYMBOLS = ["init", "initContext", "doInference", "stopInference", "status", "promptEval",
                  "getPromptTokenCount", "timing", "getSeed"]
B200_SYMBOLS = ["b200_last_error", "b200_device_count", "b200_version", "b200_model_load", "b200_model_free",
                "b200_model_info", "b200_model_weight_bytes", "b200_ctx_new", "b200_ctx_free", "b200_n_ctx",
                "b200_kv_clear", "b200_decode", "b200_generate_greedy", "b200_step_greedy", "b200_set_taps", "b200_get_tap",
                "b200_timings", "b200_reset_timings", "b200_kernel_launches", "b200_last_device_ms", "b200_profile_token", "b200_profile_kind", "b200_trace_token", "b200_trace_phases", "b200_set_token_kernel", "b200_set_prefill_batch", "b200_set_prefill_mma", "b200_set_prefill_attn_batch", "b200_job_timing_us", "b200_comm_unique_id",
                "b200_comm_init", "b200_p2p_handle", "b200_p2p_connect", "b200_p2p_disable", "b200_pipeline_generate_greedy", "b200_pipeline_decode", "b200_stage_forward", "b200_stage_batch_usable", "b200_stage_forward_batch",
                "b200_stage_logits", "b200_stage_argmax", "b200_stage_logits_view", "b200_decode_view", "b200_stage_sync", "b200_kv_write", "b200_kv_read", "b200_kv_seq_rm", "b200_kv_seq_add", "b200_kv_seq_div", "b200_op_quantize_q8_K", "b200_op_quantize_q8_0",
                "b200_op_dequantize_row", "b200_op_mul_mat_vec", "b200_op_mul_mat", "b200_op_rms_norm", "b200_op_rope",
                "b200_op_attention", "b200_set_attention_route", "b200_tokenizer_load", "b200_tokenizer_free", "b200_tokenizer_n_vocab", "b200_tokenize",
                "b200_token_to_piece", "b200_token_is_eog", "b200_token_nl", "b200_cpt_class", "b200_op_launch_shape",
                "b200_sampler_new", "b200_sampler_set_standard", "b200_sampler_free", "b200_sampler_reset", "b200_sampler_sample"]


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `make` (or __graft_entry__.build()). "
                           "booster_b200 has no fallback path.")
    L = C.CDLL(LIB_PATH)
    i32p, f32p = C.POINTER(C.c_int32), C.POINTER(C.c_float)
    vp, cp = C.c_void_p, C.c_char_p

    def sig(name, res, args):
        # (an A/B library built from an older revision lacks the newest entry points: tests/test_cabi_surface.py is what
        #  guarantees that the shipped library exports every symbol the headers declare)
        fn = getattr(L, name, None)
        if fn is not None:
            fn.restype, fn.argtypes = res, args

    # include/bridge.h
    sig("init", None, [cp, cp])
    sig("initContext", vp, [C.c_int, cp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                            C.c_int32, C.c_float, C.c_float, C.c_float, C.c_int, C.c_float, C.c_float, C.c_float,
                            C.c_int, C.c_int32, C.c_int32, C.c_float, C.c_float, C.c_float, C.c_uint32, cp])
    sig("doInference", C.c_int64, [C.c_int, vp, cp, cp, cp])
    sig("stopInference", None, [C.c_int])
    sig("status", cp, [cp])
    sig("promptEval", C.c_int64, [cp])
    sig("getPromptTokenCount", C.c_int64, [cp])
    sig("timing", C.c_int64, [cp])
    sig("getSeed", C.c_uint32, [cp])
    sig("b200_job_timing_us", C.c_int, [cp, C.POINTER(C.c_double), C.POINTER(C.c_double)])
    # include/booster_b200.h
    sig("b200_last_error", cp, [])
    sig("b200_device_count", C.c_int, [])
    sig("b200_version", cp, [])
    sig("b200_model_load", vp, [cp, C.c_int, C.c_int, C.c_int])
    sig("b200_model_free", None, [vp])
    sig("b200_model_info", C.c_int, [vp, i32p])
    sig("b200_model_weight_bytes", C.c_int64, [vp])
    sig("b200_ctx_new", vp, [vp, C.c_int])
    sig("b200_ctx_free", None, [vp])
    sig("b200_n_ctx", C.c_int, [vp])
    sig("b200_kv_clear", None, [vp])
    sig("b200_decode", C.c_int, [vp, i32p, C.c_int, C.c_int, f32p])
    sig("b200_generate_greedy", C.c_int, [vp, C.c_int32, C.c_int, C.c_int, i32p])
    sig("b200_set_taps", None, [vp, C.c_int])
    sig("b200_get_tap", C.c_int64, [vp, cp, C.c_int, f32p, C.c_int64])
    sig("b200_timings", None, [vp, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_double), C.POINTER(C.c_int64)])
    sig("b200_reset_timings", None, [vp])
    sig("b200_kernel_launches", C.c_int64, [vp])
    sig("b200_last_device_ms", C.c_float, [vp])
    sig("b200_step_greedy", C.c_int, [vp, C.c_int32, C.c_int, i32p])
    sig("b200_tokenizer_load", vp, [cp])
    sig("b200_tokenizer_free", None, [vp])
    sig("b200_tokenizer_n_vocab", C.c_int32, [vp])
    sig("b200_tokenize", C.c_int32, [vp, C.c_char_p, C.c_int32, i32p, C.c_int32, C.c_int, C.c_int])
    sig("b200_token_to_piece", C.c_int32, [vp, C.c_int32, C.c_char_p, C.c_int32, C.c_int])
    sig("b200_token_is_eog", C.c_int, [vp, C.c_int32])
    sig("b200_token_nl", C.c_int32, [vp])
    sig("b200_cpt_class", C.c_int, [C.c_uint32])
    sig("b200_op_launch_shape", C.c_int, [i32p, C.c_int, C.c_int64, C.c_int64, C.c_int, C.c_int, i32p])
    sig("b200_profile_token", C.c_int, [vp, C.c_int32, C.c_int, f32p, i32p])
    sig("b200_profile_kind", C.c_int, [vp, C.c_int, C.c_int, C.c_int, f32p, i32p])
    sig("b200_trace_token", C.c_int64, [vp, C.c_int32, C.c_int, C.c_int, C.POINTER(C.c_uint64), C.c_int64, i32p, C.c_int64])
    sig("b200_comm_unique_id", C.c_int, [C.POINTER(C.c_uint8)])
    sig("b200_comm_init", C.c_int, [vp, C.c_int, C.c_int, C.POINTER(C.c_uint8)])
    sig("b200_p2p_handle", C.c_int, [vp, C.POINTER(C.c_uint8)])
    sig("b200_p2p_connect", C.c_int, [vp, C.c_int, C.c_int, C.POINTER(C.c_uint8), C.POINTER(C.c_uint8)])
    sig("b200_p2p_disable", None, [vp])
    sig("b200_pipeline_generate_greedy", C.c_int, [vp, C.c_int32, C.c_int, C.c_int, i32p])
    sig("b200_pipeline_decode", C.c_int, [vp, i32p, C.c_int, C.c_int, f32p])
    sig("b200_stage_forward", C.c_int, [vp, C.c_int32, C.c_int, C.c_int, vp])
    sig("b200_stage_batch_usable", C.c_int, [vp, C.c_int])
    sig("b200_stage_forward_batch", C.c_int, [vp, i32p, C.c_int, C.c_int, vp])
    sig("b200_stage_logits", C.c_int, [vp, f32p])
    sig("b200_stage_argmax", C.c_int, [vp, i32p])
    sig("b200_stage_sync", C.c_int, [vp])
    sig("b200_stage_logits_view", f32p, [vp])
    sig("b200_decode_view", f32p, [vp, C.c_int32, C.c_int])
    sig("b200_kv_write", C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint16), C.POINTER(C.c_uint16)])
    sig("b200_kv_read", C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint16), C.POINTER(C.c_uint16)])
    sig("b200_op_quantize_q8_K", C.c_int, [f32p, C.c_int64, vp])
    sig("b200_op_quantize_q8_0", C.c_int, [f32p, C.c_int64, vp])
    sig("b200_op_dequantize_row", C.c_int, [C.c_int, vp, C.c_int64, f32p])
    sig("b200_op_mul_mat_vec", C.c_int, [C.c_int, vp, C.c_int64, C.c_int64, f32p, f32p])
    sig("b200_op_mul_mat", C.c_int, [C.c_int, vp, C.c_int64, C.c_int64, f32p, C.c_int64, f32p])
    sig("b200_op_rms_norm", C.c_int, [f32p, f32p, C.c_int64, C.c_float, f32p])
    sig("b200_op_rope", C.c_int, [f32p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, f32p])
    sig("b200_op_attention", C.c_int, [f32p, C.POINTER(C.c_uint16), C.POINTER(C.c_uint16), C.c_int, C.c_int, C.c_int,
                                       C.c_int, C.c_float, C.c_int, f32p])
    sig("b200_set_attention_route", None, [C.c_int])
    sig("b200_kv_seq_rm", C.c_int, [vp, C.c_int, C.c_int])
    sig("b200_kv_seq_add", C.c_int, [vp, C.c_int, C.c_int, C.c_int])
    sig("b200_kv_seq_div", C.c_int, [vp, C.c_int, C.c_int, C.c_int])
    sig("b200_sampler_new", vp, [cp, C.c_int, C.c_int32, C.c_int32, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, C.c_float, C.c_float, C.c_int])
    sig("b200_sampler_set_standard", None, [vp, C.c_int32, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float])
    sig("b200_sampler_free", None, [vp])
    sig("b200_sampler_reset", None, [vp, i32p, C.c_int32, C.c_uint32])
    sig("b200_sampler_sample", C.c_int32, [vp, f32p, C.c_int32, C.c_int32])
    sig("b200_set_token_kernel", None, [C.c_int])
    sig("b200_set_prefill_batch", None, [C.c_int])
    sig("b200_set_prefill_mma", None, [C.c_int])
    sig("b200_set_prefill_attn_batch", None, [C.c_int])
    sig("b200_trace_phases", C.c_int64, [vp, C.c_int32, C.c_int, C.c_int, C.POINTER(C.c_uint64), C.c_int64, i32p, C.c_int64, i32p])
    _lib = L
    return L


def last_error() -> str:
    return lib().b200_last_error().decode(errors="replace")


class B200Error(RuntimeError):
    pass


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        raise B200Error(f"{what}: {last_error()}")
