"""Python mirror of the token-level interface (include/booster_b200.h), named after the llama.h calls
the reference's bridge makes on the hot path (cpp/bridge.cpp:118-171, 549-560; cpp/janus.cpp:224) so that the
parity tests read like tests of the reference: load_model_from_file -> new_context_with_model -> decode ->
get_logits. All compute happens inside libbooster_b200.so on the GPU; this file only marshals arguments.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import B200Error, check, last_error

INFO = ["n_vocab", "n_embd", "n_layer", "n_head", "n_head_kv", "n_ff", "head_dim", "n_ctx_train",
        "layer_begin", "layer_end", "ftype"]
TYPE = {"F32": 0, "F16": 1, "Q8_0": 8, "Q4_K": 12, "Q5_K": 13, "Q6_K": 14}


def device_count() -> int:
    return _lib.lib().b200_device_count()


class Model:
    """llama_load_model_from_file (cpp/bridge.cpp:131) for one pipeline stage [layer_begin, layer_end)."""

    def __init__(self, path: str, device: int = 0, layer_begin: int = 0, layer_end: int = -1):
        self.L = _lib.lib()
        self.h = self.L.b200_model_load(path.encode(), device, layer_begin, layer_end)
        if not self.h:
            raise B200Error(f"b200_model_load({path}): {_lib.last_error()}")
        arr = (C.c_int32 * 16)()
        self.L.b200_model_info(self.h, arr)
        for i, k in enumerate(INFO):
            setattr(self, k, int(arr[i]))
        self.weight_bytes = int(self.L.b200_model_weight_bytes(self.h))

    def close(self):
        if self.h:
            self.L.b200_model_free(self.h)
            self.h = None


class Context:
    """llama_new_context_with_model (cpp/bridge.cpp:162) + the decode calls."""

    def __init__(self, model: Model, n_ctx: int = 2048):
        self.L, self.model = model.L, model
        self.h = self.L.b200_ctx_new(model.h, n_ctx)
        if not self.h:
            raise B200Error(f"b200_ctx_new: {_lib.last_error()}")
        self.n_ctx = self.L.b200_n_ctx(self.h)

    def close(self):
        if self.h:
            self.L.b200_ctx_free(self.h)
            self.h = None

    def kv_clear(self):
        self.L.b200_kv_clear(self.h)

    def kv_write(self, layer: int, pos0: int, k_rows, v_rows):
        """rows [pos0, pos0+n) of one layer's K / V cache from f16 arrays [n, n_head_kv*head_dim] (test support)"""
        k = np.ascontiguousarray(k_rows, dtype=np.float16).view(np.uint16)
        v = np.ascontiguousarray(v_rows, dtype=np.float16).view(np.uint16)
        u16p = C.POINTER(C.c_uint16)
        check(self.L.b200_kv_write(self.h, layer, pos0, k.shape[0], k.ctypes.data_as(u16p), v.ctypes.data_as(u16p)), "b200_kv_write")

    def kv_read(self, layer: int, pos0: int, n: int):
        kvd = self.model.n_head_kv * self.model.head_dim
        k = np.empty((n, kvd), dtype=np.uint16)
        v = np.empty((n, kvd), dtype=np.uint16)
        u16p = C.POINTER(C.c_uint16)
        check(self.L.b200_kv_read(self.h, layer, pos0, n, k.ctypes.data_as(u16p), v.ctypes.data_as(u16p)), "b200_kv_read")
        return k.view(np.float16), v.view(np.float16)

    def kv_seq_rm(self, p0: int, p1: int):
        """llama_kv_cache_seq_rm(ctx, 0, p0, p1)"""
        check(self.L.b200_kv_seq_rm(self.h, p0, p1), "b200_kv_seq_rm")

    def kv_seq_add(self, p0: int, p1: int, delta: int):
        """llama_kv_cache_seq_add(ctx, 0, p0, p1, delta); the K-shift runs inside the next decode"""
        check(self.L.b200_kv_seq_add(self.h, p0, p1, delta), "b200_kv_seq_add")

    def kv_seq_div(self, p0: int, p1: int, d: int):
        """llama_kv_cache_seq_div(ctx, 0, p0, p1, d) (cpp/bridge.cpp:518): Self-Extend's grouping of positions"""
        check(self.L.b200_kv_seq_div(self.h, p0, p1, d), "b200_kv_seq_div")

    def decode(self, tokens: Sequence[int], pos0: int, want_logits: bool = True) -> Optional[np.ndarray]:
        """llama_decode(ctx, llama_batch_get_one(tokens, n, pos0, 0)) then llama_get_logits (last token's row)."""
        toks = np.ascontiguousarray(tokens, dtype=np.int32)
        out = np.empty(self.model.n_vocab, dtype=np.float32) if want_logits else None
        check(self.L.b200_decode(self.h, toks.ctypes.data_as(C.POINTER(C.c_int32)), len(toks), pos0,
                                 out.ctypes.data_as(C.POINTER(C.c_float)) if want_logits else None), "b200_decode")
        return out

    def generate_greedy(self, first_token: int, pos0: int, n_steps: int) -> np.ndarray:
        out = np.empty(n_steps, dtype=np.int32)
        check(self.L.b200_generate_greedy(self.h, first_token, pos0, n_steps, out.ctypes.data_as(C.POINTER(C.c_int32))),
              "b200_generate_greedy")
        return out

    def greedy(self, prompt: Sequence[int], n_gen: int) -> (List[int], List[np.ndarray]):
        """Same protocol as oracle.ref.RefModel.greedy: prefill in one batch, n_gen arg-max steps with logits."""
        self.kv_clear()
        logits = self.decode(prompt, 0)
        pos = len(prompt)
        ids, all_logits = [], []
        for _ in range(n_gen):
            all_logits.append(logits)
            t = int(np.argmax(logits))
            ids.append(t)
            logits = self.decode([t], pos)
            pos += 1
        return ids, all_logits

    def set_taps(self, enable: bool):
        self.L.b200_set_taps(self.h, int(enable))

    def get_tap(self, name: str, layer: int) -> Optional[np.ndarray]:
        n = self.L.b200_get_tap(self.h, name.encode(), layer, None, 0)
        if n == 0:
            return None
        out = np.empty(n, dtype=np.float32)
        self.L.b200_get_tap(self.h, name.encode(), layer, out.ctypes.data_as(C.POINTER(C.c_float)), n)
        return out

    def last_device_ms(self) -> float:
        return float(self.L.b200_last_device_ms(self.h))

    def profile_token(self, token: int, pos: int):
        """device ms and launch count per kernel kind for one un-graphed token (event pair around every launch)"""
        ms = np.zeros(8, dtype=np.float32)
        cnt = np.zeros(8, dtype=np.int32)
        check(self.L.b200_profile_token(self.h, token, pos, ms.ctypes.data_as(C.POINTER(C.c_float)),
                                        cnt.ctypes.data_as(C.POINTER(C.c_int32))), "b200_profile_token")
        kinds = ["embed", "qkv", "attention", "wo", "gate_up", "down", "head", "attn_pv_split_route"]
        return {k: (float(ms[i]), int(cnt[i])) for i, k in enumerate(kinds)}

    KINDS = ["embed", "qkv", "attention", "wo", "gate_up", "down", "head", "attn_pv_split_route"]

    def profile_kind(self, kind: str, pos: int, reps: int = 4):
        """(ms per launch, launches) of ONE kernel kind launched back to back over every layer, `reps` times"""
        ms = C.c_float(0.0)
        n = C.c_int32(0)
        check(self.L.b200_profile_kind(self.h, self.KINDS.index(kind), pos, reps, C.byref(ms), C.byref(n)), "b200_profile_kind")
        return float(ms.value) / max(1, n.value), int(n.value)

    def trace_token(self, token: int, pos: int, reps: int = 3):
        """(stamps[n_launches, 512, 12] u64 ns, meta[n_launches, 2] = (kind, ctas)) of one graph-replayed token"""
        cap_l = 1024
        out = np.zeros((cap_l, 512, 12), dtype=np.uint64)
        meta = np.zeros((cap_l, 2), dtype=np.int32)
        n = int(self.L.b200_trace_token(self.h, token, pos, reps, out.ctypes.data_as(C.POINTER(C.c_uint64)), out.size,
                                        meta.ctypes.data_as(C.POINTER(C.c_int32)), meta.size))
        if n < 0:
            raise B200Error(f"b200_trace_token: {last_error()}")
        return out[:n], meta[:n]

    def trace_phases(self, token: int, pos: int, reps: int = 3):
        """(stamps[n_phases, n_ctas, 4] u64 ns, kinds[n_phases]) of one token through the persistent kernel; None if not in use"""
        cap_p, cap_c = 1024, 256
        out = np.zeros(cap_p * cap_c * 4, dtype=np.uint64)
        kinds = np.zeros(cap_p, dtype=np.int32)
        nc = C.c_int32(0)
        n = int(self.L.b200_trace_phases(self.h, token, pos, reps, out.ctypes.data_as(C.POINTER(C.c_uint64)), out.size,
                                         kinds.ctypes.data_as(C.POINTER(C.c_int32)), kinds.size, C.byref(nc)))
        if n < 0:
            raise B200Error(f"b200_trace_phases: {last_error()}")
        if n == 0:
            return None
        return out[:n * nc.value * 4].reshape(n, nc.value, 4), kinds[:n]

    def kernel_launches(self) -> int:
        return int(self.L.b200_kernel_launches(self.h))

    # pipeline over NCCL (one process per GPU)
    def comm_init(self, rank: int, world: int, uid: bytes):
        buf = (C.c_uint8 * 128).from_buffer_copy(uid)
        check(self.L.b200_comm_init(self.h, rank, world, buf), "b200_comm_init")

    def p2p_handle(self) -> bytes:
        """this rank's inbox as a 64-byte CUDA IPC handle"""
        buf = (C.c_uint8 * 64)()
        check(self.L.b200_p2p_handle(self.h, buf), "b200_p2p_handle")
        return bytes(buf)

    def p2p_connect(self, rank: int, world: int, next_handle: bytes, first_handle: bytes):
        nh = (C.c_uint8 * 64).from_buffer_copy(next_handle)
        fh = (C.c_uint8 * 64).from_buffer_copy(first_handle)
        check(self.L.b200_p2p_connect(self.h, rank, world, nh, fh), "b200_p2p_connect")

    def p2p_disable(self):
        self.L.b200_p2p_disable(self.h)

    def pipeline_generate_greedy(self, first_token: int, pos0: int, n_steps: int) -> np.ndarray:
        out = np.empty(n_steps, dtype=np.int32)
        check(self.L.b200_pipeline_generate_greedy(self.h, first_token, pos0, n_steps,
                                                   out.ctypes.data_as(C.POINTER(C.c_int32))), "b200_pipeline_generate_greedy")
        return out

    def pipeline_decode(self, tokens: Sequence[int], pos0: int, want_logits: bool = False) -> Optional[np.ndarray]:
        toks = np.ascontiguousarray(tokens, dtype=np.int32)
        out = np.empty(self.model.n_vocab, dtype=np.float32) if want_logits else None
        check(self.L.b200_pipeline_decode(self.h, toks.ctypes.data_as(C.POINTER(C.c_int32)), len(toks), pos0,
                                          out.ctypes.data_as(C.POINTER(C.c_float)) if want_logits else None), "b200_pipeline_decode")
        return out


def comm_unique_id() -> bytes:
    buf = (C.c_uint8 * 128)()
    check(_lib.lib().b200_comm_unique_id(buf), "b200_comm_unique_id")
    return bytes(buf)


# ---- operator-level wrappers (parity tests) ----------------------------------------------------------------
def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def op_quantize_q8_K(x) -> np.ndarray:
    x = _f32(x)
    out = np.empty(x.size // 256 * 292, dtype=np.uint8)
    check(_lib.lib().b200_op_quantize_q8_K(x.ctypes.data_as(C.POINTER(C.c_float)), x.size, out.ctypes.data), "op_quantize_q8_K")
    return out


def op_quantize_q8_0(x) -> np.ndarray:
    x = _f32(x)
    out = np.empty(x.size // 32 * 34, dtype=np.uint8)
    check(_lib.lib().b200_op_quantize_q8_0(x.ctypes.data_as(C.POINTER(C.c_float)), x.size, out.ctypes.data), "op_quantize_q8_0")
    return out


def op_dequantize_row(t: int, raw, k: int) -> np.ndarray:
    raw = np.ascontiguousarray(raw, dtype=np.uint8)
    out = np.empty(k, dtype=np.float32)
    check(_lib.lib().b200_op_dequantize_row(t, raw.ctypes.data, k, out.ctypes.data_as(C.POINTER(C.c_float))), "op_dequantize_row")
    return out


def op_mul_mat_vec(t: int, w_raw, n_rows: int, k: int, x) -> np.ndarray:
    w_raw = np.ascontiguousarray(w_raw, dtype=np.uint8)
    x = _f32(x)
    y = np.empty(n_rows, dtype=np.float32)
    check(_lib.lib().b200_op_mul_mat_vec(t, w_raw.ctypes.data, n_rows, k, x.ctypes.data_as(C.POINTER(C.c_float)),
                                         y.ctypes.data_as(C.POINTER(C.c_float))), "op_mul_mat_vec")
    return y


def op_mul_mat(t: int, w_raw, n_rows: int, k: int, x) -> np.ndarray:
    """x[T][k] -> y[T][n_rows] through the prompt-batch kernels (T <= 512)"""
    w_raw = np.ascontiguousarray(w_raw, dtype=np.uint8)
    x = np.ascontiguousarray(x, dtype=np.float32)
    T = x.shape[0]
    y = np.empty((T, n_rows), dtype=np.float32)
    check(_lib.lib().b200_op_mul_mat(t, w_raw.ctypes.data, n_rows, k, x.ctypes.data_as(C.POINTER(C.c_float)), T,
                                     y.ctypes.data_as(C.POINTER(C.c_float))), "op_mul_mat")
    return y


def op_rms_norm(x, w, eps: float) -> np.ndarray:
    x = _f32(x)
    y = np.empty_like(x)
    wp = _f32(w).ctypes.data_as(C.POINTER(C.c_float)) if w is not None else None
    check(_lib.lib().b200_op_rms_norm(x.ctypes.data_as(C.POINTER(C.c_float)), wp, x.size, eps,
                                      y.ctypes.data_as(C.POINTER(C.c_float))), "op_rms_norm")
    return y


def op_rope(x, n_heads: int, head_dim: int, pos: int, freq_base: float, freq_scale: float = 1.0, freq_factors=None) -> np.ndarray:
    y = _f32(x).copy()
    ff = _f32(freq_factors).ctypes.data_as(C.POINTER(C.c_float)) if freq_factors is not None else None
    check(_lib.lib().b200_op_rope(y.ctypes.data_as(C.POINTER(C.c_float)), n_heads, head_dim, pos, freq_base, freq_scale, ff), "op_rope")
    return y


def op_attention(q, k_cache_f16, v_cache_f16, n_kv: int, n_head: int, n_head_kv: int, head_dim: int, scale: float,
                 round_q: bool = False) -> np.ndarray:
    q = _f32(q)
    k = np.ascontiguousarray(k_cache_f16, dtype=np.float16).view(np.uint16)
    v = np.ascontiguousarray(v_cache_f16, dtype=np.float16).view(np.uint16)
    out = np.empty(n_head * head_dim, dtype=np.float32)
    check(_lib.lib().b200_op_attention(q.ctypes.data_as(C.POINTER(C.c_float)), k.ctypes.data_as(C.POINTER(C.c_uint16)),
                                       v.ctypes.data_as(C.POINTER(C.c_uint16)), n_kv, n_head, n_head_kv, head_dim, scale, int(round_q),
                                       out.ctypes.data_as(C.POINTER(C.c_float))), "op_attention")
    return out


def set_token_kernel(on: bool) -> None:
    """contexts created afterwards decode with the persistent per-token kernel (True) or one kernel per operator (False)"""
    _lib.lib().b200_set_token_kernel(int(on))


def set_prefill_batch(on: bool) -> None:
    """prompt batches through the batched kernels (True, default) or token by token (False)"""
    _lib.lib().b200_set_prefill_batch(int(on))


def set_prefill_mma(mode: int) -> None:
    """K-quant prompt batches: 1 = mma.sync kernel (default), 2 = tcgen05 / TMEM kernel, 0 = dp4a batch kernel"""
    _lib.lib().b200_set_prefill_mma(int(mode))


def set_attention_route(route: int) -> None:
    """0: automatic (cluster kernel when it fits), 1: always the long-context three-kernel route"""
    _lib.lib().b200_set_attention_route(route)


class Sampler:
    """doInference's sampler on its own (include/booster_b200.h b200_sampler_*): Janus (janus != 0) or the standard chain"""

    def __init__(self, path: str, n_ctx: int, janus: int = 1, depth: int = 200, scale: float = 0.96, hi: float = 0.99, lo: float = 0.96,
                 temperature: float = 0.8, top_k: int = 40, top_p: float = 0.95, repetition_penalty: float = 1.0, penalty_last_n: int = 64,
                 mirostat: int = 0, mirostat_tau: float = 5.0, mirostat_eta: float = 0.1, typical_p: float = 1.0, tfs_z: float = 1.0,
                 min_p: float = 0.05):
        self.L = _lib.lib()
        self.h = self.L.b200_sampler_new(path.encode(), n_ctx, janus, depth, scale, hi, lo, temperature, top_k, top_p, repetition_penalty, penalty_last_n)
        if not self.h:
            raise B200Error(f"b200_sampler_new({path}) failed")
        self.L.b200_sampler_set_standard(self.h, mirostat, mirostat_tau, mirostat_eta, typical_p, tfs_z, min_p)

    def close(self):
        if self.h:
            self.L.b200_sampler_free(self.h)
            self.h = None

    def reset(self, prompt: Sequence[int], seed: int):
        toks = np.ascontiguousarray(prompt, dtype=np.int32)
        self.L.b200_sampler_reset(self.h, toks.ctypes.data_as(C.POINTER(C.c_int32)), len(toks), seed)

    def sample(self, logits: np.ndarray, pos: int, n_predict: int = -1) -> int:
        lg = np.ascontiguousarray(logits, dtype=np.float32).copy()
        return int(self.L.b200_sampler_sample(self.h, lg.ctypes.data_as(C.POINTER(C.c_float)), pos, n_predict))


class Tokenizer:
    """llama_tokenize / llama_token_to_piece / llama_token_is_eog of a GGUF's vocabulary (include/booster_b200.h)."""

    def __init__(self, path: str):
        self.L = _lib.lib()
        self.h = self.L.b200_tokenizer_load(path.encode())
        if not self.h:
            raise B200Error(f"b200_tokenizer_load({path}): {_lib.last_error()}")
        self.n_vocab = int(self.L.b200_tokenizer_n_vocab(self.h))

    def close(self):
        if self.h:
            self.L.b200_tokenizer_free(self.h)
            self.h = None

    def tokenize(self, text: bytes, add_special: bool = False, parse_special: bool = True) -> List[int]:
        cap = len(text) + 16
        buf = (C.c_int32 * cap)()
        n = self.L.b200_tokenize(self.h, text, len(text), buf, cap, int(add_special), int(parse_special))
        if n == -2**31:
            raise B200Error(f"b200_tokenize: {_lib.last_error()}")
        if n < 0:
            raise B200Error("token buffer too small")
        return list(buf[:n])

    def piece(self, token: int, special: bool = True) -> bytes:
        buf = C.create_string_buffer(512)
        n = self.L.b200_token_to_piece(self.h, token, buf, 512, int(special))
        if n < 0:
            raise B200Error("piece buffer too small")
        return buf.raw[:n]

    def is_eog(self, token: int) -> bool:
        return bool(self.L.b200_token_is_eog(self.h, token))

    @property
    def token_nl(self) -> int:
        """llama_token_nl: the newline token (-1 if the vocabulary has none)"""
        return int(self.L.b200_token_nl(self.h))
