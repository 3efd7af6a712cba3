"""oracle/ref.py — TEST INFRASTRUCTURE, NOT PRODUCT.

ctypes driver for oracle/_ref/libbooster_cpu_ref_<variant>.so: the UNMODIFIED reference CPU path
(cpp/ggml + cpp/src/llama.cpp + cpp/bridge.cpp, built by oracle/Makefile from /root/reference where it lies)
plus the thin veneer oracle/ref_shim.cpp. Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may
import this module; the product (booster_b200/) never does.

Variant choice: "native" was compiled with -march=native in the build container (Sapphire Rapids: AVX-512,
AMX, AVX512-FP16), as Booster's own CPU build would be (booster_cpu.go:4); "v3" is an x86-64-v3 build.
If the host CPU lacks a flag the native build needs we fall back to v3 (the K-quant dot products are AVX2
in both; only ggml.c's f32 vector width and tinyBLAS tile width differ).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF_DIR = os.path.join(_HERE, "_ref")
_NATIVE_NEEDS = {"avx512f", "avx512bw", "avx512vl", "avx512dq", "avx512_vnni", "avx512_bf16", "avx512_fp16",
                 "amx_tile", "amx_int8", "amx_bf16", "avx_vnni", "avx512vbmi", "avx512_vbmi2", "avx512ifma",
                 "avx512_bitalg", "avx512_vpopcntdq", "avx2", "fma", "f16c", "bmi2"}

_lib = None
_variant = None


def _cpu_flags() -> set:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return set(line.split(":", 1)[1].split())
    except OSError:
        pass
    return set()


def available() -> bool:
    return any(os.path.exists(os.path.join(_REF_DIR, f"libbooster_cpu_ref_{v}.so")) for v in ("native", "v3"))


def variant() -> Optional[str]:
    lib()
    return _variant


def lib() -> C.CDLL:
    global _lib, _variant
    if _lib is not None:
        return _lib
    flags = _cpu_flags()
    order = ["native", "v3"] if _NATIVE_NEEDS <= flags else ["v3"]
    if os.environ.get("BOOSTER_REF_VARIANT"):
        order = [os.environ["BOOSTER_REF_VARIANT"]]
    last = None
    for v in order:
        p = os.path.join(_REF_DIR, f"libbooster_cpu_ref_{v}.so")
        if not os.path.exists(p):
            continue
        try:
            L = C.CDLL(p, mode=C.RTLD_LOCAL)
        except OSError as e:  # pragma: no cover
            last = e
            continue
        _lib, _variant = L, v
        break
    if _lib is None:
        raise RuntimeError(f"reference CPU library not found/loaded under {_REF_DIR} ({last}); run `make -C oracle ref`")
    L = _lib
    L.refshim_init.argtypes = [C.c_int]
    L.refshim_quantize.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int]
    L.refshim_quantize.restype = C.c_int
    L.refshim_load.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int]
    L.refshim_load.restype = C.c_void_p
    L.refshim_free.argtypes = [C.c_void_p]
    L.refshim_n_vocab.argtypes = [C.c_void_p]
    L.refshim_n_ctx.argtypes = [C.c_void_p]
    L.refshim_kv_clear.argtypes = [C.c_void_p]
    L.refshim_decode.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.c_int, C.c_int, C.POINTER(C.c_float)]
    L.refshim_decode.restype = C.c_int
    L.refshim_set_taps.argtypes = [C.c_void_p, C.c_char_p]
    L.refshim_get_tap.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_float), C.c_int64, C.POINTER(C.c_int64)]
    L.refshim_get_tap.restype = C.c_int64
    L.refshim_reset_timings.argtypes = [C.c_void_p]
    L.refshim_timings.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_int)]
    # raw ggml entry points used by operator-level tests (exported C symbols of the reference itself)
    for name in ("quantize_row_q8_K", "quantize_row_q8_0"):
        fn = getattr(L, name)
        fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        fn.restype = None
    for name in ("dequantize_row_q4_K", "dequantize_row_q5_K", "dequantize_row_q6_K", "dequantize_row_q8_0"):
        fn = getattr(L, name)
        fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        fn.restype = None
    for name in ("ggml_vec_dot_q4_K_q8_K", "ggml_vec_dot_q5_K_q8_K", "ggml_vec_dot_q6_K_q8_K", "ggml_vec_dot_q8_0_q8_0"):
        fn = getattr(L, name)
        fn.argtypes = [C.c_int, C.POINTER(C.c_float), C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int]
        fn.restype = None
    L.ggml_quantize_chunk.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p]
    L.ggml_quantize_chunk.restype = C.c_size_t
    # tokenizer oracle (may be absent from a library built before the tokenizer shim existed)
    if hasattr(L, "refshim_vocab_load"):
        L.refshim_vocab_load.argtypes = [C.c_char_p]
        L.refshim_vocab_load.restype = C.c_void_p
        L.refshim_vocab_free.argtypes = [C.c_void_p]
        L.refshim_tokenize.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(C.c_int32), C.c_int, C.c_int, C.c_int]
        L.refshim_tokenize.restype = C.c_int
        L.refshim_token_to_piece.argtypes = [C.c_void_p, C.c_int32, C.c_char_p, C.c_int, C.c_int]
        L.refshim_token_to_piece.restype = C.c_int
        L.refshim_token_is_eog.argtypes = [C.c_void_p, C.c_int32]
        L.refshim_token_is_eog.restype = C.c_int
        if hasattr(L, "refshim_vocab_token_nl"):
            L.refshim_vocab_token_nl.argtypes = [C.c_void_p]
            L.refshim_vocab_token_nl.restype = C.c_int
        L.refshim_cpt_flags.argtypes = [C.c_uint32]
        L.refshim_cpt_flags.restype = C.c_uint16
    if hasattr(L, "refshim_janus_generate"):
        L.refshim_kv_seq_rm.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.refshim_kv_seq_add.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        if hasattr(L, "refshim_kv_seq_div"):
            L.refshim_kv_seq_div.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.refshim_janus_generate.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float,
                                             C.c_uint32, C.c_int, C.POINTER(C.c_int32)]
        L.refshim_janus_generate.restype = C.c_int
    if hasattr(L, "refshim_standard_generate"):
        L.refshim_standard_generate.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.c_int, C.c_int, C.c_int, C.c_float, C.c_float,
                                                C.c_float, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int,
                                                C.c_uint32, C.POINTER(C.c_int32)]
        L.refshim_standard_generate.restype = C.c_int
        L.refshim_token_nl.argtypes = [C.c_void_p]
        L.refshim_token_nl.restype = C.c_int
    L.refshim_init(1)
    return L


def has_tokenizer() -> bool:
    return available() and hasattr(lib(), "refshim_vocab_load")


class RefVocab:
    """the reference's tokenizer on a vocab-only load of a GGUF: llama_tokenize / llama_token_to_piece /
    llama_token_is_eog (cpp/src/llama-vocab.cpp), the calls cpp/bridge.cpp:275-278, 630, 640 make"""

    def __init__(self, path: str):
        self.L = lib()
        self.h = self.L.refshim_vocab_load(path.encode())
        if not self.h:
            raise RuntimeError(f"reference could not load the vocabulary of {path}")

    def close(self):
        if self.h:
            self.L.refshim_vocab_free(self.h)
            self.h = None

    def tokenize(self, text: bytes, add_special: bool = False, parse_special: bool = True) -> List[int]:
        cap = len(text) + 16
        buf = (C.c_int32 * cap)()
        n = self.L.refshim_tokenize(self.h, text, len(text), buf, cap, int(add_special), int(parse_special))
        if n < 0:
            raise RuntimeError("token buffer too small")
        return list(buf[:n])

    def piece(self, token: int, special: bool = True) -> bytes:
        buf = C.create_string_buffer(512)
        n = self.L.refshim_token_to_piece(self.h, token, buf, 512, int(special))
        if n < 0:
            raise RuntimeError("piece buffer too small")
        return buf.raw[:n]

    def is_eog(self, token: int) -> bool:
        return bool(self.L.refshim_token_is_eog(self.h, token))

    def token_nl(self) -> int:
        """llama_token_nl(model)"""
        return int(self.L.refshim_vocab_token_nl(self.h))


def cpt_flags(cp: int) -> int:
    """codepoint_flags of the reference's Unicode tables (cpp/src/unicode.h:8-46) as the raw uint16"""
    return int(lib().refshim_cpt_flags(cp))


def quantize_model(f_in: str, f_out: str, ftype: int, nthread: int = 0) -> None:
    """llama_model_quantize (cpp/src/llama.cpp:15435-) — ftype 7 Q8_0, 15 Q4_K_M, 17 Q5_K_M."""
    rc = lib().refshim_quantize(f_in.encode(), f_out.encode(), ftype, nthread or (os.cpu_count() or 1))
    if rc != 0:
        raise RuntimeError(f"llama_model_quantize failed rc={rc}")


class RefModel:
    """The reference's llama_model + llama_context on the CPU (n_gpu_layers = 0, flash_attn = false by default:
    Booster's defaults, cpp/common/common.h:175)."""

    def __init__(self, path: str, n_ctx: int = 512, n_batch: int = 512, n_threads: Optional[int] = None, flash_attn: bool = False):
        self.L = lib()
        self.n_threads = n_threads or (os.cpu_count() or 1)
        self.h = self.L.refshim_load(path.encode(), n_ctx, n_batch, self.n_threads, int(flash_attn))
        if not self.h:
            raise RuntimeError(f"reference failed to load {path}")
        self.n_vocab = self.L.refshim_n_vocab(self.h)
        self.n_ctx = self.L.refshim_n_ctx(self.h)

    def close(self):
        if self.h:
            self.L.refshim_free(self.h)
            self.h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def kv_clear(self):
        self.L.refshim_kv_clear(self.h)

    def decode(self, tokens: Sequence[int], pos0: int) -> np.ndarray:
        toks = np.ascontiguousarray(tokens, dtype=np.int32)
        out = np.empty(self.n_vocab, dtype=np.float32)
        rc = self.L.refshim_decode(self.h, toks.ctypes.data_as(C.POINTER(C.c_int32)), len(toks), pos0,
                                   out.ctypes.data_as(C.POINTER(C.c_float)))
        if rc != 0:
            raise RuntimeError(f"llama_decode rc={rc}")
        return out

    def set_taps(self, names: List[str]):
        self.L.refshim_set_taps(self.h, ",".join(names).encode())

    def get_tap(self, name: str) -> Optional[np.ndarray]:
        shape = (C.c_int64 * 4)()
        n = self.L.refshim_get_tap(self.h, name.encode(), None, 0, shape)
        if n == 0:
            return None
        out = np.empty(n, dtype=np.float32)
        self.L.refshim_get_tap(self.h, name.encode(), out.ctypes.data_as(C.POINTER(C.c_float)), n, shape)
        return out

    def greedy(self, prompt: Sequence[int], n_gen: int) -> (List[int], List[np.ndarray]):
        """prefill `prompt` in one batch, then n_gen greedy (arg-max) steps; returns ids and per-step logits."""
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

    def kv_seq_rm(self, p0: int, p1: int):
        """llama_kv_cache_seq_rm(ctx, 0, p0, p1) (cpp/bridge.cpp:500)"""
        self.L.refshim_kv_seq_rm(self.h, p0, p1)

    def kv_seq_add(self, p0: int, p1: int, delta: int):
        """llama_kv_cache_seq_add(ctx, 0, p0, p1, delta) (cpp/bridge.cpp:501); the K-shift runs inside the next decode"""
        self.L.refshim_kv_seq_add(self.h, p0, p1, delta)

    def kv_seq_div(self, p0: int, p1: int, d: int):
        """llama_kv_cache_seq_div(ctx, 0, p0, p1, d) (cpp/bridge.cpp:518, Self-Extend)"""
        self.L.refshim_kv_seq_div(self.h, p0, p1, d)

    def janus_generate(self, prompt: Sequence[int], n_gen: int, depth: int, scale: float, hi: float, lo: float, seed: int,
                       n_predict: int = -1) -> List[int]:
        """the bridge's generation loop with the reference's own initJanus / sample_janus_token (cpp/janus.cpp)"""
        toks = np.ascontiguousarray(prompt, dtype=np.int32)
        out = np.empty(n_gen, dtype=np.int32)
        n = self.L.refshim_janus_generate(self.h, toks.ctypes.data_as(C.POINTER(C.c_int32)), len(toks), n_gen, depth, scale, hi, lo,
                                          seed, n_predict, out.ctypes.data_as(C.POINTER(C.c_int32)))
        if n < 0:
            raise RuntimeError("llama_decode failed")
        return out[:n].tolist()

    def standard_generate(self, prompt: Sequence[int], n_gen: int, seed: int, *, mirostat: int = 0, mirostat_tau: float = 5.0,
                          mirostat_eta: float = 0.1, temp: float = 0.8, top_k: int = 40, top_p: float = 0.95, typical_p: float = 1.0,
                          tfs_z: float = 1.0, min_p: float = 0.05, penalty_repeat: float = 1.0, penalty_last_n: int = 64) -> List[int]:
        """the bridge's generation loop with the standard chain the reference keeps commented out (cpp/bridge.cpp:598):
        llama_sampling_init / llama_sampling_sample / llama_sampling_accept of common/sampling.cpp"""
        toks = np.ascontiguousarray(prompt, dtype=np.int32)
        out = np.empty(n_gen, dtype=np.int32)
        n = self.L.refshim_standard_generate(self.h, toks.ctypes.data_as(C.POINTER(C.c_int32)), len(toks), n_gen, mirostat, mirostat_tau,
                                             mirostat_eta, temp, top_k, top_p, typical_p, tfs_z, min_p, penalty_repeat, penalty_last_n,
                                             seed, out.ctypes.data_as(C.POINTER(C.c_int32)))
        if n < 0:
            raise RuntimeError("llama_sampling_init / llama_decode failed")
        return out[:n].tolist()

    def token_nl(self) -> int:
        return int(self.L.refshim_token_nl(self.h))

    def reset_timings(self):
        self.L.refshim_reset_timings(self.h)

    def timings(self) -> Dict[str, float]:
        tp, te = C.c_double(), C.c_double()
        np_, ne = C.c_int(), C.c_int()
        self.L.refshim_timings(self.h, C.byref(tp), C.byref(np_), C.byref(te), C.byref(ne))
        return {"t_p_eval_ms": tp.value, "n_p_eval": np_.value, "t_eval_ms": te.value, "n_eval": ne.value}


# ---- raw operator access (block layouts: cpp/ggml/src/ggml-common.h:186-316) ---------------------------------
GGML_TYPE = {"Q8_0": 8, "Q4_K": 12, "Q5_K": 13, "Q6_K": 14}
_ROW = {8: (32, 34), 12: (256, 144), 13: (256, 176), 14: (256, 210)}


def quantize_weights(x: np.ndarray, t: int) -> np.ndarray:
    """ggml_quantize_chunk: the reference's weight quantizer (quantize_q4_K etc.), rows x k f32 -> raw blocks."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    rows, k = x.shape
    be, bb = _ROW[t]
    out = np.empty(rows * (k // be) * bb, dtype=np.uint8)
    n = lib().ggml_quantize_chunk(t, x.ctypes.data, out.ctypes.data, 0, rows, k, None)
    assert n == out.size
    return out


def quantize_row_q8_K(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty(x.size // 256 * 292, dtype=np.uint8)
    lib().quantize_row_q8_K(x.ctypes.data, out.ctypes.data, x.size)
    return out


def quantize_row_q8_0(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty(x.size // 32 * 34, dtype=np.uint8)
    lib().quantize_row_q8_0(x.ctypes.data, out.ctypes.data, x.size)
    return out


def dequantize_row(t: int, raw: np.ndarray, k: int) -> np.ndarray:
    fn = {8: "dequantize_row_q8_0", 12: "dequantize_row_q4_K", 13: "dequantize_row_q5_K", 14: "dequantize_row_q6_K"}[t]
    raw = np.ascontiguousarray(raw, dtype=np.uint8)
    out = np.empty(k, dtype=np.float32)
    getattr(lib(), fn)(raw.ctypes.data, out.ctypes.data, k)
    return out


def mul_mat_vec(t: int, w_raw: np.ndarray, n_rows: int, k: int, x: np.ndarray) -> np.ndarray:
    """The reference's batch-1 quantized mat-vec: quantize x to the type's vec_dot_type, then one
    ggml_vec_dot_* per row (ggml_compute_forward_mul_mat, cpp/ggml/src/ggml.c:12345-12373, 12186-12275)."""
    L = lib()
    w_raw = np.ascontiguousarray(w_raw, dtype=np.uint8)
    be, bb = _ROW[t]
    rb = k // be * bb
    if t == 8:
        xq, dot = quantize_row_q8_0(x), L.ggml_vec_dot_q8_0_q8_0
    else:
        xq = quantize_row_q8_K(x)
        dot = {12: L.ggml_vec_dot_q4_K_q8_K, 13: L.ggml_vec_dot_q5_K_q8_K, 14: L.ggml_vec_dot_q6_K_q8_K}[t]
    y = np.empty(n_rows, dtype=np.float32)
    s = C.c_float()
    for r in range(n_rows):
        dot(k, C.byref(s), 0, w_raw.ctypes.data + r * rb, 0, xq.ctypes.data, 0, 1)
        y[r] = s.value
    return y
