"""oracle/port.py — TEST INFRASTRUCTURE, NOT PRODUCT.

ctypes driver for liboracle_port.so (oracle/oracle_port.c, the plain-C restatement of the reference CPU path).
Mirrors oracle/ref.py's interface so tests can swap one for the other. Pinned against oracle/_ref and the
golden vectors in tests/golden (see tests/test_oracle_pinned.py).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle_port.so")
_lib = None


class PortMat(C.Structure):
    _fields_ = [("type", C.c_int32), ("pad", C.c_int32), ("data", C.c_void_p), ("rows", C.c_int64), ("k", C.c_int64)]


class PortLayer(C.Structure):
    _fields_ = [(n, PortMat) for n in ("wq", "wk", "wv", "wo", "gate", "up", "down")] + \
               [("attn_norm", C.c_void_p), ("ffn_norm", C.c_void_p)]


class PortModel(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("n_layer", "n_embd", "n_head", "n_head_kv", "head_dim", "n_ff", "n_vocab", "n_ctx")] + \
               [(n, C.c_float) for n in ("rms_eps", "rope_freq_base", "rope_freq_scale", "yarn_ext_factor", "yarn_attn_factor")] + \
               [("n_ctx_orig", C.c_int32)] + \
               [("rope_freq_factors", C.c_void_p), ("tok_embd", PortMat), ("output", PortMat), ("output_norm", C.c_void_p),
                ("layers", C.POINTER(PortLayer)), ("k_cache", C.c_void_p), ("v_cache", C.c_void_p),
                ("tap_l_out", C.c_void_p), ("tap_q", C.c_void_p), ("tap_kqv", C.c_void_p)]


def build() -> None:
    subprocess.check_call(["make", "-C", _HERE, "port"], stdout=subprocess.DEVNULL)


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "oracle_port.c")
        if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
            build()
        L = C.CDLL(_SO)
        L.port_quantize_row_q8_K.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        L.port_quantize_row_q8_0.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        L.port_dequantize_row.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int64]
        L.port_mul_mat_vec.argtypes = [C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]
        L.port_rms_norm.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_void_p]
        L.port_rope.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_void_p]
        L.port_attention.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_void_p]
        L.port_decode.argtypes = [C.POINTER(PortModel), C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.port_decode.restype = C.c_int
        L.port_decode_cells.argtypes = [C.POINTER(PortModel), C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.port_decode_cells.restype = C.c_int
        L.port_k_shift.argtypes = [C.POINTER(PortModel), C.c_void_p]
        _lib = L
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def quantize_row_q8_K(x) -> np.ndarray:
    x = _f32(x)
    out = np.empty(x.size // 256 * 292, dtype=np.uint8)
    lib().port_quantize_row_q8_K(x.ctypes.data, out.ctypes.data, x.size)
    return out


def quantize_row_q8_0(x) -> np.ndarray:
    x = _f32(x)
    out = np.empty(x.size // 32 * 34, dtype=np.uint8)
    lib().port_quantize_row_q8_0(x.ctypes.data, out.ctypes.data, x.size)
    return out


def dequantize_row(t: int, raw, k: int) -> np.ndarray:
    raw = np.ascontiguousarray(raw, dtype=np.uint8)
    out = np.empty(k, dtype=np.float32)
    lib().port_dequantize_row(t, raw.ctypes.data, out.ctypes.data, k)
    return out


def mul_mat_vec(t: int, w_raw, n_rows: int, k: int, x) -> np.ndarray:
    w_raw = np.ascontiguousarray(w_raw, dtype=np.uint8)
    x = _f32(x)
    y = np.empty(n_rows, dtype=np.float32)
    lib().port_mul_mat_vec(t, w_raw.ctypes.data, n_rows, k, x.ctypes.data, y.ctypes.data)
    return y


def rms_norm(x, w, eps: float) -> np.ndarray:
    x = _f32(x)
    y = np.empty_like(x)
    wv = _f32(w) if w is not None else None
    lib().port_rms_norm(x.ctypes.data, wv.ctypes.data if wv is not None else None, x.size, eps, y.ctypes.data)
    return y


def rope(x, n_heads: int, head_dim: int, pos: int, freq_base: float, freq_scale: float = 1.0, freq_factors=None) -> np.ndarray:
    y = _f32(x).copy()
    ff = _f32(freq_factors) if freq_factors is not None else None
    lib().port_rope(y.ctypes.data, n_heads, head_dim, pos, freq_base, freq_scale, ff.ctypes.data if ff is not None else None)
    return y


def attention(q, k_cache_f16, v_cache_f16, n_kv, n_head, n_head_kv, head_dim, scale, round_q=False) -> np.ndarray:
    """the default attention route in the reference's exact operation order (oracle_port.c port_attention)"""
    q = _f32(q)
    k = np.ascontiguousarray(k_cache_f16, dtype=np.float16).view(np.uint16)
    v = np.ascontiguousarray(v_cache_f16, dtype=np.float16).view(np.uint16)
    out = np.empty(n_head * head_dim, dtype=np.float32)
    lib().port_attention(q.ctypes.data, k.ctypes.data, v.ctypes.data, n_kv, n_head, n_head_kv, head_dim, scale, int(round_q), out.ctypes.data)
    return out


class PortModelRunner:
    """Same interface as oracle.ref.RefModel (decode / greedy / kv_clear), running oracle_port.c."""

    def __init__(self, path: str, n_ctx: int = 512):
        from booster_b200 import gguf_io as G   # container parsing only (no compute)
        self.L = lib()
        self.f = G.read_gguf(path)
        kv = self.f.kv
        self._keep = []
        M = PortModel()
        M.n_layer = int(kv["llama.block_count"]); M.n_embd = int(kv["llama.embedding_length"])
        M.n_head = int(kv["llama.attention.head_count"]); M.n_head_kv = int(kv.get("llama.attention.head_count_kv", M.n_head))
        M.head_dim = M.n_embd // M.n_head; M.n_ff = int(kv["llama.feed_forward_length"])
        M.n_ctx = (n_ctx + 31) // 32 * 32
        M.rms_eps = float(kv.get("llama.attention.layer_norm_rms_epsilon", 1e-5))
        M.rope_freq_base = float(kv.get("llama.rope.freq_base", 10000.0))
        sf = float(kv.get("llama.rope.scaling.factor", 0.0))
        M.rope_freq_scale = 1.0 if sf == 0.0 or kv.get("llama.rope.scaling.type", "linear") == "none" else 1.0 / sf
        # cparams.yarn_ext_factor / yarn_attn_factor / n_ctx_orig_yarn (cpp/src/llama.cpp:16670-16690)
        M.yarn_ext_factor = 1.0 if kv.get("llama.rope.scaling.type", "linear") == "yarn" else 0.0
        M.yarn_attn_factor = float(kv.get("llama.rope.scaling.attn_factor", 1.0))
        M.n_ctx_orig = int(kv.get("llama.rope.scaling.original_context_length", kv.get("llama.context_length", 4096)))

        def mat(name) -> PortMat:
            t = self.f.tensors[name]
            arr = np.ascontiguousarray(t.data)
            self._keep.append(arr)
            return PortMat(t.type, 0, arr.ctypes.data, t.ne[1], t.ne[0])

        def vec(name):
            arr = np.ascontiguousarray(self.f.tensors[name].data).view(np.float32)
            self._keep.append(arr)
            return arr.ctypes.data

        M.tok_embd = mat("token_embd.weight")
        M.n_vocab = int(self.f.tensors["token_embd.weight"].ne[1])
        M.output = mat("output.weight" if "output.weight" in self.f.tensors else "token_embd.weight")
        M.output_norm = vec("output_norm.weight")
        if "rope_freqs.weight" in self.f.tensors:
            M.rope_freq_factors = vec("rope_freqs.weight")
        layers = (PortLayer * M.n_layer)()
        for i in range(M.n_layer):
            p = f"blk.{i}."
            for fld, nm in (("wq", "attn_q"), ("wk", "attn_k"), ("wv", "attn_v"), ("wo", "attn_output"),
                            ("gate", "ffn_gate"), ("up", "ffn_up"), ("down", "ffn_down")):
                setattr(layers[i], fld, mat(p + nm + ".weight"))
            layers[i].attn_norm = vec(p + "attn_norm.weight")
            layers[i].ffn_norm = vec(p + "ffn_norm.weight")
        self._layers = layers
        M.layers = C.cast(layers, C.POINTER(PortLayer))
        kvd = M.n_head_kv * M.head_dim
        self.kc = np.zeros((M.n_layer, M.n_ctx, kvd), dtype=np.uint16)
        self.vc = np.zeros_like(self.kc)
        M.k_cache, M.v_cache = self.kc.ctypes.data, self.vc.ctypes.data
        self.tap_l_out = np.zeros((M.n_layer, M.n_embd), dtype=np.float32)
        self.tap_q = np.zeros((M.n_layer, M.n_head * M.head_dim), dtype=np.float32)
        self.tap_kqv = np.zeros_like(self.tap_q)
        M.tap_l_out, M.tap_q, M.tap_kqv = self.tap_l_out.ctypes.data, self.tap_q.ctypes.data, self.tap_kqv.ctypes.data
        self.M = M
        self.n_vocab, self.n_ctx = M.n_vocab, M.n_ctx
        self._cells_reset()

    def close(self):
        pass

    def kv_clear(self):
        """llama_kv_cache_clear (cpp/src/llama.cpp:3135-3152)"""
        self.kc[:] = 0
        self.vc[:] = 0
        self._cells_reset()

    # ---- struct llama_kv_cache's cell bookkeeping (cpp/src/llama.cpp:2495-2539), host side. Until the first seq_rm / seq_add /
    # seq_div a sequence only grows and cell == position; afterwards cells are placed by find_slot, carry a position and a pending
    # K-shift delta, and attention masks by position.
    def _cells_reset(self):
        self.managed = False
        self.n_hi = 0                                   # identity mode: highest position written + 1
        self.cell_pos = np.full(self.n_ctx, -1, dtype=np.int32)
        self.cell_delta = np.zeros(self.n_ctx, dtype=np.int32)
        self.head, self.used, self.has_shift = 0, 0, False

    def _enter_managed(self):
        if self.managed:
            return
        self.managed = True
        n = min(self.n_hi, self.n_ctx)
        self.cell_pos[:n] = np.arange(n, dtype=np.int32)
        self.used = n
        self.head = 0 if n >= self.n_ctx else n         # cpp/src/llama.cpp:14821-14826: head += n_tokens, wrapped

    def kv_seq_rm(self, p0: int, p1: int):
        """llama_kv_cache_seq_rm(ctx, 0, p0, p1) (cpp/src/llama.cpp:3154-3206)"""
        self._enter_managed()
        p0 = max(p0, 0)
        p1 = 2 ** 31 - 1 if p1 < 0 else p1
        new_head = self.n_ctx
        for i in range(self.n_ctx):
            if p0 <= self.cell_pos[i] < p1:
                self.used -= 1
                self.cell_pos[i] = -1
                if new_head == self.n_ctx:
                    new_head = i
        if new_head != self.n_ctx and new_head < self.head:
            self.head = new_head

    def kv_seq_add(self, p0: int, p1: int, delta: int):
        """llama_kv_cache_seq_add(ctx, 0, p0, p1, delta) (cpp/src/llama.cpp:3268-3314)"""
        self._enter_managed()
        p0 = max(p0, 0)
        p1 = 2 ** 31 - 1 if p1 < 0 else p1
        if p0 == p1:
            return
        new_head = self.n_ctx
        for i in range(self.n_ctx):
            if p0 <= self.cell_pos[i] < p1:
                self.has_shift = True
                self.cell_pos[i] += delta
                self.cell_delta[i] += delta
                if self.cell_pos[i] < 0:
                    self.used -= 1
                    self.cell_pos[i] = -1
                    if new_head == self.n_ctx:
                        new_head = i
        self.head = new_head if new_head != self.n_ctx else 0

    def kv_seq_div(self, p0: int, p1: int, d: int):
        """llama_kv_cache_seq_div(ctx, 0, p0, p1, d) (cpp/src/llama.cpp:3316-3349)"""
        self._enter_managed()
        p0 = max(p0, 0)
        p1 = 2 ** 31 - 1 if p1 < 0 else p1
        if p0 == p1:
            return
        for i in range(self.n_ctx):
            if p0 <= self.cell_pos[i] < p1:
                self.has_shift = True
                old = int(self.cell_pos[i])
                self.cell_pos[i] = old // d
                self.cell_delta[i] += self.cell_pos[i] - old

    def _find_slot(self, n: int) -> int:
        """llama_kv_cache_find_slot for n tokens of one sequence (cpp/src/llama.cpp:3028-3125, 14684-14688)"""
        size = self.n_ctx
        if n > size:
            return -1
        if self.head > self.used + 2 * n:
            self.head = 0
        n_tested = 0
        while True:
            if self.head + n > size:
                n_tested += size - self.head
                self.head = 0
                continue
            found = True
            for i in range(n):
                if self.cell_pos[self.head + i] >= 0:
                    found = False
                    self.head += i + 1
                    n_tested += i + 1
                    break
            if found:
                return self.head
            if n_tested >= size:
                return -1

    def decode(self, tokens: Sequence[int], pos0: int) -> np.ndarray:
        toks = np.ascontiguousarray(tokens, dtype=np.int32)
        out = np.empty(self.n_vocab, dtype=np.float32)
        if not self.managed:
            rc = self.L.port_decode(C.byref(self.M), toks.ctypes.data, len(toks), pos0, out.ctypes.data)
            if rc != 0:
                raise RuntimeError(f"port_decode rc={rc}")
            self.n_hi = max(self.n_hi, pos0 + len(toks))
            return out
        if self.has_shift:                              # llama_kv_cache_update: the pending K-shift, then the deltas are cleared
            self.L.port_k_shift(C.byref(self.M), self.cell_delta.ctypes.data)
            self.cell_delta[:] = 0
            self.has_shift = False
        n = len(toks)
        first = self._find_slot(n)
        if first < 0:
            raise RuntimeError("no free KV cells for the batch")     # llama_decode returns 1 (cpp/src/llama.cpp:14690)
        self.cell_pos[first:first + n] = np.arange(pos0, pos0 + n, dtype=np.int32)
        self.used += n
        n_kv = int(np.max(np.nonzero(self.cell_pos >= 0)[0])) + 1   # llama_kv_cache_cell_max
        cell_of = np.arange(first, first + n, dtype=np.int32)
        rc = self.L.port_decode_cells(C.byref(self.M), toks.ctypes.data, n, pos0, cell_of.ctypes.data, self.cell_pos.ctypes.data,
                                      n_kv, out.ctypes.data)
        if rc != 0:
            raise RuntimeError(f"port_decode_cells rc={rc}")
        self.head = first + n                           # cpp/src/llama.cpp:14821-14826 (after the graph ran)
        if self.head >= self.n_ctx:
            self.head = 0
        return out

    def greedy(self, prompt: Sequence[int], n_gen: int) -> (List[int], List[np.ndarray]):
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
