"""Minimal GGUF v3 writer/reader (numpy) and synthetic LLaMA-shaped model generator.

No model file exists offline, so tests and bench fabricate GGUFs with the exact
hyper-parameters of the BASELINE.json configs. Two sources of weights:

* ``synth_llama(..., source="blocks")`` writes random *quantized blocks* directly (valid fp16 scales,
  random packed quants) following the reference's tensor-type mixture for the file type
  (``llama_tensor_get_type``, cpp/src/llama.cpp:15435-15620) — fast enough for the 4.9 GB 8B file.
* ``synth_llama(..., source="f32")`` writes F32 tensors ~ N(0, 0.02²) to be quantized by the reference's
  own ``llama_model_quantize`` through oracle/_ref (bit-faithful quantized files; small models).

Container layout: cpp/ggml/src/ggml.c:20767-20788 (header / kv / tensor infos), data aligned to
general.alignment = 32 (cpp/ggml/src/ggml.c:21100-21104).
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import numpy as np

GGUF_MAGIC = b"GGUF"
ALIGN = 32

# ggml_type ids (cpp/ggml/include/ggml.h:360-375) -> (block elements, block bytes)
F32, F16, Q8_0, Q4_K, Q5_K, Q6_K = 0, 1, 8, 12, 13, 14
BLOCK = {F32: (1, 4), F16: (1, 2), Q8_0: (32, 34), Q4_K: (256, 144), Q5_K: (256, 176), Q6_K: (256, 210)}
TYPE_NAME = {F32: "F32", F16: "F16", Q8_0: "Q8_0", Q4_K: "Q4_K", Q5_K: "Q5_K", Q6_K: "Q6_K"}

# llama_ftype (cpp/include/llama.h:133-165)
FTYPE = {"F32": 0, "Q8_0": 7, "Q4_K_M": 15, "Q5_K_M": 17, "Q6_K": 18}

_GV = {"u8": 0, "i8": 1, "u16": 2, "i16": 3, "u32": 4, "i32": 5, "f32": 6, "bool": 7, "str": 8, "arr": 9,
       "u64": 10, "i64": 11, "f64": 12}
_FMT = {0: "<B", 1: "<b", 2: "<H", 3: "<h", 4: "<I", 5: "<i", 6: "<f", 7: "<B", 10: "<Q", 11: "<q", 12: "<d"}


def row_bytes(t: int, k: int) -> int:
    be, bb = BLOCK[t]
    assert k % be == 0, f"k={k} not a multiple of block {be}"
    return k // be * bb


def _s(b: bytes) -> bytes:
    return struct.pack("<Q", len(b)) + b


def _kv_bytes(key: str, val) -> bytes:
    """val = (kind, value) with kind in _GV, or ('arr', (elem_kind, list))."""
    kind, v = val
    out = _s(key.encode()) + struct.pack("<I", _GV[kind])
    if kind == "str":
        out += _s(v.encode() if isinstance(v, str) else v)
    elif kind == "arr":
        ek, items = v
        out += struct.pack("<IQ", _GV[ek], len(items))
        if ek == "str":
            out += b"".join(_s(x.encode() if isinstance(x, str) else x) for x in items)
        else:
            out += b"".join(struct.pack(_FMT[_GV[ek]], x) for x in items)
    else:
        out += struct.pack(_FMT[_GV[kind]], v)
    return out


def write_gguf(path: str, kv: Dict[str, tuple], tensors: List[Tuple[str, Tuple[int, ...], int, object]]) -> None:
    """tensors: (name, ne (ggml order: ne[0] = innermost), ggml type, data) where data is a numpy array whose
    raw bytes are the tensor in ggml layout, or a zero-arg callable returning such an array (lazy, to
    bound memory for multi-GB files)."""
    infos = []
    off = 0
    sizes = []
    where = {}
    for name, ne, t, data in tensors:
        n_rows = int(np.prod(ne[1:])) if len(ne) > 1 else 1
        nbytes = row_bytes(t, ne[0]) * n_rows
        sizes.append(nbytes)
        if isinstance(data, tuple) and data[0] == "alias":
            # the tensor shares the file bytes of an earlier tensor of the same type and shape (synthetic benchmark
            # files: every consumer still gets its own copy in HBM, the FILE just does not store it twice)
            o, nb0 = where[data[1]]
            assert nb0 == nbytes, f"{name}: alias of {data[1]} with a different size"
            infos.append((name, ne, t, o))
            continue
        where[name] = (off, nbytes)
        infos.append((name, ne, t, off))
        off += (nbytes + ALIGN - 1) // ALIGN * ALIGN
    with open(path, "wb") as f:
        f.write(GGUF_MAGIC + struct.pack("<IQQ", 3, len(tensors), len(kv)))
        for k, v in kv.items():
            f.write(_kv_bytes(k, v))
        for name, ne, t, o in infos:
            f.write(_s(name.encode()) + struct.pack("<I", len(ne)) + b"".join(struct.pack("<Q", d) for d in ne)
                    + struct.pack("<IQ", t, o))
        pos = f.tell()
        f.write(b"\0" * ((pos + ALIGN - 1) // ALIGN * ALIGN - pos))
        for (name, ne, t, data), nbytes in zip(tensors, sizes):
            if isinstance(data, tuple) and data[0] == "alias":
                continue
            arr = data() if callable(data) else data
            raw = np.ascontiguousarray(arr).view(np.uint8).reshape(-1)
            assert raw.size == nbytes, f"{name}: have {raw.size} bytes, expected {nbytes}"
            raw.tofile(f)
            pad = (nbytes + ALIGN - 1) // ALIGN * ALIGN - nbytes
            if pad:
                f.write(b"\0" * pad)


@dataclass
class GGUFTensor:
    name: str
    ne: Tuple[int, ...]
    type: int
    data: np.ndarray  # uint8 view (memmap) of the raw bytes


@dataclass
class GGUFFile:
    kv: Dict[str, object] = field(default_factory=dict)
    tensors: Dict[str, GGUFTensor] = field(default_factory=dict)


def read_gguf(path: str) -> GGUFFile:
    mm = np.memmap(path, dtype=np.uint8, mode="r")
    buf = memoryview(mm)
    p = 0

    def rd(fmt):
        nonlocal p
        v = struct.unpack_from(fmt, buf, p)
        p += struct.calcsize(fmt)
        return v if len(v) > 1 else v[0]

    def rs():
        nonlocal p
        n = rd("<Q")
        s = bytes(buf[p:p + n])
        p += n
        return s

    assert bytes(buf[:4]) == GGUF_MAGIC, "not a GGUF file"
    p = 4
    version, n_t, n_kv = rd("<IQQ")
    assert version in (2, 3)
    out = GGUFFile()
    for _ in range(n_kv):
        key = rs().decode()
        t = rd("<I")
        if t == 8:
            out.kv[key] = rs().decode(errors="replace")
        elif t == 9:
            et, n = rd("<IQ")
            if et == 8:
                out.kv[key] = [rs() for _ in range(n)]
            else:
                out.kv[key] = [rd(_FMT[et]) for _ in range(n)]
        else:
            out.kv[key] = rd(_FMT[t])
    infos = []
    for _ in range(n_t):
        name = rs().decode()
        nd = rd("<I")
        ne = tuple(rd("<Q") for _ in range(nd))
        t, o = rd("<IQ")
        infos.append((name, ne, t, o))
    align = int(out.kv.get("general.alignment", ALIGN))
    data0 = (p + align - 1) // align * align
    for name, ne, t, o in infos:
        n_rows = int(np.prod(ne[1:])) if len(ne) > 1 else 1
        nbytes = row_bytes(t, ne[0]) * n_rows if t in BLOCK else 0
        out.tensors[name] = GGUFTensor(name, ne, t, mm[data0 + o: data0 + o + nbytes])
    return out


# ----------------------------------------------------------------------------------------------------------
# synthetic models
# ----------------------------------------------------------------------------------------------------------
@dataclass
class LlamaConfig:
    n_layer: int = 32
    n_embd: int = 4096
    n_head: int = 32
    n_head_kv: int = 8
    n_ff: int = 14336
    n_vocab: int = 128256
    n_ctx_train: int = 8192
    rope_freq_base: float = 500000.0
    rms_eps: float = 1e-5
    name: str = "synthetic-llama"

    @property
    def head_dim(self) -> int:
        return self.n_embd // self.n_head


CONFIGS = {
    # BASELINE.json configs (SURVEY.md §8 model constants)
    "llama3-8b": LlamaConfig(),
    "mistral-7b": LlamaConfig(n_vocab=32000, rope_freq_base=10000.0, n_ctx_train=32768, name="synthetic-mistral-7b"),
    "llama3-70b": LlamaConfig(n_layer=80, n_embd=8192, n_head=64, n_head_kv=8, n_ff=28672, name="synthetic-llama3-70b"),
    # parity twins: same per-layer shapes, few layers / small vocab so that the CPU oracle runs in seconds
    "llama3-8b-2l": LlamaConfig(n_layer=2, n_vocab=4096, name="synthetic-llama3-8b-2layer"),
    "llama3-70b-1l": LlamaConfig(n_layer=1, n_embd=8192, n_head=64, n_head_kv=8, n_ff=28672, n_vocab=2048,
                                 name="synthetic-llama3-70b-1layer"),
    "tiny": LlamaConfig(n_layer=2, n_embd=256, n_head=2, n_head_kv=1, n_ff=512, n_vocab=320, n_ctx_train=512,
                        rope_freq_base=10000.0, name="synthetic-tiny"),
    "tiny-gqa4": LlamaConfig(n_layer=3, n_embd=512, n_head=4, n_head_kv=1, n_ff=768, n_vocab=512, n_ctx_train=512,
                             name="synthetic-tiny-gqa4"),
}


def _use_more_bits(i: int, n: int) -> bool:
    # cpp/src/llama.cpp:15442-15444
    return i < n // 8 or i >= 7 * n // 8 or (i - n // 8) % 3 == 2


def tensor_types(cfg: LlamaConfig, ftype: str) -> Dict[str, int]:
    """Per-tensor block type chosen by the reference's quantizer for `ftype` on an LLM_ARCH_LLAMA model
    (cpp/src/llama.cpp:15435-15620: output -> Q6_K, attn_v / ffn_down -> Q6_K where use_more_bits, 70B attn_v
    Q4_K -> Q5_K; Q8_0 keeps everything Q8_0)."""
    base = {"Q4_K_M": Q4_K, "Q5_K_M": Q5_K, "Q8_0": Q8_0, "Q6_K": Q6_K}[ftype]
    kq = ftype in ("Q4_K_M", "Q5_K_M")
    is70 = cfg.n_layer == 80  # llama.cpp model type heuristic: MODEL_70B <=> n_layer == 80
    out = {"token_embd.weight": base, "output.weight": Q8_0 if base == Q8_0 else Q6_K}
    for i in range(cfg.n_layer):
        p = f"blk.{i}."
        v = base
        if kq and _use_more_bits(i, cfg.n_layer):
            v = Q6_K
        if is70 and v == Q4_K:
            v = Q5_K
        d = Q6_K if (kq and _use_more_bits(i, cfg.n_layer)) else base
        out.update({p + "attn_q.weight": base, p + "attn_k.weight": base, p + "attn_v.weight": v,
                    p + "attn_output.weight": base, p + "ffn_gate.weight": base, p + "ffn_up.weight": base,
                    p + "ffn_down.weight": d})
    return out


def random_blocks(rng: np.random.Generator, t: int, n_rows: int, k: int, sigma: float = 0.02) -> np.ndarray:
    """n_rows x k weights as random quantized blocks of type t whose dequantized values have std ~ sigma.
    Every byte pattern of the packed fields is valid; only the fp16 super-block scales are chosen."""
    be, bb = BLOCK[t]
    nb = n_rows * (k // be)
    raw = np.frombuffer(rng.bytes(nb * bb), dtype=np.uint8).reshape(nb, bb).copy()

    def put_f16(col: int, vals: np.ndarray):
        raw[:, col:col + 2] = vals.astype(np.float16).view(np.uint8).reshape(nb, 2)

    jitter = rng.uniform(0.5, 1.5, size=nb)
    if t == Q4_K:      # {d, dmin, scales[12], qs[128]}: w = d*sc*q - dmin*m
        d = sigma / 257.0 * jitter
        put_f16(0, d); put_f16(2, 7.5 * d)
    elif t == Q5_K:    # q in 0..31
        d = sigma / 530.0 * jitter
        put_f16(0, d); put_f16(2, 15.5 * d)
    elif t == Q6_K:    # {ql[128], qh[64], scales i8[16], d}: w = d*sc*(q-32)
        put_f16(208, sigma / 1369.0 * jitter)
    elif t == Q8_0:    # {d, qs[32]}
        put_f16(0, sigma / 74.0 * jitter)
    else:
        raise ValueError(t)
    return raw.reshape(-1)


def llama_kv(cfg: LlamaConfig, ftype: str, extra_kv: Dict[str, tuple] | None = None) -> Dict[str, tuple]:
    kv = {
        "general.architecture": ("str", "llama"),
        "general.name": ("str", cfg.name),
        "general.alignment": ("u32", ALIGN),
        "general.file_type": ("u32", FTYPE.get(ftype, 0)),
        "llama.context_length": ("u32", cfg.n_ctx_train),
        "llama.embedding_length": ("u32", cfg.n_embd),
        "llama.block_count": ("u32", cfg.n_layer),
        "llama.feed_forward_length": ("u32", cfg.n_ff),
        "llama.attention.head_count": ("u32", cfg.n_head),
        "llama.attention.head_count_kv": ("u32", cfg.n_head_kv),
        "llama.attention.layer_norm_rms_epsilon": ("f32", cfg.rms_eps),
        "llama.rope.dimension_count": ("u32", cfg.head_dim),
        "llama.rope.freq_base": ("f32", cfg.rope_freq_base),
        "llama.vocab_size": ("u32", cfg.n_vocab),
        "tokenizer.ggml.model": ("str", "no_vocab"),   # cpp/src/llama.cpp:5267-5268
    }
    if extra_kv:
        kv.update(extra_kv)
    return kv


def synth_llama(path: str, cfg: LlamaConfig, ftype: str = "Q4_K_M", seed: int = 1234, source: str = "blocks",
                extra_kv: Dict[str, tuple] | None = None, rope_freqs: np.ndarray | None = None, share_period: int = 0) -> None:
    """Write a synthetic LLaMA-architecture GGUF. source="blocks": random quantized blocks in the reference's
    type mixture for `ftype`; source="f32": F32 weights ~ N(0, 0.02^2) (norms ~ 1 +- 0.1), ftype ignored.
    share_period > 0 (blocks only): a layer's matrix re-uses the FILE bytes of the same matrix of an earlier layer with the
    same block type (at most `share_period` distinct copies per (matrix, type)) — a 42 GB 70B-shaped file becomes ~3 GB on
    disk while every layer still has its own tiles in HBM, so timing is unaffected. Never used for parity fixtures."""
    rng = np.random.default_rng(seed)
    E, KV, FF, V = cfg.n_embd, cfg.n_head_kv * cfg.head_dim, cfg.n_ff, cfg.n_vocab
    shapes = [("token_embd.weight", (E, V)), ("output_norm.weight", (E,)), ("output.weight", (E, V))]
    for i in range(cfg.n_layer):
        p = f"blk.{i}."
        shapes += [(p + "attn_norm.weight", (E,)), (p + "attn_q.weight", (E, E)), (p + "attn_k.weight", (E, KV)),
                   (p + "attn_v.weight", (E, KV)), (p + "attn_output.weight", (E, E)), (p + "ffn_norm.weight", (E,)),
                   (p + "ffn_gate.weight", (E, FF)), (p + "ffn_up.weight", (E, FF)), (p + "ffn_down.weight", (FF, E))]
    if rope_freqs is not None:
        shapes.append(("rope_freqs.weight", (cfg.head_dim // 2,)))
    types = tensor_types(cfg, ftype) if source == "blocks" else {}
    tensors = []
    shared: Dict[tuple, List[str]] = {}
    for name, ne in shapes:
        if name == "rope_freqs.weight":
            tensors.append((name, ne, F32, np.asarray(rope_freqs, dtype=np.float32)))
        elif len(ne) == 1:
            tensors.append((name, ne, F32, (1.0 + 0.1 * rng.standard_normal(ne[0])).astype(np.float32)))
        elif source == "blocks":
            t = types[name]
            # lazy: generated when written, so only one tensor is in memory at a time
            sub = np.random.default_rng(rng.integers(0, 2**63))
            if share_period > 0 and name.startswith("blk."):
                key = (name.split(".", 2)[2], t)
                have = shared.setdefault(key, [])
                if len(have) >= share_period:
                    tensors.append((name, ne, t, ("alias", have[int(name.split(".")[1]) % share_period])))
                    continue
                have.append(name)
            tensors.append((name, ne, t, (lambda s=sub, t=t, ne=ne: random_blocks(s, t, ne[1], ne[0]))))
        else:
            sub = np.random.default_rng(rng.integers(0, 2**63))
            tensors.append((name, ne, F32, (lambda s=sub, ne=ne: (0.02 * s.standard_normal((ne[1], ne[0]), dtype=np.float32)))))
    write_gguf(path, llama_kv(cfg, ftype if source == "blocks" else "F32", extra_kv), tensors)


def weight_bytes_per_token(cfg: LlamaConfig, ftype: str) -> int:
    """Algorithmic weight bytes one decoded token reads: all per-layer matrices + output +
    norm vectors; token_embd excluded (one row)."""
    types = tensor_types(cfg, ftype)
    E, KV, FF, V = cfg.n_embd, cfg.n_head_kv * cfg.head_dim, cfg.n_ff, cfg.n_vocab
    tot = row_bytes(types["output.weight"], E) * V + 4 * E
    for i in range(cfg.n_layer):
        p = f"blk.{i}."
        tot += row_bytes(types[p + "attn_q.weight"], E) * E + 2 * 0
        tot += row_bytes(types[p + "attn_k.weight"], E) * KV + row_bytes(types[p + "attn_v.weight"], E) * KV
        tot += row_bytes(types[p + "attn_output.weight"], E) * E
        tot += row_bytes(types[p + "ffn_gate.weight"], E) * FF + row_bytes(types[p + "ffn_up.weight"], E) * FF
        tot += row_bytes(types[p + "ffn_down.weight"], FF) * E
        tot += 2 * 4 * E
    return tot


def kv_bytes_per_token(cfg: LlamaConfig, n_kv: int) -> int:
    """f16 K and V rows attended by one decoded token at kv length n_kv."""
    return 2 * cfg.n_layer * n_kv * cfg.n_head_kv * cfg.head_dim * 2
