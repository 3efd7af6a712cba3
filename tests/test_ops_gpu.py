"""Operator-level parity of the CUDA kernels (through the C-ABI, same kernels the engine launches) against
the golden vectors the reference produced and against the port oracle on seeded inputs.
EVERYTHING is bit-exact: the kernels reproduce the reference's fp32 operation order (see kernels.cuh)."""
import os

import numpy as np
import pytest

from conftest import q8k_equal, rel_err
from booster_b200 import engine, gguf_io as G
from oracle import port

pytestmark = pytest.mark.gpu
TYPES = {"Q8_0": 8, "Q4_K": 12, "Q5_K": 13, "Q6_K": 14}


@pytest.fixture(scope="module")
def ops(golden_dir):
    return np.load(os.path.join(golden_dir, "ops.npz"))


def test_quantize_q8_K_bit_exact_golden(ops):
    assert q8k_equal(engine.op_quantize_q8_K(ops["act_x"]), ops["act_q8_K"])


def test_quantize_q8_0_bit_exact_golden(ops):
    assert np.array_equal(engine.op_quantize_q8_0(ops["act_x"]), ops["act_q8_0"])


@pytest.mark.parametrize("k", [256, 4096, 14336, 28672])
def test_quantize_bit_exact_vs_port(k):
    rng = np.random.default_rng(k)
    x = (rng.standard_normal(k) * rng.choice([1e-3, 1.0, 50.0])).astype(np.float32)
    x[::97] = 0.0
    assert q8k_equal(engine.op_quantize_q8_K(x), port.quantize_row_q8_K(x))
    assert np.array_equal(engine.op_quantize_q8_0(x), port.quantize_row_q8_0(x))


@pytest.mark.parametrize("name", list(TYPES))
def test_dequantize_bit_exact_golden(ops, name):
    t, k = TYPES[name], 1536
    rb = G.row_bytes(t, k)
    for r in (0, 7, 15):
        assert np.array_equal(engine.op_dequantize_row(t, ops[f"w_{name}"][r * rb:(r + 1) * rb], k), ops[f"deq_{name}"][r])


@pytest.mark.parametrize("name", list(TYPES))
def test_mul_mat_vec_golden(ops, name):
    # the reference's own ggml_vec_dot_* results, bit for bit
    y = engine.op_mul_mat_vec(TYPES[name], ops[f"w_{name}"], 16, 1536, ops["act_x"])
    assert np.array_equal(y, ops[f"dot_{name}"]), f"max abs diff {np.abs(y - ops[f'dot_{name}']).max()}"


@pytest.mark.parametrize("name", list(TYPES))
@pytest.mark.parametrize("shape", [(64, 256), (64, 768), (1024, 4096), (6144, 4096), (256, 14336), (128, 8192), (64, 28672)])
def test_mul_mat_vec_vs_port(name, shape):
    n, k = shape
    t = TYPES[name]
    rng = np.random.default_rng(n * 31 + k)
    w = G.random_blocks(rng, t, n, k)
    x = rng.standard_normal(k).astype(np.float32)
    y, yr = engine.op_mul_mat_vec(t, w, n, k, x), port.mul_mat_vec(t, w, n, k, x)
    assert np.array_equal(y, yr), f"max abs diff {np.abs(y - yr).max()} rel {rel_err(y, yr)}"


def test_mul_mat_vec_edge_activations():
    # all-zero activation blocks, ties at +-max: the integer path must still be exact
    rng = np.random.default_rng(5)
    k, n = 1024, 32
    x = rng.standard_normal(k).astype(np.float32)
    x[:256] = 0.0
    x[300] = 9.0; x[400] = -9.0
    for t in TYPES.values():
        w = G.random_blocks(rng, t, n, k)
        assert np.array_equal(engine.op_mul_mat_vec(t, w, n, k, x), port.mul_mat_vec(t, w, n, k, x))


@pytest.mark.parametrize("k", [256, 4096, 8192])
def test_rms_norm(k):
    rng = np.random.default_rng(k)
    x = rng.standard_normal(k).astype(np.float32) * 3
    w = (1 + 0.1 * rng.standard_normal(k)).astype(np.float32)
    y, yr = engine.op_rms_norm(x, w, 1e-5), port.rms_norm(x, w, 1e-5)
    # double accumulation on both sides: identical unless the mean lands on a float rounding boundary
    assert np.array_equal(y, yr) or rel_err(y, yr) < 1e-7
    assert rel_err(engine.op_rms_norm(x, None, 1e-6), port.rms_norm(x, None, 1e-6)) < 1e-7


@pytest.mark.parametrize("pos", [0, 1, 17, 2047, 8191])
@pytest.mark.parametrize("base", [10000.0, 500000.0])
def test_rope(pos, base):
    rng = np.random.default_rng(pos)
    x = rng.standard_normal(8 * 128).astype(np.float32)
    y, yr = engine.op_rope(x, 8, 128, pos, base), port.rope(x, 8, 128, pos, base)
    # the table is built with the host libm exactly like ggml_rope_cache_init; rotation is un-fused mul/sub
    assert np.array_equal(y, yr)
    ff = (1 + rng.random(64)).astype(np.float32)
    assert np.array_equal(engine.op_rope(x, 8, 128, pos, base, 0.5, ff), port.rope(x, 8, 128, pos, base, 0.5, ff))


# (n_head, n_head_kv, n_kv): GQA 1/2/4/8; the contexts of the BASELINE configs (2048, 4096, 8192 and one position short of
# them); (64, 8, 8192) exceeds one CTA's shared memory for the GQA score rows and takes the long-context route
@pytest.mark.parametrize("cfg", [(4, 1, 1), (4, 1, 33), (32, 8, 257), (32, 8, 2048), (64, 8, 700), (8, 8, 64), (2, 1, 5), (32, 8, 3000),
                                 (16, 8, 1000), (32, 8, 4096), (32, 8, 8191), (32, 8, 8192), (64, 8, 4095), (64, 8, 4096), (64, 8, 8192),
                                 (8, 1, 12000), (8, 1, 70000)])
@pytest.mark.parametrize("round_q", [False, True])
def test_attention_bit_exact(cfg, round_q):
    _attention_case(cfg, round_q)


# the long-context three-kernel route (taken automatically only when a CTA's share of the context exceeds its shared
# memory, e.g. (8, 1, 70000) above) forced at ordinary sizes too
@pytest.mark.parametrize("cfg", [(4, 1, 33), (32, 8, 2048), (64, 8, 4095), (64, 8, 8192), (2, 1, 5), (16, 8, 1000)])
@pytest.mark.parametrize("round_q", [False, True])
def test_attention_long_context_route_bit_exact(cfg, round_q):
    engine.set_attention_route(1)
    try:
        _attention_case(cfg, round_q)
    finally:
        engine.set_attention_route(0)


def _attention_case(cfg, round_q):
    n_head, n_head_kv, n_kv = cfg
    hd = 128
    rng = np.random.default_rng(n_kv)
    q = (rng.standard_normal(n_head * hd) * 2).astype(np.float32)
    k = rng.standard_normal((n_kv, n_head_kv * hd)).astype(np.float16)
    v = rng.standard_normal((n_kv, n_head_kv * hd)).astype(np.float16)
    scale = 1.0 / np.sqrt(hd)
    y = engine.op_attention(q, k, v, n_kv, n_head, n_head_kv, hd, scale, round_q)
    yr = port.attention(q, k, v, n_kv, n_head, n_head_kv, hd, scale, round_q)
    assert np.array_equal(y, yr), f"max abs diff {np.abs(y - yr).max()} rel {rel_err(y, yr)}"


# ---- batch > 1 mat-mul (prompt batches): k_mma_batch on the tensor cores for K-quants, k_matmul_batch (dp4a) for Q8_0 ----
@pytest.mark.parametrize("name", list(TYPES))
@pytest.mark.parametrize("shape", [(64, 256, 8), (128, 768, 33), (1024, 4096, 70), (96, 14336, 31), (320, 4096, 512)])
def test_mul_mat_batch_vs_port(name, shape):
    """x[T][k] through the prompt-batch kernels equals the reference arithmetic token by token (ragged T, rows that are not
    a multiple of 64, one K step and 56 of them)"""
    n, k, T = shape
    t = TYPES[name]
    rng = np.random.default_rng(n * 7 + k + T)
    w = G.random_blocks(rng, t, n, k)
    x = (rng.standard_normal((T, k)) * rng.choice([0.01, 1.0, 30.0], size=(T, 1))).astype(np.float32)
    x[min(3, T - 1), :256] = 0.0
    y = engine.op_mul_mat(t, w, n, k, x)
    for i in sorted(set([0, 1, T // 2, T - 2, T - 1])):
        yr = port.mul_mat_vec(t, w, n, k, x[i])
        assert np.array_equal(y[i], yr), f"token {i}: max abs diff {np.abs(y[i] - yr).max()} rel {rel_err(y[i], yr)}"


def _extreme_blocks(name, n, k, rng):
    """rows of worst-case magnitudes for the exact-fp16 argument of prefill_mma.cuh: largest quants x largest sub-block
    scales (Q4_K 15*63, Q5_K 31*63, Q6_K -32*-128 and 31*127), mixed with random rows"""
    t = TYPES[name]
    w = G.random_blocks(rng, t, n, k).reshape(n, -1).copy()
    nb = k // 256
    if name == "Q4_K":
        blk = np.full(144, 0xFF, np.uint8); blk[0:4] = np.array([1.0, 0.5], np.float16).view(np.uint8)
    elif name == "Q5_K":
        blk = np.full(176, 0xFF, np.uint8); blk[0:4] = np.array([1.0, 0.5], np.float16).view(np.uint8)
    elif name == "Q6_K":
        blk = np.zeros(210, np.uint8); blk[192:208] = 0x80; blk[208:210] = np.array([1.0], np.float16).view(np.uint8)
    else:
        blk = np.full(34, 0x81, np.uint8); blk[0:2] = np.array([1.0], np.float16).view(np.uint8)
    per = blk.size
    reps = w.shape[1] // per
    w[0::4] = np.tile(blk, reps)
    if name == "Q6_K":   # q = 63 (w = +31, odd), scale +127
        b2 = np.full(210, 0xFF, np.uint8); b2[192:208] = 0x7F; b2[208:210] = np.array([1.0], np.float16).view(np.uint8)
        w[1::4] = np.tile(b2, nb)
    return w.reshape(-1)


@pytest.mark.parametrize("name", list(TYPES))
def test_mul_mat_batch_extremes(name):
    """every quant at +-127 against the largest weights: the integer sums reach their bounds (Q6_K: 32 * 4096 * 127 < 2^24)
    and must still come out of the fp16 tensor-core contraction exactly"""
    rng = np.random.default_rng(11)
    n, k, T = 64, 1024, 40
    w = _extreme_blocks(name, n, k, rng)
    x = rng.standard_normal((T, k)).astype(np.float32)
    x[0] = 3.0; x[1] = -3.0                          # every quant -127 / +127
    x[2] = np.where(np.arange(k) % 2 == 0, 5.0, -5.0)
    x[3, :] = 0.0; x[3, ::256] = 1.0                 # a single non-zero per block
    y = engine.op_mul_mat(TYPES[name], w, n, k, x)
    for i in range(T if name != "Q8_0" else 8):
        yr = port.mul_mat_vec(TYPES[name], w, n, k, x[i])
        assert np.array_equal(y[i], yr), f"token {i}: max abs diff {np.abs(y[i] - yr).max()}"


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("name", ["Q4_K", "Q5_K", "Q6_K", "Q8_0"])
def test_mul_mat_batch_kernel_variants(name, mode):
    """the batch kernels — mma.sync (1, default; Q8_0: the block-diagonal HMMA kernel), tcgen05 (2; K-quants), dp4a (0): A/B
    switch b200_set_prefill_mma — give the reference's bits, on random blocks and on the worst-case magnitudes"""
    rng = np.random.default_rng(3)
    n, k, T = 256, 2048, 100
    x = rng.standard_normal((T, k)).astype(np.float32)
    x[0] = 3.0; x[1] = -3.0
    w = _extreme_blocks(name, n, k, rng)
    engine.set_prefill_mma(mode)
    try:
        y = engine.op_mul_mat(TYPES[name], w, n, k, x)
    finally:
        engine.set_prefill_mma(1)
    for i in (0, 1, 2, 50, 99):
        yr = port.mul_mat_vec(TYPES[name], w, n, k, x[i])
        assert np.array_equal(y[i], yr), f"mode {mode} token {i}: max abs diff {np.abs(y[i] - yr).max()}"
