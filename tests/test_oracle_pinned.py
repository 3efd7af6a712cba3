"""Pin the port oracle (oracle/oracle_port.c) against (1) golden vectors produced by the unmodified reference
(tests/golden/make_golden.py) and (2) the reference itself (oracle/_ref) when it is loadable on this host.
Integer/byte results must be bit-exact; the quantized mat-vec follows the reference's AVX2 lane order and is
bit-exact too; full-model logits agree to fp32 round-off (attention/softmax order differs)."""
import os

import numpy as np
import pytest

from conftest import q8k_equal, rel_err
from booster_b200 import gguf_io as G
from oracle import port

TYPES = {"Q8_0": 8, "Q4_K": 12, "Q5_K": 13, "Q6_K": 14}


@pytest.fixture(scope="module")
def ops(golden_dir):
    return np.load(os.path.join(golden_dir, "ops.npz"))


def test_quantize_q8_K_bit_exact(ops):
    # quantize_row_q8_K: ties at +-max (first wins), all-zero block, denormal-ish values, exact .5 products
    assert q8k_equal(port.quantize_row_q8_K(ops["act_x"]), ops["act_q8_K"])


def test_quantize_q8_0_bit_exact(ops):
    assert np.array_equal(port.quantize_row_q8_0(ops["act_x"]), ops["act_q8_0"])


@pytest.mark.parametrize("name", list(TYPES))
def test_dequantize_bit_exact(ops, name):
    t, k = TYPES[name], 1536
    w = ops[f"w_{name}"]
    rb = G.row_bytes(t, k)
    for r in range(16):
        assert np.array_equal(port.dequantize_row(t, w[r * rb:(r + 1) * rb], k), ops[f"deq_{name}"][r])


@pytest.mark.parametrize("name", list(TYPES))
def test_vec_dot_bit_exact(ops, name):
    t = TYPES[name]
    y = port.mul_mat_vec(t, ops[f"w_{name}"], 16, 1536, ops["act_x"])
    assert np.array_equal(y, ops[f"dot_{name}"]), f"max abs diff {np.abs(y - ops[f'dot_{name}']).max()}"


# A 1-ulp difference upstream (the port uses libm expf, the reference a SIMD polynomial) can flip ONE int8 of a
# Q8_K/Q8_0 activation; in these tiny models (K = 256..768) a single flip moves a layer output by ~1e-3 and the
# logits by up to ~1e-2 (measured: tiny-gqa4 at position 2). So model-level criteria are flip-robust:
# most steps agree to fp32 round-off, every step stays within the flip bound, greedy ids are exact.
ROUNDOFF, FLIP_BOUND = 5e-6, 5e-2


@pytest.mark.parametrize("model", ["tiny_Q4_K_M", "tiny_Q5_K_M", "tiny_Q8_0", "tiny-gqa4_Q4_K_M"])
def test_model_logits_vs_golden(golden_dir, model):
    g = np.load(os.path.join(golden_dir, model + ".npz"))
    m = port.PortModelRunner(os.path.join(golden_dir, model + ".gguf"), n_ctx=64)
    prompt = g["prompt"].tolist()
    errs = []
    # batch (n>1) arithmetic, then batch-1 steps, feeding the REFERENCE's token ids
    lg = m.decode(prompt, 0)
    errs.append(rel_err(lg, g["logits"][0]))
    pos = len(prompt)
    for i, t in enumerate(g["ids"].tolist()):
        assert int(np.argmax(lg)) == t, f"greedy id differs at step {i}"
        lg = m.decode([t], pos)
        pos += 1
        errs.append(rel_err(lg, g["logits"][i + 1]))
    m.kv_clear()
    for i, t in enumerate(prompt[:6]):
        errs.append(rel_err(m.decode([t], i), g["single"][i]))
    errs = np.array(errs)
    assert errs.max() < FLIP_BOUND, errs
    assert np.median(errs) < ROUNDOFF, errs
    assert (errs < ROUNDOFF).mean() >= 0.6, errs


def test_port_vs_reference_live(ref_or_none, model_dir):
    """when oracle/_ref loads on this host: random-block twin with the 8B per-layer shapes' arithmetic paths
    (mixed Q4_K/Q6_K, GQA 4) — port and reference must agree."""
    ref = ref_or_none
    if ref is None:
        pytest.skip("oracle/_ref not available on this host")
    rng = np.random.default_rng(3)
    for name, t in TYPES.items():
        w = ref.quantize_weights(0.02 * rng.standard_normal((8, 512)).astype(np.float32), t)
        x = rng.standard_normal(512).astype(np.float32)
        assert np.array_equal(port.mul_mat_vec(t, w, 8, 512, x), ref.mul_mat_vec(t, w, 8, 512, x))
    path = os.path.join(model_dir, "tiny-gqa4_Q5_K_M_blocks.gguf")
    G.synth_llama(path, G.CONFIGS["tiny-gqa4"], "Q5_K_M", seed=11, source="blocks")
    r = ref.RefModel(path, n_ctx=64, n_threads=4)
    p = port.PortModelRunner(path, n_ctx=64)
    toks = [3, 1, 4, 1, 5, 9, 2, 6]
    assert rel_err(p.decode(toks, 0), r.decode(toks, 0)) < FLIP_BOUND
    assert rel_err(p.decode([7], len(toks)), r.decode([7], len(toks))) < FLIP_BOUND
    r.close()
