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


# The port restates EVERY floating-point operation of the reference's AVX-512 ("native") build in its exact
# order — AVX2-lane fma chains of the K-quant dots, tinyBLAS 16-lane chains + _mm512_reduce_add_ps for K.q and
# V.p, ggml_vec_dot_f16 for batch > 1, the ggml_v_expf polynomial in silu/softmax, double row sums — so whole-model
# logits are BIT-IDENTICAL to the reference's, step after step. Bitwise equality is the criterion: with quantized
# activations anything weaker is meaningless, because a 1-ulp upstream difference eventually flips an int8 of a
# Q8_K block and the deviation cascades to ~1e-2 and persists through the KV cache (DESIGN.md "why bit-exact").
@pytest.mark.parametrize("model", ["tiny_Q4_K_M", "tiny_Q5_K_M", "tiny_Q8_0", "tiny-gqa4_Q4_K_M", "tiny-gqa4-yarn_Q4_K_M"])
def test_model_logits_vs_golden_bitwise(golden_dir, model):
    g = np.load(os.path.join(golden_dir, model + ".npz"))
    m = port.PortModelRunner(os.path.join(golden_dir, model + ".gguf"), n_ctx=64)
    prompt = g["prompt"].tolist()
    lg = m.decode(prompt, 0)                                   # batch (n > 1) arithmetic
    assert np.array_equal(lg, g["logits"][0])
    n_layer = m.M.n_layer
    assert np.array_equal(m.tap_l_out[n_layer - 1], g[f"l_out_{n_layer - 1}"])
    pos = len(prompt)
    for i, t in enumerate(g["ids"].tolist()):
        assert int(np.argmax(lg)) == t
        lg = m.decode([t], pos)                                # batch-1 arithmetic
        pos += 1
        assert np.array_equal(lg, g["logits"][i + 1]), f"step {i}"
    m.kv_clear()
    for i, t in enumerate(prompt[:6]):
        assert np.array_equal(m.decode([t], i), g["single"][i])


@pytest.mark.parametrize("model", ["tiny-gqa4_Q4_K_M", "tiny-gqa4-yarn_Q4_K_M"])
@pytest.mark.parametrize("scenario", ["kvshift", "selfextend"])
def test_port_kv_cells_vs_golden_bitwise(golden_dir, model, scenario):
    """SURVEY §8 f-3 restated on the CPU: struct llama_kv_cache's cell bookkeeping (find_slot, seq_rm, seq_add, seq_div), the
    K-shift of the cached f16 K rows by each cell's accumulated delta (build_k_shift) and attention masked by the cells'
    positions — the port reproduces every logit of the reference's context-shift scenario (cpp/bridge.cpp:487-507) and of its
    Self-Extend scenario (:509-524), bit for bit, incl. the YaRN model whose K-shift rescales every cell."""
    import kvshift_script
    g = np.load(os.path.join(golden_dir, f"{scenario}_{model}.npz"))
    m = port.PortModelRunner(os.path.join(golden_dir, model + ".gguf"), n_ctx=64)
    run = kvshift_script.run if scenario == "kvshift" else kvshift_script.run_self_extend
    lg = run(m, g["prompt"].tolist())
    assert lg.shape == g["logits"].shape
    for i in range(lg.shape[0]):
        assert np.array_equal(lg[i], g["logits"][i]), f"step {i}"
    assert m.managed and not m.has_shift


def _random_kv_ops(r, p, rng, n_ctx):
    """one random walk of decodes (single tokens and batches) and KV operations (context shift, a removed middle range closed
    again, a window of positions divided Self-Extend style, a tail moved up) applied to the reference and to the port alike;
    returns False at the first logit difference. A cache that runs full must run full on both sides."""
    n_past, cells = 0, 0
    lg = None

    def both(fn, *a):
        getattr(r, fn)(*a)
        getattr(p, fn)(*a)

    def dec(toks):
        nonlocal n_past, lg
        ea = eb = None
        try:
            a = r.decode(toks, n_past)
        except RuntimeError as e:
            ea = e
        try:
            b = p.decode(toks, n_past)
        except RuntimeError as e:
            eb = e
        if ea or eb:
            assert bool(ea) == bool(eb), (ea, eb)
            return None                                     # full on both sides: the walk ends
        n_past += len(toks)
        lg = a
        return np.array_equal(a, b)

    ok = dec([int(t) for t in rng.integers(0, 512, size=int(rng.integers(1, 20)))])
    cells = n_past
    for _ in range(60):
        if ok is not True:
            return ok is None
        op = int(rng.integers(0, 10))
        if op < 5 or n_past < 8:
            n = int(rng.integers(1, 6)) if op < 4 else 1
            if cells + n > n_ctx - 1:                       # cpp/bridge.cpp:495-503
                n_keep = int(rng.integers(0, 4)); n_disc = (n_past - n_keep) // 2
                both("kv_seq_rm", n_keep, n_keep + n_disc); both("kv_seq_add", n_keep + n_disc, n_past, -n_disc)
                n_past -= n_disc; cells -= n_disc
                continue
            ok = dec([int(np.argmax(lg))] if n == 1 else [int(t) for t in rng.integers(0, 512, size=n)])
            cells += n
        elif op < 7:
            a = int(rng.integers(0, n_past - 2)); b = int(rng.integers(a + 1, min(n_past, a + 8)))
            both("kv_seq_rm", a, b); both("kv_seq_add", b, n_past, -(b - a))
            cells -= b - a; n_past -= b - a
        elif op < 9:
            w = int(rng.choice([4, 8]))
            if n_past >= w + 2:                             # cpp/bridge.cpp:512-523
                a = int(rng.integers(0, n_past - w)); end = (a + w - 1) // 2 + 1
                both("kv_seq_div", a, a + w, 2); both("kv_seq_add", a + w, n_past, end - (a + w))
                n_past += end - (a + w)
        else:
            both("kv_seq_add", int(rng.integers(0, n_past)), -1, int(rng.integers(1, 4)))
            n_past += 3
    return ok is not False


def test_port_kv_cells_random_ops_vs_reference_live(ref_or_none, golden_dir):
    """the port's KV-cell bookkeeping, K-shift and position mask against the LIVE reference on random walks of decodes and
    llama_kv_cache_seq_rm / seq_add / seq_div calls (freed cells re-used out of order, positions that coincide, deltas that
    accumulate over several operations before the next decode applies them): every logit bit-identical, and a full cache is
    full on both sides"""
    ref = ref_or_none
    if ref is None or not hasattr(ref.lib(), "refshim_kv_seq_div"):
        pytest.skip("oracle/_ref with the KV shim is not available")
    if ref.variant() != "native":
        pytest.skip("bitwise parity is defined against the reference's -march=native build")
    path = os.path.join(golden_dir, "tiny-gqa4_Q4_K_M.gguf")
    r = ref.RefModel(path, n_ctx=64, n_threads=2)
    for seed in range(16):
        p = port.PortModelRunner(path, n_ctx=64)
        r.kv_clear()
        assert _random_kv_ops(r, p, np.random.default_rng(seed), 64), f"seed {seed}"
    r.close()


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
    if ref.variant() != "native":
        pytest.skip("bit-exactness is defined against the AVX-512 (native) build of the reference")
    lp, lr = p.decode(toks, 0), r.decode(toks, 0)
    assert np.array_equal(lp, lr)
    pos = len(toks)
    for _ in range(6):
        t = int(np.argmax(lr))
        lp, lr = p.decode([t], pos), r.decode([t], pos)
        pos += 1
        assert np.array_equal(lp, lr)
    r.close()
