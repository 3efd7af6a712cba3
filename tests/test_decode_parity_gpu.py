"""Model-level parity of the CUDA decode path (through the C-ABI): logits BIT-IDENTICAL to
  * the golden logits the unmodified reference produced for reference-quantized tiny models (tests/golden),
  * the port oracle (itself bit-identical to the reference) on seeded random-block twins with the real 8B / 70B
    per-layer shapes,
  * the reference itself (oracle/_ref, AVX-512 build) when it loads on this host,
plus size-independent properties at BASELINE.json's full sizes (determinism, graph == un-graphed, device-greedy
== host-greedy, stage split == single stage).

Bitwise equality is the only meaningful criterion here: activations are re-quantized to int8 before every
mat-mul, so any 1-ulp deviation eventually flips a quant and cascades to ~1e-2 (DESIGN.md "why bit-exact").
It also makes greedy decoding token-id-exact by construction."""
import os

import numpy as np
import pytest

from conftest import rel_err
from booster_b200 import engine, gguf_io as G
from oracle import port

pytestmark = pytest.mark.gpu


def _same(a, b, what=""):
    assert np.array_equal(a, b), f"{what}: max abs diff {np.abs(np.asarray(a) - np.asarray(b)).max():.3e} rel {rel_err(a, b):.3e}"


def _synth(model_dir, cfg, ftype, seed=7):
    p = os.path.join(model_dir, f"{cfg}_{ftype}_s{seed}.gguf")
    if not os.path.exists(p):
        G.synth_llama(p, G.CONFIGS[cfg], ftype, seed=seed, source="blocks")
    return p


# tiny-gqa4-yarn: YaRN rope scaling with ext_factor = 1, attn_factor 1.25, a 48-token prompt past the original context of 32
@pytest.mark.parametrize("model", ["tiny_Q4_K_M", "tiny_Q5_K_M", "tiny_Q8_0", "tiny-gqa4_Q4_K_M", "tiny-gqa4-yarn_Q4_K_M"])
def test_golden_models_bitwise(golden_dir, model):
    g = np.load(os.path.join(golden_dir, model + ".npz"))
    m = engine.Model(os.path.join(golden_dir, model + ".gguf"))
    c = engine.Context(m, 64)
    prompt = g["prompt"].tolist()
    lg = c.decode(prompt, 0)                                  # batch > 1 arithmetic
    _same(lg, g["logits"][0], "prefill")
    pos = len(prompt)
    for i, t in enumerate(g["ids"].tolist()):
        assert int(np.argmax(lg)) == t, f"greedy id differs from the reference at step {i}"
        lg = c.decode([t], pos)
        pos += 1
        _same(lg, g["logits"][i + 1], f"step {i}")
    c.kv_clear()
    for i, t in enumerate(prompt[:6]):                        # batch-1 arithmetic from position 0
        _same(c.decode([t], i), g["single"][i], f"single {i}")
    c.close(); m.close()


@pytest.mark.parametrize("model", ["tiny_Q4_K_M", "tiny-gqa4_Q4_K_M"])
def test_golden_layer_taps_bitwise(golden_dir, model):
    """layer-wise: Qcur / kqv_merged_cont / l_out of the prefill call vs the reference's cb_eval taps"""
    g = np.load(os.path.join(golden_dir, model + ".npz"))
    m = engine.Model(os.path.join(golden_dir, model + ".gguf"))
    c = engine.Context(m, 64)
    c.set_taps(True)
    prompt = g["prompt"].tolist()
    c.decode(prompt, 0)
    T = len(prompt)
    for il in range(m.n_layer):
        for ours, theirs in (("Qcur", f"Qcur_{il}"), ("kqv_merged_cont", f"kqv_merged_cont_{il}"), ("l_out", f"l_out_{il}")):
            a, b = c.get_tap(ours, il), g[theirs]
            b_last = b.reshape(T, -1)[-1] if b.size == a.size * T else b     # the last layer keeps only the last row
            _same(a, b_last, f"{ours}-{il}")
    c.close(); m.close()


@pytest.mark.parametrize("cfg,ftype,n_gen", [("llama3-8b-2l", "Q4_K_M", 12), ("llama3-8b-2l", "Q8_0", 3),
                                              ("llama3-8b-2l", "Q5_K_M", 4), ("llama3-70b-1l", "Q4_K_M", 3)])
def test_fullshape_twins_vs_port_bitwise(model_dir, cfg, ftype, n_gen):
    """the real per-layer shapes (n_embd 4096/8192, n_ff 14336/28672, GQA 4/8) with few layers: CUDA vs port"""
    path = _synth(model_dir, cfg, ftype)
    conf = G.CONFIGS[cfg]
    prompt = np.random.default_rng(42).integers(0, conf.n_vocab, size=20).tolist()
    p = port.PortModelRunner(path, n_ctx=128)
    ids_p, lg_p = p.greedy(prompt, n_gen)
    m = engine.Model(path)
    c = engine.Context(m, 128)
    ids_g, lg_g = c.greedy(prompt, n_gen)
    assert ids_g == ids_p
    for i, (a, b) in enumerate(zip(lg_g, lg_p)):
        _same(a, b, f"step {i}")
    c.close(); m.close()


def test_fullshape_twin_vs_reference_live_bitwise(model_dir, ref_or_none):
    ref = ref_or_none
    if ref is None or ref.variant() != "native":
        pytest.skip("the AVX-512 build of oracle/_ref does not load on this host")
    path = _synth(model_dir, "llama3-8b-2l", "Q4_K_M")
    prompt = np.random.default_rng(42).integers(0, 4096, size=40).tolist()
    r = ref.RefModel(path, n_ctx=128, n_threads=min(16, os.cpu_count() or 1))
    ids_r, lg_r = r.greedy(prompt, 16)
    m = engine.Model(path)
    c = engine.Context(m, 128)
    ids_g, lg_g = c.greedy(prompt, 16)
    assert ids_g == ids_r
    for i, (a, b) in enumerate(zip(lg_g, lg_r)):
        _same(a, b, f"step {i}")
    r.close(); c.close(); m.close()


def test_long_context_bitwise(model_dir):
    """attention across several 64-position tiles and a non-multiple-of-32 kv length"""
    path = _synth(model_dir, "tiny-gqa4", "Q4_K_M")
    prompt = np.random.default_rng(1).integers(0, 512, size=300).tolist()
    p = port.PortModelRunner(path, n_ctx=512)
    m = engine.Model(path); c = engine.Context(m, 512)
    _same(c.decode(prompt, 0), p.decode(prompt, 0), "prefill 300")
    _same(c.decode([17], 300), p.decode([17], 300), "decode @300")
    c.close(); m.close()


@pytest.mark.parametrize("cfg,ftype,n_ctx,n_kv0", [
    ("llama3-8b-2l", "Q4_K_M", 2048, 2040),      # BASELINE config 2: 8B shapes at the end of ctx 2048
    ("llama3-8b-2l", "Q5_K_M", 8192, 8186),      # config 4: Mistral-7B shapes (same as 8B, GQA 4), Q5_K_M, ctx 8192
    ("llama3-8b-2l", "Q8_0", 1536, 1530),        # config 3: Q8_0 after 512 prefill + 1K generated
    ("llama3-70b-1l", "Q4_K_M", 4096, 4090),     # config 5: 70B shapes (GQA 8), ctx 4096
])
def test_fullshape_decode_at_full_context_vs_port_bitwise(model_dir, cfg, ftype, n_ctx, n_kv0):
    """the real per-layer shapes decoding at the END of each BASELINE config's context. Both sides start from the same
    KV cache (N(0,1) f16 rows for positions < n_kv0, injected through b200_kv_write / the port's cache arrays — a
    prefill of thousands of tokens through the scalar port would take minutes) and then decode real tokens: the new
    K/V rows, the attention over all n_kv0+ positions and the logits must be bit-identical, and stay so token after token."""
    path = _synth(model_dir, cfg, ftype)
    conf = G.CONFIGS[cfg]
    kvd = conf.n_head_kv * conf.head_dim
    rng = np.random.default_rng(n_kv0)
    p = port.PortModelRunner(path, n_ctx=n_ctx)
    m = engine.Model(path)
    c = engine.Context(m, n_ctx)
    for il in range(conf.n_layer):
        k = rng.standard_normal((n_kv0, kvd)).astype(np.float16)
        v = rng.standard_normal((n_kv0, kvd)).astype(np.float16)
        p.kc[il, :n_kv0] = k.view(np.uint16); p.vc[il, :n_kv0] = v.view(np.uint16)
        c.kv_write(il, 0, k, v)
    tok = 17
    for i in range(min(4, n_ctx - n_kv0)):
        a, b = c.decode([tok], n_kv0 + i), p.decode([tok], n_kv0 + i)
        _same(a, b, f"decode at n_kv {n_kv0 + i + 1}")
        tok = int(np.argmax(b))
    for il in range(conf.n_layer):                 # the rows this run appended to the cache
        k, v = c.kv_read(il, n_kv0, 2)
        assert np.array_equal(k.view(np.uint16), p.kc[il, n_kv0:n_kv0 + 2]) and np.array_equal(v.view(np.uint16), p.vc[il, n_kv0:n_kv0 + 2])
    c.close(); m.close()


@pytest.mark.parametrize("model", ["tiny-gqa4_Q4_K_M", "tiny-gqa4-yarn_Q4_K_M"])
def test_context_shift_bitwise_golden(golden_dir, model):
    """llama_kv_cache_seq_rm + seq_add (cpp/bridge.cpp:487-507), the K-shift re-rotation of the cached K rows
    (build_k_shift, cpp/src/llama.cpp:8482-8510) and the reference's cell bookkeeping (find_slot re-uses the freed cells in the
    middle of the cache) — every logit of the scenario in tests/kvshift_script.py equals the reference's, bit for bit.
    The YaRN model scales cos/sin by the attention factor, so its K-shift touches every cell (as the reference's does)."""
    import kvshift_script
    g = np.load(os.path.join(golden_dir, f"kvshift_{model}.npz"))
    m = engine.Model(os.path.join(golden_dir, model + ".gguf"))
    c = engine.Context(m, 64)
    lg = kvshift_script.run(c, g["prompt"].tolist())
    assert lg.shape == g["logits"].shape
    for i in range(lg.shape[0]):
        _same(lg[i], g["logits"][i], f"step {i}")
    # the device-resident greedy loop addresses cells by position: refused after a shift instead of computing something else
    with pytest.raises(engine.B200Error):
        c.generate_greedy(1, 40, 2)
    c.kv_clear()                                   # llama_kv_cache_clear: back to an unshifted cache
    assert c.generate_greedy(1, 0, 2).shape == (2,)
    c.close(); m.close()


@pytest.mark.parametrize("model", ["tiny-gqa4_Q4_K_M", "tiny-gqa4-yarn_Q4_K_M"])
def test_self_extend_bitwise_golden(golden_dir, model):
    """SURVEY §8 f-3, Self-Extend (cpp/bridge.cpp:509-524, grp_attn_n = 2, grp_attn_w = 16): llama_kv_cache_seq_add / seq_div /
    seq_add before every decode — prompt chunks and generated tokens — compress whole windows of positions; cells keep their
    places, several cells share a position, every window is re-rotated by its own delta. Every logit of
    tests/kvshift_script.py::run_self_extend equals the reference's, bit for bit."""
    import kvshift_script
    g = np.load(os.path.join(golden_dir, f"selfextend_{model}.npz"))
    m = engine.Model(os.path.join(golden_dir, model + ".gguf"))
    c = engine.Context(m, 64)
    lg = kvshift_script.run_self_extend(c, g["prompt"].tolist())
    assert lg.shape == g["logits"].shape
    for i in range(lg.shape[0]):
        _same(lg[i], g["logits"][i], f"step {i}")
    with pytest.raises(engine.B200Error):
        c.kv_seq_div(0, 8, 0)                      # a zero divisor is an error, not a crash
    c.close(); m.close()


@pytest.mark.parametrize("cfg,ftype,n_prompt", [("llama3-8b-2l", "Q4_K_M", 200), ("llama3-8b-2l", "Q8_0", 130), ("llama3-8b-2l", "Q5_K_M", 70),
                                                ("llama3-70b-1l", "Q4_K_M", 65), ("tiny-gqa4", "Q4_K_M", 450)])
def test_prompt_batch_kernels_equal_token_by_token_bitwise(model_dir, cfg, ftype, n_prompt):
    """a prompt batch through the batched kernels (prefill.cuh: weight tiles fetched once per 64 tokens,
    attention with the token in blockIdx.z) gives the logits AND the KV cache of the token-by-token path, bit for bit —
    partial last chunks of 64, two passes of 512 for the long prompt, every block type, GQA 4 and 8, then decode continues
    on the cache the batch wrote."""
    path = _synth(model_dir, cfg, ftype)
    conf = G.CONFIGS[cfg]
    prompt = np.random.default_rng(n_prompt).integers(0, conf.n_vocab, size=n_prompt).tolist()
    m = engine.Model(path)
    out = {}
    for batch in (False, True):
        engine.set_prefill_batch(batch)
        try:
            c = engine.Context(m, 512)
            l0 = c.kernel_launches()
            lg = [c.decode(prompt, 0)]
            n_launch = c.kernel_launches() - l0
            lg.append(c.decode([int(np.argmax(lg[0]))], n_prompt))
            lg.append(c.decode([7, 8, 9, 10, 11, 12, 13, 14, 15], n_prompt + 1))     # a second, short batch on top
            kv = [c.kv_read(il, 0, n_prompt + 10) for il in range(conf.n_layer)]
            out[batch] = (lg, kv, n_launch)
            c.close()
        finally:
            engine.set_prefill_batch(True)
    assert out[True][2] < out[False][2] / 4, "the batched kernels were not used"
    for a, b in zip(out[True][0], out[False][0]):
        _same(a, b, "logits")
    for (ka, va), (kb, vb) in zip(out[True][1], out[False][1]):
        assert np.array_equal(ka.view(np.uint16), kb.view(np.uint16)) and np.array_equal(va.view(np.uint16), vb.view(np.uint16))
    m.close()


def test_device_greedy_equals_host_greedy_and_is_deterministic(model_dir):
    path = _synth(model_dir, "llama3-8b-2l", "Q4_K_M")
    m = engine.Model(path)
    c = engine.Context(m, 256)
    prompt = [11, 22, 33, 44, 55]
    ids_h, _ = c.greedy(prompt, 32)
    c.kv_clear()
    lg = c.decode(prompt, 0)
    first = int(np.argmax(lg))
    dev = c.generate_greedy(first, len(prompt), 32)        # CUDA-graph loop, arg-max on device
    assert [first] + dev.tolist()[:-1] == ids_h
    dev2 = c.generate_greedy(first, len(prompt), 32)
    assert np.array_equal(dev, dev2)
    c.close(); m.close()


def test_graph_equals_ungraphed(model_dir):
    path = _synth(model_dir, "llama3-8b-2l", "Q4_K_M")
    m = engine.Model(path)
    c = engine.Context(m, 64)
    a = c.decode([1, 2, 3, 4], 0).copy()
    c.kv_clear(); c.set_taps(True)
    b = c.decode([1, 2, 3, 4], 0)
    assert np.array_equal(a, b)
    assert c.get_tap("l_out", 1) is not None
    c.close(); m.close()


def test_position_bounds_and_bad_tokens(golden_dir):
    m = engine.Model(os.path.join(golden_dir, "tiny_Q4_K_M.gguf"))
    c = engine.Context(m, 32)
    with pytest.raises(engine.B200Error):
        c.decode([1] * 40, 0)                # exceeds n_ctx: the reference's find_slot failure (llama.cpp:14690)
    with pytest.raises(engine.B200Error):
        c.decode([m.n_vocab], 0)
    with pytest.raises(engine.B200Error):
        c.decode([1], 32)
    c.close(); m.close()


def test_in_process_stage_split_prompt_back_to_back(golden_dir):
    """a 12-token prompt enqueued back to back through two stages (no host sync between tokens, as the bridge's prompt
    loop does): the consumer's copy of token i's l_out must not race with the producer's token i+1 (back edge event) —
    logits and the greedy continuation equal the single-stage run. Uses two devices when the box has them."""
    import ctypes as C
    path = os.path.join(golden_dir, "tiny-gqa4_Q4_K_M.gguf")
    g = np.load(os.path.join(golden_dir, "tiny-gqa4_Q4_K_M.npz"))
    prompt = g["prompt"].tolist()
    d1 = 1 if engine.device_count() > 1 else 0
    m0 = engine.Model(path, 0, 0, 1); m1 = engine.Model(path, d1, 1, 3)
    c0 = engine.Context(m0, 64); c1 = engine.Context(m1, 64)
    L = m0.L
    out = np.empty(m0.n_vocab, dtype=np.float32)
    for rep in range(3):
        for i, t in enumerate(prompt):
            engine.check(L.b200_stage_forward(c0.h, t, i, 1, None), "stage 0")
            engine.check(L.b200_stage_forward(c1.h, t, i, 1, c0.h), "stage 1")
        engine.check(L.b200_stage_logits(c1.h, out.ctypes.data_as(C.POINTER(C.c_float))), "logits")
        _same(out, g["logits"][0], f"prefill through 2 stages (rep {rep})")
    pos = len(prompt)
    for i, t in enumerate(g["ids"].tolist()[:4]):
        assert int(np.argmax(out)) == t
        engine.check(L.b200_stage_forward(c0.h, t, pos, 0, None), "stage 0")
        engine.check(L.b200_stage_forward(c1.h, t, pos, 0, c0.h), "stage 1")
        engine.check(L.b200_stage_logits(c1.h, out.ctypes.data_as(C.POINTER(C.c_float))), "logits")
        _same(out, g["logits"][i + 1], f"step {i} through 2 stages")
        pos += 1
    for x in (c0, c1):
        x.close()
    for x in (m0, m1):
        x.close()


def test_in_process_stage_split_equals_single_stage(golden_dir):
    """layer split [0,1) + [1,2) on the same device through the stage API == one stage"""
    path = os.path.join(golden_dir, "tiny_Q4_K_M.gguf")
    m = engine.Model(path); c = engine.Context(m, 64)
    full = c.decode([5], 0)
    m0 = engine.Model(path, 0, 0, 1); m1 = engine.Model(path, 0, 1, 2)
    c0 = engine.Context(m0, 64); c1 = engine.Context(m1, 64)
    L = m.L
    import ctypes as C
    engine.check(L.b200_stage_forward(c0.h, 5, 0, 0, None), "stage 0")
    engine.check(L.b200_stage_forward(c1.h, 5, 0, 0, c0.h), "stage 1")
    out = np.empty(m.n_vocab, dtype=np.float32)
    engine.check(L.b200_stage_logits(c1.h, out.ctypes.data_as(C.POINTER(C.c_float))), "logits")
    assert np.array_equal(out, full)
    for x in (c, c0, c1):
        x.close()
    for x in (m, m0, m1):
        x.close()


def test_in_process_stage_split_prompt_batch(model_dir):
    """a prompt chunk through the batched prompt kernels of every stage of an in-process layer split (b200_stage_forward_batch:
    one peer copy of the residual streams [n][n_embd] per boundary) == the single-stage batch == token by token: logits,
    and the greedy continuation through the per-token stage path on the KV rows the batch wrote. Two devices when present."""
    import ctypes as C
    path = _synth(model_dir, "llama3-8b-2l", "Q4_K_M")
    conf = G.CONFIGS["llama3-8b-2l"]
    prompt = np.random.default_rng(77).integers(0, conf.n_vocab, size=150).tolist()
    m = engine.Model(path); c = engine.Context(m, 256)
    full = c.decode(prompt, 0)
    nxt = c.decode([int(np.argmax(full))], len(prompt))
    d1 = 1 if engine.device_count() > 1 else 0
    m0 = engine.Model(path, 0, 0, 1); m1 = engine.Model(path, d1, 1, 2)
    c0 = engine.Context(m0, 256); c1 = engine.Context(m1, 256)
    L = m.L
    toks = (C.c_int32 * len(prompt))(*prompt)
    out = np.empty(m.n_vocab, dtype=np.float32)
    for rep in range(2):
        assert L.b200_stage_batch_usable(c0.h, len(prompt)) == 1 and L.b200_stage_batch_usable(c1.h, len(prompt)) == 1
        engine.check(L.b200_stage_forward_batch(c0.h, toks, len(prompt), 0, None), "stage 0 batch")
        engine.check(L.b200_stage_forward_batch(c1.h, toks, len(prompt), 0, c0.h), "stage 1 batch")
        engine.check(L.b200_stage_logits(c1.h, out.ctypes.data_as(C.POINTER(C.c_float))), "logits")
        _same(out, full, f"prompt batch through 2 stages (rep {rep})")
    t = int(np.argmax(out))
    engine.check(L.b200_stage_forward(c0.h, t, len(prompt), 0, None), "stage 0")
    engine.check(L.b200_stage_forward(c1.h, t, len(prompt), 0, c0.h), "stage 1")
    engine.check(L.b200_stage_logits(c1.h, out.ctypes.data_as(C.POINTER(C.c_float))), "logits")
    _same(out, nxt, "decode after the staged prompt batch")
    for x in (c, c0, c1):
        x.close()
    for x in (m, m0, m1):
        x.close()
