"""The persistent per-token kernel (booster_b200/csrc/token_kernel.cuh: one launch per token, phases joined by grid
barriers) computes with the same device code as the per-operator kernels: logits BIT-IDENTICAL to the reference's golden
vectors, to the port oracle on the full 8B / 70B per-layer shapes at the end of the BASELINE contexts, and to the default
path. It is selectable (b200_set_token_kernel / BOOSTER_B200_TOKEN_KERNEL=1), not the default — measured slower."""
import os

import numpy as np
import pytest

from booster_b200 import engine, gguf_io as G
from oracle import port

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def token_kernel_on():
    engine.set_token_kernel(True)
    yield
    engine.set_token_kernel(False)


@pytest.mark.parametrize("model", ["tiny_Q4_K_M", "tiny_Q5_K_M", "tiny_Q8_0", "tiny-gqa4_Q4_K_M", "tiny-gqa4-yarn_Q4_K_M"])
def test_golden_models_bitwise(golden_dir, model):
    g = np.load(os.path.join(golden_dir, model + ".npz"))
    m = engine.Model(os.path.join(golden_dir, model + ".gguf"))
    c = engine.Context(m, 64)
    assert c.trace_phases(1, 0, reps=1) is not None, "the persistent kernel is not in use"
    prompt = g["prompt"].tolist()
    lg = c.decode(prompt, 0)
    assert np.array_equal(lg, g["logits"][0])
    pos = len(prompt)
    for i, t in enumerate(g["ids"].tolist()):
        assert int(np.argmax(lg)) == t
        lg = c.decode([t], pos)
        pos += 1
        assert np.array_equal(lg, g["logits"][i + 1]), f"step {i}"
    first = int(np.argmax(c.decode(prompt, 0)))
    dev = c.generate_greedy(first, len(prompt), 8)          # CUDA-graph loop of the persistent kernel + device arg-max
    assert [first] + dev.tolist()[:-1] == g["ids"].tolist()
    c.close(); m.close()


@pytest.mark.parametrize("cfg,ftype,n_ctx,n_kv0", [("llama3-8b-2l", "Q4_K_M", 2048, 2040), ("llama3-8b-2l", "Q5_K_M", 8192, 8186),
                                                   ("llama3-8b-2l", "Q8_0", 1536, 1530), ("llama3-70b-1l", "Q4_K_M", 4096, 4090)])
def test_fullshape_decode_at_full_context_vs_port_bitwise(model_dir, cfg, ftype, n_ctx, n_kv0):
    path = os.path.join(model_dir, f"{cfg}_{ftype}_s7.gguf")
    if not os.path.exists(path):
        G.synth_llama(path, G.CONFIGS[cfg], ftype, seed=7, source="blocks")
    conf = G.CONFIGS[cfg]
    kvd = conf.n_head_kv * conf.head_dim
    rng = np.random.default_rng(n_kv0)
    p = port.PortModelRunner(path, n_ctx=n_ctx)
    m = engine.Model(path)
    c = engine.Context(m, n_ctx)
    assert c.trace_phases(1, 0, reps=1) is not None
    for il in range(conf.n_layer):
        k = rng.standard_normal((n_kv0, kvd)).astype(np.float16)
        v = rng.standard_normal((n_kv0, kvd)).astype(np.float16)
        p.kc[il, :n_kv0] = k.view(np.uint16); p.vc[il, :n_kv0] = v.view(np.uint16)
        c.kv_write(il, 0, k, v)
    tok = 17
    for i in range(3):
        a, b = c.decode([tok], n_kv0 + i), p.decode([tok], n_kv0 + i)
        assert np.array_equal(a, b), f"decode at n_kv {n_kv0 + i + 1}: max abs diff {np.abs(a - b).max()}"
        tok = int(np.argmax(b))
    c.close(); m.close()
