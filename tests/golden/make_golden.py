"""Generate the golden fixtures under tests/golden/ from the UNMODIFIED reference (oracle/_ref, built from
/root/reference by oracle/Makefile). Run in the build container (needs oracle/_ref); the outputs are committed
so that the GPU box — where /root/reference does not exist — can check both the port oracle and the CUDA path
against vectors the reference itself produced.

    python tests/golden/make_golden.py

Outputs:
  tiny_<ftype>.gguf     tiny LLaMA-shaped models, F32 ~ N(0, 0.02^2) weights quantized by the reference's own
                        llama_model_quantize (cpp/src/llama.cpp:15435-) to Q4_K_M / Q5_K_M / Q8_0
  tiny_<ftype>.npz      reference logits: prefill of a 12-token prompt in one llama_decode (batch arithmetic),
                        8 greedy steps (batch-1 arithmetic), plus l_out taps of the prefill call
  ops.npz               operator vectors: quantize_row_q8_K / q8_0 (incl. ties at +-max and all-zero blocks),
                        reference-quantized weight rows of each block type with their dequantization and
                        ggml_vec_dot_* results
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from booster_b200 import gguf_io as G  # noqa: E402
from oracle import ref  # noqa: E402

PROMPT = [5, 9, 200, 17, 3, 99, 42, 7, 11, 300, 1, 2]
N_GEN = 8
MODELS = [("tiny", "Q4_K_M", 15), ("tiny", "Q5_K_M", 17), ("tiny", "Q8_0", 7), ("tiny-gqa4", "Q4_K_M", 15)]
# YaRN rope scaling with ext_factor = 1 (llama.rope.scaling.type = "yarn": cpp/src/llama.cpp:16686-16690), a prompt that
# runs past the original context so that the interpolation / extrapolation ramp and the attention factor all matter
YARN_KV = {"llama.rope.scaling.type": ("str", "yarn"), "llama.rope.scaling.factor": ("f32", 4.0),
           "llama.rope.scaling.original_context_length": ("u32", 32), "llama.rope.scaling.attn_factor": ("f32", 1.25)}
YARN_PROMPT = [int(t) for t in np.random.default_rng(7).integers(0, 512, size=48)]


def make_models(only=None):
    for cfg_name, ftype, code in MODELS + [("tiny-gqa4+yarn", "Q4_K_M", 15)]:
        yarn = cfg_name.endswith("+yarn")
        PROMPT = YARN_PROMPT if yarn else globals()["PROMPT"]
        cfg = G.CONFIGS[cfg_name.split("+")[0]]
        cfg_name = cfg_name.replace("+", "-")
        if only and cfg_name not in only:
            continue
        out = os.path.join(HERE, f"{cfg_name}_{ftype}.gguf")
        with tempfile.TemporaryDirectory() as td:
            f32 = os.path.join(td, "f32.gguf")
            G.synth_llama(f32, cfg, seed=1234, source="f32", extra_kv=YARN_KV if yarn else None)
            ref.quantize_model(f32, out, code, nthread=4)
        r = ref.RefModel(out, n_ctx=64, n_threads=4)
        names = [f"l_out-{i}" for i in range(cfg.n_layer)] + [f"Qcur-{i}" for i in range(cfg.n_layer)] + \
                [f"kqv_merged_cont-{i}" for i in range(cfg.n_layer)]
        r.set_taps(names)
        r.kv_clear()
        logits = [r.decode(PROMPT, 0)]
        taps = {n.replace("-", "_"): r.get_tap(n) for n in names}
        r.set_taps([])
        ids = []
        pos = len(PROMPT)
        for _ in range(N_GEN):
            t = int(np.argmax(logits[-1]))
            ids.append(t)
            logits.append(r.decode([t], pos))
            pos += 1
        # token-by-token (batch-1 arithmetic from position 0)
        r.kv_clear()
        single = [r.decode([t], i) for i, t in enumerate(PROMPT[:6])]
        np.savez_compressed(os.path.join(HERE, f"{cfg_name}_{ftype}.npz"), prompt=np.array(PROMPT, dtype=np.int32),
                            ids=np.array(ids, dtype=np.int32), logits=np.stack(logits), single=np.stack(single), **taps)
        r.close()
        print(f"{cfg_name} {ftype}: {os.path.getsize(out) / 1e6:.2f} MB, greedy ids {ids}")


def make_ops():
    rng = np.random.default_rng(99)
    out = {}
    # activation vectors: gaussian, ties at +max/-max (first occurrence must win), an all-zero block, tiny values
    x = rng.standard_normal(256 * 6).astype(np.float32)
    x[256:512] = 0.0
    x[512 + 7] = 3.5; x[512 + 100] = -3.5; x[512:768] = np.clip(x[512:768], -3.5, 3.5)
    x[768 + 200] = -4.25; x[768 + 13] = 4.25; x[768:1024] = np.clip(x[768:1024], -4.25, 4.25)
    x[1024:1280] *= 1e-6
    x[1280:1536] = np.round(x[1280:1536] * 4) / 4   # many exact .5 products
    out["act_x"] = x
    out["act_q8_K"] = ref.quantize_row_q8_K(x)
    out["act_q8_0"] = ref.quantize_row_q8_0(x)
    for name, t in ref.GGML_TYPE.items():
        n, k = 16, 1536
        w = ref.quantize_weights(0.02 * rng.standard_normal((n, k)).astype(np.float32), t)
        out[f"w_{name}"] = w
        out[f"deq_{name}"] = np.stack([ref.dequantize_row(t, w[r * G.row_bytes(t, k):(r + 1) * G.row_bytes(t, k)], k) for r in range(n)])
        out[f"dot_{name}"] = ref.mul_mat_vec(t, w, n, k, x)
    np.savez_compressed(os.path.join(HERE, "ops.npz"), **out)
    print("ops.npz written")


def make_kvshift():
    """tests/golden/kvshift_<model>.npz: logits of every step of tests/kvshift_script.py run by the reference (context shift:
    llama_kv_cache_seq_rm / seq_add, K-shift inside the next llama_decode, freed cells re-used)"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import kvshift_script
    for model in ("tiny-gqa4_Q4_K_M", "tiny-gqa4-yarn_Q4_K_M"):
        r = ref.RefModel(os.path.join(HERE, model + ".gguf"), n_ctx=64, n_threads=4)
        prompt = [int(t) for t in np.random.default_rng(11).integers(0, 512, size=30)]
        lg = kvshift_script.run(r, prompt)
        np.savez_compressed(os.path.join(HERE, f"kvshift_{model}.npz"), prompt=np.array(prompt, dtype=np.int32), logits=lg)
        print(f"kvshift_{model}.npz: {lg.shape[0]} steps")
        # Self-Extend (cpp/bridge.cpp:509-524): 30 prompt tokens in chunks of 8 + 28 generated, windows of 16 compressed by 2
        se = kvshift_script.run_self_extend(r, prompt)
        np.savez_compressed(os.path.join(HERE, f"selfextend_{model}.npz"), prompt=np.array(prompt, dtype=np.int32), logits=se)
        print(f"selfextend_{model}.npz: {se.shape[0]} steps")
        r.close()


def make_janus():
    """tests/golden/janus.json: token ids the reference's bridge loop generates with its own Janus sampler (unmodified
    cpp/janus.cpp through refshim_janus_generate) on tiny models with the padded synthetic vocabularies, for fixed seeds.
    The models are rebuilt by the tests with the same deterministic generator (tests/test_sampler.py::_model)."""
    import dataclasses
    import json
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import tokenizer_fixtures as F
    from booster_b200 import engine
    out = []
    for kind in ("spm", "bpe"):
        extra = F.vocab_kv(kind, pad_to=30016 if kind == "spm" else 128256)
        cfg = dataclasses.replace(G.CONFIGS["tiny"], n_vocab=extra["llama.vocab_size"][1])
        with tempfile.TemporaryDirectory() as td:
            path = os.path.join(td, f"tiny_{kind}.gguf")
            G.synth_llama(path, cfg, "Q4_K_M", seed=5, extra_kv=extra)
            tok = engine.Tokenizer(path)
            r = ref.RefModel(path, n_ctx=64, n_threads=2)
            for text in ("Hello world, it's 42 tokens", "русский язык и ещё"):
                prompt = tok.tokenize(text.encode(), False, True)
                for depth, scale, hi, lo in ((200, 1.0, 1.0, 1.0), (200, 0.96, 0.99, 0.96), (8, 0.9, 0.7, 0.5)):
                    for seed in (7, 4242):
                        ids = r.janus_generate(prompt, 20, depth, scale, hi, lo, seed, n_predict=20)
                        out.append({"kind": kind, "text": text, "prompt": prompt, "depth": depth, "scale": scale, "hi": hi, "lo": lo,
                                    "seed": seed, "n_predict": 20, "ids": ids})
            r.close(); tok.close()
    json.dump(out, open(os.path.join(HERE, "janus.json"), "w"), indent=0)
    print(f"janus.json: {len(out)} cases")


# initContext-reachable settings of the standard chain (janus = 0): (mirostat, tau, eta, temperature, top_k, top_p, typical_p,
# repetition_penalty, penalty_last_n)
STANDARD_CASES = [
    (0, 0.0, 0.0, 0.8, 40, 0.95, 1.0, 1.0, 64),
    (0, 0.0, 0.0, 1.1, 200, 0.9, 1.0, 1.2, 32),
    (0, 0.0, 0.0, 1.0, 0, 1.0, 0.6, 1.0, 0),
    (0, 0.0, 0.0, 0.0, 40, 0.95, 1.0, 1.4, 64),
    (1, 3.0, 0.2, 1.0, 40, 0.95, 1.0, 1.0, 64),
    (2, 4.0, 0.3, 0.8, 40, 0.95, 1.0, 1.1, 64),
]


def make_standard():
    """tests/golden/standard_chain.json: token ids the reference's bridge loop WOULD generate with the standard chain it keeps
    commented out (cpp/bridge.cpp:598; refshim_standard_generate runs the reference's own llama_sampling_init /
    llama_sampling_sample / llama_sampling_accept) on tiny models with small synthetic vocabularies, for fixed seeds."""
    import dataclasses
    import json
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import tokenizer_fixtures as F
    from booster_b200 import engine
    out = []
    for kind in ("spm", "bpe"):
        extra = F.vocab_kv(kind, pad_to=32)
        cfg = dataclasses.replace(G.CONFIGS["tiny"], n_vocab=extra["llama.vocab_size"][1])
        with tempfile.TemporaryDirectory() as td:
            path = os.path.join(td, f"tiny_{kind}.gguf")
            G.synth_llama(path, cfg, "Q4_K_M", seed=5, extra_kv=extra)
            tok = engine.Tokenizer(path)
            r = ref.RefModel(path, n_ctx=64, n_threads=2)
            for text in ("Hello world, it's 42 tokens\nand more", "русский язык и ещё"):
                prompt = tok.tokenize(text.encode(), False, True)
                for (mi, tau, eta, temp, top_k, top_p, typ, rep, last_n) in STANDARD_CASES:
                    for seed in (7, 4242):
                        ids = r.standard_generate(prompt, 20, seed, mirostat=mi, mirostat_tau=tau, mirostat_eta=eta, temp=temp, top_k=top_k,
                                                  top_p=top_p, typical_p=typ, penalty_repeat=rep, penalty_last_n=last_n)
                        out.append({"kind": kind, "text": text, "prompt": prompt, "mirostat": mi, "mirostat_tau": tau, "mirostat_eta": eta,
                                    "temperature": temp, "top_k": top_k, "top_p": top_p, "typical_p": typ, "repetition_penalty": rep,
                                    "penalty_last_n": last_n, "seed": seed, "n_predict": 20, "ids": ids})
            r.close(); tok.close()
    json.dump(out, open(os.path.join(HERE, "standard_chain.json"), "w"), indent=0)
    print(f"standard_chain.json: {len(out)} cases")


if __name__ == "__main__":
    if not ref.available():
        sys.exit("oracle/_ref is not built: run `make -C oracle ref` in the build container")
    print("reference variant:", ref.variant())
    only = sys.argv[1:]            # e.g. `make_golden.py tiny-gqa4-yarn`: (re)generate only the named models
    if only == ["janus"]:
        make_janus()
        sys.exit(0)
    if only == ["standard"]:
        make_standard()
        sys.exit(0)
    if only == ["kvshift"]:
        make_kvshift()
        sys.exit(0)
    if not only:
        make_ops()
        make_janus()
        make_standard()
        make_kvshift()
    make_models(only)
