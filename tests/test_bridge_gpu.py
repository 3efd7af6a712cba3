"""The drop-in boundary end to end: the nine cgo entry points (include/bridge.h == cpp/bridge.h:132-165) driven
exactly as pkg/server/server.go drives them (init -> initContext -> doInference with a concurrent status poller
-> timing/promptEval/getPromptTokenCount), on a no_vocab model whose prompt is a list of token ids."""
import ctypes as C
import os
import threading
import time

import numpy as np
import pytest

from booster_b200 import _lib
from oracle import port

pytestmark = pytest.mark.gpu


def _init(path, n_ctx=64, predict=8, idx=0):
    L = _lib.lib()
    L.init(b"", b"")
    return L, L.initContext(idx, path.encode(), 1, 0, 100, 0, 0, 0, n_ctx, predict, 0, 0.0, 0.0, 0.0, 1, 1.0, 1.0,
                            1.0, 0, 1, 200, 1.0, 1.0, 1.0, 42, b"")


def test_do_inference_matches_oracle_greedy(golden_dir):
    path = os.path.join(golden_dir, "tiny_Q4_K_M.gguf")
    g = np.load(os.path.join(golden_dir, "tiny_Q4_K_M.npz"))
    L, ctx = _init(path, predict=8)
    assert ctx
    prompt = " ".join(str(t) for t in g["prompt"].tolist())
    seen = []
    stop = threading.Event()

    def poll():                       # status() is called concurrently from other goroutines (router.go:123)
        while not stop.is_set():
            seen.append(L.status(b"job-1"))
            time.sleep(0.0005)

    th = threading.Thread(target=poll); th.start()
    n = L.doInference(0, ctx, b"job-1", b"", prompt.encode())
    stop.set(); th.join()
    text = L.status(b"job-1").decode()
    ids = [int(x) for x in text.split()]
    assert ids[:12] == g["prompt"].tolist()                       # status = prompt pieces + generated pieces
    assert ids[12:] == g["ids"].tolist()                          # greedy ids == the reference's (bit-exact logits)
    assert n == 12 + 7                                            # n_p_eval + n_eval (the last sampled token is not decoded)
    assert L.getPromptTokenCount(b"job-1") == 12
    assert L.timing(b"job-1") >= 0 and L.promptEval(b"job-1") >= 0
    assert L.getSeed(b"job-1") == 42
    pu, gu = C.c_double(), C.c_double()
    assert L.b200_job_timing_us(b"job-1", C.byref(pu), C.byref(gu)) == 0 and gu.value > 0
    assert all(s is not None for s in seen)


def test_prompt_too_long_returns_zero(golden_dir):
    path = os.path.join(golden_dir, "tiny_Q4_K_M.gguf")
    L, ctx = _init(path, n_ctx=32, predict=4, idx=1)
    assert ctx
    prompt = " ".join(["7"] * 40)                                 # > n_ctx - 4 (cpp/bridge.cpp:382-386)
    assert L.doInference(1, ctx, b"job-long", b"", prompt.encode()) == 0
    assert L.getPromptTokenCount(b"job-long") == 40


def test_stop_inference(golden_dir):
    path = os.path.join(golden_dir, "tiny_Q8_0.gguf")
    L, ctx = _init(path, n_ctx=64, predict=-1, idx=2)
    assert ctx
    t = threading.Timer(0.02, lambda: L.stopInference(2))
    t.start()
    n = L.doInference(2, ctx, b"job-stop", b"", b"1 2 3")
    t.join()
    assert 0 < n <= 64


def test_bad_model_path_returns_null():
    L = _lib.lib()
    L.init(b"", b"")
    assert not L.initContext(3, b"/nonexistent.gguf", 1, 0, 100, 0, 0, 0, 64, 8, 0, 0.0, 0.0, 0.0, 1, 1.0, 1.0, 1.0,
                             0, 1, 200, 1.0, 1.0, 1.0, 42, b"")


@pytest.mark.parametrize("kind", ["spm", "bpe"])
def test_do_inference_with_text_prompt(kind, tmp_path):
    """text in, text out through the nine symbols with a real vocabulary: the prompt is tokenized as
    llama_tokenize(model, prompt, add_special=false, parse_special=true) (cpp/bridge.cpp:275-278), status() is the
    concatenation of llama_token_to_piece of prompt and generated tokens (cpp/bridge.cpp:628-632), generation stops at
    an end-of-generation token (cpp/bridge.cpp:640)."""
    import dataclasses

    import tokenizer_fixtures as F
    from booster_b200 import engine, gguf_io as G
    extra = F.vocab_kv(kind, pad_to=32)
    n_vocab = extra["llama.vocab_size"][1]
    cfg = dataclasses.replace(G.CONFIGS["tiny"], n_vocab=n_vocab)
    path = str(tmp_path / f"tiny_{kind}.gguf")
    G.synth_llama(path, cfg, "Q4_K_M", seed=5, extra_kv=extra)
    tok = engine.Tokenizer(path)
    prompt = "Hello world, it's 42 tokens<|user|>ok" if kind == "spm" else "Hello world, it's 42 tokens<|eot_id|>ok"
    ids = tok.tokenize(prompt.encode(), add_special=False, parse_special=True)
    assert 8 < len(ids) < 40
    # the same ids through the token-level API give the generated continuation
    m = engine.Model(path); c = engine.Context(m, 64)
    gen, _ = c.greedy(ids, 6)
    c.close(); m.close()
    idx = 2 if kind == "spm" else 3
    L, ctx = _init(path, n_ctx=64, predict=6, idx=idx)
    assert ctx
    job = f"job-text-{kind}".encode()
    n = L.doInference(idx, ctx, job, b"", prompt.encode())
    assert L.getPromptTokenCount(job) == len(ids)
    out = L.status(job)
    expect_ids = list(ids)
    for g in gen:
        expect_ids.append(g)
        if tok.is_eog(g):
            break
    assert out == b"".join(tok.piece(i, True) for i in expect_ids)
    assert n >= len(ids)
    tok.close()


def test_two_pods_concurrently_equal_sequential(golden_dir):
    """the Go server runs one doInference per pod, up to 8 pods at once, each on its own OS thread:
    two pods decoding at the same time on one GPU give exactly what each gives alone"""
    path = os.path.join(golden_dir, "tiny_Q4_K_M.gguf")
    g = np.load(os.path.join(golden_dir, "tiny_Q4_K_M.npz"))
    prompt = " ".join(str(t) for t in g["prompt"].tolist()).encode()
    L = _lib.lib()
    ctxs = []
    for idx in (4, 5):
        _, ctx = _init(path, n_ctx=64, predict=8, idx=idx)
        assert ctx
        ctxs.append((idx, ctx))
    results = {}

    def run(idx, ctx, rounds):
        for r in range(rounds):
            job = f"pod{idx}-r{r}".encode()
            L.doInference(idx, ctx, job, b"", prompt)
            results[job] = L.status(job).decode()

    ths = [threading.Thread(target=run, args=(idx, ctx, 6)) for idx, ctx in ctxs]
    for t in ths: t.start()
    for t in ths: t.join()
    expect = g["prompt"].tolist() + g["ids"].tolist()
    assert len(results) == 12
    for job, text in results.items():
        assert [int(x) for x in text.split()] == expect, job


@pytest.mark.parametrize("kind", ["spm", "bpe"])
def test_do_inference_janus_sampler_equals_reference_golden(kind, tmp_path, golden_dir):
    """SURVEY §8 f-2 through the nine symbols: with the seed fixed, doInference generates the token ids the reference's own
    bridge loop generates with its Janus sampler (tests/golden/janus.json, made by the unmodified cpp/janus.cpp) — logits
    bit-identical, sampler arithmetic identical, same std::mt19937 draws. Three settings: the reference's deterministic one,
    its defaults, and a wide short list (real random draws)."""
    import json

    import test_sampler
    from booster_b200 import engine
    cases = [c for c in json.load(open(os.path.join(golden_dir, "janus.json"))) if c["kind"] == kind]
    assert len(cases) >= 12
    path = test_sampler._model(tmp_path, kind)
    tok = engine.Tokenizer(path)
    L = _lib.lib()
    L.init(b"", b"")
    idx = 6
    for i, c in enumerate(cases):
        ctx = L.initContext(idx, path.encode(), 1, 0, 100, 0, 0, 0, 64, c["n_predict"], 0, 0.0, 0.0, 0.8, 40, 0.95, 1.0, 1.0, 64,
                            1, c["depth"], c["scale"], c["hi"], c["lo"], c["seed"], b"")
        assert ctx
        job = f"janus-{kind}-{i}".encode()
        L.doInference(idx, ctx, job, b"", c["text"].encode())
        assert L.getSeed(job) == c["seed"]
        want = b"".join(tok.piece(t, True) for t in c["prompt"] + c["ids"])
        assert L.status(job) == want, (c["depth"], c["scale"], c["hi"], c["lo"], c["seed"])
    tok.close()


# (the byte-level BPE fixture is covered at the sampler level on the CPU — tests/test_sampler.py; through status() its pieces can
#  contain a NUL byte (token 0 renders as b"\x00"), which a C string cannot carry: the text ends there for the Go server too)
@pytest.mark.parametrize("kind", ["spm"])
def test_do_inference_standard_chain_equals_reference_golden(kind, tmp_path, golden_dir):
    """janus = 0 through the nine symbols: doInference samples with the standard chain (repetition penalty, top-k, typical, top-p,
    min-p, temperature, mirostat 1 / 2) and publishes the token ids the reference's own llama_sampling_sample generates in the
    bridge's loop for the same seed (tests/golden/standard_chain.json, made by oracle/_ref: common/sampling.cpp +
    src/llama-sampling.cpp unmodified) — logits bit-identical, candidate order, float arithmetic and mt19937 draws identical."""
    import json

    import test_sampler
    from booster_b200 import engine
    cases = [c for c in json.load(open(os.path.join(golden_dir, "standard_chain.json"))) if c["kind"] == kind]
    assert len(cases) >= 24
    path = test_sampler._model(tmp_path, kind, big=False)
    tok = engine.Tokenizer(path)
    L = _lib.lib()
    L.init(b"", b"")
    idx = 5
    for i, c in enumerate(cases):
        assert tok.tokenize(c["text"].encode(), False, True) == c["prompt"]
        ctx = L.initContext(idx, path.encode(), 1, 0, 100, 0, 0, 0, 64, c["n_predict"], c["mirostat"], c["mirostat_tau"], c["mirostat_eta"],
                            c["temperature"], c["top_k"], c["top_p"], c["typical_p"], c["repetition_penalty"], c["penalty_last_n"],
                            0, 200, 1.0, 1.0, 1.0, c["seed"], b"")
        assert ctx
        job = f"standard-{kind}-{i}".encode()
        L.doInference(idx, ctx, job, b"", c["text"].encode())
        want = b"".join(tok.piece(t, True) for t in c["prompt"] + c["ids"]).split(b"\x00", 1)[0]   # status() is a C string
        assert L.status(job) == want, {k: v for k, v in c.items() if k not in ("prompt", "ids", "text")}
    tok.close()


def test_pod_split_over_two_gpus_equals_one_gpu(model_dir):
    """the reference's gpu1 / gpu2 proportions (pkg/server/server.go:514-530 -> tensor_split): a pod whose layers sit on two
    devices — prompt chunk through b200_stage_forward_batch (peer copy of the residual streams), generation through the
    per-token stage chain — publishes the same ids as the pod on one device. Needs two visible GPUs."""
    from booster_b200 import engine, gguf_io as G
    if engine.device_count() < 2:
        pytest.skip("one visible GPU")
    path = os.path.join(model_dir, "llama3-8b-2l_Q4_K_M_s7.gguf")
    if not os.path.exists(path):
        G.synth_llama(path, G.CONFIGS["llama3-8b-2l"], "Q4_K_M", seed=7, source="blocks")
    prompt = " ".join(str(int(t)) for t in np.random.default_rng(5).integers(0, 4096, size=120)).encode()
    L = _lib.lib()
    L.init(b"", b"")
    texts = []
    # (one GPU, 512-token chunks) / (two GPUs, one chunk) / (two GPUs, chunks of 32 kept in flight across the stages)
    for idx, (g1, g2, nb) in ((4, (100, 0, 0)), (5, (50, 50, 0)), (6, (50, 50, 32))):
        ctx = L.initContext(idx, path.encode(), 1, nb, g1, g2, 0, 0, 256, 12, 0, 0.0, 0.0, 0.0, 1, 1.0, 1.0, 1.0, 0, 1, 200,
                            1.0, 1.0, 1.0, 42, b"")
        assert ctx
        job = f"split-{idx}".encode()
        assert L.doInference(idx, ctx, job, b"", prompt) == 120 + 11
        texts.append(L.status(job).decode())
    assert texts[0] == texts[1] == texts[2]
    assert len(texts[0].split()) == 132
