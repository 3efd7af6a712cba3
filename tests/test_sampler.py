"""the samplers behind doInference, pinned on the CPU against the reference's own sampler.

Janus (booster_b200/csrc/janus.cpp) restates initJanus + sample_janus_token (cpp/janus.cpp:191-331, 405-700). The reference's
janus.cpp is compiled UNMODIFIED into oracle/_ref; oracle/ref_shim.cpp runs the bridge's generation loop with it
(refshim_janus_generate). Here the reference generates with a fixed seed, then the same logits (the reference's, step by
step) go through b200_sampler_*: the ids must be equal, token for token — including the draws from multi-candidate short
lists (same std::mt19937 / std::discrete_distribution arithmetic). No GPU involved: the sampler is host code."""
import dataclasses
import os

import numpy as np
import pytest

import tokenizer_fixtures as F
from booster_b200 import engine, gguf_io as G


def _model(tmp_path, kind, big=True):
    # The reference's initJanus writes FIXED token ids up to 29936 into its scales table whenever the vocabulary has at most
    # 128000 entries (the "LLaMA-2" branch, cpp/janus.cpp:626-693) — a heap overflow for the small synthetic vocabularies
    # (it crashes the reference). The fixtures are therefore padded: SPM to 30016 tokens (LLaMA-2 branch, all ids valid),
    # BPE to 128256 (LLaMA-3 branch: by piece text and id range).
    extra = F.vocab_kv(kind, pad_to=(30016 if kind == "spm" else 128256) if big else 32)
    cfg = dataclasses.replace(G.CONFIGS["tiny"], n_vocab=extra["llama.vocab_size"][1])
    path = str(tmp_path / f"tiny_{kind}.gguf")
    G.synth_llama(path, cfg, "Q4_K_M", seed=5, extra_kv=extra)
    return path


# (depth, scale, hi, lo): the reference's deterministic setting, its defaults, and two wide short lists (real draws)
SETTINGS = [(200, 1.0, 1.0, 1.0), (200, 0.96, 0.99, 0.96), (8, 0.9, 0.7, 0.5), (200, 0.97, 0.3, 0.2)]


@pytest.mark.parametrize("kind", ["spm", "bpe"])
def test_janus_equals_reference_token_for_token(kind, tmp_path, ref_or_none):
    ref = ref_or_none
    if ref is None or not hasattr(ref.lib(), "refshim_janus_generate"):
        pytest.skip("oracle/_ref with the sampler shim is not available")
    path = _model(tmp_path, kind)
    tok = engine.Tokenizer(path)
    prompts = [tok.tokenize(s.encode(), False, True) for s in ("Hello world, it's 42 tokens", "русский язык и ещё", "{ \"a\": [1, 2, 3] }")]
    tok.close()
    r = ref.RefModel(path, n_ctx=64, n_threads=2)
    n_multi = 0
    for depth, scale, hi, lo in SETTINGS:
        s = engine.Sampler(path, 64, janus=1, depth=depth, scale=scale, hi=hi, lo=lo)
        for pi, prompt in enumerate(prompts):
            for seed in (1, 12345):
                ids_ref = r.janus_generate(prompt, 24, depth, scale, hi, lo, seed, n_predict=24)
                # replay: the reference's logits step by step through our sampler
                r.kv_clear()
                lg = r.decode(prompt, 0)
                s.reset(prompt, seed)
                pos = len(prompt)
                ours = []
                for want in ids_ref:
                    got = s.sample(lg, pos, 24)
                    ours.append(got)
                    if got != want:
                        break
                    lg = r.decode([got], pos)
                    pos += 1
                assert ours == ids_ref, (kind, (depth, scale, hi, lo), pi, seed)
                n_multi += len(set(ids_ref)) > 1
        s.close()
    assert n_multi > 0
    r.close()


# the standard chain (janus = 0): settings of llama_sampling_params that initContext can reach, plus tfs_z / min_p
STANDARD = [
    dict(),                                                                  # the defaults: k f y p m t (temp 0.8, top-k 40, top-p 0.95, min-p 0.05)
    dict(temp=1.3, top_k=0, top_p=1.0, min_p=0.0),                           # pure temperature over the whole vocabulary
    dict(temp=1.0, top_k=200, top_p=0.9, min_p=0.02),                        # top-k > 128: the bucket pre-sort
    dict(temp=0.9, penalty_repeat=1.3, penalty_last_n=16),                   # penalties over a short window (+ newline restore)
    dict(temp=0.9, penalty_repeat=1.15, penalty_last_n=-1, top_k=60),        # the whole window
    dict(temp=1.0, typical_p=0.6, top_k=0, min_p=0.0),                       # locally typical (leaves the candidates unsorted)
    dict(temp=1.0, tfs_z=0.9, top_k=100, min_p=0.0),                         # tail-free
    dict(temp=0.0, penalty_repeat=1.5),                                      # greedy over the penalised logits
    dict(temp=-1.0, penalty_repeat=1.2),                                     # "greedy with probabilities"
    dict(temp=1.0, mirostat=1, mirostat_tau=3.0, mirostat_eta=0.2),          # mirostat (mu starts at 0: llama_sampling_init never sets it)
    dict(temp=0.8, mirostat=2, mirostat_tau=4.0, mirostat_eta=0.3, penalty_repeat=1.1),
]


@pytest.mark.parametrize("kind", ["spm", "bpe"])
def test_standard_chain_equals_reference_token_for_token(kind, tmp_path, ref_or_none):
    """janus = 0: llama_sampling_sample as the reference links it (common/sampling.cpp over src/llama-sampling.cpp) generates with
    a fixed seed; the reference's logits, step by step, through b200_sampler_* must give the same ids — the same candidate order
    (std::sort / partial_sort / the top-k bucket sort), float arithmetic and mt19937 draws"""
    ref = ref_or_none
    if ref is None or not hasattr(ref.lib(), "refshim_standard_generate"):
        pytest.skip("oracle/_ref with the standard-chain shim is not available")
    if ref.variant() != "native":
        pytest.skip("the restated chain follows the fused multiply-adds of the reference's -march=native build")
    path = _model(tmp_path, kind, big=False)
    tok = engine.Tokenizer(path)
    prompts = [tok.tokenize(s.encode(), False, True) for s in ("Hello world, it's 42 tokens\nand a second line", "русский язык и ещё")]
    n_vocab = tok.n_vocab
    tok.close()
    assert n_vocab > 200                                                     # so that top_k = 200 takes the bucket path
    r = ref.RefModel(path, n_ctx=96, n_threads=2)
    n_varied = 0
    for si, kw in enumerate(STANDARD):
        ours_kw = dict(kw)
        s = engine.Sampler(path, 96, janus=0, temperature=ours_kw.pop("temp", 0.8), top_k=ours_kw.pop("top_k", 40), top_p=ours_kw.pop("top_p", 0.95),
                           repetition_penalty=ours_kw.pop("penalty_repeat", 1.0), penalty_last_n=ours_kw.pop("penalty_last_n", 64), **ours_kw)
        for pi, prompt in enumerate(prompts):
            for seed in (3, 777):
                ids_ref = r.standard_generate(prompt, 40, seed, **kw)
                r.kv_clear()
                lg = r.decode(prompt, 0)
                s.reset(prompt, seed)
                pos = len(prompt)
                ours = []
                for want in ids_ref:
                    got = s.sample(lg, pos)
                    ours.append(got)
                    if got != want:
                        break
                    lg = r.decode([got], pos)
                    pos += 1
                assert ours == ids_ref, (kind, si, kw, pi, seed)
                n_varied += len(set(ids_ref)) > 3
        s.close()
    assert n_varied >= len(STANDARD)                                         # real draws, not one token repeated
    r.close()
    # the full-size vocabularies (30 016 / 128 256 tokens): the top-k <= 128 path selects with the heap operations of
    # std::partial_sort over the raw logits instead of materialising every candidate — same ids, same order, same draws
    path = _model(tmp_path, kind)
    r = ref.RefModel(path, n_ctx=96, n_threads=2)
    for kw in (dict(), dict(temp=1.0, top_k=128, typical_p=0.7, penalty_repeat=1.2, penalty_last_n=16)):
        ours_kw = dict(kw)
        s = engine.Sampler(path, 96, janus=0, temperature=ours_kw.pop("temp", 0.8), top_k=ours_kw.pop("top_k", 40), top_p=ours_kw.pop("top_p", 0.95),
                           repetition_penalty=ours_kw.pop("penalty_repeat", 1.0), penalty_last_n=ours_kw.pop("penalty_last_n", 64), **ours_kw)
        ids_ref = r.standard_generate(prompts[0], 24, 11, **kw)
        r.kv_clear()
        lg = r.decode(prompts[0], 0)
        s.reset(prompts[0], 11)
        pos = len(prompts[0])
        ours = []
        for want in ids_ref:
            got = s.sample(lg, pos)
            ours.append(got)
            if got != want:
                break
            lg = r.decode([got], pos)
            pos += 1
        assert ours == ids_ref and len(set(ids_ref)) > 3, (kind, kw)
        s.close()
    r.close()


@pytest.mark.parametrize("kind", ["spm", "bpe"])
def test_janus_golden_ids_on_port_logits(kind, tmp_path, golden_dir):
    """the committed ids of tests/golden/janus.json (the reference's bridge loop with its own Janus sampler) from the PORT's
    logits — needs no reference library; the GPU test drives the same cases through doInference"""
    import json

    from oracle import port
    cases = [c for c in json.load(open(os.path.join(golden_dir, "janus.json"))) if c["kind"] == kind]
    assert len(cases) >= 12
    path = _model(tmp_path, kind)
    m = port.PortModelRunner(path, n_ctx=64)
    for c in cases:
        s = engine.Sampler(path, 64, janus=1, depth=c["depth"], scale=c["scale"], hi=c["hi"], lo=c["lo"])
        m.kv_clear()
        lg = m.decode(c["prompt"], 0)
        s.reset(c["prompt"], c["seed"])
        pos = len(c["prompt"])
        ours = []
        for want in c["ids"]:
            got = s.sample(lg, pos, c["n_predict"])
            ours.append(got)
            if got != want:
                break
            lg = m.decode([got], pos)
            pos += 1
        assert ours == c["ids"], {k: v for k, v in c.items() if k not in ("prompt", "ids", "text")}
        s.close()


@pytest.mark.parametrize("kind", ["spm", "bpe"])
def test_standard_chain_golden_ids_on_port_logits(kind, tmp_path, golden_dir):
    """the committed ids of tests/golden/standard_chain.json (generated by the reference's own chain inside the bridge's loop)
    from logits the PORT computes — no reference library needed: oracle_port.c is bit-identical to the reference's logits, the
    sampler restates its chain, so the ids must come out the same. The GPU test drives the same cases through doInference."""
    import json

    from oracle import port
    cases = [c for c in json.load(open(os.path.join(golden_dir, "standard_chain.json"))) if c["kind"] == kind]
    assert len(cases) >= 24
    path = _model(tmp_path, kind, big=False)
    m = port.PortModelRunner(path, n_ctx=64)
    for c in cases:
        s = engine.Sampler(path, 64, janus=0, temperature=c["temperature"], top_k=c["top_k"], top_p=c["top_p"],
                           repetition_penalty=c["repetition_penalty"], penalty_last_n=c["penalty_last_n"], mirostat=c["mirostat"],
                           mirostat_tau=c["mirostat_tau"], mirostat_eta=c["mirostat_eta"], typical_p=c["typical_p"])
        m.kv_clear()
        lg = m.decode(c["prompt"], 0)
        s.reset(c["prompt"], c["seed"])
        pos = len(c["prompt"])
        ours = []
        for want in c["ids"]:
            got = s.sample(lg, pos)
            ours.append(got)
            if got != want:
                break
            lg = m.decode([got], pos)
            pos += 1
        assert ours == c["ids"], {k: v for k, v in c.items() if k not in ("prompt", "ids", "text")}
        s.close()


def test_standard_chain_properties(tmp_path):
    """janus = 0 (a setting the reference ignores): temperature <= 0 and top_k = 1 are arg-max; a fixed seed is deterministic;
    the repetition penalty moves a repeated arg-max; top-k bounds the support"""
    path = _model(tmp_path, "bpe", big=False)
    rng = np.random.default_rng(0)
    n_vocab = engine.Tokenizer(path).n_vocab
    lg = rng.standard_normal(n_vocab).astype(np.float32) * 3
    top = int(np.argmax(lg))
    for kw in (dict(temperature=0.0), dict(temperature=0.8, top_k=1)):
        s = engine.Sampler(path, 64, janus=0, **kw)
        s.reset([1, 2, 3], 7)
        assert [s.sample(lg, 3 + i) for i in range(4)] == [top] * 4
        s.close()
    a = engine.Sampler(path, 64, janus=0, temperature=1.0, top_k=5, top_p=1.0)
    a.reset([1, 2, 3], 99); draws = [a.sample(lg, 3 + i) for i in range(200)]
    a.reset([1, 2, 3], 99); again = [a.sample(lg, 3 + i) for i in range(200)]
    assert draws == again and len(set(draws)) > 1
    assert set(draws) <= set(np.argsort(-lg)[:5].tolist())
    a.close()
    pen = engine.Sampler(path, 64, janus=0, temperature=0.0, repetition_penalty=100.0, penalty_last_n=64)
    pen.reset([top], 1)                                    # the arg-max token is in the penalty window: it loses
    assert pen.sample(lg, 1) != top
    pen.close()
