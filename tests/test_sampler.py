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
