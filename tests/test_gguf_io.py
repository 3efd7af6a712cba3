"""Host logic: GGUF container writer/reader, the reference's tensor-type mixture, algorithmic byte counts."""
import os

import numpy as np

from booster_b200 import gguf_io as G


def test_roundtrip(tmp_path):
    p = str(tmp_path / "t.gguf")
    G.synth_llama(p, G.CONFIGS["tiny"], "Q4_K_M", seed=5, source="blocks")
    f = G.read_gguf(p)
    assert f.kv["general.architecture"] == "llama"
    assert f.kv["llama.block_count"] == 2 and f.kv["llama.attention.head_count_kv"] == 1
    t = f.tensors["blk.0.attn_q.weight"]
    assert t.ne == (256, 256) and t.type == G.Q4_K and t.data.size == 256 * 144
    assert f.tensors["output.weight"].type == G.Q6_K
    assert f.tensors["blk.1.ffn_down.weight"].type == G.Q6_K      # use_more_bits(1, 2)
    assert f.tensors["blk.0.ffn_down.weight"].type == G.Q4_K
    assert f.tensors["output_norm.weight"].type == G.F32


def test_mixture_matches_reference_rules():
    # cpp/src/llama.cpp:15442-15444, 15547-15555, 15603-15610: Q6_K layers for n_layer = 32
    cfg = G.CONFIGS["llama3-8b"]
    types = G.tensor_types(cfg, "Q4_K_M")
    more = [i for i in range(32) if types[f"blk.{i}.ffn_down.weight"] == G.Q6_K]
    assert more == [0, 1, 2, 3, 6, 9, 12, 15, 18, 21, 24, 27, 28, 29, 30, 31]
    assert all(types[f"blk.{i}.attn_v.weight"] == types[f"blk.{i}.ffn_down.weight"] for i in range(32))
    t70 = G.tensor_types(G.CONFIGS["llama3-70b"], "Q4_K_M")
    assert set(t70[f"blk.{i}.attn_v.weight"] for i in range(80)) == {G.Q5_K, G.Q6_K}   # 70B: Q4_K attn_v -> Q5_K


def test_algorithmic_bytes_match_survey_table():
    # SURVEY.md §8d / BASELINE.md §3
    assert round(G.weight_bytes_per_token(G.CONFIGS["llama3-8b"], "Q4_K_M") / 1e6, 1) == 4617.4
    assert round(G.weight_bytes_per_token(G.CONFIGS["llama3-8b"], "Q8_0") / 1e6, 1) == 7974.8
    assert round(G.weight_bytes_per_token(G.CONFIGS["mistral-7b"], "Q5_K_M") / 1e6, 1) == 5040.6
    assert round(G.weight_bytes_per_token(G.CONFIGS["llama3-70b"], "Q4_K_M") / 1e6, 1) == 41921.5
    assert round(G.kv_bytes_per_token(G.CONFIGS["llama3-8b"], 2048) / 1e6, 1) == 268.4


def test_random_blocks_are_sane(tmp_path):
    from oracle import port
    rng = np.random.default_rng(0)
    for t in (G.Q4_K, G.Q5_K, G.Q6_K, G.Q8_0):
        raw = G.random_blocks(rng, t, 4, 1024)
        w = np.stack([port.dequantize_row(t, raw[r * G.row_bytes(t, 1024):(r + 1) * G.row_bytes(t, 1024)], 1024) for r in range(4)])
        assert np.isfinite(w).all()
        assert 0.005 < w.std() < 0.08, (t, w.std())
