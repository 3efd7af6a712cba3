"""First-contact GPU script (development aid, not a test): parity of the CUDA path against the real reference
CPU library on small synthetic models, operator checks, then a rough full-size timing. Run under gpurun."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from booster_b200 import engine, gguf_io as G  # noqa: E402
from oracle import ref  # noqa: E402

TMP = os.environ.get("B200_TMP", "/tmp/b200_models")
os.makedirs(TMP, exist_ok=True)


def rel(a, b):
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


def ops():
    rng = np.random.default_rng(0)
    for k in (256, 4096, 14336):
        x = rng.standard_normal(k).astype(np.float32)
        a, b = engine.op_quantize_q8_K(x), ref.quantize_row_q8_K(x)
        print(f"q8_K k={k}: identical={np.array_equal(a, b)}")
        a, b = engine.op_quantize_q8_0(x), ref.quantize_row_q8_0(x)
        print(f"q8_0 k={k}: identical={np.array_equal(a, b)} ndiff={int((a != b).sum())}")
    for name, t in (("Q4_K", 12), ("Q5_K", 13), ("Q6_K", 14), ("Q8_0", 8)):
        for (n, k) in ((64, 256), (128, 4096), (32, 14336)):
            w = ref.quantize_weights(0.02 * rng.standard_normal((n, k)).astype(np.float32), t)
            x = rng.standard_normal(k).astype(np.float32)
            y, yr = engine.op_mul_mat_vec(t, w, n, k, x), ref.mul_mat_vec(t, w, n, k, x)
            d = engine.op_dequantize_row(t, w[: G.row_bytes(t, k)], k)
            dr = ref.dequantize_row(t, w[: G.row_bytes(t, k)], k)
            print(f"matvec {name} {n}x{k}: rel={rel(y, yr):.2e} maxabs={np.abs(y - yr).max():.2e}  dequant identical={np.array_equal(d, dr)}")


def model_parity(cfg_name, ftype, n_prompt=12, n_gen=12, n_ctx=256):
    cfg = G.CONFIGS[cfg_name]
    path = os.path.join(TMP, f"{cfg_name}_{ftype}.gguf")
    if not os.path.exists(path):
        G.synth_llama(path, cfg, ftype, seed=7, source="blocks")
    prompt = np.random.default_rng(42).integers(0, cfg.n_vocab, size=n_prompt).tolist()
    r = ref.RefModel(path, n_ctx=n_ctx)
    t0 = time.time()
    ids_r, lg_r = r.greedy(prompt, n_gen)
    t_ref = time.time() - t0
    m = engine.Model(path)
    c = engine.Context(m, n_ctx)
    ids_g, lg_g = c.greedy(prompt, n_gen)
    errs = [rel(a, b) for a, b in zip(lg_g, lg_r)]
    print(f"[{cfg_name} {ftype}] ids equal={ids_g == ids_r} max rel logits err={max(errs):.2e} (first {errs[0]:.2e}) ref {t_ref:.1f}s")
    # single-token path from scratch: decode token by token (batch-1 arithmetic on both sides)
    r.kv_clear(); c.kv_clear()
    worst = 0.0
    for i, t in enumerate(prompt[:6]):
        worst = max(worst, rel(c.decode([t], i), r.decode([t], i)))
    print(f"    token-by-token max rel err={worst:.2e}")
    c.close(); m.close(); r.close()


def timing_8b(n_ctx=2048, steps=64):
    cfg = G.CONFIGS["llama3-8b"]
    path = os.path.join(TMP, "llama3-8b_Q4_K_M.gguf")
    t0 = time.time()
    if not os.path.exists(path):
        G.synth_llama(path, cfg, "Q4_K_M", seed=1234, source="blocks")
    print(f"synth 8B: {time.time() - t0:.1f}s, {os.path.getsize(path) / 1e9:.2f} GB")
    t0 = time.time()
    m = engine.Model(path)
    print(f"load: {time.time() - t0:.1f}s weight_bytes={m.weight_bytes / 1e6:.1f} MB")
    c = engine.Context(m, n_ctx)
    out = c.generate_greedy(1, 0, 16)  # warm
    for pos0 in (0, n_ctx - steps - 1):
        t0 = time.time()
        out = c.generate_greedy(1, pos0, steps)
        dt = time.time() - t0
        nkv = pos0 + steps / 2
        by = m.weight_bytes + G.kv_bytes_per_token(cfg, int(nkv))
        print(f"8B greedy pos0={pos0}: {steps / dt:.1f} tok/s  {dt / steps * 1e3:.3f} ms/tok  {by * steps / dt / 1e9:.0f} GB/s  ids[:6]={out[:6].tolist()}")
    c.close(); m.close()


if __name__ == "__main__":
    print("devices:", engine.device_count(), "ref variant:", ref.variant())
    what = sys.argv[1:] or ["ops", "parity", "timing"]
    if "ops" in what:
        ops()
    if "parity" in what:
        for cfg, ft in (("tiny", "Q4_K_M"), ("tiny-gqa4", "Q5_K_M"), ("tiny-gqa4", "Q8_0"), ("llama3-8b-2l", "Q4_K_M")):
            model_parity(cfg, ft)
    if "timing" in what:
        timing_8b()


def taps(cfg_name, ftype, tokens, n_ctx=64):
    """layer-wise localisation: reference cb_eval taps vs engine taps for one llama_decode call."""
    cfg = G.CONFIGS[cfg_name]
    path = os.path.join(TMP, f"{cfg_name}_{ftype}.gguf")
    if not os.path.exists(path):
        G.synth_llama(path, cfg, ftype, seed=7, source="blocks")
    names = []
    for il in range(cfg.n_layer):
        names += [f"Qcur-{il}", f"kqv_merged_cont-{il}", f"ffn_inp-{il}", f"ffn_gate_par-{il}", f"l_out-{il}", f"attn_norm-{il}", f"Vcur-{il}", f"kq_soft_max_ext-{il}"]
    names += ["result_output", "result_norm"]
    r = ref.RefModel(path, n_ctx=n_ctx)
    r.set_taps(names)
    lr = r.decode(tokens, 0)
    m = engine.Model(path); c = engine.Context(m, n_ctx)
    c.set_taps(True)
    lg = c.decode(tokens, 0)
    print(f"--- taps {cfg_name} {ftype} n_tokens={len(tokens)} logits rel={rel(lg, lr):.2e}")
    T = len(tokens)
    for il in range(cfg.n_layer):
        for nm in ("Qcur", "kqv_merged_cont", "ffn_inp", "ffn_gate_par", "l_out"):
            a = c.get_tap(nm, il); b = r.get_tap(f"{nm}-{il}")
            if a is None or b is None:
                print(f"   {nm}-{il}: missing gpu={a is None} ref={b is None}"); continue
            # the reference tensors hold all T tokens ([.., T]) except after the last layer's get_rows; ours the last token
            bl = b.reshape(T, -1)[-1] if b.size == a.size * T else b
            if nm == "Qcur" and b.size == a.size * T:
                bl = b.reshape(T, -1)[-1]
            print(f"   {nm}-{il}: rel={rel(a, bl):.2e}  (ref size {b.size}, gpu size {a.size})")
    c.close(); m.close(); r.close()


def q80_debug():
    rng = np.random.default_rng(0)
    x = rng.standard_normal(256).astype(np.float32)
    a, b = engine.op_quantize_q8_0(x).reshape(-1, 34), ref.quantize_row_q8_0(x).reshape(-1, 34)
    print("q80 diff per column:", (a != b).sum(0).tolist())
    print("gpu", a[0, :8].tolist(), "ref", b[0, :8].tolist())


if __name__ == "__main__" and "taps" in sys.argv[1:]:
    q80_debug()
    taps("llama3-8b-2l", "Q4_K_M", [5])
    taps("tiny-gqa4", "Q5_K_M", [5, 9, 200, 17, 3, 99, 42, 7, 11, 300, 1, 2])
    taps("tiny-gqa4", "Q8_0", [5])
    taps("tiny", "Q4_K_M", [5, 9, 200, 17, 3, 99, 42, 7, 11, 300, 1, 2])
