import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from booster_b200 import engine, gguf_io as G
from oracle import port
rel = lambda a, b: float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))
TMP = "/tmp/b200_models"; os.makedirs(TMP, exist_ok=True)

def run(path, prompt, n_steps, n_ctx):
    p = port.PortModelRunner(path, n_ctx=n_ctx)
    m = engine.Model(path); c = engine.Context(m, n_ctx)
    for mode in ("graph", "taps"):
        c.set_taps(mode == "taps"); c.kv_clear(); p.kv_clear()
        lg = c.decode(prompt, 0); lp = p.decode(prompt, 0)
        out = [rel(lg, lp)]
        pos = len(prompt)
        for s in range(n_steps):
            t = int(np.argmax(lp))
            lg = c.decode([t], pos); lp = p.decode([t], pos); pos += 1
            out.append(rel(lg, lp))
            if mode == "taps" and out[-1] > 1e-3 and out[-2] < 1e-3:
                for il in range(m.n_layer):
                    print(f"      first bad step {s} pos {pos-1} layer {il}: q {rel(c.get_tap('Qcur', il), p.tap_q[il]):.2e} kqv {rel(c.get_tap('kqv_merged_cont', il), p.tap_kqv[il]):.2e} l_out {rel(c.get_tap('l_out', il), p.tap_l_out[il]):.2e}")
        print(f"  {mode:5s}: " + " ".join(f"{e:.1e}" for e in out))
    # single-token from scratch
    c.set_taps(False); c.kv_clear(); p.kv_clear()
    out = []
    for i, t in enumerate(prompt[:8]):
        out.append(rel(c.decode([t], i), p.decode([t], i)))
    print("  single graph: " + " ".join(f"{e:.1e}" for e in out))
    c.close(); m.close()

gold = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
print("tiny_Q5_K_M golden"); run(os.path.join(gold, "tiny_Q5_K_M.gguf"), [5, 9, 200, 17, 3, 99, 42, 7, 11, 300, 1, 2], 6, 64)
print("tiny_Q4_K_M golden"); run(os.path.join(gold, "tiny_Q4_K_M.gguf"), [5, 9, 200, 17, 3, 99, 42, 7, 11, 300, 1, 2], 6, 64)
for cfg, ft in (("llama3-8b-2l", "Q8_0"), ("llama3-8b-2l", "Q4_K_M")):
    path = os.path.join(TMP, f"{cfg}_{ft}.gguf")
    if not os.path.exists(path): G.synth_llama(path, G.CONFIGS[cfg], ft, seed=7, source="blocks")
    print(cfg, ft); run(path, np.random.default_rng(42).integers(0, 4096, size=16).tolist(), 6, 128)
