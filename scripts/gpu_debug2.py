import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from booster_b200 import engine, gguf_io as G
from oracle import port
rel = lambda a, b: float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))
TMP = "/tmp/b200_models"; os.makedirs(TMP, exist_ok=True)

def run(path, toks, n_ctx):
    p = port.PortModelRunner(path, n_ctx=n_ctx)
    m = engine.Model(path); c = engine.Context(m, n_ctx)
    c.set_taps(True)
    for i, t in enumerate(toks):
        lg = c.decode([t], i); lp = p.decode([t], i)
        line = f"  pos {i}: logits {rel(lg, lp):.1e} |"
        for il in range(m.n_layer):
            line += f" L{il}: q {rel(c.get_tap('Qcur', il), p.tap_q[il]):.1e} kqv {rel(c.get_tap('kqv_merged_cont', il), p.tap_kqv[il]):.1e} l_out {rel(c.get_tap('l_out', il), p.tap_l_out[il]):.1e} |"
        print(line)
        # compare the f16 caches of layer 0 at this position (engine cache is not exposed: compare via kqv only)
    c.close(); m.close()

gold = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
print("tiny_Q5_K_M golden"); run(os.path.join(gold, "tiny_Q5_K_M.gguf"), [5, 9, 200, 17, 3, 99, 42], 64)
for cfg, ft in (("llama3-8b-2l", "Q4_K_M"), ("llama3-8b-2l", "Q8_0")):
    path = os.path.join(TMP, f"{cfg}_{ft}.gguf")
    if not os.path.exists(path): G.synth_llama(path, G.CONFIGS[cfg], ft, seed=7, source="blocks")
    print(cfg, ft); run(path, np.random.default_rng(42).integers(0, 4096, size=4).tolist(), 128)
