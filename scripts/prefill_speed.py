"""Prompt-batch speed on a bench config's synthetic model: python scripts/prefill_speed.py [config] [n_prompt] [reps]
Prints tokens/s of b200_decode(prompt batch) (wall clock incl. the final synchronisation) and a checksum of the logits."""
import os
import sys
import time
import zlib

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from booster_b200 import engine  # noqa: E402
from booster_b200 import gguf_io as G  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "8b-q4km-2048"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 512
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
bc = bench.BENCH_CONFIGS[name]
m = engine.Model(bench.model_path(bc["model"], bc["ftype"], bc.get("share_period", 0)))
c = engine.Context(m, max(n + 64, 1536))
prompt = np.random.default_rng(42).integers(0, G.CONFIGS[bc["model"]].n_vocab, size=n).tolist()
for r in range(reps):
    c.kv_clear()
    t0 = time.perf_counter()
    lg = c.decode(prompt, 0)
    dt = time.perf_counter() - t0
    print(f"{name} prefill {n} tokens: {1e3 * dt:.2f} ms  {n / dt:.0f} tok/s  logits crc {zlib.crc32(lg.tobytes()):08x}", flush=True)
