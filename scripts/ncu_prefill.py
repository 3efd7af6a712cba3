"""Target for ncu: one 512-token prompt batch through the batched prompt kernels on the synthetic 8B model of a bench config.
usage: python scripts/ncu_prefill.py [bench config name]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from booster_b200 import engine  # noqa: E402

bc = bench.BENCH_CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "8b-q8_0-prefill512"]
m = engine.Model(bench.model_path(bc["model"], bc["ftype"], bc.get("share_period", 0)))
c = engine.Context(m, 1536)
prompt = np.random.default_rng(42).integers(0, 128256, size=512).tolist()
c.decode(prompt, 0, want_logits=False)
print("done")
