"""Prompt and generation speed of a bridge pod split over the visible GPUs (the reference's gpu1..gpu4 proportions):
python scripts/pod_prompt_speed.py [n_gpus] [n_prompt] [config]   (BOOSTER_B200_PIPELINE_CHUNKS=0: chunks one after the other)"""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from booster_b200 import _lib  # noqa: E402

n_gpus = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n_prompt = int(sys.argv[2]) if len(sys.argv) > 2 else 1980
name = sys.argv[3] if len(sys.argv) > 3 else "8b-q4km-2048"
bc = bench.BENCH_CONFIGS[name]
path = bench.model_path(bc["model"], bc["ftype"], bc.get("share_period", 0))
os.environ["BOOSTER_B200_SPLIT"] = ",".join(["1"] * n_gpus)
L = _lib.lib()
L.init(b"", b"")
h = L.initContext(0, path.encode(), 1, 0, 100, 0, 0, 0, bc["ctx"], 32, 0, 0.0, 0.0, 0.0, 1, 1.0, 1.0, 1.0, 0, 1, 200, 1.0, 1.0, 1.0, 42, b"")
assert h
rng = np.random.default_rng(1)
for rep in range(3):
    prompt = " ".join(str(int(t)) for t in rng.integers(0, 1000, size=n_prompt if rep else 64)).encode()
    job = f"pod-{rep}".encode()
    t0 = time.perf_counter()
    ret = L.doInference(0, h, job, b"", prompt)
    dt = time.perf_counter() - t0
    pu, gu = C.c_double(), C.c_double()
    L.b200_job_timing_us(job, C.byref(pu), C.byref(gu))
    if rep:
        print(f"{name} pod over {n_gpus} GPU(s), pipeline_chunks={os.environ.get('BOOSTER_B200_PIPELINE_CHUNKS', '1')}: prompt {n_prompt} tokens "
              f"{1e6 / pu.value:.0f} tok/s ({pu.value:.1f} us/token), generation {1e6 / gu.value:.1f} tok/s, total {1e3 * dt:.1f} ms, returned {ret}", flush=True)
