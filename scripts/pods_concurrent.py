"""Aggregate generation speed of P bridge pods running at once, each split over the same G GPUs (multi-stream throughput of the
layer split: while pod A's token is on stage 1, pod B's is on stage 0): python scripts/pods_concurrent.py [n_gpus] [n_pods] [n_gen]"""
import ctypes as C
import os
import sys
import threading
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from booster_b200 import _lib  # noqa: E402

n_gpus = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n_pods = int(sys.argv[2]) if len(sys.argv) > 2 else 2
n_gen = int(sys.argv[3]) if len(sys.argv) > 3 else 256
bc = bench.BENCH_CONFIGS["8b-q4km-2048"]
path = bench.model_path(bc["model"], bc["ftype"], 0)
os.environ["BOOSTER_B200_SPLIT"] = ",".join(["1"] * n_gpus)
L = _lib.lib()
L.init(b"", b"")
pods = []
for i in range(n_pods):
    h = L.initContext(i, path.encode(), 1, 0, 100, 0, 0, 0, 1024, n_gen, 0, 0.0, 0.0, 0.0, 1, 1.0, 1.0, 1.0, 0, 1, 200, 1.0, 1.0, 1.0, 42 + i, b"")
    assert h
    pods.append(h)
rng = np.random.default_rng(3)
prompts = [" ".join(str(int(t)) for t in rng.integers(0, 1000, size=64)).encode() for _ in range(n_pods)]


def run(active, tag):
    out = [0] * n_pods
    gu = [0.0] * n_pods

    def job(i):
        j = f"{tag}-{i}".encode()
        out[i] = L.doInference(i, pods[i], j, b"", prompts[i])
        p, g = C.c_double(), C.c_double()
        L.b200_job_timing_us(j, C.byref(p), C.byref(g))
        gu[i] = g.value

    th = [threading.Thread(target=job, args=(i,)) for i in range(active)]
    t0 = time.perf_counter()
    for t in th:
        t.start()
    for t in th:
        t.join()
    dt = time.perf_counter() - t0
    return dt, gu[:active], out[:active]


run(n_pods, "warm")
for active in sorted(set([1, n_pods])):
    dt, gu, out = run(active, f"run{active}")
    per = [1e6 / g for g in gu]
    print(f"{active} pod(s) over {n_gpus} GPU(s): per-pod generation {', '.join(f'{x:.1f}' for x in per)} tok/s, aggregate {sum(per):.1f} tok/s "
          f"(wall {1e3 * dt:.0f} ms for {active} x (64 prompt + {n_gen} generated))", flush=True)
