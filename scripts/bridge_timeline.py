"""Timeline of a doInference job as the status() poller sees it (diagnostic for the e2e leg of bench.py):
python scripts/bridge_timeline.py [config] [n_prompt] [n_gen]"""
import ctypes as C
import os
import sys
import threading
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from booster_b200 import _lib  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "8b-q4km-2048"
bc = bench.BENCH_CONFIGS[name]
ctx = bc["ctx"]
n_gen = int(sys.argv[3]) if len(sys.argv) > 3 else 64
n_prompt = int(sys.argv[2]) if len(sys.argv) > 2 else ctx - 4 - n_gen
path = bench.model_path(bc["model"], bc["ftype"], bc.get("share_period", 0))
L = _lib.lib()
L.init(b"", b"")
h = L.initContext(7, path.encode(), 1, 0, 100, 0, 0, 0, ctx, n_gen, 0, 0.0, 0.0, 0.0, 1, 1.0, 1.0, 1.0, 0, 1, 200, 1.0, 1.0, 1.0, 42, b"")
assert h
rng = np.random.default_rng(4242)
for rep in range(3):
    npr = n_prompt if rep else 16
    prompt = " ".join(str(int(t)) for t in rng.integers(0, 1000, size=npr)).encode()
    job = f"tl-{rep}".encode()
    ev = []
    done = threading.Event()

    def poll():
        last = -1
        while not done.is_set():
            n = L.status(job).count(b" ")
            if n != last:
                ev.append((time.perf_counter(), n)); last = n
            time.sleep(0.0002)

    th = threading.Thread(target=poll); th.start()
    t0 = time.perf_counter()
    ret = L.doInference(7, h, job, b"", prompt)
    t_end = time.perf_counter()
    done.set(); th.join()
    pu, gu = C.c_double(), C.c_double()
    L.b200_job_timing_us(job, C.byref(pu), C.byref(gu))
    print(f"rep {rep}: prompt {npr} ret {ret} total {1e3 * (t_end - t0):.1f} ms; job timing prompt {pu.value:.1f} us/tok gen {gu.value:.1f} us/tok")
    prev = t0
    for t, n in ev:
        if t - prev > 0.004 or n <= npr + 2 or n >= npr + n_gen - 1:
            print(f"   +{1e3 * (t - t0):9.2f} ms  n={n}  (gap {1e3 * (t - prev):.2f} ms)")
        prev = t
    print(f"   end +{1e3 * (t_end - t0):.2f} ms (gap {1e3 * (t_end - prev):.2f} ms)")
