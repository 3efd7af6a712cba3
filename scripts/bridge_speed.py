"""Generation speed THROUGH the nine bridge symbols (doInference) on the bench model: µs per generated token from
b200_job_timing_us. usage: python scripts/bridge_speed.py [n_predict]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from booster_b200 import _lib  # noqa: E402

n_predict = int(sys.argv[1]) if len(sys.argv) > 1 else 256
L = _lib.lib()
L.init(b"", b"")
ctx = L.initContext(0, bench.model_path().encode(), 1, 0, 100, 0, 0, 0, 2048, n_predict, 0, 0.0, 0.0, 0.0, 1, 1.0, 1.0, 1.0, 0, 1, 200, 1.0, 1.0, 1.0, 42, b"")
assert ctx
prompt = " ".join(str(7 + i) for i in range(32)).encode()
for job in (b"warm", b"timed"):
    n = L.doInference(0, ctx, job, b"", prompt)
pu, gu = C.c_double(), C.c_double()
L.b200_job_timing_us(b"timed", C.byref(pu), C.byref(gu))
print(f"doInference: {n} tokens, prompt {pu.value:.1f} us/token ({1e6 / pu.value:.1f} tok/s), generation {gu.value:.1f} us/token ({1e6 / gu.value:.1f} tok/s)")
