import os, sys, threading
sys.path.insert(0, os.getcwd())
from booster_b200 import _lib
L = _lib.lib()
L.init(b"", b"")
path = os.path.join(os.getcwd(), "tests/golden", sys.argv[1] if len(sys.argv) > 1 else "tiny_Q8_0.gguf")
ctx = L.initContext(2, path.encode(), 1, 0, 100, 0, 0, 0, 64, -1, 0, 0.0, 0.0, 0.0, 1, 1.0, 1.0, 1.0, 0, 1, 200, 1.0, 1.0, 1.0, 42, b"")
assert ctx
n = L.doInference(2, ctx, b"job-stop", b"", b"1 2 3")
print("returned", n, L.status(b"job-stop"))
