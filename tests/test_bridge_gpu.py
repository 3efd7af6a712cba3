"""The drop-in boundary end to end: the nine cgo entry points (include/bridge.h == cpp/bridge.h:132-165) driven
exactly as pkg/server/server.go drives them (init -> initContext -> doInference with a concurrent status poller
-> timing/promptEval/getPromptTokenCount), on a no_vocab model whose prompt is a list of token ids."""
import ctypes as C
import os
import threading
import time

import numpy as np
import pytest

from booster_b200 import _lib
from oracle import port

pytestmark = pytest.mark.gpu


def _init(path, n_ctx=64, predict=8, idx=0):
    L = _lib.lib()
    L.init(b"", b"")
    return L, L.initContext(idx, path.encode(), 1, 0, 100, 0, 0, 0, n_ctx, predict, 0, 0.0, 0.0, 0.0, 1, 1.0, 1.0,
                            1.0, 0, 1, 200, 1.0, 1.0, 1.0, 42, b"")


def test_do_inference_matches_oracle_greedy(golden_dir):
    path = os.path.join(golden_dir, "tiny_Q4_K_M.gguf")
    g = np.load(os.path.join(golden_dir, "tiny_Q4_K_M.npz"))
    L, ctx = _init(path, predict=8)
    assert ctx
    prompt = " ".join(str(t) for t in g["prompt"].tolist())
    seen = []
    stop = threading.Event()

    def poll():                       # status() is called concurrently from other goroutines (router.go:123)
        while not stop.is_set():
            seen.append(L.status(b"job-1"))
            time.sleep(0.0005)

    th = threading.Thread(target=poll); th.start()
    n = L.doInference(0, ctx, b"job-1", b"", prompt.encode())
    stop.set(); th.join()
    text = L.status(b"job-1").decode()
    ids = [int(x) for x in text.split()]
    assert ids[:12] == g["prompt"].tolist()                       # status = prompt pieces + generated pieces
    assert ids[12:] == g["ids"].tolist()                          # greedy ids == the reference's (bit-exact logits)
    assert n == 12 + 7                                            # n_p_eval + n_eval (the last sampled token is not decoded)
    assert L.getPromptTokenCount(b"job-1") == 12
    assert L.timing(b"job-1") >= 0 and L.promptEval(b"job-1") >= 0
    assert L.getSeed(b"job-1") == 42
    pu, gu = C.c_double(), C.c_double()
    assert L.b200_job_timing_us(b"job-1", C.byref(pu), C.byref(gu)) == 0 and gu.value > 0
    assert all(s is not None for s in seen)


def test_prompt_too_long_returns_zero(golden_dir):
    path = os.path.join(golden_dir, "tiny_Q4_K_M.gguf")
    L, ctx = _init(path, n_ctx=32, predict=4, idx=1)
    assert ctx
    prompt = " ".join(["7"] * 40)                                 # > n_ctx - 4 (cpp/bridge.cpp:382-386)
    assert L.doInference(1, ctx, b"job-long", b"", prompt.encode()) == 0
    assert L.getPromptTokenCount(b"job-long") == 40


def test_stop_inference(golden_dir):
    path = os.path.join(golden_dir, "tiny_Q8_0.gguf")
    L, ctx = _init(path, n_ctx=64, predict=-1, idx=2)
    assert ctx
    t = threading.Timer(0.02, lambda: L.stopInference(2))
    t.start()
    n = L.doInference(2, ctx, b"job-stop", b"", b"1 2 3")
    t.join()
    assert 0 < n <= 64


def test_bad_model_path_returns_null():
    L = _lib.lib()
    L.init(b"", b"")
    assert not L.initContext(3, b"/nonexistent.gguf", 1, 0, 100, 0, 0, 0, 64, 8, 0, 0.0, 0.0, 0.0, 1, 1.0, 1.0, 1.0,
                             0, 1, 200, 1.0, 1.0, 1.0, 42, b"")
