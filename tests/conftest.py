"""pytest configuration: `gpu` marker (tests that need a CUDA device + the built libbooster_b200.so) and shared
helpers. `-m "not gpu"` covers the oracle against the golden vectors, host logic and the C-ABI surface;
`-m gpu` are the parity tests proper and call through the C-ABI."""
import os
import sys

# the CPU oracles are OpenMP code: on a 128-core GPU host the default thread count oversubscribes badly
os.environ.setdefault("OMP_NUM_THREADS", "16")

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def rel_err(a, b) -> float:
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def model_dir(tmp_path_factory):
    d = os.environ.get("B200_TMP") or str(tmp_path_factory.mktemp("b200_models"))
    os.makedirs(d, exist_ok=True)
    return d


@pytest.fixture(scope="session")
def ref_or_none():
    """the real reference CPU library (oracle/_ref) if it is present and loadable on this host, else None"""
    try:
        from oracle import ref
        if not ref.available():
            return None
        ref.lib()
        return ref
    except Exception:
        return None


def q8k_equal(a, b) -> bool:
    """block_q8_K streams are equal, ignoring the bsums of all-zero blocks: the reference leaves them unwritten
    there (quantize_row_q8_K_ref zero-block branch, cpp/ggml/src/ggml-quants.c:3607-3612; they are multiplied by d = 0)."""
    a = np.asarray(a, dtype=np.uint8).reshape(-1, 292).copy()
    b = np.asarray(b, dtype=np.uint8).reshape(-1, 292).copy()
    zero = (b[:, :4] == 0).all(axis=1) & (b[:, 4:260] == 0).all(axis=1)
    a[zero, 260:] = 0
    b[zero, 260:] = 0
    return bool(np.array_equal(a, b))


def greedy_consistent(ours, theirs) -> bool:
    """arg-max agreement up to a provable near-tie: either the ids are equal, or the reference's margin between
    its own arg-max and ours is smaller than twice the max-abs deviation of our logits from the reference's at this
    step (a deviation the flip bound allows) — i.e. the two candidates are tied within the arithmetic noise."""
    ours = np.asarray(ours, dtype=np.float64)
    theirs = np.asarray(theirs, dtype=np.float64)
    a, t = int(np.argmax(ours)), int(np.argmax(theirs))
    if a == t:
        return True
    return bool(theirs[t] - theirs[a] <= 2.0 * np.abs(ours - theirs).max())
