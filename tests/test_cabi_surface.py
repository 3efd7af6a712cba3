"""The drop-in boundary without a GPU: libbooster_b200.so loads, exports every symbol include/*.h declares,
and — because there is no CPU fallback — every compute entry point fails LOUDLY when no CUDA device exists."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from booster_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", src)
    skip = {"defined", "sizeof"}
    return sorted({n for n in names if n not in skip and not n.isupper()})


def _exported():
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    return {line.split()[-1] for line in out.splitlines() if " T " in line}


def test_library_is_built():
    assert os.path.exists(_lib.LIB_PATH), "run `make` / __graft_entry__.build() first"


def test_bridge_symbols_match_reference_header():
    decl = _declared("bridge.h")
    # exactly the nine entry points of cpp/bridge.h:132-165
    assert decl == sorted(_lib.BRIDGE_SYMBOLS)
    assert set(decl) <= _exported()


def test_additive_symbols_exported():
    decl = _declared("booster_b200.h")
    assert set(decl) == set(_lib.B200_SYMBOLS), set(decl) ^ set(_lib.B200_SYMBOLS)
    missing = set(decl) - _exported()
    assert not missing, missing


def test_prototypes_bind():
    L = _lib.lib()
    assert L.b200_version().startswith(b"booster_b200")
    assert L.b200_device_count() >= 0


def _no_gpu():
    return _lib.lib().b200_device_count() == 0


@pytest.mark.skipif(not _no_gpu(), reason="only meaningful without a CUDA device")
def test_no_cpu_fallback(golden_dir):
    """the product path must fail loudly when the GPU is missing — never route through a CPU implementation"""
    L = _lib.lib()
    path = os.path.join(golden_dir, "tiny_Q4_K_M.gguf").encode()
    assert not L.b200_model_load(path, 0, 0, -1)
    assert b"no CUDA device" in L.b200_last_error()
    x = np.zeros(256, dtype=np.float32)
    out = np.zeros(292, dtype=np.uint8)
    assert L.b200_op_quantize_q8_K(x.ctypes.data_as(C.POINTER(C.c_float)), 256, out.ctypes.data) != 0
    assert b"no CUDA device" in L.b200_last_error()
    L.init(b"", b"")
    ctx = L.initContext(0, path, 1, 0, 100, 0, 0, 0, 64, 8, 0, 0.0, 0.0, 0.0, 1, 1.0, 1.0, 1.0, 0, 1, 200, 1.0, 1.0, 1.0, 42, b"")
    assert not ctx
    assert L.doInference(0, None, b"job", b"", b"1 2 3") == 0
    assert L.status(b"nope") == b""
    assert L.timing(b"nope") == 0 and L.promptEval(b"nope") == 0 and L.getPromptTokenCount(b"nope") == 0


def test_sass_shows_the_hardware_paths():
    """the shipped library is sm_100a code whose hot kernels use what DESIGN.md says they use: TMA bulk copies + mbarriers in
    the mat-vec / batch kernels, mma.sync HMMA in the default prompt-batch kernels, tcgen05.mma / TMEM loads in k_umma_batch
    (profiles/r02h_sass_summary.txt is the same listing per kernel)"""
    import shutil
    import subprocess
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    lib = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "booster_b200", "libbooster_b200.so")
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    assert "sm_100a" in sass
    for mnemonic in ("UBLKCP", "SYNCS", "HMMA.16816.F32", "LDSM", "UTCHMMA", "UTCBAR", "LDTM", "IDP.4A"):
        assert mnemonic in sass, mnemonic


def test_plain_c_client_compiles_links_and_runs(tmp_path):
    """the cgo caller's view (pkg/server/server.go:7-36) without a Go toolchain: tests/c_harness/bridge_client.c includes both
    headers from a C translation unit (-std=c11 -Wall -Werror -pedantic: the boundary is plain C), links against the shared
    library like `#cgo LDFLAGS: -lbooster_b200` and drives the nine symbols; without a model every call answers with the
    reference's failure values instead of crashing"""
    import shutil
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler on PATH")
    libdir = os.path.dirname(_lib.LIB_PATH)
    exe = str(tmp_path / "bridge_client")
    subprocess.run([cc, "-std=c11", "-D_DEFAULT_SOURCE", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "c_harness", "bridge_client.c"), "-o", exe,
                    "-L", libdir, "-lbooster_b200", "-lpthread", f"-Wl,-rpath,{libdir}"], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ok: bad configuration handled" in r.stdout and "version: booster_b200" in r.stdout
