"""The launch shape the engine picks for every mat-vec of the BASELINE.json models (warps per CTA, warps sharing a 32-row
unit, ring stages per warp) — pure host arithmetic (b200_op_launch_shape), pinned here so that a change of the selection
rule or of the shared-memory budget shows up as a deliberate diff. DESIGN.md §4 explains the half-occupied Q6_K shapes."""
import ctypes as C

import pytest

from booster_b200 import _lib

Q4_K, Q5_K, Q6_K, Q8_0 = 12, 13, 14, 8
SHAPES = [
    # name,                          types,        units, k,     norm, (W, G, S)
    ("8B qkv, all Q4_K",             [Q4_K],        192,  4096, 1, (16, 8, 2)),
    ("8B qkv, attn_v Q6_K",          [Q4_K, Q6_K],  192,  4096, 1, (8, 4, 4)),      # 6.7 KB tiles: half the warps fit
    ("8B wo",                        [Q4_K],        128,  4096, 0, (16, 16, 2)),
    ("8B gate|up",                   [Q4_K],        896,  4096, 1, (14, 2, 3)),
    ("8B ffn_down Q4_K",             [Q4_K],        128, 14336, 0, (14, 14, 2)),
    ("8B ffn_down Q6_K",             [Q6_K],        128, 14336, 0, (8, 8, 3)),      # round-2 item 1
    ("8B output Q6_K",               [Q6_K],       4008,  4096, 1, (16, 1, 2)),
    ("70B qkv, attn_v Q5_K",         [Q4_K, Q5_K],  320,  8192, 1, (12, 4, 3)),
    ("70B gate|up",                  [Q4_K],       1792,  8192, 1, (14, 1, 3)),
    ("70B ffn_down Q6_K",            [Q6_K],        256, 28672, 0, (8, 4, 3)),
    ("Mistral-7B Q5_K_M gate|up",    [Q5_K],        896,  4096, 1, (16, 2, 2)),
    ("8B Q8_0 gate|up",              [Q8_0],        896,  4096, 1, (16, 2, 4)),
]


@pytest.mark.parametrize("name,types,units,k,norm,want", SHAPES, ids=[s[0] for s in SHAPES])
def test_launch_shape(name, types, units, k, norm, want):
    L = _lib.lib()
    t = (C.c_int32 * len(types))(*types)
    out = (C.c_int32 * 3)()
    assert L.b200_op_launch_shape(t, len(types), units, k, norm, 148, out) == 0
    W, G, S = tuple(out)
    assert (W, G, S) == want
    tiles_unit = k // (32 if types[0] == Q8_0 else 256)
    assert W % G == 0 and tiles_unit % G == 0 and S >= 2          # the kernel's structural requirements
    assert G == 1 or units * G <= 148 * W                          # K-split only inside one wave


def test_launch_shape_rejects_oversized_vectors():
    L = _lib.lib()
    t = (C.c_int32 * 1)(Q6_K)
    out = (C.c_int32 * 3)()
    assert L.b200_op_launch_shape(t, 1, 128, 256 * 1024, 0, 148, out) != 0
    assert b"shared-memory" in L.b200_last_error()
