"""Target for ncu: load the synthetic 8B Q4_K_M model and run a few UN-GRAPHED decode tokens at n_kv ~ 2000
(stage API: plain kernel launches, no CUDA graph), so that `ncu -k regex:... -s ... -c ...` can pick launches.
Launch order per token: [k_embed] + 32 x [k_matvec<QKV>, k_attn_scores, k_attn_softmax_pv, k_matvec<RESID> (wo),
k_matvec<SILU> (gate/up), k_matvec<RESID> (down)] + k_matvec<STORE> (head) + k_argmax_partial + k_argmax_finish
= 196 kernels, 129 of them k_matvec."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from booster_b200 import engine  # noqa: E402

n_tokens = int(sys.argv[1]) if len(sys.argv) > 1 else 3
m = engine.Model(bench.model_path())
c = engine.Context(m, 2048)
tok = C.c_int32(1)
for i in range(n_tokens):
    assert m.L.b200_stage_forward(c.h, tok.value, 2000 + i, 0, None) == 0
    assert m.L.b200_stage_argmax(c.h, C.byref(tok)) == 0
print("done, last token", tok.value)
