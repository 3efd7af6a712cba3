"""Phase timeline of ONE graph-replayed decode token from the kernels' own %globaltimer stamps (b200_trace_token).
usage: python scripts/trace_token.py [pos] [out.txt]   — prints per layer-kernel: gap to the previous kernel's end,
phase offsets (median over CTAs) and the kernel's span; then the per-kind totals."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from booster_b200 import engine  # noqa: E402

pos = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
m = engine.Model(bench.model_path())
c = engine.Context(m, 2048)
tok = 1
out = c.generate_greedy(tok, 0, 64)         # some KV content + warm
st, meta = c.trace_token(int(out[-1]), pos, reps=3)
KINDS = ["embed", "qkv", "attn", "wo", "gate_up", "down", "head", "pv"]
lines = []
t_first = None
prev_end = None
tot = {}
for i in range(len(st)):
    kind, ctas = int(meta[i, 0]), int(meta[i, 1])
    s = st[i, :ctas, :].astype(np.int64)
    # phase columns in time order: matvec 0 start,1 ring part-filled,2 after wait,5 x consumed,6 norm scale,3 prologue done,4 end
    #   7 first tile landed, 8 its integers done, 9 first round published (G>1), 11 first round's chains done, 10 warp 0 out of work
    order = [0, 1, 2] if kind == 2 else ([0, 1, 2, 3, 4] if kind == 7 else [0, 1, 2, 5, 6, 3, 7, 8, 9, 11, 10, 4])
    nph = len(order)
    s = s[s[:, order[-1]] > 0]                   # CTAs that exited early (beyond n_kv) have no end stamp
    s = s[:, order]
    for p in range(1, nph):                      # phases a kernel variant does not stamp: carry the previous one
        s[:, p] = np.where(s[:, p] > 0, s[:, p], s[:, p - 1])
    start = s[:, 0]
    end = s[:, nph - 1]
    if t_first is None:
        t_first = start.min()
    k0, k1 = start.min(), end.max()
    gap = (k0 - prev_end) if prev_end is not None else 0
    ph = [float(np.median(s[:, p] - k0)) / 1e3 for p in range(nph)]
    phmax = [float(np.max(s[:, p] - k0)) / 1e3 for p in range(nph)]
    lines.append(f"{i:4d} {KINDS[kind]:8s} ctas {ctas:3d} t0 {(k0 - t_first) / 1e3:9.2f} us  gap_prev_end {gap / 1e3:6.2f}  span {(k1 - k0) / 1e3:6.2f}  "
                 f"phase med " + " ".join(f"{x:6.2f}" for x in ph) + "  | max " + " ".join(f"{x:6.2f}" for x in phmax))
    ctas = len(s)
    a = tot.setdefault(KINDS[kind], [0, 0.0, 0.0, np.zeros(nph)])
    a[0] += 1; a[1] += (k1 - k0) / 1e3; a[2] += gap / 1e3; a[3] += np.array(ph)
    prev_end = k1
total = (prev_end - t_first) / 1e3
hdr = [f"token at pos {pos}: {len(st)} traced launches, first start -> last end = {total:.1f} us",
       "per kind: n, mean span us, mean gap (start - previous kernel's end; negative = PDL overlap), mean phase offsets (median CTA)"]
for k, (n, sp, gp, ph) in tot.items():
    hdr.append(f"  {k:8s} n {n:3d}  span {sp / n:7.2f}  gap {gp / n:6.2f}  phases " + " ".join(f"{x / n:6.2f}" for x in ph))
text = "\n".join(hdr + lines)
print("\n".join(hdr + lines[:14]))
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(text + "\n")
