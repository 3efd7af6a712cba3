"""Phase timeline of ONE decode token through the persistent per-token kernel (b200_trace_phases).
usage: python scripts/trace_phases.py [pos] [out.txt] [bench config name]
Per phase kind: mean over layers of (run = dependent half, barrier = arrive..passed incl. the next phase's independent
half, skew = last CTA done - median CTA done), all in us; then the first layers phase by phase."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from booster_b200 import engine  # noqa: E402

pos = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
cfgname = sys.argv[3] if len(sys.argv) > 3 else bench.DEFAULT_CONFIG
bc = bench.BENCH_CONFIGS[cfgname]
m = engine.Model(bench.model_path(bc["model"], bc["ftype"], bc.get("share_period", 0)))
c = engine.Context(m, bc["ctx"])
out = c.generate_greedy(1, 0, 64)
r = c.trace_phases(int(out[-1]), pos, reps=3)
if r is None:
    sys.exit("the persistent kernel is not in use")
st, kinds = r
st = st.astype(np.int64)
NAMES = {1: "scores", 2: "softmax_pv", 10: "head", 11: "wo/down", 12: "qkv", 13: "gate_up"}
t0 = st[0, :, 0].min()
lines, tot = [], {}
n = len(kinds)
seen11 = 0
for p in range(n):
    name = NAMES.get(int(kinds[p]), str(kinds[p]))
    if kinds[p] == 11:
        name = "wo" if seen11 % 2 == 0 else "down"
        seen11 += 1
    s = st[p]
    start, done = s[:, 0], s[:, 1]
    arrived = s[:, 2] if p + 1 < n else done
    passed = s[:, 3] if p + 1 < n else done
    run_med = float(np.median(done - start)) / 1e3
    run_max = float((done.max() - start.min())) / 1e3
    bar = float(np.median(passed - done)) / 1e3
    span = float(passed.max() - start.min()) / 1e3
    skew = float(done.max() - np.median(done)) / 1e3
    lines.append(f"{p:4d} {name:10s} t0 {(start.min() - t0) / 1e3:9.2f}  run med {run_med:6.2f} (first start -> last done {run_max:6.2f})  "
                 f"done -> barrier passed med {bar:5.2f}  skew {skew:5.2f}  span {span:6.2f}")
    a = tot.setdefault(name, [0, 0.0, 0.0, 0.0, 0.0])
    a[0] += 1; a[1] += run_med; a[2] += run_max; a[3] += bar; a[4] += span
total = (st[n - 1, :, 1].max() - t0) / 1e3
hdr = [f"{cfgname}: token at pos {pos}: {n} phases, first start -> last done = {total:.1f} us",
       "per kind: n, mean run (median CTA), mean first-start->last-done, mean done->barrier-passed (median CTA), mean span"]
for k, (cnt, r1, r2, b, sp) in tot.items():
    hdr.append(f"  {k:10s} n {cnt:3d}  run {r1 / cnt:6.2f}  all-CTAs {r2 / cnt:6.2f}  barrier {b / cnt:5.2f}  span {sp / cnt:6.2f}   (total {sp:8.1f} us)")
text = "\n".join(hdr + lines)
print("\n".join(hdr + lines[:20]))
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(text + "\n")
