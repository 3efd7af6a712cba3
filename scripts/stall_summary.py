"""Per-kernel warp-stall breakdown (source-page sampling of an .ncu-rep, read here without a GPU):
python scripts/stall_summary.py file.ncu-rep"""
import collections
import csv
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
secs = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
seen = 0
for k in range(0, len(secs) - 1, 2):          # every kernel appears twice (SASS view, source view)
    a, b = secs[k], secs[k + 1]
    hdr, body = rows[a + 1], rows[a + 2:b]
    cols = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
    tot = collections.Counter()
    for r in body:
        for c in cols:
            tot[c] += int(r[hdr.index(c)] or 0)
    s = sum(tot.values()) or 1
    ie = hdr.index("Instructions Executed")
    n_inst = sum(int(r[ie]) for r in body)
    n_static = sum(1 for r in body if int(r[ie]) > 0)
    seen += 1
    print(f"#{seen} {rows[a][1][:60]}: {s} samples, {n_inst} warp-instructions executed, {n_static} distinct SASS instructions executed of {len(body)}")
    print("    " + "  ".join(f"{c[6:]} {100 * v / s:.1f}%" for c, v in tot.most_common(9)))
