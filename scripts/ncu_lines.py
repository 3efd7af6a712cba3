"""Hot source lines of an .ncu-rep (cuda,sass correlated view): python scripts/ncu_lines.py file.ncu-rep [n]"""
import csv
import subprocess
import sys

n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))


def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


cur, agg, hdr = None, [], None
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        cur = r[1].split('/')[-1]
    elif r[0] == 'Line No':
        hdr = r
    elif r[0].isdigit():
        agg.append((cur, int(r[0]), r[1], num(r[6]), num(r[7]), r))
tot_s = sum(a[3] for a in agg) or 1
tot_i = sum(a[4] for a in agg) or 1
sc = [i for i, h in enumerate(hdr) if h.startswith('stall_')]
print("samples", tot_s, "warp-instructions", tot_i)
print("\nTOP BY SAMPLES (share of samples, share of executed instructions, top stall reasons)")
for a in sorted(agg, key=lambda a: -a[3])[:n]:
    r = a[5]
    st = sorted([(num(r[i]), hdr[i][6:]) for i in sc], reverse=True)[:3]
    print(f"{100 * a[3] / tot_s:5.1f}% ex {100 * a[4] / tot_i:5.1f}% {a[0][:16]}:{a[1]:4d} {a[2].strip()[:84]}  | " + " ".join(f"{k}:{v}" for v, k in st))
print("\nTOP BY EXECUTED INSTRUCTIONS")
for a in sorted(agg, key=lambda a: -a[4])[:n]:
    print(f"{100 * a[4] / tot_i:5.1f}% smp {100 * a[3] / tot_s:5.1f}% {a[0][:16]}:{a[1]:4d} {a[2].strip()[:96]}")
