"""Print the hottest SASS instructions (by warp-stall samples) of one kernel of an .ncu-rep source page.
usage: python scripts/ncu_hot.py file.ncu-rep launch_index [n]"""
import csv, subprocess, sys, io
rep, idx = sys.argv[1], int(sys.argv[2])
n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(idx), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
print(rows[0][:2])
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) >= len(hdr) - 2]
def iv(r, name):
    try: return int(r[ci[name]] or 0)
    except (ValueError, IndexError): return 0
tot = sum(iv(r, "# Samples") for r in data)
ninst = sum(iv(r, "Instructions Executed") for r in data)
print("sass lines", len(data), "samples", tot, "warp-instr", ninst)
stalls = [h for h in hdr if h.startswith("stall_")]
agg = {s: sum(iv(r, s) for r in data) for s in stalls}
print([(k, v) for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]])
for r in sorted(data, key=lambda r: -iv(r, "# Samples"))[:n]:
    print(f"{data.index(r):5d} {r[ci['Source']].strip()[:64]:64s} smp {iv(r,'# Samples'):5d} exec {iv(r,'Instructions Executed'):7d} "
          f"lsb {iv(r,'stall_long_sb'):4d} wait {iv(r,'stall_wait'):4d} ssb {iv(r,'stall_short_sb'):4d} bar {iv(r,'stall_barrier'):4d}")
