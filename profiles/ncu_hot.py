"""Top stalled SASS instructions of an .ncu-rep (source page, warp-stall sampling): python profiles/ncu_hot.py file.ncu-rep [N]"""
import csv
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.reader(lines[start:]))
hdr = rows[0]
iS, iSrc, iAll = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
body = [r for r in rows[1:] if len(r) == len(hdr)]
tot = sum(int(r[iS]) for r in body)
print(f"total samples {tot}")
agg = {}
for r in body:
    for i in stall_cols:
        agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i] or 0)
print("stall reasons:", ", ".join(f"{k[6:]}={v}" for k, v in sorted(agg.items(), key=lambda x: -x[1]) if v))
for idx, r in sorted(enumerate(body), key=lambda x: -int(x[1][iS]))[:n]:
    top = sorted(((int(r[i] or 0), hdr[i][6:]) for i in stall_cols), reverse=True)[:2]
    print(f"{int(r[iS]):7d} {100 * int(r[iS]) / tot:5.1f}%  #{idx:5d} {r[iSrc].strip()[:90]:90s} {top}")
