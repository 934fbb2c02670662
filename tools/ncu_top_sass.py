"""Top stalled SASS instructions of the first kernel in an .ncu-rep:  python tools/ncu_top_sass.py file.ncu-rep [n=25] [context=0]"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
ctx = int(sys.argv[3]) if len(sys.argv) > 3 else 0
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]; body = rows[2:]
iS, iN, iE = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[iN]) for r in body)
print(f"{rows[0][1][:100]}: {len(body)} instructions, {tot} samples, {sum(int(r[iE]) for r in body)} warp instructions executed")
order = sorted(range(len(body)), key=lambda i: -int(body[i][iN]))[:n]
for i in order:
    r = body[i]
    why = sorted(((int(r[j]), hdr[j][6:]) for j in stall), reverse=True)[:2]
    print(f"{100 * int(r[iN]) / tot:5.1f}%  exec {int(r[iE]):>9}  line {i:5d}  {r[iS].strip()[:90]:90s} {why}")
    for k in range(max(0, i - ctx), i):
        print(f"          {'':>14}  line {k:5d}  {body[k][iS].strip()[:90]}")
