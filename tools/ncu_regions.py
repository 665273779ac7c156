"""Stall samples of an .ncu-rep source page bucketed by code address (bucket size in bytes).
   python tools/ncu_regions.py rep.ncu-rep [bucket=0x200]"""
import csv, subprocess, sys, collections, re
rep = sys.argv[1]; bucket = int(sys.argv[2], 0) if len(sys.argv) > 2 else 0x200
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]; data = [r for r in rows[hi + 1:] if len(r) >= len(h)]
ci = {k: i for i, k in enumerate(h)}
reasons = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
base = min(int(r[ci["Address"]], 16) for r in data)
B = collections.defaultdict(lambda: [0, 0, 0, collections.Counter(), collections.Counter()])
tot = 0
for r in data:
    a = int(r[ci["Address"]], 16) - base
    b = B[a // bucket]
    n = int(r[ci["# Samples"]] or 0); ex = int(r[ci["Instructions Executed"]] or 0)
    b[0] += n; b[1] += ex; tot += n
    src = r[ci["Source"]]
    if re.search(r'\bD(FMA|MUL|ADD)\b', src): b[2] += ex
    for k in reasons: b[3][k[6:]] += int(r[ci[k]] or 0)
print("offset   samples%  warp-instr  fp64-instr  top stall reasons")
for k in sorted(B):
    n, ex, fp, st, _ = B[k]
    if n == 0 and ex == 0: continue
    top = ", ".join(f"{a}={100*v/max(n,1):.0f}%" for a, v in st.most_common(4))
    print(f"{k*bucket:06x}  {100*n/tot:6.2f}  {ex:10d}  {fp:10d}  {top}")
