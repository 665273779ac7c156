"""Per-instruction stall samples of an .ncu-rep source page: top instructions by samples and totals per stall reason
and per opcode class.   python tools/ncu_hot.py rep.ncu-rep [top]"""
import csv, subprocess, sys, collections, re
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]; data = rows[hi + 1:]
ci = {k: i for i, k in enumerate(h)}
reasons = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
tot = collections.Counter(); byop = collections.Counter(); byop_n = collections.Counter()
recs = []
for r in data:
    if len(r) < len(h): continue
    n = int(r[ci["# Samples"]] or 0)
    src = r[ci["Source"]]
    op = re.sub(r'^@!?U?P\d+\s+', '', src).split()[0].split('.')[0] if src else "?"
    ex = int(r[ci["Instructions Executed"]] or 0)
    st = {k: int(r[ci[k]] or 0) for k in reasons}
    for k, v in st.items(): tot[k] += v
    byop[op] += n; byop_n[op] += ex
    recs.append((n, r[ci["Address"]], src, ex, st))
S = sum(tot.values())
print("total samples", S)
for k, v in tot.most_common(12): print(f"  {k:24s}{100*v/S:6.1f} %")
print("by opcode (samples %, executed warp-instr %):")
E = sum(byop_n.values())
for k, v in byop.most_common(16): print(f"  {k:10s}{100*v/S:6.1f} %   {100*byop_n[k]/E:6.1f} %")
print("top instructions:")
for n, a, src, ex, st in sorted(recs, reverse=True)[:top]:
    big = ", ".join(f"{k[6:]}={v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3] if v)
    print(f"  {n:6d} {a[-5:]} ex={ex:8d} {src[:60]:60s} {big}")
