"""Join the encoded stall counts (cuobjdump) with executed counts (ncu source page) for one kernel: execution-weighted
average issue distance after FP64 instructions, and the hot dependent pairs."""
import collections, csv, re, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
so = "trixicuda.jl_b200/libtrixib200.so"
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout.split("\n")
ins, on, i = [], False, 0
while i < len(out):
    l = out[i]
    if "Function :" in l: on = pat in l
    if on:
        m = re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);\s+/\* (0x[0-9a-f]+) \*/', l)
        if m and i + 1 < len(out):
            m2 = re.match(r'\s+/\* (0x[0-9a-f]+) \*/', out[i + 1])
            if m2:
                hi = int(m2.group(1), 16)
                ins.append((int(m.group(1), 16), m.group(2).strip(), (hi >> 41) & 0xf)); i += 2; continue
    i += 1
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
h, data = rows[1], rows[2:]
ix = {k: i for i, k in enumerate(h)}
base = int(data[0][ix['Address']], 16)
ex = {}
for r in data:
    ex[int(r[ix['Address']], 16) - base] = (float(r[ix['Instructions Executed']].replace(',', '') or 0), float(r[ix['# Samples']].replace(',', '') or 0))
tot_e = tot_s = 0; hist = collections.Counter(); allst = 0; alle = 0
for a, s, st in ins:
    e, sm = ex.get(a, (0, 0))
    allst += e * max(st, 1); alle += e
    if re.search(r'\bD(FMA|MUL|ADD)\b', s):
        tot_e += e; tot_s += e * st; hist[st] += e
print("executed-weighted avg stall after FP64: %.2f" % (tot_s / tot_e))
print("FP64 exec by stall count:", {k: round(100 * v / tot_e, 1) for k, v in sorted(hist.items())})
print("sum of encoded stalls per executed instruction (min issue cycles/inst): %.2f" % (allst / alle))
byop = collections.Counter(); byop_n = collections.Counter()
for a, s, st in ins:
    e, sm = ex.get(a, (0, 0))
    op = re.sub(r'^@!?U?P\d+\s+', '', s).split()[0].split('.')[0]
    byop[op] += e * max(st, 1); byop_n[op] += e
print("encoded issue cycles by opcode (share of total, avg stall):")
for op, v in byop.most_common(16):
    print("  %-8s %5.1f%%  avg %.2f  n/iter %.0f" % (op, 100 * v / allst, v / byop_n[op], byop_n[op] / (alle / (allst and 1)) if False else byop_n[op]))
