"""Static SASS statistics of one kernel in libtrixib200.so: code size, opcode mix, encoded stall counts."""
import collections, re, subprocess, sys
pat = sys.argv[1] if len(sys.argv) > 1 else "k_line6ILi5ELi5ELb0ELi4ELi2ELi8ELb0ELb1ELb0ELi0E"
so = sys.argv[2] if len(sys.argv) > 2 else "trixicuda.jl_b200/libtrixib200.so"
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout.split("\n")
ins, on = [], False
i = 0
while i < len(out):
    l = out[i]
    if "Function :" in l:
        on = pat in l
    if on:
        m = re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);\s+/\* (0x[0-9a-f]+) \*/', l)
        if m and i + 1 < len(out):
            m2 = re.match(r'\s+/\* (0x[0-9a-f]+) \*/', out[i + 1])
            if m2:
                hi = int(m2.group(1), 16)
                ins.append((int(m.group(1), 16), m.group(2).strip(), (hi >> 41) & 0xf))
                i += 2
                continue
    i += 1
print(len(ins), "instructions,", ins[-1][0] + 16, "bytes")
hist = collections.defaultdict(collections.Counter)
for a, s, st in ins:
    op = re.sub(r'^@!?U?P\d+\s+', '', s).split()[0].split('.')[0]
    hist[op][st] += 1
for op in ("DFMA", "DMUL", "DADD", "MUFU", "LDS", "STS", "LDGSTS", "IMAD", "BRA", "BSYNC", "CALL", "LDL", "STL"):
    print(op, sum(hist[op].values()), sorted(hist[op].items()))
fp = [(st) for a, s, st in ins if re.search(r'\bD(FMA|MUL|ADD)\b', s)]
print("avg encoded stall after an FP64 instruction: %.2f" % (sum(fp) / max(len(fp), 1)))
