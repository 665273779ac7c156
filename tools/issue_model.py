"""Additive issue-cost model of a kernel from an .ncu-rep (source page: executed warp instructions per SASS line).

    python tools/issue_model.py rep.ncu-rep [n_smsp=592]

Every executed warp instruction is charged the issue interval of its pipe as measured by the probes in tools/
(FP64 2 cycles, 3 with three distinct register sources; IMAD 2; 64/128-bit shared-memory and global accesses 2;
MUFU 4; everything else 1), the charges are summed per SM sub-partition and compared with the measured duration.
profiles/r1_line6_notes.md uses this to show that k_line6 is bound by the SUM of its issue costs."""
import collections, csv, re, subprocess, sys

rep = sys.argv[1]
nsmsp = int(sys.argv[2]) if len(sys.argv) > 2 else 592
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]; ci = {k: i for i, k in enumerate(h)}
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
rh, rv = rr[0], rr[2]
cycles = float(rv[rh.index("sm__cycles_elapsed.max")].replace(",", ""))


def cost(s):
    op = re.sub(r'^@!?U?P\d+\s+', '', s).split()[0]
    base = op.split('.')[0]
    if base in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"):
        regs = set(re.findall(r'(?<![U\w])R(\d+)', s.split(None, 1)[1].split(',', 1)[1] if ',' in s else ""))
        return ("fp64_3reg", 3) if len(regs) >= 3 else ("fp64", 2)
    if base == "IMAD":
        return ("imad", 2)
    if base == "MUFU":
        return ("mufu", 4)
    if base in ("LDS", "STS", "LDG", "STG", "LDL", "STL", "LDGSTS", "LDC", "ATOMS", "RED", "SHFL"):
        return ("memory", 2)
    return ("other", 1)


cnt, cyc = collections.Counter(), collections.Counter()
for r in rows[hi + 1:]:
    if len(r) < len(h) or not r[ci["Source"]]:
        continue
    ex = int(r[ci["Instructions Executed"]] or 0)
    k, c = cost(r[ci["Source"]])
    cnt[k] += ex
    cyc[k] += ex * c
tot_i, tot_c = sum(cnt.values()), sum(cyc.values())
print(f"measured: {cycles:.0f} cycles elapsed; {tot_i / nsmsp:.0f} warp instructions per sub-partition")
for k in sorted(cyc, key=lambda k: -cyc[k]):
    print(f"  {k:10s} {cnt[k] / nsmsp:10.0f} instr  x cost -> {cyc[k] / nsmsp:10.0f} cycles ({100 * cyc[k] / tot_c:4.1f} %)")
print(f"sum of issue costs per sub-partition: {tot_c / nsmsp:.0f} cycles = {100 * tot_c / nsmsp / cycles:.1f} % of the measured time")
