"""Static estimate of the instructions one iteration of k_line6's persistent loop executes (no GPU needed): instructions
between the loop head and its back edge, minus those whose source line belongs to a cold construct (the logarithmic-branch
correction, source terms, the multi-GPU pack / flag wait) -- checked against ncu's executed-instruction counts of the default
shape (profiles/r2_ncu_line6_l6_mix.txt: 2461 warp instructions per element pair).
    python tools/sass_hot.py <mangled-name-substring> [--unrolled | --peeled]
(--unrolled: the kernel has no phase loop; --peeled: the loop runs the y and x phases only)

The single-copy shape is over-counted (its run-time switches between the three directions are counted with all arms),
the unrolled shape is not."""
import collections, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "trixicuda.jl_b200", "libtrixib200.so")
pat = sys.argv[1]
def lines_of(name):
    return open(os.path.join(ROOT, "trixicuda.jl_b200", "csrc", name)).read().split("\n")
def find(src, sub, start=0):
    for i in range(start, len(src)):
        if sub in src[i]:
            return i + 1
    raise SystemExit("marker not found: " + sub)
# cold source ranges (1-based, inclusive): the logarithmic-branch correction and the source terms in the phase body,
# the multi-GPU pack / flag wait in the kernel
inc, cuh = lines_of("kernels_line6_phase.inc"), lines_of("kernels_line6.cuh")
c0 = find(inc, "if (FAST && worst >= L6_ROUGH_HI)")
c1 = find(inc, "if (PP && K0 + NP > 7)", c0) - 1
s0 = find(inc, "if (d.src != TRIXIB200_SRC_NONE)")
COLD_INC = [(c0, c1), (s0, s0 + 5)]
p0 = find(cuh, "if (p2p.npeers != 0) {")
p1 = find(cuh, "static_assert(!PP ||", p0) - 1
COLD6 = [(p0, p1)]
COLD_FILES = {"equations.cuh", "device.cuh", "kernels_warp3d.cuh"}          # exact means, face_node of the pack loop
def cold(f, n):
    if f in COLD_FILES:
        return True
    if f == "kernels_line3d.cuh" and n >= 100:       # l3_means_exact, l3_pair_correction, l3_source
        return True
    if f == "kernels_line6_phase.inc":
        return any(a <= n <= b for a, b in COLD_INC)
    if f == "kernels_line6.cuh":
        return any(a <= n <= b for a, b in COLD6) or n <= 40 or (140 <= n <= 150)
    return False
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.split("\n")
ins, on, line, labels, pend = [], False, None, {}, []
for l in txt:
    if l.startswith(".text."):
        on = pat in l
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        line = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r'^(\.L_x_\d+):', l)
    if m:
        pend.append(m.group(1))
        continue
    m = re.match(r'\s+/\*([0-9a-f]+)\*/\s+(.*?);', l)
    if m:
        for lb in pend:
            labels[lb] = len(ins)
        pend = []
        ins.append((int(m.group(1), 16), m.group(2).strip(), line))
# back edges: branch to a label at a lower index
back = []
for i, (a, s, ln) in enumerate(ins):
    m = re.search(r'BRA\S*\s+.*?`\((\.L_x_\d+)\)', s)
    if m and m.group(1) in labels and labels[m.group(1)] <= i:
        back.append((labels[m.group(1)], i))
back.sort(key=lambda t: t[0] - t[1])
outer = back[0]
if '--unrolled' in sys.argv:
    back = back[:1]
inner = sorted([b for b in back[1:] if outer[0] < b[0] and b[1] < outer[1] and b[1] - b[0] > 300], key=lambda t: t[1] - t[0])[:1]
def count(lo, hi, skip=()):
    c, ops = 0, collections.Counter()
    for i in range(lo, hi + 1):
        if any(a <= i <= b for a, b in skip):
            continue
        a, s, ln = ins[i]
        if ln and cold(*ln):
            continue
        op = re.sub(r'^@!?U?P\d+\s+', '', s).split()[0].split('.')[0]
        c += 1; ops[op] += 1
    return c, ops
print("kernel instructions", len(ins), "outer loop", outer, "inner loops", inner)
tot, ops = count(outer[0], outer[1], skip=inner[:1])
if inner:
    ib, iops = count(*inner[0])
    reps = 2 if "--peeled" in sys.argv else 3
    print("per-pair part", tot, "phase body", ib, "x", reps, "-> per element pair", tot + reps * ib)
    for k in iops: ops[k] += reps * iops[k]
else:
    print("per element pair (phases unrolled)", tot)
fp = sum(ops[k] for k in ("DFMA", "DMUL", "DADD"))
print("FP64", fp, "other", sum(ops.values()) - fp)
print(" ".join(f"{k}={v}" for k, v in ops.most_common(30)))
