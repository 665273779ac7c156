"""Dynamic instruction mix and stall attribution per opcode from an .ncu-rep captured with --import-source on."""
import collections, csv, re, subprocess, sys
rep, ndofs = sys.argv[1], float(sys.argv[2])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, data = rows[1], rows[2:]
ix = {k: i for i, k in enumerate(h)}
def f(r, k):
    try: return float(r[ix[k]].replace(',', ''))
    except Exception: return 0.0
cnt, samp = collections.Counter(), collections.Counter()
st = collections.defaultdict(collections.Counter)
for r in data:
    m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[ix['Source']])
    op = m.group(2).split('.')[0] if m else '?'
    cnt[op] += f(r, 'Instructions Executed'); samp[op] += f(r, '# Samples')
    for k in ('stall_wait', 'stall_short_sb', 'stall_long_sb', 'stall_no_inst', 'stall_math', 'stall_mio', 'stall_branch_resolving', 'stall_dispatch'):
        st[op][k[6:]] += f(r, k)
N = ndofs / 32
tot = sum(samp.values())
print("per DOF:", {o: round(c / N, 1) for o, c in cnt.most_common(28)})
print("fp64/DOF %.1f  total/DOF %.1f" % (sum(cnt[o] for o in ('DFMA', 'DMUL', 'DADD', 'DSETP', 'DMNMX')) / N, sum(cnt.values()) / N))
for op, s in samp.most_common(12):
    print("%-8s samp %5.1f%%  %s" % (op, 100 * s / tot, " ".join(f"{k}={100*v/tot:.1f}" for k, v in st[op].most_common(3))))
