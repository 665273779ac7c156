"""Static SASS instruction counts per source line of one kernel (nvdisasm -g -c on the cubin of libtrixib200.so):
which source constructs the instructions of a kernel come from, without running it.
    python tools/sass_lines.py <mangled-name-substring> [lo-hi source line range to print]"""
import collections, os, re, subprocess, sys, tempfile
so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "trixicuda.jl_b200", "libtrixib200.so")
pat = sys.argv[1]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.split("\n")
on, line, per, ops = False, None, collections.Counter(), collections.defaultdict(collections.Counter)
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
    m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(.*?);', l)
    if m and line:
        s = re.sub(r'^@!?U?P\d+\s+', '', m.group(1).strip())
        op = s.split()[0].split('.')[0]
        per[line] += 1
        ops[line][op] += 1
print("total", sum(per.values()))
for (f, n), c in sorted(per.items()):
    print(f"{f}:{n:4d} {c:5d}  " + " ".join(f"{k}={v}" for k, v in ops[(f, n)].most_common(6)))
