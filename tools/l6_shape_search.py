"""Compile k_line6 for every combination of the code-shape knobs (kernels_line6.cuh) and report registers, stack and
spill bytes: ptxas's allocation at the 168-register cap is chaotic, so the variant is picked by measurement.
   python tools/l6_shape_search.py [NP ...]"""
import itertools, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KNOBS = ["L6_Q_RELOAD", "L6_OLD_EARLY", "L6_HI_EARLY", "L6_CN_EARLY", "L6_ACC_VOUTER"]
nps = [int(a) for a in sys.argv[1:]] or [4, 2]
src = """#include <algorithm>
#include <cstdint>
#include <cuda_runtime.h>
#include "trixib200.h"
#include "kernels_line6.cuh"
void* p[] = {%s};
"""
with tempfile.TemporaryDirectory() as td:
    cu = os.path.join(td, "p.cu")
    open(cu, "w").write(src % ", ".join(f"(void*)tb::k_line6<5,5,false,4,3,{n}>" for n in nps))
    for combo in itertools.product((0, 1), repeat=len(KNOBS)):
        defs = [f"-D{k}={v}" for k, v in zip(KNOBS, combo)]
        r = subprocess.run(["nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-I" + ROOT + "/include",
                            "-I" + ROOT + "/trixicuda.jl_b200/csrc", "-Xptxas", "-v", "-c", cu, "-o", os.path.join(td, "p.o")] + defs,
                           capture_output=True, text=True)
        name = None
        for l in r.stderr.splitlines():
            m = re.search(r"Compiling entry function '(\S+)'", l)
            if m: name = m.group(1); continue
            m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", l)
            if m: st = m.groups(); continue
            m = re.search(r"Used (\d+) registers", l)
            t = re.search(r"k_line6ILi5ELi5ELb0ELi4ELi3ELi(\d)ELi0EEEv", name or "")
            if m and t:
                npv = t.group(1)
                print("".join(map(str, combo)), "NP", npv, "regs", m.group(1), "stack", st[0], "spill", st[1], st[2], flush=True)
                name = None
