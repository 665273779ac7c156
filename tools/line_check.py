"""Developer check of the line-owner kernel (whatever TRIXIB200_LINE_SHAPE selects): du against the CPU oracle at small
levels (odd element counts per warp included through level 2), then timing. Usage: line_check.py [check levels] -- [time levels]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]
import numpy as np
import torch
import cases

args = sys.argv[1:]
split = args.index("--") if "--" in args else len(args)
check = [int(a) for a in args[:split]] or ([] if "--" in args else [2, 3, 4])
timed = [int(a) for a in args[split + 1:]]
shape = os.environ.get("TRIXIB200_LINE_SHAPE", "default")
ok = True
for lv in check:
    c = dict(cases.CASES["c5_euler_ec_3d"], level=lv)
    o = cases.make_oracle(c)
    u = o.compute_coefficients(0.0)
    rng = np.random.default_rng(lv)
    for kind in ("ic", "rough"):
        if kind == "rough":      # both ln_mean branches, no symmetry
            U = u.reshape(-1, 5).copy()
            U[:, 0] *= rng.uniform(0.7, 1.4, len(U)); U[:, 1:4] += rng.uniform(-0.2, 0.2, (len(U), 3))
            U[:, 4] += rng.uniform(0.0, 1.0, len(U))
            u = U.ravel()
        ref = o.rhs(u, 0.0)
        semi = cases.make_semi(c, node_coordinates=False)
        assert semi.line3d
        ud = torch.from_numpy(u).cuda()
        du = semi.new_vector()
        errs = []
        for rep in range(3):
            du.fill_(float("nan"))
            semi.rhs(du, ud, 0.0)
            torch.cuda.synchronize()
            errs.append(cases.rel_max_err(du.cpu().numpy(), ref))
        good = max(errs) <= 1e-12
        ok &= good
        print(f"shape={shape} level {lv} {kind}: rel max err {max(errs):.2e} {'ok' if good else 'FAIL'}", flush=True)
        del semi
for lv in timed:
    c = dict(cases.CASES["c5_euler_ec_3d"], level=lv)
    semi = cases.make_semi(c, node_coordinates=False)
    u = semi.compute_coefficients_gpu(0.0, on_device=True)
    du = semi.new_vector()
    for _ in range(3):
        semi.rhs(du, u, 0.0)
    torch.cuda.synchronize()
    best = min(semi.time_rhs(du, u, 0.0, 10) / 10 for _ in range(3))
    nd = semi.ndofs()
    print(f"shape={shape} level {lv}: rhs {best:.4f} ms  {nd / best / 1e6:.2f} GDOF/s  finite={bool(torch.isfinite(du).all())}",
          flush=True)
    del semi, u, du
    torch.cuda.empty_cache()
sys.exit(0 if ok else 1)
