"""Developer timing loop (not the contract bench): rhs! time for 3D Euler EC p=3 at a few levels."""
import sys, os, time
sys.path[:0] = [os.path.dirname(os.path.dirname(os.path.abspath(__file__)))] 
sys.path[:0] = [os.path.join(sys.path[0], "tests"), os.path.join(sys.path[0], "oracle")]
import torch, numpy as np
import cases

levels = [int(a) for a in sys.argv[1:]] or [4, 5, 6]
kern = os.environ.get("KERN", "auto")
for staged in (False,):
    for lv in levels:
        if staged and lv > 6:
            continue
        c = dict(cases.CASES["c5_euler_ec_3d"], level=lv)
        t0 = time.time()
        semi = cases.make_semi(c, staged_only=(kern == "staged"), no_warp_kernel=(kern == "node"),
                               no_line_kernel=(kern == "warp"), node_coordinates=False)
        t1 = time.time()
        u = semi.compute_coefficients_gpu(0.0, on_device=True)
        du = semi.new_vector()
        for _ in range(3):
            semi.rhs(du, u, 0.0)
        torch.cuda.synchronize()
        reps = 20
        ms = semi.time_rhs(du, u, 0.0, reps) / reps
        nd = semi.ndofs()
        print(f"level {lv} kern={kern} line3d={semi.line3d} E={semi.nelements} setup {t1-t0:.1f}s rhs {ms:.4f} ms  "
              f"{nd/ms/1e6:.2f} GDOF/s  algo-HBM {81*nd/ms/1e6:.0f} GB/s", flush=True)
        del semi, u, du
        torch.cuda.empty_cache()
