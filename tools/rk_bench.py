"""Developer timing: one CarpenterKennedy2N54 stage as rhs! + update kernel vs the fused trixib200_rk2n_stage."""
import sys, os, json
sys.path[:0] = [os.path.dirname(os.path.dirname(os.path.abspath(__file__)))]
sys.path[:0] = [os.path.join(sys.path[0], "tests"), os.path.join(sys.path[0], "oracle")]
import torch
import cases

for lv in [int(a) for a in sys.argv[1:]] or [6]:
    c = dict(cases.CASES["c5_euler_ec_3d"], level=lv)
    semi = cases.make_semi(c, node_coordinates=False)
    u = semi.compute_coefficients_gpu(0.0, on_device=True)
    u2, du, tmp = semi.new_vector(), semi.new_vector(), semi.new_vector().zero_()
    a, b, dt = -0.4178904745, 0.3792103129999, 1e-4
    def timed(fn, reps=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(torch.cuda.current_stream())
        for _ in range(reps):
            fn()
        e1.record(torch.cuda.current_stream())
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    def unfused():
        semi.rhs(du, u, 0.0)
        semi.rk2n_update(u2, tmp, du, a, b, dt)
    def fused():
        semi.rk2n_stage(u2, u, tmp, 0.0, a, b, dt)
    ms_rhs = timed(lambda: semi.rhs(du, u, 0.0))
    ms_un = timed(unfused)
    ms_fu = timed(fused)
    nd = semi.ndofs()
    print(json.dumps({"level": lv, "ndofs": nd, "rhs_ms": ms_rhs, "rhs_plus_update_ms": ms_un, "fused_stage_ms": ms_fu,
                      "stage_speedup": ms_un / ms_fu, "fused_dof_stages_per_s": nd / ms_fu * 1e3}), flush=True)
    del semi, u, u2, du, tmp
    torch.cuda.empty_cache()
