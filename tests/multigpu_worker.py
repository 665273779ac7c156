"""Worker of test_multi_gpu_rhs_matches_oracle (torchrun, one process per GPU, NCCL): every rank owns one
contiguous range of the Morton order, computes rhs! and max_dt on it and compares with the oracle's global
result; both the fused and the staged kernels, plus the C-ABI host-vector entry point."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), HERE]

import cases  # noqa: E402


def main():
    from trixib200 import distributed as D
    rank, local_rank, world = D.init_process_group(backend="nccl")
    lv = int(os.environ.get("TRIXIB200_MULTI_LEVEL", "3"))     # level of the line-kernel case (4-5 for the 8-rank log)
    for name, level in (("c5_euler_ec_3d", lv), ("c2_euler_ec_2d", 4), ("euler_source_terms_3d", 2),
                        ("euler_nonperiodic_3d", 2), ("advection_basic_3d", 3), ("mhd_ec_3d", 2),
                        ("c3_euler_sc_3d_nosmooth", 3),
                        # alpha smoothing across cuts, mortars replicated across cuts
                        ("c3_euler_sc_3d", 3), ("c4_mhd_alfven_mortar_3d", 2), ("euler_ec_mortar_off_3d", 2),
                        ("mhd_alfven_mortar_off_3d", 2), ("euler_shock_mortar_off_3d", 2), ("euler_ec_mortar_2d", 3),
                        ("euler_shock_2d", 4)):
        if name == "c3_euler_sc_3d_nosmooth":
            c = dict(cases.CASES["c3_euler_sc_3d"], alpha_smooth=False, level=level)
        else:
            c = dict(cases.CASES[name], level=level)
        o = cases.make_oracle(c)
        u = o.compute_coefficients(0.0)
        du_ref = o.rhs(u, 0.1)
        dt_ref = o.max_dt(u)
        for staged in (False, True):
            comm_id = D.broadcast_comm_id()
            semi = cases.make_semi(c, staged_only=staged, device=local_rank, rank=rank, nranks=world, comm_id=comm_id)
            if rank == 0:
                print(f"[multigpu] {name} staged={staged} in-kernel peer-memory halo exchange: "
                      f"{bool(semi.size('p2p_halo'))}", flush=True)
            lo = semi.local_slice(u)
            u_d = torch.from_numpy(np.ascontiguousarray(lo)).to(semi.device)
            du_d = semi.new_vector().fill_(float("nan"))
            for _ in range(3):          # repeated calls reuse the halo buffers
                semi.rhs(du_d, u_d, 0.1)
            torch.cuda.synchronize()
            got = du_d.cpu().numpy()
            ref = semi.local_slice(du_ref)
            err = np.abs(got - ref).max() / np.abs(du_ref).max()
            assert err <= 1e-12, (name, staged, rank, err)
            dt = semi.max_dt(u_d, 0.0)
            assert abs(dt - dt_ref) <= 1e-13 * dt_ref, (name, dt, dt_ref)
            if not staged:
                # device-side analysis reductions are reduced over the ranks (trixib200_calc_error_norms / integrate)
                import trixib200 as T
                l2, linf = semi.calc_error_norms(u_d, 0.4, T.SolutionAnalyzer(semi.solver.basis))
                l2_ref, linf_ref = o.error_norms(u, 0.4)
                assert np.abs(l2 - l2_ref).max() <= 1e-13 * max(1.0, np.abs(l2_ref).max()), (name, l2, l2_ref)
                assert np.abs(linf - linf_ref).max() <= 1e-12 * max(1.0, np.abs(linf_ref).max()), (name, linf, linf_ref)
                integ, integ_ref = semi.integrate(u_d, normalize=False), o.integrate(u)
                # (sums of ~1e5 terms in a different order on every rank count: round-off of the summation)
                assert np.abs(integ - integ_ref).max() <= 2e-12 * max(1.0, np.abs(integ_ref).max()), (name, integ - integ_ref)
                # fused Runge-Kutta stage on the partition: same two element lists, halo exchange inside
                tmp = semi.new_vector().zero_()
                u_out = semi.new_vector().fill_(float("nan"))
                semi.rk2n_stage(u_out, u_d, tmp, 0.1, 0.0, 0.25, 1e-3)
                torch.cuda.synchronize()
                ref_out = semi.local_slice(u + 0.25 * 1e-3 * du_ref)
                assert np.abs(u_out.cpu().numpy() - ref_out).max() <= 1e-12 * 1e-3 * np.abs(du_ref).max() + 4e-16 * np.abs(ref_out).max(), (name, "rk2n_stage")
            du_h = np.full_like(lo, np.nan)
            semi.rhs_host(du_h, np.ascontiguousarray(lo), 0.1)
            assert np.array_equal(du_h, got), (name, staged, "rhs_host")
            del semi
        dist.barrier()
        if rank == 0:
            print(f"[multigpu] {name} done", flush=True)      # (a hang is then attributable to one case)
    print("MULTIGPU_OK", rank, flush=True)
    dist.barrier()
    sys.stdout.flush()
    os._exit(0)      # like bench.py: leave without communicator / IPC teardown in garbage-collection order


if __name__ == "__main__":
    main()
