"""Worker of tests/test_cpu_host.py::test_halo_exchange_world_size_2_gloo (launched by torchrun, gloo backend).

Emulates on the CPU what trixib200_rhs does between ranks: pack the face traces listed by the partition plan
(send_elem / send_dir, peer-major), exchange them per peer, and interpret the received buffer through the halo
codes of the local interface list. The expected traces come from the oracle's global interfaces.u."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), HERE]

import cases  # noqa: E402


def face_nodes(N, nd, dim, fixed):
    idx = np.arange(N ** nd).reshape((N,) * nd, order="F")
    sl = [slice(None)] * nd
    sl[dim] = fixed
    return idx[tuple(sl)].ravel(order="F")


def main():
    import trixib200 as T
    from trixib200 import distributed as D
    rank, _, world = D.init_process_group(backend="gloo")
    for case_name, level in (("c5_euler_ec_3d", 2), ("c2_euler_ec_2d", 3)):
        c = dict(cases.CASES[case_name], level=level)
        nd, N = c["ndim"], c["polydeg"] + 1
        o = cases.make_oracle(c)
        nv, nn, nf = o.nvars, N ** nd, N ** (nd - 1)
        u = o.compute_coefficients(0.0)
        du = o.new_u()
        o.stage("prolong2interfaces", du, u, 0.0)
        iu = o.f64("interfaces.u").reshape(-1, nf, nv, 2)          # [I, f, v, side]
        mesh = T.TreeMesh(c["cmin"], c["cmax"], initial_refinement_level=level, periodicity=True)
        basis = T.LobattoLegendreBasisGPU(c["polydeg"])
        plan, _ = D.partition_plan(mesh, basis.nodes, rank, world)
        first = plan.scalar("first_element")
        U = u.reshape(-1, nn, nv)
        se, sd = plan.array("send_elem"), plan.array("send_dir")
        send = np.empty((se.shape[0], nf, nv))
        for k in range(se.shape[0]):
            dim, side = sd[k] // 2, sd[k] % 2
            send[k] = U[first + se[k], face_nodes(N, nd, dim, N - 1 if side == 1 else 0)]
        recv = np.empty_like(send)
        peers, cnt = plan.array("peers"), plan.array("peer_count")
        reqs, off = [], 0
        st, rt = torch.from_numpy(send), torch.from_numpy(recv)
        for q, k in zip(peers.tolist(), cnt.tolist()):
            reqs.append(dist.isend(st[off: off + k].contiguous(), dst=q))
            reqs.append(dist.irecv(rt[off: off + k], src=q))
            off += k
        for r in reqs:
            r.wait()
        L, R, ifg = plan.array("if_left"), plan.array("if_right"), plan.array("if_global")
        nhalo = 0
        for s in range(ifg.shape[0]):
            if R[s] <= -2:
                assert np.array_equal(recv[-2 - R[s]], iu[ifg[s], :, :, 1]), (case_name, s)
                nhalo += 1
            if L[s] <= -2:
                assert np.array_equal(recv[-2 - L[s]], iu[ifg[s], :, :, 0]), (case_name, s)
                nhalo += 1
        assert nhalo == se.shape[0] and nhalo > 0
    dist.barrier()
    print("HALO_OK", rank, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
