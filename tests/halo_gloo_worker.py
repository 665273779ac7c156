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
    for case_name, level in (("c5_euler_ec_3d", 2), ("c2_euler_ec_2d", 3), ("euler_ec_mortar_off_3d", 2)):
        c = dict(cases.CASES[case_name], level=level)
        nd, N = c["ndim"], c["polydeg"] + 1
        o = cases.make_oracle(c)
        nv, nn, nf = o.nvars, N ** nd, N ** (nd - 1)
        u = o.compute_coefficients(0.0)
        du = o.new_u()
        o.stage("prolong2interfaces", du, u, 0.0)
        iu = o.f64("interfaces.u").reshape(-1, nf, nv, 2)          # [I, f, v, side]
        mesh = T.TreeMesh(c["cmin"], c["cmax"], initial_refinement_level=level, periodicity=True,
                          refinement_patches=c["patches"])
        basis = T.LobattoLegendreBasisGPU(c["polydeg"])
        plan, _ = D.partition_plan(mesh, basis.nodes, rank, world)
        first = plan.scalar("first_element")
        U = u.reshape(-1, nn, nv)
        se, sd = plan.array("send_elem"), plan.array("send_dir")
        send = np.empty((se.shape[0], nf, nv))
        for k in range(se.shape[0]):
            dim, side = sd[k] // 2, sd[k] % 2
            send[k] = U[first + se[k], face_nodes(N, nd, dim, N - 1 if side == 1 else 0)]
        peers, cnt, scnt = plan.array("peers"), plan.array("peer_count"), plan.array("peer_send_count")
        recv = np.empty((int(cnt.sum()), nf, nv))
        reqs, soff, roff = [], 0, 0
        st, rt = torch.from_numpy(send), torch.from_numpy(recv)
        for q, ks, kr in zip(peers.tolist(), scnt.tolist(), cnt.tolist()):   # as halo_begin posts them (runtime.cu)
            if ks:
                reqs.append(dist.isend(st[soff: soff + ks].contiguous(), dst=q))
            if kr:
                reqs.append(dist.irecv(rt[roff: roff + kr], src=q))
            soff, roff = soff + ks, roff + kr
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
        # mortars replicated across the cut: rows of elements on the other rank name the slot with that element's face
        ids = plan.array("mo_ids")
        nmortar_faces = 0
        if ids.size:
            rows = (4 if nd == 3 else 2) + 1
            ids = ids.reshape(-1, rows)
            mg, side, mdim = plan.array("mo_global"), plan.array("mo_side"), plan.array("mo_dim")
            gl = T.init_containers(mesh, basis.nodes).mortars.neighbor_ids - 1
            for ml in range(ids.shape[0]):
                for row in range(rows):
                    if ids[ml, row] <= -2:
                        large = row == rows - 1
                        fixed = (N - 1 if side[ml] == 1 else 0) if large else (0 if side[ml] == 1 else N - 1)
                        want = U[gl[row, mg[ml]], face_nodes(N, nd, mdim[ml], fixed)]
                        assert np.array_equal(recv[-2 - ids[ml, row]], want), (case_name, ml, row)
                        nmortar_faces += 1
            assert nmortar_faces > 0 or world == 1, case_name
        assert nhalo + nmortar_faces == recv.shape[0] and nhalo > 0
    dist.barrier()
    print("HALO_OK", rank, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
