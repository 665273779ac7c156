"""Multi-GPU plumbing: one process per GPU (torchrun), torch.distributed only carries the rendezvous.

The reference is single-GPU (reference src/auxiliary/auxiliary.jl:7-9); the partition along the TreeMesh Morton
curve and the halo-face exchange (ncclSend/ncclRecv inside trixib200_rhs) are new. Each rank passes the WHOLE
host mesh to trixib200_create together with (rank, nranks); the library keeps its contiguous range of the leaf
order. This module only (1) initialises the process group, (2) broadcasts the 128-byte NCCL unique id that
trixib200_comm_init needs, (3) exposes the host-only partition plan (trixib200_plan_*) for CPU tests.
"""
import ctypes as C
import os

import numpy as np

from . import _lib


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def init_process_group(backend=None):
    """torch.distributed rendezvous from the torchrun environment (127.0.0.1 default master)."""
    import torch
    import torch.distributed as dist
    rank, local_rank, world = env_rank()
    if world == 1:
        return rank, local_rank, world
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29531")
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(local_rank)
    if not dist.is_initialized():
        kw = {}
        if backend == "nccl":
            kw["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, local_rank, world


def broadcast_comm_id():
    """Rank 0 creates the NCCL unique id of the library's own communicator; everyone receives it."""
    import torch
    import torch.distributed as dist
    buf = C.create_string_buffer(128)
    if dist.get_rank() == 0:
        _lib.check(_lib.lib().trixib200_comm_unique_id(buf))
    t = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.broadcast(t, src=0)
    return bytes(t.cpu().numpy().tobytes())


class PartitionPlan:
    """Host-only view of the plan trixib200_create builds for one rank (no CUDA needed)."""

    def __init__(self, cfg, mesh_host, keep=None):
        self._L = _lib.lib()
        self._p = C.c_void_p()
        self._keep = keep
        _lib.check(self._L.trixib200_plan_create(C.byref(cfg), C.byref(mesh_host), C.byref(self._p)))

    def __del__(self):
        if getattr(self, "_p", None):
            self._L.trixib200_plan_destroy(self._p)
            self._p = None

    def scalar(self, name):
        return int(self._L.trixib200_plan_len(self._p, name.encode()))

    def array(self, name):
        n = self._L.trixib200_plan_len(self._p, name.encode())
        if n < 0:
            raise KeyError(name)
        out = np.empty(n, dtype=np.int64)
        if n:
            _lib.check(self._L.trixib200_plan_get(self._p, name.encode(), _lib.fptr(out), n))
        return out


def mesh_host_struct(containers, mesh):
    """trixib200_mesh_host from numpy containers (connectivity only: enough for the partition plan)."""
    keep = []

    def i64(a, fortran=False):
        a = np.ascontiguousarray(np.asarray(a, dtype=np.int64).ravel(order="F" if fortran else "C"))
        keep.append(a)
        return _lib.fptr(a)

    def f64(a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        keep.append(a)
        return _lib.fptr(a)

    c = containers
    mh = _lib.MeshHost()
    mh.nelements = c.elements.inverse_jacobian.shape[0]
    mh.ninterfaces = c.interfaces.orientations.shape[0]
    mh.nboundaries = c.boundaries.neighbor_ids.shape[0]
    mh.nmortars = c.mortars.orientations.shape[0]
    mh.inverse_jacobian = f64(c.elements.inverse_jacobian)
    mh.cell_centers = f64(mesh.cell_centers()[:, : mesh.ndim].ravel())
    mh.interfaces_neighbor_ids = i64(c.interfaces.neighbor_ids, True)
    mh.interfaces_orientations = i64(c.interfaces.orientations)
    mh.boundaries_neighbor_ids = i64(c.boundaries.neighbor_ids)
    mh.boundaries_orientations = i64(c.boundaries.orientations)
    mh.boundaries_neighbor_sides = i64(c.boundaries.neighbor_sides)
    mh.n_boundaries_per_direction = i64(c.boundaries.n_boundaries_per_direction)
    mh.mortars_neighbor_ids = i64(c.mortars.neighbor_ids, True)
    mh.mortars_large_sides = i64(c.mortars.large_sides)
    mh.mortars_orientations = i64(c.mortars.orientations)
    return mh, keep


def partition_plan(mesh, nodes, rank, nranks, ndim=None, bc_periodic=True):
    """Plan of `rank` for a TreeMesh (host only)."""
    from .treemesh import init_containers
    c = init_containers(mesh, nodes)
    mh, keep = mesh_host_struct(c, mesh)
    cfg = _lib.Config()
    cfg.ndim, cfg.rank, cfg.nranks = mesh.ndim, rank, nranks
    for i in range(6):
        cfg.boundary_conditions[i] = _lib.BC_PERIODIC if bc_periodic else _lib.BC_DIRICHLET_IC
    return PartitionPlan(cfg, mh, keep), c
