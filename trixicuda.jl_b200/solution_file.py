"""Trixi's on-disk format for the solution and the mesh: `SaveSolutionCallback`, `save_solution_file`, `save_mesh_file`.

The reference example writes solution files through Trixi's callback (reference examples/euler_ec_3d.jl:38-41:
`SaveSolutionCallback(interval = 100, save_initial_solution = true, save_final_solution = true, solution_variables =
cons2prim)`), which copies `u` to the host and calls Trixi's `save_solution_file` / `save_mesh_file`. This module writes
the same files from the Python mirror (layout restated from Trixi.jl's callbacks_step/save_solution_dg.jl and
meshes/mesh_io.jl):

solution_%09d.h5 -- root attributes ndims, equations, polydeg, n_vars, n_elements, mesh_type, mesh_file, time, dt, timestep;
    datasets variables_1 .. variables_n = vec(data[v, .., :]) (node index i fastest, element slowest) with attribute
    "name"; element_variables_1 .. with attribute "name" (e.g. the shock-capturing alpha).
mesh.h5 -- root attributes mesh_type, ndims, n_cells, capacity, n_leaf_cells, minimum_level, maximum_level, center_level_0,
    length_level_0, periodicity; datasets parent_ids, child_ids [2^d, n_cells], neighbor_ids [2 d, n_cells], levels,
    coordinates [d, n_cells] of the WHOLE tree (parents included) in Trixi's depth-first order, ids 1-based, 0 = none.

The HDF5 container is written by hdf5_lite (no h5py / libhdf5 in this image). Multi-rank: rank 0 writes after a gather
(`semi.gather_to_host`) -- the files describe the global mesh, as Trixi's do.
"""
import os

import numpy as np

from . import hdf5_lite, _lib

_KIND = {_lib.EQ_ADVECTION: "advection", _lib.EQ_EULER: "euler", _lib.EQ_MHD: "mhd"}

_NAMES = {
    ("advection", "cons"): ("scalar",), ("advection", "prim"): ("scalar",),
    ("euler", 1, "cons"): ("rho", "rho_v1", "rho_e"), ("euler", 1, "prim"): ("rho", "v1", "p"),
    ("euler", 2, "cons"): ("rho", "rho_v1", "rho_v2", "rho_e"), ("euler", 2, "prim"): ("rho", "v1", "v2", "p"),
    ("euler", 3, "cons"): ("rho", "rho_v1", "rho_v2", "rho_v3", "rho_e"),
    ("euler", 3, "prim"): ("rho", "v1", "v2", "v3", "p"),
    ("mhd", 3, "cons"): ("rho", "rho_v1", "rho_v2", "rho_v3", "rho_e", "B1", "B2", "B3", "psi"),
    ("mhd", 3, "prim"): ("rho", "v1", "v2", "v3", "p", "B1", "B2", "B3", "psi"),
}


def cons2prim(u, equations):
    """Nodewise conversion on an array whose LAST axis is the variable (Trixi cons2prim)."""
    kind = _KIND[equations.kind]
    if kind == "advection":
        return u
    nd = equations.ndim
    out = np.array(u, dtype=np.float64, copy=True)
    rho = u[..., 0]
    v = u[..., 1:1 + nd] / rho[..., None]
    out[..., 1:1 + nd] = v
    kin = 0.5 * rho * (v * v).sum(-1)
    if kind == "euler":
        out[..., nd + 1] = (equations.gamma - 1) * (u[..., nd + 1] - kin)
    else:   # GLM-MHD: p = (gamma - 1) (rho_e - kin - |B|^2 / 2 - psi^2 / 2)
        mag = 0.5 * (u[..., 5:8] ** 2).sum(-1)
        out[..., 4] = (equations.gamma - 1) * (u[..., 4] - kin - mag - 0.5 * u[..., 8] ** 2)
    return out


def cons2cons(u, equations):
    return u


def varnames(solution_variables, equations):
    which = "prim" if solution_variables is cons2prim else "cons"
    kind = _KIND[equations.kind]
    if kind == "advection":
        return _NAMES[("advection", which)]
    return _NAMES[(kind, equations.ndim, which)]


def solution_filename(output_directory, timestep):
    return os.path.join(output_directory, "solution_%09d.h5" % timestep)


def save_solution_file(u_ode, time, dt, timestep, semi, solution_variables=cons2prim, output_directory="out",
                       element_variables=None, mesh_file="mesh.h5"):
    """Trixi `save_solution_file(u, time, dt, timestep, mesh, equations, dg, cache, solution_callback, ...)`."""
    u = u_ode.detach().cpu().numpy() if hasattr(u_ode, "detach") else np.asarray(u_ode)
    mesh, eq = semi.mesh, semi.equations
    nd, n, nv = mesh.ndim, semi.nnodes, semi.nvars
    E = u.size // (nv * n ** nd)
    data = solution_variables(u.reshape((E,) + (n,) * nd + (nv,)), eq)
    f = hdf5_lite.File()
    f.attrs["ndims"] = nd
    f.attrs["equations"] = type(eq).__name__
    f.attrs["polydeg"] = n - 1
    f.attrs["n_vars"] = nv
    f.attrs["n_elements"] = E
    f.attrs["mesh_type"] = "TreeMesh"
    f.attrs["mesh_file"] = mesh_file
    f.attrs["time"] = float(time)
    f.attrs["dt"] = float(dt)
    f.attrs["timestep"] = int(timestep)
    names = varnames(solution_variables, eq)
    for v in range(nv):
        d = f.create_dataset(f"variables_{v + 1}", np.ascontiguousarray(data[..., v]).ravel())
        d.attrs["name"] = names[v]
    for v, (key, arr) in enumerate((element_variables or {}).items()):
        d = f.create_dataset(f"element_variables_{v + 1}", np.asarray(arr, dtype=np.float64).ravel())
        d.attrs["name"] = str(key)
    os.makedirs(output_directory, exist_ok=True)
    path = solution_filename(output_directory, timestep)
    f.write(path)
    return path


def load_solution_file(path):
    """-> (attributes, data [n_elements, n, (n, (n,)) n_vars], variable names, element variables)."""
    attrs, ds = hdf5_lite.File.read(path)
    nd, n, nv, E = attrs["ndims"], attrs["polydeg"] + 1, attrs["n_vars"], attrs["n_elements"]
    data = np.stack([ds[f"variables_{v + 1}"][0].reshape((E,) + (n,) * nd) for v in range(nv)], axis=-1)
    names = [ds[f"variables_{v + 1}"][1]["name"] for v in range(nv)]
    elem = {a["name"]: arr for k, (arr, a) in ds.items() if k.startswith("element_variables_")}
    return attrs, data, names, elem


# ------------------------------------------------------------------------------------------------- mesh file
def tree_arrays(mesh):
    """The whole 2^d-tree behind the leaves of `mesh` (parents included) in Trixi's storage order (depth first, a
    parent directly before its children, children in Morton order with x fastest): parent_ids, child_ids,
    neighbor_ids (same-level neighbour or 0), levels, coordinates -- ids 1-based like Trixi's."""
    nd = mesh.ndim
    lv_leaf, ic_leaf = mesh.levels, mesh.icoords[:, :nd]
    lmax = int(lv_leaf.max()) if lv_leaf.size else 0
    cells = set()
    for l, ic in zip(lv_leaf.tolist(), ic_leaf.tolist()):
        while (l, tuple(ic)) not in cells:
            cells.add((l, tuple(ic)))
            if l == 0:
                break
            l, ic = l - 1, [c >> 1 for c in ic]

    def key(l, ic):
        k = 0
        for b in range(l):
            for d in range(nd):
                k |= ((ic[d] >> b) & 1) << (b * nd + d)
        return k << (nd * (lmax - l))
    order = sorted(cells, key=lambda c: (key(*c), c[0]))
    ids = {c: i + 1 for i, c in enumerate(order)}
    n = len(order)
    parent = np.zeros(n, dtype=np.int64)
    child = np.zeros((n, 1 << nd), dtype=np.int64)
    nbr = np.zeros((n, 2 * nd), dtype=np.int64)
    levels = np.zeros(n, dtype=np.int64)
    coords = np.zeros((n, nd))
    for i, (l, ic) in enumerate(order):
        levels[i] = l
        if l > 0:
            p = ids[(l - 1, tuple(c >> 1 for c in ic))]
            parent[i] = p
            child[p - 1, sum((ic[d] & 1) << d for d in range(nd))] = i + 1
        m = 1 << l
        for d in range(nd):
            for s, side in ((-1, 0), (1, 1)):
                q = list(ic)
                q[d] += s
                if not 0 <= q[d] < m:
                    if not mesh.periodicity[d]:
                        continue
                    q[d] %= m
                nbr[i, 2 * d + side] = ids.get((l, tuple(q)), 0)
        # cell centre accumulated root to leaf with the float operations of Trixi's refine!
        x = np.array(mesh.center_level_0[:nd], dtype=np.float64)
        for lev in range(1, l + 1):
            dx = mesh.length_level_0 / float(1 << lev)
            for d in range(nd):
                x[d] += (1.0 if (ic[d] >> (l - lev)) & 1 else -1.0) * dx / 2
        coords[i] = x
    return dict(parent_ids=parent, child_ids=child, neighbor_ids=nbr, levels=levels, coordinates=coords)


def save_mesh_file(mesh, output_directory="out", filename="mesh.h5"):
    """Trixi `save_mesh_file(mesh::TreeMesh, output_directory)` (serial)."""
    t = tree_arrays(mesh)
    f = hdf5_lite.File()
    f.attrs["mesh_type"] = "TreeMesh"
    f.attrs["ndims"] = mesh.ndim
    f.attrs["n_cells"] = int(t["levels"].size)
    f.attrs["capacity"] = int(mesh.n_cells_max if mesh.n_cells_max is not None else t["levels"].size)
    f.attrs["n_leaf_cells"] = int(mesh.n_leaf_cells)
    f.attrs["minimum_level"] = int(mesh.levels.min())
    f.attrs["maximum_level"] = int(mesh.levels.max())
    f.attrs["center_level_0"] = np.asarray(mesh.center_level_0[:mesh.ndim], dtype=np.float64)
    f.attrs["length_level_0"] = float(mesh.length_level_0)
    f.attrs["periodicity"] = np.asarray(mesh.periodicity, dtype=bool)
    for k in ("parent_ids", "child_ids", "neighbor_ids", "levels", "coordinates"):
        f.create_dataset(k, t[k])
    os.makedirs(output_directory, exist_ok=True)
    path = os.path.join(output_directory, filename)
    f.write(path)
    return path


def gather_to_host(u, semi):
    """The global solution vector on the host of rank 0 (None elsewhere): the ranks own contiguous ranges of the
    Morton order, so the global vector is the concatenation of the local ones."""
    loc = u.detach().cpu().numpy() if hasattr(u, "detach") else np.asarray(u)
    if getattr(semi, "nranks", 1) == 1:
        return loc
    import torch.distributed as dist
    parts = [None] * semi.nranks if semi.rank == 0 else None
    dist.gather_object(loc, parts, dst=0)
    return np.concatenate(parts) if semi.rank == 0 else None


class SaveSolutionCallback:
    """`SaveSolutionCallback(interval = ..., save_initial_solution = true, save_final_solution = true,
    solution_variables = cons2prim, output_directory = "out")` as in the reference example; called by `solve` with
    (u, t, dt, timestep, finished)."""

    def __init__(self, interval=0, save_initial_solution=True, save_final_solution=True, solution_variables=cons2prim,
                 output_directory="out"):
        self.interval, self.save_initial_solution = int(interval), bool(save_initial_solution)
        self.save_final_solution, self.solution_variables = bool(save_final_solution), solution_variables
        self.output_directory = output_directory
        self.files = []
        self._mesh_saved = False

    def __call__(self, u, t, dt, timestep, semi, finished=False):
        due = (timestep == 0 and self.save_initial_solution) or (finished and self.save_final_solution) or \
              (self.interval > 0 and timestep > 0 and timestep % self.interval == 0)
        if not due or (self.files and self.files[-1][0] == timestep):
            return None
        u_host = gather_to_host(u, semi)
        if u_host is None:              # not rank 0
            return None
        if not self._mesh_saved:
            save_mesh_file(semi.mesh, self.output_directory)
            self._mesh_saved = True
        elem = {}
        if getattr(semi.solver.volume_integral, "indicator", None) is not None and getattr(semi, "nranks", 1) == 1:
            try:
                elem["indicator_shock_capturing"] = semi.cache("alpha")
            except Exception:
                pass
        path = save_solution_file(u_host, t, dt, timestep, semi, self.solution_variables, self.output_directory, elem)
        self.files.append((timestep, path))
        return path
