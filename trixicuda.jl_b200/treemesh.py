"""Host-side TreeMesh and DG containers (numpy), the stand-in for what the Julia shim obtains from Trixi.jl.

In the drop-in, the Julia shim runs Trixi's own CPU ``init_elements / init_interfaces / init_boundaries /
init_mortars`` exactly as the reference does (/root/reference/src/solvers/cache.jl:130-158) and hands the host
arrays to ``trixib200_create``. Julia is not available in this build environment, so this module produces
the same arrays (same ordering, 1-based Int64 ids) with a vectorised algorithm on Morton keys instead of
Trixi's pointer tree: leaves in depth-first order == ascending normalised Morton key with x as the fastest
bit. Field layouts follow /root/reference/src/solvers/containers_3d.jl:6-28,57-75,99-128,160-191.
"""
from dataclasses import dataclass, field

import numpy as np


def _interleave(ic, ndim, nbits):
    """Morton key with dimension 0 as the fastest bit."""
    key = np.zeros(ic.shape[0], dtype=np.int64)
    for b in range(nbits):
        for d in range(ndim):
            key |= ((ic[:, d] >> b) & 1) << (b * ndim + d)
    return key


def _deinterleave(key, ndim, nbits):
    ic = np.zeros((key.shape[0], 3), dtype=np.int64)
    for b in range(nbits):
        for d in range(ndim):
            ic[:, d] |= ((key >> (b * ndim + d)) & 1) << b
    return ic


class TreeMesh:
    """``TreeMesh(coordinates_min, coordinates_max; initial_refinement_level, refinement_patches, periodicity,
    n_cells_max)`` -- hypercube 2^d-tree mesh (Trixi semantics)."""

    MAX_LEVEL = 20

    def __init__(self, coordinates_min, coordinates_max, initial_refinement_level=0, refinement_patches=(),
                 periodicity=True, n_cells_max=None):
        cmin = np.atleast_1d(np.asarray(coordinates_min, dtype=np.float64))
        cmax = np.atleast_1d(np.asarray(coordinates_max, dtype=np.float64))
        self.ndim = int(cmin.shape[0])
        if self.ndim not in (1, 2, 3):
            raise ValueError("TreeMesh: ndim must be 1, 2 or 3")
        if np.isscalar(periodicity) or isinstance(periodicity, (bool, np.bool_)):
            periodicity = (bool(periodicity),) * self.ndim
        self.periodicity = tuple(bool(p) for p in periodicity)
        self.center_level_0 = (cmin + cmax) / 2
        self.length_level_0 = float(np.max(cmax - cmin))
        self.initial_refinement_level = int(initial_refinement_level)
        self.n_cells_max = n_cells_max
        nd, L = self.ndim, self.initial_refinement_level
        n = 1 << (nd * L)
        self.levels = np.full(n, L, dtype=np.int64)
        self.icoords = _deinterleave(np.arange(n, dtype=np.int64), nd, L)
        for patch in refinement_patches:
            if patch.get("type", "box") != "box":
                raise NotImplementedError("only refinement_patches of type 'box' are supported")
            self._refine_box(np.asarray(patch["coordinates_min"], dtype=np.float64),
                             np.asarray(patch["coordinates_max"], dtype=np.float64))
        if n_cells_max is not None and self.n_leaf_cells > n_cells_max:
            raise ValueError("TreeMesh: n_cells_max exceeded")

    # ------------------------------------------------------------------ geometry
    @property
    def n_leaf_cells(self):
        return int(self.levels.shape[0])

    def length_at_level(self, levels):
        return self.length_level_0 / (2.0 ** np.asarray(levels, dtype=np.float64))

    def cell_centers(self):
        """Cell midpoints, accumulated root-to-leaf like Trixi's `refine!` does (exact same float ops)."""
        nd = self.ndim
        x = np.tile(self.center_level_0, (self.n_leaf_cells, 1))
        maxlev = int(self.levels.max()) if self.n_leaf_cells else 0
        for lev in range(1, maxlev + 1):
            m = self.levels >= lev
            dx = self.length_level_0 / float(1 << lev)
            for d in range(nd):
                bit = (self.icoords[m, d] >> (self.levels[m] - lev)) & 1
                x[m, d] += np.where(bit == 1, 1.0, -1.0) * dx / 2
        return x

    def _keys(self, levels, ic):
        nd = self.ndim
        return _interleave(ic, nd, self.MAX_LEVEL) << (nd * (self.MAX_LEVEL - levels))

    def _sort(self):
        order = np.argsort(self._keys(self.levels, self.icoords), kind="stable")
        self.levels = self.levels[order]
        self.icoords = self.icoords[order]

    def _locate(self, levels, ic):
        """Index of the leaf covering the cell (level, ic), -1 where the position is outside the domain
        (non-periodic). Also returns that leaf's level."""
        nd = self.ndim
        ic = ic.copy()
        n = (np.int64(1) << levels)
        valid = np.ones(levels.shape[0], dtype=bool)
        for d in range(nd):
            out = (ic[:, d] < 0) | (ic[:, d] >= n)
            if self.periodicity[d]:
                ic[:, d] = np.where(out, (ic[:, d] + n) % n, ic[:, d])
            else:
                valid &= ~out
                ic[:, d] = np.clip(ic[:, d], 0, n - 1)
        keys = self._keys(levels, ic)
        idx = np.searchsorted(self._leaf_keys, keys, side="right") - 1
        return np.where(valid, idx, -1), np.where(valid, self.levels[idx], -1)

    def _refine(self, mask):
        nd = self.ndim
        nchild = 1 << nd
        keep_l, keep_ic = self.levels[~mask], self.icoords[~mask]
        pl, pic = self.levels[mask], self.icoords[mask]
        cl = np.repeat(pl + 1, nchild)
        cic = np.repeat(pic * 2, nchild, axis=0)
        k = np.tile(np.arange(nchild, dtype=np.int64), pl.shape[0])
        for d in range(nd):
            cic[:, d] += (k >> d) & 1
        self.levels = np.concatenate([keep_l, cl])
        self.icoords = np.concatenate([keep_ic, cic])
        self._sort()

    def _rebalance(self):
        nd = self.ndim
        while True:
            self._leaf_keys = self._keys(self.levels, self.icoords)
            need = np.zeros(self.n_leaf_cells, dtype=bool)
            for direction in range(1, 2 * nd + 1):
                d = (direction - 1) // 2
                ic = self.icoords.copy()
                ic[:, d] += 1 if direction % 2 == 0 else -1
                idx, lev = self._locate(self.levels, ic)
                bad = (idx >= 0) & (lev < self.levels - 1)
                need[idx[bad]] = True
            if not need.any():
                return
            self._refine(need)

    def _refine_box(self, lo, hi):
        x = self.cell_centers()
        inside = np.ones(self.n_leaf_cells, dtype=bool)
        for d in range(self.ndim):
            inside &= (lo[d] < x[:, d]) & (x[:, d] < hi[d])
        if inside.any():
            self._refine(inside)
        self._rebalance()


@dataclass
class ElementContainer:
    inverse_jacobian: np.ndarray          # [E]
    node_coordinates: np.ndarray          # [ndim, N.., E]  (Fortran order, flat view via .ravel(order="F"))
    cell_ids: np.ndarray                  # [E] (1-based leaf numbers; Trixi stores tree cell ids)
    levels: np.ndarray


@dataclass
class InterfaceContainer:
    neighbor_ids: np.ndarray              # [2, I] int64 1-based
    orientations: np.ndarray              # [I]


@dataclass
class BoundaryContainer:
    neighbor_ids: np.ndarray              # [B]
    orientations: np.ndarray
    neighbor_sides: np.ndarray
    node_coordinates: np.ndarray          # [ndim, N.., B]
    n_boundaries_per_direction: np.ndarray  # [2*ndim]


@dataclass
class MortarContainer:
    neighbor_ids: np.ndarray              # [2^(ndim-1)+1, M]
    large_sides: np.ndarray
    orientations: np.ndarray


@dataclass
class Containers:
    elements: ElementContainer
    interfaces: InterfaceContainer
    boundaries: BoundaryContainer
    mortars: MortarContainer
    ndim: int = 3
    nnodes: int = 4
    extra: dict = field(default_factory=dict)


# small-children tables, 1-based child numbers as in Trixi
_CHILD3 = np.array([[2, 4, 6, 8], [1, 3, 5, 7], [3, 4, 7, 8], [1, 2, 5, 6], [5, 6, 7, 8], [1, 2, 3, 4]]) - 1
_CHILD2 = np.array([[2, 4], [1, 3], [3, 4], [1, 2]]) - 1


def init_containers(mesh: TreeMesh, nodes: np.ndarray) -> Containers:
    """Trixi's init_elements / init_interfaces / init_boundaries / init_mortars on the leaf cells."""
    nd, N = mesh.ndim, int(nodes.shape[0])
    E = mesh.n_leaf_cells
    mesh._leaf_keys = mesh._keys(mesh.levels, mesh.icoords)
    levels, ic = mesh.levels, mesh.icoords

    # --- elements
    dx = mesh.length_at_level(levels)
    jac = dx / 2
    inverse_jacobian = 1.0 / jac
    centers = mesh.cell_centers()
    shape = (nd,) + (N,) * nd + (E,)
    node_coordinates = np.empty(shape, dtype=np.float64, order="F")
    for d in range(nd):
        sh = [1] * (nd + 1)
        sh[d] = N
        xd = centers[:, d].reshape([1] * nd + [E]) + jac.reshape([1] * nd + [E]) * nodes.reshape(sh)
        node_coordinates[d] = np.broadcast_to(xd, shape[1:])
    elements = ElementContainer(inverse_jacobian, node_coordinates, np.arange(1, E + 1, dtype=np.int64), levels.copy())

    # --- neighbours per direction: leaf index covering the same-level position and its level
    nb_idx = np.empty((2 * nd, E), dtype=np.int64)
    nb_lev = np.empty((2 * nd, E), dtype=np.int64)
    for direction in range(1, 2 * nd + 1):
        d = (direction - 1) // 2
        q = ic.copy()
        q[:, d] += 1 if direction % 2 == 0 else -1
        nb_idx[direction - 1], nb_lev[direction - 1] = mesh._locate(levels, q)

    # --- interfaces: element-outer, positive directions, neighbour is a leaf of the same level
    pos = np.arange(1, 2 * nd, 2)                      # rows of directions 2,4,6
    m = (nb_idx[pos] >= 0) & (nb_lev[pos] == levels[None, :])   # [nd, E]
    mt = m.T                                            # [E, nd] -> C-order flatten = element-outer
    e_idx, d_idx = np.nonzero(mt)
    left = e_idx + 1
    right = nb_idx[pos].T[mt] + 1
    interfaces = InterfaceContainer(np.asfortranarray(np.stack([left, right]).astype(np.int64)),
                                    (d_idx + 1).astype(np.int64))

    # --- boundaries: direction-outer, element-inner; no neighbour and no coarse neighbour
    mb = nb_idx < 0                                     # [2nd, E]
    dir_idx, e_idx = np.nonzero(mb)
    B = e_idx.shape[0]
    bshape = (nd,) + (N,) * (nd - 1) + (B,)
    bnc = np.empty(bshape, dtype=np.float64, order="F")
    for b in range(B):                                  # boundaries are few (surface of the domain)
        d = dir_idx[b] // 2
        sl = [slice(None)] * (nd + 2)
        sl[1 + d] = N - 1 if dir_idx[b] % 2 == 1 else 0
        sl[-1] = e_idx[b]
        bnc[..., b] = node_coordinates[tuple(sl)]
    boundaries = BoundaryContainer((e_idx + 1).astype(np.int64), (dir_idx // 2 + 1).astype(np.int64),
                                   np.where(dir_idx % 2 == 1, 1, 2).astype(np.int64), bnc,
                                   np.bincount(dir_idx, minlength=2 * nd).astype(np.int64))

    # --- mortars: element-outer, all directions; same-level neighbour position is covered by finer leaves
    nsmall = 1 << (nd - 1) if nd > 1 else 0
    if nd > 1:
        mm = ((nb_idx >= 0) & (nb_lev > levels[None, :])).T          # [E, 2nd]
        e_idx, dir0 = np.nonzero(mm)
        M = e_idx.shape[0]
        nids = np.empty((nsmall + 1, M), dtype=np.int64, order="F")
        nids[nsmall] = e_idx + 1
        table = _CHILD3 if nd == 3 else _CHILD2
        # neighbour cell (same level as large) integer coords, periodic-wrapped
        q = ic[e_idx].copy()
        dd = dir0 // 2
        q[np.arange(M), dd] += np.where(dir0 % 2 == 1, 1, -1)
        n = np.int64(1) << levels[e_idx]
        for d in range(nd):
            q[:, d] = (q[:, d] + n) % n
        for s in range(nsmall):
            k = table[dir0, s]
            cq = q * 2
            for d in range(nd):
                cq[:, d] += (k >> d) & 1
            idx, lev = mesh._locate(levels[e_idx] + 1, cq)
            if not np.all(lev == levels[e_idx] + 1):
                raise RuntimeError("TreeMesh is not 2:1 balanced")
            nids[s] = idx + 1
        mortars = MortarContainer(nids, np.where(dir0 % 2 == 1, 1, 2).astype(np.int64),
                                  (dir0 // 2 + 1).astype(np.int64))
    else:
        mortars = MortarContainer(np.zeros((1, 0), dtype=np.int64, order="F"), np.zeros(0, dtype=np.int64),
                                  np.zeros(0, dtype=np.int64))
    return Containers(elements, interfaces, boundaries, mortars, nd, N)
