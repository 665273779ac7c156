"""Solution / mesh files in Trixi's HDF5 layout (reference examples/euler_ec_3d.jl:38-41).
CPU-only: the container writer, the tree export and the callback's file contents; the files are read back with the
module's own reader (no libhdf5 in the image -- DESIGN.md section 7 says what that does and does not prove)."""
import os
import struct

import numpy as np
import pytest

import trixib200 as T
from trixib200 import hdf5_lite, solution_file as S


def test_hdf5_container_structure_and_round_trip(tmp_path):
    f = hdf5_lite.File()
    f.attrs["ndims"] = 3
    f.attrs["equations"] = "CompressibleEulerEquations3D"
    f.attrs["time"] = 0.4
    f.attrs["periodicity"] = np.array([True, False, True])
    f.attrs["center_level_0"] = np.array([0.0, 0.5, -1.0])
    rng = np.random.default_rng(0)
    arrays = {f"variables_{v + 1}": rng.standard_normal(37) for v in range(11)}       # 11 links: "variables_10" < "variables_2"
    arrays["child_ids"] = np.arange(24, dtype=np.int64).reshape(3, 8)
    arrays["levels32"] = np.arange(5, dtype=np.int32)
    for k, a in arrays.items():
        d = f.create_dataset(k, a)
        d.attrs["name"] = "name of " + k
    p = str(tmp_path / "t.h5")
    f.write(p)
    b = open(p, "rb").read()
    # superblock: signature, version 0, 8-byte offsets and lengths, end-of-file address = file size
    assert b[:8] == b"\x89HDF\r\n\x1a\n" and b[8] == 0 and b[13] == 8 and b[14] == 8
    assert struct.unpack_from("<Q", b, 40)[0] == len(b)
    a_root, _, a_btree, a_heap = struct.unpack_from("<QI4xQQ", b, 64)
    assert a_root % 8 == 0 and b[a_btree:a_btree + 4] == b"TREE" and b[a_heap:a_heap + 4] == b"HEAP"
    attrs, ds = hdf5_lite.File.read(p)          # the reader asserts sorted links, aligned data, header sizes
    assert attrs["ndims"] == 3 and isinstance(attrs["ndims"], int)
    assert attrs["equations"] == "CompressibleEulerEquations3D" and attrs["time"] == 0.4
    assert attrs["periodicity"].tolist() == [True, False, True]
    assert np.array_equal(attrs["center_level_0"], [0.0, 0.5, -1.0])
    assert sorted(ds) == sorted(arrays)
    for k, a in arrays.items():
        got, at = ds[k]
        assert got.dtype == a.dtype and np.array_equal(got, a) and at["name"] == "name of " + k


@pytest.mark.parametrize("nd,level,patch", [(3, 2, True), (2, 3, True), (1, 4, False), (3, 2, False)])
def test_tree_export_matches_the_leaf_mesh(nd, level, patch):
    """Whole tree in Trixi's storage order: leaves in mesh order with identical levels / centres, parents directly
    before their first child, child and neighbour links consistent (periodic wrap included)."""
    patches = (dict(type="box", coordinates_min=(-0.5,) * nd, coordinates_max=(0.5,) * nd),) if patch else ()
    per = (True, False, True)[:nd] if not patch else True
    mesh = T.TreeMesh((-1.0,) * nd, (1.0,) * nd, initial_refinement_level=level, refinement_patches=patches,
                      periodicity=per, n_cells_max=10 ** 5)
    t = S.tree_arrays(mesh)
    n = t["levels"].size
    leaf = (t["child_ids"] == 0).all(1)
    assert leaf.sum() == mesh.n_leaf_cells
    assert np.array_equal(t["levels"][leaf], mesh.levels)
    assert np.array_equal(t["coordinates"][leaf], mesh.cell_centers()[:, :nd])
    assert t["parent_ids"][0] == 0 and t["levels"][0] == 0 and (t["parent_ids"][1:] > 0).all()
    for i in range(n):
        kids = t["child_ids"][i]
        if kids.any():
            assert (kids > 0).all() and kids[0] == i + 2                 # first child directly behind its parent
            assert (t["parent_ids"][kids - 1] == i + 1).all() and (t["levels"][kids - 1] == t["levels"][i] + 1).all()
            assert (np.diff(kids) > 0).all()
        for d in range(nd):
            for side in (0, 1):
                q = t["neighbor_ids"][i, 2 * d + side]
                if q:
                    assert t["levels"][q - 1] == t["levels"][i]
                    assert t["neighbor_ids"][q - 1, 2 * d + 1 - side] == i + 1       # mutual
                    dx = 2.0 / 2 ** t["levels"][i]
                    step = (t["coordinates"][q - 1, d] - t["coordinates"][i, d]) * (1 if side else -1)
                    assert abs(step - dx) < 1e-14 or abs(step - dx + 2.0) < 1e-14    # neighbour or periodic image
    if per is not True and nd >= 2:
        root_level_cells = np.where(t["levels"] == level)[0]
        lo = t["coordinates"][root_level_cells, 1].min()
        on_wall = root_level_cells[t["coordinates"][root_level_cells, 1] == lo]
        assert (t["neighbor_ids"][on_wall, 2] == 0).all()                # direction 3 (-y) of a non-periodic wall


def test_solution_file_layout(tmp_path):
    """`save_solution_file` with cons2prim: attributes and the per-variable vectors (node index i fastest, element
    slowest = Julia's vec(data[v, .., :])) -- written from a host vector, no GPU needed."""
    class Semi:          # the attributes save_solution_file reads
        pass
    semi = Semi()
    semi.mesh = T.TreeMesh((-2.0,) * 3, (2.0,) * 3, initial_refinement_level=1, n_cells_max=1000)
    semi.equations = T.CompressibleEulerEquations3D(1.4)
    semi.nnodes, semi.nvars = 4, 5
    E = 8
    rng = np.random.default_rng(1)
    prim = np.empty((E, 4, 4, 4, 5))
    prim[..., 0] = rng.uniform(0.5, 2.0, (E, 4, 4, 4))
    prim[..., 1:4] = rng.uniform(-1, 1, (E, 4, 4, 4, 3))
    prim[..., 4] = rng.uniform(0.5, 2.0, (E, 4, 4, 4))
    cons = prim.copy()
    cons[..., 1:4] = prim[..., :1] * prim[..., 1:4]
    cons[..., 4] = prim[..., 4] / 0.4 + 0.5 * prim[..., 0] * (prim[..., 1:4] ** 2).sum(-1)
    path = S.save_solution_file(cons.ravel(), 0.25, 1e-3, 17, semi, S.cons2prim, str(tmp_path),
                                element_variables={"indicator_shock_capturing": np.linspace(0, 0.5, E)})
    assert os.path.basename(path) == "solution_000000017.h5"
    attrs, data, names, elem = S.load_solution_file(path)
    assert attrs == {"ndims": 3, "equations": "CompressibleEulerEquations3D", "polydeg": 3, "n_vars": 5,
                     "n_elements": 8, "mesh_type": "TreeMesh", "mesh_file": "mesh.h5", "time": 0.25, "dt": 1e-3,
                     "timestep": 17}
    assert names == ["rho", "v1", "v2", "v3", "p"]
    assert np.abs(data - prim).max() <= 1e-14
    assert np.array_equal(elem["indicator_shock_capturing"], np.linspace(0, 0.5, E))
    # raw layout: variables_2 is v1 with i fastest
    _, ds = hdf5_lite.File.read(path)
    assert np.abs(ds["variables_2"][0] - prim[..., 1].ravel()).max() <= 1e-14
    assert S.varnames(S.cons2cons, semi.equations) == ("rho", "rho_v1", "rho_v2", "rho_v3", "rho_e")
