"""CPU checks of the drop-in boundary: libtrixib200.so loads and exports every symbol include/trixib200.h
declares; the ctypes structs match the header; on a machine without a GPU create() fails loudly (no CPU
fallback exists); the product package never touches oracle/."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "trixib200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(trixib200_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported():
    from trixib200 import _lib
    _lib.build()
    L = _lib.lib()
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), f"libtrixib200.so lacks {n}"
    assert sorted(_lib.EXPORTS) == names, "python EXPORTS list and header disagree"
    assert L.trixib200_version() == 100


def test_struct_layouts_match_header():
    """Field names and order of the ctypes mirrors follow the C structs (same field names in the header)."""
    from trixib200 import _lib
    src = open(os.path.join(ROOT, "include", "trixib200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    for cname, ct in (("trixib200_config", _lib.Config), ("trixib200_basis_host", _lib.BasisHost),
                      ("trixib200_mesh_host", _lib.MeshHost)):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (cname, cname), src, flags=re.S).group(1)
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            for piece in decl.split(","):
                m = re.search(r"([A-Za-z_][A-Za-z_0-9]*)\s*(\[\d+\])?\s*$", piece.strip())
                fields.append(m.group(1))
        assert fields == [f[0] for f in ct._fields_], cname
    assert C.sizeof(_lib.Config) == 4 * 22 + 8 * 7


def test_create_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a machine without a CUDA device")
    import trixib200 as T
    import cases
    with pytest.raises(RuntimeError):
        cases.make_semi(cases.CASES["c5_euler_ec_3d"])
    # and straight through the C ABI: create() returns an error code and a message, never a handle
    from trixib200 import _lib
    L = _lib.lib()
    cfg, bh, mh = _lib.Config(), _lib.BasisHost(), _lib.MeshHost()
    cfg.ndim, cfg.polydeg, cfg.equations, cfg.nranks = 3, 3, _lib.EQ_EULER, 1
    cfg.surface_flux = _lib.FLUX["flux_ranocha"]
    bh.nnodes = 4
    mh.nelements = 8
    h = C.c_void_p()
    rc = L.trixib200_create(C.byref(cfg), C.byref(bh), C.byref(mh), C.byref(h))
    assert rc < 0 and not h.value
    assert b"CUDA" in L.trixib200_last_error() or b"cuda" in L.trixib200_last_error()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "trixicuda.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "liboracle" not in text and "oracle/" not in text, f
