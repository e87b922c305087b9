"""CPU-side checks: the C-ABI library loads, exports every symbol that
include/pinnfem.h declares, refuses to compute without a GPU, and builds the
mesh plan (pure integer work) bit-exactly like the oracle."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest

from oracle import pinnfem_oracle as O

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "pinnfem.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pf_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    from pinn_fem_b200 import _lib

    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"libpinnfem.so does not export {n}"
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)
    assert lib.pf_version() == 100


def test_gd_config_struct_matches_header():
    from pinn_fem_b200 import _lib

    # 2 int32 + 6 double + 4*3 int32 + 3 double + 2 int32 = 8 + 48 + 48 + 24 + 8
    assert ctypes.sizeof(_lib.GDConfig) == 136


def _mesh(name):
    if name == "lattice":
        return O.lattice_truss(9, 6)
    if name == "ex1":
        return (np.array([[0.0, 0], [1, 0], [2, 0], [3, 0]]), np.array([[0, 1], [1, 2], [2, 3]]),
                np.array([0, 1, 3, 5, 7]))
    rng = np.random.default_rng(4)
    nodes = rng.normal(size=(40, 2))
    el = []
    while len(el) < 120:
        i, j = rng.integers(0, 40, size=2)
        if i != j:
            el.append((i, j))  # random graph: duplicates and both orientations happen
    return nodes, np.array(el), rng.integers(0, 80, size=13)


@pytest.mark.parametrize("name", ["ex1", "lattice", "random_dups"])
def test_plan_index_arrays_bit_exact(name):
    from pinn_fem_b200 import AssemblyPlan

    nodes, el, fixed = _mesh(name)
    nnode = len(nodes)
    p = AssemblyPlan(nodes, el, fixed)
    assert (p.nnode, p.nelem, p.ndof, p.dim) == (nnode, len(el), 2 * nnode, 2)
    assert np.array_equal(p.elem_dofs, O.all_element_dofs(el)) and p.elem_dofs.dtype == np.int64
    free, fx = O.free_and_fixed_dofs(2 * nnode, fixed)
    assert np.array_equal(p.free_dofs, free) and np.array_equal(p.fixed_dofs, fx)
    rowptr, colind, slots = O.bsr_pattern(nnode, el)
    assert np.array_equal(p.bsr_rowptr, rowptr) and np.array_equal(p.bsr_colind, colind)
    assert np.array_equal(p.elem_slots, slots)
    ptr, inc_elem, inc_end = O.node_incidence(nnode, el)
    assert np.array_equal(p.inc_ptr, ptr) and np.array_equal(p.inc_elem, inc_elem)
    assert np.array_equal(p.inc_nbr, el[inc_elem, 1 - inc_end])
    assert p.nnzb == len(colind) and p.ninc == 2 * len(el)
    assert p.has_duplicate_edges == (name == "random_dups")
    # every incidence's slot is the (node, nbr) block; the diagonal slot is (node, node)
    rows = np.repeat(np.arange(nnode), np.diff(p.inc_ptr))
    assert np.array_equal(p.bsr_colind[p.inc_slot], p.inc_nbr)
    assert np.all((p.inc_slot >= p.bsr_rowptr[rows]) & (p.inc_slot < p.bsr_rowptr[rows + 1]))
    assert np.array_equal(p.bsr_colind[p.diag_slot], np.arange(nnode))
    g = O.element_geometry(nodes, el)
    assert np.array_equal(p.geometry("l0"), g.l0) and np.array_equal(p.geometry("cos"), g.c)
    assert np.array_equal(p.geometry("centroid"), (nodes[el[:, 0]] + nodes[el[:, 1]]) / 2.0)


def test_plan_1d_and_errors():
    from pinn_fem_b200 import AssemblyPlan

    p = AssemblyPlan(np.array([0.0, 0.7, 1.5, 3.0]), [[0, 1], [1, 2], [2, 3], [0, 2]], [0])
    assert p.dim == 1 and p.ndof == 4 and np.array_equal(p.elem_dofs, [[0, 1], [1, 2], [2, 3], [0, 2]])
    assert p.free_dofs.tolist() == [1, 2, 3]
    with pytest.raises(ValueError, match="zero initial length"):
        AssemblyPlan(np.array([[0.0, 0], [0, 0]]), [[0, 1]], [])
    with pytest.raises(ValueError, match="out-of-range"):
        AssemblyPlan(np.array([[0.0, 0], [1, 0]]), [[0, 1]], [9])
    with pytest.raises(ValueError, match="outside"):
        AssemblyPlan(np.array([[0.0, 0], [1, 0]]), [[0, 2]], [])


def test_no_cpu_fallback():
    """Without a CUDA device every compute entry point must fail loudly."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from pinn_fem_b200 import AssemblyPlan, PinnFemError

    nodes, el, fixed = _mesh("ex1")
    p = AssemblyPlan(nodes, el, fixed)
    with pytest.raises(PinnFemError, match="no CUDA device"):
        p.to("cuda")
    u = torch.zeros(8, dtype=torch.float64)
    with pytest.raises(PinnFemError, match="no CPU fallback"):
        p.internal_force(u, torch.ones(3, dtype=torch.float64), torch.ones(3, dtype=torch.float64))


def test_product_does_not_import_oracle():
    for path in (ROOT / "pinn_fem_b200").rglob("*.py"):
        src = path.read_text()
        assert "oracle" not in re.sub(r"#.*", "", src).replace("pinnfem_oracle", "oracle") or \
            not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), path
