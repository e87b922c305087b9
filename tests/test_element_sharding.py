"""Element sharding: the partition / halo / exchange-list logic is pure integer work and
is checked here on CPU; the oracle evaluated on each rank's local mesh must reproduce the owned rows of
the global internal force bit for bit (same element order per node)."""
import numpy as np
import pytest

from oracle import pinnfem_oracle as O
from pinn_fem_b200.element_sharding import contiguous_node_partition, coordinate_bisection_partition, partition_mesh
from pinn_fem_b200.meshes import lattice_truss


def _random_mesh(seed):
    rng = np.random.default_rng(seed)
    nodes = rng.normal(size=(57, 2))
    el = []
    while len(el) < 160:
        i, j = rng.integers(0, 57, size=2)
        if i != j:
            el.append((i, j))
    return nodes, np.array(el), rng.integers(0, 114, size=9)


@pytest.mark.parametrize("world", [1, 2, 3, 8])
@pytest.mark.parametrize("mesh", ["lattice", "random", "random_part", "random_rcb"])
def test_partition_is_consistent(world, mesh):
    part = None
    if mesh == "lattice":
        nodes, el, fixed = lattice_truss(12, 9)
    else:
        nodes, el, fixed = _random_mesh(3)
        if mesh == "random_part":
            part = np.random.default_rng(1).integers(0, world, size=len(nodes)).astype(np.int32)
        elif mesh == "random_rcb":
            part = coordinate_bisection_partition(nodes, world)
    locs = partition_mesh(nodes, el, fixed, world, part)
    p = contiguous_node_partition(len(nodes), world) if part is None else part
    # every node owned once, every element owned once, local elements ascending in global id
    assert np.array_equal(np.sort(np.concatenate([m.owned for m in locs])), np.arange(len(nodes)))
    owned_el = np.concatenate([m.elements_global[m.elem_owned == 1] for m in locs])
    assert np.array_equal(np.sort(owned_el), np.arange(len(el)))
    fixed_u = np.unique(fixed)
    for m in locs:
        assert np.all(np.diff(m.elements_global) > 0) and np.all(np.diff(m.owned) > 0) and np.all(np.diff(m.halo) > 0)
        assert np.all(p[m.owned] == m.rank) and np.all(p[m.halo] != m.rank)
        lg = m.local_nodes_global
        assert np.array_equal(lg[m.elements], np.asarray(el)[m.elements_global])          # same edges, same orientation
        # an element is local iff it touches an owned node
        touches = (p[np.asarray(el)[:, 0]] == m.rank) | (p[np.asarray(el)[:, 1]] == m.rank)
        assert np.array_equal(np.flatnonzero(touches), m.elements_global)
        assert np.array_equal(np.sort(m.local_dofs_global()[m.fixed_dofs]), np.intersect1d(m.local_dofs_global(), fixed_u))
        assert m.nfree_global == 2 * len(nodes) - fixed_u.size
        # exchange lists: what r sends to q is exactly what q expects from r, in the same order
        for k, q in enumerate(m.peers):
            other = locs[int(q)]
            kk = int(np.flatnonzero(other.peers == m.rank)[0])
            sent = lg[m.send_nodes[m.send_ptr[k]:m.send_ptr[k + 1]]]
            expected = other.local_nodes_global[other.recv_nodes[other.recv_ptr[kk]:other.recv_ptr[kk + 1]]]
            assert np.array_equal(sent, expected)
            assert np.all(m.send_nodes[m.send_ptr[k]:m.send_ptr[k + 1]] < m.n_owned)     # owned rows go out
            assert np.all(m.recv_nodes[m.recv_ptr[k]:m.recv_ptr[k + 1]] >= m.n_owned)    # halo rows come in
        assert np.array_equal(np.sort(m.recv_nodes), np.arange(m.n_owned, m.n_owned + m.halo.size))  # every halo node once


@pytest.mark.parametrize("world", [2, 3, 5])
def test_local_meshes_reproduce_global_internal_force_bitwise(world):
    nodes, el, fixed = lattice_truss(11, 7)
    rng = np.random.default_rng(0)
    u = rng.uniform(-1e-3, 1e-3, 2 * len(nodes))
    E = rng.uniform(0.5, 1.5, len(el))
    A = rng.uniform(0.5, 1.5, len(el))
    f_ref, _ = O.assemble_residual(nodes, el, E, A, u)
    got = np.full_like(f_ref, np.nan)
    for m in partition_mesh(nodes, el, fixed, world):
        f_loc, _ = O.assemble_residual(m.nodes, m.elements, E[m.elements_global], A[m.elements_global],
                                       m.to_local_vector(u))
        nd = m.n_owned * m.dim
        got[m.local_dofs_global()[:nd]] = f_loc[:nd]
    assert np.array_equal(got, f_ref)


@pytest.mark.parametrize("world", [1, 2, 3, 5, 8])
def test_coordinate_bisection_partition_is_balanced_and_compact(world):
    """Recursive coordinate bisection: rank sizes as balanced as contiguous ranges, deterministic, and -- on a
    lattice whose nodes are numbered at random -- far fewer halo nodes than contiguous id ranges; the local
    meshes still reproduce the global internal force bit for bit."""
    nodes, el, fixed = lattice_truss(16, 12)
    rng = np.random.default_rng(4)
    perm = rng.permutation(len(nodes))            # new id of old node i
    inv = np.argsort(perm)
    nodes_s, el_s = nodes[inv], perm[np.asarray(el)]
    fixed_s = (perm[np.asarray(fixed) // 2] * 2 + np.asarray(fixed) % 2)
    part = coordinate_bisection_partition(nodes_s, world)
    assert np.array_equal(part, coordinate_bisection_partition(nodes_s, world))
    sizes = np.bincount(part, minlength=world)
    assert sizes.max() - sizes.min() <= 1 and sizes.sum() == len(nodes)
    locs = partition_mesh(nodes_s, el_s, fixed_s, world, part)
    if world > 1:
        halo_rcb = sum(m.halo.size for m in locs)
        halo_ids = sum(m.halo.size for m in partition_mesh(nodes_s, el_s, fixed_s, world))
        assert halo_rcb < 0.5 * halo_ids
    u = rng.uniform(-1e-3, 1e-3, 2 * len(nodes))
    E = rng.uniform(0.5, 1.5, len(el))
    A = rng.uniform(0.5, 1.5, len(el))
    f_ref, _ = O.assemble_residual(nodes_s, el_s, E, A, u)
    got = np.full_like(f_ref, np.nan)
    for m in locs:
        f_loc, _ = O.assemble_residual(m.nodes, m.elements, E[m.elements_global], A[m.elements_global], m.to_local_vector(u))
        nd = m.n_owned * m.dim
        got[m.local_dofs_global()[:nd]] = f_loc[:nd]
    assert np.array_equal(got, f_ref)
    # 1-D coordinates and the error contract
    assert np.array_equal(coordinate_bisection_partition(np.arange(10.0), 2), np.repeat([0, 1], 5))
    with pytest.raises(ValueError):
        coordinate_bisection_partition(nodes[:3], 4)


def test_bad_partition_is_rejected():
    nodes, el, fixed = lattice_truss(4)
    with pytest.raises(ValueError):
        partition_mesh(nodes, el, fixed, 2, part=np.full(len(nodes), 2, dtype=np.int32))
    with pytest.raises(ValueError):  # a rank without nodes
        partition_mesh(nodes, el, fixed, 2, part=np.zeros(len(nodes), dtype=np.int32))
    with pytest.raises(ValueError):
        partition_mesh(nodes[:3], el[:0], [], 4)
