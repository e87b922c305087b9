"""CPU oracle for the PINN-FEM hot path -- TEST INFRASTRUCTURE ONLY.

This module is a NumPy restatement of the reference's algorithm for the path
BASELINE.json names.  It is the *checker*: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  The product (``pinn_fem_b200``)
never imports anything under ``oracle/``; it fails loudly without its CUDA
library.

Parity status: PINNED.  ``tests/golden/make_golden.py`` imports the real
reference from ``/root/reference/FEM/python`` (it is pure Python) and stores
its outputs under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks
every function here against those fixtures (fp64 NumPy path: ~1e-14; fp32
torch path: fp32 round-off) and against the known-answer cases of the
reference's own ``test_torch_element.py``.

Every function cites the reference file:line (relative to
``/root/reference/FEM/python``) whose arithmetic it restates.  Arithmetic is
dtype-parameterised: ``float64`` is what the CUDA kernels are compared
against (north star: fp64, 1e-10 relative), ``float32`` mimics the reference's
hard-coded torch dtype.

Layout convention shared with the CUDA library: batched arrays carry the
problem index LAST: ``u[ndof, B]``, ``E[nelem, B]``, ``f_int[ndof, B]``.
All functions also accept the un-batched 1-D forms.
"""

from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

LINEAR = 0  # fem/element.py:15-102 (small displacement)
GREEN_LAGRANGE = 1  # fem/element.py:105-133 (defined, never wired: SURVEY D1)


# ----------------------------------------------------------------------------
# Integer / index work (must be bit-exact)
# ----------------------------------------------------------------------------


def element_dofs(node_i: int, node_j: int, dim: int = 2) -> np.ndarray:
    """fem/geometry.py:8-9 (2-D) and fem/assembly.py:26-28 (1-D: dof == node)."""
    if dim == 1:
        return np.array([node_i, node_j], dtype=np.int64)
    return np.array([2 * node_i, 2 * node_i + 1, 2 * node_j, 2 * node_j + 1], dtype=np.int64)


def all_element_dofs(elements: np.ndarray, dim: int = 2) -> np.ndarray:
    """Vectorised ``element_dofs`` for every element: int64 [nelem, 2*dim]."""
    el = np.asarray(elements, dtype=np.int64).reshape(-1, 2)
    if dim == 1:
        return el.copy()
    return np.stack([2 * el[:, 0], 2 * el[:, 0] + 1, 2 * el[:, 1], 2 * el[:, 1] + 1], axis=1)


def free_and_fixed_dofs(ndof: int, fixed_dofs) -> Tuple[np.ndarray, np.ndarray]:
    """fem/boundary.py:8-13: sorted-unique fixed, ascending complement free."""
    fixed = np.unique(np.asarray(fixed_dofs, dtype=np.int64).reshape(-1))
    mask = np.ones(ndof, dtype=bool)
    mask[fixed] = False
    return np.flatnonzero(mask).astype(np.int64), fixed.astype(np.int64)


def structural_pattern(nnode: int, elements: np.ndarray, dim: int = 2) -> np.ndarray:
    """Boolean dense pattern touched by ``K[np.ix_(dofs, dofs)] += ke``
    (fem/assembly.py:48,:71; fem/nn_assembly.py:226-229)."""
    ndof = nnode * dim
    pat = np.zeros((ndof, ndof), dtype=bool)
    for dofs in all_element_dofs(elements, dim):
        pat[np.ix_(dofs, dofs)] = True
    return pat


def bsr_pattern(nnode: int, elements: np.ndarray) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Node-block CSR pattern of the assembled tangent.

    Row ``n`` holds block columns ``{n} U neighbours(n)`` in ascending order.
    Returns ``(rowptr[nnode+1], colind[nnzb], elem_slots[nelem,4])`` where
    ``elem_slots[e] = (slot(i,i), slot(i,j), slot(j,i), slot(j,j))`` are the
    positions in ``colind`` that element ``e=(i,j)`` adds into -- the scatter
    map of fem/assembly.py:71 expressed on 2x2 (or 1x1) node blocks.
    """
    el = np.asarray(elements, dtype=np.int64).reshape(-1, 2)
    cols: List[set] = [{n} for n in range(nnode)]
    for i, j in el:
        cols[i].add(int(j))
        cols[j].add(int(i))
    rowptr = np.zeros(nnode + 1, dtype=np.int64)
    for n in range(nnode):
        rowptr[n + 1] = rowptr[n] + len(cols[n])
    colind = np.empty(rowptr[-1], dtype=np.int64)
    lookup: List[Dict[int, int]] = []
    for n in range(nnode):
        srt = sorted(cols[n])
        colind[rowptr[n] : rowptr[n + 1]] = srt
        lookup.append({c: int(rowptr[n]) + k for k, c in enumerate(srt)})
    slots = np.empty((len(el), 4), dtype=np.int64)
    for e, (i, j) in enumerate(el):
        slots[e] = (lookup[i][i], lookup[i][j], lookup[j][i], lookup[j][j])
    return rowptr, colind, slots


def node_incidence(nnode: int, elements: np.ndarray) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Node -> incident element lists in ascending element order.

    ``f_int[dofs] += fe`` (fem/assembly.py:72) visits elements in order, so for
    every DOF the floating-point sum runs over its incident elements in
    ascending element id; ``inc_elem`` encodes exactly that order.  ``inc_end``
    is 0 when the node is the element's first node, 1 when it is the second.
    """
    el = np.asarray(elements, dtype=np.int64).reshape(-1, 2)
    deg = np.zeros(nnode + 1, dtype=np.int64)
    for i, j in el:
        deg[i + 1] += 1
        deg[j + 1] += 1
    ptr = np.cumsum(deg)
    fill = ptr[:-1].copy()
    inc_elem = np.empty(ptr[-1], dtype=np.int64)
    inc_end = np.empty(ptr[-1], dtype=np.int64)
    for e, (i, j) in enumerate(el):
        inc_elem[fill[i]] = e
        inc_end[fill[i]] = 0
        fill[i] += 1
        inc_elem[fill[j]] = e
        inc_end[fill[j]] = 1
        fill[j] += 1
    return ptr, inc_elem, inc_end


# ----------------------------------------------------------------------------
# Element routines (scalar, one element) -- line-by-line restatements
# ----------------------------------------------------------------------------


def truss1d_linear_element(x_i0, x_j0, u_i, u_j, young, area, dtype=np.float64):
    """fem/element.py:15-42 / fem/nn_assembly.py:18-47."""
    t = dtype
    l0 = abs(t(x_j0) - t(x_i0))
    if l0 <= 0.0:
        raise ValueError("Element with zero initial length detected")
    eps = (t(u_j) - t(u_i)) / l0
    k = (t(young) * t(area)) / l0
    ke = k * np.array([[1.0, -1.0], [-1.0, 1.0]], dtype=t)
    fe = k * np.array([t(u_i) - t(u_j), t(u_j) - t(u_i)], dtype=t)
    return ke, fe, eps


def truss2d_linear_element(x_i0, x_j0, u_i, u_j, young, area, dtype=np.float64):
    """fem/element.py:45-102 / fem/nn_assembly.py:50-102.

    Geometry (l0, cx, cy and their products) is always evaluated in float64 --
    the torch path computes it with Python floats (nn_assembly.py:65-82) and
    only the final pattern is cast to float32.
    """
    t = dtype
    dx0 = np.asarray(x_j0, dtype=np.float64) - np.asarray(x_i0, dtype=np.float64)
    l0 = float(np.linalg.norm(dx0))
    if l0 <= 0.0:
        raise ValueError("Element with zero initial length detected")
    cx = dx0[0] / l0
    cy = dx0[1] / l0
    c2, s2, cs = cx * cx, cy * cy, cx * cy
    pattern = np.array(
        [[c2, cs, -c2, -cs], [cs, s2, -cs, -s2], [-c2, -cs, c2, cs], [-cs, -s2, cs, s2]],
        dtype=t,
    )
    k = (t(young) * t(area)) / t(l0)
    ke = k * pattern
    ue = np.array([u_i[0], u_i[1], u_j[0], u_j[1]], dtype=t)
    fe = ke @ ue
    du = np.asarray(u_j, dtype=np.float64) - np.asarray(u_i, dtype=np.float64)
    eps = (cx * du[0] + cy * du[1]) / l0
    return ke, fe, eps


def truss2d_element_state(x_i0, x_j0, u_i, u_j, young, area, dtype=np.float64):
    """fem/element.py:105-133 (Green-Lagrange)."""
    t = dtype
    x_i0 = np.asarray(x_i0, dtype=t)
    x_j0 = np.asarray(x_j0, dtype=t)
    dx0 = x_j0 - x_i0
    l0 = t(np.linalg.norm(dx0))
    if l0 <= 0.0:
        raise ValueError("Element with zero initial length detected")
    dx = (x_j0 + np.asarray(u_j, dtype=t)) - (x_i0 + np.asarray(u_i, dtype=t))
    l = t(np.linalg.norm(dx))
    d = np.array([dx[0], dx[1], -dx[0], -dx[1]], dtype=t)
    d0 = np.array([dx0[0], dx0[1], -dx0[0], -dx0[1]], dtype=t)
    e_gl = (l * l - l0 * l0) / (t(2.0) * l0 * l0)
    ea = t(young) * t(area)
    ke = (ea / (l0**3)) * np.outer(d0, d0) + (ea / l0) * e_gl * np.outer(d, d)
    fe = (ea / l0) * e_gl * d
    return ke, fe, e_gl


# ----------------------------------------------------------------------------
# Global assembly
# ----------------------------------------------------------------------------


def _per_element(val, nelem, dtype):
    a = np.asarray(val, dtype=dtype)
    if a.ndim == 0:
        a = np.full(nelem, a, dtype=dtype)
    return a


def assemble_system_loop(nodes, elements, young, area, disp, dim=2, kind=LINEAR, dtype=np.float64):
    """fem/assembly.py:16-75 and fem/nn_assembly.py:105-231: the Python loop,
    element by element, dense ``K[ndof,ndof]``, ``f_int[ndof]``, ``max|strain|``.
    ``young``/``area`` may be scalars (the reference's ScalarProperty) or
    per-element arrays (what an NNProperty evaluated at centroids gives)."""
    nodes = np.asarray(nodes, dtype=np.float64)
    el = np.asarray(elements, dtype=np.int64).reshape(-1, 2)
    nnode = nodes.shape[0]
    ndof = nnode * dim
    E = _per_element(young, len(el), dtype)
    A = _per_element(area, len(el), dtype)
    K = np.zeros((ndof, ndof), dtype=dtype)
    f = np.zeros(ndof, dtype=dtype)
    max_eps = 0.0
    for e, (i, j) in enumerate(el):
        dofs = element_dofs(i, j, dim)
        if dim == 1:
            ke, fe, eps = truss1d_linear_element(nodes[i], nodes[j], disp[i], disp[j], E[e], A[e], dtype)
        else:
            u_i = np.array([disp[2 * i], disp[2 * i + 1]])
            u_j = np.array([disp[2 * j], disp[2 * j + 1]])
            fn = truss2d_linear_element if kind == LINEAR else truss2d_element_state
            ke, fe, eps = fn(nodes[i], nodes[j], u_i, u_j, E[e], A[e], dtype)
        K[np.ix_(dofs, dofs)] += ke
        f[dofs] += fe
        max_eps = max(max_eps, abs(float(eps)))
    return K, f, max_eps


@dataclass
class ElementGeometry:
    """Initial-configuration geometry per element (fp64), as the reference
    computes it inside each element call."""

    l0: np.ndarray
    c: np.ndarray  # cos (1-D: sign is irrelevant, c = 1)
    s: np.ndarray
    dx0: np.ndarray  # [nelem, dim]


def element_geometry(nodes, elements, dim=2) -> ElementGeometry:
    nodes = np.asarray(nodes, dtype=np.float64)
    el = np.asarray(elements, dtype=np.int64).reshape(-1, 2)
    if dim == 1:
        dx = (nodes[el[:, 1]] - nodes[el[:, 0]]).reshape(-1, 1)
        l0 = np.abs(dx[:, 0])
        return ElementGeometry(l0, np.ones_like(l0), np.zeros_like(l0), dx)
    dx = nodes[el[:, 1]] - nodes[el[:, 0]]
    l0 = np.sqrt(dx[:, 0] * dx[:, 0] + dx[:, 1] * dx[:, 1])
    return ElementGeometry(l0, dx[:, 0] / l0, dx[:, 1] / l0, dx)


def _bcast(a, like):
    """Give per-element array ``a`` a trailing batch axis if ``like`` has one."""
    a = np.asarray(a)
    if like.ndim == 2 and a.ndim == 1:
        return a[:, None]
    return a


def element_forces(nodes, elements, young, area, disp, dim=2, kind=LINEAR, dtype=np.float64):
    """Vectorised element internal forces ``fe[nelem, 2*dim(,B)]`` and strain.

    Same arithmetic as ``truss*_linear_element`` (``fe = ke @ u_e`` expanded in
    the row order NumPy's matmul uses) / ``truss2d_element_state``."""
    el = np.asarray(elements, dtype=np.int64).reshape(-1, 2)
    g = element_geometry(nodes, el, dim)
    u = np.asarray(disp, dtype=dtype)
    E = _per_element(young, len(el), dtype) if np.ndim(young) < 2 else np.asarray(young, dtype=dtype)
    A = _per_element(area, len(el), dtype) if np.ndim(area) < 2 else np.asarray(area, dtype=dtype)
    if u.ndim == 2:
        E, A = _bcast(E, u), _bcast(A, u)
    l0 = _bcast(g.l0.astype(dtype), u)
    k = (E * A) / l0
    if dim == 1:
        ui, uj = u[el[:, 0]], u[el[:, 1]]
        fe = np.stack([k * (ui - uj), k * (uj - ui)], axis=1)
        eps = (uj - ui) / l0
        return fe, eps
    uix, uiy = u[2 * el[:, 0]], u[2 * el[:, 0] + 1]
    ujx, ujy = u[2 * el[:, 1]], u[2 * el[:, 1] + 1]
    if kind == LINEAR:
        c, s = _bcast(g.c, u), _bcast(g.s, u)
        c2 = (c * c).astype(dtype)
        s2 = (s * s).astype(dtype)
        cs = (c * s).astype(dtype)
        r0 = (k * c2) * uix + (k * cs) * uiy + (k * -c2) * ujx + (k * -cs) * ujy
        r1 = (k * cs) * uix + (k * s2) * uiy + (k * -cs) * ujx + (k * -s2) * ujy
        r2 = (k * -c2) * uix + (k * -cs) * uiy + (k * c2) * ujx + (k * cs) * ujy
        r3 = (k * -cs) * uix + (k * -s2) * uiy + (k * cs) * ujx + (k * s2) * ujy
        fe = np.stack([r0, r1, r2, r3], axis=1)
        eps = (c * (ujx - uix) + s * (ujy - uiy)) / l0
        return fe, eps
    xy = np.asarray(nodes, dtype=dtype)
    xi0, yi0 = _bcast(xy[el[:, 0], 0], u), _bcast(xy[el[:, 0], 1], u)
    xj0, yj0 = _bcast(xy[el[:, 1], 0], u), _bcast(xy[el[:, 1], 1], u)
    dxx = (xj0 + ujx) - (xi0 + uix)  # element.py:119-121
    dxy = (yj0 + ujy) - (yi0 + uiy)
    l = np.sqrt(dxx * dxx + dxy * dxy)
    e_gl = (l * l - l0 * l0) / (dtype(2.0) * l0 * l0)
    n = k * e_gl
    fe = np.stack([n * dxx, n * dxy, -(n * dxx), -(n * dxy)], axis=1)
    return fe, e_gl


def assemble_residual(nodes, elements, young, area, disp, dim=2, kind=LINEAR, dtype=np.float64):
    """``f_int`` only, vectorised; accumulation in ascending element order per
    DOF exactly like fem/assembly.py:72 (``np.add.at`` is ordered)."""
    el = np.asarray(elements, dtype=np.int64).reshape(-1, 2)
    fe, eps = element_forces(nodes, el, young, area, disp, dim, kind, dtype)
    u = np.asarray(disp)
    f = np.zeros(u.shape, dtype=dtype)
    dofs = all_element_dofs(el, dim)
    np.add.at(f, dofs.reshape(-1), fe.reshape((-1,) + fe.shape[2:]))
    max_eps = np.max(np.abs(eps), axis=0) if len(el) else 0.0
    return f, max_eps


def element_stiffness(nodes, elements, young, area, disp=None, dim=2, kind=LINEAR, dtype=np.float64):
    """Vectorised ``ke[nelem, 2*dim, 2*dim]`` for one (un-batched) problem."""
    el = np.asarray(elements, dtype=np.int64).reshape(-1, 2)
    g = element_geometry(nodes, el, dim)
    E = _per_element(young, len(el), dtype)
    A = _per_element(area, len(el), dtype)
    l0 = g.l0.astype(dtype)
    k = (E * A) / l0
    if dim == 1:
        pat = np.array([[1.0, -1.0], [-1.0, 1.0]], dtype=dtype)
        return k[:, None, None] * pat[None]
    if kind == LINEAR:
        d = np.stack([g.c, g.s, -g.c, -g.s], axis=1)
        pat = (d[:, :, None] * d[:, None, :]).astype(dtype)
        # the reference builds c2, cs, s2 then negates; products of +/-c, +/-s
        # give the same bits.
        return k[:, None, None] * pat
    u = np.asarray(disp, dtype=dtype)
    xy = np.asarray(nodes, dtype=dtype)
    ui = np.stack([u[2 * el[:, 0]], u[2 * el[:, 0] + 1]], axis=1)
    uj = np.stack([u[2 * el[:, 1]], u[2 * el[:, 1] + 1]], axis=1)
    dx0 = g.dx0.astype(dtype)
    dx = (xy[el[:, 1]] + uj) - (xy[el[:, 0]] + ui)
    d = np.concatenate([dx, -dx], axis=1)
    d0 = np.concatenate([dx0, -dx0], axis=1)
    l = np.sqrt(dx[:, 0] * dx[:, 0] + dx[:, 1] * dx[:, 1])
    e_gl = (l * l - l0 * l0) / (dtype(2.0) * l0 * l0)
    ea = E * A
    ke = (ea / l0**3)[:, None, None] * (d0[:, :, None] * d0[:, None, :])
    ke = ke + ((ea / l0) * e_gl)[:, None, None] * (d[:, :, None] * d[:, None, :])
    return ke


def assemble_tangent_bsr(nodes, elements, young, area, disp=None, dim=2, kind=LINEAR, dtype=np.float64):
    """Tangent in node-block CSR: ``vals[nnzb, dim, dim]``; per block the
    element contributions are added in ascending element order (same order the
    dense ``+=`` of fem/assembly.py:71 produces for each entry)."""
    el = np.asarray(elements, dtype=np.int64).reshape(-1, 2)
    nnode = np.asarray(nodes).shape[0]
    rowptr, colind, slots = bsr_pattern(nnode, el)
    ke = element_stiffness(nodes, el, young, area, disp, dim, kind, dtype)
    vals = np.zeros((len(colind), dim, dim), dtype=dtype)
    blocks = np.stack(
        [ke[:, :dim, :dim], ke[:, :dim, dim:], ke[:, dim:, :dim], ke[:, dim:, dim:]], axis=1
    )  # [nelem, 4, dim, dim]
    np.add.at(vals, slots.reshape(-1), blocks.reshape(-1, dim, dim))
    return rowptr, colind, vals


def bsr_to_dense(rowptr, colind, vals, dim=2):
    nnode = len(rowptr) - 1
    K = np.zeros((nnode * dim, nnode * dim), dtype=vals.dtype)
    for n in range(nnode):
        for p in range(rowptr[n], rowptr[n + 1]):
            m = colind[p]
            K[n * dim : (n + 1) * dim, m * dim : (m + 1) * dim] = vals[p]
    return K


def tangent_matvec(nodes, elements, young, area, v, disp=None, dim=2, kind=LINEAR, dtype=np.float64):
    """``K_t(u) @ v`` matrix-free (un-batched or batched ``v[ndof,B]``)."""
    el = np.asarray(elements, dtype=np.int64).reshape(-1, 2)
    v = np.asarray(v, dtype=dtype)
    if kind == LINEAR or dim == 1:
        return assemble_residual(nodes, el, young, area, v, dim, LINEAR, dtype)[0]
    out = np.zeros_like(v)
    if v.ndim == 1:
        ke = element_stiffness(nodes, el, young, area, disp, dim, kind, dtype)
        dofs = all_element_dofs(el, dim)
        np.add.at(out, dofs.reshape(-1), np.einsum("eab,eb->ea", ke, v[dofs]).reshape(-1))
        return out
    for b in range(v.shape[1]):
        Eb = young[:, b] if np.ndim(young) == 2 else young
        Ab = area[:, b] if np.ndim(area) == 2 else area
        out[:, b] = tangent_matvec(nodes, el, Eb, Ab, v[:, b], disp[:, b], dim, kind, dtype)
    return out


def material_vjp(nodes, elements, young, area, disp, g, dim=2, kind=LINEAR, dtype=np.float64):
    """Given ``g = dL/df_int`` return ``(dL/dE_e, dL/dA_e)`` per element: what
    torch autograd produces through fem/nn_assembly.py:72,:96-100,:226-229.
    ``f_e = (E*A) * h_e(u)`` so ``dL/dE = A * <g_e, h_e>`` and vice versa."""
    el = np.asarray(elements, dtype=np.int64).reshape(-1, 2)
    u = np.asarray(disp, dtype=dtype)
    one = np.ones(len(el), dtype=dtype)
    h, _ = element_forces(nodes, el, one, one, u, dim, kind, dtype)  # unit-EA forces
    gd = np.asarray(g, dtype=dtype)[all_element_dofs(el, dim)]  # [nelem, 2dim(,B)]
    gh = np.sum(gd * h, axis=1)
    E = _per_element(young, len(el), dtype) if np.ndim(young) < 2 else np.asarray(young, dtype=dtype)
    A = _per_element(area, len(el), dtype) if np.ndim(area) < 2 else np.asarray(area, dtype=dtype)
    if u.ndim == 2:
        E, A = _bcast(E, u), _bcast(A, u)
    return A * gh, E * gh


# ----------------------------------------------------------------------------
# Material networks (SimpleNN + NNProperty)
# ----------------------------------------------------------------------------


@dataclass(frozen=True)
class NetSpec:
    """examples/json/generic.py:118-142: Linear(in,h) Tanh [Linear(h,h) Tanh]x(L-1) Linear(h,1)."""

    input_dim: int
    hidden_layers: int
    width: int

    @property
    def layer_shapes(self) -> List[Tuple[int, int]]:
        shapes = [(self.width, self.input_dim)]
        shapes += [(self.width, self.width)] * (self.hidden_layers - 1)
        shapes.append((1, self.width))
        return shapes

    @property
    def n_params(self) -> int:
        return sum(o * i + o for o, i in self.layer_shapes)


def unpack_theta(spec: NetSpec, theta: np.ndarray):
    """Flat theta in ``nn.Module.parameters()`` order: W1,b1,W2,b2,...,Wout,bout
    (row-major ``[out,in]`` weights)."""
    out, off = [], 0
    for o, i in spec.layer_shapes:
        W = theta[off : off + o * i].reshape(o, i)
        off += o * i
        b = theta[off : off + o]
        off += o
        out.append((W, b))
    return out


def nn_inputs(nodes, elements, load_factor, dim=2, dtype=np.float64) -> np.ndarray:
    """Centroid inputs ``[load_factor, x(, y)]`` -- the dict keys are sorted
    (fem/properties.py:116-125), centroids from fem/nn_assembly.py:198-205."""
    nodes = np.asarray(nodes, dtype=np.float64)
    el = np.asarray(elements, dtype=np.int64).reshape(-1, 2)
    cen = (nodes[el[:, 0]] + nodes[el[:, 1]]) / 2.0
    cen = cen.reshape(len(el), -1)
    lam = np.full((len(el), 1), float(load_factor))
    return np.concatenate([lam, cen], axis=1).astype(dtype)


def softplus(x):
    """torch.nn.functional.softplus (beta=1, threshold=20), fem/properties.py:154."""
    x = np.asarray(x)
    safe = np.minimum(x, 20.0)
    return np.where(x > 20.0, x, np.log1p(np.exp(safe)))


def mlp_forward(spec: NetSpec, theta, X, scale=1.0, enforce_positive=True, keep=False):
    """fem/properties.py:150-156: ``softplus(net(x)) * scale``; X is [n, in]."""
    layers = unpack_theta(spec, np.asarray(theta))
    a = np.asarray(X, dtype=np.asarray(theta).dtype)
    acts = [a]
    for W, b in layers[:-1]:
        a = np.tanh(a @ W.T + b)
        acts.append(a)
    W, b = layers[-1]
    z = (a @ W.T + b)[:, 0]
    y = softplus(z) if enforce_positive else z
    out = y * np.asarray(theta).dtype.type(scale)
    if keep:
        return out, (acts, z)
    return out


def mlp_backward(spec: NetSpec, theta, X, g_out, scale=1.0, enforce_positive=True):
    """``dL/dtheta`` (flat) given ``g_out = dL/d(value)`` per point: the reverse
    pass torch autograd runs through NNProperty.value."""
    theta = np.asarray(theta)
    t = theta.dtype.type
    layers = unpack_theta(spec, theta)
    _, (acts, z) = mlp_forward(spec, theta, X, scale, enforce_positive, keep=True)
    gz = np.asarray(g_out, dtype=theta.dtype) * t(scale)
    if enforce_positive:
        sig = np.where(z > 20.0, 1.0, 1.0 / (1.0 + np.exp(-np.minimum(z, 20.0))))
        gz = gz * sig.astype(theta.dtype)
    grads = [None] * len(layers)
    delta = gz[:, None]  # [n, 1]
    for li in range(len(layers) - 1, -1, -1):
        W, _ = layers[li]
        a_in = acts[li]
        grads[li] = (delta.T @ a_in, delta.sum(axis=0))
        if li > 0:
            delta = (delta @ W) * (1.0 - a_in * a_in)
    flat = []
    for gW, gb in grads:
        flat.append(gW.reshape(-1))
        flat.append(gb.reshape(-1))
    return np.concatenate(flat).astype(theta.dtype)


# ----------------------------------------------------------------------------
# Newton-Raphson drivers
# ----------------------------------------------------------------------------


@dataclass
class Mesh:
    nodes: np.ndarray
    elements: np.ndarray
    loads: np.ndarray
    fixed_dofs: np.ndarray
    dim: int = 2

    @property
    def nnode(self):
        return np.asarray(self.nodes).shape[0]

    @property
    def ndof(self):
        return self.nnode * self.dim


def _nr_increment(mesh: Mesh, E, A, u, f_ext, max_iterations, tolerance, min_den, kind):
    free, fixed = free_and_fixed_dofs(mesh.ndof, mesh.fixed_dofs)
    converged, res, eps, n_it = False, np.inf, 0.0, 0
    for ite in range(max_iterations):
        K, f_int, eps = assemble_system_loop(mesh.nodes, mesh.elements, E, A, u, mesh.dim, kind)
        rhs = f_ext - f_int
        try:
            du_f = np.linalg.solve(K[np.ix_(free, free)], rhs[free])
        except np.linalg.LinAlgError as exc:
            raise RuntimeError("Tangent stiffness became singular during solve") from exc
        du = np.zeros_like(u)
        du[free] = du_f
        u += du
        res = np.linalg.norm(du) / max(np.linalg.norm(u), min_den)
        n_it = ite + 1
        if res <= tolerance:
            converged = True
            break
    return converged, res, eps, n_it


def solve_incremental_newton(mesh: Mesh, E, A, n_increments=10, max_iterations=80, tolerance=1e-6,
                             min_denominator=1e-12, kind=LINEAR):
    """fem/core.py:10-79: true incremental NR, ``u`` carried across increments."""
    u = np.zeros(mesh.ndof)
    free, _ = free_and_fixed_dofs(mesh.ndof, mesh.fixed_dofs)
    history, ok_all = [], True
    for iinc in range(1, n_increments + 1):
        lam = iinc / n_increments
        ok, res, eps, n_it = _nr_increment(mesh, E, A, u, lam * mesh.loads, max_iterations, tolerance,
                                           min_denominator, kind)
        history.append({"increment": float(iinc), "load_factor": float(lam), "iterations": float(n_it),
                        "residual": float(res), "max_strain": float(eps), "converged": float(ok)})
        ok_all = ok_all and ok
    K, _, _ = assemble_system_loop(mesh.nodes, mesh.elements, E, A, u, mesh.dim, kind)
    reactions = K @ u - mesh.loads
    reactions[free] = 0.0
    return u, reactions, ok_all, history


def solve_nr(mesh: Mesh, E, A, load_factor=1.0, max_iterations=1000, tolerance=1e-6,
             min_denominator=1e-10, kind=LINEAR):
    """fem/solver.py:408-512: one load factor, always restarts from u = 0."""
    u = np.zeros(mesh.ndof)
    free, _ = free_and_fixed_dofs(mesh.ndof, mesh.fixed_dofs)
    ok, res, eps, n_it = _nr_increment(mesh, E, A, u, load_factor * mesh.loads, max_iterations, tolerance,
                                       min_denominator, kind)
    history = [{"load_factor": float(load_factor), "iterations": float(n_it), "residual": float(res),
                "max_strain": float(eps), "converged": float(ok)}]
    K, _, _ = assemble_system_loop(mesh.nodes, mesh.elements, E, A, u, mesh.dim, kind)
    reactions = K @ u - load_factor * mesh.loads
    reactions[free] = 0.0
    return u, reactions, ok, history


# ----------------------------------------------------------------------------
# PINN gradient descent (solve_gd inner loop, fem/solver.py:200-400)
# ----------------------------------------------------------------------------


@dataclass
class MaterialNets:
    """The three material properties: each either a scalar or (NetSpec, theta, scale).
    Order young, area, density == Material.get_all_torch_params (fem/model.py:36-43)."""

    young: object
    area: object
    density: object = 0.0

    def nets(self):
        return [(n, p) for n, p in (("young", self.young), ("area", self.area), ("density", self.density))
                if isinstance(p, tuple)]

    def evaluate(self, name, X):
        p = getattr(self, name)
        if isinstance(p, tuple):
            spec, theta, scale = p
            return mlp_forward(spec, theta, X, scale)
        return np.full(len(X), float(p), dtype=X.dtype)


class Adam:
    """torch.optim.Adam defaults (betas 0.9/0.999, eps 1e-8, no decay/amsgrad),
    single-tensor update order of torch/optim/adam.py (torch 2.11)."""

    def __init__(self, lr, n, dtype=np.float64):
        self.lr, self.t = lr, 0
        self.m = np.zeros(n, dtype=dtype)
        self.v = np.zeros(n, dtype=dtype)

    def step(self, p, g):
        self.t += 1
        b1, b2, eps = 0.9, 0.999, 1e-8
        self.m += (g - self.m) * (1.0 - b1)  # lerp_
        self.v *= b2
        self.v += (1.0 - b2) * g * g
        bc1 = 1.0 - b1**self.t
        bc2 = 1.0 - b2**self.t
        step_size = self.lr / bc1
        denom = np.sqrt(self.v) / math.sqrt(bc2) + eps
        p += (-step_size * self.m) / denom  # addcdiv_(exp_avg, denom, value=-step_size)


def gd_loss_and_grads(mesh: Mesh, mat: MaterialNets, u, lam, meas_dofs, meas_vals, alpha_p, alpha_d,
                      kind=LINEAR, dtype=np.float64):
    """One forward/backward of fem/solver.py:262-289 in closed form.

    Returns dict with losses, residual, ``g_u`` and per-net ``g_theta``."""
    free, _ = free_and_fixed_dofs(mesh.ndof, mesh.fixed_dofs)
    X = nn_inputs(mesh.nodes, mesh.elements, lam, mesh.dim, dtype)
    E = mat.evaluate("young", X)
    A = mat.evaluate("area", X)
    f_int, _ = assemble_residual(mesh.nodes, mesh.elements, E, A, u, mesh.dim, kind, dtype)
    r = f_int[free] - dtype(lam) * mesh.loads.astype(dtype)[free]
    loss_p = dtype(0.5) * np.sum(r * r)
    has_meas = meas_dofs is not None and alpha_d > 0
    if has_meas:
        rd = np.asarray(meas_vals, dtype=dtype) - u[meas_dofs]
        loss_d = np.mean(rd * rd)
        loss = dtype(alpha_p) * loss_p + dtype(alpha_d) * loss_d
    else:
        loss_d = dtype(0.0)
        loss = dtype(alpha_p) * loss_p
    g_f = np.zeros(mesh.ndof, dtype=dtype)
    g_f[free] = dtype(alpha_p) * r
    g_u = tangent_matvec(mesh.nodes, mesh.elements, E, A, g_f, u, mesh.dim, kind, dtype)
    if has_meas:
        np.add.at(g_u, np.asarray(meas_dofs), -dtype(2.0 * alpha_d / len(meas_dofs)) * rd)
    gE, gA = material_vjp(mesh.nodes, mesh.elements, E, A, u, g_f, mesh.dim, kind, dtype)
    g_theta = {}
    for name, g_e in (("young", gE), ("area", gA)):
        p = getattr(mat, name)
        if isinstance(p, tuple):
            spec, theta, scale = p
            g_theta[name] = mlp_backward(spec, theta, X, g_e, scale)
    return {"loss": loss, "loss_physics": loss_p, "loss_data": loss_d, "residual": r, "g_u": g_u,
            "g_theta": g_theta, "E": E, "A": A, "f_int": f_int}


def solve_gd(mesh: Mesh, mat: MaterialNets, max_iterations=1000, tolerance=1e-6, lr_u=1e-7, lr_theta=1e-4,
             alpha_p=1.0, alpha_d=100.0, meas_dofs=None, meas_vals=None, lam=1.0, u_initial=None,
             kind=LINEAR, dtype=np.float64):
    """fem/solver.py:200-400 without the preconditioning wrapper.  Mutates the
    ``theta`` arrays inside ``mat`` (the reference mutates the caller's model)."""
    free, fixed = free_and_fixed_dofs(mesh.ndof, mesh.fixed_dofs)
    u = np.zeros(mesh.ndof, dtype=dtype) if u_initial is None else np.array(u_initial, dtype=dtype)
    opt_u = Adam(lr_u, mesh.ndof, dtype)
    opt_t = {name: Adam(lr_theta, p[0].n_params, dtype) for name, p in mat.nets() if name != "density"}
    # one torch optimiser over all tensors shares the step counter; per-net Adam
    # objects stepped together are equivalent.  Density has grad=None -> skipped.
    history, converged = [], False
    has_meas = meas_dofs is not None and meas_vals is not None
    for it in range(max_iterations):
        out = gd_loss_and_grads(mesh, mat, u, lam, meas_dofs if has_meas else None, meas_vals, alpha_p,
                                alpha_d, kind, dtype)
        opt_u.step(u, out["g_u"])
        for name, g in out["g_theta"].items():
            opt_t[name].step(getattr(mat, name)[1], g)
        u[fixed] = 0.0
        res_norm = float(np.sqrt(np.sum(out["residual"] ** 2)))
        entry = {"iteration": float(it + 1), "loss_total": float(out["loss"]),
                 "loss_physics": float(out["loss_physics"]),
                 "loss_data": float(out["loss_data"]) if has_meas else 0.0,
                 "u_norm": float(np.sqrt(np.sum(u[free] ** 2))), "residual_norm": res_norm}
        if mat.nets():
            tn = 0.0
            for _, (spec, theta, _s) in mat.nets():
                for W, b in unpack_theta(spec, theta):
                    tn += float(np.sqrt(np.sum(W * W))) + float(np.sqrt(np.sum(b * b)))
            entry["theta_norm"] = tn
        history.append(entry)
        if it > 10:
            if res_norm < tolerance:
                converged = True
                break
            if not np.isnan(entry["loss_total"]) and entry["loss_total"] < tolerance:
                converged = True
                break
    K_unused = None
    X = nn_inputs(mesh.nodes, mesh.elements, lam, mesh.dim, dtype)
    f_int, _ = assemble_residual(mesh.nodes, mesh.elements, mat.evaluate("young", X), mat.evaluate("area", X),
                                 u, mesh.dim, kind, dtype)
    reactions = f_int - dtype(lam) * mesh.loads.astype(dtype)
    reactions[free] = 0.0
    return u, reactions, converged, history


def solve_gd_preconditioned(mesh, mat, max_iterations, tolerance, **kw):
    """fem/solver.py:114-198: relaxed phase (min(300, max//3) its, tol
    max(1e-4, 10 tol)), then the main phase warm-started through float32."""
    pre_max = min(300, max_iterations // 3)
    pre_tol = max(1e-4, tolerance * 10)
    u, reac, conv, hist = solve_gd(mesh, mat, pre_max, pre_tol, **kw)
    if conv and hist[-1]["residual_norm"] < tolerance:
        return u, reac, conv, hist
    kw = dict(kw)
    kw["u_initial"] = u.astype(np.float32).astype(u.dtype)  # solver.py:147-149
    u2, reac2, conv2, hist2 = solve_gd(mesh, mat, max_iterations - pre_max, tolerance, **kw)
    n_pre = hist[-1]["iteration"] if hist else 0
    merged = list(hist) + [dict(h, iteration=h["iteration"] + n_pre) for h in hist2]
    return u2, reac2, conv2, merged


def solve_incremental_gd(mesh, mat, n_increments=10, preconditioning=False, max_iterations=1000,
                         tolerance=1e-6, **kw):
    """fem/solver.py:1045-1167 with method "gd": load factor schedule, warm
    start carried through float32 (solver.py:1110), stop at first failure."""
    u_cur, result = None, None
    for iinc in range(1, n_increments + 1):
        lam = 0.0 + (iinc / n_increments) * (1.0 - 0.0)
        args = dict(kw, lam=lam, u_initial=None if u_cur is None else u_cur.astype(np.float32).astype(np.float64))
        fn = solve_gd_preconditioned if preconditioning else solve_gd
        result = fn(mesh, mat, max_iterations, tolerance, **args)
        u_cur = result[0]
        if not result[2]:
            break
    return result


# ----------------------------------------------------------------------------
# Gauss-Newton / Levenberg-Marquardt (fem/nn_solver.py)
# ----------------------------------------------------------------------------


def jacobian_blocks(mesh: Mesh, mat: MaterialNets, u, f_ext, meas_dofs=None, lam=1.0, kind=LINEAR,
                    dtype=np.float64):
    """fem/nn_solver.py:50-135 in closed form.

    ``J_utheta[i, :] = d f_int[free_i] / d theta`` -- the reference obtains each
    row with a reverse pass (``r_physics[i].backward``); analytically
    ``d f_int / d theta = sum_e (d f_e / dE_e)(dE_e/dtheta) + (.. A ..)``, so row
    ``i`` is an MLP backward with upstream ``g_e = A_e * h_e[local(i)]``.
    Columns follow ``get_all_torch_params`` order: young | area | density(zeros)."""
    free, _ = free_and_fixed_dofs(mesh.ndof, mesh.fixed_dofs)
    X = nn_inputs(mesh.nodes, mesh.elements, lam, mesh.dim, dtype)
    E = mat.evaluate("young", X)
    A = mat.evaluate("area", X)
    rowptr, colind, vals = assemble_tangent_bsr(mesh.nodes, mesh.elements, E, A, u, mesh.dim, kind, dtype)
    K = bsr_to_dense(rowptr, colind, vals, mesh.dim)
    f_int, _ = assemble_residual(mesh.nodes, mesh.elements, E, A, u, mesh.dim, kind, dtype)
    r = f_int[free] - np.asarray(f_ext, dtype=dtype)[free]
    j_uu = K[np.ix_(free, free)]
    cols = []
    for name in ("young", "area", "density"):
        p = getattr(mat, name)
        if not isinstance(p, tuple):
            continue
        spec, theta, scale = p
        blk = np.zeros((len(free), spec.n_params), dtype=dtype)
        if name != "density":
            for row, dof in enumerate(free):
                g_f = np.zeros(mesh.ndof, dtype=dtype)
                g_f[dof] = 1.0
                gE, gA = material_vjp(mesh.nodes, mesh.elements, E, A, u, g_f, mesh.dim, kind, dtype)
                blk[row] = mlp_backward(spec, theta, X, gE if name == "young" else gA, scale)
        cols.append(blk)
    n_theta = sum(p[0].n_params for _, p in mat.nets())
    j_ut = np.concatenate(cols, axis=1) if cols else np.zeros((len(free), n_theta), dtype=dtype)
    j_du = None
    if meas_dofs is not None:
        j_du = np.zeros((len(meas_dofs), len(free)), dtype=dtype)
        for i, d in enumerate(meas_dofs):
            hit = np.where(free == d)[0]
            if len(hit):
                j_du[i, hit[0]] = -1.0
    return j_uu, j_ut, r, j_du


def gauss_newton_system(j_uu, j_ut, r_phys, j_du, r_data, alpha_p=1.0, alpha_d=1.0):
    """fem/nn_solver.py:223-248: stack J and R with the loss weights."""
    if j_du is not None:
        top = np.concatenate([alpha_p * j_uu, alpha_p * j_ut], axis=1)
        bot = np.concatenate([alpha_d * j_du, np.zeros((j_du.shape[0], j_ut.shape[1]), dtype=j_uu.dtype)], axis=1)
        return np.concatenate([top, bot], axis=0), np.concatenate([alpha_p * r_phys, alpha_d * r_data])
    return np.concatenate([j_uu, j_ut], axis=1), alpha_p * r_phys


def lm_step(J, R):
    """fem/nn_solver.py:266-277: ``(J^T J + 1e-6 tr(J^T J)/n I) dx = -J^T R``."""
    jtj = J.T @ J
    jtr = J.T @ R
    damping = 1e-6 * np.trace(jtj) / jtj.shape[0]
    reg = jtj + damping * np.eye(jtj.shape[0], dtype=J.dtype)
    return np.linalg.solve(reg, -jtr), jtj, jtr, damping


def lm_step_dual(J, R):
    """The same step through the m x m dual system: ``(J^T J + d I)^-1 J^T = J^T (J J^T + d I)^-1`` with the
    reference's damping ``d = 1e-6 tr(J^T J)/n`` (= 1e-6 tr(J J^T)/n).  Identical in exact arithmetic; the n x n
    matrix is a rank-m matrix plus a 1e-6-relative ridge (condition ~1e6 n/m and worse), so the two agree only as
    far as that conditioning allows."""
    g = J @ J.T
    damping = 1e-6 * np.trace(g) / J.shape[1]
    y = np.linalg.solve(g + damping * np.eye(g.shape[0], dtype=J.dtype), R)
    return -(J.T @ y), damping


def _theta_add(mat: MaterialNets, delta, step):
    off = 0
    for _, (spec, theta, _s) in mat.nets():
        theta += step * delta[off : off + spec.n_params]
        off += spec.n_params


def solve_pinn_newton_raphson(mesh: Mesh, mat: MaterialNets, f_ext, meas_vals=None, meas_dofs=None,
                              max_iterations=50, tolerance=1e-6, alpha_p=1.0, alpha_d=1.0,
                              min_denominator=1e-12, line_search=True, kind=LINEAR, dtype=np.float64, dual=False):
    """fem/nn_solver.py:138-426 including its quirks: on an accepted trial
    theta keeps the trial update and is advanced again (:309-313 + :366-371).  ``dual``: take the LM step through
    :func:`lm_step_dual` (what the CUDA path does when there are fewer residuals than unknowns)."""
    free, fixed = free_and_fixed_dofs(mesh.ndof, mesh.fixed_dofs)
    u = np.zeros(mesh.ndof, dtype=dtype)
    f_ext = np.asarray(f_ext, dtype=dtype)
    has_meas = meas_vals is not None and meas_dofs is not None
    md = np.asarray(meas_dofs, dtype=np.int64) if has_meas else None
    mv = np.asarray(meas_vals, dtype=dtype) if has_meas else None
    history, converged = [], False
    n_free = len(free)

    def total_residual(uu):
        X = nn_inputs(mesh.nodes, mesh.elements, 1.0, mesh.dim, dtype)
        f, _ = assemble_residual(mesh.nodes, mesh.elements, mat.evaluate("young", X), mat.evaluate("area", X),
                                 uu, mesh.dim, kind, dtype)
        rp = f[free] - f_ext[free]
        if has_meas:
            return np.concatenate([alpha_p * rp, alpha_d * (mv - uu[md])])
        return alpha_p * rp

    for it in range(max_iterations):
        j_uu, j_ut, r_p, j_du = jacobian_blocks(mesh, mat, u, f_ext, md, 1.0, kind, dtype)
        r_d = (mv - u[md]) if has_meas else np.zeros(0, dtype=dtype)
        J, R = gauss_newton_system(j_uu, j_ut, r_p, j_du if has_meas else None, r_d, alpha_p, alpha_d)
        rp_n, rd_n, rt_n = (float(np.linalg.norm(r_p)), float(np.linalg.norm(r_d)) if has_meas else 0.0,
                            float(np.linalg.norm(R)))
        try:
            dx = lm_step_dual(J, R)[0] if dual else lm_step(J, R)[0]
        except np.linalg.LinAlgError:
            break
        du_f, dth = dx[:n_free], dx[n_free:]
        step = 1.0
        if line_search:
            for _ in range(15):
                u_t = u.copy()
                u_t[free] += step * du_f
                u_t[fixed] = 0.0
                backup = [theta.copy() for _, (_sp, theta, _s) in mat.nets()]
                _theta_add(mat, dth, step)
                if float(np.linalg.norm(total_residual(u_t))) < rt_n * (1.0 - 1e-4 * step):
                    break
                for (_, (_sp, theta, _s)), b in zip(mat.nets(), backup):
                    theta[:] = b
                step *= 0.7
                if step < 1e-10:
                    step = 0.0
                    break
            if 0 < step < 1e-8:
                step = 1e-6
        if step > 0:
            u[free] += step * du_f
            u[fixed] = 0.0
            _theta_add(mat, dth, step)
        rel = rt_n / max(float(np.linalg.norm(u[free])), min_denominator)
        history.append({"iteration": float(it + 1), "r_physics": rp_n, "r_data": rd_n, "r_total": rt_n,
                        "relative_error": rel, "step_size": float(step)})
        if rel < tolerance and step > 0:
            converged = True
            break
        if step == 0.0:
            break
    return u, converged, history


# ----------------------------------------------------------------------------
# Synthetic meshes
# ----------------------------------------------------------------------------


def lattice_truss(nx: int, ny: Optional[int] = None) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """C5 generator: nx x ny nodes at integer coordinates, node id = j*nx + i.
    Members, row-major over cells, in this order: all horizontals
    (i,j)-(i+1,j), then all verticals (i,j)-(i,j+1), then one diagonal
    (i,j)-(i+1,j+1) per cell.  Left column (i == 0) fully fixed.
    nx = ny = 578 gives nnode 334084, nelem 999941, ndof 668168."""
    ny = nx if ny is None else ny
    jj, ii = np.meshgrid(np.arange(ny), np.arange(nx), indexing="ij")
    nid = (jj * nx + ii).astype(np.int64)
    nodes = np.stack([ii.reshape(-1), jj.reshape(-1)], axis=1).astype(np.float64)
    hor = np.stack([nid[:, :-1].reshape(-1), nid[:, 1:].reshape(-1)], axis=1)
    ver = np.stack([nid[:-1, :].reshape(-1), nid[1:, :].reshape(-1)], axis=1)
    dia = np.stack([nid[:-1, :-1].reshape(-1), nid[1:, 1:].reshape(-1)], axis=1)
    elements = np.concatenate([hor, ver, dia], axis=0)
    left = nid[:, 0]
    fixed = np.sort(np.concatenate([2 * left, 2 * left + 1]))
    return nodes, elements, fixed
