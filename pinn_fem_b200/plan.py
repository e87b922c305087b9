"""AssemblyPlan: Python handle on a ``pf_plan`` (include/pinnfem.h).

Holds the per-mesh integer work (DOF maps, free/fixed partition, node-block
CSR pattern, node->element incidence) and launches the fp64 assembly kernels
on torch CUDA tensors.  Tensors are passed by ``data_ptr()`` on torch's current
stream; layouts are ``[rows]`` or ``[rows, B]`` (problem index last, contiguous).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import ELEM_GREEN_LAGRANGE, ELEM_LINEAR, check

KINDS = {"linear": ELEM_LINEAR, "green_lagrange": ELEM_GREEN_LAGRANGE, "gl": ELEM_GREEN_LAGRANGE,
         ELEM_LINEAR: ELEM_LINEAR, ELEM_GREEN_LAGRANGE: ELEM_GREEN_LAGRANGE}


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream_ptr(device: torch.device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class AssemblyPlan:
    def __init__(self, nodes, elements, fixed_dofs=(), dim: Optional[int] = None, device=None):
        self._lib = _lib.load()
        self._handle = C.c_void_p()
        nodes = np.ascontiguousarray(np.asarray(nodes, dtype=np.float64))
        if dim is None:
            dim = 1 if nodes.ndim == 1 else int(nodes.shape[1])
        elements = np.ascontiguousarray(np.asarray(elements, dtype=np.int64).reshape(-1, 2))
        fixed = np.ascontiguousarray(np.asarray(fixed_dofs, dtype=np.int64).reshape(-1))
        nnode = nodes.shape[0]
        check(self._lib.pf_plan_create(int(dim), nnode, elements.shape[0], elements.ctypes.data_as(C.c_void_p),
                                       nodes.ctypes.data_as(C.c_void_p), fixed.ctypes.data_as(C.c_void_p),
                                       fixed.size, C.byref(self._handle)))
        self.device: Optional[torch.device] = None
        self._cache = {}
        if device is not None:
            self.to(device)

    # -- lifetime ---------------------------------------------------------
    def __del__(self):
        h = getattr(self, "_handle", None)
        if h is not None and h.value:
            self._lib.pf_plan_destroy(h)
            self._handle = C.c_void_p()

    def to(self, device) -> "AssemblyPlan":
        device = torch.device(device)
        if device.type != "cuda":
            raise ValueError("AssemblyPlan computes on CUDA devices only (no CPU fallback)")
        if self.device is not None:
            if self.device == device:
                return self
            raise ValueError(f"plan already lives on {self.device}")
        if device.index is not None:
            index = device.index
        else:
            index = torch.cuda.current_device() if torch.cuda.is_available() else 0
        check(self._lib.pf_plan_upload(self._handle, index))
        self.device = torch.device("cuda", index)
        return self

    # -- sizes and index arrays --------------------------------------------
    def _size(self, what: int) -> int:
        return int(self._lib.pf_plan_size(self._handle, what))

    dim = property(lambda s: s._size(_lib.PLAN_DIM))
    nnode = property(lambda s: s._size(_lib.PLAN_NNODE))
    nelem = property(lambda s: s._size(_lib.PLAN_NELEM))
    ndof = property(lambda s: s._size(_lib.PLAN_NDOF))
    nfree = property(lambda s: s._size(_lib.PLAN_NFREE))
    nfixed = property(lambda s: s._size(_lib.PLAN_NFIXED))
    nnzb = property(lambda s: s._size(_lib.PLAN_NNZB))
    ninc = property(lambda s: s._size(_lib.PLAN_NINC))
    max_degree = property(lambda s: s._size(_lib.PLAN_MAX_DEGREE))
    has_duplicate_edges = property(lambda s: bool(s._size(_lib.PLAN_HAS_DUPLICATE_EDGES)))

    def _array(self, which: int) -> np.ndarray:
        if which not in self._cache:
            n = int(self._lib.pf_plan_array_len(self._handle, which))
            out = np.empty(n, dtype=np.int64)
            check(self._lib.pf_plan_get_array(self._handle, which, out.ctypes.data_as(C.c_void_p), n))
            out.setflags(write=False)
            self._cache[which] = out
        return self._cache[which]

    elem_dofs = property(lambda s: s._array(_lib.ARR_ELEM_DOFS).reshape(s.nelem, 2 * s.dim))
    free_dofs = property(lambda s: s._array(_lib.ARR_FREE_DOFS))
    fixed_dofs = property(lambda s: s._array(_lib.ARR_FIXED_DOFS))
    bsr_rowptr = property(lambda s: s._array(_lib.ARR_BSR_ROWPTR))
    bsr_colind = property(lambda s: s._array(_lib.ARR_BSR_COLIND))
    elem_slots = property(lambda s: s._array(_lib.ARR_ELEM_SLOTS).reshape(s.nelem, 4))
    inc_ptr = property(lambda s: s._array(_lib.ARR_INC_PTR))
    inc_elem = property(lambda s: s._array(_lib.ARR_INC_ELEM))
    inc_nbr = property(lambda s: s._array(_lib.ARR_INC_NBR))
    inc_slot = property(lambda s: s._array(_lib.ARR_INC_SLOT))
    diag_slot = property(lambda s: s._array(_lib.ARR_DIAG_SLOT))

    def geometry(self, which: str) -> np.ndarray:
        idx = {"l0": 0, "cos": 1, "sin": 2, "centroid": 3}[which]
        n = self.nelem * (self.dim if idx == 3 else 1)
        out = np.empty(n, dtype=np.float64)
        check(self._lib.pf_plan_get_geometry(self._handle, idx, out.ctypes.data_as(C.c_void_p), n))
        return out.reshape(self.nelem, self.dim) if idx == 3 else out

    # -- helpers ------------------------------------------------------------
    def _need_device(self):
        if self.device is None:
            raise _lib.PinnFemError(_lib.PF_ERR_NO_DEVICE,
                                    "plan is not on a CUDA device (use .to('cuda')); there is no CPU fallback")

    def _chk(self, t: torch.Tensor, rows: int, name: str, B: Optional[int] = None) -> int:
        if not isinstance(t, torch.Tensor) or t.device != self.device or t.dtype != torch.float64:
            raise ValueError(f"{name} must be a float64 tensor on {self.device}")
        if not t.is_contiguous():
            raise ValueError(f"{name} must be contiguous")
        if t.dim() not in (1, 2) or t.shape[0] != rows:
            raise ValueError(f"{name} must have shape [{rows}] or [{rows}, B], got {tuple(t.shape)}")
        b = 1 if t.dim() == 1 else int(t.shape[1])
        if B is not None and b != B:
            raise ValueError(f"{name} has batch {b}, expected {B}")
        return b

    def _materials(self, E, A, B) -> int:
        """Return the C-ABI ``mat_batched`` flag: 2-D E/A are per problem, 1-D are shared."""
        bE = self._chk(E, self.nelem, "E")
        bA = self._chk(A, self.nelem, "A")
        if E.dim() != A.dim() or bE != bA:
            raise ValueError("E and A must have the same shape")
        if E.dim() == 2:
            if bE != B:
                raise ValueError(f"E/A batch {bE} does not match B={B}")
            return 1
        return 0

    # -- kernels ------------------------------------------------------------
    def residual(self, u, E, A, f_ext=None, load_factor: float = 1.0, kind="linear", f_int=True, r=False,
                 half_sq=False, max_strain=False):
        """Internal force and (optionally) the masked residual, 0.5*sum(r^2) and
        max|strain| -- one fused launch.  Returns a dict of the requested outputs."""
        self._need_device()
        B = self._chk(u, self.ndof, "u")
        mb = self._materials(E, A, B)
        out = {}
        if f_int:
            out["f_int"] = torch.empty_like(u)
        fext_b = 0
        if r or half_sq:
            if f_ext is None:
                raise ValueError("f_ext is required for the residual")
            fext_b = 1 if (f_ext.dim() == 2 and B > 1) else 0
            self._chk(f_ext, self.ndof, "f_ext", B if fext_b else None)
        if r:
            out["r"] = torch.empty_like(u)
        if half_sq:
            out["half_sq"] = torch.empty(B, dtype=torch.float64, device=self.device)
        if max_strain:
            out["max_strain"] = torch.empty(B, dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            check(self._lib.pf_residual(self._handle, KINDS[kind], B, _ptr(u), _ptr(E), _ptr(A), mb,
                                        _ptr(out.get("f_int")), _ptr(f_ext), fext_b, float(load_factor),
                                        _ptr(out.get("r")), _ptr(out.get("half_sq")), _ptr(out.get("max_strain")),
                                        _stream_ptr(self.device)))
        return out

    def internal_force(self, u, E, A, kind="linear") -> torch.Tensor:
        return self.residual(u, E, A, kind=kind)["f_int"]

    def residual_into(self, u, E, A, f_ext, load_factor, r_out, kind="linear"):
        """Residual only, into a preallocated tensor (the bench's hot call)."""
        B = self._chk(u, self.ndof, "u")
        mb = self._materials(E, A, B)
        fext_b = 1 if (f_ext.dim() == 2 and B > 1) else 0
        check(self._lib.pf_residual(self._handle, KINDS[kind], B, _ptr(u), _ptr(E), _ptr(A), mb, None, _ptr(f_ext),
                                    fext_b, float(load_factor), _ptr(r_out), None, None, _stream_ptr(self.device)))
        return r_out

    def tangent_matvec(self, v, E, A, u=None, kind="linear") -> torch.Tensor:
        self._need_device()
        B = self._chk(v, self.ndof, "v")
        if u is not None:
            self._chk(u, self.ndof, "u", B)
        mb = self._materials(E, A, B)
        out = torch.empty_like(v)
        with torch.cuda.device(self.device):
            check(self._lib.pf_tangent_matvec(self._handle, KINDS[kind], B, _ptr(u), _ptr(E), _ptr(A), mb, _ptr(v),
                                              _ptr(out), _stream_ptr(self.device)))
        return out

    def material_vjp(self, u, E, A, g, kind="linear") -> Tuple[torch.Tensor, torch.Tensor]:
        self._need_device()
        B = self._chk(u, self.ndof, "u")
        self._chk(g, self.ndof, "g", B)
        mb = self._materials(E, A, B)
        shape = (self.nelem,) if u.dim() == 1 else (self.nelem, B)
        gE = torch.empty(shape, dtype=torch.float64, device=self.device)
        gA = torch.empty(shape, dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            check(self._lib.pf_material_vjp(self._handle, KINDS[kind], B, _ptr(u), _ptr(E), _ptr(A), mb, _ptr(g),
                                            _ptr(gE), _ptr(gA), _stream_ptr(self.device)))
        return gE, gA

    def tangent_bsr(self, E, A, u=None, kind="linear", B: Optional[int] = None,
                    out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Node-block CSR values ``[nnzb, dim, dim]`` (or ``[nnzb, dim*dim, B]``); ``out`` reuses a
        buffer of that shape."""
        self._need_device()
        if u is not None:
            B = self._chk(u, self.ndof, "u")
        elif B is None:
            B = 1 if E.dim() == 1 else int(E.shape[1])
        mb = self._materials(E, A, B)
        d = self.dim
        batched = (u is not None and u.dim() == 2) or (u is None and E.dim() == 2)
        shape = (self.nnzb, d * d, B) if batched else (self.nnzb, d, d)
        if out is not None:
            if tuple(out.shape) != shape or out.dtype != torch.float64 or not out.is_contiguous() \
                    or out.device != self.device:
                raise ValueError(f"out must be a contiguous float64 tensor of shape {shape} on {self.device}")
            vals = out
        else:
            vals = torch.empty(shape, dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            check(self._lib.pf_tangent_bsr(self._handle, KINDS[kind], B, _ptr(u), _ptr(E), _ptr(A), mb, _ptr(vals),
                                           _stream_ptr(self.device)))
        return vals

    def element_strain(self, u, measure="linear") -> torch.Tensor:
        """Signed strain per element ``[nelem(, B)]``: "linear", "green_lagrange" or "engineering"."""
        self._need_device()
        B = self._chk(u, self.ndof, "u")
        m = {"linear": 0, "green_lagrange": 1, "gl": 1, "engineering": 2}[measure]
        out = torch.empty((self.nelem,) if u.dim() == 1 else (self.nelem, B), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            check(self._lib.pf_element_strain(self._handle, m, B, _ptr(u), _ptr(out), _stream_ptr(self.device)))
        return out

    def bsr_to_dense(self, vals, free_only: bool = False) -> torch.Tensor:
        self._need_device()
        n = self.nfree if free_only else self.ndof
        K = torch.empty((n, n), dtype=torch.float64, device=self.device)
        fn = self._lib.pf_bsr_to_free_dense if free_only else self._lib.pf_bsr_to_dense
        with torch.cuda.device(self.device):
            check(fn(self._handle, _ptr(vals), _ptr(K), _stream_ptr(self.device)))
        return K

    def tangent_dense(self, E, A, u=None, kind="linear", free_only=False) -> torch.Tensor:
        return self.bsr_to_dense(self.tangent_bsr(E, A, u, kind), free_only)

    def residual_host(self, u_host, E_host, A_host, f_ext_host, load_factor=1.0, r_host=None, kind="linear",
                      chunk: int = 64):
        """End-to-end call on HOST tensors ``[rows, B]`` (pinned for full speed):
        chunks are copied in, evaluated and copied back on overlapping streams."""
        self._need_device()
        for t in (u_host, E_host, A_host, f_ext_host):
            if t.device.type != "cpu" or t.dtype != torch.float64 or not t.is_contiguous():
                raise ValueError("residual_host expects contiguous float64 CPU tensors")
        B = 1 if u_host.dim() == 1 else int(u_host.shape[1])
        if r_host is None:
            r_host = torch.empty_like(u_host, pin_memory=u_host.is_pinned())
        check(self._lib.pf_residual_host(self._handle, KINDS[kind], B, _ptr(u_host), _ptr(E_host), _ptr(A_host),
                                         _ptr(f_ext_host), float(load_factor), _ptr(r_host), int(chunk)))
        return r_host
