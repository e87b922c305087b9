"""Element sharding of ONE large mesh over the GPUs of a box (second sharding).

Nodes are partitioned between ranks (default: contiguous node-id ranges, which for the benchmark
lattice are row bands).  A rank's *local mesh* holds

* its owned nodes first (ascending global id), then its halo nodes (ascending global id): the nodes
  that its elements touch but other ranks own;
* every element incident to an owned node, in ascending global element id -- so the owned rows of
  ``f_int`` are summed in the same order as on a single GPU and carry the same bits.

An element on a partition interface is local to both neighbours; its *owner* is the owner of its
first node (only the owner contributes to ``dL/dtheta``).  The partition, halo and exchange lists
are pure integer work done identically on every rank from the replicated global mesh: no
communication is needed to set up.  At run time the ranks swap halo rows of ``u`` and ``r`` and
all-reduce one short buffer per iteration (``pf_halo_exchange`` / ``pf_comm_allreduce_sum``, called from
inside ``pf_gd_solve_sharded``): through peer memory over NVLink -- the sending kernel stores into the
receiver's mailbox, the all-reduce runs inside the Adam kernel -- or, without peer access, NCCL.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import sharding


def contiguous_node_partition(nnode: int, world: int) -> np.ndarray:
    """``part[n]`` = owner rank of node n: balanced contiguous id ranges."""
    part = np.empty(nnode, dtype=np.int32)
    for r in range(world):
        s = sharding.shard_range(nnode, r, world)
        part[s.start:s.stop] = r
    return part


def coordinate_bisection_partition(nodes, world: int) -> np.ndarray:
    """``part[n]`` = owner rank of node n by recursive coordinate bisection: the node set is cut across its longest
    extent into two groups sized in proportion to the ranks they will hold, recursively, so every rank gets
    ``nnode // world`` or one more nodes in a compact region whatever the node numbering ("general
    meshes: graph partition").  Ties are broken by node id, so the result is deterministic; contiguous id ranges
    (``contiguous_node_partition``) are the special case of a mesh numbered along one axis."""
    xy = np.asarray(nodes, dtype=np.float64)
    if xy.ndim == 1:
        xy = xy[:, None]
    nnode = xy.shape[0]
    if world < 1 or world > max(nnode, 1):
        raise ValueError(f"cannot give {world} ranks at least one of {nnode} nodes each")
    part = np.empty(nnode, dtype=np.int32)
    counts = sharding.shard_counts(nnode, world)  # nodes per rank, in rank order
    stack = [(np.arange(nnode, dtype=np.int64), 0, world)]
    while stack:
        ids, r0, nr = stack.pop()
        if nr == 1:
            part[ids] = r0
            continue
        half = nr // 2
        n_left = int(sum(counts[r0:r0 + half]))
        ext = xy[ids].max(axis=0) - xy[ids].min(axis=0)
        axis = int(np.argmax(ext))
        order = np.lexsort((ids, xy[ids, axis]))  # by coordinate, then by id
        stack.append((ids[order[:n_left]], r0, half))
        stack.append((ids[order[n_left:]], r0 + half, nr - half))
    return part


@dataclass
class LocalMesh:
    """One rank's view of the partitioned mesh (host arrays; all ids int64 unless noted)."""
    rank: int
    world: int
    owned: np.ndarray            # global ids of owned nodes, ascending
    halo: np.ndarray             # global ids of halo nodes, ascending
    elements_global: np.ndarray  # global ids of local elements, ascending
    elements: np.ndarray         # [nelem_local, 2] in LOCAL node ids
    elem_owned: np.ndarray       # uint8 [nelem_local]
    nodes: np.ndarray            # coordinates of local nodes [nlocal(, dim)]
    fixed_dofs: np.ndarray       # LOCAL ids of fixed DOFs (owned and halo nodes)
    peers: np.ndarray            # int32 ranks exchanged with, ascending
    send_ptr: np.ndarray         # int64 [n_peers + 1]
    send_nodes: np.ndarray       # int32 LOCAL ids of owned nodes each peer needs
    recv_ptr: np.ndarray
    recv_nodes: np.ndarray       # int32 LOCAL ids of halo nodes each peer owns
    dim: int = 2
    nfree_global: int = 0

    @property
    def n_owned(self) -> int:
        return int(self.owned.size)

    @property
    def local_nodes_global(self) -> np.ndarray:
        return np.concatenate([self.owned, self.halo])

    def local_dofs_global(self) -> np.ndarray:
        g = self.local_nodes_global
        return (g[:, None] * self.dim + np.arange(self.dim)[None, :]).reshape(-1)

    def to_local_vector(self, x_global: np.ndarray) -> np.ndarray:
        """Rows of a global DOF vector ``[ndof(, B)]`` in local order (owned then halo)."""
        return np.ascontiguousarray(np.asarray(x_global)[self.local_dofs_global()])


def partition_mesh(nodes, elements, fixed_dofs, world: int, part: Optional[np.ndarray] = None) -> List[LocalMesh]:
    """Local meshes of all ranks (every rank can compute every rank's: the exchange lists of two
    neighbours are derived from the same data, so they match by construction)."""
    nodes = np.asarray(nodes, dtype=np.float64)
    dim = 1 if nodes.ndim == 1 else nodes.shape[1]
    nnode = nodes.shape[0]
    el = np.asarray(elements, dtype=np.int64).reshape(-1, 2)
    part = contiguous_node_partition(nnode, world) if part is None else np.asarray(part, dtype=np.int32)
    if part.shape != (nnode,) or part.min() < 0 or part.max() >= world:
        raise ValueError("part must assign every node a rank in [0, world)")
    if np.bincount(part, minlength=world).min() == 0:
        raise ValueError(f"every rank must own at least one node ({nnode} nodes over {world} ranks)")
    fixed = np.unique(np.asarray(fixed_dofs, dtype=np.int64))
    ndof = nnode * dim
    nfree_global = ndof - fixed.size
    is_fixed = np.zeros(ndof, dtype=bool)
    is_fixed[fixed] = True
    own_i, own_j = part[el[:, 0]], part[el[:, 1]]
    out: List[LocalMesh] = []
    halos: Dict[int, np.ndarray] = {}
    owned_of: Dict[int, np.ndarray] = {}
    for r in range(world):
        owned = np.flatnonzero(part == r).astype(np.int64)
        eg = np.flatnonzero((own_i == r) | (own_j == r)).astype(np.int64)  # ascending global element id
        touched = np.unique(el[eg].reshape(-1)) if eg.size else np.empty(0, dtype=np.int64)
        halo = touched[part[touched] != r]
        owned_of[r], halos[r] = owned, halo
        local_global = np.concatenate([owned, halo])
        g2l = np.full(nnode, -1, dtype=np.int64)
        g2l[local_global] = np.arange(local_global.size)
        el_local = g2l[el[eg]]
        gd = (local_global[:, None] * dim + np.arange(dim)[None, :]).reshape(-1)
        fixed_local = np.flatnonzero(is_fixed[gd]).astype(np.int64)
        out.append(LocalMesh(rank=r, world=world, owned=owned, halo=halo, elements_global=eg, elements=el_local,
                             elem_owned=(own_i[eg] == r).astype(np.uint8), nodes=nodes[local_global],
                             fixed_dofs=fixed_local, peers=np.empty(0, np.int32), send_ptr=np.zeros(1, np.int64),
                             send_nodes=np.empty(0, np.int32), recv_ptr=np.zeros(1, np.int64),
                             recv_nodes=np.empty(0, np.int32), dim=dim, nfree_global=int(nfree_global)))
    # exchange lists: rank r receives its halo nodes from their owners; q sends exactly those, in the same order
    for r in range(world):
        lm = out[r]
        need_from = {int(q): lm.halo[part[lm.halo] == q] for q in np.unique(part[lm.halo])} if lm.halo.size else {}
        give_to = {}
        for q in range(world):
            if q == r or not halos[q].size:
                continue
            mine = halos[q][part[halos[q]] == r]
            if mine.size:
                give_to[q] = mine
        peers = sorted(set(need_from) | set(give_to))
        g2l = np.full(nnode, -1, dtype=np.int64)
        g2l[lm.local_nodes_global] = np.arange(lm.local_nodes_global.size)
        sp, rp, sn, rn = [0], [0], [], []
        for q in peers:
            s = g2l[give_to.get(q, np.empty(0, np.int64))]
            t = g2l[need_from.get(q, np.empty(0, np.int64))]
            sn.append(s)
            rn.append(t)
            sp.append(sp[-1] + s.size)
            rp.append(rp[-1] + t.size)
        lm.peers = np.asarray(peers, dtype=np.int32)
        lm.send_ptr, lm.recv_ptr = np.asarray(sp, dtype=np.int64), np.asarray(rp, dtype=np.int64)
        lm.send_nodes = (np.concatenate(sn) if sn else np.empty(0)).astype(np.int32)
        lm.recv_nodes = (np.concatenate(rn) if rn else np.empty(0)).astype(np.int32)
    return out


class Communicator:
    """Communicator owned by libpinnfem (``pf_comm``): NCCL plus, when the GPUs can map each other's memory,
    the peer-memory transport of ``csrc/pf_peer.cuh`` (halo rows stored straight into the neighbours'
    mailboxes over NVLink, all-reduce fused into the consuming kernel).  The 128-byte NCCL id is created on
    rank 0 and broadcast, the 64-byte CUDA IPC handles of the mailboxes are all-gathered, both with
    ``torch.distributed`` (any backend); collective over all ranks.

    ``transport``: ``"auto"`` (peer memory if every rank can set it up, else NCCL), ``"peer"`` (raise if it
    cannot be set up) or ``"nccl"``; the environment variable ``PF_COMM_TRANSPORT`` overrides ``"auto"``.
    ``self.transport`` holds the one in use, identical on every rank."""

    HALO_SLOT_DOUBLES = 1 << 16  # per (parity, source rank): 32768 halo nodes of a 2-D single-problem vector
    AR_SLOT_DOUBLES = 1 << 12

    def __init__(self, device, group=None, transport="auto"):
        import torch
        import torch.distributed as dist

        from . import _lib

        self._lib = _lib.load()
        self.device = torch.device(device)
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        idbuf = (C.c_ubyte * 128)()
        if self.rank == 0:
            _lib.check(self._lib.pf_comm_unique_id(idbuf))
        if self.world > 1:
            on_gpu = dist.get_backend(group) == "nccl"
            t = torch.tensor(list(idbuf), dtype=torch.uint8, device=self.device if on_gpu else "cpu")
            dist.broadcast(t, src=0, group=group)
            idbuf = (C.c_ubyte * 128)(*t.cpu().tolist())
        self._handle = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self._lib.pf_comm_create(self.world, self.rank, idbuf, self.device.index or 0,
                                                C.byref(self._handle)))
        self._group = group
        if transport == "auto":
            transport = os.environ.get("PF_COMM_TRANSPORT", "auto")
        if transport not in ("auto", "peer", "nccl"):
            raise ValueError(f"transport must be 'auto', 'peer' or 'nccl', got {transport!r}")
        self.transport = "nccl"
        if self.world > 1 and transport != "nccl":
            self._setup_peer_transport(group, required=transport == "peer")

    def _setup_peer_transport(self, group, required):
        """Export this rank's mailbox, all-gather the IPC handles, map the peers' mailboxes.  Every step is
        agreed on by all ranks (a rank that fails makes everybody fall back to NCCL, or raise)."""
        import torch
        import torch.distributed as dist

        from . import _lib

        on_gpu = dist.get_backend(group) == "nccl"
        where = self.device if on_gpu else "cpu"

        def all_ok(ok):
            t = torch.tensor([1 if ok else 0], dtype=torch.int32, device=where)
            dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
            return bool(t.item())

        hbuf = (C.c_ubyte * 64)()
        err = ""
        with torch.cuda.device(self.device):
            rc = self._lib.pf_comm_peer_export(self._handle, self.HALO_SLOT_DOUBLES, self.AR_SLOT_DOUBLES, hbuf)
            if rc != 0:
                err = _lib.last_error()
            mine = torch.tensor(list(hbuf), dtype=torch.uint8, device=where)
            parts = [torch.empty_like(mine) for _ in range(self.world)]
            dist.all_gather(parts, mine, group=group)  # also orders every rank's mailbox reset before any use
            ok = all_ok(rc == 0)
            if ok:
                handles = (C.c_ubyte * (64 * self.world))(*torch.cat(parts).cpu().tolist())
                rc = self._lib.pf_comm_peer_import(self._handle, handles)
                if rc != 0:
                    err = _lib.last_error()
                ok = all_ok(rc == 0)
            if ok:
                self.transport = "peer"
            else:
                self._lib.pf_comm_peer_detach(self._handle)
                if required:
                    raise RuntimeError(f"peer-memory transport unavailable on rank {self.rank}: {err or 'a peer failed'}")

    def check(self):
        """Synchronise and raise if a peer-memory wait timed out (a neighbour never arrived)."""
        import torch

        from . import _lib

        with torch.cuda.device(self.device):
            _lib.check(self._lib.pf_comm_peer_check(self._handle, C.c_void_p(
                torch.cuda.current_stream(self.device).cuda_stream)))

    def allreduce_sum_(self, x):
        import torch

        from . import _lib

        if x.dtype != torch.float64 or not x.is_contiguous() or x.device != self.device:
            raise ValueError("allreduce_sum_ needs a contiguous float64 tensor on the communicator's device")
        with torch.cuda.device(self.device):
            _lib.check(self._lib.pf_comm_allreduce_sum(self._handle, C.c_void_p(x.data_ptr()), x.numel(),
                                                       C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))
        return x

    def close(self, _collective=True):
        """Collective when the peer transport is on: every rank unmaps its peers' mailboxes before any rank
        frees its own (CUDA IPC requires the importers to close first)."""
        if getattr(self, "_handle", None) and self._handle.value:
            if getattr(self, "transport", "nccl") == "peer":
                import torch.distributed as dist

                self._lib.pf_comm_peer_detach(self._handle)
                self.transport = "nccl"
                if _collective and dist.is_available() and dist.is_initialized():
                    dist.barrier(group=self._group)
            self._lib.pf_comm_destroy(self._handle)
            self._handle = C.c_void_p()

    def __del__(self):
        try:
            self.close(_collective=False)  # garbage collection is not synchronised between ranks: no barrier here
        except Exception:
            pass


class ShardedMesh:
    """This rank's local mesh on the device: assembly plan + halo exchange."""

    def __init__(self, nodes, elements, fixed_dofs, comm: Communicator, part: Optional[np.ndarray] = None):
        import torch

        from . import _lib
        from .plan import AssemblyPlan

        self.comm = comm
        self.local = partition_mesh(nodes, elements, fixed_dofs, comm.world, part)[comm.rank]
        lm = self.local
        self.plan = AssemblyPlan(lm.nodes, lm.elements, lm.fixed_dofs, dim=lm.dim, device=comm.device)
        self._lib = _lib.load()
        self._halo = C.c_void_p()
        keep = [np.ascontiguousarray(a) for a in (lm.peers, lm.send_ptr, lm.send_nodes, lm.recv_ptr, lm.recv_nodes)]
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        with torch.cuda.device(comm.device):
            _lib.check(self._lib.pf_halo_create(comm._handle, lm.dim, int(lm.peers.size), p(keep[0]), p(keep[1]),
                                                p(keep[2]), p(keep[3]), p(keep[4]), C.byref(self._halo)))
        if comm.world > 1:  # every rank must pick the same transport for an exchange: agree on the largest message
            import torch.distributed as dist

            on_gpu = dist.get_backend(comm._group) == "nccl"
            t = torch.tensor([int(self._lib.pf_halo_max_message_nodes(self._halo))], dtype=torch.int64,
                             device=comm.device if on_gpu else "cpu")
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=comm._group)
            _lib.check(self._lib.pf_halo_set_max_message_nodes(self._halo, int(t.item())))
        self.elem_owned = torch.as_tensor(lm.elem_owned).to(comm.device)
        self.n_owned_dofs = lm.n_owned * lm.dim

    def exchange_(self, x):
        """Overwrite the halo rows of ``x`` ``[ndof_local(, B)]`` with the owners' values."""
        import torch

        from . import _lib

        B = 1 if x.dim() == 1 else int(x.shape[1])
        if x.dtype != torch.float64 or not x.is_contiguous() or x.shape[0] != self.plan.ndof:
            raise ValueError(f"x must be a contiguous float64 [{self.plan.ndof}(, B)] tensor")
        with torch.cuda.device(self.comm.device):
            _lib.check(self._lib.pf_halo_exchange(self._halo, C.c_void_p(x.data_ptr()), B, C.c_void_p(
                torch.cuda.current_stream(self.comm.device).cuda_stream)))
        return x

    def internal_force_owned(self, u_local, E_local, A_local):
        """``f_int`` of this rank's owned DOFs ``[n_owned_dofs(, B)]``: halo exchange of ``u`` + local gather."""
        self.exchange_(u_local)
        return self.plan.internal_force(u_local, E_local, A_local)[: self.n_owned_dofs]

    def close(self):
        if getattr(self, "_halo", None) and self._halo.value:
            self._lib.pf_halo_destroy(self._halo)
            self._halo = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def gd_solve_element_sharded(mesh: ShardedMesh, nets, scales, theta, u_global, f_ext_global, meas_dofs=None,
                             meas_vals=None, *, max_iterations=1000, tolerance=1e-6, learning_rate_u=1e-7,
                             learning_rate_theta=1e-4, alpha_physics=1.0, alpha_data=100.0, load_factor=1.0,
                             record_history=True, legacy_loss=False):
    """PINN gradient descent on one mesh sharded by element over all ranks (``pf_gd_solve_sharded``).
    ``theta`` ``[n_theta]`` (replicated), ``u_global`` / ``f_ext_global`` ``[ndof]`` host or device arrays of the
    WHOLE mesh, ``meas_dofs`` / ``meas_vals`` global measurement lists.  Returns a dict with this rank's owned
    rows (``u_owned``, ``reactions_owned``, ``owned_dofs`` = their global ids) and the replicated ``theta``,
    ``history``, ``n_iters``, ``converged``."""
    import torch

    from . import _lib
    from .ops import _dev_f64

    lm, plan, dev = mesh.local, mesh.plan, mesh.comm.device
    cfg = _lib.GDConfig()
    cfg.max_iterations, cfg.kind, cfg.tolerance = int(max_iterations), _lib.ELEM_LINEAR, float(tolerance)
    cfg.learning_rate_u, cfg.learning_rate_theta = float(learning_rate_u), float(learning_rate_theta)
    cfg.alpha_physics, cfg.alpha_data, cfg.load_factor = float(alpha_physics), float(alpha_data), float(load_factor)
    cfg.loss_mode = 1 if legacy_loss else 0
    ntheta = 0
    for k in range(3):
        spec = nets[k] if k < len(nets) else None
        cfg.net_enabled[k] = 1 if spec is not None else 0
        cfg.net_scale[k] = float(scales[k]) if k < len(scales) else 0.0
        if spec is not None:
            cfg.net_input_dim[k], cfg.net_hidden_layers[k], cfg.net_width[k] = spec.input_dim, spec.hidden_layers, spec.width
            ntheta += spec.n_params
    theta_d = _dev_f64(torch.as_tensor(theta, dtype=torch.float64).to(dev).reshape(-1).clone(), "theta") if ntheta else None
    if ntheta and theta_d.numel() != ntheta:
        raise ValueError(f"theta must have {ntheta} entries")
    to_np = lambda a: a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a, dtype=np.float64)
    u = torch.as_tensor(lm.to_local_vector(to_np(u_global))).to(dev).contiguous()
    f_ext = torch.as_tensor(lm.to_local_vector(to_np(f_ext_global))).to(dev).contiguous()
    n_meas_all, md_l, mv_l = 0, None, None
    if meas_dofs is not None and meas_vals is not None and len(meas_dofs) > 0:
        md_g, mv_g = np.asarray(meas_dofs, dtype=np.int64), to_np(meas_vals).reshape(-1)
        n_meas_all = int(md_g.size)
        g2l = np.full(int(max(md_g.max() + 1, lm.local_dofs_global().max() + 1)), -1, dtype=np.int64)
        g2l[lm.local_dofs_global()[: mesh.n_owned_dofs]] = np.arange(mesh.n_owned_dofs)
        sel = g2l[md_g] >= 0  # measurements on my owned DOFs, in their global order
        if sel.any():
            md_l = torch.as_tensor(g2l[md_g[sel]].astype(np.int32)).to(dev)
            mv_l = torch.as_tensor(mv_g[sel]).to(dev).contiguous()
    cfg.n_measured = int(md_l.numel()) if md_l is not None else 0
    shard = _lib.GDShard(halo=mesh._halo.value, n_owned_nodes=lm.n_owned, elem_owned=mesh.elem_owned.data_ptr(),
                         nfree_global=lm.nfree_global, n_measured_global=n_meas_all, reserved=0)
    history = (torch.zeros((max(int(max_iterations), 1), _lib.GD_HISTORY_COLS), dtype=torch.float64, device=dev)
               if record_history else None)
    n_iters = torch.zeros(1, dtype=torch.int32, device=dev)
    converged = torch.zeros(1, dtype=torch.int32, device=dev)
    reactions = torch.zeros(plan.ndof, dtype=torch.float64, device=dev)
    p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.device(dev):
        e0.record()
        _lib.check(mesh._lib.pf_gd_solve_sharded(plan._handle, C.byref(cfg), C.byref(shard), p(theta_d), p(u), p(f_ext),
                                                 p(md_l), p(mv_l), p(history), p(n_iters), p(converged), p(reactions),
                                                 C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
        e1.record()
    n = int(n_iters[0])
    return {"u_owned": u[: mesh.n_owned_dofs], "u_local": u, "reactions_owned": reactions[: mesh.n_owned_dofs],
            "owned_dofs": lm.local_dofs_global()[: mesh.n_owned_dofs], "theta": theta_d,
            "history": history[:n] if history is not None else None, "n_iters": n, "converged": bool(converged[0]),
            "solve_ms": e0.elapsed_time(e1)}  # device time of the loop itself (staging of the local vectors excluded)
