"""ctypes binding of libpinnfem.so (the C ABI declared in include/pinnfem.h).

There is deliberately no fallback: if the shared library is missing, or a
kernel is requested without a CUDA device, the call raises.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("PINNFEM_LIB", _HERE / "libpinnfem.so"))

PF_OK = 0
PF_ERR_ARG, PF_ERR_CUDA, PF_ERR_NO_DEVICE, PF_ERR_SINGULAR, PF_ERR_GEOMETRY = 1, 2, 3, 4, 5
ELEM_LINEAR, ELEM_GREEN_LAGRANGE = 0, 1

(PLAN_DIM, PLAN_NNODE, PLAN_NELEM, PLAN_NDOF, PLAN_NFREE, PLAN_NFIXED, PLAN_NNZB, PLAN_NINC, PLAN_MAX_DEGREE,
 PLAN_HAS_DUPLICATE_EDGES, PLAN_DEVICE) = range(11)
(ARR_ELEM_DOFS, ARR_FREE_DOFS, ARR_FIXED_DOFS, ARR_BSR_ROWPTR, ARR_BSR_COLIND, ARR_ELEM_SLOTS, ARR_INC_PTR,
 ARR_INC_ELEM, ARR_INC_NBR, ARR_INC_SLOT, ARR_DIAG_SLOT) = range(11)
GD_HISTORY_COLS = 7


class PinnFemError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(message)
        self.code = code


class GDConfig(C.Structure):
    """Mirror of ``pf_gd_config`` (include/pinnfem.h)."""

    _fields_ = [
        ("max_iterations", C.c_int32),
        ("kind", C.c_int32),
        ("tolerance", C.c_double),
        ("learning_rate_u", C.c_double),
        ("learning_rate_theta", C.c_double),
        ("alpha_physics", C.c_double),
        ("alpha_data", C.c_double),
        ("load_factor", C.c_double),
        ("net_enabled", C.c_int32 * 3),
        ("net_input_dim", C.c_int32 * 3),
        ("net_hidden_layers", C.c_int32 * 3),
        ("net_width", C.c_int32 * 3),
        ("net_scale", C.c_double * 3),
        ("n_measured", C.c_int32),
        ("loss_mode", C.c_int32),
    ]


class GDShard(C.Structure):
    """Mirror of ``pf_gd_shard`` (include/pinnfem.h)."""

    _fields_ = [
        ("halo", C.c_void_p),
        ("n_owned_nodes", C.c_int64),
        ("elem_owned", C.c_void_p),
        ("nfree_global", C.c_int64),
        ("n_measured_global", C.c_int32),
        ("reserved", C.c_int32),
    ]


_vp, _i64, _i32, _dbl, _int = C.c_void_p, C.c_int64, C.c_int32, C.c_double, C.c_int

# name -> (restype, argtypes); must list every symbol of include/pinnfem.h
SIGNATURES = {
    "pf_last_error": (C.c_char_p, []),
    "pf_version": (_int, []),
    "pf_device_count": (_int, []),
    "pf_plan_create": (_int, [_int, _i64, _i64, _vp, _vp, _vp, _i64, C.POINTER(_vp)]),
    "pf_plan_upload": (_int, [_vp, _int]),
    "pf_plan_destroy": (None, [_vp]),
    "pf_plan_size": (_i64, [_vp, _int]),
    "pf_plan_array_len": (_i64, [_vp, _int]),
    "pf_plan_get_array": (_int, [_vp, _int, _vp, _i64]),
    "pf_plan_get_geometry": (_int, [_vp, _int, _vp, _i64]),
    "pf_residual": (_int, [_vp, _int, _i64, _vp, _vp, _vp, _int, _vp, _vp, _int, _dbl, _vp, _vp, _vp, _vp]),
    "pf_tangent_matvec": (_int, [_vp, _int, _i64, _vp, _vp, _vp, _int, _vp, _vp, _vp]),
    "pf_material_vjp": (_int, [_vp, _int, _i64, _vp, _vp, _vp, _int, _vp, _vp, _vp, _vp]),
    "pf_tangent_bsr": (_int, [_vp, _int, _i64, _vp, _vp, _vp, _int, _vp, _vp]),
    "pf_element_strain": (_int, [_vp, _int, _i64, _vp, _vp, _vp]),
    "pf_bsr_to_dense": (_int, [_vp, _vp, _vp, _vp]),
    "pf_bsr_to_free_dense": (_int, [_vp, _vp, _vp, _vp]),
    "pf_mlp_num_params": (_i64, [_int, _int, _int]),
    "pf_mlp_forward": (_int, [_vp, _int, _int, _int, _vp, _i64, _vp, _dbl, _dbl, _int, _vp, _vp]),
    "pf_mlp_backward": (_int, [_vp, _int, _int, _int, _vp, _i64, _vp, _dbl, _dbl, _int, _vp, _vp, _vp]),
    "pf_mlp_acts_len": (_i64, [_int, _int, _int, _i64]),
    "pf_mlp_forward_batched": (_int, [_vp, _int, _int, _int, _vp, _i64, _i64, _i64, _vp, _dbl, _dbl, _int, _vp, _i64, _vp,
                                      _vp]),
    "pf_mlp_backward_batched": (_int, [_vp, _int, _int, _int, _vp, _i64, _i64, _i64, _vp, _dbl, _vp, _i64, _vp, _vp, _i64,
                                       _vp]),
    "pf_mlp_param_jacobian": (_int, [_vp, _int, _int, _int, _vp, _i64, _vp, _dbl, _dbl, _int, _vp, _vp]),
    "pf_gd_solve": (_int, [_vp, C.POINTER(GDConfig), _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pf_solve_dense": (_int, [_i64, _i64, _vp, _vp, _vp, _vp]),
    "pf_solve_spd": (_int, [_i64, _vp, _vp, _vp, _vp]),
    "pf_cg_solve": (_int, [_vp, _int, _i64, _vp, _vp, _vp, _int, _vp, _vp, _dbl, _int, _vp, _i64, _vp, _vp, _vp]),
    "pf_cg_work_len": (_i64, [_vp, _i64]),
    "pf_gn_normal_equations": (_int, [_i64, _i64, _vp, _vp, _dbl, _vp, _vp, _vp, _vp]),
    "pf_gn_lm_step": (_int, [_i64, _i64, _vp, _vp, _dbl, _int, _vp, _vp, _vp, _vp]),
    "pf_gn_jacobian": (_int, [_vp, _int, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _i64, _i64, _dbl, _dbl, _vp, _i64, _vp, _vp]),
    "pf_residual_host": (_int, [_vp, _int, _i64, _vp, _vp, _vp, _vp, _dbl, _vp, _i64]),
    "pf_measure_fp64_peak": (_int, [_int, C.POINTER(_dbl)]),
    "pf_debug_tanh": (_int, [_i64, _vp, _vp, _vp]),
    "pf_debug_tanh_table": (_int, [_i64, _vp, _vp, _vp]),
    "pf_comm_available": (_int, []),
    "pf_comm_unique_id": (_int, [_vp]),
    "pf_comm_create": (_int, [_int, _int, _vp, _int, C.POINTER(_vp)]),
    "pf_comm_destroy": (None, [_vp]),
    "pf_comm_allreduce_sum": (_int, [_vp, _vp, _i64, _vp]),
    "pf_comm_peer_export": (_int, [_vp, _i64, _i64, _vp]),
    "pf_comm_peer_import": (_int, [_vp, _vp]),
    "pf_comm_peer_enabled": (_int, [_vp]),
    "pf_comm_peer_detach": (None, [_vp]),
    "pf_comm_peer_check": (_int, [_vp, _vp]),
    "pf_halo_create": (_int, [_vp, _int, _int, _vp, _vp, _vp, _vp, _vp, C.POINTER(_vp)]),
    "pf_halo_destroy": (None, [_vp]),
    "pf_halo_exchange": (_int, [_vp, _vp, _i64, _vp]),
    "pf_halo_max_message_nodes": (_i64, [_vp]),
    "pf_halo_set_max_message_nodes": (_int, [_vp, _i64]),
    "pf_gd_solve_sharded": (_int, [_vp, C.POINTER(GDConfig), C.POINTER(GDShard), _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                   _vp, _vp]),
}

_lib = None


def load() -> C.CDLL:
    """Load libpinnfem.so (built by ``__graft_entry__.build()``); raise if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). pinn_fem_b200 has no CPU fallback."
        )
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    msg = load().pf_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(code: int) -> None:
    """Map C-ABI status codes to the exceptions the reference raises."""
    if code == PF_OK:
        return
    msg = last_error()
    if code == PF_ERR_SINGULAR:
        raise RuntimeError(msg or "Tangent stiffness became singular during solve")
    if code == PF_ERR_GEOMETRY:
        raise ValueError(msg or "Element with zero initial length detected")
    if code == PF_ERR_ARG:
        raise ValueError(msg)
    raise PinnFemError(code, msg)
