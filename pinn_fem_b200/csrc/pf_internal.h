// Internal declarations shared by the translation units of libpinnfem.so.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>
#include <string>
#include <vector>

#include "pinnfem.h"

// ---------------------------------------------------------------------------
// error reporting
// ---------------------------------------------------------------------------
void pf_set_error(const char* fmt, ...);

#define PF_CUDA_CHECK(expr)                                                                      \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess) {                                                                 \
            pf_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return (_e == cudaErrorNoDevice || _e == cudaErrorInsufficientDriver) ? PF_ERR_NO_DEVICE \
                                                                                  : PF_ERR_CUDA; \
        }                                                                                        \
    } while (0)

#define PF_REQUIRE(cond, ...)         \
    do {                              \
        if (!(cond)) {                \
            pf_set_error(__VA_ARGS__); \
            return PF_ERR_ARG;        \
        }                             \
    } while (0)

// ---------------------------------------------------------------------------
// plan
// ---------------------------------------------------------------------------

// One node->element incidence as the device sees it (16 bytes of indices plus
// 32 bytes of geometry, read warp-uniformly by the gather kernels).
struct PfIncidence {
    int32_t elem;  // element id
    int32_t nbr;   // the element's other node
    int32_t slot;  // BSR slot of block (node, nbr)
    int32_t flags; // bit0: first incidence of this row that targets `slot`
};

// A patch is a compact set of <= kPatchNodes nodes (recursive coordinate
// bisection) whose element data is staged once per CTA in shared memory.
constexpr int kPatchNodes = 32;

struct PfPatch {
    int32_t node_off;   // into patch_nodes (owned nodes first, then halo nodes)
    int32_t n_owned;
    int32_t n_local;    // owned + halo
    int32_t elem_off;   // into patch_elems
    int32_t n_elem;     // elements incident to owned nodes (ascending element id)
    int32_t inc_off;    // into patch_inc; patch_inc_ptr is indexed by node_off + local owned index
    int32_t ptr_off;    // into patch_inc_ptr (n_owned + 1 entries, relative to inc_off)
    int32_t pad;
};

struct PfPatchInc {
    int16_t lelem;  // index into the patch's element list
    int16_t lnbr;   // local index of the other node
};

struct pf_plan {
    int dim = 2;
    int64_t nnode = 0, nelem = 0, ndof = 0, nfree = 0, nfixed = 0, nnzb = 0, ninc = 0;
    int max_degree = 0;
    bool has_dup = false;

    // host arrays (int64 where they are exposed for bit-exact checks)
    std::vector<int32_t> conn;        // [nelem][2]
    std::vector<double> nodes;        // [nnode][dim]
    std::vector<int64_t> elem_dofs;   // [nelem][2*dim]
    std::vector<int64_t> free_dofs, fixed_dofs;
    std::vector<int64_t> bsr_rowptr, bsr_colind, elem_slots;
    std::vector<int64_t> inc_ptr, inc_elem, inc_nbr, inc_slot, diag_slot;
    std::vector<uint8_t> inc_first;
    std::vector<uint8_t> dof_free;    // [ndof] 1 = free
    std::vector<double> l0, cosv, sinv, centroid;  // per element ([nelem][dim] for centroid)

    // device copies
    int device = -1;
    int sm_count = 148;
    int32_t* d_inc_ptr = nullptr;       // [nnode+1]
    PfIncidence* d_inc = nullptr;       // [ninc]
    double4* d_inc_geo = nullptr;       // [ninc] {cos, sin, 1/l0, l0}
    double4* d_inc_xy = nullptr;        // [ninc] {x_nbr, y_nbr, x_self, y_self}
    int2* d_inc2 = nullptr;             // [ninc] {elem, nbr}: the compact table of the single-problem kernels
    double* d_nodes = nullptr;          // [nnode][dim] coordinates
    int32_t* d_diag_slot = nullptr;     // [nnode]
    int2* d_conn = nullptr;             // [nelem]
    double4* d_elem_geo = nullptr;      // [nelem] {cos, sin, 1/l0, l0}
    double4* d_elem_xy = nullptr;       // [nelem] {x_i, y_i, x_j, y_j}
    double* d_centroid = nullptr;       // [nelem][dim]
    uint8_t* d_dof_free = nullptr;      // [ndof]
    int32_t* d_free_dofs = nullptr;     // [nfree]
    int32_t* d_free_index = nullptr;    // [ndof] position in free_dofs or -1
    int32_t* d_bsr_rowptr = nullptr;    // [nnode+1]
    int32_t* d_bsr_colind = nullptr;    // [nnzb]

    // patches (host + device)
    std::vector<PfPatch> patches;
    std::vector<int32_t> patch_nodes, patch_elems, patch_inc_ptr;
    std::vector<PfPatchInc> patch_inc;
    int max_patch_elems = 0, max_patch_local = 0, max_patch_inc = 0;
    bool patch_ok = false;  // patch tables fit the shared-memory kernel
    PfPatch* d_patches = nullptr;
    int32_t* d_patch_nodes = nullptr;
    int32_t* d_patch_elems = nullptr;
    int32_t* d_patch_inc_ptr = nullptr;   // [sum(n_owned + 1)] offsets relative to the patch's inc_off
    PfPatchInc* d_patch_inc = nullptr;
    double4* d_patch_inc_geo = nullptr;   // {cos, sin, 1/l0, l0} per patch incidence

    // growable scratch for deterministic two-stage reductions
    double* d_work = nullptr;
    size_t work_bytes = 0;

    // streams / staging for pf_residual_host
    cudaStream_t host_streams[3] = {nullptr, nullptr, nullptr};
    double* d_stage[3][4] = {{nullptr}};
    size_t stage_elems[4] = {0, 0, 0, 0};
};

int pf_plan_reserve_work(pf_plan* plan, size_t bytes);
// Raise the release threshold of the current device's default stream-ordered pool (once per device): freed
// scratch stays cached instead of going back to the driver at every synchronisation.
void pf_keep_pool_cached();

// pf_patch.cu: shared-memory staged gather for wide batches (linear element).
// Returns PF_OK and sets *handled when it ran; otherwise the caller uses the generic kernel.
struct PfGatherCall {
    int mode;  // 0 force/residual, 1 mat-vec
    int64_t B, ldb;
    const double *u, *v, *E, *A, *f_ext;
    int mat_batched, fext_batched;
    double load_factor;
    double *f_out, *r_out, *half_sq, *max_strain;
};
int pf_patch_gather(pf_plan* plan, const PfGatherCall& c, cudaStream_t st, int64_t* columns_done);
int pf_plan_activate(const pf_plan* plan);  // cudaSetDevice + uploaded check

// pf_gd_large.cu: the PINN-GD loop as a device-resident sequence of kernels (large meshes)
int pf_gd_solve_large(pf_plan* plan, const pf_gd_config* cfg, int64_t nprob, double* theta, double* u,
                      const double* f_ext, const int32_t* meas_dofs, const double* meas_vals, double* history,
                      int32_t* n_iters, int32_t* converged, double* reactions, cudaStream_t st,
                      const pf_gd_shard* shard = nullptr);

// pf_comm.cu
int pf_comm_world(const pf_comm* c);
pf_comm* pf_halo_comm(pf_halo* h);
bool pf_halo_uses_peer(const pf_halo* h, int64_t B);  // this exchange goes through the peers' mailboxes
struct PfPeerView;
bool pf_comm_peer_view(const pf_comm* c, PfPeerView* out);  // false: NCCL transport

static inline cudaStream_t pf_stream_of(void* s) { return reinterpret_cast<cudaStream_t>(s); }
