/* pinnfem.h -- C ABI of the B200-native PINN-FEM hot path (libpinnfem.so).
 *
 * The reference (rpacheco-blazquez/PINN-FEM) is pure Python and has no FFI of
 * its own; its seam is the Python call signatures of
 * FEM/python/fem/*.py.  Every entry point below names the reference routine
 * (file:line under FEM/python/) whose arithmetic it replaces.  The Python host
 * shim in pinn_fem_b200/ binds these with ctypes and re-exposes the
 * reference's own signatures (INTEGRATION.md shows the binding).
 *
 * Conventions
 *  - All functions return PF_OK (0) or an error code and never throw; the
 *    message is available from pf_last_error() (thread local).
 *  - "host" pointers are ordinary CPU memory; "dev" pointers are CUDA device
 *    memory owned by the caller (e.g. torch tensor .data_ptr()).  `stream` is
 *    a cudaStream_t passed as void* (NULL = default stream); kernels are
 *    enqueued and the call returns without synchronising unless stated.
 *  - Floating point is IEEE fp64 throughout.
 *  - Batched arrays carry the problem index LAST and contiguous:
 *    u[ndof][B], E[nelem][B], f_int[ndof][B]; B = 1 is the plain vector.
 *  - A plan is immutable after pf_plan_upload(); concurrent calls on
 *    different streams are safe as long as they do not share `workspace`
 *    outputs (the reductions that need scratch say so).
 */
#ifndef PINNFEM_H
#define PINNFEM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PF_VERSION 100

/* status codes */
#define PF_OK 0
#define PF_ERR_ARG 1        /* bad argument (message says which) */
#define PF_ERR_CUDA 2       /* CUDA runtime error */
#define PF_ERR_NO_DEVICE 3  /* no usable CUDA device: there is no CPU fallback */
#define PF_ERR_SINGULAR 4   /* "Tangent stiffness became singular during solve" (fem/core.py:36) */
#define PF_ERR_GEOMETRY 5   /* "Element with zero initial length detected" (fem/element.py:28,:62,:116) */

/* element kinds */
#define PF_ELEM_LINEAR 0          /* fem/element.py:15-102, fem/nn_assembly.py:18-102 */
#define PF_ELEM_GREEN_LAGRANGE 1  /* fem/element.py:105-133 (verbatim, 2-D only) */

typedef struct pf_plan pf_plan;

const char* pf_last_error(void);
int pf_version(void);
/* number of visible CUDA devices (0 when there is no driver/GPU) */
int pf_device_count(void);

/* ------------------------------------------------------------------------
 * Mesh plan: all integer/index work, done once per mesh on the host.
 * Replaces fem/geometry.py:8-9 (element_dofs), fem/boundary.py:8-13
 * (free_and_fixed_dofs) and the implicit scatter map of
 * fem/assembly.py:48,:71-72 / fem/nn_assembly.py:226-229.
 *
 *   elements  host int64 [nelem][2]
 *   nodes     host fp64  [nnode][dim]   (dim 1: [nnode])
 *   fixed     host int64 [nfixed_in]    (any order, duplicates allowed)
 * pf_plan_create needs no GPU; pf_plan_upload copies the plan to `device`.
 * ------------------------------------------------------------------------ */
int pf_plan_create(int dim, int64_t nnode, int64_t nelem, const int64_t* elements, const double* nodes,
                   const int64_t* fixed, int64_t nfixed_in, pf_plan** out);
int pf_plan_upload(pf_plan* plan, int device);
void pf_plan_destroy(pf_plan* plan);

/* scalar properties of a plan */
#define PF_PLAN_DIM 0
#define PF_PLAN_NNODE 1
#define PF_PLAN_NELEM 2
#define PF_PLAN_NDOF 3
#define PF_PLAN_NFREE 4
#define PF_PLAN_NFIXED 5
#define PF_PLAN_NNZB 6       /* number of dim x dim blocks in the node-block CSR pattern */
#define PF_PLAN_NINC 7       /* 2*nelem node->element incidences */
#define PF_PLAN_MAX_DEGREE 8
#define PF_PLAN_HAS_DUPLICATE_EDGES 9
#define PF_PLAN_DEVICE 10    /* -1 until uploaded */
int64_t pf_plan_size(const pf_plan* plan, int what);

/* host copies of the plan's index arrays, for bit-exact checks.  All int64. */
#define PF_ARR_ELEM_DOFS 0   /* [nelem][2*dim]  == element_dofs() per element */
#define PF_ARR_FREE_DOFS 1   /* [nfree]  ascending */
#define PF_ARR_FIXED_DOFS 2  /* [nfixed] sorted unique */
#define PF_ARR_BSR_ROWPTR 3  /* [nnode+1] */
#define PF_ARR_BSR_COLIND 4  /* [nnzb] ascending inside each row, diagonal included */
#define PF_ARR_ELEM_SLOTS 5  /* [nelem][4]: BSR slots of blocks (i,i),(i,j),(j,i),(j,j) */
#define PF_ARR_INC_PTR 6     /* [nnode+1] */
#define PF_ARR_INC_ELEM 7    /* [ninc] incident elements of each node, ascending element id */
#define PF_ARR_INC_NBR 8     /* [ninc] the element's other node */
#define PF_ARR_INC_SLOT 9    /* [ninc] BSR slot of block (node, other node) */
#define PF_ARR_DIAG_SLOT 10  /* [nnode] BSR slot of block (node, node) */
int64_t pf_plan_array_len(const pf_plan* plan, int which);
int pf_plan_get_array(const pf_plan* plan, int which, int64_t* dst_host, int64_t dst_len);
/* fp64 geometry the plan precomputed: which = 0 l0[nelem], 1 cos[nelem], 2 sin[nelem],
 * 3 centroids[nelem][dim] */
int pf_plan_get_geometry(const pf_plan* plan, int which, double* dst_host, int64_t dst_len);

/* ------------------------------------------------------------------------
 * Element internal force / residual (fp64, deterministic, matrix-free).
 * Replaces the loop of fem/assembly.py:52-73 and fem/nn_assembly.py:183-229
 * (f_int part): per element k = E*A/l0, fe = ke @ u_e, f_int[dofs] += fe.
 * The sum for every DOF runs over its incident elements in ascending element
 * id, the order the reference's `+=` produces.
 *
 *   u        dev [ndof][B]
 *   E, A     dev [nelem][B] when mat_batched != 0, else [nelem] shared by all problems
 *   f_int    dev [ndof][B] out, may be NULL
 *   f_ext    dev [ndof] (fext_batched == 0) or [ndof][B]; may be NULL when r is NULL
 *   r        dev [ndof][B] out, may be NULL:  r = f_int - load_factor*f_ext on
 *            free DOFs, 0 on fixed DOFs (fem/solver.py:267-269, fem/core.py:29)
 *   half_sq  dev [B] out, may be NULL: 0.5*sum(r^2) per problem (fem/solver.py:270);
 *            uses the plan's workspace (one call at a time per plan)
 *   max_strain dev [B] out, may be NULL: max |strain| (fem/assembly.py:73)
 * ------------------------------------------------------------------------ */
int pf_residual(pf_plan* plan, int kind, int64_t B, const double* u, const double* E, const double* A,
                int mat_batched, double* f_int, const double* f_ext, int fext_batched, double load_factor,
                double* r, double* half_sq, double* max_strain, void* stream);

/* out = K_t(u) @ v without forming K (linear kind: u may be NULL).  For the
 * linear element this is also the VJP dL/du = K^T g that torch autograd
 * computes through fem/nn_assembly.py:96-100.  v, out: dev [ndof][B]. */
int pf_tangent_matvec(pf_plan* plan, int kind, int64_t B, const double* u, const double* E, const double* A,
                      int mat_batched, const double* v, double* out, void* stream);

/* dL/dE_e and dL/dA_e given g = dL/df_int (dev [ndof][B]); gE, gA dev
 * [nelem][B].  The reverse pass of fem/nn_assembly.py:72,:96-100,:226-229. */
int pf_material_vjp(pf_plan* plan, int kind, int64_t B, const double* u, const double* E, const double* A,
                    int mat_batched, const double* g, double* gE, double* gA, void* stream);

/* Tangent stiffness in node-block CSR: vals dev [nnzb][dim*dim][B] (row-major
 * dim x dim blocks; B = 1 gives a standard BSR value array).  Replaces the
 * K[np.ix_(dofs,dofs)] += ke scatter (fem/assembly.py:71, nn_assembly.py:226-229);
 * per block the element contributions are added in ascending element id. */
int pf_tangent_bsr(pf_plan* plan, int kind, int64_t B, const double* u, const double* E, const double* A,
                   int mat_batched, double* vals, void* stream);

/* Signed strain per element, strain dev [nelem][B]: measure 0 = small-strain axial
 * (fem/element.py:36,:69), 1 = Green-Lagrange (fem/element.py:126), 2 = engineering
 * (L - L0)/L0 from the deformed length (api_fem_solver.py:100-108). */
int pf_element_strain(pf_plan* plan, int measure, int64_t B, const double* u, double* strain, void* stream);

/* Dense views for the reference's dense-K callers (small meshes, B = 1):
 *   K_dense dev [ndof][ndof] row-major  (assemble_system's first return value)
 *   K_ff    dev [nfree][nfree]          (K[np.ix_(free, free)], fem/core.py:32) */
int pf_bsr_to_dense(pf_plan* plan, const double* vals, double* K_dense, void* stream);
int pf_bsr_to_free_dense(pf_plan* plan, const double* vals, double* K_ff, void* stream);

/* ------------------------------------------------------------------------
 * Material networks: SimpleNN (examples/json/generic.py:118-142) wrapped by
 * NNProperty.value (fem/properties.py:97-179): softplus(net(x)) * scale with
 * x = [load_factor, x_c(, y_c)] at element centroids (sorted dict keys).
 * theta is the flat parameter vector in nn.Module.parameters() order.
 *
 *   X        dev [n][input_dim] row-major, or NULL to use the plan's element
 *            centroids with `load_factor` (then n must equal nelem and
 *            input_dim == dim + 1)
 *   out      dev [n]
 *   g_out    dev [n]    upstream gradient dL/d(value)
 *   g_theta  dev [n_params] out (overwritten; deterministic reduction order)
 * ------------------------------------------------------------------------ */
int64_t pf_mlp_num_params(int input_dim, int hidden_layers, int width);
int pf_mlp_forward(pf_plan* plan, int input_dim, int hidden_layers, int width, const double* theta, int64_t n,
                   const double* X, double load_factor, double scale, int enforce_positive, double* out,
                   void* stream);
int pf_mlp_backward(pf_plan* plan, int input_dim, int hidden_layers, int width, const double* theta, int64_t n,
                    const double* X, double load_factor, double scale, int enforce_positive,
                    const double* g_out, double* g_theta, void* stream);
/* Batched networks with saved activations -- the kernels of the large-mesh / batched PINN loops (one launch per
 * network for all problems, DMMA, register-resident layer chain).  Problem p uses theta + p * theta_stride and
 * column p of out / g_out, both dev [n][ldb] (ldb >= B).  The forward leaves, per problem, a record of
 * pf_mlp_acts_len(...) doubles (hidden activations and d value / d z) in acts[p]; the backward consumes it instead
 * of recomputing the forward pass (tanh' = 1 - a^2).  acts may be NULL in the forward (nothing saved), out may be
 * NULL when acts is given.  Shapes covered: input_dim <= 3, width <= 24, 1-3 hidden layers; pf_mlp_acts_len
 * returns 0 for other shapes and the batched calls then fail with PF_ERR_ARG.
 *   g_theta  dev [B][gt_stride] out (first n_params entries of every row; deterministic reduction order) */
int64_t pf_mlp_acts_len(int input_dim, int hidden_layers, int width, int64_t n);
int pf_mlp_forward_batched(pf_plan* plan, int input_dim, int hidden_layers, int width, const double* theta,
                           int64_t theta_stride, int64_t B, int64_t n, const double* X, double load_factor,
                           double scale, int enforce_positive, double* out, int64_t ldb, double* acts, void* stream);
int pf_mlp_backward_batched(pf_plan* plan, int input_dim, int hidden_layers, int width, const double* theta,
                            int64_t theta_stride, int64_t B, int64_t n, const double* X, double load_factor,
                            const double* g_out, int64_t ldb, const double* acts, double* g_theta, int64_t gt_stride,
                            void* stream);
/* y[i] = the tanh the hidden layers use (branch-free, absolute error <= 4.5e-16); exposed for its
 * accuracy test.  x, y dev [n]. */
int pf_debug_tanh(int64_t n, const double* x, double* y, void* stream);
/* the same for the table-based tanh of the batched / large-point-set kernels (absolute error <= 4.5e-16,
 * +-1 from |x| = 20, NaN propagates) */
int pf_debug_tanh_table(int64_t n, const double* x, double* y, void* stream);
/* Jacobian rows d value_p / d theta for every point p: jac dev [n][n_params]
 * (what fem/nn_solver.py:91-110 obtains with one reverse pass per row). */
int pf_mlp_param_jacobian(pf_plan* plan, int input_dim, int hidden_layers, int width, const double* theta,
                          int64_t n, const double* X, double load_factor, double scale, int enforce_positive,
                          double* jac, void* stream);

/* ------------------------------------------------------------------------
 * PINN gradient-descent loop on the device (fem/solver.py:252-355: assemble,
 * residual, loss, backward, two Adam steps, BC zeroing, convergence test).
 * One call runs up to max_iterations iterations of `nprob` independent
 * problems that share the mesh plan (problem p uses theta[p], u[p], ...).
 * ------------------------------------------------------------------------ */
typedef struct pf_gd_config {
    int32_t max_iterations;
    int32_t kind;              /* PF_ELEM_* */
    double tolerance;
    double learning_rate_u;
    double learning_rate_theta;
    double alpha_physics;
    double alpha_data;
    double load_factor;
    /* three property slots in Material order: young, area, density
     * (fem/model.py:36-43).  enabled = 0 means a scalar property of value
     * `scale`; otherwise an MLP whose output is multiplied by `scale`. */
    int32_t net_enabled[3];
    int32_t net_input_dim[3];
    int32_t net_hidden_layers[3];
    int32_t net_width[3];
    double net_scale[3];
    int32_t n_measured;
    /* 0: fem/solver.py losses (0.5*sum r^2, converge on ||r|| or loss);
     * 1: legacy fem/nn_solver_gd.py:113-124,:171 (mean r^2, converge on loss only) */
    int32_t loss_mode;
} pf_gd_config;

#define PF_GD_HISTORY_COLS 7 /* iteration, loss_total, loss_physics, loss_data, u_norm, residual_norm, theta_norm */

/*   theta      dev [nprob][n_theta_total] in/out (young|area|density params, enabled nets only)
 *   u          dev [nprob][ndof] in/out  (problem-major: each problem is one small system)
 *   f_ext      dev [ndof] shared loads
 *   meas_dofs  dev int32 [n_measured], meas_vals dev [nprob][n_measured] (NULL when n_measured == 0)
 *   history    dev [nprob][max_iterations][PF_GD_HISTORY_COLS] out, may be NULL
 *   n_iters    dev int32 [nprob] out: iterations executed
 *   converged  dev int32 [nprob] out
 *   reactions  dev [nprob][ndof] out, may be NULL (fem/solver.py:374-380)
 */
int pf_gd_solve(pf_plan* plan, const pf_gd_config* cfg, int64_t nprob, double* theta, double* u,
                const double* f_ext, const int32_t* meas_dofs, const double* meas_vals, double* history,
                int32_t* n_iters, int32_t* converged, double* reactions, void* stream);

/* ------------------------------------------------------------------------
 * Element-sharded meshes over the GPUs of one box (one process per GPU).
 * A rank's local mesh lists its owned nodes first, then its halo nodes, and
 * holds every element incident to an owned node in ascending global element
 * id, so the owned rows of f_int carry the same bits as on a single GPU.
 * The halo rows move between neighbours, and dL/dtheta and the loss scalars are
 * all-reduced, through peer memory over NVLink (default) or NCCL (resolved with
 * dlopen at first use).  The reference
 * has no counterpart: its dense K cannot hold such meshes (fem/assembly.py:19).
 *
 *   pf_comm_unique_id   rank 0 creates the 128-byte NCCL id, the host
 *                       broadcasts it (torch.distributed), every rank calls
 *                       pf_comm_create(world, rank, id, device)
 *   pf_halo_create      peers[n_peers]; send_ptr/recv_ptr [n_peers+1] offsets
 *                       into send_nodes / recv_nodes (host int32, LOCAL node ids)
 *   pf_halo_exchange    x dev [ndof_local][B]: owned rows on the send lists go
 *                       out, halo rows are overwritten with the owners' values
 * ------------------------------------------------------------------------ */
typedef struct pf_comm pf_comm;
typedef struct pf_halo pf_halo;
int pf_comm_available(void);
int pf_comm_unique_id(unsigned char* id128);
int pf_comm_create(int world, int rank, const unsigned char* id128, int device, pf_comm** out);
void pf_comm_destroy(pf_comm* comm);
int pf_comm_allreduce_sum(pf_comm* comm, double* buf, int64_t n, void* stream);
/* Peer-memory transport (NVLink / NVSwitch, CUDA IPC; csrc/pf_peer.cuh).  Optional, collective:
 *   pf_comm_peer_export   allocate this rank's mailbox (2 x world slots of halo_slot_doubles and of
 *                         ar_slot_doubles, double buffered) and return its 64-byte IPC handle; the host
 *                         all-gathers the handles (torch.distributed)
 *   pf_comm_peer_import   handles [world][64] in rank order: map every peer's mailbox.  From here on
 *                         pf_halo_exchange is ONE kernel that stores the halo rows straight into the
 *                         neighbours' mailboxes, pf_comm_allreduce_sum sums in rank order (identical bits on
 *                         every rank) and pf_gd_solve_sharded fuses the all-reduce into its Adam kernel;
 *                         messages larger than a slot keep using NCCL
 *   pf_comm_peer_detach   unmap the peers' mailboxes (NCCL transport from here on); call it on every rank,
 *                         then synchronise the ranks, before pf_comm_destroy frees the own mailbox
 *   pf_comm_peer_check    synchronise `stream`; PF_ERR_CUDA when a wait for a neighbour timed out (60 s) */
int pf_comm_peer_export(pf_comm* comm, int64_t halo_slot_doubles, int64_t ar_slot_doubles, unsigned char* handle64);
int pf_comm_peer_import(pf_comm* comm, const unsigned char* handles);
int pf_comm_peer_enabled(const pf_comm* comm);
void pf_comm_peer_detach(pf_comm* comm);
int pf_comm_peer_check(pf_comm* comm, void* stream);
int pf_halo_create(pf_comm* comm, int dim, int n_peers, const int32_t* peers, const int64_t* send_ptr,
                   const int32_t* send_nodes, const int64_t* recv_ptr, const int32_t* recv_nodes, pf_halo** out);
void pf_halo_destroy(pf_halo* halo);
int pf_halo_exchange(pf_halo* halo, double* x, int64_t B, void* stream);
/* The transport of an exchange is chosen from the largest message (nodes) any rank sends or receives, so that
 * all ranks choose alike: the host reduces that number over the ranks (MAX) and hands it back here. */
int64_t pf_halo_max_message_nodes(const pf_halo* halo);
int pf_halo_set_max_message_nodes(pf_halo* halo, int64_t global_max_nodes);

typedef struct pf_gd_shard {
    pf_halo* halo;              /* exchange lists of this rank */
    int64_t n_owned_nodes;      /* owned nodes come first in the local numbering */
    const uint8_t* elem_owned;  /* dev [nelem_local]: 1 where this rank owns the element (owner of its first node) */
    int64_t nfree_global;       /* free DOFs of the whole mesh (divisor of the legacy mean loss) */
    int32_t n_measured_global;  /* measurements of the whole mesh; cfg->n_measured counts this rank's */
    int32_t reserved;
} pf_gd_shard;

/* pf_gd_solve's large-mesh loop on this rank's local mesh (plan built from it): theta dev [n_theta]
 * (replicated, identical on every rank on return), u_local dev [ndof_local], f_ext_local dev [ndof_local],
 * meas_dofs_local / meas_vals_local: the measurements on this rank's owned DOFs (local DOF ids),
 * history dev [max_iterations][PF_GD_HISTORY_COLS] (identical on every rank), reactions_local dev
 * [ndof_local] (owned rows).  Every rank must call it. */
int pf_gd_solve_sharded(pf_plan* plan, const pf_gd_config* cfg, const pf_gd_shard* shard, double* theta,
                        double* u_local, const double* f_ext_local, const int32_t* meas_dofs_local,
                        const double* meas_vals_local, double* history, int32_t* n_iters, int32_t* converged,
                        double* reactions_local, void* stream);

/* ------------------------------------------------------------------------
 * Linear solves.
 * pf_solve_dense: batched dense solve A x = b by LU with partial pivoting
 *   (np.linalg.solve in fem/core.py:35, fem/solver.py:464; torch.linalg.solve
 *   in fem/nn_solver.py:277).  A dev [nbatch][n][n] row-major (destroyed),
 *   b dev [nbatch][n] (overwritten with x), info dev int32 [nbatch]
 *   (0 ok, k>0: zero pivot at step k -> singular).  n <= 120: one CTA per
 *   system in shared memory; larger n: blocked, the panel factorised by a
 *   thread-block cluster in distributed shared memory.
 * pf_cg_solve: Jacobi-preconditioned conjugate gradients on the free DOFs of
 *   K_t(u) x = rhs, matrix-free, batched (x, rhs dev [ndof][B]; fixed DOFs are
 *   held at 0).  iters_out/resid_out host outputs; synchronises the stream.
 * ------------------------------------------------------------------------ */
int pf_solve_dense(int64_t nbatch, int64_t n, double* A, double* b, int32_t* info, void* stream);
/* pf_solve_spd: Cholesky solve of one symmetric positive definite system -- the damped normal equations
 *   (J^T J + d I) dx = -J^T R of fem/nn_solver.py:273-277 are SPD by construction.  No pivoting, so every step
 *   runs on all SMs (panel solve, DMMA trailing update); 25x faster than the pivoted LU at n = 4096.
 *   A dev [n][n] row-major (lower triangle read, overwritten by L), b dev [n] (overwritten by x),
 *   info dev int32 [1]: 0, or k > 0 when the k-th pivot is not positive. */
int pf_solve_spd(int64_t n, double* A, double* b, int32_t* info, void* stream);
int pf_cg_solve(pf_plan* plan, int kind, int64_t B, const double* u, const double* E, const double* A,
                int mat_batched, const double* rhs, double* x, double rel_tol, int max_iters, double* work,
                int64_t work_len, int32_t* iters_out, double* resid_out, void* stream);
int64_t pf_cg_work_len(const pf_plan* plan, int64_t B);

/* ------------------------------------------------------------------------
 * Gauss-Newton / Levenberg-Marquardt normal equations (fem/nn_solver.py:266-277):
 *   JtJ = J^T J, Jtr = J^T R, damping = 1e-6 * trace(JtJ)/n, JtJ += damping*I.
 * J dev [m][n] row-major, R dev [m]; jtj dev [n][n], jtr dev [n] out;
 * damping_out dev [1] out (may be NULL).  fp64 tensor-core (DMMA) GEMM.
 * ------------------------------------------------------------------------ */
int pf_gn_normal_equations(int64_t m, int64_t n, const double* J, const double* R, double damping_factor,
                           double* jtj, double* jtr, double* damping_out, void* stream);

/* One damped Gauss-Newton / Levenberg-Marquardt step (fem/nn_solver.py:266-277):
 *   dx = -(J^T J + d I)^-1 J^T R,  d = damping_factor * trace(J^T J) / n.
 * With fewer residuals than unknowns (m < n: every inverse problem of the reference, 6 x 1001 on example 10) the
 * step is taken through the m x m dual system (J J^T + d I) y = R, dx = -J^T y -- the same vector, O(m^2 n) work
 * instead of O(n^3).  path: 0 = choose, 1 = force the n x n system, 2 = force the dual.  J dev [m][n] row-major,
 * R dev [m], dx dev [n] out, damping_out dev [1] out (may be NULL), info dev int32 [1] (pivot report of
 * pf_solve_spd). */
int pf_gn_lm_step(int64_t m, int64_t n, const double* J, const double* R, double damping_factor, int path, double* dx,
                  double* damping_out, int32_t* info, void* stream);

/* Stacked Gauss-Newton Jacobian (fem/nn_solver.py:50-135, :223-239), dense row-major
 *   J dev [(nfree + n_meas)][nfree + nE + nA + n_rest]:
 *     rows 0..nfree-1     [ alpha_physics*K_ff | alpha_physics*d f_int[free]/d theta ]
 *     rows nfree..        [ alpha_data*(-1 at the measured free DOF) | 0 ]
 *   vals    dev BSR tangent values from pf_tangent_bsr (B = 1)
 *   jacE    dev [nelem][nE] = d E_e / d theta_E from pf_mlp_param_jacobian (NULL when nE == 0), jacA alike
 *   n_rest  trailing all-zero parameter columns (density: it never enters the physics)
 * replaces the n_free x n_tensors reverse passes of compute_jacobian_blocks. */
int pf_gn_jacobian(pf_plan* plan, int kind, const double* u, const double* E, const double* A, const double* vals,
                   const double* jacE, int64_t nE, const double* jacA, int64_t nA, int64_t n_rest,
                   double alpha_physics, double alpha_data, const int32_t* meas_dofs, int64_t n_meas, double* J,
                   void* stream);

/* fp64 peak probe for the rooflines of the compute-bound kernels (not on any solver path):
 * kind 0 = DFMA pipe, 1 = DMMA (fp64 tensor pipe); *tflops_out (host) = measured TFLOP/s on the
 * current device.  Synchronous, takes a few milliseconds. */
int pf_measure_fp64_peak(int kind, double* tflops_out);

/* ------------------------------------------------------------------------
 * Host-buffer convenience entry point (the end-to-end path): residual of B
 * problems whose u/E/A live in (pinned) host memory, streamed through the
 * GPU in chunks of `chunk` problems with copies overlapped with compute on
 * internal streams; result written to host r.  Host arrays use the same
 * [row][B] layout.  Synchronous.
 * ------------------------------------------------------------------------ */
int pf_residual_host(pf_plan* plan, int kind, int64_t B, const double* u_host, const double* E_host,
                     const double* A_host, const double* f_ext_host, double load_factor, double* r_host,
                     int64_t chunk);

#ifdef __cplusplus
}
#endif
#endif /* PINNFEM_H */
