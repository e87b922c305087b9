// Device-resident PINN gradient descent for meshes that do not fit one CTA's shared memory
// (fem/solver.py:252-355 on ~10^6 elements).
//
// One iteration is a fixed sequence of launches on the caller's stream -- material networks at the
// centroids (pf_mlp_forward), residual + 0.5*sum r^2 (pf_residual, patch/node gather), dL/du = K r
// (pf_tangent_matvec), dL/dE, dL/dA (pf_material_vjp), dL/dtheta (pf_mlp_backward: DMMA), two Adam
// updates with BC zeroing, monitoring norms, history row and convergence test -- with every scalar
// (losses, norms, the converged flag, beta^t) living in device memory.  The host only enqueues; it
// looks at the converged flag every kPoll iterations.  After convergence the remaining enqueued
// iterations are no-ops for the state (the update kernels test the flag).
//
// Element-sharded meshes (pf_gd_solve_sharded): every rank runs the same sequence on its local mesh
// (owned nodes first, then halo nodes; all elements incident to an owned node).  Per iteration the
// ranks swap the halo rows of u and of r (pf_halo_exchange) and all-reduce ONE buffer
// [dL/dtheta | 0.5 sum r^2 | sum data^2 | sum u_free^2]; theta, the losses, the history and the
// convergence flag are then identical on every rank.
#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <vector>

#include "pf_internal.h"
#include "pf_mlp.cuh"
#include "pf_mlp_frag.h"
#include "pf_peer.cuh"

namespace {

constexpr int kPoll = 8;
constexpr int kRedThreads = 256;

// device scalars
enum { S_POW_B1 = 0, S_POW_B2, S_DONE, S_ITERS, S_CONV, S_COUNT };
// reduction buffer: [n_theta gradient entries | L_HALF_SQ | L_DATA_SQ | L_UNORM_SQ]
enum { L_HALF_SQ = 0, L_DATA_SQ, L_UNORM_SQ, L_COUNT };

struct LargeCfg {
    double tolerance, lr_u, lr_t, alpha_p, alpha_d, load_factor, gscale, cdata;
    int legacy, has_meas, n_meas, max_iterations;  // n_meas, nfree: of the whole mesh
    int64_t nfree;
};

// sum_j (m_j - u[dof_j])^2 in a fixed order (one block)
__global__ void __launch_bounds__(kRedThreads) data_loss_kernel(const int32_t* __restrict__ md,
                                                                const double* __restrict__ mv, int n_meas,
                                                                const double* __restrict__ u, double* __restrict__ losses) {
    __shared__ double red[kRedThreads];
    double s = 0.0;
    for (int j = threadIdx.x; j < n_meas; j += kRedThreads) {
        const double d = mv[j] - u[md[j]];
        s += d * d;
    }
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = kRedThreads / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) losses[L_DATA_SQ] = red[0];
}

// Adam on u (torch defaults), data-loss gradient, BC zeroing (solver.py:292-298), partial ||u_free||^2
__global__ void __launch_bounds__(kRedThreads) adam_u_kernel(LargeCfg c, int64_t ndof, const double* __restrict__ gu,
                                                             const double* __restrict__ msum,
                                                             const double* __restrict__ mcnt,
                                                             const uint8_t* __restrict__ dof_free,
                                                             double* __restrict__ u, double* __restrict__ m,
                                                             double* __restrict__ v, const double* __restrict__ sc,
                                                             double* __restrict__ part) {
    __shared__ double red[kRedThreads];
    const bool done = sc[S_DONE] != 0.0;
    const double b1 = 0.9, b2 = 0.999, eps = 1e-8;
    const double pb1 = sc[S_POW_B1] * b1, pb2 = sc[S_POW_B2] * b2;  // beta^t of this iteration
    const double bc1 = 1.0 - pb1, bc2s = sqrt(1.0 - pb2);
    const double step = c.lr_u / bc1;
    double s = 0.0;
    for (int64_t d = (int64_t)blockIdx.x * kRedThreads + threadIdx.x; d < ndof; d += (int64_t)gridDim.x * kRedThreads) {
        double ud = u[d];
        if (!done) {
            double g = c.gscale * gu[d];
            if (c.has_meas && mcnt[d] != 0.0) g += c.cdata * (msum[d] - mcnt[d] * ud);
            const double mm = m[d] + (g - m[d]) * (1.0 - b1);
            const double vv = v[d] * b2 + (1.0 - b2) * g * g;
            m[d] = mm;
            v[d] = vv;
            ud = dof_free[d] ? ud + (-step * mm) / (sqrt(vv) / bc2s + eps) : 0.0;
            u[d] = ud;
        }
        if (dof_free[d]) s += ud * ud;
    }
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = kRedThreads / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) part[blockIdx.x] = red[0];
}

struct TensorList {
    int n;
    int off[3 * 2 * (PF_MLP_MAX_LAYERS + 1)];
    int cnt[3 * 2 * (PF_MLP_MAX_LAYERS + 1)];
};

// One block: Adam on the active parameters, per-tensor norms, u-norm second stage, history row,
// convergence test (solver.py:308-355), iteration bookkeeping.
// Element-sharded runs on the peer-memory transport (ar_n > 0): the all-reduce of the reduction buffer
// g = [dL/dtheta | losses] over the ranks happens HERE, through the peers' mailboxes, before the update --
// the collective and its consumer are one kernel (`losses` aliases the tail of g).
__global__ void __launch_bounds__(1024) adam_theta_finish_kernel(LargeCfg c, int n_active, int n_theta, TensorList tl,
                                                                 double* g, double* __restrict__ theta,
                                                                 double* __restrict__ m, double* __restrict__ v,
                                                                 const double* losses,
                                                                 const double* __restrict__ upart, int n_upart,
                                                                 double* __restrict__ sc, double* __restrict__ history,
                                                                 PfPeerView pv, int ar_n,
                                                                 const double* __restrict__ rpart, int n_rpart) {
    __shared__ double tn[3 * 2 * (PF_MLP_MAX_LAYERS + 1)];
    __shared__ double ured[1024];
    if (ar_n > 0) {
        // this rank's 0.5 sum r^2 and sum u_free^2 (second stage of sq_partial_kernel / adam_u_kernel, fixed tree)
        // go into the reduction buffer, then the buffer is summed over the ranks
        for (int which = 0; which < 2; ++which) {
            const double* part = which ? upart : rpart;
            const int np = which ? n_upart : n_rpart;
            double s = 0.0;
            for (int b = threadIdx.x; b < np; b += blockDim.x) s += part[b];
            ured[threadIdx.x] = s;
            __syncthreads();
            for (int o = blockDim.x / 2; o > 0; o >>= 1) {
                if (threadIdx.x < o) ured[threadIdx.x] += ured[threadIdx.x + o];
                __syncthreads();
            }
            if (threadIdx.x == 0) g[n_theta + (which ? L_UNORM_SQ : L_HALF_SQ)] = which ? ured[0] : 0.5 * ured[0];
            __syncthreads();
        }
        pf_peer_allreduce_block(pv, g, ar_n);  // ends with a block barrier
        n_upart = 0;                           // the u norm now comes from the reduced buffer
    }
    const bool done = sc[S_DONE] != 0.0;
    {   // ||u_free||^2: the partial sums of adam_u_kernel, folded by a fixed tree
        double s = 0.0;
        for (int b = threadIdx.x; b < n_upart; b += blockDim.x) s += upart[b];
        ured[threadIdx.x] = s;
        __syncthreads();
        for (int o = blockDim.x / 2; o > 0; o >>= 1) {
            if (threadIdx.x < o) ured[threadIdx.x] += ured[threadIdx.x + o];
            __syncthreads();
        }
    }
    const double b1 = 0.9, b2 = 0.999, eps = 1e-8;
    const double pb1 = sc[S_POW_B1] * b1, pb2 = sc[S_POW_B2] * b2;
    const double bc1 = 1.0 - pb1, bc2s = sqrt(1.0 - pb2);
    const double step = c.lr_t / bc1;
    if (!done) {
        for (int q = threadIdx.x; q < n_active; q += blockDim.x) {
            const double gq = c.gscale * g[q];
            const double mm = m[q] + (gq - m[q]) * (1.0 - b1);
            const double vv = v[q] * b2 + (1.0 - b2) * gq * gq;
            m[q] = mm;
            v[q] = vv;
            theta[q] += (-step * mm) / (sqrt(vv) / bc2s + eps);
        }
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    for (int t = warp; t < tl.n; t += nwarp) {
        double s = 0.0;
        for (int i = lane; i < tl.cnt[t]; i += 32) s = fma(theta[tl.off[t] + i], theta[tl.off[t] + i], s);
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) tn[t] = sqrt(s);
    }
    __syncthreads();
    if (threadIdx.x == 0 && !done) {
        double tsum = 0.0;
        for (int t = 0; t < tl.n; ++t) tsum += tn[t];
        const double un = n_upart ? ured[0] : losses[L_UNORM_SQ];  // sharded runs pre-reduce it into the buffer
        const int it = (int)sc[S_ITERS];  // 0-based index of this iteration
        const double s2 = 2.0 * losses[L_HALF_SQ];
        const double loss_p = c.legacy ? s2 / (double)c.nfree : 0.5 * s2;
        const double loss_d = c.has_meas ? losses[L_DATA_SQ] / c.n_meas : 0.0;
        const double loss = c.alpha_p * loss_p + (c.has_meas ? c.alpha_d * loss_d : 0.0);
        const double rn = sqrt(s2);
        if (history) {
            double* h = history + (int64_t)it * PF_GD_HISTORY_COLS;
            h[0] = (double)(it + 1);
            h[1] = loss;
            h[2] = loss_p;
            h[3] = c.n_meas > 0 ? loss_d : 0.0;
            h[4] = sqrt(un);
            h[5] = rn;
            h[6] = tsum;
        }
        sc[S_POW_B1] = pb1;
        sc[S_POW_B2] = pb2;
        sc[S_ITERS] = (double)(it + 1);
        int conv = 0;
        if (it > 10) {  // solver.py:341-355 (legacy nn_solver_gd.py:171: loss only)
            if (!c.legacy && rn < c.tolerance) conv = 1;
            else if (!isnan(loss) && loss < c.tolerance) conv = 1;
        }
        if (conv) {
            sc[S_CONV] = 1.0;
            sc[S_DONE] = 1.0;
        } else if (it + 1 >= c.max_iterations) {
            sc[S_DONE] = 1.0;
        }
    }
}

// part[block] = sum of x^2 over the block's grid-stride slice (fixed order)
__global__ void __launch_bounds__(kRedThreads) sq_partial_kernel(const double* __restrict__ x, int64_t n,
                                                                 double* __restrict__ part) {
    __shared__ double red[kRedThreads];
    double s = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * kRedThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kRedThreads)
        s += x[i] * x[i];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = kRedThreads / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) part[blockIdx.x] = red[0];
}

// out = scale * sum(part[0..n)): one block, strided partial sums folded by a fixed tree
__global__ void __launch_bounds__(kRedThreads) sum_partials_kernel(const double* __restrict__ part, int n, double scale,
                                                                   double* __restrict__ out) {
    __shared__ double red[kRedThreads];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += kRedThreads) s += part[i];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = kRedThreads / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = scale * red[0];
}

// elements on a partition interface are local to both ranks: only the owner contributes to dL/dtheta
__global__ void mask_elements_kernel(const uint8_t* __restrict__ owned, int64_t n, double* __restrict__ gE,
                                     double* __restrict__ gA) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && !owned[i]) {
        gE[i] = 0.0;
        gA[i] = 0.0;
    }
}

__global__ void fill_kernel(double* __restrict__ x, int64_t n, double v) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = v;
}

// reactions = f_int - lambda f_ext on fixed DOFs, 0 on free DOFs (solver.py:374-380)
__global__ void reactions_kernel(const double* __restrict__ f, const double* __restrict__ fext, double lam,
                                 const uint8_t* __restrict__ dof_free, int64_t ndof, double* __restrict__ out) {
    const int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (d < ndof) out[d] = dof_free[d] ? 0.0 : __dsub_rn(f[d], __dmul_rn(lam, fext[d]));
}

// internal stream joined to the caller's stream on both ends; graph objects of one problem
struct WorkStream {
    cudaStream_t st = nullptr, caller = nullptr;
    cudaEvent_t ev = nullptr;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    int open(cudaStream_t caller_st) {
        caller = caller_st;
        PF_CUDA_CHECK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        PF_CUDA_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        PF_CUDA_CHECK(cudaEventRecord(ev, caller));
        PF_CUDA_CHECK(cudaStreamWaitEvent(st, ev, 0));
        return PF_OK;
    }
    void drop_graph() {
        if (exec) cudaGraphExecDestroy(exec);
        if (graph) cudaGraphDestroy(graph);
        exec = nullptr;
        graph = nullptr;
    }
    ~WorkStream() {
        drop_graph();
        if (st) {
            cudaStreamSynchronize(st);  // all paths (errors too) leave nothing in flight on freed buffers
            if (ev && cudaEventRecord(ev, st) == cudaSuccess) cudaStreamWaitEvent(caller, ev, 0);
            cudaStreamDestroy(st);
        }
        if (ev) cudaEventDestroy(ev);
    }
};

// Scratch of one solve, from the device's stream-ordered pool: cudaMalloc / cudaFree of ~15 buffers cost
// more than 100 iterations of the loop when the process already holds tens of GB (measured inside bench.py:
// 2.8 ms per iteration instead of 1.15), the pool hands the same blocks back on the next call.
struct DevBuf {
    std::vector<void*> ptrs;
    cudaStream_t st = nullptr;
    ~DevBuf() {
        for (void* p : ptrs) cudaFreeAsync(p, st);
    }
    int open(cudaStream_t s) {
        st = s;
        pf_keep_pool_cached();
        return PF_OK;
    }
    template <typename T>
    int alloc(T** out, size_t count) {
        void* p = nullptr;
        PF_CUDA_CHECK(cudaMallocAsync(&p, std::max<size_t>(count, 1) * sizeof(T), st));
        ptrs.push_back(p);
        *out = static_cast<T*>(p);
        return PF_OK;
    }
};


// ---------------------------------------------------------------------------------------------------------------
// Batched problems on one large mesh (BASELINE config 5: ~10^6 elements x many independent inverse problems).
// A group of G problems advances together: every array carries the problem index last ([ndof][ld], [nelem][ld]),
// so the residual, K r and the material VJP run on the patch-staged batch kernels (pf_patch.cu) and the material
// networks on the batched fragment kernels (pf_mlp_frag.cu, per-problem theta).  Adam moments, beta^t, losses,
// history rows and the converged flag are per problem; a problem that has converged stops changing while the rest
// of its group runs on.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kURowBlocks = 128;  // row blocks of the batched u update (partial ||u||^2 rows per problem)

// dst[d][ld] column p = src[p][n] (rows of the caller's problem-major array), zero in the padding columns
__global__ void __launch_bounds__(1024) to_batch_layout_kernel(const double* __restrict__ src, int64_t n, int B, int64_t ld,
                                                               double* __restrict__ dst) {
    __shared__ double t[32][33];
    const int64_t d0 = (int64_t)blockIdx.x * 32;
    const int p0 = blockIdx.y * 32;
    {
        const int p = p0 + threadIdx.y;
        const int64_t d = d0 + threadIdx.x;
        t[threadIdx.y][threadIdx.x] = (p < B && d < n) ? src[(int64_t)p * n + d] : 0.0;
    }
    __syncthreads();
    const int p = p0 + threadIdx.x;
    const int64_t d = d0 + threadIdx.y;
    if (p < ld && d < n) dst[d * ld + p] = t[threadIdx.x][threadIdx.y];
}

// dst[p][n] = src[d][ld] column p.  REACT: dst = reactions = f_int - lambda f_ext on fixed DOFs, 0 on free DOFs
template <bool REACT>
__global__ void __launch_bounds__(1024) from_batch_layout_kernel(const double* __restrict__ src, int64_t n, int B, int64_t ld,
                                                                 const double* __restrict__ fext, double lam,
                                                                 const uint8_t* __restrict__ dof_free,
                                                                 double* __restrict__ dst) {
    __shared__ double t[32][33];
    const int64_t d0 = (int64_t)blockIdx.x * 32;
    const int p0 = blockIdx.y * 32;
    {
        const int p = p0 + threadIdx.x;
        const int64_t d = d0 + threadIdx.y;
        double v = 0.0;
        if (p < B && d < n) {
            v = src[d * ld + p];
            if (REACT) v = dof_free[d] ? 0.0 : __dsub_rn(v, __dmul_rn(lam, fext[d]));
        }
        t[threadIdx.y][threadIdx.x] = v;
    }
    __syncthreads();
    const int p = p0 + threadIdx.y;
    const int64_t d = d0 + threadIdx.x;
    if (p < B && d < n) dst[(int64_t)p * n + d] = t[threadIdx.x][threadIdx.y];
}

// msum[k][p] = sum of problem p's targets on the k-th distinct measured DOF, in ascending measurement index
__global__ void meas_sum_kernel(const int32_t* __restrict__ slot_ptr, const int32_t* __restrict__ slot_js, int nslot,
                                const double* __restrict__ mv, int n_meas, int B, int64_t ld, double* __restrict__ msum) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y;
    if (p >= B || k >= nslot) return;
    double s = 0.0;
    for (int q = slot_ptr[k]; q < slot_ptr[k + 1]; ++q) s += mv[(int64_t)p * n_meas + slot_js[q]];
    msum[(int64_t)k * ld + p] = s;
}

// dl[p] = sum_j (m[p][j] - u[dof_j][p])^2, one block per problem, fixed order
__global__ void __launch_bounds__(kRedThreads) data_loss_batched_kernel(const int32_t* __restrict__ md,
                                                                        const double* __restrict__ mv, int n_meas,
                                                                        const double* __restrict__ u, int64_t ld,
                                                                        double* __restrict__ dl) {
    __shared__ double red[kRedThreads];
    const int p = blockIdx.x;
    double s = 0.0;
    for (int j = threadIdx.x; j < n_meas; j += kRedThreads) {
        const double d = mv[(int64_t)p * n_meas + j] - u[(int64_t)md[j] * ld + p];
        s += d * d;
    }
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = kRedThreads / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) dl[p] = red[0];
}

// Adam on u for every problem of the group (same arithmetic as adam_u_kernel): thread = (32 problems x 8 rows),
// block rb owns the rows [rb * rows_per_block, ...).  upart[rb][p] = the block's share of ||u_free||^2.
__global__ void __launch_bounds__(256) adam_u_batched_kernel(LargeCfg c, int64_t ndof, int64_t ld, int B,
                                                             int64_t rows_per_block, const double* __restrict__ gu,
                                                             const int32_t* __restrict__ slot_of_dof,
                                                             const double* __restrict__ msum,
                                                             const double* __restrict__ mcnt,
                                                             const uint8_t* __restrict__ dof_free, double* __restrict__ u,
                                                             double* __restrict__ m, double* __restrict__ v,
                                                             const double* __restrict__ sc, double* __restrict__ upart) {
    __shared__ double red[8][33];
    const int p = blockIdx.y * 32 + threadIdx.x;
    const bool live = p < B;
    const double* scp = sc + (int64_t)(live ? p : 0) * S_COUNT;
    const bool done = scp[S_DONE] != 0.0;
    const double b1 = 0.9, b2 = 0.999, eps = 1e-8;
    const double pb1 = scp[S_POW_B1] * b1, pb2 = scp[S_POW_B2] * b2;
    const double bc1 = 1.0 - pb1, bc2s = sqrt(1.0 - pb2);
    const double step = c.lr_u / bc1;
    const int64_t d_end = min(ndof, (int64_t)(blockIdx.x + 1) * rows_per_block);
    double s = 0.0;
    if (live) {
        for (int64_t d = (int64_t)blockIdx.x * rows_per_block + threadIdx.y; d < d_end; d += 8) {
            const int64_t i = d * ld + p;
            double ud = u[i];
            const bool fr = dof_free[d] != 0;
            if (!done) {
                double g = c.gscale * gu[i];
                if (c.has_meas) {
                    const int k = slot_of_dof[d];
                    if (k >= 0) g += c.cdata * (msum[(int64_t)k * ld + p] - mcnt[k] * ud);
                }
                const double mm = m[i] + (g - m[i]) * (1.0 - b1);
                const double vv = v[i] * b2 + (1.0 - b2) * g * g;
                m[i] = mm;
                v[i] = vv;
                ud = fr ? ud + (-step * mm) / (sqrt(vv) / bc2s + eps) : 0.0;
                u[i] = ud;
            }
            if (fr) s += ud * ud;
        }
    }
    red[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && live) {
        double t = 0.0;
        for (int y = 0; y < 8; ++y) t += red[y][threadIdx.x];
        upart[(int64_t)blockIdx.x * ld + p] = t;
    }
}

// One block per problem: Adam on its active parameters, per-tensor norms, history row, convergence test
// (solver.py:308-355) -- adam_theta_finish_kernel with per-problem state.
__global__ void __launch_bounds__(256) adam_theta_finish_batched_kernel(LargeCfg c, int n_active, int n_theta, TensorList tl,
                                                                        const double* __restrict__ g, double* __restrict__ theta,
                                                                        double* __restrict__ m, double* __restrict__ v,
                                                                        const double* __restrict__ half_sq,
                                                                        const double* __restrict__ dl,
                                                                        const double* __restrict__ upart, int n_upart,
                                                                        int64_t ld, double* __restrict__ sc,
                                                                        double* __restrict__ history, int64_t hist_stride) {
    __shared__ double tn[3 * 2 * (PF_MLP_MAX_LAYERS + 1)];
    __shared__ double ured[256];
    const int p = blockIdx.x;
    g += (int64_t)p * n_theta;
    theta += (int64_t)p * n_theta;
    m += (int64_t)p * n_theta;
    v += (int64_t)p * n_theta;
    sc += (int64_t)p * S_COUNT;
    const bool done = sc[S_DONE] != 0.0;
    {
        double s = 0.0;
        for (int b = threadIdx.x; b < n_upart; b += blockDim.x) s += upart[(int64_t)b * ld + p];
        ured[threadIdx.x] = s;
        __syncthreads();
        for (int o = blockDim.x / 2; o > 0; o >>= 1) {
            if (threadIdx.x < o) ured[threadIdx.x] += ured[threadIdx.x + o];
            __syncthreads();
        }
    }
    const double b1 = 0.9, b2 = 0.999, eps = 1e-8;
    const double pb1 = sc[S_POW_B1] * b1, pb2 = sc[S_POW_B2] * b2;
    const double bc1 = 1.0 - pb1, bc2s = sqrt(1.0 - pb2);
    const double step = c.lr_t / bc1;
    if (!done) {
        for (int q = threadIdx.x; q < n_active; q += blockDim.x) {
            const double gq = c.gscale * g[q];
            const double mm = m[q] + (gq - m[q]) * (1.0 - b1);
            const double vv = v[q] * b2 + (1.0 - b2) * gq * gq;
            m[q] = mm;
            v[q] = vv;
            theta[q] += (-step * mm) / (sqrt(vv) / bc2s + eps);
        }
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    for (int t = warp; t < tl.n; t += nwarp) {
        double s = 0.0;
        for (int i = lane; i < tl.cnt[t]; i += 32) s = fma(theta[tl.off[t] + i], theta[tl.off[t] + i], s);
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) tn[t] = sqrt(s);
    }
    __syncthreads();
    if (threadIdx.x == 0 && !done) {
        double tsum = 0.0;
        for (int t = 0; t < tl.n; ++t) tsum += tn[t];
        const int it = (int)sc[S_ITERS];
        const double s2 = 2.0 * half_sq[p];
        const double loss_p = c.legacy ? s2 / (double)c.nfree : 0.5 * s2;
        const double loss_d = c.has_meas ? dl[p] / c.n_meas : 0.0;
        const double loss = c.alpha_p * loss_p + (c.has_meas ? c.alpha_d * loss_d : 0.0);
        const double rn = sqrt(s2);
        if (history) {
            double* h = history + (int64_t)p * hist_stride + (int64_t)it * PF_GD_HISTORY_COLS;
            h[0] = (double)(it + 1);
            h[1] = loss;
            h[2] = loss_p;
            h[3] = c.n_meas > 0 ? loss_d : 0.0;
            h[4] = sqrt(ured[0]);
            h[5] = rn;
            h[6] = tsum;
        }
        sc[S_POW_B1] = pb1;
        sc[S_POW_B2] = pb2;
        sc[S_ITERS] = (double)(it + 1);
        int conv = 0;
        if (it > 10) {
            if (!c.legacy && rn < c.tolerance) conv = 1;
            else if (!isnan(loss) && loss < c.tolerance) conv = 1;
        }
        if (conv) {
            sc[S_CONV] = 1.0;
            sc[S_DONE] = 1.0;
        } else if (it + 1 >= c.max_iterations) {
            sc[S_DONE] = 1.0;
        }
    }
}

__global__ void finish_flags_kernel(const double* __restrict__ sc, int B, int32_t* __restrict__ n_iters,
                                    int32_t* __restrict__ converged, int* __restrict__ n_done) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= B) return;
    n_iters[p] = (int32_t)sc[(int64_t)p * S_COUNT + S_ITERS];
    converged[p] = (int32_t)sc[(int64_t)p * S_COUNT + S_CONV];
    if (sc[(int64_t)p * S_COUNT + S_DONE] != 0.0) atomicAdd(n_done, 1);
}

struct BatchNets {
    PfMlpDesc desc[3];
    int theta_off[3];
    int ntheta, n_active;
    TensorList tl;
};

// the shapes / sizes the batched loop covers
bool batched_capable(pf_plan* plan, const pf_gd_config* cfg, const BatchNets& bn) {
    static const int off = getenv("PF_GD_NO_BATCH") ? atoi(getenv("PF_GD_NO_BATCH")) : 0;
    if (off) return false;
    for (int k = 0; k < 2; ++k)
        if (cfg->net_enabled[k] && !pf_mlp_frag_supported(bn.desc[k], true)) return false;
    return plan->nelem > 0;
}

int gd_solve_batched(pf_plan* plan, const pf_gd_config* cfg, const BatchNets& bn, const LargeCfg& c, int64_t nprob,
                     double* theta_all, double* u_all, const double* f_ext, const int32_t* meas_dofs,
                     const double* meas_vals_all, double* history_all, int32_t* n_iters, int32_t* converged,
                     double* reactions_all, WorkStream& ws, const std::vector<int32_t>& h_md, int no_graph) {
    cudaStream_t st = ws.st;
    const int64_t ndof = plan->ndof, nelem = plan->nelem;
    const int n_meas = cfg->n_measured, ntheta = bn.ntheta;
    const bool has_meas = c.has_meas != 0;
    const bool net[2] = {cfg->net_enabled[0] != 0, cfg->net_enabled[1] != 0};
    const bool mat_batched = net[0] || net[1];
    int64_t alen[2] = {0, 0};
    for (int k = 0; k < 2; ++k)
        if (net[k]) alen[k] = pf_mlp_frag_acts_len(bn.desc[k], nelem);

    // group size: PF_GD_BATCH (default 64 problems), halved until the group's arrays fit the free memory
    static const int env_group = getenv("PF_GD_BATCH") ? atoi(getenv("PF_GD_BATCH")) : 0;
    int64_t G = std::min<int64_t>(nprob, env_group > 0 ? env_group : 64);
    const double per_problem = 8.0 * (5.0 * ndof + (mat_batched ? 4.0 : 0.0) * nelem + alen[0] + alen[1] + 4.0 * ntheta) +
                               8.0 * std::max<int64_t>(plan->nnode, (int64_t)plan->patches.size());
    size_t free_b = 0, total_b = 0;
    PF_CUDA_CHECK(cudaMemGetInfo(&free_b, &total_b));
    while (G > 1 && per_problem * ((G + 15) / 16 * 16) > 0.85 * (double)free_b) G = (G + 1) / 2;
    PF_REQUIRE(per_problem * G <= 0.95 * (double)free_b, "not enough device memory for one problem of this size");
    const int64_t ld = G >= 16 ? (G + 15) / 16 * 16 : (G + 1) / 2 * 2;  // full 16-problem chunks for the patch kernels

    // distinct measured DOFs -> slots; the targets of a DOF measured several times are summed in measurement order
    std::vector<int32_t> h_slot_of(has_meas ? ndof : 0, -1), h_slot_ptr, h_slot_js;
    std::vector<double> h_cnt;
    int nslot = 0;
    if (has_meas) {
        std::vector<int32_t> first;
        for (int j = 0; j < n_meas; ++j)
            if (h_slot_of[h_md[j]] < 0) h_slot_of[h_md[j]] = nslot++;
        h_cnt.assign(nslot, 0.0);
        h_slot_ptr.assign(nslot + 1, 0);
        for (int j = 0; j < n_meas; ++j) h_slot_ptr[h_slot_of[h_md[j]] + 1]++;
        for (int k = 0; k < nslot; ++k) h_slot_ptr[k + 1] += h_slot_ptr[k];
        h_slot_js.resize(n_meas);
        std::vector<int32_t> fill(h_slot_ptr.begin(), h_slot_ptr.end() - 1);
        for (int j = 0; j < n_meas; ++j) {
            const int k = h_slot_of[h_md[j]];
            h_slot_js[fill[k]++] = j;
            h_cnt[k] += 1.0;
        }
    }

    DevBuf buf;
    buf.open(st);
    int rc;
    double *U, *r, *gu, *mu, *vu, *E = nullptr, *A = nullptr, *gE = nullptr, *gA = nullptr, *mt, *vt, *gt, *sc, *hs, *dl, *upart;
    double *acts[2] = {nullptr, nullptr}, *fpart = nullptr, *msum = nullptr, *mcnt = nullptr;
    int32_t *slot_of = nullptr, *slot_ptr = nullptr, *slot_js = nullptr;
    int* n_done;
    if ((rc = buf.alloc(&U, ndof * ld)) || (rc = buf.alloc(&r, ndof * ld)) || (rc = buf.alloc(&gu, ndof * ld)) ||
        (rc = buf.alloc(&mu, ndof * ld)) || (rc = buf.alloc(&vu, ndof * ld)) || (rc = buf.alloc(&mt, G * ntheta)) ||
        (rc = buf.alloc(&vt, G * ntheta)) || (rc = buf.alloc(&gt, G * ntheta)) || (rc = buf.alloc(&sc, G * S_COUNT)) ||
        (rc = buf.alloc(&hs, ld)) || (rc = buf.alloc(&dl, ld)) || (rc = buf.alloc(&upart, kURowBlocks * ld)) ||
        (rc = buf.alloc(&n_done, 1)))
        return rc;
    const int64_t mat_len = mat_batched ? nelem * ld : nelem;
    if ((rc = buf.alloc(&E, mat_len)) || (rc = buf.alloc(&A, mat_len))) return rc;
    if (bn.n_active && ((rc = buf.alloc(&gE, nelem * ld)) || (rc = buf.alloc(&gA, nelem * ld)))) return rc;
    size_t part_len = 0;
    for (int k = 0; k < 2; ++k) {
        if (!net[k]) continue;
        if ((rc = buf.alloc(&acts[k], (size_t)(alen[k] * G)))) return rc;
        part_len = std::max(part_len, (size_t)G * pf_mlp_frag_chunks(bn.desc[k], nelem, G, plan->sm_count) * bn.desc[k].n_params);
    }
    if (part_len && (rc = buf.alloc(&fpart, part_len))) return rc;
    if (has_meas) {
        if ((rc = buf.alloc(&msum, (size_t)nslot * ld)) || (rc = buf.alloc(&mcnt, nslot)) || (rc = buf.alloc(&slot_of, ndof)) ||
            (rc = buf.alloc(&slot_ptr, nslot + 1)) || (rc = buf.alloc(&slot_js, n_meas)))
            return rc;
        PF_CUDA_CHECK(cudaMemcpyAsync(mcnt, h_cnt.data(), nslot * sizeof(double), cudaMemcpyHostToDevice, st));
        PF_CUDA_CHECK(cudaMemcpyAsync(slot_of, h_slot_of.data(), ndof * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        PF_CUDA_CHECK(cudaMemcpyAsync(slot_ptr, h_slot_ptr.data(), (nslot + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        PF_CUDA_CHECK(cudaMemcpyAsync(slot_js, h_slot_js.data(), n_meas * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        PF_CUDA_CHECK(cudaStreamSynchronize(st));  // host vectors stay alive anyway; keeps the ordering simple
    }
    // residual partial sums live in the plan workspace: size it once, before anything is captured
    if ((rc = pf_plan_reserve_work(plan, (size_t)std::max<int64_t>(plan->nnode, (int64_t)plan->patches.size()) * ld * sizeof(double))))
        return rc;

    const int64_t rows_per_block = (ndof + kURowBlocks - 1) / kURowBlocks;
    const int n_upart = (int)((ndof + rows_per_block - 1) / rows_per_block);
    const int64_t hist_stride = (int64_t)std::max(cfg->max_iterations, 1) * PF_GD_HISTORY_COLS;
    const dim3 tgrid_dof((unsigned)((ndof + 31) / 32), (unsigned)((ld + 31) / 32)), tblock(32, 32);
    std::vector<double> h_sc;

    for (int64_t p0 = 0; p0 < nprob; p0 += G) {
        const int B = (int)std::min<int64_t>(G, nprob - p0);
        double* theta = theta_all ? theta_all + p0 * ntheta : nullptr;
        const double* mv = has_meas ? meas_vals_all + p0 * n_meas : nullptr;
        double* history = history_all ? history_all + p0 * hist_stride : nullptr;
        to_batch_layout_kernel<<<tgrid_dof, tblock, 0, st>>>(u_all + p0 * ndof, ndof, B, ld, U);
        PF_CUDA_CHECK(cudaMemsetAsync(mu, 0, ndof * ld * sizeof(double), st));
        PF_CUDA_CHECK(cudaMemsetAsync(vu, 0, ndof * ld * sizeof(double), st));
        if (ntheta) {
            PF_CUDA_CHECK(cudaMemsetAsync(mt, 0, G * ntheta * sizeof(double), st));
            PF_CUDA_CHECK(cudaMemsetAsync(vt, 0, G * ntheta * sizeof(double), st));
            PF_CUDA_CHECK(cudaMemsetAsync(gt, 0, G * ntheta * sizeof(double), st));
        }
        PF_CUDA_CHECK(cudaMemsetAsync(dl, 0, ld * sizeof(double), st));
        h_sc.assign((size_t)G * S_COUNT, 0.0);
        for (int p = 0; p < G; ++p) {
            h_sc[p * S_COUNT + S_POW_B1] = h_sc[p * S_COUNT + S_POW_B2] = 1.0;
            if (cfg->max_iterations == 0 || p >= B) h_sc[p * S_COUNT + S_DONE] = 1.0;
        }
        PF_CUDA_CHECK(cudaMemcpyAsync(sc, h_sc.data(), h_sc.size() * sizeof(double), cudaMemcpyHostToDevice, st));
        // scalar properties: one constant array for the whole solve (the padding columns of a network's array too)
        for (int k = 0; k < 2; ++k) {
            double* dst = k == 0 ? E : A;
            fill_kernel<<<(unsigned)((mat_len + 255) / 256), 256, 0, st>>>(dst, mat_len, cfg->net_scale[k]);
        }
        if (has_meas)
            meas_sum_kernel<<<dim3((unsigned)((B + 63) / 64), (unsigned)nslot), 64, 0, st>>>(slot_ptr, slot_js, nslot, mv, n_meas,
                                                                                             B, ld, msum);
        PF_CUDA_CHECK(cudaGetLastError());
        PF_CUDA_CHECK(cudaStreamSynchronize(st));  // h_sc is reused by the next group

        auto materials = [&](bool save) -> int {
            for (int k = 0; k < 2; ++k) {
                if (!net[k]) continue;
                int e = pf_mlp_frag_forward(bn.desc[k], theta + bn.theta_off[k], ntheta, B, nelem, nullptr, plan->d_centroid,
                                            cfg->load_factor, cfg->net_scale[k], 1, k == 0 ? E : A, ld, save ? acts[k] : nullptr,
                                            alen[k], plan->sm_count, st);
                if (e) return e;
            }
            return PF_OK;
        };
        auto iteration = [&]() -> int {
            if ((rc = materials(true))) return rc;
            if ((rc = pf_residual(plan, PF_ELEM_LINEAR, ld, U, E, A, mat_batched ? 1 : 0, nullptr, f_ext, 0, cfg->load_factor, r,
                                  hs, nullptr, st)))
                return rc;
            if ((rc = pf_tangent_matvec(plan, PF_ELEM_LINEAR, ld, nullptr, E, A, mat_batched ? 1 : 0, r, gu, st))) return rc;
            if (bn.n_active) {
                if ((rc = pf_material_vjp(plan, PF_ELEM_LINEAR, ld, U, E, A, mat_batched ? 1 : 0, r, gE, gA, st))) return rc;
                for (int k = 0; k < 2; ++k)
                    if (net[k] && (rc = pf_mlp_frag_backward(bn.desc[k], theta + bn.theta_off[k], ntheta, B, nelem, nullptr,
                                                             plan->d_centroid, cfg->load_factor, k == 0 ? gE : gA, ld, acts[k],
                                                             alen[k], fpart, gt + bn.theta_off[k], ntheta, plan->sm_count, st)))
                        return rc;
            }
            if (has_meas) data_loss_batched_kernel<<<B, kRedThreads, 0, st>>>(meas_dofs, mv, n_meas, U, ld, dl);
            adam_u_batched_kernel<<<dim3((unsigned)n_upart, (unsigned)((B + 31) / 32)), dim3(32, 8), 0, st>>>(
                c, ndof, ld, B, rows_per_block, gu, slot_of, msum, mcnt, plan->d_dof_free, U, mu, vu, sc, upart);
            adam_theta_finish_batched_kernel<<<B, 256, 0, st>>>(c, bn.n_active, ntheta, bn.tl, gt, theta, mt, vt, hs, dl, upart,
                                                               n_upart, ld, sc, history, hist_stride);
            PF_CUDA_CHECK(cudaGetLastError());
            return PF_OK;
        };
        bool finished = cfg->max_iterations == 0;
        ws.drop_graph();
        int h_done = 0;
        for (int it = 0; it < cfg->max_iterations && !finished; ++it) {
            if (it == 1 && !no_graph && cfg->max_iterations > 2) {
                if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
                    const int crc = iteration();
                    const cudaError_t ce = cudaStreamEndCapture(st, &ws.graph);
                    if (crc != PF_OK || ce != cudaSuccess || !ws.graph ||
                        cudaGraphInstantiate(&ws.exec, ws.graph, 0) != cudaSuccess) {
                        ws.drop_graph();
                        cudaGetLastError();
                    }
                } else {
                    cudaGetLastError();
                }
            }
            if (ws.exec) {
                PF_CUDA_CHECK(cudaGraphLaunch(ws.exec, st));
            } else if ((rc = iteration())) {
                return rc;
            }
            if ((it + 1) % kPoll == 0 || it + 1 == cfg->max_iterations) {
                PF_CUDA_CHECK(cudaMemsetAsync(n_done, 0, sizeof(int), st));
                finish_flags_kernel<<<(B + 127) / 128, 128, 0, st>>>(sc, B, n_iters + p0, converged + p0, n_done);
                PF_CUDA_CHECK(cudaMemcpyAsync(&h_done, n_done, sizeof(int), cudaMemcpyDeviceToHost, st));
                PF_CUDA_CHECK(cudaStreamSynchronize(st));
                finished = h_done >= B;
            }
        }
        PF_CUDA_CHECK(cudaMemsetAsync(n_done, 0, sizeof(int), st));
        finish_flags_kernel<<<(B + 127) / 128, 128, 0, st>>>(sc, B, n_iters + p0, converged + p0, n_done);
        from_batch_layout_kernel<false><<<tgrid_dof, tblock, 0, st>>>(U, ndof, B, ld, nullptr, 0.0, nullptr, u_all + p0 * ndof);
        if (reactions_all) {
            if ((rc = materials(false))) return rc;
            if ((rc = pf_residual(plan, PF_ELEM_LINEAR, ld, U, E, A, mat_batched ? 1 : 0, r, nullptr, 0, 0.0, nullptr, nullptr,
                                  nullptr, st)))
                return rc;
            from_batch_layout_kernel<true><<<tgrid_dof, tblock, 0, st>>>(r, ndof, B, ld, f_ext, cfg->load_factor,
                                                                        plan->d_dof_free, reactions_all + p0 * ndof);
        }
        PF_CUDA_CHECK(cudaGetLastError());
        PF_CUDA_CHECK(cudaStreamSynchronize(st));
    }
    return PF_OK;
}

}  // namespace

// One problem at a time (or, for several problems on shapes the batched kernels cover, groups of problems through
// gd_solve_batched); called by pf_gd_solve when the single-CTA kernel does not fit (sh == NULL)
// and by pf_gd_solve_sharded with this rank's local mesh.
int pf_gd_solve_large(pf_plan* plan, const pf_gd_config* cfg, int64_t nprob, double* theta_all, double* u_all,
                      const double* f_ext, const int32_t* meas_dofs, const double* meas_vals_all, double* history_all,
                      int32_t* n_iters, int32_t* converged, double* reactions_all, cudaStream_t caller_st,
                      const pf_gd_shard* sh) {
    // The loop runs on its own stream (ordered after the caller's work, and the caller's stream waits for it
    // at the end): the caller's stream is usually the legacy default stream, which cannot be captured.
    WorkStream ws;
    int wrc = ws.open(caller_st);
    if (wrc) return wrc;
    cudaStream_t st = ws.st;
    const int64_t ndof = plan->ndof, nelem = plan->nelem;
    const int64_t nd_own = sh ? sh->n_owned_nodes * plan->dim : ndof;  // rows this rank updates
    pf_comm* comm = sh ? pf_halo_comm(sh->halo) : nullptr;
    PfPeerView pv;  // peer-memory transport: the all-reduce is fused into adam_theta_finish_kernel
    const bool peer = comm && pf_comm_world(comm) > 1 && pf_comm_peer_view(comm, &pv);
    PF_REQUIRE(!sh || (sh->halo && nprob == 1 && sh->n_owned_nodes >= 0 && sh->n_owned_nodes <= plan->nnode),
               "pf_gd_solve_sharded: bad shard description");
    PfMlpDesc desc[3];
    int theta_off[3] = {0, 0, 0}, ntheta = 0;
    TensorList tl{};
    for (int k = 0; k < 3; ++k) {
        theta_off[k] = ntheta;
        if (!cfg->net_enabled[k]) continue;
        int rc = pf_mlp_make_desc(cfg->net_input_dim[k], cfg->net_hidden_layers[k], cfg->net_width[k], &desc[k]);
        if (rc) return rc;
        PF_REQUIRE(desc[k].in_dim == plan->dim + 1,
                   "network input_dim %d does not match [load_factor, centroid] = %d inputs (fem/properties.py:116-125)",
                   desc[k].in_dim, plan->dim + 1);
        for (int l = 0; l <= desc[k].L; ++l) {
            const int in = l == 0 ? desc[k].in_dim : desc[k].w;
            tl.off[tl.n] = ntheta + desc[k].w_off[l];
            tl.cnt[tl.n++] = l == desc[k].L ? desc[k].w : desc[k].w * in;
            tl.off[tl.n] = ntheta + desc[k].b_off[l];
            tl.cnt[tl.n++] = l == desc[k].L ? 1 : desc[k].w;
        }
        ntheta += desc[k].n_params;
    }
    // density never enters the physics: its gradient is None in the reference and Adam skips it
    const int n_active = cfg->net_enabled[2] ? theta_off[2] : ntheta;
    const bool fused_ar = peer && ntheta + L_COUNT <= pv.ar_slot;
    // The iteration is captured once and replayed as a CUDA graph (every argument is iteration-independent; the
    // peer-memory epochs live in device memory).  Measured on B200, eager -> replayed: 1.144 -> 1.111 ms per
    // iteration on the 10^6-element lattice, 0.464 -> 0.435 on 3.5 x 10^5 elements, 0.245 -> 0.213 on 1.25 x 10^5,
    // 0.285 -> 0.260 element-sharded over 8 GPUs.  Runs with NCCL calls inside the iteration stay eager unless
    // PF_GD_GRAPH=1 asks for capture; PF_GD_GRAPH=0 turns replay off.
    // Under a kernel profiler (ncu / nsys attach through a CUDA injection library) the loop also stays eager, so the
    // profiler sees ordinary launches: ncu 2025.2 fails with LaunchFailed on the replayed mlp_tc_kernel node
    // (101 KB of opt-in dynamic shared memory) although the same graph runs correctly -- and bitwise equal to the
    // eager loop -- without it.
    const char* graph_env = getenv("PF_GD_GRAPH");  // read per solve
    const bool nccl_in_loop = sh && pf_comm_world(comm) > 1 && !(fused_ar && pf_halo_uses_peer(sh->halo, 1));
    const bool profiler = getenv("CUDA_INJECTION64_PATH") || getenv("NV_NSIGHT_INJECTION_PORT_BASE") ||
                          getenv("NV_COMPUTE_PROFILER_PERFWORKS_DIR") || getenv("NSYS_PROFILING_SESSION_ID");
    const int no_graph = graph_env ? !atoi(graph_env) : (nccl_in_loop || profiler ? 1 : 0);
    const int n_meas = cfg->n_measured;                                   // measurements this rank holds
    const int n_meas_all = sh ? sh->n_measured_global : n_meas;           // of the whole mesh (the mean's divisor)
    const bool has_meas = n_meas_all > 0 && cfg->alpha_data > 0.0 && (sh || (meas_dofs && meas_vals_all));
    const bool local_meas = has_meas && n_meas > 0 && meas_dofs && meas_vals_all;
    const int64_t nfree_all = sh ? sh->nfree_global : plan->nfree;

    LargeCfg c{};
    c.tolerance = cfg->tolerance;
    c.lr_u = cfg->learning_rate_u;
    c.lr_t = cfg->learning_rate_theta;
    c.alpha_p = cfg->alpha_physics;
    c.alpha_d = cfg->alpha_data;
    c.load_factor = cfg->load_factor;
    c.legacy = cfg->loss_mode == 1;
    c.gscale = c.legacy ? 2.0 * cfg->alpha_physics / (double)nfree_all : cfg->alpha_physics;
    c.has_meas = has_meas ? 1 : 0;
    c.n_meas = n_meas_all;
    c.cdata = has_meas ? -2.0 * cfg->alpha_data / n_meas_all : 0.0;
    c.nfree = nfree_all;
    c.max_iterations = cfg->max_iterations;

    if (!sh && nprob > 1) {  // several problems on one mesh: advance them in groups on the batched kernels
        BatchNets bn{};
        for (int k = 0; k < 3; ++k) {
            bn.desc[k] = desc[k];
            bn.theta_off[k] = theta_off[k];
        }
        bn.ntheta = ntheta;
        bn.n_active = n_active;
        bn.tl = tl;
        if (batched_capable(plan, cfg, bn)) {
            std::vector<int32_t> md(local_meas ? n_meas : 0);
            if (local_meas) {
                PF_CUDA_CHECK(cudaMemcpyAsync(md.data(), meas_dofs, n_meas * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
                PF_CUDA_CHECK(cudaStreamSynchronize(st));
                for (int j = 0; j < n_meas; ++j)
                    PF_REQUIRE(md[j] >= 0 && md[j] < ndof, "measured DOF %d out of range", md[j]);
            }
            return gd_solve_batched(plan, cfg, bn, c, nprob, theta_all, u_all, f_ext, meas_dofs, meas_vals_all, history_all,
                                    n_iters, converged, reactions_all, ws, md, no_graph);
        }
    }

    DevBuf buf;  // declared after `ws`: released (stream-ordered) before the work stream is joined and destroyed
    buf.open(st);
    double *E, *A, *r, *gu, *gE, *gA, *mu, *vu, *mt, *vt, *gt, *sc, *upart, *msum = nullptr, *mcnt = nullptr, *fint;
    int rc;
    if ((rc = buf.alloc(&E, nelem)) || (rc = buf.alloc(&A, nelem)) || (rc = buf.alloc(&r, ndof)) ||
        (rc = buf.alloc(&gu, ndof)) || (rc = buf.alloc(&gE, nelem)) || (rc = buf.alloc(&gA, nelem)) ||
        (rc = buf.alloc(&mu, ndof)) || (rc = buf.alloc(&vu, ndof)) || (rc = buf.alloc(&mt, ntheta)) ||
        (rc = buf.alloc(&vt, ntheta)) || (rc = buf.alloc(&gt, ntheta + L_COUNT)) || (rc = buf.alloc(&sc, S_COUNT)) ||
        (rc = buf.alloc(&fint, ndof)))
        return rc;
    const int ublocks = (int)std::min<int64_t>((ndof + kRedThreads - 1) / kRedThreads, 1024);
    double* rpart;
    if ((rc = buf.alloc(&upart, ublocks)) || (rc = buf.alloc(&rpart, ublocks))) return rc;
    if (has_meas && ((rc = buf.alloc(&msum, ndof)) || (rc = buf.alloc(&mcnt, ndof)))) return rc;
    // size the plan workspace once (residual partial sums, MLP gradient partials) so it is never
    // reallocated while iterations are in flight
    {
        size_t need = (size_t)std::max<int64_t>(plan->nnode, (int64_t)plan->patches.size()) * sizeof(double);
        for (int k = 0; k < 2; ++k) {
            if (!cfg->net_enabled[k]) continue;
            const size_t rows = nelem >= 2048 ? (size_t)plan->sm_count * 8 : (size_t)((nelem + 31) / 32);
            need = std::max(need, rows * desc[k].n_params * sizeof(double));
        }
        if ((rc = pf_plan_reserve_work(plan, need))) return rc;
    }

    // Networks the fragment kernels cover (pf_mlp_frag.cu) keep their hidden activations from the forward pass for
    // the backward pass of the same iteration (no forward recompute); other shapes use pf_mlp_forward / _backward.
    static const int no_frag = getenv("PF_MLP_NO_FRAG") ? atoi(getenv("PF_MLP_NO_FRAG")) : 0;
    double *acts[2] = {nullptr, nullptr}, *fpart = nullptr;
    bool frag[2] = {false, false};
    {
        size_t part_len = 0;
        for (int k = 0; k < 2; ++k) {
            frag[k] = cfg->net_enabled[k] && !no_frag && nelem >= 2048 && pf_mlp_frag_supported(desc[k], true);
            if (!frag[k]) continue;
            if ((rc = buf.alloc(&acts[k], (size_t)pf_mlp_frag_acts_len(desc[k], nelem)))) return rc;
            part_len = std::max(part_len, (size_t)pf_mlp_frag_chunks(desc[k], nelem, 1, plan->sm_count) * desc[k].n_params);
        }
        if (part_len && (rc = buf.alloc(&fpart, part_len))) return rc;
    }

    std::vector<int32_t> h_md;
    std::vector<double> h_mv, h_sum, h_cnt;
    double* losses = gt + ntheta;  // tail of the reduction buffer
    if (local_meas) {
        h_md.resize(n_meas);
        PF_CUDA_CHECK(cudaMemcpyAsync(h_md.data(), meas_dofs, n_meas * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        PF_CUDA_CHECK(cudaStreamSynchronize(st));
        for (int j = 0; j < n_meas; ++j)
            PF_REQUIRE(h_md[j] >= 0 && h_md[j] < ndof, "measured DOF %d out of range", h_md[j]);
    }
    const unsigned eb = (unsigned)((nelem + 255) / 256), db = (unsigned)((ndof + 255) / 256);
    double h_sc[S_COUNT];

    for (int64_t p = 0; p < nprob; ++p) {
        double* theta = theta_all ? theta_all + p * ntheta : nullptr;
        double* u = u_all + p * ndof;
        const double* mv = local_meas ? meas_vals_all + p * n_meas : nullptr;
        double* history = history_all ? history_all + p * (int64_t)std::max(cfg->max_iterations, 1) * PF_GD_HISTORY_COLS
                                      : nullptr;
        if (has_meas) {  // per-DOF sums of the targets: d/du of mean((m - u)^2) in closed form
            h_mv.resize(n_meas);
            if (local_meas) {
                PF_CUDA_CHECK(cudaMemcpyAsync(h_mv.data(), mv, n_meas * sizeof(double), cudaMemcpyDeviceToHost, st));
                PF_CUDA_CHECK(cudaStreamSynchronize(st));
            }
            h_sum.assign(ndof, 0.0);
            h_cnt.assign(ndof, 0.0);
            for (int j = 0; j < (local_meas ? n_meas : 0); ++j) {
                h_sum[h_md[j]] += h_mv[j];
                h_cnt[h_md[j]] += 1.0;
            }
            PF_CUDA_CHECK(cudaMemcpyAsync(msum, h_sum.data(), ndof * sizeof(double), cudaMemcpyHostToDevice, st));
            PF_CUDA_CHECK(cudaMemcpyAsync(mcnt, h_cnt.data(), ndof * sizeof(double), cudaMemcpyHostToDevice, st));
        }
        PF_CUDA_CHECK(cudaMemsetAsync(mu, 0, ndof * sizeof(double), st));
        PF_CUDA_CHECK(cudaMemsetAsync(vu, 0, ndof * sizeof(double), st));
        if (ntheta) {
            PF_CUDA_CHECK(cudaMemsetAsync(mt, 0, ntheta * sizeof(double), st));
            PF_CUDA_CHECK(cudaMemsetAsync(vt, 0, ntheta * sizeof(double), st));
        }
        PF_CUDA_CHECK(cudaMemsetAsync(gt, 0, (ntheta + L_COUNT) * sizeof(double), st));
        for (int i = 0; i < S_COUNT; ++i) h_sc[i] = 0.0;
        h_sc[S_POW_B1] = h_sc[S_POW_B2] = 1.0;
        if (cfg->max_iterations == 0) h_sc[S_DONE] = 1.0;
        PF_CUDA_CHECK(cudaMemcpyAsync(sc, h_sc, sizeof(h_sc), cudaMemcpyHostToDevice, st));
        PF_CUDA_CHECK(cudaStreamSynchronize(st));  // h_sum / h_cnt / h_sc are reused by the next problem

        auto materials = [&]() -> int {
            for (int k = 0; k < 2; ++k) {
                double* dst = k == 0 ? E : A;
                if (frag[k]) {
                    int e = pf_mlp_frag_forward(desc[k], theta + theta_off[k], 0, 1, nelem, nullptr, plan->d_centroid,
                                                cfg->load_factor, cfg->net_scale[k], 1, dst, 1, acts[k], 0, plan->sm_count,
                                                st);
                    if (e) return e;
                } else if (cfg->net_enabled[k]) {
                    int e = pf_mlp_forward(plan, desc[k].in_dim, desc[k].L, desc[k].w, theta + theta_off[k], nelem, nullptr,
                                           cfg->load_factor, cfg->net_scale[k], 1, dst, st);
                    if (e) return e;
                } else {
                    fill_kernel<<<eb, 256, 0, st>>>(dst, nelem, cfg->net_scale[k]);
                }
            }
            return PF_OK;
        };

        // one iteration = a fixed launch sequence whose arguments do not depend on the iteration index (the
        // history row and beta^t come from device scalars), so it is captured once into a CUDA graph and
        // replayed (iteration 0 always runs eagerly: lazy allocations, function attributes).
        auto iteration = [&]() -> int {
            if ((rc = materials())) return rc;
            // r = f_int - lambda f_ext on free DOFs, 0.5 sum r^2 (solver.py:262-270)
            if (!sh) {
                if ((rc = pf_residual(plan, PF_ELEM_LINEAR, 1, u, E, A, 0, nullptr, f_ext, 0, cfg->load_factor, r,
                                      losses + L_HALF_SQ, nullptr, st)))
                    return rc;
            } else {
                if ((rc = pf_halo_exchange(sh->halo, u, 1, st))) return rc;  // neighbours' latest u on the halo rows
                if ((rc = pf_residual(plan, PF_ELEM_LINEAR, 1, u, E, A, 0, nullptr, f_ext, 0, cfg->load_factor, r, nullptr,
                                      nullptr, st)))
                    return rc;
                // halo rows of the local residual are incomplete sums: drop them, then fetch the owners' values
                if (ndof > nd_own) PF_CUDA_CHECK(cudaMemsetAsync(r + nd_own, 0, (ndof - nd_own) * sizeof(double), st));
                sq_partial_kernel<<<ublocks, kRedThreads, 0, st>>>(r, nd_own, rpart);
                if (!fused_ar) sum_partials_kernel<<<1, kRedThreads, 0, st>>>(rpart, ublocks, 0.5, losses + L_HALF_SQ);
                if ((rc = pf_halo_exchange(sh->halo, r, 1, st))) return rc;
            }
            // reverse pass: dL/du = gscale K r, dL/dE, dL/dA, dL/dtheta (closed form of the autograd graph)
            if ((rc = pf_tangent_matvec(plan, PF_ELEM_LINEAR, 1, nullptr, E, A, 0, r, gu, st))) return rc;
            if (n_active) {
                if ((rc = pf_material_vjp(plan, PF_ELEM_LINEAR, 1, u, E, A, 0, r, gE, gA, st))) return rc;
                if (sh && sh->elem_owned) mask_elements_kernel<<<eb, 256, 0, st>>>(sh->elem_owned, nelem, gE, gA);
                for (int k = 0; k < 2; ++k) {
                    if (frag[k]) {
                        if ((rc = pf_mlp_frag_backward(desc[k], theta + theta_off[k], 0, 1, nelem, nullptr, plan->d_centroid,
                                                       cfg->load_factor, k == 0 ? gE : gA, 1, acts[k], 0, fpart,
                                                       gt + theta_off[k], 0, plan->sm_count, st)))
                            return rc;
                    } else if (cfg->net_enabled[k] &&
                               (rc = pf_mlp_backward(plan, desc[k].in_dim, desc[k].L, desc[k].w, theta + theta_off[k], nelem,
                                                     nullptr, cfg->load_factor, cfg->net_scale[k], 1, k == 0 ? gE : gA,
                                                     gt + theta_off[k], st)))
                        return rc;
                }
            }
            // sharded: a rank without measurements of its own still resets its slot (the all-reduce left the
            // global sum of the previous iteration there)
            if (local_meas || (sh && has_meas))
                data_loss_kernel<<<1, kRedThreads, 0, st>>>(meas_dofs, mv, local_meas ? n_meas : 0, u, losses);
            adam_u_kernel<<<ublocks, kRedThreads, 0, st>>>(c, nd_own, gu, msum, mcnt, plan->d_dof_free, u, mu, vu, sc, upart);
            // sharded: one all-reduce per iteration of [dL/dtheta | 0.5 sum r^2 | sum data^2 | sum u_free^2] -- inside
            // adam_theta_finish_kernel on the peer-memory transport (which also folds the two partial-sum
            // buffers), an NCCL call otherwise
            if (sh && !fused_ar) {
                sum_partials_kernel<<<1, kRedThreads, 0, st>>>(upart, ublocks, 1.0, losses + L_UNORM_SQ);
                if ((rc = pf_comm_allreduce_sum(comm, gt, ntheta + L_COUNT, st))) return rc;
            }
            adam_theta_finish_kernel<<<1, 1024, 0, st>>>(c, n_active, ntheta, tl, gt, theta, mt, vt, losses, upart,
                                                        sh && !fused_ar ? 0 : ublocks, sc, history, pv,
                                                        fused_ar ? ntheta + L_COUNT : 0, rpart, ublocks);
            PF_CUDA_CHECK(cudaGetLastError());
            return PF_OK;
        };
        bool finished = cfg->max_iterations == 0;
        ws.drop_graph();
        for (int it = 0; it < cfg->max_iterations && !finished; ++it) {
            if (it == 1 && !no_graph && cfg->max_iterations > 2) {
                if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
                    const int crc = iteration();
                    const cudaError_t ce = cudaStreamEndCapture(st, &ws.graph);
                    if (crc != PF_OK || ce != cudaSuccess || !ws.graph ||
                        cudaGraphInstantiate(&ws.exec, ws.graph, 0) != cudaSuccess) {
                        ws.drop_graph();
                        cudaGetLastError();  // capture is an optimisation: fall back to eager launches
                    }
                } else {
                    cudaGetLastError();
                }
            }
            if (ws.exec) {
                PF_CUDA_CHECK(cudaGraphLaunch(ws.exec, st));
            } else if ((rc = iteration())) {
                return rc;
            }
            if ((it + 1) % kPoll == 0 || it + 1 == cfg->max_iterations) {
                PF_CUDA_CHECK(cudaMemcpyAsync(h_sc, sc, sizeof(h_sc), cudaMemcpyDeviceToHost, st));
                PF_CUDA_CHECK(cudaStreamSynchronize(st));
                finished = h_sc[S_DONE] != 0.0;
                if (peer && (rc = pf_comm_peer_check(comm, st))) return rc;  // a neighbour never arrived
            }
        }
        if (cfg->max_iterations == 0) {
            PF_CUDA_CHECK(cudaMemcpyAsync(h_sc, sc, sizeof(h_sc), cudaMemcpyDeviceToHost, st));
            PF_CUDA_CHECK(cudaStreamSynchronize(st));
        }
        const int32_t iters = (int32_t)h_sc[S_ITERS], conv = (int32_t)h_sc[S_CONV];
        PF_CUDA_CHECK(cudaMemcpyAsync(n_iters + p, &iters, sizeof(int32_t), cudaMemcpyHostToDevice, st));
        PF_CUDA_CHECK(cudaMemcpyAsync(converged + p, &conv, sizeof(int32_t), cudaMemcpyHostToDevice, st));
        if (sh && (rc = pf_halo_exchange(sh->halo, u, 1, st))) return rc;  // halo rows follow the owners' last step
        if (reactions_all) {
            if ((rc = materials())) return rc;
            if ((rc = pf_residual(plan, PF_ELEM_LINEAR, 1, u, E, A, 0, fint, nullptr, 0, 0.0, nullptr, nullptr, nullptr, st)))
                return rc;
            reactions_kernel<<<db, 256, 0, st>>>(fint, f_ext, cfg->load_factor, plan->d_dof_free, ndof,
                                                 reactions_all + p * ndof);
            if (ndof > nd_own)  // halo rows belong to the neighbours
                PF_CUDA_CHECK(cudaMemsetAsync(reactions_all + p * ndof + nd_own, 0, (ndof - nd_own) * sizeof(double), st));
            PF_CUDA_CHECK(cudaGetLastError());
        }
        PF_CUDA_CHECK(cudaStreamSynchronize(st));  // iters / conv are stack variables
    }
    return PF_OK;
}

extern "C" int pf_gd_solve_sharded(pf_plan* plan, const pf_gd_config* cfg, const pf_gd_shard* shard, double* theta,
                                   double* u_local, const double* f_ext_local, const int32_t* meas_dofs_local,
                                   const double* meas_vals_local, double* history, int32_t* n_iters,
                                   int32_t* converged, double* reactions_local, void* stream) {
    int rc = pf_plan_activate(plan);
    if (rc) return rc;
    PF_REQUIRE(cfg && shard && u_local && f_ext_local && n_iters && converged, "pf_gd_solve_sharded: NULL argument");
    PF_REQUIRE(cfg->kind == PF_ELEM_LINEAR, "pf_gd_solve_sharded supports the linear element only");
    PF_REQUIRE(cfg->max_iterations >= 0, "max_iterations must be >= 0");
    return pf_gd_solve_large(plan, cfg, 1, theta, u_local, f_ext_local, meas_dofs_local, meas_vals_local, history, n_iters,
                             converged, reactions_local, pf_stream_of(stream), shard);
}
