// Gauss-Newton / Levenberg-Marquardt building blocks (fem/nn_solver.py:50-135, :223-277).
//
//  * pf_gn_jacobian: the stacked Jacobian J = [[a_p K_ff, a_p J_utheta], [a_d J_du, 0]].
//    The reference obtains J_utheta with n_free x n_tensors reverse passes
//    (nn_solver.py:91-110); here d f_int / d theta is assembled in closed form,
//    row (free DOF) by row, from the per-element parameter Jacobians of the
//    material networks: d f_d / d theta = sum_{e ni d} (A_e h_ed) dE_e/dtheta + (E_e h_ed) dA_e/dtheta.
//  * pf_gn_normal_equations: JtJ = J^T J with fp64 tensor-core MMA (DMMA m8n8k4),
//    Jtr = J^T R, damping = factor * trace(JtJ) / n added to the diagonal
//    (nn_solver.py:268-274).
#include "pf_element.cuh"
#include "pf_internal.h"

namespace {

// ---------------------------------------------------------------------------------------
// J^T J with mma.sync.aligned.m8n8k4.f64 (DMMA).  CTA tile 64x64 of C, 4 warps of 32x32,
// K (rows of J) consumed 32 at a time from a double-buffered shared-memory stage.  Only tiles
// with j0 >= i0 are computed; the mirror tile is written from the same accumulators.
// ---------------------------------------------------------------------------------------
constexpr int TN = 64, KC = 32, LDS_ = TN + 4;  // row stride 68 doubles: conflict-free fragment loads
constexpr int kJtjStage = 2 * KC * LDS_;         // doubles per pipeline stage (I tile + J tile)

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

// Two-stage cp.async pipeline over the rows of J: while the warps run the 8 k4-steps (128 DMMA each) of
// chunk c, the 2 x 32 x 64 doubles of chunk c+1 stream into the other stage.  Edge tiles (ragged n or m,
// odd n: 16-byte alignment) take synchronous zero-filled loads instead.
__global__ void __launch_bounds__(128) jtj_dmma_kernel(int64_t m, int64_t n, const double* __restrict__ J,
                                                       double* __restrict__ C) {
    const int bi = blockIdx.y, bj = blockIdx.x;
    if (bj < bi) return;  // symmetric: upper tiles only
    extern __shared__ __align__(16) double jsm[];
    const int64_t i0 = (int64_t)bi * TN, j0 = (int64_t)bj * TN;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wi = (warp >> 1) * 32, wj = (warp & 1) * 32;  // warp's 32x32 corner inside the tile
    const int g = lane >> 2, t = lane & 3;
    double acc[4][4][2];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;

    const bool interior = i0 + TN <= n && j0 + TN <= n && (n & 1) == 0 &&
                          (reinterpret_cast<uintptr_t>(J) & 15) == 0;
    auto load = [&](int64_t r0, int stage) {
        double* sI = jsm + stage * kJtjStage;
        double* sJ = sI + KC * LDS_;
        if (interior && r0 + KC <= m) {
            // 2 tiles x 32 rows x 32 16-byte pieces = 2048 copies, 16 per thread
            for (int q = tid; q < 2 * KC * (TN / 2); q += 128) {
                const int which = q / (KC * (TN / 2)), rem = q - which * (KC * (TN / 2));
                const int k = rem / (TN / 2), c2 = rem - k * (TN / 2);
                const double* src = J + (r0 + k) * n + (which ? j0 : i0) + 2 * c2;
                double* dst = (which ? sJ : sI) + k * LDS_ + 2 * c2;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)),
                             "l"(src)
                             : "memory");
            }
        } else {
            for (int q = tid; q < KC * TN; q += 128) {
                const int k = q / TN, c = q % TN;
                const int64_t r = r0 + k;
                sI[k * LDS_ + c] = (r < m && i0 + c < n) ? J[r * n + i0 + c] : 0.0;
                sJ[k * LDS_ + c] = (r < m && j0 + c < n) ? J[r * n + j0 + c] : 0.0;
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    const int64_t nchunk = (m + KC - 1) / KC;
    load(0, 0);
    for (int64_t c = 0; c < nchunk; ++c) {
        const int stage = (int)(c & 1);
        if (c + 1 < nchunk) {
            load((c + 1) * KC, stage ^ 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        const double* sI = jsm + stage * kJtjStage;
        const double* sJ = sI + KC * LDS_;
#pragma unroll
        for (int k4 = 0; k4 < KC; k4 += 4) {
            double af[4], bf[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) af[a] = sI[(k4 + t) * LDS_ + wi + a * 8 + g];  // A[row g][k t] = J[k][i]
#pragma unroll
            for (int b = 0; b < 4; ++b) bf[b] = sJ[(k4 + t) * LDS_ + wj + b * 8 + g];  // B[k t][col g] = J[k][j]
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) dmma(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
        }
        __syncthreads();  // the stage is refilled by the next iteration's load
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int64_t i = i0 + wi + a * 8 + g, j = j0 + wj + b * 8 + t * 2 + h;
                if (i < n && j < n) {
                    C[i * n + j] = acc[a][b][h];
                    if (bi != bj) C[j * n + i] = acc[a][b][h];
                }
            }
}

__global__ void jtr_kernel(int64_t m, int64_t n, const double* __restrict__ J, const double* __restrict__ R,
                           double* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double acc = 0.0;
    for (int64_t r = 0; r < m; ++r) acc = fma(J[r * n + i], R[r], acc);
    out[i] = acc;
}

// damping = factor * trace(C) / n; C += damping * I   (single CTA, fixed-order reduction)
__global__ void __launch_bounds__(1024) lm_damping_kernel(int64_t n, double factor, double* __restrict__ C,
                                                          double* __restrict__ damping_out) {
    __shared__ double s[1024];
    __shared__ double s_d;
    double acc = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) acc += C[i * n + i];
    s[threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tr = 0.0;
        for (int q = 0; q < (int)blockDim.x; ++q) tr += s[q];
        s_d = factor * tr / (double)n;
        if (damping_out) *damping_out = s_d;
    }
    __syncthreads();
    const double d = s_d;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) C[i * n + i] += d;
}

// ---------------------------------------------------------------------------------------
// Jacobian assembly
// ---------------------------------------------------------------------------------------
template <int DIM>
__global__ void jac_kff_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colind,
                               const double* __restrict__ vals, const int32_t* __restrict__ free_index,
                               int64_t nnode, int64_t ld, double alpha, double* __restrict__ J) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= nnode) return;
    for (int p = rowptr[n]; p < rowptr[n + 1]; ++p) {
        const int64_t mcol = colind[p];
        for (int r = 0; r < DIM; ++r) {
            const int row = free_index[n * DIM + r];
            if (row < 0) continue;
            for (int c = 0; c < DIM; ++c) {
                const int col = free_index[mcol * DIM + c];
                if (col >= 0) J[(int64_t)row * ld + col] = alpha * vals[(int64_t)p * DIM * DIM + r * DIM + c];
            }
        }
    }
}

// one CTA per node; rows = the node's free DOFs; threads stride over the parameter columns
template <int DIM>
__global__ void __launch_bounds__(256) jac_utheta_kernel(const int32_t* __restrict__ inc_ptr,
                                                         const PfIncidence* __restrict__ inc,
                                                         const double4* __restrict__ inc_geo,
                                                         const int32_t* __restrict__ free_index,
                                                         const double* __restrict__ u, const double* __restrict__ E,
                                                         const double* __restrict__ A,
                                                         const double* __restrict__ jacE, int64_t nE,
                                                         const double* __restrict__ jacA, int64_t nA, int64_t col0,
                                                         int64_t ld, double alpha, double* __restrict__ J) {
    const int64_t n = blockIdx.x;
    const int rx = free_index[n * DIM], ry = DIM == 2 ? free_index[n * DIM + 1] : -1;
    if (rx < 0 && ry < 0) return;
    const double xs = DIM == 2 ? u[2 * n] : u[n], ys = DIM == 2 ? u[2 * n + 1] : 0.0;
    for (int64_t q = threadIdx.x; q < nE + nA; q += blockDim.x) {
        double ax = 0.0, ay = 0.0;
        for (int k = inc_ptr[n]; k < inc_ptr[n + 1]; ++k) {
            const PfIncidence ic = inc[k];
            const double4 geo = inc_geo[k];
            double hx = 0.0, hy = 0.0;  // unit-EA force of this element on the node
            const double xo = DIM == 2 ? u[2 * ic.nbr] : u[ic.nbr], yo = DIM == 2 ? u[2 * ic.nbr + 1] : 0.0;
            pf_linear_incidence<DIM>(1.0, 1.0, geo, xs, ys, xo, yo, hx, hy);
            // d f / d theta_q = (A h) dE/dtheta_q   or   (E h) dA/dtheta_q
            const double w = q < nE ? A[ic.elem] * jacE[(int64_t)ic.elem * nE + q]
                                    : E[ic.elem] * jacA[(int64_t)ic.elem * nA + (q - nE)];
            ax = fma(hx, w, ax);
            ay = fma(hy, w, ay);
        }
        if (rx >= 0) J[(int64_t)rx * ld + col0 + q] = alpha * ax;
        if (ry >= 0) J[(int64_t)ry * ld + col0 + q] = alpha * ay;
    }
}

__global__ void jac_data_kernel(const int32_t* __restrict__ meas_dofs, int64_t n_meas,
                                const int32_t* __restrict__ free_index, int64_t row0, int64_t ld, double alpha,
                                double* __restrict__ J) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_meas) return;
    const int col = free_index[meas_dofs[i]];
    if (col >= 0) J[(row0 + i) * ld + col] = -alpha;  // d(m - u)/du = -1 (nn_solver.py:121-130)
}

}  // namespace

extern "C" int pf_gn_normal_equations(int64_t m, int64_t n, const double* J, const double* R, double damping_factor,
                                      double* jtj, double* jtr, double* damping_out, void* stream) {
    PF_REQUIRE(m >= 1 && n >= 1, "pf_gn_normal_equations: empty system");
    PF_REQUIRE(J && jtj, "pf_gn_normal_equations: NULL argument");
    cudaStream_t st = pf_stream_of(stream);
    const unsigned tiles = (unsigned)((n + TN - 1) / TN);
    constexpr size_t kJtjSmem = 2 * kJtjStage * sizeof(double);
    PF_CUDA_CHECK(cudaFuncSetAttribute(jtj_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kJtjSmem));
    jtj_dmma_kernel<<<dim3(tiles, tiles), 128, kJtjSmem, st>>>(m, n, J, jtj);
    PF_CUDA_CHECK(cudaGetLastError());
    if (R && jtr) {
        jtr_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(m, n, J, R, jtr);
        PF_CUDA_CHECK(cudaGetLastError());
    }
    lm_damping_kernel<<<1, 1024, 0, st>>>(n, damping_factor, jtj, damping_out);
    PF_CUDA_CHECK(cudaGetLastError());
    return PF_OK;
}

extern "C" int pf_gn_jacobian(pf_plan* plan, int kind, const double* u, const double* E, const double* A,
                              const double* vals, const double* jacE, int64_t nE, const double* jacA, int64_t nA,
                              int64_t n_rest, double alpha_physics, double alpha_data, const int32_t* meas_dofs,
                              int64_t n_meas, double* J, void* stream) {
    int rc = pf_plan_activate(plan);
    if (rc) return rc;
    PF_REQUIRE(kind == PF_ELEM_LINEAR, "pf_gn_jacobian supports the linear element only");
    PF_REQUIRE(u && E && A && vals && J, "pf_gn_jacobian: NULL argument");
    PF_REQUIRE((nE == 0 || jacE) && (nA == 0 || jacA), "pf_gn_jacobian: parameter Jacobian is NULL");
    PF_REQUIRE(n_meas == 0 || meas_dofs, "pf_gn_jacobian: meas_dofs is NULL");
    cudaStream_t st = pf_stream_of(stream);
    const int64_t ncols = plan->nfree + nE + nA + n_rest, nrows = plan->nfree + n_meas;
    PF_CUDA_CHECK(cudaMemsetAsync(J, 0, (size_t)nrows * ncols * sizeof(double), st));
    const unsigned nb = (unsigned)((plan->nnode + 127) / 128);
    if (plan->dim == 1)
        jac_kff_kernel<1><<<nb, 128, 0, st>>>(plan->d_bsr_rowptr, plan->d_bsr_colind, vals, plan->d_free_index,
                                              plan->nnode, ncols, alpha_physics, J);
    else
        jac_kff_kernel<2><<<nb, 128, 0, st>>>(plan->d_bsr_rowptr, plan->d_bsr_colind, vals, plan->d_free_index,
                                              plan->nnode, ncols, alpha_physics, J);
    if (nE + nA > 0) {
        if (plan->dim == 1)
            jac_utheta_kernel<1><<<(unsigned)plan->nnode, 256, 0, st>>>(plan->d_inc_ptr, plan->d_inc, plan->d_inc_geo,
                                                                        plan->d_free_index, u, E, A, jacE, nE, jacA, nA,
                                                                        plan->nfree, ncols, alpha_physics, J);
        else
            jac_utheta_kernel<2><<<(unsigned)plan->nnode, 256, 0, st>>>(plan->d_inc_ptr, plan->d_inc, plan->d_inc_geo,
                                                                        plan->d_free_index, u, E, A, jacE, nE, jacA, nA,
                                                                        plan->nfree, ncols, alpha_physics, J);
    }
    if (n_meas > 0)
        jac_data_kernel<<<(unsigned)((n_meas + 127) / 128), 128, 0, st>>>(meas_dofs, n_meas, plan->d_free_index,
                                                                          plan->nfree, ncols, alpha_data, J);
    PF_CUDA_CHECK(cudaGetLastError());
    return PF_OK;
}

namespace {
__global__ void transpose_small_kernel(int64_t m, int64_t n, const double* __restrict__ J, double* __restrict__ Jt) {
    __shared__ double t[32][33];
    const int64_t c0 = (int64_t)blockIdx.x * 32, r0 = (int64_t)blockIdx.y * 32;
    if (r0 + threadIdx.y < m && c0 + threadIdx.x < n) t[threadIdx.y][threadIdx.x] = J[(r0 + threadIdx.y) * n + c0 + threadIdx.x];
    __syncthreads();
    if (c0 + threadIdx.y < n && r0 + threadIdx.x < m) Jt[(c0 + threadIdx.y) * m + r0 + threadIdx.x] = t[threadIdx.x][threadIdx.y];
}
// dx[c] = -sum_i J[i][c] y[i] (ascending i: fixed order)
__global__ void neg_jt_times_kernel(int64_t m, int64_t n, const double* __restrict__ J, const double* __restrict__ y,
                                    double* __restrict__ dx) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    double acc = 0.0;
    for (int64_t i = 0; i < m; ++i) acc = fma(J[i * n + c], y[i], acc);
    dx[c] = -acc;
}
__global__ void negate_kernel(int64_t n, const double* __restrict__ x, double* __restrict__ y) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = -x[i];
}
}  // namespace

// One Levenberg-Marquardt step dx = -(J^T J + d I)^-1 J^T R, d = damping_factor * trace(J^T J) / n
// (fem/nn_solver.py:266-277).  The inverse problems of this code base have far fewer residuals than unknowns
// (6 x 1001 on example 10): then the step is taken through the m x m dual system
//     (J J^T + d I) y = R,   dx = -J^T y            [(J^T J + d I)^-1 J^T = J^T (J J^T + d I)^-1]
// which costs O(m^2 n) instead of O(n^3) and is better conditioned (the n x n matrix is J^T J plus a 1e-6-relative
// ridge on an (n - m)-dimensional null space).  path: 0 = choose (dual when m < n), 1 = n x n system, 2 = dual.
extern "C" int pf_gn_lm_step(int64_t m, int64_t n, const double* J, const double* R, double damping_factor, int path,
                             double* dx, double* damping_out, int32_t* info, void* stream) {
    PF_REQUIRE(m >= 1 && n >= 1 && J && R && dx && info, "pf_gn_lm_step: bad argument");
    PF_REQUIRE(path >= 0 && path <= 2, "pf_gn_lm_step: path must be 0, 1 or 2");
    cudaStream_t st = pf_stream_of(stream);
    pf_keep_pool_cached();
    const bool dual = path == 2 || (path == 0 && m < n);
    double* work = nullptr;
    int rc = PF_OK;
    if (dual) {
        PF_CUDA_CHECK(cudaMallocAsync((void**)&work, (size_t)(n * m + m * m + m) * sizeof(double), st));
        double *Jt = work, *G = Jt + n * m, *y = G + m * m;
        transpose_small_kernel<<<dim3((unsigned)((n + 31) / 32), (unsigned)((m + 31) / 32)), dim3(32, 32), 0, st>>>(m, n, J, Jt);
        // Gram matrix of the rows = (J^T)^T (J^T); its trace equals trace(J^T J), the ridge is rescaled to the n unknowns
        rc = pf_gn_normal_equations(n, m, Jt, nullptr, damping_factor * (double)m / (double)n, G, nullptr, damping_out, stream);
        if (!rc) {
            cudaMemcpyAsync(y, R, m * sizeof(double), cudaMemcpyDeviceToDevice, st);
            rc = pf_solve_spd(m, G, y, info, stream);
        }
        if (!rc) neg_jt_times_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(m, n, J, y, dx);
    } else {
        PF_CUDA_CHECK(cudaMallocAsync((void**)&work, (size_t)(n * n + n) * sizeof(double), st));
        double *jtj = work, *jtr = jtj + n * n;
        rc = pf_gn_normal_equations(m, n, J, R, damping_factor, jtj, jtr, damping_out, stream);
        if (!rc) {
            negate_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, jtr, dx);
            rc = pf_solve_spd(n, jtj, dx, info, stream);
        }
    }
    const cudaError_t le = cudaGetLastError();
    cudaFreeAsync(work, st);
    if (rc) return rc;
    PF_CUDA_CHECK(le);
    return PF_OK;
}
