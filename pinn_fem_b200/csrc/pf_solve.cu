// Linear solves (fp64).
//
//  * pf_solve_dense: LU with partial pivoting, the algorithm behind
//    np.linalg.solve (fem/core.py:35, fem/solver.py:464) and torch.linalg.solve
//    (fem/nn_solver.py:277).
//      - n <= 120: one CTA per system, the augmented matrix [A | b] lives in
//        shared memory (batched Newton steps on small meshes);
//      - larger n: blocked right-looking LU in global memory: the 32-column panel
//        is factorised by a thread-block CLUSTER of 8 CTAs that keeps the whole
//        panel in distributed shared memory (pivot candidates and the pivot row
//        travel over DSMEM, two cluster barriers per column), then row swaps +
//        triangular solve and a tiled rank-32 update spread over all SMs.  Panels
//        too tall for the cluster's shared memory (> ~6000 rows) use the
//        single-CTA panel kernel.
//  * pf_cg_solve: Jacobi-preconditioned conjugate gradients on the free DOFs,
//    matrix-free through the assembly mat-vec kernel, batched over problems,
//    with fixed-order (deterministic) two-stage dot products.
#include <cooperative_groups.h>

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdlib>

#include "pf_internal.h"

namespace cg = cooperative_groups;

namespace {

constexpr int kSmallMax = 120;

// argmax |x| with the lowest index winning ties (LAPACK idamax semantics)
__device__ __forceinline__ void warp_argmax(double& v, int& i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double v2 = __shfl_xor_sync(0xffffffffu, v, o);
        const int i2 = __shfl_xor_sync(0xffffffffu, i, o);
        if (v2 > v || (v2 == v && i2 < i)) {
            v = v2;
            i = i2;
        }
    }
}

__global__ void __launch_bounds__(128) lu_small_kernel(int n, double* __restrict__ Ag, double* __restrict__ bg,
                                                       int32_t* __restrict__ info) {
    extern __shared__ double s[];
    const int ld = n + 1 + (n & 1);  // n+1 columns (b is the last), odd row stride: conflict-free column walks
    double* A = s;                   // [n][ld], column n holds b
    __shared__ int s_piv;
    __shared__ int s_info;
    const int tid = threadIdx.x, nt = blockDim.x;
    double* Asrc = Ag + (size_t)blockIdx.x * n * n;
    double* bsrc = bg + (size_t)blockIdx.x * n;
    for (int q = tid; q < n * n; q += nt) A[(q / n) * ld + (q % n)] = Asrc[q];
    for (int q = tid; q < n; q += nt) A[q * ld + n] = bsrc[q];
    if (tid == 0) s_info = 0;
    __syncthreads();
    for (int k = 0; k < n; ++k) {
        if (tid < 32) {
            double best = -1.0;
            int bi = k;
            for (int i = k + tid; i < n; i += 32) {
                const double v = fabs(A[i * ld + k]);
                if (v > best) {
                    best = v;
                    bi = i;
                }
            }
            warp_argmax(best, bi);
            if (tid == 0) {
                s_piv = bi;
                if (!(best > 0.0) && s_info == 0) s_info = k + 1;  // zero column, or nothing but NaNs (best stays -1)
            }
        }
        __syncthreads();
        const int p = s_piv;
        if (p != k)
            for (int c = tid; c <= n; c += nt) {
                const double t = A[k * ld + c];
                A[k * ld + c] = A[p * ld + c];
                A[p * ld + c] = t;
            }
        __syncthreads();
        const double piv = A[k * ld + k];
        if (piv != 0.0) {
            const int rows = n - k - 1, cols = n - k;  // columns k+1..n (incl. rhs)
            for (int q = tid; q < rows * cols; q += nt) {
                const int i = k + 1 + q / cols, c = k + 1 + q % cols;
                const double l = A[i * ld + k] / piv;
                A[i * ld + c] = fma(-l, A[k * ld + c], A[i * ld + c]);
            }
        }
        __syncthreads();
    }
    // back substitution (column oriented: one barrier per unknown)
    for (int i = n - 1; i >= 0; --i) {
        const double d = A[i * ld + i];
        if (tid == 0) A[i * ld + n] = d != 0.0 ? A[i * ld + n] / d : A[i * ld + n];
        __syncthreads();
        const double xi = A[i * ld + n];
        for (int r = tid; r < i; r += nt) A[r * ld + n] = fma(-A[r * ld + i], xi, A[r * ld + n]);
        __syncthreads();
    }
    for (int q = tid; q < n; q += nt) bsrc[q] = A[q * ld + n];
    if (tid == 0) info[blockIdx.x] = s_info;
}

// ---- blocked LU for one large system -------------------------------------------------
constexpr int NB = 32;

// Factor the panel A[k0:n, k0:k0+nb) in place with partial pivoting; piv[j] = chosen row for column k0+j.
__global__ void __launch_bounds__(1024) lu_panel_kernel(int n, int k0, int nb, double* __restrict__ A,
                                                        int32_t* __restrict__ piv, int32_t* __restrict__ info) {
    __shared__ double s_val[32];
    __shared__ int s_idx[32];
    __shared__ int s_p;
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
    for (int j = 0; j < nb; ++j) {
        const int col = k0 + j;
        double best = -1.0;
        int bi = col;
        for (int i = col + tid; i < n; i += nt) {
            const double v = fabs(A[(size_t)i * n + col]);
            if (v > best) {
                best = v;
                bi = i;
            }
        }
        warp_argmax(best, bi);
        if (lane == 0) {
            s_val[warp] = best;
            s_idx[warp] = bi;
        }
        __syncthreads();
        if (warp == 0) {
            best = lane < nw ? s_val[lane] : -1.0;
            bi = lane < nw ? s_idx[lane] : 0x7fffffff;
            warp_argmax(best, bi);
            if (lane == 0) {
                s_p = bi;
                piv[j] = bi;
                if (!(best > 0.0) && *info == 0) *info = col + 1;  // zero column, or nothing but NaNs (best stays -1)
            }
        }
        __syncthreads();
        const int p = s_p;
        if (p != col && tid < nb) {  // swap inside the panel only; the rest is swapped by lu_swap_trsm_kernel
            const double t = A[(size_t)col * n + k0 + tid];
            A[(size_t)col * n + k0 + tid] = A[(size_t)p * n + k0 + tid];
            A[(size_t)p * n + k0 + tid] = t;
        }
        __syncthreads();
        const double pv = A[(size_t)col * n + col];
        if (pv != 0.0) {
            // one warp per row: lane c updates column col+1+c of the panel
            for (int i = col + 1 + warp; i < n; i += nw) {
                double l = 0.0;
                if (lane == 0) {
                    l = A[(size_t)i * n + col] / pv;
                    A[(size_t)i * n + col] = l;
                }
                l = __shfl_sync(0xffffffffu, l, 0);
                const int c = col + 1 + lane;
                if (c < k0 + nb) A[(size_t)i * n + c] = fma(-l, A[(size_t)col * n + c], A[(size_t)i * n + c]);
            }
        }
        __syncthreads();
    }
}

// The same panel factorisation on a thread-block cluster.  The single-CTA kernel above walks the panel 32 times
// through one SM (1.0-1.3 ms per panel at n = 4096, 87 % of the whole solve).  Here the rows k0..n-1 are dealt in
// contiguous slabs to the kLuCluster CTAs of a cluster and stay in their shared memory for all nb column steps;
// global memory is touched twice (load, store).  Per column: local arg-max, every CTA pushes its candidate into
// every other CTA's shared memory, cluster barrier, every CTA picks the winner (lowest index on ties: idamax)
// and fetches the pivot row from its owner over DSMEM, cluster barrier, the two owners write the swapped rows,
// local rank-1 update.  Arithmetic and pivot choice are those of lu_panel_kernel: same bits.
constexpr int kLuCluster = 8;
constexpr int kLuLd = NB + 1;  // odd row stride: column walks are bank-conflict free

__global__ void __cluster_dims__(kLuCluster, 1, 1) __launch_bounds__(1024)
    lu_panel_cluster_kernel(int n, int k0, int nb, int rows_per, double* __restrict__ A, int32_t* __restrict__ piv,
                            int32_t* __restrict__ info) {
    extern __shared__ double s_slab[];  // [rows_per][kLuLd]
    __shared__ double s_cval[kLuCluster];
    __shared__ int s_cidx[kLuCluster];
    __shared__ double s_wval[32];
    __shared__ int s_widx[32];
    __shared__ double s_prow[NB], s_crow[NB];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
    const int g0 = k0 + rank * rows_per;  // first global row of this CTA's slab
    const int myrows = max(0, min(rows_per, n - g0));
    for (int q = tid; q < myrows * nb; q += nt) {
        const int r = q / nb, c = q - r * nb;
        s_slab[r * kLuLd + c] = A[(size_t)(g0 + r) * n + k0 + c];
    }
    cluster.sync();  // every CTA of the cluster is running and has its slab before anyone touches remote shared memory
    for (int j = 0; j < nb; ++j) {
        const int col = k0 + j;
        double best = -1.0;
        int bi = INT_MAX;
        for (int r = max(0, col - g0) + tid; r < myrows; r += nt) {
            const double v = fabs(s_slab[r * kLuLd + j]);
            if (v > best) {
                best = v;
                bi = g0 + r;
            }
        }
        warp_argmax(best, bi);
        if (lane == 0) {
            s_wval[warp] = best;
            s_widx[warp] = bi;
        }
        __syncthreads();
        if (warp == 0) {
            best = lane < nw ? s_wval[lane] : -1.0;
            bi = lane < nw ? s_widx[lane] : INT_MAX;
            warp_argmax(best, bi);
            if (lane < kLuCluster) {  // this CTA's candidate goes into slot `rank` of CTA `lane`
                cluster.map_shared_rank(s_cval, lane)[rank] = best;
                cluster.map_shared_rank(s_cidx, lane)[rank] = bi;
            }
        }
        cluster.sync();  // (1) all candidates have landed everywhere
        double gb = -1.0;
        int p = INT_MAX;
#pragma unroll
        for (int c = 0; c < kLuCluster; ++c) {
            const double v = s_cval[c];
            const int i = s_cidx[c];
            if (v > gb || (v == gb && i < p)) {
                gb = v;
                p = i;
            }
        }
        // every remaining entry of the column is NaN: no candidate ever compares greater, all CTAs publish
        // (-1, INT_MAX).  Keep row `col` as the pivot (an in-range owner for the DSMEM reads below) and report the
        // column as singular, like a zero pivot: the host raises the reference's error instead of the GPU faulting.
        const bool no_pivot = p == INT_MAX;
        if (no_pivot) p = col;
        if (rank == 0 && tid == 0) {
            piv[j] = p;
            if ((gb == 0.0 || no_pivot) && *info == 0) *info = col + 1;
        }
        const int op = (p - k0) / rows_per, oc = (col - k0) / rows_per;  // owners of the pivot row and of row `col`
        if (tid < nb) {  // everyone needs the pivot row; its owner also needs the row it trades places with
            s_prow[tid] = cluster.map_shared_rank(s_slab, op)[(p - k0 - op * rows_per) * kLuLd + tid];
            if (rank == op && p != col)
                s_crow[tid] = cluster.map_shared_rank(s_slab, oc)[(col - k0 - oc * rows_per) * kLuLd + tid];
        }
        cluster.sync();  // (2) all remote reads are done before the owners overwrite the two rows
        if (tid < nb && p != col) {
            if (rank == oc) s_slab[(col - g0) * kLuLd + tid] = s_prow[tid];
            if (rank == op) s_slab[(p - g0) * kLuLd + tid] = s_crow[tid];
        }
        __syncthreads();
        const double pv = s_prow[j];
        if (pv != 0.0) {
            // one warp per row: lane c updates panel column j+1+c
            for (int r = max(0, col + 1 - g0) + warp; r < myrows; r += nw) {
                double l = 0.0;
                if (lane == 0) {
                    l = s_slab[r * kLuLd + j] / pv;
                    s_slab[r * kLuLd + j] = l;
                }
                l = __shfl_sync(0xffffffffu, l, 0);
                const int c = j + 1 + lane;
                if (c < nb) s_slab[r * kLuLd + c] = fma(-l, s_prow[c], s_slab[r * kLuLd + c]);
            }
        }
        __syncthreads();
    }
    for (int q = tid; q < myrows * nb; q += nt) {
        const int r = q / nb, c = q - r * nb;
        A[(size_t)(g0 + r) * n + k0 + c] = s_slab[r * kLuLd + c];
    }
    cluster.sync();  // no CTA leaves while a neighbour could still be reading its shared memory
}

// Apply the panel's row swaps to the columns outside the panel (and to b), then solve
// L11 * U12 = A12 for the columns right of the panel (and for b's top block).
__global__ void lu_swap_trsm_kernel(int n, int k0, int nb, double* __restrict__ A, double* __restrict__ b,
                                    const int32_t* __restrict__ piv) {
    const int c0 = blockIdx.x * blockDim.x + threadIdx.x;  // column index over [0, n] without the panel
    if (c0 > n - nb) return;
    const int c = c0 < k0 ? c0 : c0 + nb;  // c == n is the right-hand side
    auto at = [&](int r) -> double& { return c == n ? b[r] : A[(size_t)r * n + c]; };
    for (int j = 0; j < nb; ++j) {
        const int p = piv[j];
        if (p != k0 + j) {
            const double t = at(k0 + j);
            at(k0 + j) = at(p);
            at(p) = t;
        }
    }
    if (c < k0) return;
    for (int j = 1; j < nb; ++j) {  // unit lower triangular forward substitution
        double acc = at(k0 + j);
        for (int q = 0; q < j; ++q) acc = fma(-A[(size_t)(k0 + j) * n + k0 + q], at(k0 + q), acc);
        at(k0 + j) = acc;
    }
}

// A22 -= L21 * U12 (and b2 -= L21 * y1): 64x64 output tile per CTA, K = nb <= 32.
__global__ void __launch_bounds__(256) lu_update_kernel(int n, int k0, int nb, double* __restrict__ A,
                                                        double* __restrict__ b) {
    __shared__ double sL[64][NB + 1];
    __shared__ double sU[NB][64 + 1];
    const int r0 = k0 + nb + blockIdx.y * 64, c0 = k0 + nb + blockIdx.x * 64;  // columns run to n (inclusive: rhs)
    const int tid = threadIdx.x;
    for (int q = tid; q < 64 * nb; q += 256) {
        const int i = q / nb, j = q % nb;
        sL[i][j] = (r0 + i < n) ? A[(size_t)(r0 + i) * n + k0 + j] : 0.0;
    }
    for (int q = tid; q < nb * 64; q += 256) {
        const int j = q / 64, c = q % 64;
        const int cc = c0 + c;
        sU[j][c] = cc < n ? A[(size_t)(k0 + j) * n + cc] : (cc == n ? b[k0 + j] : 0.0);
    }
    __syncthreads();
    const int ty = tid / 16, tx = tid % 16;  // 16x16 threads, 4x4 outputs each
    double acc[4][4] = {};
    for (int j = 0; j < nb; ++j) {
        double l[4], uu[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) l[a] = sL[ty * 4 + a][j];
#pragma unroll
        for (int c = 0; c < 4; ++c) uu[c] = sU[j][tx * 4 + c];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[a][c] = fma(l[a], uu[c], acc[a][c]);
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int r = r0 + ty * 4 + a;
        if (r >= n) continue;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int cc = c0 + tx * 4 + c;
            if (cc < n)
                A[(size_t)r * n + cc] -= acc[a][c];
            else if (cc == n)
                b[r] -= acc[a][c];
        }
    }
}

// x = U^{-1} y in place.  One CTA of 32 warps walks the 32-row panels from the bottom: warp r forms the dot product of
// row k0 + r with the part already solved (lanes stride along the row: coalesced; the column-oriented version read
// every column with an n-double stride and took 7.5 ms at n = 4096), then warp 0 finishes the 32 x 32 upper triangle
// out of shared memory.  Fixed summation order.  A zero diagonal entry divides by nothing (the caller reports
// `info`), as before.
__global__ void __launch_bounds__(1024) lu_backsolve_kernel(int n, const double* __restrict__ A, double* b) {
    __shared__ double sT[NB][NB + 1];
    __shared__ double sR[NB][33];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int npanel = (n + NB - 1) / NB;
    for (int p = npanel - 1; p >= 0; --p) {
        const int k0 = p * NB, nb = min(NB, n - k0);
        for (int q = tid; q < nb * nb; q += 1024) {
            const int r = q / nb, c = q % nb;
            sT[r][c] = c >= r ? A[(size_t)(k0 + r) * n + k0 + c] : 0.0;
        }
        double acc = 0.0;
        if (warp < nb)
            for (int j = k0 + nb + lane; j < n; j += 32) acc = fma(A[(size_t)(k0 + warp) * n + j], b[j], acc);
        sR[warp][lane] = acc;
        __syncthreads();
        if (warp == 0) {
            double s = 0.0;
            if (lane < nb)
                for (int q = 0; q < 32; ++q) s += sR[lane][q];
            double rhs = lane < nb ? b[k0 + lane] - s : 0.0;
            for (int c = nb - 1; c >= 0; --c) {
                const double d = sT[c][c];
                const double num = __shfl_sync(0xffffffffu, rhs, c);
                const double xc = d != 0.0 ? num / d : num;
                if (lane == c) rhs = xc;
                else if (lane < c) rhs = fma(-sT[lane][c], xc, rhs);
            }
            if (lane < nb) b[k0 + lane] = rhs;
        }
        __syncthreads();  // the solved block of b is visible to every warp before the next panel's dot products
    }
}

}  // namespace

extern "C" int pf_solve_dense(int64_t nbatch, int64_t n, double* A, double* b, int32_t* info, void* stream) {
    PF_REQUIRE(nbatch >= 1 && n >= 1, "pf_solve_dense: nbatch and n must be >= 1");
    PF_REQUIRE(A && b && info, "pf_solve_dense: NULL argument");
    PF_REQUIRE(n < (1 << 15), "pf_solve_dense: n too large (%lld)", (long long)n);
    cudaStream_t st = pf_stream_of(stream);
    if (n <= kSmallMax) {
        const int ld = (int)n + 1 + ((int)n & 1);
        const size_t smem = (size_t)n * ld * sizeof(double);
        if (smem > 48 * 1024)
            PF_CUDA_CHECK(cudaFuncSetAttribute(lu_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        lu_small_kernel<<<(unsigned)nbatch, 128, smem, st>>>((int)n, A, b, info);
        PF_CUDA_CHECK(cudaGetLastError());
        return PF_OK;
    }
    // cluster panel: the first (tallest) panel decides the shared-memory opt-in; PF_LU_CLUSTER=0 keeps the single-CTA panel
    const int env_cluster = getenv("PF_LU_CLUSTER") ? atoi(getenv("PF_LU_CLUSTER")) : 1;  // read per call
    const size_t cluster_smem_max = 200 * 1024;
    bool use_cluster = env_cluster != 0;
    if (use_cluster) {
        const size_t want = std::min(cluster_smem_max, (size_t)((n + kLuCluster - 1) / kLuCluster) * kLuLd * sizeof(double));
        if (cudaFuncSetAttribute(lu_panel_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)want) !=
            cudaSuccess) {
            cudaGetLastError();
            use_cluster = false;
        }
    }
    int32_t* piv = nullptr;
    pf_keep_pool_cached();
    PF_CUDA_CHECK(cudaMallocAsync((void**)&piv, NB * sizeof(int32_t), st));
    PF_CUDA_CHECK(cudaMemsetAsync(info, 0, nbatch * sizeof(int32_t), st));
    for (int64_t m = 0; m < nbatch; ++m) {
        double* Am = A + (size_t)m * n * n;
        double* bm = b + (size_t)m * n;
        for (int k0 = 0; k0 < (int)n; k0 += NB) {
            const int nb = std::min<int>(NB, (int)n - k0);
            const int rows = (int)n - k0, rows_per = (rows + kLuCluster - 1) / kLuCluster;
            const size_t slab = (size_t)rows_per * kLuLd * sizeof(double);
            if (use_cluster && slab <= cluster_smem_max) {
                lu_panel_cluster_kernel<<<kLuCluster, 1024, slab, st>>>((int)n, k0, nb, rows_per, Am, piv, info + m);
            } else {
                lu_panel_kernel<<<1, 1024, 0, st>>>((int)n, k0, nb, Am, piv, info + m);
            }
            const int ncols = (int)n + 1 - nb;
            lu_swap_trsm_kernel<<<(ncols + 127) / 128, 128, 0, st>>>((int)n, k0, nb, Am, bm, piv);
            const int rem = (int)n - k0 - nb;
            if (rem > 0) {
                dim3 grid((rem + 1 + 63) / 64, (rem + 63) / 64);
                lu_update_kernel<<<grid, 256, 0, st>>>((int)n, k0, nb, Am, bm);
            }
        }
        lu_backsolve_kernel<<<1, 1024, 0, st>>>((int)n, Am, bm);
    }
    PF_CUDA_CHECK(cudaGetLastError());
    PF_CUDA_CHECK(cudaFreeAsync(piv, st));
    return PF_OK;
}

// ---------------------------------------------------------------------------------------
// Jacobi-preconditioned conjugate gradients, matrix-free, batched over problems
// ---------------------------------------------------------------------------------------
namespace {

constexpr int CG_BX = 32, CG_BY = 8, CG_RPT = 32;  // a CTA covers 256 rows x 32 problems
constexpr int CG_TILE_ROWS = CG_BY * CG_RPT;

struct CgVec {
    const uint8_t* __restrict__ dof_free;
    int64_t ndof, B;
};

// per-column partial sums of a*b over the CTA's rows (fixed order) -> part[blockIdx.x][b]
__device__ __forceinline__ void cg_block_reduce(double acc, double* __restrict__ part, int64_t B, int64_t b) {
    __shared__ double s[CG_BY][CG_BX];
    s[threadIdx.y][threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.y == 0 && b < B) {
        double t = 0.0;
        for (int y = 0; y < CG_BY; ++y) t += s[y][threadIdx.x];
        part[(int64_t)blockIdx.x * B + b] = t;
    }
    __syncthreads();
}

// dinv = 1 / diag(K) on free DOFs (0 on fixed): node-centric, linear element
template <int DIM>
__global__ void cg_diag_kernel(const int32_t* __restrict__ inc_ptr, const PfIncidence* __restrict__ inc,
                               const double4* __restrict__ inc_geo, const uint8_t* __restrict__ dof_free,
                               const double* __restrict__ E, const double* __restrict__ A, int64_t mat_stride,
                               int64_t mat_bmul, int64_t nnode, int64_t B, double* __restrict__ dinv) {
    const int64_t b = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
    const int64_t n = (int64_t)blockIdx.x * blockDim.y + threadIdx.y;
    if (b >= B || n >= nnode) return;
    double dx = 0.0, dy = 0.0;
    for (int k = inc_ptr[n]; k < inc_ptr[n + 1]; ++k) {
        const PfIncidence ic = inc[k];
        const double4 g = inc_geo[k];
        const double kk = E[(int64_t)ic.elem * mat_stride + b * mat_bmul] * A[(int64_t)ic.elem * mat_stride + b * mat_bmul] * g.z;
        dx += kk * (DIM == 2 ? g.x * g.x : 1.0);
        dy += kk * g.y * g.y;
    }
    const int64_t d0 = n * DIM;
    dinv[d0 * B + b] = (dof_free[d0] && dx != 0.0) ? 1.0 / dx : 0.0;
    if (DIM == 2) dinv[(d0 + 1) * B + b] = (dof_free[d0 + 1] && dy != 0.0) ? 1.0 / dy : 0.0;
}

// r = mask*rhs, x = 0, p = z = dinv*r; partial sums of r.z and rhs.rhs
__global__ void cg_init_kernel(CgVec v, const double* __restrict__ rhs, const double* __restrict__ dinv,
                               double* __restrict__ x, double* __restrict__ r, double* __restrict__ p,
                               double* __restrict__ part_rz, double* __restrict__ part_bb) {
    const int64_t b = (int64_t)blockIdx.y * CG_BX + threadIdx.x;
    double rz = 0.0, bb = 0.0;
    for (int t = 0; t < CG_RPT; ++t) {
        const int64_t d = ((int64_t)blockIdx.x * CG_BY + threadIdx.y) * CG_RPT + t;
        if (d < v.ndof && b < v.B) {
            const int64_t q = d * v.B + b;
            const double rr = v.dof_free[d] ? rhs[q] : 0.0;
            const double z = dinv[q] * rr;
            x[q] = 0.0;
            r[q] = rr;
            p[q] = z;
            rz += rr * z;
            bb += rr * rr;
        }
    }
    cg_block_reduce(rz, part_rz, v.B, b);
    cg_block_reduce(bb, part_bb, v.B, b);
}

__global__ void cg_dot_kernel(CgVec v, const double* __restrict__ p, const double* __restrict__ Ap,
                              double* __restrict__ part) {
    const int64_t b = (int64_t)blockIdx.y * CG_BX + threadIdx.x;
    double acc = 0.0;
    for (int t = 0; t < CG_RPT; ++t) {
        const int64_t d = ((int64_t)blockIdx.x * CG_BY + threadIdx.y) * CG_RPT + t;
        if (d < v.ndof && b < v.B && v.dof_free[d]) acc += p[d * v.B + b] * Ap[d * v.B + b];
    }
    cg_block_reduce(acc, part, v.B, b);
}

// x += alpha p; r -= alpha Ap (free rows); partial sums of r.z (z = dinv r) and r.r
__global__ void cg_update_kernel(CgVec v, const double* __restrict__ rz, const double* __restrict__ pAp,
                                 const double* __restrict__ p, const double* __restrict__ Ap,
                                 const double* __restrict__ dinv, double* __restrict__ x, double* __restrict__ r,
                                 double* __restrict__ part_rz, double* __restrict__ part_rr) {
    const int64_t b = (int64_t)blockIdx.y * CG_BX + threadIdx.x;
    double alpha = 0.0;
    if (b < v.B && pAp[b] != 0.0) alpha = rz[b] / pAp[b];
    double s_rz = 0.0, s_rr = 0.0;
    for (int t = 0; t < CG_RPT; ++t) {
        const int64_t d = ((int64_t)blockIdx.x * CG_BY + threadIdx.y) * CG_RPT + t;
        if (d < v.ndof && b < v.B && v.dof_free[d]) {
            const int64_t q = d * v.B + b;
            x[q] = fma(alpha, p[q], x[q]);
            const double rr = fma(-alpha, Ap[q], r[q]);
            r[q] = rr;
            s_rz += rr * (dinv[q] * rr);
            s_rr += rr * rr;
        }
    }
    cg_block_reduce(s_rz, part_rz, v.B, b);
    cg_block_reduce(s_rr, part_rr, v.B, b);
}

// p = z + (rz_new / rz) p
__global__ void cg_direction_kernel(CgVec v, const double* __restrict__ rz_new, const double* __restrict__ rz,
                                    const double* __restrict__ r, const double* __restrict__ dinv,
                                    double* __restrict__ p) {
    const int64_t b = (int64_t)blockIdx.y * CG_BX + threadIdx.x;
    double beta = 0.0;
    if (b < v.B && rz[b] != 0.0) beta = rz_new[b] / rz[b];
    for (int t = 0; t < CG_RPT; ++t) {
        const int64_t d = ((int64_t)blockIdx.x * CG_BY + threadIdx.y) * CG_RPT + t;
        if (d < v.ndof && b < v.B) {
            const int64_t q = d * v.B + b;
            p[q] = v.dof_free[d] ? fma(beta, p[q], dinv[q] * r[q]) : 0.0;
        }
    }
}

// out[b] = sum_rows part[row][b]: a CTA per 32 columns, 32 strided partial sums per column folded in a
// fixed order (a single thread per column walking thousands of rows serially dominated the CG iteration)
constexpr int CS_Y = 32;
__global__ void __launch_bounds__(32 * CS_Y) cg_colsum_kernel(const double* __restrict__ part, int64_t rows, int64_t B,
                                                              double* __restrict__ out) {
    __shared__ double s[CS_Y][33];
    const int64_t b = (int64_t)blockIdx.x * 32 + threadIdx.x;
    double acc = 0.0;
    if (b < B)
        for (int64_t r = threadIdx.y; r < rows; r += CS_Y) acc += part[r * B + b];
    s[threadIdx.y][threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.y == 0 && b < B) {
        double t = 0.0;
        for (int y = 0; y < CS_Y; ++y) t += s[y][threadIdx.x];
        out[b] = t;
    }
}

}  // namespace

extern "C" int64_t pf_cg_work_len(const pf_plan* plan, int64_t B) {
    if (!plan || B < 1) return -1;
    const int64_t tiles = (plan->ndof + CG_TILE_ROWS - 1) / CG_TILE_ROWS;
    return 4 * plan->ndof * B + 2 * tiles * B + 8 * B;
}

extern "C" int pf_cg_solve(pf_plan* plan, int kind, int64_t B, const double* u, const double* E, const double* A,
                           int mat_batched, const double* rhs, double* x, double rel_tol, int max_iters,
                           double* work, int64_t work_len, int32_t* iters_out, double* resid_out, void* stream) {
    int rc = pf_plan_activate(plan);
    if (rc) return rc;
    PF_REQUIRE(kind == PF_ELEM_LINEAR || plan->dim == 1, "pf_cg_solve: linear element only");
    PF_REQUIRE(B >= 1 && E && A && rhs && x && work, "pf_cg_solve: bad argument");
    PF_REQUIRE(work_len >= pf_cg_work_len(plan, B), "pf_cg_solve: work buffer too small");
    cudaStream_t st = pf_stream_of(stream);
    const int64_t nd = plan->ndof, tiles = (nd + CG_TILE_ROWS - 1) / CG_TILE_ROWS;
    double* r = work;
    double* p = r + nd * B;
    double* Ap = p + nd * B;
    double* dinv = Ap + nd * B;
    double* part0 = dinv + nd * B;
    double* part1 = part0 + tiles * B;
    double* rz = part1 + tiles * B;
    double* pAp = rz + B;
    double* rz_new = pAp + B;
    double* bb = rz_new + B;
    double* rr = bb + B;
    CgVec v{plan->d_dof_free, nd, B};
    dim3 blk(CG_BX, CG_BY), grd((unsigned)tiles, (unsigned)((B + CG_BX - 1) / CG_BX));
    const unsigned cb = (unsigned)((B + 31) / 32);
    const dim3 cblk(32, CS_Y);
    {
        dim3 dblk(32, 8), dgrd((unsigned)((plan->nnode + 7) / 8), (unsigned)((B + 31) / 32));
        const int64_t ms = mat_batched ? B : 1, mb = mat_batched ? 1 : 0;
        if (plan->dim == 1)
            cg_diag_kernel<1><<<dgrd, dblk, 0, st>>>(plan->d_inc_ptr, plan->d_inc, plan->d_inc_geo, plan->d_dof_free, E, A, ms, mb, plan->nnode, B, dinv);
        else
            cg_diag_kernel<2><<<dgrd, dblk, 0, st>>>(plan->d_inc_ptr, plan->d_inc, plan->d_inc_geo, plan->d_dof_free, E, A, ms, mb, plan->nnode, B, dinv);
    }
    cg_init_kernel<<<grd, blk, 0, st>>>(v, rhs, dinv, x, r, p, part0, part1);
    cg_colsum_kernel<<<cb, cblk, 0, st>>>(part0, tiles, B, rz);
    cg_colsum_kernel<<<cb, cblk, 0, st>>>(part1, tiles, B, bb);
    PF_CUDA_CHECK(cudaGetLastError());
    std::vector<double> h_rr(B), h_bb(B);
    PF_CUDA_CHECK(cudaMemcpyAsync(h_bb.data(), bb, B * sizeof(double), cudaMemcpyDeviceToHost, st));
    PF_CUDA_CHECK(cudaStreamSynchronize(st));
    int it = 0;
    double worst = 0.0;
    bool done = true;
    for (int64_t b = 0; b < B; ++b) done = done && h_bb[b] == 0.0;
    while (!done && it < max_iters) {
        rc = pf_tangent_matvec(plan, kind, B, u, E, A, mat_batched, p, Ap, stream);
        if (rc) return rc;
        cg_dot_kernel<<<grd, blk, 0, st>>>(v, p, Ap, part0);
        cg_colsum_kernel<<<cb, cblk, 0, st>>>(part0, tiles, B, pAp);
        cg_update_kernel<<<grd, blk, 0, st>>>(v, rz, pAp, p, Ap, dinv, x, r, part0, part1);
        cg_colsum_kernel<<<cb, cblk, 0, st>>>(part0, tiles, B, rz_new);
        cg_colsum_kernel<<<cb, cblk, 0, st>>>(part1, tiles, B, rr);
        cg_direction_kernel<<<grd, blk, 0, st>>>(v, rz_new, rz, r, dinv, p);
        std::swap(rz, rz_new);
        ++it;
        if (it % 8 == 0 || it == max_iters) {
            PF_CUDA_CHECK(cudaMemcpyAsync(h_rr.data(), rr, B * sizeof(double), cudaMemcpyDeviceToHost, st));
            PF_CUDA_CHECK(cudaStreamSynchronize(st));
            done = true;
            worst = 0.0;
            for (int64_t b = 0; b < B; ++b) {
                const double rel = h_bb[b] > 0.0 ? std::sqrt(h_rr[b] / h_bb[b]) : 0.0;
                worst = std::max(worst, rel);
                if (!(rel <= rel_tol)) done = false;
            }
        }
    }
    PF_CUDA_CHECK(cudaGetLastError());
    if (iters_out) *iters_out = it;
    if (resid_out) *resid_out = worst;
    return PF_OK;
}
