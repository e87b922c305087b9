// Dense symmetric positive definite solve (fp64) for the Levenberg-Marquardt system
// (J^T J + d I) dx = -J^T R of fem/nn_solver.py:266-277, which is SPD by construction.
//
// LU with partial pivoting (pf_solve.cu) serialises on a single-CTA panel (pivot search + row swaps):
// 1 ms per 32-column panel at n = 4096.  Cholesky needs no pivoting, so every step is parallel:
//   per 32-column panel   chol_panel_kernel   every CTA factors the 32 x 32 diagonal block in shared memory
//                                             (redundantly, 32 steps) and solves its own 128 rows of the panel
//                         chol_syrk_kernel    trailing update C -= P P^T on the fp64 tensor pipe (DMMA),
//                                             lower 64 x 64 tiles only
//   then                  chol_trisolve_kernel  L y = b and L^T x = y, one CTA walking the panels
// A is overwritten by L (lower triangle), b by x.  Everything is deterministic (fixed summation orders).
#include <algorithm>

#include "pf_internal.h"

namespace {

constexpr int NB = 32;        // panel width
constexpr int PROWS = 128;    // panel rows per CTA in the solve
constexpr int TS = 64;        // trailing-update tile
constexpr int PLD = NB + 4;   // shared row stride of a panel tile: 36 doubles = 8 words (mod 32), conflict free

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

// Factor the nb x nb diagonal block held in sL (row stride NB + 1, lower triangle) in place; all 128 threads
// take part.  Returns through *bad the 1-based local index of the first non-positive pivot (0 = fine).
__device__ void factor_diag(double (*sL)[NB + 1], int nb, int* bad) {
    const int tid = threadIdx.x;
    for (int c = 0; c < nb; ++c) {
        __syncthreads();
        const double piv = sL[c][c];
        if (!(piv > 0.0)) {
            if (tid == 0 && *bad == 0) *bad = c + 1;
            return;  // uniform: every thread reads the same pivot
        }
        const double d = sqrt(piv);
        __syncthreads();
        if (tid == 0) sL[c][c] = d;
        for (int r = c + 1 + tid; r < nb; r += blockDim.x) sL[r][c] /= d;
        __syncthreads();
        // rank-1 update of the remaining lower triangle
        const int m = nb - c - 1;
        for (int q = tid; q < m * m; q += blockDim.x) {
            const int r = c + 1 + q / m, cc = c + 1 + q % m;
            if (cc <= r) sL[r][cc] -= sL[r][c] * sL[cc][c];
        }
    }
    __syncthreads();
}

// Panel step: diagonal block + the rows below it.  grid.x = 1 + ceil((n - k0 - nb) / PROWS).  Every CTA factors the
// (unfactored) diagonal block it reads from A; CTA 0 stores the factor into the scratch `Ld` -- NOT into A: a CTA
// that is scheduled late must still find the unfactored block there -- and reports a non-positive pivot.  The next
// launch in stream order (chol_syrk_kernel, or chol_diag_store_kernel for the last panel) copies Ld into A.
__global__ void __launch_bounds__(PROWS) chol_panel_kernel(int n, int k0, int nb, double* __restrict__ A,
                                                           double* __restrict__ Ld, int32_t* __restrict__ info) {
    __shared__ double sL[NB][NB + 1];
    __shared__ double sP[PROWS][NB + 1];
    __shared__ int bad;
    const int tid = threadIdx.x;
    if (tid == 0) bad = 0;
    for (int q = tid; q < nb * nb; q += PROWS) {
        const int r = q / nb, c = q % nb;
        sL[r][c] = c <= r ? A[(size_t)(k0 + r) * n + k0 + c] : 0.0;
    }
    __syncthreads();
    factor_diag(sL, nb, &bad);
    __syncthreads();
    if (bad) {
        if (blockIdx.x == 0 && tid == 0 && *info == 0) *info = k0 + bad;
        return;
    }
    if (blockIdx.x == 0) {
        for (int q = tid; q < nb * nb; q += PROWS) Ld[q] = sL[q / nb][q % nb];
        return;
    }
    // rows i0 .. i0 + PROWS of the panel: X = A[i][k0:k0+nb] L^-T, staged through shared memory so that
    // global accesses are whole 256-byte row pieces
    const int i0 = k0 + nb + (blockIdx.x - 1) * PROWS;
    const int rows = min(PROWS, n - i0);
    for (int q = tid; q < rows * nb; q += PROWS) {
        const int r = q / nb, c = q % nb;
        sP[r][c] = A[(size_t)(i0 + r) * n + k0 + c];
    }
    __syncthreads();
    if (tid < rows) {
        double x[NB];
#pragma unroll
        for (int c = 0; c < NB; ++c) {
            if (c < nb) {
                double acc = sP[tid][c];
#pragma unroll
                for (int d = 0; d < NB; ++d)
                    if (d < c) acc = fma(-x[d], sL[c][d], acc);
                x[c] = acc / sL[c][c];
            }
        }
#pragma unroll
        for (int c = 0; c < NB; ++c)
            if (c < nb) sP[tid][c] = x[c];
    }
    __syncthreads();
    for (int q = tid; q < rows * nb; q += PROWS) {
        const int r = q / nb, c = q % nb;
        A[(size_t)(i0 + r) * n + k0 + c] = sP[r][c];
    }
}

// Trailing update: C[i][j] -= sum_k P[i][k] P[j][k] for i >= j in [t0, n), P = A[:, k0:k0+NB] (full panels only;
// the ragged last panel has no trailing matrix).  CTA tile 64 x 64, 4 warps of 32 x 32, lower tiles only.
__device__ __forceinline__ void store_diag(int n, int k0, int nb, const double* __restrict__ Ld, double* __restrict__ A) {
    for (int q = threadIdx.x; q < nb * nb; q += blockDim.x) {
        const int r = q / nb, c = q % nb;
        if (c <= r) A[(size_t)(k0 + r) * n + k0 + c] = Ld[q];
    }
}

__global__ void __launch_bounds__(128) chol_diag_store_kernel(int n, int k0, int nb, const double* __restrict__ Ld,
                                                              double* __restrict__ A) {
    store_diag(n, k0, nb, Ld, A);
}

__global__ void __launch_bounds__(128) chol_syrk_kernel(int n, int k0, int t0, const double* __restrict__ Ld,
                                                        double* __restrict__ A) {
    const int bi = blockIdx.y, bj = blockIdx.x;
    if (bj > bi) return;
    // the panel kernel has finished (stream order): its factored diagonal block goes into A now; no tile of this
    // launch touches those entries
    if (bi == 0 && bj == 0) store_diag(n, k0, NB, Ld, A);
    __shared__ double sI[TS][PLD];
    __shared__ double sJ[TS][PLD];
    const int i0 = t0 + bi * TS, j0 = t0 + bj * TS;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int q = tid; q < TS * NB; q += 128) {
        const int r = q / NB, c = q % NB;
        sI[r][c] = i0 + r < n ? A[(size_t)(i0 + r) * n + k0 + c] : 0.0;
        sJ[r][c] = j0 + r < n ? A[(size_t)(j0 + r) * n + k0 + c] : 0.0;
    }
    __syncthreads();
    const int wi = (warp >> 1) * 32, wj = (warp & 1) * 32;
    const int g = lane >> 2, t = lane & 3;
    double acc[4][4][2];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
#pragma unroll
    for (int k4 = 0; k4 < NB; k4 += 4) {
        double af[4], bf[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) af[a] = sI[wi + a * 8 + g][k4 + t];  // A[row g][k t] = P[i][k]
#pragma unroll
        for (int b = 0; b < 4; ++b) bf[b] = sJ[wj + b * 8 + g][k4 + t];  // B[k t][col g] = P[j][k]
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) dmma(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int i = i0 + wi + a * 8 + g, j = j0 + wj + b * 8 + t * 2 + h;
                if (i < n && j <= i) A[(size_t)i * n + j] -= acc[a][b][h];
            }
}

// L y = b (BACK = false) or L^T x = y (BACK = true), in place in b.  One CTA of 32 warps walks the panels:
// warp w owns row / column w of the current 32-block for the dot products with the part already solved,
// then warp 0 finishes the 32 x 32 triangle out of shared memory.
template <bool BACK>
__global__ void __launch_bounds__(1024) chol_trisolve_kernel(int n, const double* __restrict__ A, double* __restrict__ b) {
    __shared__ double sT[NB][NB + 1];
    __shared__ double sR[NB][33];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int npanel = (n + NB - 1) / NB;
    for (int pp = 0; pp < npanel; ++pp) {
        const int p = BACK ? npanel - 1 - pp : pp;
        const int k0 = p * NB, nb = min(NB, n - k0);
        // diagonal block (lower triangle, L[k0+r][k0+c])
        for (int q = tid; q < nb * nb; q += 1024) {
            const int r = q / nb, c = q % nb;
            sT[r][c] = c <= r ? A[(size_t)(k0 + r) * n + k0 + c] : 0.0;
        }
        if (!BACK) {
            // s_r = sum_{j < k0} L[k0 + r][j] y[j]: warp r, lanes stride along the row (coalesced)
            double acc = 0.0;
            if (warp < nb)
                for (int j = lane; j < k0; j += 32) acc = fma(A[(size_t)(k0 + warp) * n + j], b[j], acc);
            sR[warp][lane] = acc;
        } else {
            // s_c = sum_{i >= k0 + nb} L[i][k0 + c] x[i]: lane c, warps stride over the rows (256-byte row pieces)
            double acc = 0.0;
            if (lane < nb)
                for (int i = k0 + nb + warp; i < n; i += 32) acc = fma(A[(size_t)i * n + k0 + lane], b[i], acc);
            sR[lane][warp] = acc;
        }
        __syncthreads();
        if (warp == 0) {
            double s = 0.0;
            if (lane < nb)
                for (int q = 0; q < 32; ++q) s += sR[lane][q];
            double rhs = lane < nb ? b[k0 + lane] - s : 0.0;
            if (!BACK) {
                for (int c = 0; c < nb; ++c) {
                    const double xc = __shfl_sync(0xffffffffu, rhs, c) / sT[c][c];
                    if (lane == c) rhs = xc;
                    else if (lane > c && lane < nb) rhs = fma(-sT[lane][c], xc, rhs);
                }
            } else {
                for (int c = nb - 1; c >= 0; --c) {
                    const double xc = __shfl_sync(0xffffffffu, rhs, c) / sT[c][c];
                    if (lane == c) rhs = xc;
                    else if (lane < c) rhs = fma(-sT[c][lane], xc, rhs);
                }
            }
            if (lane < nb) b[k0 + lane] = rhs;
        }
        __syncthreads();
    }
}

}  // namespace

// A dev [n][n] row-major symmetric positive definite (only the lower triangle is read; overwritten by L),
// b dev [n] (overwritten by x), info dev int32 [1]: 0, or k > 0 when the k-th pivot is not positive.
extern "C" int pf_solve_spd(int64_t n64, double* A, double* b, int32_t* info, void* stream) {
    PF_REQUIRE(n64 >= 1 && A && b && info, "pf_solve_spd: bad argument");
    PF_REQUIRE(n64 < (1 << 15), "pf_solve_spd: n too large (%lld)", (long long)n64);
    const int n = (int)n64;
    cudaStream_t st = pf_stream_of(stream);
    PF_CUDA_CHECK(cudaMemsetAsync(info, 0, sizeof(int32_t), st));
    pf_keep_pool_cached();
    double* Ld = nullptr;  // factored diagonal block of the current panel (see chol_panel_kernel)
    PF_CUDA_CHECK(cudaMallocAsync((void**)&Ld, NB * NB * sizeof(double), st));
    for (int k0 = 0; k0 < n; k0 += NB) {
        const int nb = std::min(NB, n - k0);
        const int below = n - k0 - nb;
        chol_panel_kernel<<<1 + (below + PROWS - 1) / PROWS, PROWS, 0, st>>>(n, k0, nb, A, Ld, info);
        if (below > 0) {  // then nb == NB
            const int tiles = (below + TS - 1) / TS;
            chol_syrk_kernel<<<dim3(tiles, tiles), 128, 0, st>>>(n, k0, k0 + nb, Ld, A);
        } else {
            chol_diag_store_kernel<<<1, 128, 0, st>>>(n, k0, nb, Ld, A);
        }
    }
    chol_trisolve_kernel<false><<<1, 1024, 0, st>>>(n, A, b);
    chol_trisolve_kernel<true><<<1, 1024, 0, st>>>(n, A, b);
    const cudaError_t le = cudaGetLastError();
    PF_CUDA_CHECK(cudaFreeAsync(Ld, st));
    PF_CUDA_CHECK(le);
    return PF_OK;
}
