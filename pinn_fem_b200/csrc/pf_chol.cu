// Dense symmetric positive definite solve (fp64): the Newton step on K_ff (fem/core.py:32-35, fem/solver.py:464;
// K_ff of a constrained truss is SPD) and the Levenberg-Marquardt system of fem/nn_solver.py:266-277.
//
// Blocked right-looking Cholesky, 64-column panels, with panel look-ahead over two streams:
//
//   stream A (critical path)                               stream B (bulk)
//   diag_k    64 x 64 diagonal block, one CTA, one block barrier per column
//   trsm_k    rows below: X = A L^-T, one CTA per 64 rows ---> fwd_k    y_k = L_kk^-1 b_k,  b_i -= L_ik y_k (i below)
//   syrk_k(a) DMMA update of block column k+1 only         +-> syrk_k(b) DMMA update of the remaining lower tiles
//   diag_k+1, trsm_k+1 ...  <--- syrk_k+1(a) waits for syrk_k(b)
//
// so the serial part (diagonal block, panel solve) of panel k+1 runs while the bulk of panel k's trailing update is
// still in flight, and the forward substitution L y = b rides along for free.  The back substitution L^T x = y
// follows, one launch per panel (every CTA solves the 64 x 64 triangle redundantly, then updates its own columns).
// A is overwritten by L (lower triangle), b by x.  Everything is deterministic (fixed summation orders).
// No kernel writes a location another CTA of the same launch reads (round 1's panel kernel did: ADVICE r1).
#include <algorithm>
#include <cstdlib>

#include "pf_internal.h"

namespace {

constexpr int NB = 64;        // panel width
constexpr int TS = 64;        // trailing-update tile
constexpr int PLD = NB + 4;   // shared row stride of a panel tile: 68 doubles = 8 words (mod 32): conflict-free fragments

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

// Diagonal block A[k0:k0+nb, k0:k0+nb] = L L^T in place (lower triangle).  One CTA; one barrier per column: the
// rank-1 update of column c reads the UNSCALED column c (a_r a_cc / pivot), the scaled column goes to a separate
// output array, so nothing written in a step is read in the same step.
constexpr size_t kTwoBlocksSmem = 2 * NB * (NB + 1) * sizeof(double);  // 66.5 KB: dynamic (opt-in) shared memory
constexpr size_t kOneBlockSmem = NB * (NB + 1) * sizeof(double);

__global__ void __launch_bounds__(256) chol_diag_kernel(int n, int k0, int nb, double* __restrict__ A, int32_t* __restrict__ info) {
    // 16 x 16 threads; thread (ty, tx) keeps the elements (r = ty + 16 i, cc = tx + 16 j) of the block in registers.
    // Per column: its owners publish the current column through shared memory (double buffered by parity), one
    // barrier, everybody reads the pivot and the 4 + 4 column entries it needs and updates its 16 registers.
    // (Measured alternatives: shared-memory resident block 62 us, 128 threads with half a row each 75 us, this 45 us
    // per 64 x 64 block -- the kernel is a chain of 64 dependent steps, bound by instruction latency.)
    __shared__ double sC[2][NB];
    extern __shared__ __align__(16) double sm2[];
    double(*sL)[NB + 1] = reinterpret_cast<double(*)[NB + 1]>(sm2);
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    double a[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int r = ty + 16 * i, cc = tx + 16 * j;
            a[i][j] = (r < nb && cc <= r) ? A[(size_t)(k0 + r) * n + k0 + cc] : 0.0;
        }
    for (int c = 0; c < nb; ++c) {
        const int par = c & 1, jc = c >> 4;
        if (tx == (c & 15)) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (j == jc) sC[par][ty + 16 * i] = a[i][j];
        }
        __syncthreads();
        const double piv = sC[par][c];
        if (!(piv > 0.0)) {  // uniform: every thread reads the same pivot
            if (tid == 0 && *info == 0) *info = k0 + c + 1;
            return;
        }
        const double rs = rsqrt(piv), inv = rs * rs;  // one transcendental per column on the critical path
        double ar[4], ac[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            ar[i] = sC[par][ty + 16 * i];
            ac[i] = sC[par][tx + 16 * i];
        }
        if (tx == (c & 15)) {  // the owners of column c store L[:, c] (read again only after the loop)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = ty + 16 * i;
                if (r >= c) sL[r][c] = r == c ? piv * rs : ar[i] * rs;
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = ty + 16 * i;
            const double lr = ar[i] * inv;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int cc = tx + 16 * j;
                if (r > c && cc > c && cc <= r) a[i][j] = fma(-lr, ac[j], a[i][j]);
            }
        }
    }
    __syncthreads();
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int r0 = 0; r0 < NB; r0 += 8) {
        const int r = r0 + warp;
        if (r < nb) {
            if (lane <= r) A[(size_t)(k0 + r) * n + k0 + lane] = sL[r][lane];
            if (lane + 32 <= r) A[(size_t)(k0 + r) * n + k0 + lane + 32] = sL[r][lane + 32];
        }
    }
}

// Rows i0 .. i0 + 64 of the panel: X = A[i][k0:k0+NB] L^-T (L = the factored diagonal block).  Four lanes share a
// row: each keeps the x_d with d = lane (mod 4) in registers and forms its quarter of the dot product, two shuffles
// fold the quarters (one row solve is a chain of 64 dependent steps: 28 cycles per FMA with one lane per row).
__global__ void __launch_bounds__(4 * NB) chol_trsm_kernel(int n, int k0, double* __restrict__ A) {
    extern __shared__ __align__(16) double sm2[];
    double(*sL)[NB + 1] = reinterpret_cast<double(*)[NB + 1]>(sm2);
    double(*sP)[NB + 1] = reinterpret_cast<double(*)[NB + 1]>(sm2 + NB * (NB + 1));
    __shared__ double sInv[NB];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int i0 = k0 + NB + blockIdx.x * NB;
    const int rows = min(NB, n - i0);
#pragma unroll
    for (int r0 = 0; r0 < NB; r0 += 8) {  // all 32 loads of a thread in flight together
        const int r = r0 + warp;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int c = lane + 32 * h;
            sL[r][c] = c <= r ? A[(size_t)(k0 + r) * n + k0 + c] : 0.0;
            sP[r][c] = r < rows ? A[(size_t)(i0 + r) * n + k0 + c] : 0.0;
        }
    }
    __syncthreads();
    if (tid < NB) sInv[tid] = 1.0 / sL[tid][tid];
    __syncthreads();
    const int row = tid >> 2, sub = tid & 3;
    double x[NB / 4];
#pragma unroll
    for (int c = 0; c < NB; ++c) {
        double a0 = 0.0, a1 = 0.0;
#pragma unroll
        for (int k = 0; 4 * k < c; ++k) {  // d = 4 k + sub < c
            const int d = 4 * k + sub;
            const double term = d < c ? x[k] * sL[c][d] : 0.0;
            if (k & 1) a1 += term; else a0 += term;
        }
        double s = a0 + a1;
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        const double xc = (sP[row][c] - s) * sInv[c];
        if (sub == (c & 3)) x[c >> 2] = xc;  // no shared-memory store in here: the loads of later steps stay hoistable
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NB / 4; ++k) sP[row][4 * k + sub] = x[k];
    __syncthreads();
#pragma unroll
    for (int r0 = 0; r0 < NB; r0 += 8) {
        const int r = r0 + warp;
        if (r < rows) {
            A[(size_t)(i0 + r) * n + k0 + lane] = sP[r][lane];
            A[(size_t)(i0 + r) * n + k0 + lane + 32] = sP[r][lane + 32];
        }
    }
}

// Trailing update: C[i][j] -= sum_k P[i][k] P[j][k] for i >= j in [t0, n), P = A[:, k0:k0+NB].  CTA tile 64 x 64,
// 4 warps of 32 x 32 on DMMA.  blockIdx.x + bj0 = tile column, blockIdx.y = tile row; tiles above the diagonal exit.
__global__ void __launch_bounds__(128) chol_syrk_kernel(int n, int k0, int t0, int bj0, double* __restrict__ A) {
    const int bi = blockIdx.y, bj = blockIdx.x + bj0;
    if (bj > bi) return;
    extern __shared__ __align__(16) double sm[];
    double(*sI)[PLD] = reinterpret_cast<double(*)[PLD]>(sm);
    double(*sJ)[PLD] = reinterpret_cast<double(*)[PLD]>(sm + TS * PLD);
    const int i0 = t0 + bi * TS, j0 = t0 + bj * TS;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // both operand tiles in one round trip: 16-byte cp.async pieces (n is even whenever a tile is ragged-free;
    // rows beyond n and odd leading dimensions take the plain path)
    const bool vec = (n & 1) == 0 && i0 + TS <= n && j0 + TS <= n;
    if (vec) {
        for (int q = tid; q < TS * (NB / 2); q += 128) {
            const int r = q / (NB / 2), c = (q % (NB / 2)) * 2;
            const uint32_t di = (uint32_t)__cvta_generic_to_shared(&sI[r][c]), dj = (uint32_t)__cvta_generic_to_shared(&sJ[r][c]);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(di), "l"(A + (size_t)(i0 + r) * n + k0 + c) : "memory");
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dj), "l"(A + (size_t)(j0 + r) * n + k0 + c) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    } else {
        for (int q = tid; q < TS * NB; q += 128) {
            const int r = q / NB, c = q % NB;
            sI[r][c] = i0 + r < n ? A[(size_t)(i0 + r) * n + k0 + c] : 0.0;
            sJ[r][c] = j0 + r < n ? A[(size_t)(j0 + r) * n + k0 + c] : 0.0;
        }
    }
    const int wi = (warp >> 1) * 32, wj = (warp & 1) * 32;
    const int g = lane >> 2, t = lane & 3;
    // the C tile is read-modify-written at the end: pull it towards L2 now
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int i = i0 + wi + a * 8 + g;
        if (i < n && t == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(A + (size_t)i * n + j0 + wj));
        if (i < n && t == 2) asm volatile("prefetch.global.L2 [%0];" ::"l"(A + (size_t)i * n + min(j0 + wj + 16, n - 1)));
    }
    if (vec) asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    double acc[4][4][2];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
#pragma unroll 4
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

// 64 x 64 triangular solve by one warp (two rows per lane) out of shared memory: rhs[0..nb) -> solution, in place.
// TRANS = false: L x = rhs;  TRANS = true: L^T x = rhs.  inv = reciprocals of the diagonal.
template <bool TRANS>
__device__ __forceinline__ void warp_trisolve(const double (*sT)[NB + 1], const double* __restrict__ inv, int nb,
                                              double* __restrict__ rhs, int lane) {
    double r0 = lane < nb ? rhs[lane] : 0.0, r1 = lane + 32 < nb ? rhs[lane + 32] : 0.0;
    if (!TRANS) {
        for (int c = 0; c < nb; ++c) {
            const double l0 = sT[lane][c], l1 = sT[lane + 32][c];  // zero above the diagonal / beyond nb
            const double xc = __shfl_sync(0xffffffffu, c < 32 ? r0 : r1, c & 31) * inv[c];
            r0 = lane == c ? xc : (lane > c ? fma(-l0, xc, r0) : r0);
            r1 = lane + 32 == c ? xc : (lane + 32 > c ? fma(-l1, xc, r1) : r1);
        }
    } else {
        for (int c = nb - 1; c >= 0; --c) {
            const double l0 = sT[c][lane], l1 = sT[c][lane + 32];
            const double xc = __shfl_sync(0xffffffffu, c < 32 ? r0 : r1, c & 31) * inv[c];
            r0 = lane == c ? xc : (lane < c ? fma(-l0, xc, r0) : r0);
            r1 = lane + 32 == c ? xc : (lane + 32 < c ? fma(-l1, xc, r1) : r1);
        }
    }
    if (lane < nb) rhs[lane] = r0;
    if (lane + 32 < nb) rhs[lane + 32] = r1;
}

// stage the diagonal block (zero outside the lower triangle and beyond nb) and the reciprocals of its diagonal
__device__ __forceinline__ void load_diag_block(double (*sT)[NB + 1], double* __restrict__ sInv, const double* __restrict__ A,
                                                int n, int k0, int nb) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;  // 128 threads
#pragma unroll
    for (int r0 = 0; r0 < NB; r0 += 4) {  // all 32 loads of a thread in flight together
        const int r = r0 + warp;
        sT[r][lane] = (r < nb && lane <= r) ? A[(size_t)(k0 + r) * n + k0 + lane] : 0.0;
        sT[r][lane + 32] = (r < nb && lane + 32 <= r) ? A[(size_t)(k0 + r) * n + k0 + lane + 32] : 0.0;
    }
    __syncthreads();
    if (threadIdx.x < NB) sInv[threadIdx.x] = threadIdx.x < nb ? 1.0 / sT[threadIdx.x][threadIdx.x] : 0.0;
}

// Forward substitution step of panel k: y_k = L_kk^-1 b_k (every CTA, redundantly; CTA 0 stores it into `sol`),
// then b_i -= sum_c L[i][k0+c] y_k[c] for this CTA's rows below the panel.  b[k0..k0+nb) is only read here.
__global__ void __launch_bounds__(128) chol_fwd_kernel(int n, int k0, int nb, const double* __restrict__ A,
                                                       double* __restrict__ b, double* __restrict__ sol) {
    __shared__ double sT[NB][NB + 1];
    __shared__ double sy[NB], sInv[NB];
    const int tid = threadIdx.x;
    load_diag_block(sT, sInv, A, n, k0, nb);
    if (tid < nb) sy[tid] = b[k0 + tid];
    __syncthreads();
    if (tid < 32) warp_trisolve<false>(sT, sInv, nb, sy, tid);
    __syncthreads();
    if (blockIdx.x == 0 && tid < nb) sol[k0 + tid] = sy[tid];
    const int i = k0 + nb + blockIdx.x * 128 + tid;
    if (i < n) {
        const double* row = A + (size_t)i * n + k0;  // rows below exist only under full panels: nb == NB
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
        for (int c = 0; c < NB; c += 4) {
            a0 = fma(row[c], sy[c], a0);
            a1 = fma(row[c + 1], sy[c + 1], a1);
            a2 = fma(row[c + 2], sy[c + 2], a2);
            a3 = fma(row[c + 3], sy[c + 3], a3);
        }
        b[i] -= (a0 + a1) + (a2 + a3);
    }
}

// Back substitution step of panel k: x_k = L_kk^-T y_k (every CTA; CTA 0 stores it into `sol`), then
// y_j -= sum_c L[k0+c][j] x_k[c] for this CTA's columns j < k0 (rows of L: coalesced).  y[k0..k0+nb) is only read.
__global__ void __launch_bounds__(128) chol_bwd_kernel(int n, int k0, int nb, const double* __restrict__ A,
                                                       double* __restrict__ y, double* __restrict__ sol) {
    __shared__ double sT[NB][NB + 1];
    __shared__ double sx[NB], sInv[NB];
    const int tid = threadIdx.x;
    load_diag_block(sT, sInv, A, n, k0, nb);
    if (tid < nb) sx[tid] = y[k0 + tid];
    __syncthreads();
    if (tid < 32) warp_trisolve<true>(sT, sInv, nb, sx, tid);
    __syncthreads();
    if (blockIdx.x == 0 && tid < nb) sol[k0 + tid] = sx[tid];
    const int j = blockIdx.x * 128 + tid;
    if (j < k0) {
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll 4
        for (int c = 0; c < NB; c += 4) {  // rows beyond nb of the last panel do not exist: guarded
            if (c < nb) a0 = fma(A[(size_t)(k0 + c) * n + j], sx[c], a0);
            if (c + 1 < nb) a1 = fma(A[(size_t)(k0 + c + 1) * n + j], sx[c + 1], a1);
            if (c + 2 < nb) a2 = fma(A[(size_t)(k0 + c + 2) * n + j], sx[c + 2], a2);
            if (c + 3 < nb) a3 = fma(A[(size_t)(k0 + c + 3) * n + j], sx[c + 3], a3);
        }
        y[j] -= (a0 + a1) + (a2 + a3);
    }
}

// second stream + events of one solve; everything is joined back into the caller's stream before returning
struct Lookahead {
    cudaStream_t sb = nullptr;
    cudaEvent_t ev_trsm = nullptr, ev_bulk = nullptr, ev_join = nullptr;
    int open() {
        PF_CUDA_CHECK(cudaStreamCreateWithFlags(&sb, cudaStreamNonBlocking));
        PF_CUDA_CHECK(cudaEventCreateWithFlags(&ev_trsm, cudaEventDisableTiming));
        PF_CUDA_CHECK(cudaEventCreateWithFlags(&ev_bulk, cudaEventDisableTiming));
        PF_CUDA_CHECK(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
        return PF_OK;
    }
    ~Lookahead() {
        if (ev_trsm) cudaEventDestroy(ev_trsm);
        if (ev_bulk) cudaEventDestroy(ev_bulk);
        if (ev_join) cudaEventDestroy(ev_join);
        if (sb) cudaStreamDestroy(sb);  // work already enqueued completes; the caller's stream waits on ev_join
    }
};

}  // namespace

// A dev [n][n] row-major symmetric positive definite (only the lower triangle is read; overwritten by L),
// b dev [n] (overwritten by x), info dev int32 [1]: 0, or k > 0 when the k-th pivot is not positive.
extern "C" int pf_solve_spd(int64_t n64, double* A, double* b, int32_t* info, void* stream) {
    PF_REQUIRE(n64 >= 1 && A && b && info, "pf_solve_spd: bad argument");
    PF_REQUIRE(n64 < (1 << 15), "pf_solve_spd: n too large (%lld)", (long long)n64);
    const int n = (int)n64;
    cudaStream_t sa = pf_stream_of(stream);
    PF_CUDA_CHECK(cudaMemsetAsync(info, 0, sizeof(int32_t), sa));
    pf_keep_pool_cached();
    double* sol = nullptr;  // solution blocks (the substitution kernels never write what other CTAs still read)
    PF_CUDA_CHECK(cudaMallocAsync((void**)&sol, (size_t)n * sizeof(double), sa));
    constexpr size_t kSyrkSmem = 2 * TS * PLD * sizeof(double);
    PF_CUDA_CHECK(cudaFuncSetAttribute(chol_syrk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSyrkSmem));
    PF_CUDA_CHECK(cudaFuncSetAttribute(chol_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTwoBlocksSmem));
    PF_CUDA_CHECK(cudaFuncSetAttribute(chol_trsm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTwoBlocksSmem));
    static const int env_one = getenv("PF_CHOL_ONE_STREAM") ? atoi(getenv("PF_CHOL_ONE_STREAM")) : 0;
    const bool two_streams = n > 4 * NB && !env_one;  // small systems: one stream, nothing worth overlapping
    Lookahead la;
    if (two_streams) {
        int rc = la.open();
        if (rc) return rc;
        PF_CUDA_CHECK(cudaEventRecord(la.ev_join, sa));
        PF_CUDA_CHECK(cudaStreamWaitEvent(la.sb, la.ev_join, 0));  // B starts after the caller's earlier work
    }
    cudaStream_t sb = two_streams ? la.sb : sa;
    bool bulk_pending = false;
    for (int k0 = 0; k0 < n; k0 += NB) {
        const int nb = std::min(NB, n - k0);
        const int below = n - k0 - nb;
        chol_diag_kernel<<<1, 256, kOneBlockSmem, sa>>>(n, k0, nb, A, info);
        if (below > 0) chol_trsm_kernel<<<(below + NB - 1) / NB, 4 * NB, kTwoBlocksSmem, sa>>>(n, k0, A);  // then nb == NB
        if (two_streams) {
            PF_CUDA_CHECK(cudaEventRecord(la.ev_trsm, sa));
            PF_CUDA_CHECK(cudaStreamWaitEvent(sb, la.ev_trsm, 0));
        }
        chol_fwd_kernel<<<1 + (below + 127) / 128, 128, 0, sb>>>(n, k0, nb, A, b, sol);
        if (below > 0) {
            const int tiles = (below + TS - 1) / TS;
            // (a) the next panel's block column first, on the critical path; it overlaps tiles the previous bulk
            // update may still be writing, so it waits for that one
            if (two_streams && bulk_pending) PF_CUDA_CHECK(cudaStreamWaitEvent(sa, la.ev_bulk, 0));
            chol_syrk_kernel<<<dim3(1, tiles), 128, kSyrkSmem, sa>>>(n, k0, k0 + nb, 0, A);
            // (b) the rest of the trailing matrix, overlapped with the next panel's diagonal block and panel solve
            if (tiles > 1) {
                chol_syrk_kernel<<<dim3(tiles - 1, tiles), 128, kSyrkSmem, sb>>>(n, k0, k0 + nb, 1, A);
                if (two_streams) {
                    PF_CUDA_CHECK(cudaEventRecord(la.ev_bulk, sb));
                    bulk_pending = true;
                }
            }
        }
    }
    if (two_streams) {  // join: the back substitution needs all of L and y
        PF_CUDA_CHECK(cudaEventRecord(la.ev_join, sb));
        PF_CUDA_CHECK(cudaStreamWaitEvent(sa, la.ev_join, 0));
    }
    PF_CUDA_CHECK(cudaMemcpyAsync(b, sol, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, sa));  // y
    const int npanel = (n + NB - 1) / NB;
    for (int p = npanel - 1; p >= 0; --p) {
        const int k0 = p * NB, nb = std::min(NB, n - k0);
        chol_bwd_kernel<<<std::max(1, (k0 + 127) / 128), 128, 0, sa>>>(n, k0, nb, A, b, sol);
    }
    PF_CUDA_CHECK(cudaMemcpyAsync(b, sol, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, sa));  // x
    const cudaError_t le = cudaGetLastError();
    PF_CUDA_CHECK(cudaFreeAsync(sol, sa));
    PF_CUDA_CHECK(le);
    return PF_OK;
}
