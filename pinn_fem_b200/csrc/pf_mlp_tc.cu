// Material networks on the fp64 tensor pipe (DMMA, mma.sync.m8n8k4.f64) for large point sets.
//
// SimpleNN (examples/json/generic.py:118-142) wrapped by NNProperty.value
// (fem/properties.py:150-156): softplus(net(x)) * scale, evaluated at ~10^6 element
// centroids per call in the large-mesh PINN loop.  fp64 has no tcgen05 kind on sm_100a; the
// fp64 tensor path is DMMA.  Every dense contraction of the forward and reverse pass is a small
// GEMM with the point index as one dimension:
//   forward   z_l[t][o]   = sum_i a_l[i][t]  Wt_l[i][o]     M = points, N = out, K = in
//   backprop  e_l[t][i]   = sum_o D_l[o][t]  W_l[o][i]      M = points, N = in,  K = out
//   gradient  dW_l[o][i]  = sum_t D_l[o][t]  a_l[i][t]      M = out,    N = in + 1 (ones row -> bias), K = points
// A CTA (8 warps) is persistent over tiles of 128 points.  Each warp owns 16 points through the
// forward and backprop GEMMs (no block barrier); the gradient GEMMs span all 128 points, their output
// tiles are dealt to the warps and added into a per-CTA accumulator in shared memory, so the
// reduction order is fixed: tile order inside a CTA, CTA order across the grid (reduce_rows).
// Activations live in shared memory as [row][point] with row stride 132 doubles, weights with a
// row stride = 4 (mod 16) doubles: every DMMA fragment load is bank-conflict free.
#include "pf_internal.h"
#include "pf_mlp.cuh"

namespace {

constexpr int PTS = 128;      // points per tile
constexpr int PS = PTS + 4;   // activation row stride (doubles): 264 words = 8 (mod 32)
constexpr int kWarps = 8;              // warps per CTA: each owns PTS / kWarps points of a tile
constexpr int kThreads = kWarps * 32;
constexpr int WPTS = PTS / kWarps;     // points per warp (16)
constexpr int MTW = WPTS / 8;          // m8 tiles per warp (2)

__host__ __device__ inline int ceil4(int x) { return (x + 3) & ~3; }
__host__ __device__ inline int ceil8(int x) { return (x + 7) & ~7; }
// smallest stride >= n that is 4 (mod 16)
__host__ __device__ inline int wstride(int n) { return ((n + 11) / 16) * 16 + 4; }

struct TcPlan {  // shared-memory plan in doubles
    int wt_off[PF_MLP_MAX_LAYERS];   // Wt_l[ceil4(in_l)][WS]   (forward B operand; zero padded)
    int wn_off[PF_MLP_MAX_LAYERS];   // W_l[ceil4(w)][WSI]      (backprop B operand; l >= 1, backward only)
    int b_off[PF_MLP_MAX_LAYERS];    // bias_l[ceil8(w)]
    int wo_off, bo_off;              // output layer weights [ceil8(w)] (zero padded), bias
    int act_off[PF_MLP_MAX_LAYERS + 1];  // activation block l: rows_l x PS
    int rows[PF_MLP_MAX_LAYERS + 1];
    int d_off;                       // (unused)
    int dz_off;                      // PS
    int g_off;                       // n_params (backward only)
    int WS, WSI;
    int total;
};

__host__ __device__ inline TcPlan tc_plan(const PfMlpDesc& d, bool backward) {
    TcPlan s;
    const int w8 = ceil8(d.w);
    s.WS = wstride(w8);
    s.WSI = wstride(w8);
    int off = 0;
    for (int l = 0; l < d.L; ++l) {
        const int in = l == 0 ? d.in_dim : d.w;
        s.wt_off[l] = off;
        off += ceil4(in) * s.WS;
        s.wn_off[l] = off;
        if (backward && l > 0) off += ceil4(d.w) * s.WSI;
        s.b_off[l] = off;
        off += w8;
    }
    s.wo_off = off;
    off += w8;
    s.bo_off = off;
    off += 2;
    const int r0 = ceil8(d.in_dim + 1), rh = ceil8(d.w + 1);
    if (backward) {
        // The deltas D_l overwrite activation block l + 1 in place once it has been consumed, and the input block
        // keeps only its ceil4(in + 1) real rows (the gradient tiles mask the padding rows): 56 instead of 80 rows
        // for 3-20-20-1 = 73 KB instead of 103 KB per CTA, i.e. three CTAs (24 warps) per SM instead of two.
        for (int l = 0; l <= d.L; ++l) {
            s.rows[l] = l == 0 ? ceil4(d.in_dim + 1) : rh;
            s.act_off[l] = off;
            off += s.rows[l] * PS;
        }
        s.d_off = off;  // unused
    } else {  // forward only: two ping-pong blocks
        const int r = r0 > rh ? r0 : rh;
        for (int l = 0; l <= d.L; ++l) {
            s.rows[l] = r;
            s.act_off[l] = off + (l & 1) * r * PS;
        }
        off += 2 * r * PS;
        s.d_off = off;
    }
    s.dz_off = off;
    off += PS;
    s.g_off = off;
    if (backward) off += (d.n_params + 1) & ~1;
    s.total = off;
    return s;
}

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

// C[t][n] (+)= sum_k A[k][t] B[k][n] for the calling warp's points (MTW m-tiles) and one n-tile.
// A is [k][point] with stride PS, B is [k][n] with stride ldb.
__device__ __forceinline__ void warp_gemm_tile(const double* __restrict__ A, const double* __restrict__ Bm, int ldb,
                                               int K, int tw0, int n0, int g, int t4, double (&c)[MTW][2]) {
    for (int k0 = 0; k0 < K; k0 += 4) {
        const double b = Bm[(k0 + t4) * ldb + n0 + g];
        const double* ap = A + (k0 + t4) * PS + tw0 + g;
#pragma unroll
        for (int mt = 0; mt < MTW; ++mt) dmma(c[mt][0], c[mt][1], ap[mt * 8], b);
    }
}

// dW tile: C[o][i] = sum_t Dm[o][t] Am[i][t] over the 128 points of the tile, 4 interleaved partial sums
// (rows n0 + g >= arows of Am do not exist: they count as zero)
__device__ __forceinline__ void grad_tile(const double* __restrict__ Dm, const double* __restrict__ Am, int m0, int n0,
                                          int g, int t4, bool single_row, int arows, double& c0, double& c1) {
    double p[4][2] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};
    const double* dp = Dm + (single_row ? 0 : (m0 + g) * PS) + t4;
    const bool brow = n0 + g < arows;
    const double* ap = Am + (brow ? n0 + g : 0) * PS + t4;
    const bool arow = !single_row || g == 0;
#pragma unroll 2
    for (int t0 = 0; t0 < PTS; t0 += 16) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const double a = arow ? dp[t0 + 4 * q] : 0.0;
            const double bv = brow ? ap[t0 + 4 * q] : 0.0;
            dmma(p[q][0], p[q][1], a, bv);
        }
    }
    c0 = (p[0][0] + p[1][0]) + (p[2][0] + p[3][0]);
    c1 = (p[0][1] + p[1][1]) + (p[2][1] + p[3][1]);
}

template <bool BWD>
__global__ void __launch_bounds__(kThreads) mlp_tc_kernel(PfMlpDesc d, const double* __restrict__ theta, int64_t n,
                                                          const double* __restrict__ X,
                                                          const double* __restrict__ centroid, double load_factor,
                                                          double scale, int positive,
                                                          const double* __restrict__ g_out, double* __restrict__ out,
                                                          double* __restrict__ part) {
    extern __shared__ __align__(16) double sm[];
    const TcPlan s = tc_plan(d, BWD);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t4 = lane & 3;
    const int w = d.w, w8 = ceil8(w), L = d.L;

    // ---- stage the parameters (zero padded) and clear the activation blocks once ----
    for (int q = tid; q < s.total; q += kThreads) sm[q] = 0.0;
    __syncthreads();
    for (int l = 0; l < L; ++l) {
        const int in = l == 0 ? d.in_dim : w;
        const double* W = theta + d.w_off[l];
        for (int q = tid; q < w * in; q += kThreads) {
            const int o = q / in, i = q - o * in;
            const double v = W[q];
            sm[s.wt_off[l] + i * s.WS + o] = v;
            if (BWD && l > 0) sm[s.wn_off[l] + o * s.WSI + i] = v;
        }
        for (int o = tid; o < w; o += kThreads) sm[s.b_off[l] + o] = theta[d.b_off[l] + o];
    }
    for (int o = tid; o < w; o += kThreads) sm[s.wo_off + o] = theta[d.w_off[L] + o];
    if (tid == 0) sm[s.bo_off] = theta[d.b_off[L]];
    // ones row of every hidden block (bias gradients); when w is a multiple of 8 the forward epilogue
    // never touches it, otherwise it rewrites it every tile
    for (int l = 1; l <= L; ++l)
        for (int t = tid; t < PTS; t += kThreads) sm[s.act_off[l] + w * PS + t] = 1.0;
    __syncthreads();

    const int64_t ntiles = (n + PTS - 1) / PTS;
    const int tw0 = warp * WPTS;        // first point of the warp inside the tile
    const bool ptlane = lane < WPTS;     // per-point phases: lane i of a warp handles the warp's point i
    const int pt = tw0 + (ptlane ? lane : 0);
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t p = tile * PTS + pt;
        // ---- inputs: [load_factor, x_c(, y_c)] (sorted dict keys of fem/properties.py:116-125), ones row ----
        if (ptlane) {
            double* a0 = sm + s.act_off[0] + pt;
            if (X) {
                for (int i = 0; i < d.in_dim; ++i) a0[i * PS] = p < n ? X[p * d.in_dim + i] : 0.0;
            } else {
                a0[0] = load_factor;
                for (int i = 1; i < d.in_dim; ++i) a0[i * PS] = p < n ? centroid[p * (d.in_dim - 1) + (i - 1)] : 0.0;
            }
            a0[d.in_dim * PS] = 1.0;
        }
        __syncwarp();
        // ---- forward: the warp's 32 points through every hidden layer ----
        for (int l = 0; l < L; ++l) {
            const int K = ceil4(l == 0 ? d.in_dim : w);
            const double* A = sm + s.act_off[l];
            double* An = sm + s.act_off[l + 1];
            for (int n0 = 0; n0 < w8; n0 += 8) {
                double c[MTW][2];
                const double b0 = sm[s.b_off[l] + n0 + 2 * t4], b1 = sm[s.b_off[l] + n0 + 2 * t4 + 1];
#pragma unroll
                for (int mt = 0; mt < MTW; ++mt) {
                    c[mt][0] = b0;
                    c[mt][1] = b1;
                }
                warp_gemm_tile(A, sm + s.wt_off[l], s.WS, K, tw0, n0, g, t4, c);
#pragma unroll
                for (int mt = 0; mt < MTW; ++mt)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int o = n0 + 2 * t4 + h;
                        // row w carries the ones the bias gradient needs; rows beyond hold tanh(0) = 0
                        An[o * PS + tw0 + mt * 8 + g] = o == w ? 1.0 : pf_tanh(c[mt][h]);
                    }
            }
            __syncwarp();
        }
        // ---- output layer, softplus, dL/dz ----
        const double* aL = sm + s.act_off[L] + pt;
        double z = sm[s.bo_off];
        for (int i = 0; i < w; ++i) z = fma(sm[s.wo_off + i], aL[i * PS], z);
        if (!BWD) {
            if (ptlane && p < n) out[p] = pf_mlp_output(z, scale, positive);
            __syncwarp();
            continue;
        }
        const double dz = (ptlane && p < n) ? g_out[p] * pf_mlp_output_grad(z, scale, positive) : 0.0;
        if (ptlane) sm[s.dz_off + pt] = dz;
        double* gs = sm + s.g_off;
        // D_l lives in activation block l + 1, written in place once that block has been consumed
        auto dptr = [&](int l) { return sm + s.act_off[l + 1]; };
        __syncthreads();  // dz and a_L complete for all 128 points
        // output layer: dWo[i] = sum_t dz[t] a_L[i][t], dbo = sum_t dz[t] (ones row)
        for (int nt = warp; nt < w8 / 8 + (w8 == w ? 1 : 0); nt += kWarps) {
            double c0, c1;
            grad_tile(sm + s.dz_off, sm + s.act_off[L], 0, nt * 8, g, t4, true, s.rows[L], c0, c1);
            if (g == 0) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int i = nt * 8 + 2 * t4 + h;
                    const double v = h ? c1 : c0;
                    if (i < w)
                        gs[d.w_off[L] + i] += v;
                    else if (i == w)
                        gs[d.b_off[L]] += v;
                }
            }
        }
        __syncthreads();  // a_L consumed by every warp: it becomes D_{L-1}
        if (ptlane) {
            double* D = dptr(L - 1) + pt;
            for (int o = 0; o < w8; ++o) {
                const double a = D[o * PS];
                D[o * PS] = sm[s.wo_off + o] * dz * (1.0 - a * a);  // rows >= w: zero weight
            }
        }
        for (int l = L - 1; l >= 0; --l) {
            __syncthreads();  // D_l complete for all 128 points
            const int in = l == 0 ? d.in_dim : w;
            const double* D = dptr(l);
            double* A = sm + s.act_off[l];
            const int MT = w8 / 8, NT = ceil8(in + 1) / 8;
            for (int tl = warp; tl < MT * NT; tl += kWarps) {
                double c0, c1;
                const int mt = tl / NT, nt = tl - mt * NT;
                grad_tile(D, A, mt * 8, nt * 8, g, t4, false, s.rows[l], c0, c1);
                const int o = mt * 8 + g;
                if (o < w) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int i = nt * 8 + 2 * t4 + h;
                        const double v = h ? c1 : c0;
                        if (i < in)
                            gs[d.w_off[l] + o * in + i] += v;
                        else if (i == in)
                            gs[d.b_off[l] + o] += v;
                    }
                }
            }
            if (l > 0) {
                __syncthreads();  // a_l consumed as the A operand by every warp: it becomes D_{l-1}
                // D_{l-1}[i][t] = (sum_o W_l[o][i] D_l[o][t]) (1 - a_l[i][t]^2) for the warp's own points, in place;
                // the ones row (i = w) gives 1 - 1 = 0, padded columns have zero weights.
                for (int n0 = 0; n0 < w8; n0 += 8) {
                    double c[MTW][2];
#pragma unroll
                    for (int mt = 0; mt < MTW; ++mt) c[mt][0] = c[mt][1] = 0.0;
                    warp_gemm_tile(D, sm + s.wn_off[l], s.WSI, ceil4(w), tw0, n0, g, t4, c);
#pragma unroll
                    for (int mt = 0; mt < MTW; ++mt)
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int i = n0 + 2 * t4 + h;
                            const int t = tw0 + mt * 8 + g;
                            const double a = A[i * PS + t];
                            A[i * PS + t] = c[mt][h] * (1.0 - a * a);
                        }
                }
            }
        }
        __syncthreads();  // gradient tiles of layer 0 read block 0 of every point: done before the next tile's inputs
    }
    if (BWD) {
        __syncthreads();
        for (int q = tid; q < d.n_params; q += kThreads) part[(int64_t)blockIdx.x * d.n_params + q] = sm[s.g_off + q];
    }
}

}  // namespace

size_t pf_mlp_tc_smem(const PfMlpDesc& d, bool backward) { return (size_t)tc_plan(d, backward).total * sizeof(double); }

// Returns the grid size used (number of partial-gradient rows written to `part` when backward).
int pf_mlp_tc_launch(const PfMlpDesc& d, bool backward, const double* theta, int64_t n, const double* X,
                     const double* centroid, double load_factor, double scale, int positive, const double* g_out,
                     double* out, double* part, int grid, cudaStream_t st) {
    const size_t smem = pf_mlp_tc_smem(d, backward);
    if (backward) {
        PF_CUDA_CHECK(cudaFuncSetAttribute(mlp_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        mlp_tc_kernel<true><<<grid, kThreads, smem, st>>>(d, theta, n, X, centroid, load_factor, scale, positive, g_out,
                                                         nullptr, part);
    } else {
        PF_CUDA_CHECK(cudaFuncSetAttribute(mlp_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        mlp_tc_kernel<false><<<grid, kThreads, smem, st>>>(d, theta, n, X, centroid, load_factor, scale, positive,
                                                          nullptr, out, nullptr);
    }
    PF_CUDA_CHECK(cudaGetLastError());
    return PF_OK;
}

int pf_mlp_tc_grid(const PfMlpDesc& d, bool backward, int64_t n, int sm_count) {
    const size_t smem = pf_mlp_tc_smem(d, backward) + 1024;
    int per_sm = (int)((227 * 1024) / smem);
    if (per_sm < 1) return 0;  // does not fit: caller uses the generic kernels
    if (per_sm > 8) per_sm = 8;
    const int64_t ntiles = (n + PTS - 1) / PTS;
    const int64_t cap = (int64_t)sm_count * per_sm;
    return (int)(ntiles < cap ? ntiles : cap);
}
