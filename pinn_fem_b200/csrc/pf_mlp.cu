// Material networks (fp64): SimpleNN (examples/json/generic.py:118-142) wrapped by
// NNProperty.value (fem/properties.py:150-156): softplus(net(x)) * scale.
//
// One CTA handles a tile of points; the (tiny) weight set lives transposed and
// padded in shared memory, activations live in shared memory with the point
// index innermost (conflict-free), each thread owns one point in the forward
// and delta passes and a strided set of (out,in) weight pairs in the gradient
// pass.  All reductions run in a fixed order (ascending point inside a CTA,
// ascending CTA across the grid) so results are bitwise reproducible.
#include <cstdlib>

#include "pf_internal.h"
#include "pf_mlp.cuh"
#include "pf_mlp_frag.h"

namespace {

// PTS = points per CTA (128, 64 or 32: the largest whose tiles fit in shared memory);
// activation tiles use a padded row stride PTS + 1.

// shared-memory plan of one CTA (in doubles)
struct SmemPlan {
    int w_off[PF_MLP_MAX_LAYERS];  // transposed padded weights Wt[in][wp] of hidden layer l
    int b_off[PF_MLP_MAX_LAYERS];
    int wo_off, bo_off;            // output layer
    int act_off;                   // activations: rows = in_dim + L*wp (backward) or 2*max rows (forward)
    int delta_off;                 // 2 * wp rows (backward only)
    int dz_off;                    // PTS
    int total;
};

__host__ __device__ inline SmemPlan smem_plan(const PfMlpDesc& d, bool backward, int PTS) {
    const int PS = PTS + 1;
    SmemPlan s;
    int off = 0;
    for (int l = 0; l < d.L; ++l) {
        const int in = l == 0 ? d.in_dim : d.w;
        s.w_off[l] = off;
        off += in * d.wp;
        s.b_off[l] = off;
        off += d.wp;
    }
    s.wo_off = off;
    off += d.wp;
    s.bo_off = off;
    off += 2;  // keep 16-byte alignment of what follows
    s.act_off = off;
    const int maxrow = d.wp > d.in_dim ? d.wp : d.in_dim;
    off += (backward ? (d.in_dim + d.L * d.wp) : 2 * maxrow) * PS;
    s.delta_off = off;
    if (backward) off += 2 * d.wp * PS;
    s.dz_off = off;
    off += PTS;
    s.total = off;
    return s;
}

// Stage theta into shared memory: hidden weights transposed to [in][wp] with zero padding.
__device__ void stage_weights(const PfMlpDesc& d, const SmemPlan& s, const double* __restrict__ theta,
                              double* __restrict__ sm) {
    for (int l = 0; l < d.L; ++l) {
        const int in = l == 0 ? d.in_dim : d.w;
        const double* W = theta + d.w_off[l];
        const double* b = theta + d.b_off[l];
        for (int q = threadIdx.x; q < in * d.wp; q += blockDim.x) {
            const int i = q / d.wp, o = q % d.wp;
            sm[s.w_off[l] + q] = o < d.w ? W[o * in + i] : 0.0;
        }
        for (int o = threadIdx.x; o < d.wp; o += blockDim.x) sm[s.b_off[l] + o] = o < d.w ? b[o] : 0.0;
    }
    for (int o = threadIdx.x; o < d.wp; o += blockDim.x) sm[s.wo_off + o] = o < d.w ? theta[d.w_off[d.L] + o] : 0.0;
    if (threadIdx.x == 0) sm[s.bo_off] = theta[d.b_off[d.L]];
}

// One hidden layer for the calling thread's point: out[o][t] = tanh(b[o] + sum_i Wt[i][o] in[i][t]).
template <int PS, int NO>
__device__ __forceinline__ void hidden_block(const double* __restrict__ Wt, const double* __restrict__ b, int in, int wp,
                                             const double* __restrict__ a_in, double* __restrict__ a_out, int t, int o0) {
    double acc[NO];
#pragma unroll
    for (int j = 0; j < NO; ++j) acc[j] = b[o0 + j];
    for (int i = 0; i < in; ++i) {
        const double a = a_in[i * PS + t];
#pragma unroll
        for (int j = 0; j < NO; j += 2) {
            const double2 w = *reinterpret_cast<const double2*>(Wt + i * wp + o0 + j);
            acc[j] = fma(w.x, a, acc[j]);
            acc[j + 1] = fma(w.y, a, acc[j + 1]);
        }
    }
#pragma unroll
    for (int j = 0; j < NO; ++j) a_out[(o0 + j) * PS + t] = pf_tanh(acc[j]);
}

// One hidden layer for the calling thread's point: out[o][t] = tanh(b[o] + sum_i Wt[i][o] in[i][t]); eight
// outputs at a time (eight independent accumulator / tanh chains: the kernel is fp64-latency bound), then four.
template <int PS>
__device__ __forceinline__ void hidden_layer(const double* __restrict__ Wt, const double* __restrict__ b, int in,
                                             int wp, const double* __restrict__ a_in, double* __restrict__ a_out,
                                             int t) {
    int o0 = 0;
    for (; o0 + 8 <= wp; o0 += 8) hidden_block<PS, 8>(Wt, b, in, wp, a_in, a_out, t, o0);
    for (; o0 < wp; o0 += 4) hidden_block<PS, 4>(Wt, b, in, wp, a_in, a_out, t, o0);
}

template <int PS>
__device__ __forceinline__ void load_input(const PfMlpDesc& d, const double* __restrict__ X,
                                           const double* __restrict__ centroid, double load_factor, int64_t p,
                                           int64_t n, double* __restrict__ a0, int t) {
    if (X) {
        for (int i = 0; i < d.in_dim; ++i) a0[i * PS + t] = p < n ? X[p * d.in_dim + i] : 0.0;
    } else {  // [load_factor, x_c(, y_c)]: sorted dict keys of fem/properties.py:116-125
        a0[t] = load_factor;
        for (int i = 1; i < d.in_dim; ++i) a0[i * PS + t] = p < n ? centroid[p * (d.in_dim - 1) + (i - 1)] : 0.0;
    }
}

template <int PTS>
__global__ void __launch_bounds__(PTS) mlp_forward_kernel(PfMlpDesc d, const double* __restrict__ theta, int64_t n,
                                                          const double* __restrict__ X,
                                                          const double* __restrict__ centroid, double load_factor,
                                                          double scale, int positive, double* __restrict__ out) {
    constexpr int PS = PTS + 1;
    extern __shared__ __align__(16) double sm[];
    const SmemPlan s = smem_plan(d, false, PTS);
    stage_weights(d, s, theta, sm);
    const int t = threadIdx.x;
    const int64_t p = (int64_t)blockIdx.x * PTS + t;
    const int maxrow = d.wp > d.in_dim ? d.wp : d.in_dim;
    double* buf0 = sm + s.act_off;
    double* buf1 = buf0 + maxrow * PS;
    load_input<PS>(d, X, centroid, load_factor, p, n, buf0, t);
    __syncthreads();
    for (int l = 0; l < d.L; ++l) {
        hidden_layer<PS>(sm + s.w_off[l], sm + s.b_off[l], l == 0 ? d.in_dim : d.w, d.wp, buf0, buf1, t);
        double* tmp = buf0;
        buf0 = buf1;
        buf1 = tmp;
    }
    double z = sm[s.bo_off];
    for (int i = 0; i < d.w; ++i) z = fma(sm[s.wo_off + i], buf0[i * PS + t], z);
    if (p < n) out[p] = pf_mlp_output(z, scale, positive);
}

// mode 0: reduce parameter gradients over the CTA's points -> gpart[blockIdx.x][n_params]
// mode 1: per-point Jacobian rows                            -> jac[p][n_params]
template <int MODE, int PTS>
__global__ void __launch_bounds__(PTS) mlp_backward_kernel(PfMlpDesc d, const double* __restrict__ theta, int64_t n,
                                                           const double* __restrict__ X,
                                                           const double* __restrict__ centroid, double load_factor,
                                                           double scale, int positive,
                                                           const double* __restrict__ g_out,
                                                           double* __restrict__ dst) {
    constexpr int PS = PTS + 1;
    extern __shared__ __align__(16) double sm[];
    const SmemPlan s = smem_plan(d, true, PTS);
    stage_weights(d, s, theta, sm);
    const int t = threadIdx.x;
    const int64_t p0 = (int64_t)blockIdx.x * PTS;
    const int64_t p = p0 + t;
    const int npts = (int)((n - p0) < PTS ? (n - p0) : PTS);
    double* acts = sm + s.act_off;  // layer 0 rows: in_dim; layer l>=1 rows: wp each
    auto act = [&](int l) { return acts + (l == 0 ? 0 : (d.in_dim + (l - 1) * d.wp)) * PS; };
    load_input<PS>(d, X, centroid, load_factor, p, n, act(0), t);
    __syncthreads();
    for (int l = 0; l < d.L; ++l)
        hidden_layer<PS>(sm + s.w_off[l], sm + s.b_off[l], l == 0 ? d.in_dim : d.w, d.wp, act(l), act(l + 1), t);
    double z = sm[s.bo_off];
    for (int i = 0; i < d.w; ++i) z = fma(sm[s.wo_off + i], act(d.L)[i * PS + t], z);
    double dz = 0.0;
    if (p < n) dz = (MODE == 0 ? g_out[p] : 1.0) * pf_mlp_output_grad(z, scale, positive);
    sm[s.dz_off + t] = dz;
    double* D = sm + s.delta_off;
    double* Dn = D + d.wp * PS;
    // delta of the last hidden layer
    for (int o = 0; o < d.wp; ++o) {
        const double a = act(d.L)[o * PS + t];
        D[o * PS + t] = sm[s.wo_off + o] * dz * (1.0 - a * a);
    }
    __syncthreads();

    double* gdst = MODE == 0 ? dst + (int64_t)blockIdx.x * d.n_params : nullptr;
    // output layer parameters
    for (int q = threadIdx.x; q <= d.w; q += blockDim.x) {
        if (MODE == 0) {
            double acc = 0.0;
            if (q < d.w)
                for (int tt = 0; tt < npts; ++tt) acc = fma(sm[s.dz_off + tt], act(d.L)[q * PS + tt], acc);
            else
                for (int tt = 0; tt < npts; ++tt) acc += sm[s.dz_off + tt];
            gdst[(q < d.w ? d.w_off[d.L] + q : d.b_off[d.L])] = acc;
        }
    }
    if (MODE == 1) {
        for (int tt = 0; tt < npts; ++tt) {
            double* row = dst + (p0 + tt) * d.n_params;
            for (int q = threadIdx.x; q <= d.w; q += blockDim.x)
                row[q < d.w ? d.w_off[d.L] + q : d.b_off[d.L]] =
                    q < d.w ? sm[s.dz_off + tt] * act(d.L)[q * PS + tt] : sm[s.dz_off + tt];
        }
    }
    // hidden layers, last to first
    for (int l = d.L - 1; l >= 0; --l) {
        const int in = l == 0 ? d.in_dim : d.w;
        const double* a_in = act(l);
        const int npair = d.w * in;
        if (MODE == 0) {
            for (int q = threadIdx.x; q < npair + d.w; q += blockDim.x) {
                double acc = 0.0;
                if (q < npair) {
                    const int o = q / in, i = q % in;
                    for (int tt = 0; tt < npts; ++tt) acc = fma(D[o * PS + tt], a_in[i * PS + tt], acc);
                    gdst[d.w_off[l] + q] = acc;
                } else {
                    const int o = q - npair;
                    for (int tt = 0; tt < npts; ++tt) acc += D[o * PS + tt];
                    gdst[d.b_off[l] + o] = acc;
                }
            }
        } else {
            for (int tt = 0; tt < npts; ++tt) {
                double* row = dst + (p0 + tt) * d.n_params;
                for (int q = threadIdx.x; q < npair + d.w; q += blockDim.x) {
                    if (q < npair) {
                        const int o = q / in, i = q % in;
                        row[d.w_off[l] + q] = D[o * PS + tt] * a_in[i * PS + tt];
                    } else {
                        row[d.b_off[l] + (q - npair)] = D[(q - npair) * PS + tt];
                    }
                }
            }
        }
        if (l > 0) {
            // delta of layer l-1's output: Dn[i][t] = (sum_o W_l[o][i] D[o][t]) (1 - a^2)
            const double* Wt = sm + s.w_off[l];  // [in][wp]
            for (int i = 0; i < d.w; ++i) {
                double acc = 0.0;
                for (int o = 0; o < d.w; ++o) acc = fma(Wt[i * d.wp + o], D[o * PS + t], acc);
                const double a = a_in[i * PS + t];
                Dn[i * PS + t] = acc * (1.0 - a * a);
            }
            for (int i = d.w; i < d.wp; ++i) Dn[i * PS + t] = 0.0;
            __syncthreads();
            double* tmp = D;
            D = Dn;
            Dn = tmp;
        }
    }
}

// out[c] = sum over rows of part[r][c]: a CTA per 32 columns, 32 strided partial sums per column folded
// in a fixed order (rows = CTAs of the gradient kernel)
__global__ void __launch_bounds__(1024) reduce_rows_kernel(const double* __restrict__ part, int64_t rows, int64_t cols,
                                                           double* __restrict__ out) {
    __shared__ double s[32][33];
    const int64_t c = (int64_t)blockIdx.x * 32 + threadIdx.x;
    double acc = 0.0;
    if (c < cols)
        for (int64_t r = threadIdx.y; r < rows; r += 32) acc += part[r * cols + c];
    s[threadIdx.y][threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.y == 0 && c < cols) {
        double t = 0.0;
        for (int y = 0; y < 32; ++y) t += s[y][threadIdx.x];
        out[c] = t;
    }
}

}  // namespace

int pf_mlp_make_desc(int input_dim, int hidden_layers, int width, PfMlpDesc* d) {
    PF_REQUIRE(input_dim >= 1 && input_dim <= 16, "MLP input_dim must be in [1,16], got %d", input_dim);
    PF_REQUIRE(hidden_layers >= 1 && hidden_layers <= PF_MLP_MAX_LAYERS, "MLP hidden_layers must be in [1,%d]",
               PF_MLP_MAX_LAYERS);
    PF_REQUIRE(width >= 1 && width <= 128, "MLP width must be in [1,128], got %d", width);
    d->in_dim = input_dim;
    d->L = hidden_layers;
    d->w = width;
    d->wp = (width + 3) & ~3;
    int off = 0;
    for (int l = 0; l < hidden_layers; ++l) {
        const int in = l == 0 ? input_dim : width;
        d->w_off[l] = off;
        off += width * in;
        d->b_off[l] = off;
        off += width;
    }
    d->w_off[hidden_layers] = off;
    off += width;
    d->b_off[hidden_layers] = off;
    off += 1;
    d->n_params = off;
    return PF_OK;
}

extern "C" int64_t pf_mlp_num_params(int input_dim, int hidden_layers, int width) {
    PfMlpDesc d;
    if (pf_mlp_make_desc(input_dim, hidden_layers, width, &d) != PF_OK) return -1;
    return d.n_params;
}

static int mlp_common(pf_plan* plan, int input_dim, int hidden_layers, int width, const double* theta, int64_t n,
                      const double* X, PfMlpDesc* d, const double** centroid) {
    int rc = pf_mlp_make_desc(input_dim, hidden_layers, width, d);
    if (rc) return rc;
    PF_REQUIRE(theta != nullptr, "theta is NULL");
    PF_REQUIRE(n >= 0, "n must be >= 0");
    *centroid = nullptr;
    if (X == nullptr) {
        PF_REQUIRE(plan != nullptr, "X is NULL and no plan given");
        rc = pf_plan_activate(plan);
        if (rc) return rc;
        PF_REQUIRE(n == plan->nelem, "centroid mode needs n == nelem");
        PF_REQUIRE(input_dim == plan->dim + 1,
                   "network input_dim %d does not match [load_factor, centroid] = %d inputs (fem/properties.py:116-125)",
                   input_dim, plan->dim + 1);
        *centroid = plan->d_centroid;
    } else if (plan) {
        rc = pf_plan_activate(plan);
        if (rc) return rc;
    }
    return PF_OK;
}

constexpr size_t kSmemLimit = 220 * 1024;

// pf_mlp_tc.cu: DMMA kernels for large point sets
int pf_mlp_tc_grid(const PfMlpDesc& d, bool backward, int64_t n, int sm_count);
int pf_mlp_tc_launch(const PfMlpDesc& d, bool backward, const double* theta, int64_t n, const double* X,
                     const double* centroid, double load_factor, double scale, int positive, const double* g_out,
                     double* out, double* part, int grid, cudaStream_t st);
constexpr int64_t kTcMinPoints = 2048;  // below this the per-CTA weight staging of the DMMA kernel does not pay

static int sm_count_of(pf_plan* plan) {
    if (plan) return plan->sm_count;
    int dev = 0, n = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n;
}
static bool tc_enabled() {
    static const int off = getenv("PF_MLP_NO_TC") ? atoi(getenv("PF_MLP_NO_TC")) : 0;
    return !off;
}
static bool frag_enabled() {
    const char* e = getenv("PF_MLP_NO_FRAG");  // read per call: the tests flip it
    return !(e && atoi(e));
}

static int pick_pts(const PfMlpDesc& d, bool backward, int* pts, size_t* smem) {
    for (int cand : {128, 64, 32}) {
        const size_t bytes = (size_t)smem_plan(d, backward, cand).total * sizeof(double);
        if (bytes <= kSmemLimit) {
            *pts = cand;
            *smem = bytes;
            return PF_OK;
        }
    }
    pf_set_error("network too large for the shared-memory MLP kernels");
    return PF_ERR_ARG;
}

template <typename Kernel>
static int set_smem(Kernel k, size_t bytes) {
    if (bytes > 48 * 1024) PF_CUDA_CHECK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return PF_OK;
}

#define PF_MLP_DISPATCH(PTSV, CALL) \
    do {                            \
        if ((PTSV) == 128) {        \
            CALL(128);              \
        } else if ((PTSV) == 64) {  \
            CALL(64);               \
        } else {                    \
            CALL(32);               \
        }                           \
    } while (0)

extern "C" int pf_mlp_forward(pf_plan* plan, int input_dim, int hidden_layers, int width, const double* theta,
                              int64_t n, const double* X, double load_factor, double scale, int enforce_positive,
                              double* out, void* stream) {
    PfMlpDesc d;
    const double* cen;
    int rc = mlp_common(plan, input_dim, hidden_layers, width, theta, n, X, &d, &cen);
    if (rc) return rc;
    PF_REQUIRE(out != nullptr, "out is NULL");
    if (n == 0) return PF_OK;
    cudaStream_t st = pf_stream_of(stream);
    // large point sets: the register-resident DMMA kernel (pf_mlp_frag.cu); PF_MLP_NO_FRAG=1 keeps the older kernels
    if (n >= kTcMinPoints && frag_enabled() && pf_mlp_frag_supported(d, false))
        return pf_mlp_frag_forward(d, theta, 0, 1, n, X, cen, load_factor, scale, enforce_positive, out, 1, nullptr, 0,
                                   sm_count_of(plan), st);
    // forward alone is bound by the 2 x w tanh evaluations per point, not by the contractions: the FMA
    // kernel (no padded columns) is faster there; PF_MLP_FWD_TC=1 selects the DMMA kernel anyway
    const char* fwd_env = getenv("PF_MLP_FWD_TC");  // read per call: the tests flip it
    const int fwd_tc = fwd_env ? atoi(fwd_env) : 0;
    if (fwd_tc && n >= kTcMinPoints && tc_enabled()) {
        const int grid = pf_mlp_tc_grid(d, false, n, sm_count_of(plan));
        if (grid > 0)
            return pf_mlp_tc_launch(d, false, theta, n, X, cen, load_factor, scale, enforce_positive, nullptr, out,
                                    nullptr, grid, st);
    }
    int pts;
    size_t smem;
    if ((rc = pick_pts(d, false, &pts, &smem))) return rc;
#define PF_FWD(P)                                                                                             \
    do {                                                                                                      \
        if ((rc = set_smem(mlp_forward_kernel<P>, smem))) return rc;                                          \
        mlp_forward_kernel<P><<<(unsigned)((n + P - 1) / P), P, smem, st>>>(d, theta, n, X, cen, load_factor, \
                                                                           scale, enforce_positive, out);     \
    } while (0)
    PF_MLP_DISPATCH(pts, PF_FWD);
#undef PF_FWD
    PF_CUDA_CHECK(cudaGetLastError());
    return PF_OK;
}

// scratch for cross-CTA partial gradients when no plan is available
static thread_local double* g_scratch = nullptr;
static thread_local size_t g_scratch_bytes = 0;

static int scratch(pf_plan* plan, size_t bytes, double** out) {
    if (plan) {
        int rc = pf_plan_reserve_work(plan, bytes);
        *out = plan->d_work;
        return rc;
    }
    if (bytes > g_scratch_bytes) {
        if (g_scratch) PF_CUDA_CHECK(cudaFree(g_scratch));
        g_scratch = nullptr;
        g_scratch_bytes = 0;
        PF_CUDA_CHECK(cudaMalloc((void**)&g_scratch, bytes));
        g_scratch_bytes = bytes;
    }
    *out = g_scratch;
    return PF_OK;
}

extern "C" int pf_mlp_backward(pf_plan* plan, int input_dim, int hidden_layers, int width, const double* theta,
                               int64_t n, const double* X, double load_factor, double scale, int enforce_positive,
                               const double* g_out, double* g_theta, void* stream) {
    PfMlpDesc d;
    const double* cen;
    int rc = mlp_common(plan, input_dim, hidden_layers, width, theta, n, X, &d, &cen);
    if (rc) return rc;
    PF_REQUIRE(g_out != nullptr && g_theta != nullptr, "g_out/g_theta is NULL");
    cudaStream_t st = pf_stream_of(stream);
    if (n == 0) {
        PF_CUDA_CHECK(cudaMemsetAsync(g_theta, 0, d.n_params * sizeof(double), st));
        return PF_OK;
    }
    if (n >= kTcMinPoints && frag_enabled() && pf_mlp_frag_supported(d, true)) {
        // forward into a scratch activation record (values not written), then the backward that consumes it
        const int sms = sm_count_of(plan);
        const int64_t acts_len = pf_mlp_frag_acts_len(d, n);
        const size_t part_len = ((size_t)pf_mlp_frag_chunks(d, n, 1, sms) * d.n_params + 1) & ~size_t(1);
        double* buf = nullptr;
        if ((rc = scratch(plan, (part_len + (size_t)acts_len) * sizeof(double), &buf))) return rc;
        if ((rc = pf_mlp_frag_forward(d, theta, 0, 1, n, X, cen, load_factor, scale, enforce_positive, nullptr, 1,
                                      buf + part_len, 0, sms, st)))
            return rc;
        return pf_mlp_frag_backward(d, theta, 0, 1, n, X, cen, load_factor, g_out, 1, buf + part_len, 0, buf, g_theta, 0,
                                    sms, st);
    }
    if (n >= kTcMinPoints && tc_enabled()) {
        const int grid = pf_mlp_tc_grid(d, true, n, sm_count_of(plan));
        if (grid > 0) {
            double* tpart = nullptr;
            if ((rc = scratch(plan, (size_t)grid * d.n_params * sizeof(double), &tpart))) return rc;
            rc = pf_mlp_tc_launch(d, true, theta, n, X, cen, load_factor, scale, enforce_positive, g_out, nullptr,
                                  tpart, grid, st);
            if (rc) return rc;
            reduce_rows_kernel<<<(d.n_params + 31) / 32, dim3(32, 32), 0, st>>>(tpart, grid, d.n_params, g_theta);
            PF_CUDA_CHECK(cudaGetLastError());
            return PF_OK;
        }
    }
    int pts;
    size_t smem;
    if ((rc = pick_pts(d, true, &pts, &smem))) return rc;
    const unsigned blocks = (unsigned)((n + pts - 1) / pts);
    double* part = g_theta;
    if (blocks > 1 && (rc = scratch(plan, (size_t)blocks * d.n_params * sizeof(double), &part))) return rc;
#define PF_BWD(P)                                                                                               \
    do {                                                                                                        \
        if ((rc = set_smem(mlp_backward_kernel<0, P>, smem))) return rc;                                        \
        mlp_backward_kernel<0, P><<<blocks, P, smem, st>>>(d, theta, n, X, cen, load_factor, scale,             \
                                                           enforce_positive, g_out, part);                      \
    } while (0)
    PF_MLP_DISPATCH(pts, PF_BWD);
#undef PF_BWD
    PF_CUDA_CHECK(cudaGetLastError());
    if (blocks > 1) {
        reduce_rows_kernel<<<(d.n_params + 31) / 32, dim3(32, 32), 0, st>>>(part, blocks, d.n_params, g_theta);
        PF_CUDA_CHECK(cudaGetLastError());
    }
    return PF_OK;
}

extern "C" int64_t pf_mlp_acts_len(int input_dim, int hidden_layers, int width, int64_t n) {
    PfMlpDesc d;
    if (pf_mlp_make_desc(input_dim, hidden_layers, width, &d) != PF_OK || n < 0) return -1;
    return pf_mlp_frag_supported(d, true) ? pf_mlp_frag_acts_len(d, n) : 0;
}

extern "C" int pf_mlp_forward_batched(pf_plan* plan, int input_dim, int hidden_layers, int width, const double* theta,
                                      int64_t theta_stride, int64_t B, int64_t n, const double* X, double load_factor,
                                      double scale, int enforce_positive, double* out, int64_t ldb, double* acts,
                                      void* stream) {
    PfMlpDesc d;
    const double* cen;
    int rc = mlp_common(plan, input_dim, hidden_layers, width, theta, n, X, &d, &cen);
    if (rc) return rc;
    PF_REQUIRE(B >= 0 && ldb >= B && theta_stride >= 0, "pf_mlp_forward_batched: bad batch arguments");
    PF_REQUIRE(out != nullptr || acts != nullptr, "out and acts are both NULL");
    PF_REQUIRE(pf_mlp_frag_supported(d, acts != nullptr),
               "network shape %d-%dx%d not covered by the batched kernels (input_dim <= 3, width <= 24, <= 3 hidden layers)",
               input_dim, hidden_layers, width);
    return pf_mlp_frag_forward(d, theta, theta_stride, B, n, X, cen, load_factor, scale, enforce_positive, out, ldb, acts,
                               pf_mlp_frag_acts_len(d, n), sm_count_of(plan), pf_stream_of(stream));
}

extern "C" int pf_mlp_backward_batched(pf_plan* plan, int input_dim, int hidden_layers, int width, const double* theta,
                                       int64_t theta_stride, int64_t B, int64_t n, const double* X, double load_factor,
                                       const double* g_out, int64_t ldb, const double* acts, double* g_theta,
                                       int64_t gt_stride, void* stream) {
    PfMlpDesc d;
    const double* cen;
    int rc = mlp_common(plan, input_dim, hidden_layers, width, theta, n, X, &d, &cen);
    if (rc) return rc;
    PF_REQUIRE(B >= 0 && ldb >= B && theta_stride >= 0 && gt_stride >= d.n_params, "pf_mlp_backward_batched: bad batch arguments");
    PF_REQUIRE(g_out && g_theta && acts, "pf_mlp_backward_batched: NULL argument (acts comes from pf_mlp_forward_batched)");
    PF_REQUIRE(pf_mlp_frag_supported(d, true),
               "network shape %d-%dx%d not covered by the batched kernels (input_dim <= 3, width <= 24, <= 3 hidden layers)",
               input_dim, hidden_layers, width);
    const int sms = sm_count_of(plan);
    double* part = nullptr;
    if ((rc = scratch(plan, (size_t)B * pf_mlp_frag_chunks(d, n, B, sms) * d.n_params * sizeof(double), &part))) return rc;
    return pf_mlp_frag_backward(d, theta, theta_stride, B, n, X, cen, load_factor, g_out, ldb, acts,
                                pf_mlp_frag_acts_len(d, n), part, g_theta, gt_stride, sms, pf_stream_of(stream));
}

extern "C" int pf_mlp_param_jacobian(pf_plan* plan, int input_dim, int hidden_layers, int width,
                                     const double* theta, int64_t n, const double* X, double load_factor,
                                     double scale, int enforce_positive, double* jac, void* stream) {
    PfMlpDesc d;
    const double* cen;
    int rc = mlp_common(plan, input_dim, hidden_layers, width, theta, n, X, &d, &cen);
    if (rc) return rc;
    PF_REQUIRE(jac != nullptr, "jac is NULL");
    if (n == 0) return PF_OK;
    int pts;
    size_t smem;
    if ((rc = pick_pts(d, true, &pts, &smem))) return rc;
    cudaStream_t st = pf_stream_of(stream);
#define PF_JAC(P)                                                                                              \
    do {                                                                                                       \
        if ((rc = set_smem(mlp_backward_kernel<1, P>, smem))) return rc;                                       \
        mlp_backward_kernel<1, P><<<(unsigned)((n + P - 1) / P), P, smem, st>>>(d, theta, n, X, cen,          \
                                                                                load_factor, scale,            \
                                                                                enforce_positive, nullptr, jac); \
    } while (0)
    PF_MLP_DISPATCH(pts, PF_JAC);
#undef PF_JAC
    PF_CUDA_CHECK(cudaGetLastError());
    return PF_OK;
}

namespace {
__global__ void tanh_probe_kernel(int64_t n, const double* __restrict__ x, double* __restrict__ y) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = pf_tanh(x[i]);
}
}  // namespace

// y[i] = pf_tanh(x[i]) (dev pointers): the hidden-layer activation of every MLP kernel, exposed for its accuracy test.
extern "C" int pf_debug_tanh(int64_t n, const double* x, double* y, void* stream) {
    PF_REQUIRE(n >= 0 && (n == 0 || (x && y)), "pf_debug_tanh: bad argument");
    if (n == 0) return PF_OK;
    tanh_probe_kernel<<<(unsigned)((n + 255) / 256), 256, 0, pf_stream_of(stream)>>>(n, x, y);
    PF_CUDA_CHECK(cudaGetLastError());
    return PF_OK;
}
