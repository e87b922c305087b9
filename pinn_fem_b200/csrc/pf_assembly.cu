// Element internal force / residual / tangent kernels (fp64, sm_100a).
//
// Design (DESIGN.md "K1"): the scatter `f_int[dofs] += fe` of the reference
// (fem/assembly.py:72, fem/nn_assembly.py:226-229) is turned into a GATHER.
// One thread owns one (node, problem) pair and walks the node's incident
// elements in ascending element id -- the order the reference's `+=` sums
// them -- so the result is deterministic, needs no atomics and no colouring,
// and every output is written exactly once.  Batched arrays are [row][B] with
// the problem index innermost: a warp covers 32 consecutive problems of the
// same node, every data access is a coalesced 256-byte row segment, and all
// mesh indices / geometry are warp-uniform broadcast loads amortised over the
// batch.  Each E/A row is read from DRAM once; its second use (from the
// element's other end node) hits L1/L2.
#include "pf_internal.h"

namespace {

constexpr int kMaxBlockThreads = 256;

struct GatherArgs {
    const int32_t* __restrict__ inc_ptr;
    const PfIncidence* __restrict__ inc;
    const double4* __restrict__ inc_geo;  // {cos, sin, 1/l0, l0}
    const double4* __restrict__ inc_xy;   // {x_nbr, y_nbr, x_self, y_self}
    const uint8_t* __restrict__ dof_free;
    const double* __restrict__ u;      // [ndof][B] displacement state
    const double* __restrict__ v;      // [ndof][B] vector for the mat-vec (MODE_MATVEC)
    const double* __restrict__ E;
    const double* __restrict__ A;
    const double* __restrict__ f_ext;
    double* __restrict__ f_out;        // f_int or K v
    double* __restrict__ r_out;        // masked residual
    double* __restrict__ half_sq_part; // [gridDim.x][B] partial sums of 0.5 r^2
    unsigned long long* __restrict__ max_strain_bits;  // [B]
    int64_t nnode;
    int64_t B;
    int64_t mat_stride;   // B when per-problem materials, 0 when shared
    int64_t mat_bmul;     // 1 when per-problem, 0 when shared
    int64_t fext_stride;  // B or 1
    int64_t fext_bmul;    // 1 or 0
    double load_factor;
    int nodes_per_thread;
};

enum { MODE_FORCE = 0, MODE_MATVEC = 1 };

template <int DIM>
struct Vec {
    double x, y;
};

template <int DIM>
__device__ __forceinline__ Vec<DIM> load_vec(const double* __restrict__ p, int64_t node, int64_t B, int64_t b) {
    Vec<DIM> r;
    if (DIM == 2) {
        r.x = __ldg(p + (2 * node) * B + b);
        r.y = __ldg(p + (2 * node + 1) * B + b);
    } else {
        r.x = __ldg(p + node * B + b);
        r.y = 0.0;
    }
    return r;
}

// Contribution of one incident element to the owning node's force (or K v).
template <int DIM, int KIND, int MODE>
__device__ __forceinline__ void incidence(double Ee, double Ae, const double4& geo, const double4& xy,
                                          const Vec<DIM>& us, const Vec<DIM>& uo, const Vec<DIM>& vs,
                                          const Vec<DIM>& vo, double& fx, double& fy, double& eps_abs) {
    const double ea = Ee * Ae;
    if (KIND == PF_ELEM_LINEAR || DIM == 1) {
        // fe = ke @ u_e with ke = (EA/l0) * d (x) d, d = [c, s, -c, -s]
        // (fem/element.py:80-100): row of the owning node = k * (d . (u_self - u_other)) * (c, s)
        const double k = ea * geo.z;
        const Vec<DIM>& a = (MODE == MODE_MATVEC) ? vs : us;
        const Vec<DIM>& o = (MODE == MODE_MATVEC) ? vo : uo;
        if (DIM == 2) {
            const double axial = geo.x * (a.x - o.x) + geo.y * (a.y - o.y);
            const double t = k * axial;
            fx += t * geo.x;
            fy += t * geo.y;
            eps_abs = fmax(eps_abs, fabs(axial * geo.z));
        } else {
            const double du = a.x - o.x;
            fx += k * du;
            eps_abs = fmax(eps_abs, fabs(du * geo.z));
        }
    } else {
        // Green-Lagrange, verbatim fem/element.py:119-131 seen from the owning node:
        // d = (x_o + u_o) - (x_s + u_s), e = (l^2 - l0^2) / (2 l0^2)
        const double dx = (xy.x + uo.x) - (xy.z + us.x);
        const double dy = (xy.y + uo.y) - (xy.w + us.y);
        const double l0 = geo.w;
        const double e = (dx * dx + dy * dy - l0 * l0) * (0.5 * geo.z * geo.z);
        if (MODE == MODE_FORCE) {
            const double n = ea * geo.z * e;  // fe = (EA/l0) e d
            fx += n * dx;
            fy += n * dy;
            eps_abs = fmax(eps_abs, fabs(e));
        } else {
            // ke = EA/l0^3 d0 (x) d0 + EA/l0 e d (x) d ; row of owning node acts on (v_s - v_o)
            const double dx0 = xy.x - xy.z, dy0 = xy.y - xy.w;
            const double wx = vs.x - vo.x, wy = vs.y - vo.y;
            const double a = ea * geo.z * geo.z * geo.z * (dx0 * wx + dy0 * wy);
            const double b = ea * geo.z * e * (dx * wx + dy * wy);
            fx += a * dx0 + b * dx;
            fy += a * dy0 + b * dy;
        }
    }
}

template <int DIM, int KIND, int MODE>
__global__ void __launch_bounds__(kMaxBlockThreads) node_gather_kernel(GatherArgs a) {
    const int64_t b = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
    const bool b_ok = b < a.B;
    const int64_t bb = b_ok ? b : 0;
    const int64_t node0 = ((int64_t)blockIdx.x * blockDim.y + threadIdx.y) * a.nodes_per_thread;
    double sq = 0.0, eps_abs = 0.0;

    for (int t = 0; t < a.nodes_per_thread; ++t) {
        const int64_t n = node0 + t;
        if (n >= a.nnode) break;
        const int k0 = __ldg(a.inc_ptr + n), k1 = __ldg(a.inc_ptr + n + 1);
        Vec<DIM> us = {0.0, 0.0}, vs = {0.0, 0.0};
        if (KIND != PF_ELEM_LINEAR || MODE == MODE_FORCE) us = load_vec<DIM>(a.u, n, a.B, bb);
        if (MODE == MODE_MATVEC) vs = load_vec<DIM>(a.v, n, a.B, bb);
        double fx = 0.0, fy = 0.0;
#pragma unroll 4
        for (int k = k0; k < k1; ++k) {
            const PfIncidence inc = a.inc[k];
            const double4 geo = a.inc_geo[k];
            double4 xy = make_double4(0, 0, 0, 0);
            if (KIND == PF_ELEM_GREEN_LAGRANGE && DIM == 2) xy = a.inc_xy[k];
            const double Ee = __ldg(a.E + (int64_t)inc.elem * a.mat_stride + bb * a.mat_bmul);
            const double Ae = __ldg(a.A + (int64_t)inc.elem * a.mat_stride + bb * a.mat_bmul);
            Vec<DIM> uo = {0.0, 0.0}, vo = {0.0, 0.0};
            if (KIND != PF_ELEM_LINEAR || MODE == MODE_FORCE) uo = load_vec<DIM>(a.u, inc.nbr, a.B, bb);
            if (MODE == MODE_MATVEC) vo = load_vec<DIM>(a.v, inc.nbr, a.B, bb);
            incidence<DIM, KIND, MODE>(Ee, Ae, geo, xy, us, uo, vs, vo, fx, fy, eps_abs);
        }
        if (b_ok) {
            const int64_t d0 = (DIM == 2) ? 2 * n : n;
            if (a.f_out) {
                a.f_out[d0 * a.B + b] = fx;
                if (DIM == 2) a.f_out[(d0 + 1) * a.B + b] = fy;
            }
            if (MODE == MODE_FORCE && (a.r_out || a.half_sq_part)) {
                double rx = 0.0, ry = 0.0;
                if (a.dof_free[d0]) rx = fx - a.load_factor * __ldg(a.f_ext + d0 * a.fext_stride + b * a.fext_bmul);
                if (DIM == 2 && a.dof_free[d0 + 1])
                    ry = fy - a.load_factor * __ldg(a.f_ext + (d0 + 1) * a.fext_stride + b * a.fext_bmul);
                if (a.r_out) {
                    a.r_out[d0 * a.B + b] = rx;
                    if (DIM == 2) a.r_out[(d0 + 1) * a.B + b] = ry;
                }
                sq += rx * rx;  // node order inside a thread is ascending
                sq += ry * ry;
            }
        }
    }

    if (MODE == MODE_FORCE && (a.half_sq_part || a.max_strain_bits)) {
        // deterministic block reduction over threadIdx.y for each problem column
        __shared__ double s_sq[kMaxBlockThreads];
        __shared__ double s_eps[kMaxBlockThreads];
        const int tid = threadIdx.y * blockDim.x + threadIdx.x;
        s_sq[tid] = sq;
        s_eps[tid] = eps_abs;
        __syncthreads();
        if (threadIdx.y == 0 && b_ok) {
            double acc = 0.0, m = 0.0;
            for (int y = 0; y < (int)blockDim.y; ++y) {
                acc += s_sq[y * blockDim.x + threadIdx.x];
                m = fmax(m, s_eps[y * blockDim.x + threadIdx.x]);
            }
            if (a.half_sq_part) a.half_sq_part[(int64_t)blockIdx.x * a.B + b] = acc;
            // non-negative doubles order like their bit patterns; max is exact and order independent
            if (a.max_strain_bits) atomicMax(a.max_strain_bits + b, (unsigned long long)__double_as_longlong(m));
        }
    }
}

// out[b] = scale * sum_rows part[row][b], rows added in ascending order (deterministic).
__global__ void column_sum_kernel(const double* __restrict__ part, int64_t rows, int64_t B, double scale,
                                  double* __restrict__ out) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double acc = 0.0;
    for (int64_t r = 0; r < rows; ++r) acc += part[r * B + b];
    out[b] = scale * acc;
}

// ---------------------------------------------------------------------------
// material VJP (element-centric, pure gather)
// ---------------------------------------------------------------------------
struct VjpArgs {
    const int2* __restrict__ conn;
    const double4* __restrict__ elem_geo;
    const double4* __restrict__ elem_xy;
    const double* __restrict__ u;
    const double* __restrict__ g;
    const double* __restrict__ E;
    const double* __restrict__ A;
    double* __restrict__ gE;
    double* __restrict__ gA;
    int64_t nelem, B, mat_stride, mat_bmul;
};

template <int DIM, int KIND>
__global__ void __launch_bounds__(kMaxBlockThreads) material_vjp_kernel(VjpArgs a) {
    const int64_t b = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
    const int64_t e = (int64_t)blockIdx.x * blockDim.y + threadIdx.y;
    if (b >= a.B || e >= a.nelem) return;
    const int2 c = a.conn[e];
    const double4 geo = a.elem_geo[e];
    const Vec<DIM> ui = load_vec<DIM>(a.u, c.x, a.B, b), uj = load_vec<DIM>(a.u, c.y, a.B, b);
    const Vec<DIM> gi = load_vec<DIM>(a.g, c.x, a.B, b), gj = load_vec<DIM>(a.g, c.y, a.B, b);
    double gh;  // <g_e, f_e / (E A)>
    if (KIND == PF_ELEM_LINEAR || DIM == 1) {
        if (DIM == 2) {
            const double axial = geo.x * (ui.x - uj.x) + geo.y * (ui.y - uj.y);
            gh = geo.z * axial * (geo.x * (gi.x - gj.x) + geo.y * (gi.y - gj.y));
        } else {
            gh = geo.z * (ui.x - uj.x) * (gi.x - gj.x);
        }
    } else {
        const double4 xy = a.elem_xy[e];
        const double dx = (xy.z + uj.x) - (xy.x + ui.x);
        const double dy = (xy.w + uj.y) - (xy.y + ui.y);
        const double l0 = geo.w;
        const double eg = (dx * dx + dy * dy - l0 * l0) * (0.5 * geo.z * geo.z);
        gh = geo.z * eg * (dx * (gi.x - gj.x) + dy * (gi.y - gj.y));
    }
    const double Ee = __ldg(a.E + e * a.mat_stride + b * a.mat_bmul);
    const double Ae = __ldg(a.A + e * a.mat_stride + b * a.mat_bmul);
    a.gE[e * a.B + b] = Ae * gh;
    a.gA[e * a.B + b] = Ee * gh;
}

// ---------------------------------------------------------------------------
// tangent in node-block CSR (node-centric: a thread owns one block row)
// ---------------------------------------------------------------------------
struct TangentArgs {
    const int32_t* __restrict__ inc_ptr;
    const PfIncidence* __restrict__ inc;
    const double4* __restrict__ inc_geo;
    const double4* __restrict__ inc_xy;
    const int32_t* __restrict__ diag_slot;
    const double* __restrict__ u;
    const double* __restrict__ E;
    const double* __restrict__ A;
    double* __restrict__ vals;  // [nnzb][DIM*DIM][B]
    int64_t nnode, B, mat_stride, mat_bmul;
};

template <int DIM, int KIND>
__global__ void __launch_bounds__(kMaxBlockThreads) tangent_bsr_kernel(TangentArgs a) {
    constexpr int DD = DIM * DIM;
    const int64_t b = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
    const int64_t n = (int64_t)blockIdx.x * blockDim.y + threadIdx.y;
    if (b >= a.B || n >= a.nnode) return;
    const int k0 = a.inc_ptr[n], k1 = a.inc_ptr[n + 1];
    Vec<DIM> us = {0.0, 0.0};
    if (KIND != PF_ELEM_LINEAR && DIM == 2) us = load_vec<DIM>(a.u, n, a.B, b);
    double d00 = 0.0, d01 = 0.0, d10 = 0.0, d11 = 0.0;
    for (int k = k0; k < k1; ++k) {
        const PfIncidence inc = a.inc[k];
        const double4 geo = a.inc_geo[k];
        const double ea = __ldg(a.E + (int64_t)inc.elem * a.mat_stride + b * a.mat_bmul) *
                          __ldg(a.A + (int64_t)inc.elem * a.mat_stride + b * a.mat_bmul);
        double k00, k01, k10, k11;
        if (KIND == PF_ELEM_LINEAR || DIM == 1) {
            const double kk = ea * geo.z;  // (E*A)/l0, fem/element.py:72
            k00 = kk * (geo.x * geo.x);
            k01 = kk * (geo.x * geo.y);
            k10 = k01;
            k11 = kk * (geo.y * geo.y);
            if (DIM == 1) k00 = kk;
        } else {
            const double4 xy = a.inc_xy[k];
            const Vec<DIM> uo = load_vec<DIM>(a.u, inc.nbr, a.B, b);
            const double dx0 = xy.x - xy.z, dy0 = xy.y - xy.w;
            const double dx = (xy.x + uo.x) - (xy.z + us.x);
            const double dy = (xy.y + uo.y) - (xy.w + us.y);
            const double l0 = geo.w;
            const double e = (dx * dx + dy * dy - l0 * l0) * (0.5 * geo.z * geo.z);
            const double ca = ea * geo.z * geo.z * geo.z;  // EA/l0^3
            const double cb = ea * geo.z * e;              // EA/l0 * e
            k00 = ca * (dx0 * dx0) + cb * (dx * dx);
            k01 = ca * (dx0 * dy0) + cb * (dx * dy);
            k10 = k01;
            k11 = ca * (dy0 * dy0) + cb * (dy * dy);
        }
        d00 += k00;
        d01 += k01;
        d10 += k10;
        d11 += k11;
        double* v = a.vals + ((int64_t)inc.slot * DD) * a.B + b;
        if (inc.flags & 1) {
            v[0] = -k00;
            if (DIM == 2) {
                v[a.B] = -k01;
                v[2 * a.B] = -k10;
                v[3 * a.B] = -k11;
            }
        } else {  // several elements join the same node pair: accumulate in element order
            v[0] -= k00;
            if (DIM == 2) {
                v[a.B] -= k01;
                v[2 * a.B] -= k10;
                v[3 * a.B] -= k11;
            }
        }
    }
    double* v = a.vals + ((int64_t)a.diag_slot[n] * DD) * a.B + b;
    v[0] = d00;
    if (DIM == 2) {
        v[a.B] = d01;
        v[2 * a.B] = d10;
        v[3 * a.B] = d11;
    }
}

template <int DIM>
__global__ void bsr_to_dense_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colind,
                                    const double* __restrict__ vals, int64_t nnode, int64_t ld,
                                    const int32_t* __restrict__ index_map, double* __restrict__ K) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= nnode) return;
    for (int p = rowptr[n]; p < rowptr[n + 1]; ++p) {
        const int64_t m = colind[p];
        for (int r = 0; r < DIM; ++r) {
            int64_t row = n * DIM + r;
            if (index_map) row = index_map[row];
            if (row < 0) continue;
            for (int c = 0; c < DIM; ++c) {
                int64_t col = m * DIM + c;
                if (index_map) col = index_map[col];
                if (col < 0) continue;
                K[row * ld + col] = vals[(int64_t)p * DIM * DIM + r * DIM + c];
            }
        }
    }
}

// launch geometry: x covers problems, y covers nodes/elements
inline void pick_block(int64_t B, dim3& block) {
    int bx = 1;
    while (bx < B && bx < 128) bx <<= 1;
    block = dim3(bx, kMaxBlockThreads / bx, 1);
}

}  // namespace

static int check_common(pf_plan* plan, int kind, int64_t B, const double* E, const double* A) {
    int rc = pf_plan_activate(plan);
    if (rc) return rc;
    PF_REQUIRE(kind == PF_ELEM_LINEAR || kind == PF_ELEM_GREEN_LAGRANGE, "unknown element kind %d", kind);
    PF_REQUIRE(!(kind == PF_ELEM_GREEN_LAGRANGE && plan->dim != 2), "Green-Lagrange element is 2-D only");
    PF_REQUIRE(B >= 1, "B must be >= 1");
    PF_REQUIRE(E != nullptr && A != nullptr, "E/A is NULL");
    return PF_OK;
}

#define PF_DISPATCH_DIM_KIND(plan, kind, CALL)                         \
    do {                                                               \
        if ((plan)->dim == 1) {                                        \
            CALL(1, PF_ELEM_LINEAR);                                   \
        } else if ((kind) == PF_ELEM_LINEAR) {                         \
            CALL(2, PF_ELEM_LINEAR);                                   \
        } else {                                                       \
            CALL(2, PF_ELEM_GREEN_LAGRANGE);                           \
        }                                                              \
    } while (0)

static int launch_gather(pf_plan* plan, int kind, int mode, int64_t B, const double* u, const double* v,
                         const double* E, const double* A, int mat_batched, double* f_out, const double* f_ext,
                         int fext_batched, double load_factor, double* r, double* half_sq, double* max_strain,
                         cudaStream_t st) {
    GatherArgs a{};
    a.inc_ptr = plan->d_inc_ptr;
    a.inc = plan->d_inc;
    a.inc_geo = plan->d_inc_geo;
    a.inc_xy = plan->d_inc_xy;
    a.dof_free = plan->d_dof_free;
    a.u = u;
    a.v = v;
    a.E = E;
    a.A = A;
    a.f_ext = f_ext;
    a.f_out = f_out;
    a.r_out = r;
    a.nnode = plan->nnode;
    a.B = B;
    a.mat_stride = mat_batched ? B : 1;
    a.mat_bmul = mat_batched ? 1 : 0;
    a.fext_stride = fext_batched ? B : 1;
    a.fext_bmul = fext_batched ? 1 : 0;
    a.load_factor = load_factor;

    dim3 block;
    pick_block(B, block);
    // nodes per thread: keep >= ~8 blocks per SM in flight but bound the partial-sum rows
    int npt = 1;
    {
        const int64_t rows1 = (plan->nnode + block.y - 1) / block.y;
        while (npt < 8 && rows1 / (npt * 2) >= (int64_t)plan->sm_count * 16) npt *= 2;
    }
    a.nodes_per_thread = npt;
    const int64_t nodes_per_block = (int64_t)block.y * npt;
    dim3 grid((unsigned)((plan->nnode + nodes_per_block - 1) / nodes_per_block), (unsigned)((B + block.x - 1) / block.x), 1);
    PF_REQUIRE(grid.y <= 65535, "batch too large for one launch: B=%lld", (long long)B);

    if (half_sq) {
        int rc = pf_plan_reserve_work(plan, (size_t)grid.x * B * sizeof(double));
        if (rc) return rc;
        a.half_sq_part = plan->d_work;
    }
    if (max_strain) {
        PF_CUDA_CHECK(cudaMemsetAsync(max_strain, 0, B * sizeof(double), st));
        a.max_strain_bits = reinterpret_cast<unsigned long long*>(max_strain);
    }
#define PF_GATHER_CALL(D, K)                                                       \
    do {                                                                           \
        if (mode == MODE_FORCE)                                                    \
            node_gather_kernel<D, K, MODE_FORCE><<<grid, block, 0, st>>>(a);       \
        else                                                                       \
            node_gather_kernel<D, K, MODE_MATVEC><<<grid, block, 0, st>>>(a);      \
    } while (0)
    PF_DISPATCH_DIM_KIND(plan, kind, PF_GATHER_CALL);
#undef PF_GATHER_CALL
    PF_CUDA_CHECK(cudaGetLastError());
    if (half_sq) {
        column_sum_kernel<<<(unsigned)((B + 127) / 128), 128, 0, st>>>(plan->d_work, grid.x, B, 0.5, half_sq);
        PF_CUDA_CHECK(cudaGetLastError());
    }
    return PF_OK;
}

extern "C" int pf_residual(pf_plan* plan, int kind, int64_t B, const double* u, const double* E, const double* A,
                           int mat_batched, double* f_int, const double* f_ext, int fext_batched,
                           double load_factor, double* r, double* half_sq, double* max_strain, void* stream) {
    int rc = check_common(plan, kind, B, E, A);
    if (rc) return rc;
    PF_REQUIRE(u != nullptr, "u is NULL");
    PF_REQUIRE(f_int || r || half_sq || max_strain, "pf_residual: no output requested");
    PF_REQUIRE(!(r || half_sq) || f_ext, "pf_residual: f_ext is required for r / half_sq");
    return launch_gather(plan, kind, MODE_FORCE, B, u, nullptr, E, A, mat_batched, f_int, f_ext, fext_batched,
                         load_factor, r, half_sq, max_strain, pf_stream_of(stream));
}

extern "C" int pf_tangent_matvec(pf_plan* plan, int kind, int64_t B, const double* u, const double* E,
                                 const double* A, int mat_batched, const double* v, double* out, void* stream) {
    int rc = check_common(plan, kind, B, E, A);
    if (rc) return rc;
    PF_REQUIRE(v != nullptr && out != nullptr, "v/out is NULL");
    PF_REQUIRE(kind == PF_ELEM_LINEAR || u != nullptr, "u is required for the Green-Lagrange tangent");
    return launch_gather(plan, kind, MODE_MATVEC, B, u, v, E, A, mat_batched, out, nullptr, 0, 0.0, nullptr,
                         nullptr, nullptr, pf_stream_of(stream));
}

extern "C" int pf_material_vjp(pf_plan* plan, int kind, int64_t B, const double* u, const double* E,
                               const double* A, int mat_batched, const double* g, double* gE, double* gA,
                               void* stream) {
    int rc = check_common(plan, kind, B, E, A);
    if (rc) return rc;
    PF_REQUIRE(u && g && gE && gA, "pf_material_vjp: NULL argument");
    if (plan->nelem == 0) return PF_OK;
    VjpArgs a{plan->d_conn, plan->d_elem_geo, plan->d_elem_xy, u, g, E, A, gE, gA, plan->nelem, B,
              mat_batched ? B : 1, mat_batched ? 1 : 0};
    dim3 block;
    pick_block(B, block);
    dim3 grid((unsigned)((plan->nelem + block.y - 1) / block.y), (unsigned)((B + block.x - 1) / block.x), 1);
    cudaStream_t st = pf_stream_of(stream);
#define PF_VJP_CALL(D, K) material_vjp_kernel<D, K><<<grid, block, 0, st>>>(a)
    PF_DISPATCH_DIM_KIND(plan, kind, PF_VJP_CALL);
#undef PF_VJP_CALL
    PF_CUDA_CHECK(cudaGetLastError());
    return PF_OK;
}

extern "C" int pf_tangent_bsr(pf_plan* plan, int kind, int64_t B, const double* u, const double* E,
                              const double* A, int mat_batched, double* vals, void* stream) {
    int rc = check_common(plan, kind, B, E, A);
    if (rc) return rc;
    PF_REQUIRE(vals != nullptr, "vals is NULL");
    PF_REQUIRE(kind == PF_ELEM_LINEAR || u != nullptr, "u is required for the Green-Lagrange tangent");
    TangentArgs a{plan->d_inc_ptr, plan->d_inc, plan->d_inc_geo, plan->d_inc_xy, plan->d_diag_slot, u, E, A, vals,
                  plan->nnode, B, mat_batched ? B : 1, mat_batched ? 1 : 0};
    dim3 block;
    pick_block(B, block);
    dim3 grid((unsigned)((plan->nnode + block.y - 1) / block.y), (unsigned)((B + block.x - 1) / block.x), 1);
    cudaStream_t st = pf_stream_of(stream);
#define PF_TAN_CALL(D, K) tangent_bsr_kernel<D, K><<<grid, block, 0, st>>>(a)
    PF_DISPATCH_DIM_KIND(plan, kind, PF_TAN_CALL);
#undef PF_TAN_CALL
    PF_CUDA_CHECK(cudaGetLastError());
    return PF_OK;
}

static int bsr_dense(pf_plan* plan, const double* vals, double* K, bool free_only, cudaStream_t st) {
    int rc = pf_plan_activate(plan);
    if (rc) return rc;
    PF_REQUIRE(vals && K, "bsr_to_dense: NULL argument");
    const int64_t n = free_only ? plan->nfree : plan->ndof;
    PF_CUDA_CHECK(cudaMemsetAsync(K, 0, (size_t)n * n * sizeof(double), st));
    const int threads = 128;
    const unsigned blocks = (unsigned)((plan->nnode + threads - 1) / threads);
    const int32_t* map = free_only ? plan->d_free_index : nullptr;
    if (plan->dim == 1)
        bsr_to_dense_kernel<1><<<blocks, threads, 0, st>>>(plan->d_bsr_rowptr, plan->d_bsr_colind, vals, plan->nnode, n, map, K);
    else
        bsr_to_dense_kernel<2><<<blocks, threads, 0, st>>>(plan->d_bsr_rowptr, plan->d_bsr_colind, vals, plan->nnode, n, map, K);
    PF_CUDA_CHECK(cudaGetLastError());
    return PF_OK;
}

extern "C" int pf_bsr_to_dense(pf_plan* plan, const double* vals, double* K_dense, void* stream) {
    return bsr_dense(plan, vals, K_dense, false, pf_stream_of(stream));
}

extern "C" int pf_bsr_to_free_dense(pf_plan* plan, const double* vals, double* K_ff, void* stream) {
    return bsr_dense(plan, vals, K_ff, true, pf_stream_of(stream));
}
