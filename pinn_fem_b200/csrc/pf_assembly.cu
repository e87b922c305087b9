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
#include <algorithm>
#include <cstdlib>

#include "pf_element.cuh"
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
    int64_t B;            // problems covered by this launch
    int64_t ldb;          // row stride of the batched arrays
    int64_t mat_stride;   // B when per-problem materials, 0 when shared
    int64_t mat_bmul;     // 1 when per-problem, 0 when shared
    int64_t fext_stride;  // B or 1
    int64_t fext_bmul;    // 1 or 0
    double load_factor;
    int nodes_per_thread;
};

enum { MODE_FORCE = 0, MODE_MATVEC = 1 };

// scalar (one problem) node vector, used by the element-centric kernels
struct Vec2 {
    double x, y;
};

template <int DIM>
__device__ __forceinline__ Vec2 load_vec2(const double* __restrict__ p, int64_t node, int64_t B, int64_t b) {
    Vec2 r;
    if (DIM == 2) {
        r.x = __ldg(p + (2 * node) * B + b);
        r.y = __ldg(p + (2 * node + 1) * B + b);
    } else {
        r.x = __ldg(p + node * B + b);
        r.y = 0.0;
    }
    return r;
}

// VEC consecutive problems of one row, loaded with a single 8*VEC-byte access.
template <int VEC>
struct Pack {
    double v[VEC];
};

template <int VEC>
__device__ __forceinline__ Pack<VEC> load_pack(const double* __restrict__ p) {
    Pack<VEC> r;
    if (VEC == 2) {
        const double2 t = __ldg(reinterpret_cast<const double2*>(p));
        r.v[0] = t.x;
        r.v[VEC - 1] = t.y;
    } else {
        r.v[0] = __ldg(p);
    }
    return r;
}

template <int VEC>
__device__ __forceinline__ void store_pack(double* __restrict__ p, const Pack<VEC>& r) {
    if (VEC == 2)
        *reinterpret_cast<double2*>(p) = make_double2(r.v[0], r.v[VEC - 1]);
    else
        p[0] = r.v[0];
}

template <int DIM, int VEC>
struct Vec {
    Pack<VEC> x, y;
};

template <int DIM, int VEC>
__device__ __forceinline__ Vec<DIM, VEC> load_vec(const double* __restrict__ p, int64_t node, int64_t B, int64_t b) {
    Vec<DIM, VEC> r;
    if (DIM == 2) {
        r.x = load_pack<VEC>(p + (2 * node) * B + b);
        r.y = load_pack<VEC>(p + (2 * node + 1) * B + b);
    } else {
        r.x = load_pack<VEC>(p + node * B + b);
#pragma unroll
        for (int i = 0; i < VEC; ++i) r.y.v[i] = 0.0;
    }
    return r;
}

// Contribution of one incident element to the owning node's force (or K v), one problem.
template <int DIM, int KIND, int MODE>
__device__ __forceinline__ void incidence(double Ee, double Ae, const double4& geo, const double4& xy, double usx,
                                          double usy, double uox, double uoy, double vsx, double vsy, double vox,
                                          double voy, double& fx, double& fy, double& eps_abs) {
    const double ea = Ee * Ae;
    if (KIND == PF_ELEM_LINEAR || DIM == 1) {
        const double ax = (MODE == MODE_MATVEC) ? vsx : usx, ay = (MODE == MODE_MATVEC) ? vsy : usy;
        const double ox = (MODE == MODE_MATVEC) ? vox : uox, oy = (MODE == MODE_MATVEC) ? voy : uoy;
        const double eps = pf_linear_incidence<DIM>(Ee, Ae, geo, ax, ay, ox, oy, fx, fy);
        if (MODE == MODE_FORCE) eps_abs = fmax(eps_abs, eps);
    } else {
        // Green-Lagrange, verbatim fem/element.py:119-131 seen from the owning node:
        // d = (x_o + u_o) - (x_s + u_s), e = (l^2 - l0^2) / (2 l0^2)
        const double dx = (xy.x + uox) - (xy.z + usx);
        const double dy = (xy.y + uoy) - (xy.w + usy);
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
            const double wx = vsx - vox, wy = vsy - voy;
            const double a = ea * geo.z * geo.z * geo.z * (dx0 * wx + dy0 * wy);
            const double b = ea * geo.z * e * (dx * wx + dy * wy);
            fx += a * dx0 + b * dx;
            fy += a * dy0 + b * dy;
        }
    }
}

// Incidence records of the CTA's node tile are staged in shared memory first
// (one coalesced cooperative copy), so the data loads below do not wait on a
// dependent index load from L2/DRAM; the loads of up to UD incident elements
// (4 rows each) are issued back to back before the first use, and every lane
// moves VEC problems per access (LDG.128 for VEC = 2): the SM's budget of
// outstanding load requests, not DRAM, is what limits bytes in flight here.
constexpr int kIncCap = 192;  // staged incidences per CTA (more are read from global)

template <int DIM, int KIND, int MODE, int UD, int VEC>
__global__ void __launch_bounds__(kMaxBlockThreads) node_gather_kernel(GatherArgs a) {
    __shared__ PfIncidence s_inc[kIncCap];
    __shared__ double4 s_geo[kIncCap];
    __shared__ double4 s_xy[(KIND == PF_ELEM_GREEN_LAGRANGE && DIM == 2) ? kIncCap : 1];
    __shared__ int s_ptr[kMaxBlockThreads + 1];
    constexpr bool kNeedU = (KIND != PF_ELEM_LINEAR && DIM == 2) || MODE == MODE_FORCE;
    constexpr bool kNeedV = MODE == MODE_MATVEC;
    constexpr bool kNeedXY = KIND == PF_ELEM_GREEN_LAGRANGE && DIM == 2;

    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    const int nthreads = blockDim.x * blockDim.y;
    const int tile_nodes = blockDim.y * a.nodes_per_thread;
    const int64_t tile0 = (int64_t)blockIdx.x * tile_nodes;
    const int tile_n = (int)(a.nnode - tile0 < tile_nodes ? a.nnode - tile0 : tile_nodes);
    for (int i = tid; i <= tile_n; i += nthreads) s_ptr[i] = __ldg(a.inc_ptr + tile0 + i);
    __syncthreads();
    const int kbase = s_ptr[0];
    const int kstaged = min(s_ptr[tile_n] - kbase, kIncCap);
    for (int i = tid; i < kstaged; i += nthreads) {
        s_inc[i] = a.inc[kbase + i];
        s_geo[i] = a.inc_geo[kbase + i];
        if (kNeedXY) s_xy[i] = a.inc_xy[kbase + i];
    }
    __syncthreads();

    const int64_t b = ((int64_t)blockIdx.y * blockDim.x + threadIdx.x) * VEC;
    const bool b_ok = b < a.B;  // B is a multiple of VEC (checked by the launcher)
    const int64_t bb = b_ok ? b : 0;
    const double* __restrict__ Eb = a.E + bb * a.mat_bmul;
    const double* __restrict__ Ab = a.A + bb * a.mat_bmul;
    double sq[VEC], eps_abs[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) sq[i] = eps_abs[i] = 0.0;

    for (int t = 0; t < a.nodes_per_thread; ++t) {
        const int ln = threadIdx.y * a.nodes_per_thread + t;
        if (ln >= tile_n) break;
        const int64_t n = tile0 + ln;
        const int k0 = s_ptr[ln] - kbase, k1 = s_ptr[ln + 1] - kbase;
        Vec<DIM, VEC> us, vs;
        if (kNeedU) us = load_vec<DIM, VEC>(a.u, n, a.ldb, bb);
        if (kNeedV) vs = load_vec<DIM, VEC>(a.v, n, a.ldb, bb);
        double fx[VEC], fy[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) fx[i] = fy[i] = 0.0;
        for (int kb = k0; kb < k1; kb += UD) {
            Pack<VEC> Ee[UD], Ae[UD];
            Vec<DIM, VEC> uo[UD], vo[UD];
#pragma unroll
            for (int j = 0; j < UD; ++j) {
                const int k = kb + j;
                if (k < k1) {
                    const PfIncidence inc = k < kIncCap ? s_inc[k] : a.inc[kbase + k];
                    const int64_t eo = (int64_t)inc.elem * a.mat_stride;
                    if (a.mat_bmul) {
                        Ee[j] = load_pack<VEC>(Eb + eo);
                        Ae[j] = load_pack<VEC>(Ab + eo);
                    } else {  // materials shared by all problems: one scalar per element
                        const double e1 = __ldg(a.E + eo), a1 = __ldg(a.A + eo);
#pragma unroll
                        for (int i = 0; i < VEC; ++i) {
                            Ee[j].v[i] = e1;
                            Ae[j].v[i] = a1;
                        }
                    }
                    if (kNeedU) uo[j] = load_vec<DIM, VEC>(a.u, inc.nbr, a.ldb, bb);
                    if (kNeedV) vo[j] = load_vec<DIM, VEC>(a.v, inc.nbr, a.ldb, bb);
                }
            }
#pragma unroll
            for (int j = 0; j < UD; ++j) {
                const int k = kb + j;
                if (k < k1) {
                    const double4 geo = k < kIncCap ? s_geo[k] : a.inc_geo[kbase + k];
                    double4 xy = make_double4(0, 0, 0, 0);
                    if (kNeedXY) xy = k < kIncCap ? s_xy[k] : a.inc_xy[kbase + k];
#pragma unroll
                    for (int i = 0; i < VEC; ++i)
                        incidence<DIM, KIND, MODE>(Ee[j].v[i], Ae[j].v[i], geo, xy, kNeedU ? us.x.v[i] : 0.0,
                                                   kNeedU ? us.y.v[i] : 0.0, kNeedU ? uo[j].x.v[i] : 0.0,
                                                   kNeedU ? uo[j].y.v[i] : 0.0, kNeedV ? vs.x.v[i] : 0.0,
                                                   kNeedV ? vs.y.v[i] : 0.0, kNeedV ? vo[j].x.v[i] : 0.0,
                                                   kNeedV ? vo[j].y.v[i] : 0.0, fx[i], fy[i], eps_abs[i]);
                }
            }
        }
        if (b_ok) {
            const int64_t d0 = (DIM == 2) ? 2 * n : n;
            if (a.f_out) {
                Pack<VEC> px, py;
#pragma unroll
                for (int i = 0; i < VEC; ++i) {
                    px.v[i] = fx[i];
                    py.v[i] = fy[i];
                }
                store_pack<VEC>(a.f_out + d0 * a.ldb + b, px);
                if (DIM == 2) store_pack<VEC>(a.f_out + (d0 + 1) * a.ldb + b, py);
            }
            if (MODE == MODE_FORCE && (a.r_out || a.half_sq_part)) {
                Pack<VEC> rx, ry;
                const bool free_x = a.dof_free[d0], free_y = DIM == 2 ? a.dof_free[d0 + 1] : false;
#pragma unroll
                for (int i = 0; i < VEC; ++i) {
                    rx.v[i] = ry.v[i] = 0.0;
                    // product rounded first, like the reference's f_int[free] - lambda * f_ext[free]
                    if (free_x)
                        rx.v[i] = __dsub_rn(fx[i], __dmul_rn(a.load_factor, __ldg(a.f_ext + d0 * a.fext_stride + (b + i) * a.fext_bmul)));
                    if (free_y)
                        ry.v[i] = __dsub_rn(fy[i], __dmul_rn(a.load_factor, __ldg(a.f_ext + (d0 + 1) * a.fext_stride + (b + i) * a.fext_bmul)));
                    sq[i] += rx.v[i] * rx.v[i];  // node order inside a thread is ascending
                    sq[i] += ry.v[i] * ry.v[i];
                }
                if (a.r_out) {
                    store_pack<VEC>(a.r_out + d0 * a.ldb + b, rx);
                    if (DIM == 2) store_pack<VEC>(a.r_out + (d0 + 1) * a.ldb + b, ry);
                }
            }
        }
    }

    if (MODE == MODE_FORCE && (a.half_sq_part || a.max_strain_bits)) {
        // deterministic block reduction over threadIdx.y for each problem column
        __shared__ double s_sq[kMaxBlockThreads * VEC];
        __shared__ double s_eps[kMaxBlockThreads * VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            s_sq[tid * VEC + i] = sq[i];
            s_eps[tid * VEC + i] = eps_abs[i];
        }
        __syncthreads();
        if (threadIdx.y == 0 && b_ok) {
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                double acc = 0.0, m = 0.0;
                for (int y = 0; y < (int)blockDim.y; ++y) {
                    acc += s_sq[(y * blockDim.x + threadIdx.x) * VEC + i];
                    m = fmax(m, s_eps[(y * blockDim.x + threadIdx.x) * VEC + i]);
                }
                if (a.half_sq_part) a.half_sq_part[(int64_t)blockIdx.x * a.ldb + b + i] = acc;
                // non-negative doubles order like their bit patterns; max is exact and order independent
                if (a.max_strain_bits)
                    atomicMax(a.max_strain_bits + b + i, (unsigned long long)__double_as_longlong(m));
            }
        }
    }
}

// ---------------------------------------------------------------------------
// Single problem (B = 1), linear element: the reference's own use case on a large mesh (one NR / CG /
// GD state).  The batched kernels read 48 bytes of plan data per incidence (indices + precomputed
// geometry), which at B = 1 is three times the algorithmic traffic; here a thread owns a node, reads
// the compact {elem, nbr} table (8 bytes per incidence) and recomputes {cos, sin, 1/l0} from the node
// coordinates with the same correctly-rounded operations the plan used on the host, so the result
// is bit-identical to the batched kernels.  Orientation does not matter: negating (cos, sin) leaves
// every product of pf_linear_incidence unchanged.
// ---------------------------------------------------------------------------
struct B1Args {
    const int32_t* __restrict__ inc_ptr;
    const int2* __restrict__ inc2;
    const double* __restrict__ nodes;
    const uint8_t* __restrict__ dof_free;
    const double* __restrict__ x;  // u (force) or v (mat-vec)
    const double* __restrict__ E;
    const double* __restrict__ A;
    const double* __restrict__ f_ext;
    double* __restrict__ f_out;
    double* __restrict__ r_out;
    double* __restrict__ half_sq_part;  // [gridDim.x]
    unsigned long long* __restrict__ max_strain_bits;
    int64_t nnode;
    double load_factor;
};

template <int DIM, int MODE>
__global__ void __launch_bounds__(256) node_gather_b1_kernel(B1Args a) {
    __shared__ double s_red[8];
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double sq = 0.0, eps_abs = 0.0;
    if (n < a.nnode) {
        const double xs = a.x[n * DIM], ys = DIM == 2 ? a.x[n * DIM + 1] : 0.0;
        const double px = a.nodes[n * DIM], py = DIM == 2 ? a.nodes[n * DIM + 1] : 0.0;
        double fx = 0.0, fy = 0.0;
        const int k1 = a.inc_ptr[n + 1];
#pragma unroll 3
        for (int k = a.inc_ptr[n]; k < k1; ++k) {
            const int2 ic = __ldg(a.inc2 + k);
            double4 geo;
            double xo, yo = 0.0;
            if (DIM == 2) {
                const double2 pn = __ldg(reinterpret_cast<const double2*>(a.nodes) + ic.y);
                const double2 un = __ldg(reinterpret_cast<const double2*>(a.x) + ic.y);
                const double dx = __dsub_rn(pn.x, px), dy = __dsub_rn(pn.y, py);
                const double l0 = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
                geo = make_double4(__ddiv_rn(dx, l0), __ddiv_rn(dy, l0), __ddiv_rn(1.0, l0), l0);
                xo = un.x;
                yo = un.y;
            } else {
                const double l0 = fabs(__dsub_rn(__ldg(a.nodes + ic.y), px));
                geo = make_double4(1.0, 0.0, __ddiv_rn(1.0, l0), l0);
                xo = __ldg(a.x + ic.y);
            }
            const double eps = pf_linear_incidence<DIM>(__ldg(a.E + ic.x), __ldg(a.A + ic.x), geo, xs, ys, xo, yo, fx, fy);
            if (MODE == MODE_FORCE) eps_abs = fmax(eps_abs, eps);
        }
        const int64_t d0 = n * DIM;
        if (a.f_out) {
            a.f_out[d0] = fx;
            if (DIM == 2) a.f_out[d0 + 1] = fy;
        }
        if (MODE == MODE_FORCE && (a.r_out || a.half_sq_part)) {
            const double rx = a.dof_free[d0] ? __dsub_rn(fx, __dmul_rn(a.load_factor, a.f_ext[d0])) : 0.0;
            const double ry = (DIM == 2 && a.dof_free[d0 + 1]) ? __dsub_rn(fy, __dmul_rn(a.load_factor, a.f_ext[d0 + 1])) : 0.0;
            if (a.r_out) {
                a.r_out[d0] = rx;
                if (DIM == 2) a.r_out[d0 + 1] = ry;
            }
            sq = rx * rx;
            sq += ry * ry;
        }
    }
    if (MODE == MODE_FORCE && (a.half_sq_part || a.max_strain_bits)) {  // fixed-order block reduction
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            sq += __shfl_xor_sync(0xffffffffu, sq, o);
            eps_abs = fmax(eps_abs, __shfl_xor_sync(0xffffffffu, eps_abs, o));
        }
        if (a.half_sq_part) {
            if (lane == 0) s_red[warp] = sq;
            __syncthreads();
            if (threadIdx.x == 0) {
                double t = 0.0;
                for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_red[w];
                a.half_sq_part[blockIdx.x] = t;
            }
        }
        if (a.max_strain_bits && lane == 0)
            atomicMax(a.max_strain_bits, (unsigned long long)__double_as_longlong(eps_abs));
    }
}

// out[b] = scale * sum_rows part[row][b], rows added in ascending order (deterministic).
__global__ void column_sum_kernel(const double* __restrict__ part, int64_t rows, int64_t ld, int64_t B,
                                  double scale, double* __restrict__ out) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double acc = 0.0;
    for (int64_t r = 0; r < rows; ++r) acc += part[r * ld + b];
    out[b] = scale * acc;
}

// Same sum for narrow batches: one CTA per column, 256 strided partial sums folded by a fixed tree
// (a single thread walking ~10^3 rows serially costs more than the residual kernel itself at B = 1).
__global__ void __launch_bounds__(256) column_sum_wide_kernel(const double* __restrict__ part, int64_t rows, int64_t ld,
                                                              double scale, double* __restrict__ out) {
    __shared__ double red[256];
    const int64_t b = blockIdx.x;
    double acc = 0.0;
    for (int64_t r = threadIdx.x; r < rows; r += 256) acc += part[r * ld + b];
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[b] = scale * red[0];
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
    const Vec2 ui = load_vec2<DIM>(a.u, c.x, a.B, b), uj = load_vec2<DIM>(a.u, c.y, a.B, b);
    const Vec2 gi = load_vec2<DIM>(a.g, c.x, a.B, b), gj = load_vec2<DIM>(a.g, c.y, a.B, b);
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
    Vec2 us = {0.0, 0.0};
    if (KIND != PF_ELEM_LINEAR && DIM == 2) us = load_vec2<DIM>(a.u, n, a.B, b);
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
            const Vec2 uo = load_vec2<DIM>(a.u, inc.nbr, a.B, b);
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

// signed strain per element: measure 0 = small-strain axial (fem/element.py:69), 1 = Green-Lagrange
// (element.py:126), 2 = engineering (L - L0)/L0 from the deformed length (api_fem_solver.py:100-108)
template <int DIM>
__global__ void element_strain_kernel(const int2* __restrict__ conn, const double4* __restrict__ elem_geo,
                                      const double4* __restrict__ elem_xy, const double* __restrict__ u, int measure,
                                      int64_t nelem, int64_t B, double* __restrict__ out) {
    const int64_t b = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
    const int64_t e = (int64_t)blockIdx.x * blockDim.y + threadIdx.y;
    if (b >= B || e >= nelem) return;
    const int2 c = conn[e];
    const double4 geo = elem_geo[e];
    const Vec2 ui = load_vec2<DIM>(u, c.x, B, b), uj = load_vec2<DIM>(u, c.y, B, b);
    double eps;
    if (measure == 0 || DIM == 1) {
        eps = DIM == 2 ? (geo.x * (uj.x - ui.x) + geo.y * (uj.y - ui.y)) * geo.z : (uj.x - ui.x) * geo.z;
    } else {
        const double4 xy = elem_xy[e];
        const double dx = (xy.z + uj.x) - (xy.x + ui.x), dy = (xy.w + uj.y) - (xy.y + ui.y);
        const double l0 = geo.w;
        if (measure == 1)
            eps = (dx * dx + dy * dy - l0 * l0) / (2.0 * l0 * l0);
        else
            eps = (sqrt(dx * dx + dy * dy) - l0) / l0;
    }
    out[e * B + b] = eps;
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
    PF_REQUIRE(plan->nelem == 0 || (E != nullptr && A != nullptr), "E/A is NULL");
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

// Generic node gather on the column range [b0, b0 + nb) of arrays whose row stride is ldb.
static int launch_gather_generic(pf_plan* plan, int kind, int mode, int64_t ldb, int64_t b0, int64_t nb,
                                 const double* u, const double* v, const double* E, const double* A,
                                 int mat_batched, double* f_out, const double* f_ext, int fext_batched,
                                 double load_factor, double* r, double* half_sq, double* max_strain,
                                 cudaStream_t st) {
    auto off = [&](const double* q, bool batched) { return (q && batched) ? q + b0 : q; };
    auto offw = [&](double* q) { return q ? q + b0 : q; };
    GatherArgs a{};
    a.inc_ptr = plan->d_inc_ptr;
    a.inc = plan->d_inc;
    a.inc_geo = plan->d_inc_geo;
    a.inc_xy = plan->d_inc_xy;
    a.dof_free = plan->d_dof_free;
    a.u = off(u, true);
    a.v = off(v, true);
    a.E = off(E, mat_batched);
    a.A = off(A, mat_batched);
    a.f_ext = off(f_ext, fext_batched);
    a.f_out = offw(f_out);
    a.r_out = offw(r);
    a.nnode = plan->nnode;
    a.B = nb;
    a.ldb = ldb;
    a.mat_stride = mat_batched ? ldb : 1;
    a.mat_bmul = mat_batched ? 1 : 0;
    a.fext_stride = fext_batched ? ldb : 1;
    a.fext_bmul = fext_batched ? 1 : 0;
    a.load_factor = load_factor;

    // tuning knobs (environment, read once): lanes per CTA row, nodes per thread, problems per lane
    static const int env_bx = getenv("PF_GATHER_BX") ? atoi(getenv("PF_GATHER_BX")) : 0;
    static const int env_npt = getenv("PF_GATHER_NPT") ? atoi(getenv("PF_GATHER_NPT")) : 0;
    static const int env_vec = getenv("PF_GATHER_VEC") ? atoi(getenv("PF_GATHER_VEC")) : 0;
    static const int env_ud = getenv("PF_GATHER_UD") ? atoi(getenv("PF_GATHER_UD")) : 0;
    static const int env_threads = getenv("PF_GATHER_THREADS") ? atoi(getenv("PF_GATHER_THREADS")) : 0;
    // two problems per lane (128-bit accesses) whenever the batch is wide, even and 16-byte aligned
    auto aligned16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    int vec = (nb % 2 == 0 && ldb % 2 == 0 && nb >= 64 && aligned16(a.u) && aligned16(a.v) && aligned16(a.f_out) &&
               aligned16(a.r_out) && (!mat_batched || (aligned16(a.E) && aligned16(a.A))))
                  ? 2
                  : 1;
    if (env_vec == 1) vec = 1;
    const int64_t lanes = nb / vec;
    dim3 block;
    pick_block(lanes, block);
    if (vec == 2) block.y = 128 / block.x > 0 ? 128 / block.x : 1;  // 154 regs/thread: 128-thread CTAs keep 3 per SM
    if (env_bx > 0 && lanes >= env_bx) block = dim3(env_bx, kMaxBlockThreads / env_bx, 1);
    if (env_threads > 0 && env_threads <= kMaxBlockThreads && env_threads >= (int)block.x)
        block.y = env_threads / block.x;
    // nodes per thread: amortise the per-CTA index staging but keep the grid large
    int npt = 1;
    if (env_npt > 0 && (int)block.y * env_npt <= kMaxBlockThreads) {
        npt = env_npt;
    } else {
        const int64_t rows1 = (plan->nnode + block.y - 1) / block.y;
        while (npt < 8 && (int)block.y * npt * 2 <= kMaxBlockThreads &&
               rows1 / (npt * 2) >= (int64_t)plan->sm_count * 16)
            npt *= 2;
    }
    a.nodes_per_thread = npt;
    const int ud = env_ud > 0 ? env_ud : 3;  // 3 incidences per load batch: best register/occupancy balance
    const int64_t nodes_per_block = (int64_t)block.y * npt;
    dim3 grid((unsigned)((plan->nnode + nodes_per_block - 1) / nodes_per_block),
              (unsigned)((lanes + block.x - 1) / block.x), 1);
    PF_REQUIRE(grid.y <= 65535, "batch too large for one launch: B=%lld", (long long)nb);

    if (half_sq) {
        int rc = pf_plan_reserve_work(plan, (size_t)grid.x * ldb * sizeof(double));
        if (rc) return rc;
        a.half_sq_part = plan->d_work + b0;
    }
    if (max_strain) a.max_strain_bits = reinterpret_cast<unsigned long long*>(max_strain + b0);
#define PF_GATHER_CALL2(D, K, M)                                                        \
    do {                                                                                \
        if (vec == 2) {                                                                 \
            if (ud <= 3)                                                                \
                node_gather_kernel<D, K, M, 3, 2><<<grid, block, 0, st>>>(a);           \
            else                                                                        \
                node_gather_kernel<D, K, M, 6, 2><<<grid, block, 0, st>>>(a);           \
        } else {                                                                        \
            if (ud <= 3)                                                                \
                node_gather_kernel<D, K, M, 3, 1><<<grid, block, 0, st>>>(a);           \
            else                                                                        \
                node_gather_kernel<D, K, M, 6, 1><<<grid, block, 0, st>>>(a);           \
        }                                                                               \
    } while (0)
#define PF_GATHER_CALL(D, K)                           \
    do {                                               \
        if (mode == MODE_FORCE)                        \
            PF_GATHER_CALL2(D, K, MODE_FORCE);         \
        else                                           \
            PF_GATHER_CALL2(D, K, MODE_MATVEC);        \
    } while (0)
    PF_DISPATCH_DIM_KIND(plan, kind, PF_GATHER_CALL);
#undef PF_GATHER_CALL
#undef PF_GATHER_CALL2
    PF_CUDA_CHECK(cudaGetLastError());
    if (half_sq) {
        if (nb <= 64)
            column_sum_wide_kernel<<<(unsigned)nb, 256, 0, st>>>(plan->d_work + b0, grid.x, ldb, 0.5, half_sq + b0);
        else
            column_sum_kernel<<<(unsigned)((nb + 127) / 128), 128, 0, st>>>(plan->d_work + b0, grid.x, ldb, nb, 0.5,
                                                                            half_sq + b0);
        PF_CUDA_CHECK(cudaGetLastError());
    }
    return PF_OK;
}

// Wide batches of the linear element go through the patch-staged kernel in
// full 32-problem chunks; everything else (ragged tail, tiny batches,
// Green-Lagrange) through the generic gather.  Both produce identical bits.
static int launch_gather(pf_plan* plan, int kind, int mode, int64_t B, const double* u, const double* v,
                         const double* E, const double* A, int mat_batched, double* f_out, const double* f_ext,
                         int fext_batched, double load_factor, double* r, double* half_sq, double* max_strain,
                         cudaStream_t st) {
    if (max_strain) PF_CUDA_CHECK(cudaMemsetAsync(max_strain, 0, B * sizeof(double), st));
    if (half_sq) {
        // both kernels park partial sums in the plan workspace: size it once for the larger user
        const size_t rows = std::max<size_t>(plan->patches.size(), (size_t)plan->nnode);
        int rc = pf_plan_reserve_work(plan, rows * B * sizeof(double));
        if (rc) return rc;
    }
    static const int no_b1 = getenv("PF_NO_B1") ? atoi(getenv("PF_NO_B1")) : 0;
    const bool x_al16 = (reinterpret_cast<uintptr_t>(mode == MODE_FORCE ? u : v) & 15) == 0;  // double2 loads of x
    if (B == 1 && (kind == PF_ELEM_LINEAR || plan->dim == 1) && plan->nnode > 0 && !no_b1 && x_al16) {
        // all B = 1 arrays are plain vectors whatever mat_batched / fext_batched say
        const unsigned blocks = (unsigned)((plan->nnode + 255) / 256);
        B1Args a{plan->d_inc_ptr, plan->d_inc2, plan->d_nodes, plan->d_dof_free, mode == MODE_FORCE ? u : v, E, A, f_ext,
                 f_out, r, nullptr, max_strain ? reinterpret_cast<unsigned long long*>(max_strain) : nullptr,
                 plan->nnode, load_factor};
        if (half_sq) a.half_sq_part = plan->d_work;  // reserved above: >= nnode doubles
        if (plan->dim == 1) {
            if (mode == MODE_FORCE) node_gather_b1_kernel<1, MODE_FORCE><<<blocks, 256, 0, st>>>(a);
            else node_gather_b1_kernel<1, MODE_MATVEC><<<blocks, 256, 0, st>>>(a);
        } else {
            if (mode == MODE_FORCE) node_gather_b1_kernel<2, MODE_FORCE><<<blocks, 256, 0, st>>>(a);
            else node_gather_b1_kernel<2, MODE_MATVEC><<<blocks, 256, 0, st>>>(a);
        }
        PF_CUDA_CHECK(cudaGetLastError());
        if (half_sq) {
            column_sum_wide_kernel<<<1, 256, 0, st>>>(plan->d_work, blocks, 1, 0.5, half_sq);
            PF_CUDA_CHECK(cudaGetLastError());
        }
        return PF_OK;
    }
    int64_t done = 0;
    if (kind == PF_ELEM_LINEAR || plan->dim == 1) {
        PfGatherCall c{mode, B, B, u, v, E, A, f_ext, mat_batched, fext_batched, load_factor, f_out, r, half_sq, max_strain};
        int rc = pf_patch_gather(plan, c, st, &done);
        if (rc) return rc;
    }
    if (done < B)
        return launch_gather_generic(plan, kind, mode, B, done, B - done, u, v, E, A, mat_batched, f_out, f_ext,
                                     fext_batched, load_factor, r, half_sq, max_strain, st);
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

extern "C" int pf_element_strain(pf_plan* plan, int measure, int64_t B, const double* u, double* strain, void* stream) {
    int rc = pf_plan_activate(plan);
    if (rc) return rc;
    PF_REQUIRE(measure >= 0 && measure <= 2, "unknown strain measure %d", measure);
    PF_REQUIRE(B >= 1 && u && strain, "pf_element_strain: bad argument");
    if (plan->nelem == 0) return PF_OK;
    dim3 block;
    pick_block(B, block);
    dim3 grid((unsigned)((plan->nelem + block.y - 1) / block.y), (unsigned)((B + block.x - 1) / block.x), 1);
    cudaStream_t st = pf_stream_of(stream);
    if (plan->dim == 1)
        element_strain_kernel<1><<<grid, block, 0, st>>>(plan->d_conn, plan->d_elem_geo, plan->d_elem_xy, u, measure, plan->nelem, B, strain);
    else
        element_strain_kernel<2><<<grid, block, 0, st>>>(plan->d_conn, plan->d_elem_geo, plan->d_elem_xy, u, measure, plan->nelem, B, strain);
    PF_CUDA_CHECK(cudaGetLastError());
    return PF_OK;
}
