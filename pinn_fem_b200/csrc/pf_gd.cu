// Device-resident PINN gradient-descent loop (fem/solver.py:200-400).
//
// One CTA runs the complete solve_gd loop of one problem out of shared memory:
// MLP forward at the element centroids, assembly of f_int, residual and losses,
// the closed-form reverse pass (dL/du = K g, dL/dE, dL/dA, MLP backward), two
// Adam updates, boundary-condition zeroing, the history row and the convergence
// test -- without leaving the kernel.  The reference spends ~18 ms per iteration
// in Python/autograd overhead on its 3-element examples; here an iteration is a
// few microseconds, and independent problems (batched inverse problems, load
// cases) run one per CTA across the SMs.
//
// Everything is fp64 and every reduction runs in a fixed order, so a run is
// bitwise reproducible.  Meshes/networks that do not fit the shared-memory
// budget run the same iteration as a sequence of kernels (pf_gd_large.cu).
#include <algorithm>
#include <cstdlib>

#include "pf_element.cuh"
#include "pf_internal.h"
#include "pf_mlp.cuh"

namespace {

constexpr int kMaxNets = 3;

struct NetDev {
    int enabled;     // MLP (1) or scalar (0)
    int active;      // enters the physics (young, area); density never does (fem/nn_assembly.py:207-208)
    int theta_off;   // offset of this net's parameters in the problem's theta vector
    int act_off;     // smem offset (doubles) of activations: layer 0 [nelem][in], layers 1..L [nelem][w]
    int val_off;     // smem: value per element; +nelem: d value / dz per element
    double scale;
    unsigned div_w, div_in;  // magic multipliers: x / w and x / in_dim for x < 65536 (see fast_div)
    PfMlpDesc d;
};

// x / d for x < 2^16 and d < 256 with m = ceil(2^24 / d): the error x (m - 2^24/d) / 2^24 stays below 2^-8,
// less than the gap (1/d) between frac(x/d) and the next integer.  A runtime integer division costs ~25
// instructions in every item loop of this latency-bound kernel.
__device__ __forceinline__ int fast_div(int x, unsigned m) { return (int)__umulhi((unsigned)x << 8, m); }

struct GdArgs {
    // mesh
    const int2* __restrict__ conn;
    const double4* __restrict__ elem_geo;
    const double* __restrict__ centroid;
    const int32_t* __restrict__ inc_ptr;
    const PfIncidence* __restrict__ inc;
    const double4* __restrict__ inc_geo;
    const uint8_t* __restrict__ dof_free;
    int dim, nnode, nelem, ndof, nfree;
    // problems
    double* theta;
    double* u;
    const double* __restrict__ f_ext;
    const int32_t* __restrict__ meas_dofs;
    const double* __restrict__ meas_vals;
    double* history;
    int32_t* n_iters;
    int32_t* converged;
    double* reactions;
    pf_gd_config cfg;
    NetDev nets[kMaxNets];
    int ntheta;       // parameters per problem (enabled nets)
    int n_tensors;    // parameter tensors per problem
    // smem offsets (doubles)
    int o_theta, o_m, o_v, o_g, o_u, o_mu, o_vu, o_gu, o_f, o_r, o_gf, o_fext, o_gval, o_delta, o_red, o_tn;
    // staged mesh (doubles / ints, offsets in doubles)
    int o_incgeo, o_elemgeo, o_mvals, o_ints;
    int ninc, wmax;
    int smem_doubles;
};

// mesh tables staged in shared memory (they are read in every phase of every iteration)
struct MeshS {
    const double4* inc_geo;   // [ninc]
    const double4* elem_geo;  // [nelem]
    const int* inc_ptr;       // [nnode+1]
    const int2* inc;          // [ninc] {elem, nbr}
    const int2* conn;         // [nelem]
    const int* dof_free;      // [ndof]
    const int* meas_dofs;     // [n_measured]
    const int* meas_first;    // [ndof] first measurement of a DOF or -1
    const int* meas_next;     // [n_measured] next measurement of the same DOF or -1
    const double* mvals;      // [n_measured]
};

__device__ __forceinline__ MeshS mesh_view(const GdArgs& a, double* sm) {
    MeshS m;
    m.inc_geo = reinterpret_cast<const double4*>(sm + a.o_incgeo);
    m.elem_geo = reinterpret_cast<const double4*>(sm + a.o_elemgeo);
    m.mvals = sm + a.o_mvals;
    const int* ip = reinterpret_cast<const int*>(sm + a.o_ints);
    m.inc_ptr = ip;
    ip += (a.nnode + 2) & ~1;
    m.inc = reinterpret_cast<const int2*>(ip);
    ip += 2 * a.ninc;
    m.conn = reinterpret_cast<const int2*>(ip);
    ip += 2 * a.nelem;
    m.dof_free = ip;
    ip += a.ndof;
    m.meas_first = ip;
    ip += a.ndof;
    m.meas_dofs = ip;
    ip += a.cfg.n_measured;
    m.meas_next = ip;
    return m;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// dot product with four independent accumulators: the fp64 FMA dependency chain, not the
// instruction count, is what a 20-term dot product costs in this latency-bound kernel
__device__ __forceinline__ double dot4(const double* __restrict__ w, int ws, const double* __restrict__ x, int xs,
                                       int n) {
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    int i = 0;
    for (; i + 3 < n; i += 4) {
        a0 = fma(w[i * ws], x[i * xs], a0);
        a1 = fma(w[(i + 1) * ws], x[(i + 1) * xs], a1);
        a2 = fma(w[(i + 2) * ws], x[(i + 2) * xs], a2);
        a3 = fma(w[(i + 3) * ws], x[(i + 3) * xs], a3);
    }
    for (; i < n; ++i) a0 = fma(w[i * ws], x[i * xs], a0);
    return (a0 + a1) + (a2 + a3);
}

__device__ __forceinline__ const double* act_ptr(const GdArgs& a, const NetDev& n, const double* sm, int l) {
    return sm + n.act_off + (l == 0 ? 0 : a.nelem * n.d.in_dim + (l - 1) * a.nelem * n.d.w);
}

// hidden layer l of both physics nets at every element, one (net, element, neuron) item per thread
__device__ __forceinline__ void forward_hidden(const GdArgs& a, double* sm, int l) {
    const NetDev& n0 = a.nets[0];
    const NetDev& n1 = a.nets[1];
    const int c0 = (n0.enabled && l < n0.d.L) ? a.nelem * n0.d.w : 0;
    const int c1 = (n1.enabled && l < n1.d.L) ? a.nelem * n1.d.w : 0;
    for (int q = threadIdx.x; q < c0 + c1; q += blockDim.x) {
        const NetDev& n = q < c0 ? n0 : n1;
        const int qq = q < c0 ? q : q - c0;
        const int e = fast_div(qq, n.div_w), o = qq - e * n.d.w;
        const int in = l == 0 ? n.d.in_dim : n.d.w;
        const double* th = sm + a.o_theta + n.theta_off;
        const double z = th[n.d.b_off[l] + o] + dot4(th + n.d.w_off[l] + o * in, 1, act_ptr(a, n, sm, l) + e * in, 1, in);
        sm[n.act_off + a.nelem * n.d.in_dim + l * a.nelem * n.d.w + qq] = pf_tanh(z);
    }
}

// output layer + softplus of both nets (scalar slots just broadcast their value)
__device__ __forceinline__ void forward_output(const GdArgs& a, double* sm) {
    for (int q = threadIdx.x; q < 2 * a.nelem; q += blockDim.x) {
        const NetDev& n = a.nets[q < a.nelem ? 0 : 1];
        const int e = q < a.nelem ? q : q - a.nelem;
        double* val = sm + n.val_off;
        if (!n.enabled) {
            val[e] = n.scale;
            continue;
        }
        const double* th = sm + a.o_theta + n.theta_off;
        const double z = th[n.d.b_off[n.d.L]] + dot4(th + n.d.w_off[n.d.L], 1, act_ptr(a, n, sm, n.d.L) + e * n.d.w, 1, n.d.w);
        val[e] = pf_mlp_output(z, n.scale, 1);
        val[a.nelem + e] = pf_mlp_output_grad(z, n.scale, 1);
    }
}

// per node: out = sum over incident elements of (E A / l0) (d . (x_self - x_other)) d  (f_int(x) or K x)
template <class Epilogue>
__device__ __forceinline__ void gather_linear(const GdArgs& a, const MeshS& ms, const double* sm, const double* x,
                                              Epilogue epi) {
    const double* E = sm + a.nets[0].val_off;
    const double* A = sm + a.nets[1].val_off;
    for (int n = threadIdx.x; n < a.nnode; n += blockDim.x) {
        double fx = 0.0, fy = 0.0;
        const double xs = a.dim == 2 ? x[2 * n] : x[n];
        const double ys = a.dim == 2 ? x[2 * n + 1] : 0.0;
        for (int k = ms.inc_ptr[n]; k < ms.inc_ptr[n + 1]; ++k) {
            const int2 inc = ms.inc[k];
            const double4 geo = ms.inc_geo[k];
            if (a.dim == 2)
                pf_linear_incidence<2>(E[inc.x], A[inc.x], geo, xs, ys, x[2 * inc.y], x[2 * inc.y + 1], fx, fy);
            else
                pf_linear_incidence<1>(E[inc.x], A[inc.x], geo, xs, 0.0, x[inc.y], 0.0, fx, fy);
        }
        if (a.dim == 2) {
            epi(2 * n, fx);
            epi(2 * n + 1, fy);
        } else {
            epi(n, fx);
        }
    }
}

__global__ void __launch_bounds__(1024) gd_solve_kernel(GdArgs a) {
    extern __shared__ __align__(16) double sm[];
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarp = nt >> 5;
    const int64_t prob = blockIdx.x;
    const pf_gd_config& cfg = a.cfg;
    double* th = sm + a.o_theta;
    double* u = sm + a.o_u;
    double* r = sm + a.o_r;
    double* gf = sm + a.o_gf;
    double* gu = sm + a.o_gu;
    double* red = sm + a.o_red;
    const double lam = cfg.load_factor;
    const bool has_meas = cfg.n_measured > 0 && a.meas_dofs && a.meas_vals && cfg.alpha_data > 0.0;
    const bool legacy = cfg.loss_mode == 1;  // fem/nn_solver_gd.py: mean-squared physics loss
    const double gscale = legacy ? 2.0 * cfg.alpha_physics / (double)a.nfree : cfg.alpha_physics;
    const MeshS ms = mesh_view(a, sm);

    // ---- stage the mesh tables and the problem in shared memory ----
    {
        double4* g1 = const_cast<double4*>(ms.inc_geo);
        int2* i1 = const_cast<int2*>(ms.inc);
        for (int i = tid; i < a.ninc; i += nt) {
            g1[i] = a.inc_geo[i];
            i1[i] = make_int2(a.inc[i].elem, a.inc[i].nbr);
        }
        double4* g2 = const_cast<double4*>(ms.elem_geo);
        int2* c2 = const_cast<int2*>(ms.conn);
        for (int i = tid; i < a.nelem; i += nt) {
            g2[i] = a.elem_geo[i];
            c2[i] = a.conn[i];
        }
        int* p1 = const_cast<int*>(ms.inc_ptr);
        for (int i = tid; i <= a.nnode; i += nt) p1[i] = a.inc_ptr[i];
        int* f1 = const_cast<int*>(ms.dof_free);
        int* mf = const_cast<int*>(ms.meas_first);
        for (int i = tid; i < a.ndof; i += nt) {
            f1[i] = a.dof_free[i];
            mf[i] = -1;
        }
        int* m1 = const_cast<int*>(ms.meas_dofs);
        double* m2 = const_cast<double*>(ms.mvals);
        for (int i = tid; i < cfg.n_measured; i += nt) {
            m1[i] = a.meas_dofs[i];
            m2[i] = a.meas_vals[prob * cfg.n_measured + i];
        }
    }
    for (int i = tid; i < a.ntheta; i += nt) {
        th[i] = a.theta[prob * a.ntheta + i];
        sm[a.o_m + i] = 0.0;
        sm[a.o_v + i] = 0.0;
        sm[a.o_g + i] = 0.0;
    }
    for (int i = tid; i < a.ndof; i += nt) {
        u[i] = a.u[prob * a.ndof + i];
        sm[a.o_mu + i] = 0.0;
        sm[a.o_vu + i] = 0.0;
        sm[a.o_fext + i] = a.f_ext[i];
    }
    for (int k = 0; k < 2; ++k) {
        const NetDev& n = a.nets[k];
        if (!n.enabled) continue;
        double* x = sm + n.act_off;  // inputs [load_factor, centroid...]: sorted dict keys (fem/properties.py:116-125)
        for (int q = tid; q < a.nelem * n.d.in_dim; q += nt) {
            const int e = q / n.d.in_dim, i = q % n.d.in_dim;
            x[q] = i == 0 ? lam : a.centroid[e * (n.d.in_dim - 1) + (i - 1)];
        }
    }
    __syncthreads();
    if (tid == 0 && has_meas) {  // per-DOF measurement lists in ascending measurement order (DOFs may repeat)
        int* mf = const_cast<int*>(ms.meas_first);
        int* mn = const_cast<int*>(ms.meas_next);
        for (int j = cfg.n_measured - 1; j >= 0; --j) {
            mn[j] = mf[ms.meas_dofs[j]];
            mf[ms.meas_dofs[j]] = j;
        }
    }
    if (tid == 0) red[5] = 0.0;
    __syncthreads();

    // parameter tensors in nn.Module.parameters() order: offset and size of each (monitoring norms)
    __shared__ int s_toff[3 * 2 * (PF_MLP_MAX_LAYERS + 1)], s_tcnt[3 * 2 * (PF_MLP_MAX_LAYERS + 1)];
    if (tid == 0) {
        int tix = 0;
        for (int k = 0; k < kMaxNets; ++k) {
            const NetDev& n = a.nets[k];
            if (!n.enabled) continue;
            for (int l = 0; l <= n.d.L; ++l) {
                const int in = l == 0 ? n.d.in_dim : n.d.w;
                s_toff[tix] = n.theta_off + n.d.w_off[l];
                s_tcnt[tix++] = l == n.d.L ? n.d.w : n.d.w * in;
                s_toff[tix] = n.theta_off + n.d.b_off[l];
                s_tcnt[tix++] = l == n.d.L ? 1 : n.d.w;
            }
        }
    }
    int maxL = 0;
    for (int k = 0; k < 2; ++k)
        if (a.nets[k].enabled) maxL = max(maxL, a.nets[k].d.L);
    const int n_active_theta = a.nets[2].enabled ? a.nets[2].theta_off : a.ntheta;  // density: grad None, never stepped
    const double cdata = has_meas ? -2.0 * cfg.alpha_data / cfg.n_measured : 0.0;

    int it_done = 0, conv = 0;
    double pow_b1 = 1.0, pow_b2 = 1.0;  // beta^t, updated multiplicatively
    for (int it = 0; it <= cfg.max_iterations; ++it) {
        // ---- forward, first layer.  Thread 0 first closes the previous iteration: history row and
        //      convergence flag (solver.py:308-355); everyone reads the flag after the barrier. ----
        if (tid == 0 && it > 0) {
            double tn = 0.0;
            for (int q = 0; q < a.n_tensors; ++q) tn += sm[a.o_tn + q];
            if (a.history) {
                double* h = a.history + (prob * cfg.max_iterations + (it - 1)) * PF_GD_HISTORY_COLS;
                h[0] = (double)it;
                h[1] = red[2];
                h[2] = red[0];
                h[3] = cfg.n_measured > 0 ? red[1] : 0.0;
                h[4] = red[4];
                h[5] = red[3];
                h[6] = tn;
            }
            int c = 0;
            if (it - 1 > 10) {  // solver.py:341-355 (legacy nn_solver_gd.py:171: loss only)
                if (!legacy && red[3] < cfg.tolerance) c = 1;
                else if (!isnan(red[2]) && red[2] < cfg.tolerance) c = 1;
            }
            red[5] = (double)c;
        }
        if (maxL > 0) forward_hidden(a, sm, 0);
        __syncthreads();
        it_done = it;
        if (red[5] != 0.0) {
            conv = 1;
            break;
        }
        if (it == cfg.max_iterations) break;
        for (int l = 1; l < maxL; ++l) {
            forward_hidden(a, sm, l);
            __syncthreads();
        }
        forward_output(a, sm);
        __syncthreads();
        // ---- f_int, residual and dL/df_int in one pass (solver.py:262-270) ----
        gather_linear(a, ms, sm, u, [&](int d, double f) {
            const double rr = ms.dof_free[d] ? __dsub_rn(f, __dmul_rn(lam, sm[a.o_fext + d])) : 0.0;
            r[d] = rr;
            gf[d] = gscale * rr;
        });
        __syncthreads();
        // ---- reverse pass: dL/du = K g_f + data term, dL/dE, dL/dA; the last warp reduces the losses ----
        gather_linear(a, ms, sm, gf, [&](int d, double g) {
            for (int j = ms.meas_first[d]; j >= 0; j = ms.meas_next[j]) g += cdata * (ms.mvals[j] - u[d]);
            gu[d] = g;
        });
        {
            const double* E = sm + a.nets[0].val_off;
            const double* A = sm + a.nets[1].val_off;
            double* gE = sm + a.o_gval;
            double* gA = gE + a.nelem;
            for (int e = tid; e < a.nelem; e += nt) {
                const int2 c = ms.conn[e];
                const double4 geo = ms.elem_geo[e];
                double gh;
                if (a.dim == 2) {
                    const double axial = geo.x * (u[2 * c.x] - u[2 * c.y]) + geo.y * (u[2 * c.x + 1] - u[2 * c.y + 1]);
                    gh = geo.z * axial * (geo.x * (gf[2 * c.x] - gf[2 * c.y]) + geo.y * (gf[2 * c.x + 1] - gf[2 * c.y + 1]));
                } else {
                    gh = geo.z * (u[c.x] - u[c.y]) * (gf[c.x] - gf[c.y]);
                }
                // dL/dz of the two output layers: dL/dvalue * dvalue/dz
                gE[e] = A[e] * gh * (a.nets[0].enabled ? sm[a.nets[0].val_off + a.nelem + e] : 0.0);
                gA[e] = E[e] * gh * (a.nets[1].enabled ? sm[a.nets[1].val_off + a.nelem + e] : 0.0);
            }
        }
        if (warp == nwarp - 1) {
            double s = 0.0;
            for (int d = lane; d < a.ndof; d += 32) s += r[d] * r[d];
            s = warp_sum(s);
            double sd = 0.0;
            if (has_meas) {
                for (int j = lane; j < cfg.n_measured; j += 32) {
                    const double dd = ms.mvals[j] - u[ms.meas_dofs[j]];
                    sd += dd * dd;
                }
                sd = warp_sum(sd) / cfg.n_measured;
            }
            if (lane == 0) {
                red[0] = legacy ? s / (double)a.nfree : 0.5 * s;                              // loss_physics
                red[1] = sd;                                                                  // loss_data
                red[2] = cfg.alpha_physics * red[0] + (has_meas ? cfg.alpha_data * sd : 0.0);  // loss_total
                red[3] = sqrt(s);                                                             // ||r||
            }
        }
        __syncthreads();
        // ---- MLP backward of both nets (closed form of autograd through NNProperty.value) ----
        {
            // output-layer gradients and the delta of the last hidden layer
            int c0[2];
            for (int k = 0; k < 2; ++k) {
                c0[k] = a.nets[k].enabled ? a.nets[k].d.w + 1 + a.nelem * a.nets[k].d.w : 0;
            }
            for (int q = tid; q < c0[0] + c0[1]; q += nt) {
                const int k = q < c0[0] ? 0 : 1;
                const NetDev& n = a.nets[k];
                int qq = q - (k ? c0[0] : 0);
                const int w = n.d.w, L = n.d.L;
                const double* dz = sm + a.o_gval + k * a.nelem;
                const double* aL = act_ptr(a, n, sm, L);
                double* g = sm + a.o_g + n.theta_off;
                if (qq < w) {
                    g[n.d.w_off[L] + qq] = dot4(dz, 1, aL + qq, w, a.nelem);
                } else if (qq == w) {
                    double acc = 0.0;
                    for (int e = 0; e < a.nelem; ++e) acc += dz[e];
                    g[n.d.b_off[L]] = acc;
                } else {
                    qq -= w + 1;
                    const int e = fast_div(qq, n.div_w), o = qq - e * w;
                    const double av = aL[qq];
                    (sm + a.o_delta + k * 2 * a.nelem * a.wmax)[qq] = th[n.theta_off + n.d.w_off[L] + o] * dz[e] * (1.0 - av * av);
                }
            }
            __syncthreads();
            for (int step = 0; step < maxL; ++step) {
                // net k is at layer l = L_k - 1 - step; ping-pong delta buffers per net
                int cnt[2];
                for (int k = 0; k < 2; ++k) {
                    const NetDev& n = a.nets[k];
                    const int l = n.enabled ? n.d.L - 1 - step : -1;
                    const int in = l == 0 ? n.d.in_dim : n.d.w;
                    cnt[k] = l >= 0 ? n.d.w * in + n.d.w + (l > 0 ? a.nelem * n.d.w : 0) : 0;
                }
                for (int q = tid; q < cnt[0] + cnt[1]; q += nt) {
                    const int k = q < cnt[0] ? 0 : 1;
                    const NetDev& n = a.nets[k];
                    int qq = q - (k ? cnt[0] : 0);
                    const int w = n.d.w, l = n.d.L - 1 - step;
                    const int in = l == 0 ? n.d.in_dim : w;
                    const double* D = sm + a.o_delta + k * 2 * a.nelem * a.wmax + (step & 1) * a.nelem * a.wmax;
                    double* Dn = sm + a.o_delta + k * 2 * a.nelem * a.wmax + ((step + 1) & 1) * a.nelem * a.wmax;
                    const double* ain = act_ptr(a, n, sm, l);
                    double* g = sm + a.o_g + n.theta_off;
                    if (qq < w * in) {
                        const int o = fast_div(qq, l == 0 ? n.div_in : n.div_w), i = qq - o * in;
                        g[n.d.w_off[l] + qq] = dot4(D + o, w, ain + i, in, a.nelem);
                    } else if (qq < w * in + w) {
                        const int o = qq - w * in;
                        double acc = 0.0;
                        for (int e = 0; e < a.nelem; ++e) acc += D[e * w + o];
                        g[n.d.b_off[l] + o] = acc;
                    } else {
                        qq -= w * in + w;
                        const int e = fast_div(qq, n.div_w), i = qq - e * w;
                        const double av = ain[qq];
                        Dn[qq] = dot4(th + n.theta_off + n.d.w_off[l] + i, w, D + e * w, 1, w) * (1.0 - av * av);
                    }
                }
                __syncthreads();
            }
        }
        // ---- two Adam steps (torch.optim.Adam defaults) and BC zeroing (solver.py:292-298) ----
        {
            const double b1 = 0.9, b2 = 0.999, eps = 1e-8;
            pow_b1 *= b1;
            pow_b2 *= b2;
            const double bc1 = 1.0 - pow_b1, bc2s = sqrt(1.0 - pow_b2);
            const double step_u = cfg.learning_rate_u / bc1, step_t = cfg.learning_rate_theta / bc1;
            for (int q = tid; q < n_active_theta + a.ndof; q += nt) {
                const bool is_u = q >= n_active_theta;
                const int p = is_u ? q - n_active_theta : q;
                double* pm = sm + (is_u ? a.o_mu : a.o_m) + p;
                double* pv = sm + (is_u ? a.o_vu : a.o_v) + p;
                const double g = is_u ? gu[p] : sm[a.o_g + p];
                const double m = *pm + (g - *pm) * (1.0 - b1);
                const double v = *pv * b2 + (1.0 - b2) * g * g;
                *pm = m;
                *pv = v;
                const double upd = (-(is_u ? step_u : step_t) * m) / (sqrt(v) / bc2s + eps);  // addcdiv_: (value*t1)/t2
                if (is_u)
                    u[p] = ms.dof_free[p] ? u[p] + upd : 0.0;
                else
                    th[p] += upd;
            }
        }
        __syncthreads();
        // ---- monitoring (solver.py:304-322): ||u_free||, per-tensor parameter norms; read by thread 0
        //      at the top of the next iteration ----
        {
            for (int tix = warp; tix < a.n_tensors; tix += nwarp) {  // tensor table built once before the loop
                const int off = s_toff[tix], cnt = s_tcnt[tix];
                double s0 = 0.0, s1 = 0.0;
                for (int i = lane; i < cnt; i += 64) {
                    s0 = fma(th[off + i], th[off + i], s0);
                    if (i + 32 < cnt) s1 = fma(th[off + i + 32], th[off + i + 32], s1);
                }
                const double s = warp_sum(s0 + s1);
                if (lane == 0) sm[a.o_tn + tix] = sqrt(s);
            }
            if (warp == nwarp - 1) {
                double s = 0.0;
                for (int d = lane; d < a.ndof; d += 32) s += ms.dof_free[d] ? u[d] * u[d] : 0.0;
                s = warp_sum(s);
                if (lane == 0) red[4] = sqrt(s);
            }
        }
        __syncthreads();  // norms visible to thread 0, which closes this iteration at the top of the next one
    }

    // ---- results: theta, u, reactions f_int - lambda f_ext on fixed DOFs (solver.py:374-380) ----
    __syncthreads();
    for (int l = 0; l < maxL; ++l) {
        forward_hidden(a, sm, l);
        __syncthreads();
    }
    forward_output(a, sm);
    __syncthreads();
    gather_linear(a, ms, sm, u, [&](int d, double f) { sm[a.o_f + d] = f; });
    __syncthreads();
    for (int i = tid; i < a.ntheta; i += nt) a.theta[prob * a.ntheta + i] = th[i];
    for (int d = tid; d < a.ndof; d += nt) {
        a.u[prob * a.ndof + d] = u[d];
        if (a.reactions)
            a.reactions[prob * a.ndof + d] = ms.dof_free[d] ? 0.0 : __dsub_rn(sm[a.o_f + d], __dmul_rn(lam, sm[a.o_fext + d]));
    }
    if (tid == 0) {
        a.n_iters[prob] = it_done;
        a.converged[prob] = conv;
    }
}

}  // namespace

extern "C" int pf_gd_solve(pf_plan* plan, const pf_gd_config* cfg, int64_t nprob, double* theta, double* u,
                           const double* f_ext, const int32_t* meas_dofs, const double* meas_vals, double* history,
                           int32_t* n_iters, int32_t* converged, double* reactions, void* stream) {
    int rc = pf_plan_activate(plan);
    if (rc) return rc;
    PF_REQUIRE(cfg != nullptr, "cfg is NULL");
    PF_REQUIRE(nprob >= 1, "nprob must be >= 1");
    PF_REQUIRE(u && f_ext && n_iters && converged, "pf_gd_solve: NULL argument");
    PF_REQUIRE(cfg->kind == PF_ELEM_LINEAR, "pf_gd_solve supports the linear element only (the reference wires nothing else)");
    PF_REQUIRE(cfg->max_iterations >= 0, "max_iterations must be >= 0");
    PF_REQUIRE(cfg->n_measured == 0 || (meas_dofs && meas_vals), "measurements are NULL");

    GdArgs a{};
    a.conn = plan->d_conn;
    a.elem_geo = plan->d_elem_geo;
    a.centroid = plan->d_centroid;
    a.inc_ptr = plan->d_inc_ptr;
    a.inc = plan->d_inc;
    a.inc_geo = plan->d_inc_geo;
    a.dof_free = plan->d_dof_free;
    a.dim = plan->dim;
    a.nnode = (int)plan->nnode;
    a.nelem = (int)plan->nelem;
    a.ndof = (int)plan->ndof;
    a.nfree = (int)plan->nfree;
    a.theta = theta;
    a.u = u;
    a.f_ext = f_ext;
    a.meas_dofs = meas_dofs;
    a.meas_vals = meas_vals;
    a.history = history;
    a.n_iters = n_iters;
    a.converged = converged;
    a.reactions = reactions;
    a.cfg = *cfg;

    int off = 0, ntheta = 0, ntens = 0, max_w = 1;
    auto take = [&](int n) {
        const int o = off;
        off += (n + 1) & ~1;  // keep 16-byte alignment
        return o;
    };
    for (int k = 0; k < kMaxNets; ++k) {
        NetDev& n = a.nets[k];
        n.enabled = cfg->net_enabled[k] != 0;
        n.active = k < 2;
        n.scale = cfg->net_scale[k];
        n.theta_off = ntheta;
        if (n.enabled) {
            rc = pf_mlp_make_desc(cfg->net_input_dim[k], cfg->net_hidden_layers[k], cfg->net_width[k], &n.d);
            if (rc) return rc;
            PF_REQUIRE(n.d.in_dim == plan->dim + 1,
                       "network input_dim %d does not match [load_factor, centroid] = %d inputs "
                       "(fem/properties.py:116-125)", n.d.in_dim, plan->dim + 1);
            n.div_w = ((1u << 24) + n.d.w - 1) / n.d.w;
            n.div_in = ((1u << 24) + n.d.in_dim - 1) / n.d.in_dim;
            ntheta += n.d.n_params;
            ntens += 2 * (n.d.L + 1);
            max_w = std::max(max_w, n.d.w);
        }
    }
    PF_REQUIRE(ntheta == 0 || theta != nullptr, "theta is NULL");
    a.ntheta = ntheta;
    a.n_tensors = ntens;
    a.o_theta = take(ntheta);
    a.o_m = take(ntheta);
    a.o_v = take(ntheta);
    a.o_g = take(ntheta);
    a.o_u = take(a.ndof);
    a.o_mu = take(a.ndof);
    a.o_vu = take(a.ndof);
    a.o_gu = take(a.ndof);
    a.o_f = take(a.ndof);
    a.o_r = take(a.ndof);
    a.o_gf = take(a.ndof);
    a.o_fext = take(a.ndof);
    a.o_gval = take(2 * a.nelem);
    a.wmax = max_w;
    a.o_delta = take(4 * a.nelem * max_w);
    a.o_red = take(8);
    a.o_tn = take(std::max(ntens, 1));
    a.ninc = (int)plan->ninc;
    a.o_incgeo = take(4 * a.ninc);
    a.o_elemgeo = take(4 * a.nelem);
    a.o_mvals = take(std::max(cfg->n_measured, 1));
    {
        const int nints = ((a.nnode + 2) & ~1) + 2 * a.ninc + 2 * a.nelem + 2 * a.ndof + 2 * cfg->n_measured + 2;
        a.o_ints = take((nints + 1) / 2);
    }
    for (int k = 0; k < kMaxNets; ++k) {
        NetDev& n = a.nets[k];
        n.val_off = take(2 * a.nelem);
        n.act_off = (n.enabled && n.active) ? take(a.nelem * (n.d.in_dim + n.d.L * n.d.w)) : 0;
    }
    a.smem_doubles = off;
    const size_t smem = (size_t)off * sizeof(double);
    if (smem > 200 * 1024 || (int64_t)a.nelem * max_w >= 65536)  // does not fit one CTA (fast_div range: items < 2^16): the multi-kernel device-resident loop (pf_gd_large.cu)
        return pf_gd_solve_large(plan, cfg, nprob, theta, u, f_ext, meas_dofs, meas_vals, history, n_iters, converged,
                                 reactions, pf_stream_of(stream), nullptr);
    if (smem > 48 * 1024)
        PF_CUDA_CHECK(cudaFuncSetAttribute(gd_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int work = std::max({a.nelem * max_w, a.ndof, max_w * max_w + max_w, 32});
    int threads = work <= 64 ? 64 : work <= 128 ? 128 : 256;
    // few problems: the SMs are mostly idle and each phase is a dependent fp64 chain per item (tanh, sqrt,
    // division), so more threads = fewer items per thread = shorter phases (10.4 vs 12.6 us per iteration at
    // one problem); many problems: 256-thread CTAs keep more problems resident per SM (30.8 vs 12.1 M it/s).
    if (work > 256 && nprob <= plan->sm_count) threads = 1024;
    else if (work > 256 && nprob <= 2 * (int64_t)plan->sm_count) threads = 512;
    static const int env_threads = getenv("PF_GD_THREADS") ? atoi(getenv("PF_GD_THREADS")) : 0;
    if (env_threads >= 32 && env_threads <= 1024 && env_threads % 32 == 0) threads = env_threads;
    gd_solve_kernel<<<(unsigned)nprob, threads, smem, pf_stream_of(stream)>>>(a);
    PF_CUDA_CHECK(cudaGetLastError());
    return PF_OK;
}
