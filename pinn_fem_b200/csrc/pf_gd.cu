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
// budget are rejected with PF_ERR_ARG (the host then uses the multi-kernel path).
#include <algorithm>

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
    PfMlpDesc d;
};

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
    int smem_doubles;
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// forward of all active nets at every element centroid
__device__ void nets_forward(const GdArgs& a, double* sm) {
    const int tid = threadIdx.x, nt = blockDim.x;
    int maxL = 0;
    for (int k = 0; k < kMaxNets; ++k)
        if (a.nets[k].enabled && a.nets[k].active) maxL = max(maxL, a.nets[k].d.L);
    for (int l = 0; l < maxL; ++l) {
        for (int k = 0; k < kMaxNets; ++k) {
            const NetDev& n = a.nets[k];
            if (!n.enabled || !n.active || l >= n.d.L) continue;
            const int in = l == 0 ? n.d.in_dim : n.d.w;
            const double* W = sm + a.o_theta + n.theta_off + n.d.w_off[l];
            const double* b = sm + a.o_theta + n.theta_off + n.d.b_off[l];
            const double* ain = sm + n.act_off + (l == 0 ? 0 : a.nelem * n.d.in_dim + (l - 1) * a.nelem * n.d.w);
            double* aout = sm + n.act_off + a.nelem * n.d.in_dim + l * a.nelem * n.d.w;
            for (int it = tid; it < a.nelem * n.d.w; it += nt) {
                const int e = it / n.d.w, o = it % n.d.w;
                double acc = b[o];
                for (int i = 0; i < in; ++i) acc = fma(W[o * in + i], ain[e * in + i], acc);
                aout[e * n.d.w + o] = tanh(acc);
            }
        }
        __syncthreads();
    }
    for (int k = 0; k < kMaxNets; ++k) {
        const NetDev& n = a.nets[k];
        if (!n.active) continue;
        double* val = sm + n.val_off;
        if (!n.enabled) {
            for (int e = tid; e < a.nelem; e += nt) val[e] = n.scale;
            continue;
        }
        const double* wo = sm + a.o_theta + n.theta_off + n.d.w_off[n.d.L];
        const double bo = sm[a.o_theta + n.theta_off + n.d.b_off[n.d.L]];
        const double* aL = sm + n.act_off + a.nelem * n.d.in_dim + (n.d.L - 1) * a.nelem * n.d.w;
        for (int e = tid; e < a.nelem; e += nt) {
            double z = bo;
            for (int i = 0; i < n.d.w; ++i) z = fma(wo[i], aL[e * n.d.w + i], z);
            val[e] = pf_mlp_output(z, n.scale, 1);
            val[a.nelem + e] = pf_mlp_output_grad(z, n.scale, 1);
        }
    }
    __syncthreads();
}

// out[dof] = sum over incident elements of (E A / l0) (d . (x_self - x_other)) d  -- f_int(x) or K x
__device__ void gather_linear(const GdArgs& a, const double* sm, const double* x, double* out) {
    const double* E = sm + a.nets[0].val_off;
    const double* A = sm + a.nets[1].val_off;
    for (int n = threadIdx.x; n < a.nnode; n += blockDim.x) {
        double fx = 0.0, fy = 0.0;
        const double xs = a.dim == 2 ? x[2 * n] : x[n];
        const double ys = a.dim == 2 ? x[2 * n + 1] : 0.0;
        for (int k = a.inc_ptr[n]; k < a.inc_ptr[n + 1]; ++k) {
            const PfIncidence inc = a.inc[k];
            const double4 geo = a.inc_geo[k];
            if (a.dim == 2)
                pf_linear_incidence<2>(E[inc.elem], A[inc.elem], geo, xs, ys, x[2 * inc.nbr], x[2 * inc.nbr + 1], fx, fy);
            else
                pf_linear_incidence<1>(E[inc.elem], A[inc.elem], geo, xs, 0.0, x[inc.nbr], 0.0, fx, fy);
        }
        if (a.dim == 2) {
            out[2 * n] = fx;
            out[2 * n + 1] = fy;
        } else {
            out[n] = fx;
        }
    }
}

__global__ void __launch_bounds__(256) gd_solve_kernel(GdArgs a) {
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
    const double* mvals = a.meas_vals ? a.meas_vals + prob * cfg.n_measured : nullptr;

    // ---- load the problem into shared memory ----
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
    for (int k = 0; k < kMaxNets; ++k) {
        const NetDev& n = a.nets[k];
        if (!n.enabled || !n.active) continue;
        double* x = sm + n.act_off;  // inputs [load_factor, centroid...]: sorted dict keys (fem/properties.py:116-125)
        for (int it = tid; it < a.nelem * n.d.in_dim; it += nt) {
            const int e = it / n.d.in_dim, i = it % n.d.in_dim;
            x[it] = i == 0 ? lam : a.centroid[e * (n.d.in_dim - 1) + (i - 1)];
        }
    }
    __syncthreads();

    int it_done = 0, conv = 0;
    for (int it = 0; it < cfg.max_iterations; ++it) {
        // ---- forward: materials, internal force, residual (solver.py:262-269) ----
        nets_forward(a, sm);
        gather_linear(a, sm, u, sm + a.o_f);
        __syncthreads();
        const bool legacy = cfg.loss_mode == 1;  // fem/nn_solver_gd.py: mean-squared physics loss
        const double gscale = legacy ? 2.0 * cfg.alpha_physics / (double)a.nfree : cfg.alpha_physics;
        for (int d = tid; d < a.ndof; d += nt) {
            const double rr = a.dof_free[d] ? __dsub_rn(sm[a.o_f + d], __dmul_rn(lam, sm[a.o_fext + d])) : 0.0;
            r[d] = rr;
            gf[d] = gscale * rr;  // dL/df_int
        }
        __syncthreads();
        // ---- losses (solver.py:270-283): warp 0 reduces in a fixed order ----
        if (warp == 0) {
            double s = 0.0;
            for (int d = lane; d < a.ndof; d += 32) s += r[d] * r[d];
            s = warp_sum(s);
            double sd = 0.0;
            if (has_meas) {
                for (int j = lane; j < cfg.n_measured; j += 32) {
                    const double dd = mvals[j] - u[a.meas_dofs[j]];
                    sd += dd * dd;
                }
                sd = warp_sum(sd) / cfg.n_measured;
            }
            if (lane == 0) {
                red[0] = legacy ? s / (double)a.nfree : 0.5 * s;          // loss_physics
                red[1] = sd;                                               // loss_data
                red[2] = cfg.alpha_physics * red[0] + (has_meas ? cfg.alpha_data * sd : 0.0);  // loss_total
                red[3] = sqrt(s);                                          // ||r||
            }
        }
        // ---- reverse pass: dL/du = K g_f (+ data term), dL/dE, dL/dA ----
        gather_linear(a, sm, gf, gu);
        {
            const double* E = sm + a.nets[0].val_off;
            const double* A = sm + a.nets[1].val_off;
            double* gE = sm + a.o_gval;
            double* gA = gE + a.nelem;
            for (int e = tid; e < a.nelem; e += nt) {
                const int2 c = a.conn[e];
                const double4 geo = a.elem_geo[e];
                double gh;
                if (a.dim == 2) {
                    const double axial = geo.x * (u[2 * c.x] - u[2 * c.y]) + geo.y * (u[2 * c.x + 1] - u[2 * c.y + 1]);
                    gh = geo.z * axial * (geo.x * (gf[2 * c.x] - gf[2 * c.y]) + geo.y * (gf[2 * c.x + 1] - gf[2 * c.y + 1]));
                } else {
                    gh = geo.z * (u[c.x] - u[c.y]) * (gf[c.x] - gf[c.y]);
                }
                gE[e] = A[e] * gh;
                gA[e] = E[e] * gh;
            }
        }
        __syncthreads();
        if (has_meas && tid == 0) {  // d mean((m - u)^2)/du, measured DOFs may repeat: serial, ordered
            const double c = -2.0 * cfg.alpha_data / cfg.n_measured;
            for (int j = 0; j < cfg.n_measured; ++j) gu[a.meas_dofs[j]] += c * (mvals[j] - u[a.meas_dofs[j]]);
        }
        // ---- MLP backward (closed form of the autograd pass through NNProperty.value) ----
        for (int k = 0; k < 2; ++k) {
            const NetDev& n = a.nets[k];
            if (!n.enabled) continue;
            const int w = n.d.w, L = n.d.L;
            const double* gval = sm + a.o_gval + k * a.nelem;
            const double* sg = sm + n.val_off + a.nelem;
            double* g = sm + a.o_g + n.theta_off;
            const double* aL = sm + n.act_off + a.nelem * n.d.in_dim + (L - 1) * a.nelem * w;
            const double* wo = th + n.theta_off + n.d.w_off[L];
            double* D = sm + a.o_delta;
            double* Dn = D + a.nelem * w;
            for (int q = tid; q <= w; q += nt) {  // output layer: W_out[i], b_out
                double acc = 0.0;
                for (int e = 0; e < a.nelem; ++e) acc = fma(gval[e] * sg[e], q < w ? aL[e * w + q] : 1.0, acc);
                g[q < w ? n.d.w_off[L] + q : n.d.b_off[L]] = acc;
            }
            for (int q = tid; q < a.nelem * w; q += nt) {
                const int e = q / w, o = q % w;
                const double av = aL[q];
                D[q] = wo[o] * (gval[e] * sg[e]) * (1.0 - av * av);
            }
            __syncthreads();
            for (int l = L - 1; l >= 0; --l) {
                const int in = l == 0 ? n.d.in_dim : w;
                const double* ain = sm + n.act_off + (l == 0 ? 0 : a.nelem * n.d.in_dim + (l - 1) * a.nelem * w);
                for (int q = tid; q < w * in + w; q += nt) {
                    double acc = 0.0;
                    if (q < w * in) {
                        const int o = q / in, i = q % in;
                        for (int e = 0; e < a.nelem; ++e) acc = fma(D[e * w + o], ain[e * in + i], acc);
                        g[n.d.w_off[l] + q] = acc;
                    } else {
                        const int o = q - w * in;
                        for (int e = 0; e < a.nelem; ++e) acc += D[e * w + o];
                        g[n.d.b_off[l] + o] = acc;
                    }
                }
                if (l > 0) {
                    const double* W = th + n.theta_off + n.d.w_off[l];
                    for (int q = tid; q < a.nelem * w; q += nt) {
                        const int e = q / w, i = q % w;
                        double acc = 0.0;
                        for (int o = 0; o < w; ++o) acc = fma(W[o * w + i], D[e * w + o], acc);
                        const double av = ain[q];
                        Dn[q] = acc * (1.0 - av * av);
                    }
                }
                __syncthreads();
                double* tmp = D;
                D = Dn;
                Dn = tmp;
            }
        }
        __syncthreads();
        // ---- two Adam steps (torch.optim.Adam defaults) and BC zeroing (solver.py:292-298) ----
        {
            const double b1 = 0.9, b2 = 0.999, eps = 1e-8;
            const double t = (double)(it + 1);
            const double bc1 = 1.0 - pow(b1, t), bc2s = sqrt(1.0 - pow(b2, t));
            const double step_u = cfg.learning_rate_u / bc1, step_t = cfg.learning_rate_theta / bc1;
            for (int d = tid; d < a.ndof; d += nt) {
                const double g = gu[d];
                const double m = sm[a.o_mu + d] + (g - sm[a.o_mu + d]) * (1.0 - b1);
                const double v = sm[a.o_vu + d] * b2 + (1.0 - b2) * g * g;
                sm[a.o_mu + d] = m;
                sm[a.o_vu + d] = v;
                const double un = u[d] + (-step_u * m) / (sqrt(v) / bc2s + eps);  // addcdiv_: (value*t1)/t2
                u[d] = a.dof_free[d] ? un : 0.0;
            }
            for (int k = 0; k < 2; ++k) {  // density has grad None: Adam skips it
                const NetDev& n = a.nets[k];
                if (!n.enabled) continue;
                for (int i = tid; i < n.d.n_params; i += nt) {
                    const int p = n.theta_off + i;
                    const double g = sm[a.o_g + p];
                    const double m = sm[a.o_m + p] + (g - sm[a.o_m + p]) * (1.0 - b1);
                    const double v = sm[a.o_v + p] * b2 + (1.0 - b2) * g * g;
                    sm[a.o_m + p] = m;
                    sm[a.o_v + p] = v;
                    th[p] += (-step_t * m) / (sqrt(v) / bc2s + eps);
                }
            }
        }
        __syncthreads();
        // ---- monitoring (solver.py:304-322): ||u_free||, sum of per-tensor parameter norms ----
        {
            int tix = 0;
            for (int k = 0; k < kMaxNets; ++k) {
                const NetDev& n = a.nets[k];
                if (!n.enabled) continue;
                for (int l = 0; l <= n.d.L; ++l) {
                    for (int wb = 0; wb < 2; ++wb, ++tix) {
                        if (tix % nwarp != warp) continue;
                        const int off = n.theta_off + (wb == 0 ? n.d.w_off[l] : n.d.b_off[l]);
                        const int in = l == 0 ? n.d.in_dim : n.d.w;
                        const int cnt = wb == 0 ? (l == n.d.L ? n.d.w : n.d.w * in) : (l == n.d.L ? 1 : n.d.w);
                        double s = 0.0;
                        for (int i = lane; i < cnt; i += 32) s += th[off + i] * th[off + i];
                        s = warp_sum(s);
                        if (lane == 0) sm[a.o_tn + tix] = sqrt(s);
                    }
                }
            }
            if (warp == nwarp - 1) {
                double s = 0.0;
                for (int d = lane; d < a.ndof; d += 32) s += a.dof_free[d] ? u[d] * u[d] : 0.0;
                s = warp_sum(s);
                if (lane == 0) red[4] = sqrt(s);
            }
        }
        __syncthreads();
        if (tid == 0) {
            double tn = 0.0;
            for (int q = 0; q < a.n_tensors; ++q) tn += sm[a.o_tn + q];
            if (a.history) {
                double* h = a.history + (prob * cfg.max_iterations + it) * PF_GD_HISTORY_COLS;
                h[0] = (double)(it + 1);
                h[1] = red[2];
                h[2] = red[0];
                h[3] = cfg.n_measured > 0 ? red[1] : 0.0;
                h[4] = red[4];
                h[5] = red[3];
                h[6] = tn;
            }
            int c = 0;
            if (it > 10) {  // solver.py:341-355 (legacy nn_solver_gd.py:171: loss only)
                if (cfg.loss_mode == 0 && red[3] < cfg.tolerance) c = 1;
                else if (!isnan(red[2]) && red[2] < cfg.tolerance) c = 1;
            }
            red[5] = (double)c;
        }
        __syncthreads();
        it_done = it + 1;
        if (red[5] != 0.0) {
            conv = 1;
            break;
        }
    }

    // ---- results: theta, u, reactions f_int - lambda f_ext on fixed DOFs (solver.py:374-380) ----
    nets_forward(a, sm);
    gather_linear(a, sm, u, sm + a.o_f);
    __syncthreads();
    for (int i = tid; i < a.ntheta; i += nt) a.theta[prob * a.ntheta + i] = th[i];
    for (int d = tid; d < a.ndof; d += nt) {
        a.u[prob * a.ndof + d] = u[d];
        if (a.reactions)
            a.reactions[prob * a.ndof + d] = a.dof_free[d] ? 0.0 : __dsub_rn(sm[a.o_f + d], __dmul_rn(lam, sm[a.o_fext + d]));
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
    a.o_delta = take(2 * a.nelem * max_w);
    a.o_red = take(8);
    a.o_tn = take(std::max(ntens, 1));
    for (int k = 0; k < kMaxNets; ++k) {
        NetDev& n = a.nets[k];
        n.val_off = take(2 * a.nelem);
        n.act_off = (n.enabled && n.active) ? take(a.nelem * (n.d.in_dim + n.d.L * n.d.w)) : 0;
    }
    a.smem_doubles = off;
    const size_t smem = (size_t)off * sizeof(double);
    PF_REQUIRE(smem <= 200 * 1024,
               "problem too large for the single-CTA gradient-descent kernel (%zu bytes of shared memory needed)", smem);
    if (smem > 48 * 1024)
        PF_CUDA_CHECK(cudaFuncSetAttribute(gd_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int work = std::max({a.nelem * max_w, a.ndof, max_w * max_w + max_w, 32});
    const int threads = work <= 64 ? 64 : work <= 128 ? 128 : 256;
    gd_solve_kernel<<<(unsigned)nprob, threads, smem, pf_stream_of(stream)>>>(a);
    PF_CUDA_CHECK(cudaGetLastError());
    return PF_OK;
}
