// Shared definitions of the material-network kernels.
#pragma once
#include <stdint.h>

#define PF_MLP_MAX_LAYERS 8

// Flat parameter layout of SimpleNN in nn.Module.parameters() order:
// W_0[w][in], b_0[w], W_1[w][w], b_1[w], ..., W_out[1][w], b_out[1].
struct PfMlpDesc {
    int in_dim, L, w, wp;  // wp = w rounded up to a multiple of 4
    int n_params;
    int w_off[PF_MLP_MAX_LAYERS + 1];
    int b_off[PF_MLP_MAX_LAYERS + 1];
};

int pf_mlp_make_desc(int input_dim, int hidden_layers, int width, PfMlpDesc* d);

// softplus(z) * scale with torch's defaults (beta 1, threshold 20), fem/properties.py:153-156
__host__ __device__ __forceinline__ double pf_mlp_output(double z, double scale, int positive) {
    if (!positive) return z * scale;
    const double sp = z > 20.0 ? z : log1p(exp(z));
    return sp * scale;
}

// d(value)/dz
__host__ __device__ __forceinline__ double pf_mlp_output_grad(double z, double scale, int positive) {
    if (!positive || z > 20.0) return scale;
    return scale / (1.0 + exp(-z));
}

#ifdef __CUDACC__
// tanh for the hidden layers.  CUDA's tanh() costs ~120 SASS instructions (two branches, generic division);
// ncu put 62 % of the MLP backward kernel's samples in it.  This one is branch-free and ~35 instructions:
//   tanh(x) = sign(x) (1 - 2 / (exp(2|x|) + 1)),
// exp by range reduction (y = n ln2 + r, |r| <= ln2/2) and a degree-13 Taylor polynomial (remainder < 4e-18, Estrin),
// the reciprocal from the hardware approximation plus two Newton steps.  Absolute error 3.3e-16 measured
// (3 ulp at 1); the relative error of results much smaller than 1 is larger (1 - 2r cancels), which is
// irrelevant here: activations feed dense layers with O(1) weights.  NaN propagates.
__device__ __forceinline__ double pf_tanh(double x) {
    const double ax = fabs(x);
    const double y = ax < 20.0 ? ax + ax : 40.0;  // tanh(20) rounds to 1
    const double shifter = 6755399441055744.0;    // 1.5 * 2^52: rint(t) lands in the low word
    const double t = fma(y, 1.4426950408889634, shifter);
    const int n = __double2loint(t);
    const double nf = t - shifter;
    double r = fma(nf, -6.93147180369123816490e-01, y);
    r = fma(nf, -1.90821492927058770002e-10, r);
    // degree-13 Taylor polynomial of exp(r) in Estrin form: dependency depth 4 instead of 13 (the kernels that
    // call this are bound by fp64 dependency chains, not by instruction count)
    const double r2 = r * r, r4 = r2 * r2, r8 = r4 * r4;
    const double a0 = fma(1.0, r, 1.0);
    const double a1 = fma(1.6666666666666666e-01, r, 0.5);
    const double a2 = fma(8.333333333333333e-03, r, 4.1666666666666664e-02);
    const double a3 = fma(1.984126984126984e-04, r, 1.388888888888889e-03);
    const double a4 = fma(2.7557319223985893e-06, r, 2.48015873015873e-05);
    const double a5 = fma(2.505210838544172e-08, r, 2.755731922398589e-07);
    const double a6 = fma(1.6059043836821613e-10, r, 2.08767569878681e-09);
    const double b0 = fma(a1, r2, a0), b1 = fma(a3, r2, a2), b2 = fma(a5, r2, a4);
    const double c0 = fma(b1, r4, b0), c1 = fma(a6, r4, b2);
    const double p = fma(c1, r8, c0);
    const double e = __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));  // p * 2^n, n <= 58
    const double d = e + 1.0;
    double q;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(q) : "d"(d));
    q = fma(q, fma(-d, q, 1.0), q);  // two Newton steps: 2^-20 -> 2^-40 -> full precision
    q = fma(q, fma(-d, q, 1.0), q);
    const double res = fma(-2.0, q, 1.0);
    return x != x ? x : copysign(res, x);
}
#endif
