// Element arithmetic shared by every kernel that evaluates the linear bar, with
// explicit rounding intrinsics so that all code paths (generic gather, patch
// gather, GD loop) produce the same bits for the same inputs.
#pragma once

// Row of the owning node of fe = ke @ u_e (fem/element.py:72-100):
//   k = (E*A)/l0, axial = c (xs - xo) + s (ys - yo), f += k * axial * (c, s)
// geo = {cos, sin, 1/l0, l0}.  Returns |strain| of the element.
template <int DIM>
__device__ __forceinline__ double pf_linear_incidence(double Ee, double Ae, const double4& geo, double xs, double ys,
                                                      double xo, double yo, double& fx, double& fy) {
    const double k = __dmul_rn(__dmul_rn(Ee, Ae), geo.z);
    if (DIM == 2) {
        const double axial = __fma_rn(geo.x, __dsub_rn(xs, xo), __dmul_rn(geo.y, __dsub_rn(ys, yo)));
        const double t = __dmul_rn(k, axial);
        fx = __fma_rn(t, geo.x, fx);
        fy = __fma_rn(t, geo.y, fy);
        return fabs(__dmul_rn(axial, geo.z));
    }
    const double du = __dsub_rn(xs, xo);
    fx = __fma_rn(k, du, fx);
    return fabs(__dmul_rn(du, geo.z));
}
