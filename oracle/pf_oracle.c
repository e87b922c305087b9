/* pf_oracle.c -- CPU restatement of the reference's element loop in plain C.
 * TEST INFRASTRUCTURE ONLY: used by tests/ (as a second checker) and by
 * bench.py's cpu_baseline / --impl reference legs.  The product never links it.
 *
 * It follows the reference's scatter form literally (fem/assembly.py:52-73,
 * fem/nn_assembly.py:183-229): loop over elements in order, evaluate
 * k = (E*A)/l0 and fe = ke @ u_e (fem/element.py:59-100), then
 * f_int[dofs] += fe.  Problems are independent, so OpenMP parallelises over
 * blocks of problems; within a problem the summation order is the reference's.
 * Pinned against the golden vectors through tests/test_oracle_golden.py
 * (test_c_oracle_matches_numpy_oracle).
 *
 * Layout: batched arrays are [row][B] (problem index last), like the CUDA ABI.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define PB 8 /* problems per inner block (one cache line of doubles) */

int pfo_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* r = f_int - lam * f_ext on free dofs, 0 on fixed dofs (fem/solver.py:267-269);
 * f_int may be NULL, r may be NULL.  kind: 0 linear, 1 Green-Lagrange (2-D). */
int pfo_residual(int dim, int kind, int64_t nnode, int64_t nelem, const int64_t* elements, const double* nodes,
                 int64_t B, const double* u, const double* E, const double* A, int mat_batched,
                 const double* f_ext, double lam, const uint8_t* dof_free, double* f_int, double* r,
                 int nthreads) {
    const int64_t ndof = nnode * dim;
    const int64_t nblk = (B + PB - 1) / PB;
    if (nthreads < 1) nthreads = 1;
    int fail = 0;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
    for (int64_t blk = 0; blk < nblk; ++blk) {
        const int64_t b0 = blk * PB;
        const int nb = (int)((B - b0) < PB ? (B - b0) : PB);
        double* acc = (double*)calloc((size_t)ndof * PB, sizeof(double));
        if (!acc) {
            fail = 1;
            continue;
        }
        for (int64_t e = 0; e < nelem; ++e) {
            const int64_t i = elements[2 * e], j = elements[2 * e + 1];
            if (dim == 1) {
                const double l0 = fabs(nodes[j] - nodes[i]);
                for (int q = 0; q < nb; ++q) {
                    const int64_t b = b0 + q;
                    const double Ee = mat_batched ? E[e * B + b] : E[e], Ae = mat_batched ? A[e * B + b] : A[e];
                    const double k = (Ee * Ae) / l0;
                    const double ui = u[i * B + b], uj = u[j * B + b];
                    acc[i * PB + q] += k * (ui - uj);
                    acc[j * PB + q] += k * (uj - ui);
                }
                continue;
            }
            const double dx0 = nodes[2 * j] - nodes[2 * i], dy0 = nodes[2 * j + 1] - nodes[2 * i + 1];
            const double l0 = sqrt(dx0 * dx0 + dy0 * dy0);
            const double cx = dx0 / l0, cy = dy0 / l0;
            const double c2 = cx * cx, s2 = cy * cy, cs = cx * cy;
            for (int q = 0; q < nb; ++q) {
                const int64_t b = b0 + q;
                const double Ee = mat_batched ? E[e * B + b] : E[e], Ae = mat_batched ? A[e * B + b] : A[e];
                const double uix = u[(2 * i) * B + b], uiy = u[(2 * i + 1) * B + b];
                const double ujx = u[(2 * j) * B + b], ujy = u[(2 * j + 1) * B + b];
                double f0, f1, f2, f3;
                if (kind == 0) {
                    const double k = (Ee * Ae) / l0;
                    /* fe = ke @ u_e, ke = k * pattern (fem/element.py:80-100) */
                    f0 = (k * c2) * uix + (k * cs) * uiy + (k * -c2) * ujx + (k * -cs) * ujy;
                    f1 = (k * cs) * uix + (k * s2) * uiy + (k * -cs) * ujx + (k * -s2) * ujy;
                    f2 = (k * -c2) * uix + (k * -cs) * uiy + (k * c2) * ujx + (k * cs) * ujy;
                    f3 = (k * -cs) * uix + (k * -s2) * uiy + (k * cs) * ujx + (k * s2) * ujy;
                } else {
                    /* fem/element.py:119-131 */
                    const double dx = (nodes[2 * j] + ujx) - (nodes[2 * i] + uix);
                    const double dy = (nodes[2 * j + 1] + ujy) - (nodes[2 * i + 1] + uiy);
                    const double l = sqrt(dx * dx + dy * dy);
                    const double eg = (l * l - l0 * l0) / (2.0 * l0 * l0);
                    const double n = (Ee * Ae / l0) * eg;
                    f0 = n * dx;
                    f1 = n * dy;
                    f2 = n * -dx;
                    f3 = n * -dy;
                }
                acc[(2 * i) * PB + q] += f0;
                acc[(2 * i + 1) * PB + q] += f1;
                acc[(2 * j) * PB + q] += f2;
                acc[(2 * j + 1) * PB + q] += f3;
            }
        }
        for (int64_t d = 0; d < ndof; ++d)
            for (int q = 0; q < nb; ++q) {
                const int64_t b = b0 + q;
                const double f = acc[d * PB + q];
                if (f_int) f_int[d * B + b] = f;
                if (r) r[d * B + b] = dof_free[d] ? f - lam * f_ext[d] : 0.0;
            }
        free(acc);
    }
    return fail;
}
