// Entry points that are declared in pinnfem.h but not implemented yet.  They
// fail loudly; nothing here computes on the CPU.
#include "pf_internal.h"

#define PF_TODO(name)                               \
    pf_set_error(name ": not implemented yet");     \
    return PF_ERR_ARG

extern "C" int pf_solve_dense(int64_t, int64_t, double*, double*, int32_t*, void*) { PF_TODO("pf_solve_dense"); }
extern "C" int pf_cg_solve(pf_plan*, int, int64_t, const double*, const double*, const double*, int, const double*,
                           double*, double, int, double*, int64_t, int32_t*, double*, void*) { PF_TODO("pf_cg_solve"); }
extern "C" int64_t pf_cg_work_len(const pf_plan*, int64_t) { return -1; }
extern "C" int pf_gn_normal_equations(int64_t, int64_t, const double*, const double*, double, double*, double*,
                                      double*, void*) { PF_TODO("pf_gn_normal_equations"); }
