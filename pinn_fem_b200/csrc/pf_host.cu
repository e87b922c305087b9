// Host-buffer (end-to-end) entry point: streams B problems held in pinned host
// memory through the GPU in batch chunks, overlapping the H2D copies, the
// residual kernel and the D2H copy on three rotating streams.
#include <algorithm>

#include "pf_internal.h"

static int ensure_stage(pf_plan* p, int which, size_t elems) {
    if (p->stage_elems[which] >= elems) return PF_OK;
    for (int s = 0; s < 3; ++s) {
        if (p->d_stage[s][which]) PF_CUDA_CHECK(cudaFree(p->d_stage[s][which]));
        p->d_stage[s][which] = nullptr;
    }
    p->stage_elems[which] = 0;
    for (int s = 0; s < 3; ++s) PF_CUDA_CHECK(cudaMalloc((void**)&p->d_stage[s][which], elems * sizeof(double)));
    p->stage_elems[which] = elems;
    return PF_OK;
}

extern "C" int pf_residual_host(pf_plan* plan, int kind, int64_t B, const double* u_host, const double* E_host,
                                const double* A_host, const double* f_ext_host, double load_factor,
                                double* r_host, int64_t chunk) {
    int rc = pf_plan_activate(plan);
    if (rc) return rc;
    PF_REQUIRE(u_host && E_host && A_host && f_ext_host && r_host, "pf_residual_host: NULL argument");
    PF_REQUIRE(B >= 1, "B must be >= 1");
    if (chunk <= 0) chunk = 64;  // 512-byte row pieces: best H2D || D2H overlap measured (52 GB/s in, 13 GB/s out)
    chunk = std::min(chunk, B);
    const int64_t ndof = plan->ndof, nelem = plan->nelem;

    for (int s = 0; s < 3; ++s)
        if (!plan->host_streams[s]) PF_CUDA_CHECK(cudaStreamCreateWithFlags(&plan->host_streams[s], cudaStreamNonBlocking));
    // stage buffers: 0 = u, 1 = E, 2 = A, 3 = r  (per pipeline slot)
    if ((rc = ensure_stage(plan, 0, (size_t)ndof * chunk))) return rc;
    if ((rc = ensure_stage(plan, 1, (size_t)std::max<int64_t>(nelem, 1) * chunk))) return rc;
    if ((rc = ensure_stage(plan, 2, (size_t)std::max<int64_t>(nelem, 1) * chunk))) return rc;
    if ((rc = ensure_stage(plan, 3, (size_t)ndof * chunk))) return rc;
    // shared loads live in the plan workspace
    if ((rc = pf_plan_reserve_work(plan, (size_t)ndof * sizeof(double)))) return rc;
    double* d_fext = plan->d_work;
    PF_CUDA_CHECK(cudaMemcpy(d_fext, f_ext_host, ndof * sizeof(double), cudaMemcpyHostToDevice));

    int slot = 0;
    for (int64_t b0 = 0; b0 < B; b0 += chunk, slot = (slot + 1) % 3) {
        const int64_t cw = std::min(chunk, B - b0);
        cudaStream_t st = plan->host_streams[slot];
        double* du = plan->d_stage[slot][0];
        double* dE = plan->d_stage[slot][1];
        double* dA = plan->d_stage[slot][2];
        double* dr = plan->d_stage[slot][3];
        const size_t w = (size_t)cw * sizeof(double), hp = (size_t)B * sizeof(double);
        PF_CUDA_CHECK(cudaMemcpy2DAsync(du, w, u_host + b0, hp, w, ndof, cudaMemcpyHostToDevice, st));
        if (nelem) {
            PF_CUDA_CHECK(cudaMemcpy2DAsync(dE, w, E_host + b0, hp, w, nelem, cudaMemcpyHostToDevice, st));
            PF_CUDA_CHECK(cudaMemcpy2DAsync(dA, w, A_host + b0, hp, w, nelem, cudaMemcpyHostToDevice, st));
        }
        rc = pf_residual(plan, kind, cw, du, dE, dA, 1, nullptr, d_fext, 0, load_factor, dr, nullptr, nullptr, st);
        if (rc) return rc;
        PF_CUDA_CHECK(cudaMemcpy2DAsync(r_host + b0, hp, dr, w, w, ndof, cudaMemcpyDeviceToHost, st));
    }
    for (int s = 0; s < 3; ++s) PF_CUDA_CHECK(cudaStreamSynchronize(plan->host_streams[s]));
    return PF_OK;
}
