// fp64 peak probes for the rooflines of the compute-bound kernels (MEASURED_PEAKS.json carries HBM and bf16
// only; SURVEY.md 8d asks for the fp64 figure to be measured on the box).  Register-resident loops with
// enough independent chains to hide the pipe latency, timed with CUDA events.
#include "pf_internal.h"

namespace {

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(256) dfma_probe_kernel(int iters, double* out) {
    double a[8], x = 1.0 + 1e-9 * threadIdx.x, y = 1e-9;
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fma(a[i], x, y);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    if (s == 12345.678) out[0] = s;  // keep the chains alive
}

__global__ void __launch_bounds__(256) dmma_probe_kernel(int iters, double* out) {
    double c[8][2], a = 1.0 + 1e-9 * threadIdx.x, b = 1e-9 * (threadIdx.x & 7);
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) dmma(c[i][0], c[i][1], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    if (s == 12345.678) out[0] = s;
}

}  // namespace

// kind 0: DFMA (fp64 FMA pipe), 1: DMMA (fp64 tensor pipe, mma.sync.m8n8k4).  tflops_out: host double.
extern "C" int pf_measure_fp64_peak(int kind, double* tflops_out) {
    PF_REQUIRE(tflops_out && (kind == 0 || kind == 1), "pf_measure_fp64_peak: bad argument");
    int dev = 0, sms = 0;
    PF_CUDA_CHECK(cudaGetDevice(&dev));
    PF_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    double* d_out = nullptr;
    PF_CUDA_CHECK(cudaMalloc((void**)&d_out, sizeof(double)));
    cudaEvent_t e0, e1;
    PF_CUDA_CHECK(cudaEventCreate(&e0));
    PF_CUDA_CHECK(cudaEventCreate(&e1));
    const int iters = 20000, blocks = sms * 8, threads = 256;
    double best_ms = 1e30;
    for (int rep = 0; rep < 4; ++rep) {
        PF_CUDA_CHECK(cudaEventRecord(e0));
        if (kind == 0)
            dfma_probe_kernel<<<blocks, threads>>>(iters, d_out);
        else
            dmma_probe_kernel<<<blocks, threads>>>(iters, d_out);
        PF_CUDA_CHECK(cudaEventRecord(e1));
        PF_CUDA_CHECK(cudaEventSynchronize(e1));
        float ms = 0.f;
        PF_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best_ms) best_ms = ms;
    }
    PF_CUDA_CHECK(cudaGetLastError());
    // flop per thread-iteration: 8 FMAs = 16 flop; per warp-iteration of DMMA: 8 x (8*8*4*2) = 4096 flop
    const double total = kind == 0 ? (double)blocks * threads * iters * 16.0
                                   : (double)blocks * (threads / 32) * iters * 4096.0;
    *tflops_out = total / (best_ms * 1e-3) / 1e12;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_out);
    return PF_OK;
}
