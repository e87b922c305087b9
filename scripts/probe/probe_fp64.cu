// Micro-probes of the B200 fp64 pipes that decide the structure of the material-network kernels:
//   1. DFMA and DMMA (mma.sync.m8n8k4.f64) issue rates alone and interleaved (do the two pipes overlap?)
//   2. dependent-chain latency of DFMA and DMMA
//   3. throughput of the hidden-layer tanh variants
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../pinn_fem_b200/csrc -I../../include
//        probe_fp64.cu -o probe_fp64.bin
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "pf_mlp.cuh"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// NM DMMA chains and NF DFMA chains per loop iteration
template <int NM, int NF>
__global__ void __launch_bounds__(256) mix_kernel(int iters, double* out) {
    double c[NM > 0 ? NM : 1][2], f[NF > 0 ? NF : 1];
    const double a = 1.0 + 1e-9 * threadIdx.x, b = 1e-9 * (threadIdx.x & 7), x = 1.0 + 1e-9 * threadIdx.x, y = 1e-9;
#pragma unroll
    for (int i = 0; i < NM; ++i) c[i][0] = c[i][1] = i;
#pragma unroll
    for (int i = 0; i < NF; ++i) f[i] = i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < (NM > NF ? NM : NF); ++i) {
            if (i < NM) dmma(c[i][0], c[i][1], a, b);
            if (i < NF) f[i] = fma(f[i], x, y);
        }
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < NM; ++i) s += c[i][0] + c[i][1];
#pragma unroll
    for (int i = 0; i < NF; ++i) s += f[i];
    if (s == 12345.678) out[0] = s;
}

template <int NM, int NF>
double run_mix(int sms, int iters, double* d_out, int warps_per_sm = 64) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int blocks = sms * (warps_per_sm / 8);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        CK(cudaEventRecord(e0));
        mix_kernel<NM, NF><<<blocks, 256>>>(iters, d_out);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    CK(cudaGetLastError());
    const double warp_iters = (double)blocks * 8 * iters;
    const double tf_mma = warp_iters * NM * 512.0 / (best * 1e-3) / 1e12;
    const double tf_fma = warp_iters * NF * 64.0 / (best * 1e-3) / 1e12;
    printf("mix NM=%d NF=%d warps/SM=%d: %.3f ms  DMMA %.2f TF/s + DFMA %.2f TF/s = %.2f TF/s\n", NM, NF, warps_per_sm, best,
           tf_mma, tf_fma, tf_mma + tf_fma);
    return best;
}

// latency: one warp per SM sub-partition, single dependent chain
__global__ void lat_kernel(int kind, int iters, double* out, long long* cyc) {
    double c0 = 1.0, c1 = 2.0, f = 1.0;
    const double a = 1.0 + 1e-9 * threadIdx.x, b = 1e-9;
    long long t0 = clock64();
    if (kind == 0) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) f = fma(f, a, b);
        }
    } else {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) dmma(c0, c1, a, b);
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
    if (f + c0 + c1 == 12345.678) out[0] = f;
}

// table-based exp variant of the hidden-layer tanh: 64-entry 2^(j/64) table in shared memory, degree-5 polynomial
__device__ __forceinline__ double tanh_table(double x, const double* __restrict__ tab) {
    const double ax = fabs(x);
    const double y = ax < 20.0 ? ax + ax : 40.0;
    const double shifter = 6755399441055744.0;
    const double t = fma(y, 1.4426950408889634 * 64.0, shifter);
    const int n = __double2loint(t);
    const double nf = t - shifter;
    double r = fma(nf, -6.93147180369123816490e-01 / 64.0, y);
    r = fma(nf, -1.90821492927058770002e-10 / 64.0, r);
    // exp(r), |r| <= ln2/128: degree 5
    double p = fma(r, 8.333333333333333e-03, 4.1666666666666664e-02);
    p = fma(p, r, 1.6666666666666666e-01);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    const double tj = tab[n & 63];
    p *= tj;
    const double e = __hiloint2double(__double2hiint(p) + ((n >> 6) << 20), __double2loint(p));
    const double d = e + 1.0;
    double q;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(q) : "d"(d));
    q = fma(q, fma(-d, q, 1.0), q);
    q = fma(q, fma(-d, q, 1.0), q);
    const double res = fma(-2.0, q, 1.0);
    return x != x ? x : copysign(res, x);
}

// variant: 1 - 2/(e+1) with a single division-free refinement path: (e-1)/(e+1) via rcp + 2 Newton (same) -- and
// an alternative with Horner instead of Estrin (fewer instructions: no r2/r4/r8)
__device__ __forceinline__ double tanh_horner(double x) {
    const double ax = fabs(x);
    const double y = ax < 20.0 ? ax + ax : 40.0;
    const double shifter = 6755399441055744.0;
    const double t = fma(y, 1.4426950408889634, shifter);
    const int n = __double2loint(t);
    const double nf = t - shifter;
    double r = fma(nf, -6.93147180369123816490e-01, y);
    r = fma(nf, -1.90821492927058770002e-10, r);
    double p = 1.6059043836821613e-10;
    p = fma(p, r, 2.08767569878681e-09);
    p = fma(p, r, 2.505210838544172e-08);
    p = fma(p, r, 2.755731922398589e-07);
    p = fma(p, r, 2.7557319223985893e-06);
    p = fma(p, r, 2.48015873015873e-05);
    p = fma(p, r, 1.984126984126984e-04);
    p = fma(p, r, 1.388888888888889e-03);
    p = fma(p, r, 8.333333333333333e-03);
    p = fma(p, r, 4.1666666666666664e-02);
    p = fma(p, r, 1.6666666666666666e-01);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    const double e = __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));
    const double d = e + 1.0;
    double q;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(q) : "d"(d));
    q = fma(q, fma(-d, q, 1.0), q);
    q = fma(q, fma(-d, q, 1.0), q);
    const double res = fma(-2.0, q, 1.0);
    return x != x ? x : copysign(res, x);
}

template <int VAR, int ILP>
__global__ void __launch_bounds__(256) tanh_kernel(int iters, double* out) {
    __shared__ double tab[64];
    if (threadIdx.x < 64) tab[threadIdx.x] = exp2((double)threadIdx.x / 64.0);
    __syncthreads();
    double v[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) v[i] = 0.1 * i + 1e-3 * threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            const double xin = v[i] * 3.0 - 0.7;
            v[i] = VAR == 0 ? pf_tanh(xin) : VAR == 1 ? tanh_table(xin, tab) : VAR == 2 ? tanh_horner(xin) : tanh(xin);
        }
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += v[i];
    if (s == 12345.678) out[0] = s;
}

template <int VAR, int ILP>
void run_tanh(const char* name, int sms, double* d_out, int warps_per_sm) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int iters = 2000, blocks = sms * (warps_per_sm / 8);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        CK(cudaEventRecord(e0));
        tanh_kernel<VAR, ILP><<<blocks, 256>>>(iters, d_out);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    CK(cudaGetLastError());
    const double n = (double)blocks * 256 * iters * ILP;
    printf("tanh %-8s ILP=%d warps/SM=%2d: %.3f ms  %.1f Gtanh/s  (%.2f SM-cycles per warp-tanh at 1.9 GHz)\n", name, ILP,
           warps_per_sm, best, n / (best * 1e-3) / 1e9, best * 1e-3 * 1.9e9 / (n / 32 / sms));
}

__global__ void tanh_check_kernel(int n, double* err) {
    __shared__ double tab[64];
    if (threadIdx.x < 64) tab[threadIdx.x] = exp2((double)threadIdx.x / 64.0);
    __syncthreads();
    double m1 = 0, m2 = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double x = -22.0 + 44.0 * i / n;
        const double ref = tanh(x);
        m1 = fmax(m1, fabs(tanh_table(x, tab) - ref));
        m2 = fmax(m2, fabs(tanh_horner(x) - ref));
    }
    atomicMax((unsigned long long*)&err[0], (unsigned long long)__double_as_longlong(m1));
    atomicMax((unsigned long long*)&err[1], (unsigned long long)__double_as_longlong(m2));
}

int main() {
    int dev = 0, sms = 0, clk = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev));
    printf("SMs %d, clock %d kHz\n", sms, clk);
    double* d_out; long long* d_cyc;
    CK(cudaMalloc(&d_out, 64)); CK(cudaMalloc(&d_cyc, 64));
    CK(cudaMemset(d_out, 0, 64));
    const int iters = 4000;
    run_mix<8, 0>(sms, iters, d_out);
    run_mix<0, 8>(sms, iters * 4, d_out);
    run_mix<8, 8>(sms, iters, d_out);     // DMMA pipe time 8*16, DFMA 8*2 per SMSP-warp
    run_mix<8, 16>(sms, iters, d_out);
    run_mix<4, 16>(sms, iters, d_out);
    run_mix<4, 32>(sms, iters, d_out);    // equal pipe time if DMMA = 16 clk, DFMA = 2 clk
    run_mix<2, 16>(sms, iters, d_out);
    run_mix<2, 32>(sms, iters, d_out);
    run_mix<1, 32>(sms, iters, d_out);
    for (int w : {8, 16, 32}) { run_mix<8, 0>(sms, iters, d_out, w); run_mix<0, 8>(sms, iters * 4, d_out, w); run_mix<2, 16>(sms, iters, d_out, w); }
    for (int kind = 0; kind < 2; ++kind) {
        lat_kernel<<<1, 32>>>(kind, 1000, d_out, d_cyc);
        long long c; CK(cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost));
        printf("latency %s: %.2f cycles per dependent op (1 warp)\n", kind ? "DMMA" : "DFMA", (double)c / 16000.0);
        lat_kernel<<<sms, 128>>>(kind, 1000, d_out, d_cyc);
        CK(cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost));
        printf("latency %s: %.2f cycles per dependent op (1 warp per SMSP, all SMs)\n", kind ? "DMMA" : "DFMA", (double)c / 16000.0);
    }
    for (int w : {8, 16, 32}) {
        run_tanh<0, 1>("estrin", sms, d_out, w);
        run_tanh<0, 4>("estrin", sms, d_out, w);
        run_tanh<1, 1>("table", sms, d_out, w);
        run_tanh<1, 4>("table", sms, d_out, w);
        run_tanh<2, 4>("horner", sms, d_out, w);
        run_tanh<3, 4>("cuda", sms, d_out, w);
    }
    double* d_err; CK(cudaMalloc(&d_err, 16)); CK(cudaMemset(d_err, 0, 16));
    tanh_check_kernel<<<sms * 4, 256>>>(1 << 24, d_err);
    double h[2]; CK(cudaMemcpy(h, d_err, 16, cudaMemcpyDeviceToHost));
    printf("max abs err vs CUDA tanh: table %.3e, horner %.3e\n", h[0], h[1]);
    return 0;
}
