// Patch-staged gather: the B200 fast path of the residual / mat-vec kernels.
//
// ncu on the plain node gather (profiles/r1_gather_v1.md) showed DRAM at 45 %
// while L2->SM traffic sat at the L2 throughput cap: every E/A value crossed
// L2->SM twice (once per end node) and every u value ~7 times.  Here a CTA owns
// one compact patch of <= 32 nodes (recursive coordinate bisection, built in
// the plan) and one chunk of 32 problems.  It brings each row piece it needs
// -- E and A of every element incident to the patch, u of every patch + halo
// node; 256 contiguous bytes each -- into shared memory exactly once with
// cp.async.bulk (UBLKCP, completion on an mbarrier), then gathers from shared
// memory.  Sums still run over each node's incident elements in ascending
// element id, so results are bitwise identical to the generic kernel.
#include <algorithm>
#include <cstdlib>

#include "pf_element.cuh"
#include "pf_internal.h"

namespace {

constexpr int kChunk = 32;          // problems per CTA (one 256-byte row piece per row)
constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;
constexpr int kRowBytes = kChunk * 8;

struct PatchArgs {
    const PfPatch* __restrict__ patches;
    const int32_t* __restrict__ patch_nodes;
    const int32_t* __restrict__ patch_elems;
    const int32_t* __restrict__ patch_inc_ptr;
    const PfPatchInc* __restrict__ patch_inc;
    const double4* __restrict__ patch_inc_geo;
    const uint8_t* __restrict__ dof_free;
    const double* __restrict__ x;   // u (force) or v (mat-vec), [ndof][ldb]
    const double* __restrict__ E;
    const double* __restrict__ A;
    const double* __restrict__ f_ext;
    double* __restrict__ f_out;
    double* __restrict__ r_out;
    double* __restrict__ half_sq_part;  // [npatch][B]
    unsigned long long* __restrict__ max_strain_bits;
    int64_t ldb;          // row stride (total problems)
    int64_t B;            // problems covered by this launch (multiple of kChunk)
    int64_t fext_stride, fext_bmul;
    double load_factor;
    int mat_batched;
    int max_elems, max_local, max_inc;
    int chunks_per_cta;
    int x_rows;  // rows reserved for x: max(max_local * DIM, 16) so the tile can host the block reduction
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 1-D bulk async copy global -> shared, completion counted in bytes on the mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// Persistent over the problem chunks of one patch, two pipeline stages:
// while the CTA gathers chunk c out of stage c&1, the row pieces of chunk c+1
// stream into the other stage.  Index data is staged once per patch (already
// scaled to shared-memory offsets) and every thread keeps the global row
// pointers of "its" copies in registers, so the steady state issues ~3
// instructions per 512 bytes moved and ~20 per element evaluation.
//
// Stage layout (rows of kChunk doubles): [E of the patch's elements][A ...][x of local nodes * DIM]
template <int DIM, int MODE, bool MATB, bool STRAIN, bool BULK>
__global__ void __launch_bounds__(kThreads) patch_gather_kernel(PatchArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int n_mat_cap = MATB ? 2 * a.max_elems : 0;
    const int stage_rows = n_mat_cap + a.x_rows;
    double* stage0 = reinterpret_cast<double*>(smem_raw);
    double4* s_geo = reinterpret_cast<double4*>(stage0 + 2 * (size_t)stage_rows * kChunk);
    double* s_fext = reinterpret_cast<double*>(s_geo + a.max_inc);   // [kPatchNodes*2] load_factor * f_ext (shared loads)
    double* s_red = s_fext + kPatchNodes * 2;                        // [2][kWarps][32]
    int2* s_inc = reinterpret_cast<int2*>(s_red + 2 * kWarps * 32);  // {E row offset, x row offset} in doubles
    int32_t* s_nodes = reinterpret_cast<int32_t*>(s_inc + a.max_inc);
    int32_t* s_elems = s_nodes + a.max_local;
    int32_t* s_ptr = s_elems + a.max_elems;
    int32_t* s_free = s_ptr + kPatchNodes + 1;                       // [kPatchNodes*2]
    uint64_t* bars = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(s_free + kPatchNodes * 2) + 7) & ~uintptr_t(7));

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const PfPatch pt = a.patches[blockIdx.x];
    const int n_mat_rows = MATB ? 2 * pt.n_elem : 0;
    const int n_rows = n_mat_rows + DIM * pt.n_local;
    if (BULK && tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
    }
    for (int i = tid; i < pt.n_local; i += kThreads) s_nodes[i] = a.patch_nodes[pt.node_off + i];
    for (int i = tid; i < pt.n_elem; i += kThreads) s_elems[i] = a.patch_elems[pt.elem_off + i];
    for (int i = tid; i <= pt.n_owned; i += kThreads) s_ptr[i] = a.patch_inc_ptr[pt.ptr_off + i];
    __syncthreads();
    const int n_inc = s_ptr[pt.n_owned];
    for (int i = tid; i < n_inc; i += kThreads) {
        const PfPatchInc pi = a.patch_inc[pt.inc_off + i];
        s_inc[i] = make_int2(MATB ? pi.lelem * kChunk : a.patch_elems[pt.elem_off + pi.lelem],
                             (n_mat_rows + pi.lnbr * DIM) * kChunk);
        s_geo[i] = a.patch_inc_geo[pt.inc_off + i];
    }
    if (MODE == 0 && (a.r_out || a.half_sq_part)) {
        for (int i = tid; i < pt.n_owned * DIM; i += kThreads) {
            const int64_t dof = (int64_t)s_nodes[i / DIM] * DIM + (i % DIM);
            s_free[i] = a.dof_free[dof];
            s_fext[i] = a.fext_bmul ? 0.0 : __dmul_rn(a.load_factor, a.f_ext[dof]);
        }
    }

    // rows this thread copies: r = tid / kSegs + i * kRowsPerPass; 16-byte segment tid % kSegs
    constexpr int kSegs = kRowBytes / 16;
    constexpr int kRowsPerPass = kThreads / kSegs;
    constexpr int kMaxPass = 14;
    auto row_src = [&](int r) -> const double* {
        if (r < n_mat_rows) {
            const int q = r < pt.n_elem ? r : r - pt.n_elem;
            return (r < pt.n_elem ? a.E : a.A) + (int64_t)s_elems[q] * a.ldb;
        }
        const int q = r - n_mat_rows;
        return a.x + ((int64_t)s_nodes[q / DIM] * DIM + (q % DIM)) * a.ldb;
    };
    const int seg = tid % kSegs;
    const char* src[kMaxPass];
#pragma unroll
    for (int i = 0; i < kMaxPass; ++i) {
        const int r = tid / kSegs + i * kRowsPerPass;
        src[i] = r < n_rows ? reinterpret_cast<const char*>(row_src(r)) + seg * 16 : nullptr;
    }
    const uint32_t stage_bytes = (uint32_t)stage_rows * kRowBytes;
    const uint32_t my_dst = smem_u32(stage0) + (tid / kSegs) * kRowBytes + seg * 16;
    auto issue = [&](int chunk, int stage) {
        const int64_t boff = (int64_t)chunk * kRowBytes;
        if (BULK) {
            if (tid == 0) mbar_expect_tx(&bars[stage], (uint32_t)n_rows * kRowBytes);
            for (int r = tid; r < n_rows; r += kThreads)
                bulk_g2s(reinterpret_cast<char*>(stage0) + stage * stage_bytes + r * kRowBytes,
                         reinterpret_cast<const char*>(row_src(r)) + boff, kRowBytes, &bars[stage]);
        } else {
#pragma unroll
            for (int i = 0; i < kMaxPass; ++i)
                if (src[i])
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(my_dst + stage * stage_bytes +
                                                                                  i * kRowsPerPass * kRowBytes),
                                 "l"(src[i] + boff)
                                 : "memory");
            for (int r = tid / kSegs + kMaxPass * kRowsPerPass; r < n_rows; r += kRowsPerPass)  // oversize patches
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(stage0) + stage * stage_bytes +
                                                                              r * kRowBytes + seg * 16),
                             "l"(reinterpret_cast<const char*>(row_src(r)) + boff + seg * 16)
                             : "memory");
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
    };

    const int nchunk = (int)(a.B / kChunk);
    const int c_begin = blockIdx.y * a.chunks_per_cta;
    const int c_end = min(nchunk, c_begin + a.chunks_per_cta);
    const int a_off = pt.n_elem * kChunk;  // A rows follow the E rows
    __syncthreads();  // index staging done (and mbarriers initialised)
    issue(c_begin, 0);
    for (int c = c_begin; c < c_end; ++c) {
        const int stage = (c - c_begin) & 1;
        if (c + 1 < c_end) {
            issue(c + 1, stage ^ 1);
            if (!BULK) asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            if (!BULK) asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        if (BULK) {
            mbar_wait(&bars[stage], ((c - c_begin) >> 1) & 1);
        } else {
            __syncthreads();
        }
        const double* __restrict__ sb = stage0 + (size_t)stage * stage_rows * kChunk + lane;
        const int64_t b = (int64_t)c * kChunk + lane;
        double sq = 0.0, eps_abs = 0.0;
        for (int l = warp; l < pt.n_owned; l += kWarps) {
            const int xoff = (n_mat_rows + l * DIM) * kChunk;
            const double xs = sb[xoff];
            const double ys = DIM == 2 ? sb[xoff + kChunk] : 0.0;
            const int64_t d0 = (int64_t)s_nodes[l] * DIM;
            double fex = 0.0, fey = 0.0;
            if (MODE == 0 && (a.r_out || a.half_sq_part)) {
                if (a.fext_bmul) {  // per-problem loads: fetch early, used after the incidence loop
                    fex = __dmul_rn(a.load_factor, __ldg(a.f_ext + d0 * a.ldb + b));
                    if (DIM == 2) fey = __dmul_rn(a.load_factor, __ldg(a.f_ext + (d0 + 1) * a.ldb + b));
                } else {
                    fex = s_fext[l * DIM];
                    if (DIM == 2) fey = s_fext[l * DIM + 1];
                }
            }
            double fx = 0.0, fy = 0.0;
            const int k1 = s_ptr[l + 1];
#pragma unroll 2
            for (int k = s_ptr[l]; k < k1; ++k) {
                const int2 inc = s_inc[k];
                const double4 geo = s_geo[k];
                double Ee, Ae;
                if (MATB) {
                    Ee = sb[inc.x];
                    Ae = sb[inc.x + a_off];
                } else {
                    Ee = __ldg(a.E + inc.x);
                    Ae = __ldg(a.A + inc.x);
                }
                const double xo = sb[inc.y];
                const double yo = DIM == 2 ? sb[inc.y + kChunk] : 0.0;
                const double eps = pf_linear_incidence<DIM>(Ee, Ae, geo, xs, ys, xo, yo, fx, fy);
                if (STRAIN) eps_abs = fmax(eps_abs, eps);
            }
            if (a.f_out) {
                a.f_out[d0 * a.ldb + b] = fx;
                if (DIM == 2) a.f_out[(d0 + 1) * a.ldb + b] = fy;
            }
            if (MODE == 0 && (a.r_out || a.half_sq_part)) {
                // r = f_int - load_factor * f_ext, product rounded first like the reference (solver.py:267-269)
                const double rx = s_free[l * DIM] ? __dsub_rn(fx, fex) : 0.0;
                const double ry = (DIM == 2 && s_free[l * DIM + 1]) ? __dsub_rn(fy, fey) : 0.0;
                if (a.r_out) {
                    a.r_out[d0 * a.ldb + b] = rx;
                    if (DIM == 2) a.r_out[(d0 + 1) * a.ldb + b] = ry;
                }
                sq += rx * rx;
                sq += ry * ry;
            }
        }
        const bool reduce = MODE == 0 && (a.half_sq_part || (STRAIN && a.max_strain_bits));
        if (reduce) {
            s_red[warp * 32 + lane] = sq;
            s_red[(kWarps + warp) * 32 + lane] = eps_abs;
        }
        __syncthreads();  // stage may be refilled by the next iteration's issue; reduction inputs visible
        if (reduce && warp == 0) {
            double acc = 0.0, m = 0.0;
            for (int w = 0; w < kWarps; ++w) {
                acc += s_red[w * 32 + lane];
                m = fmax(m, s_red[(kWarps + w) * 32 + lane]);
            }
            if (a.half_sq_part) a.half_sq_part[(int64_t)blockIdx.x * a.ldb + b] = acc;
            if (STRAIN && a.max_strain_bits)
                atomicMax(a.max_strain_bits + b, (unsigned long long)__double_as_longlong(m));
        }
    }
}

__global__ void patch_column_sum_kernel(const double* __restrict__ part, int64_t rows, int64_t ld, int64_t ncol,
                                        double scale, double* __restrict__ out) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= ncol) return;
    double acc = 0.0;
    for (int64_t r = 0; r < rows; ++r) acc += part[r * ld + b];
    out[b] = scale * acc;
}

}  // namespace

int pf_patch_gather(pf_plan* plan, const PfGatherCall& c, cudaStream_t st, int64_t* columns_done) {
    *columns_done = 0;
    static const int disabled = getenv("PF_NO_PATCH") ? atoi(getenv("PF_NO_PATCH")) : 0;
    if (disabled || !plan->patch_ok || plan->patches.empty()) return PF_OK;
    const int64_t ncol = (c.B / kChunk) * kChunk;  // full chunks only; the tail goes to the generic kernel
    if (ncol < 4 * kChunk || c.ldb % 2 != 0) return PF_OK;  // too few chunks to pipeline: generic kernel
    const double* x = c.mode == 0 ? c.u : c.v;
    auto aligned16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    if (!aligned16(x) || (c.mat_batched && (!aligned16(c.E) || !aligned16(c.A)))) return PF_OK;

    const int dim = plan->dim;
    const int x_rows = plan->max_patch_local * dim;
    const size_t stage_rows = (size_t)(c.mat_batched ? 2 * plan->max_patch_elems : 0) + x_rows;
    size_t smem = 2 * stage_rows * kRowBytes;
    smem += (size_t)plan->max_patch_inc * (sizeof(double4) + sizeof(int2));
    smem += (size_t)(kPatchNodes * 2 + 2 * kWarps * 32) * sizeof(double);
    smem += (size_t)(plan->max_patch_local + plan->max_patch_elems + kPatchNodes + 1 + kPatchNodes * 2) * sizeof(int32_t);
    smem += 64;  // mbarriers + alignment slack
    if (smem > 225 * 1024) return PF_OK;

    PatchArgs a{};
    a.patches = plan->d_patches;
    a.patch_nodes = plan->d_patch_nodes;
    a.patch_elems = plan->d_patch_elems;
    a.patch_inc_ptr = plan->d_patch_inc_ptr;
    a.patch_inc = plan->d_patch_inc;
    a.patch_inc_geo = plan->d_patch_inc_geo;
    a.dof_free = plan->d_dof_free;
    a.x = x;
    a.E = c.E;
    a.A = c.A;
    a.f_ext = c.f_ext;
    a.f_out = c.f_out;
    a.r_out = c.r_out;
    a.ldb = c.ldb;
    a.B = ncol;
    a.fext_stride = c.fext_batched ? c.ldb : 1;
    a.fext_bmul = c.fext_batched ? 1 : 0;
    a.load_factor = c.load_factor;
    a.mat_batched = c.mat_batched;
    a.max_elems = plan->max_patch_elems;
    a.max_local = plan->max_patch_local;
    a.max_inc = plan->max_patch_inc;
    a.x_rows = x_rows;
    const int nchunk = (int)(ncol / kChunk);
    // split the chunk loop only as far as needed to give every SM a few CTAs
    int ysplit = 1;
    while ((int64_t)plan->patches.size() * ysplit < (int64_t)plan->sm_count * 4 && ysplit * 2 <= nchunk) ysplit *= 2;
    a.chunks_per_cta = (nchunk + ysplit - 1) / ysplit;
    const unsigned npatch = (unsigned)plan->patches.size();
    if (c.half_sq) {
        int rc = pf_plan_reserve_work(plan, (size_t)npatch * c.ldb * sizeof(double));
        if (rc) return rc;
        a.half_sq_part = plan->d_work;
    }
    if (c.max_strain) a.max_strain_bits = reinterpret_cast<unsigned long long*>(c.max_strain);

    dim3 grid(npatch, (unsigned)((nchunk + a.chunks_per_cta - 1) / a.chunks_per_cta), 1);
    static const int use_bulk = getenv("PF_PATCH_TMA") ? atoi(getenv("PF_PATCH_TMA")) : 0;
    const bool strain = c.mode == 0 && c.max_strain != nullptr;
#define PF_PATCH_LAUNCH(D, M, MB, S, T)                                                                            \
    do {                                                                                                           \
        PF_CUDA_CHECK(cudaFuncSetAttribute(patch_gather_kernel<D, M, MB, S, T>,                                    \
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));               \
        patch_gather_kernel<D, M, MB, S, T><<<grid, kThreads, smem, st>>>(a);                                      \
    } while (0)
#define PF_PATCH_SEL_T(D, M, MB, S)                                                    \
    do {                                                                               \
        if (use_bulk) PF_PATCH_LAUNCH(D, M, MB, S, true); else PF_PATCH_LAUNCH(D, M, MB, S, false); \
    } while (0)
#define PF_PATCH_SEL_MB(D, M, S)                                                       \
    do {                                                                               \
        if (c.mat_batched) PF_PATCH_SEL_T(D, M, true, S); else PF_PATCH_SEL_T(D, M, false, S); \
    } while (0)
#define PF_PATCH_SEL_M(D)                                                              \
    do {                                                                               \
        if (c.mode == 1) PF_PATCH_SEL_MB(D, 1, false);                                 \
        else if (strain) PF_PATCH_SEL_MB(D, 0, true);                                  \
        else PF_PATCH_SEL_MB(D, 0, false);                                             \
    } while (0)
    if (dim == 1) PF_PATCH_SEL_M(1); else PF_PATCH_SEL_M(2);
#undef PF_PATCH_SEL_M
#undef PF_PATCH_SEL_MB
#undef PF_PATCH_SEL_T
#undef PF_PATCH_LAUNCH
    PF_CUDA_CHECK(cudaGetLastError());
    if (c.half_sq) {
        patch_column_sum_kernel<<<(unsigned)((ncol + 127) / 128), 128, 0, st>>>(plan->d_work, npatch, c.ldb, ncol, 0.5,
                                                                                c.half_sq);
        PF_CUDA_CHECK(cudaGetLastError());
    }
    *columns_done = ncol;
    return PF_OK;
}
