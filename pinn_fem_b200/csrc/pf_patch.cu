// Patch-staged gather: the B200 fast path of the residual / mat-vec kernels.
//
// ncu on the plain node gather (profiles/r1_residual_kernel_evolution.md) showed DRAM at 45 %
// while L2->SM traffic sat at the L2 throughput cap: every E/A value crossed
// L2->SM twice (once per end node) and every u value ~7 times.  Here a CTA owns
// one compact patch of <= 32 nodes (recursive coordinate bisection, built in
// the plan) and is persistent over chunks of CH problems.  It brings each row
// piece it needs -- E and A of every element incident to the patch, u of every
// patch + halo node; CH*8 contiguous bytes each -- into shared memory exactly
// once with cp.async (two stages: chunk c+1 streams in while chunk c is
// gathered), then gathers from shared memory.  Sums still run over each node's
// incident elements in ascending element id, so results are bitwise identical
// to the generic kernel.
//
// CH = 32: 512 threads, one CTA per SM (2 x 85 KB stages for a 32-node lattice patch).
// CH = 16: 256 threads, two CTAs per SM: the index staging, barriers and copy waits of one CTA
//          overlap the gather of the other.
#include <algorithm>
#include <cstdlib>

#include "pf_element.cuh"
#include "pf_internal.h"

namespace {

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
    int64_t B;            // problems covered by this launch (multiple of CH)
    int64_t fext_bmul;    // 1: f_ext is [ndof][ldb], 0: shared [ndof]
    double load_factor;
    int max_elems, max_local, max_inc;
    int chunks_per_cta;
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

template <int CH>
struct PatchCfg {
    static constexpr int kThreads = CH * 16;       // 512 / 256
    static constexpr int kWarps = kThreads / 32;
    static constexpr int kRowBytes = CH * 8;
    static constexpr int kSegs = kRowBytes / 16;   // 16-byte segments per row piece
    static constexpr int kRowsPerPass = kThreads / kSegs;
    static constexpr int kLanesPerNode = CH / 2;   // a lane carries two adjacent problems
    static constexpr int kNodesPerWarp = 32 / kLanesPerNode;
    static constexpr int kMinCtas = CH == 32 ? 1 : 2;
};

// shared-memory footprint of one CTA (the kernel carves it in this order)
template <int CH>
size_t patch_smem_bytes(int stage_rows, int max_inc, int max_local, int max_elems) {
    using C = PatchCfg<CH>;
    size_t smem = 2 * (size_t)stage_rows * C::kRowBytes;                       // two stages
    smem += (size_t)max_inc * (sizeof(double4) + sizeof(int2));               // incidence geometry + offsets
    smem += (size_t)(kPatchNodes * 2 + 2 * C::kWarps * CH) * sizeof(double);  // loads, block reduction
    smem += (size_t)(max_local + max_elems + kPatchNodes + 1 + kPatchNodes * 2) * sizeof(int32_t);
    smem += 64;                                                               // mbarriers + alignment slack
    smem += (size_t)stage_rows * sizeof(void*) + kPatchNodes * sizeof(int64_t);  // row sources, output offsets
    return smem;
}

// Stage layout (rows of CH doubles): [E of the patch's elements][A ...][x of local nodes * DIM]
template <int DIM, int MODE, bool MATB, bool STRAIN, bool BULK, int CH>
__global__ void __launch_bounds__(PatchCfg<CH>::kThreads, PatchCfg<CH>::kMinCtas) patch_gather_kernel(PatchArgs a) {
    using C = PatchCfg<CH>;
    constexpr int kThreads = C::kThreads, kWarps = C::kWarps, kRowBytes = C::kRowBytes;
    constexpr int kSegs = C::kSegs, kRowsPerPass = C::kRowsPerPass;
    constexpr int LPN = C::kLanesPerNode, NPW = C::kNodesPerWarp;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int stage_rows = (MATB ? 2 * a.max_elems : 0) + a.max_local * DIM;
    double* stage0 = reinterpret_cast<double*>(smem_raw);
    double4* s_geo = reinterpret_cast<double4*>(stage0 + 2 * (size_t)stage_rows * CH);
    double* s_fext = reinterpret_cast<double*>(s_geo + a.max_inc);   // [kPatchNodes*2] load_factor * f_ext (shared loads)
    double* s_red = s_fext + kPatchNodes * 2;                        // [2][kWarps][CH]
    int2* s_inc = reinterpret_cast<int2*>(s_red + 2 * kWarps * CH);  // {E row, x row} offsets in double2 units
    int32_t* s_nodes = reinterpret_cast<int32_t*>(s_inc + a.max_inc);
    int32_t* s_elems = s_nodes + a.max_local;
    int32_t* s_ptr = s_elems + a.max_elems;
    int32_t* s_free = s_ptr + kPatchNodes + 1;                       // [kPatchNodes*2]
    // aligned by OFFSET from the shared array, not through uintptr_t: an integer round trip makes the compiler forget
    // the address space and every access below becomes a generic LD/ST (long-scoreboard stalls in the copy loop)
    const size_t bars_off =
        ((size_t)(reinterpret_cast<unsigned char*>(s_free + kPatchNodes * 2) - smem_raw) + 7) & ~size_t(7);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + bars_off);
    const char** s_rowsrc = reinterpret_cast<const char**>(bars + 2);       // [stage_rows] global address of each staged row
    int64_t* s_outoff = reinterpret_cast<int64_t*>(s_rowsrc + stage_rows);  // [kPatchNodes] first DOF of the node * ldb

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
        s_inc[i] = make_int2(MATB ? pi.lelem * (CH / 2) : s_elems[pi.lelem], (n_mat_rows + pi.lnbr * DIM) * (CH / 2));
        s_geo[i] = a.patch_inc_geo[pt.inc_off + i];
    }
    if (MODE == 0 && (a.r_out || a.half_sq_part)) {
        for (int i = tid; i < pt.n_owned * DIM; i += kThreads) {
            const int64_t dof = (int64_t)s_nodes[i / DIM] * DIM + (i % DIM);
            s_free[i] = a.dof_free[dof];
            s_fext[i] = a.fext_bmul ? 0.0 : __dmul_rn(a.load_factor, a.f_ext[dof]);
        }
    }
    // The global address of every staged row sits in shared memory (not in registers: the gather
    // needs them), so a copy is LDS.64 + add + LDGSTS.  Thread t copies 16-byte segment t % kSegs of
    // rows t / kSegs, t / kSegs + kRowsPerPass, ...
    for (int r = tid; r < n_rows; r += kThreads) {
        const double* src;
        if (r < n_mat_rows) {
            const int q = r < pt.n_elem ? r : r - pt.n_elem;
            src = (r < pt.n_elem ? a.E : a.A) + (int64_t)s_elems[q] * a.ldb;
        } else {
            const int q = r - n_mat_rows;
            src = a.x + ((int64_t)s_nodes[q / DIM] * DIM + (q % DIM)) * a.ldb;
        }
        s_rowsrc[r] = reinterpret_cast<const char*>(src);
    }
    for (int l = tid; l < pt.n_owned; l += kThreads) s_outoff[l] = (int64_t)s_nodes[l] * DIM * a.ldb;
    const int seg = tid % kSegs;
    const uint32_t stage_bytes = (uint32_t)stage_rows * kRowBytes;
    const uint32_t my_dst = smem_u32(stage0) + (tid / kSegs) * kRowBytes + seg * 16;
    auto issue = [&](int chunk, int stage) {
        if (BULK) {
            if (tid == 0) mbar_expect_tx(&bars[stage], (uint32_t)n_rows * kRowBytes);
            for (int r = tid; r < n_rows; r += kThreads)
                bulk_g2s(reinterpret_cast<char*>(stage0) + stage * stage_bytes + r * kRowBytes,
                         s_rowsrc[r] + (int64_t)chunk * kRowBytes, kRowBytes, &bars[stage]);
        } else {
            const int64_t boff = (int64_t)chunk * kRowBytes + seg * 16;
            uint32_t dst = my_dst + stage * stage_bytes;
#pragma unroll 4
            for (int r = tid / kSegs; r < n_rows; r += kRowsPerPass, dst += kRowsPerPass * kRowBytes)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(s_rowsrc[r] + boff) : "memory");
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
    };

    const int nchunk = (int)(a.B / CH);
    const int c_begin = blockIdx.y * a.chunks_per_cta;
    const int c_end = min(nchunk, c_begin + a.chunks_per_cta);
    const int a_off = pt.n_elem * (CH / 2);  // A rows follow the E rows (double2 units)
    // Lane = (node of the warp's group, problem pair): LPN lanes share a node, each carries two adjacent
    // problems through 128-bit shared-memory loads -- half the load and index instructions per element
    // evaluation and two independent dependency chains per incidence.
    const int sub = lane / LPN, pj = lane % LPN;
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
        const double2* __restrict__ sb = reinterpret_cast<const double2*>(stage0 + (size_t)stage * stage_rows * CH) + pj;
        const int64_t b = (int64_t)c * CH + 2 * pj;
        double2 sq = make_double2(0.0, 0.0), eps_abs = make_double2(0.0, 0.0);
        for (int l = NPW * warp + sub; l < pt.n_owned; l += NPW * kWarps) {
            const int xoff = (n_mat_rows + l * DIM) * (CH / 2);
            const double2 xs = sb[xoff];
            const double2 ys = DIM == 2 ? sb[xoff + CH / 2] : make_double2(0.0, 0.0);
            const int64_t o0 = s_outoff[l] + b;
            double2 fex = make_double2(0.0, 0.0), fey = make_double2(0.0, 0.0);
            if (MODE == 0 && (a.r_out || a.half_sq_part)) {
                if (a.fext_bmul) {  // per-problem loads: fetch early, used after the incidence loop
                    fex = __ldg(reinterpret_cast<const double2*>(a.f_ext + o0));
                    fex = make_double2(__dmul_rn(a.load_factor, fex.x), __dmul_rn(a.load_factor, fex.y));
                    if (DIM == 2) {
                        fey = __ldg(reinterpret_cast<const double2*>(a.f_ext + o0 + a.ldb));
                        fey = make_double2(__dmul_rn(a.load_factor, fey.x), __dmul_rn(a.load_factor, fey.y));
                    }
                } else {
                    fex.x = fex.y = s_fext[l * DIM];
                    if (DIM == 2) fey.x = fey.y = s_fext[l * DIM + 1];
                }
            }
            double2 fx = make_double2(0.0, 0.0), fy = make_double2(0.0, 0.0);
            const int k1 = s_ptr[l + 1];
#pragma unroll 2
            for (int k = s_ptr[l]; k < k1; ++k) {
                const int2 inc = s_inc[k];
                const double4 geo = s_geo[k];
                double2 Ee, Ae;
                if (MATB) {
                    Ee = sb[inc.x];
                    Ae = sb[inc.x + a_off];
                } else {
                    Ee.x = Ee.y = __ldg(a.E + inc.x);
                    Ae.x = Ae.y = __ldg(a.A + inc.x);
                }
                const double2 xo = sb[inc.y];
                const double2 yo = DIM == 2 ? sb[inc.y + CH / 2] : make_double2(0.0, 0.0);
                const double e0 = pf_linear_incidence<DIM>(Ee.x, Ae.x, geo, xs.x, ys.x, xo.x, yo.x, fx.x, fy.x);
                const double e1 = pf_linear_incidence<DIM>(Ee.y, Ae.y, geo, xs.y, ys.y, xo.y, yo.y, fx.y, fy.y);
                if (STRAIN) {
                    eps_abs.x = fmax(eps_abs.x, e0);
                    eps_abs.y = fmax(eps_abs.y, e1);
                }
            }
            if (a.f_out) {
                *reinterpret_cast<double2*>(a.f_out + o0) = fx;
                if (DIM == 2) *reinterpret_cast<double2*>(a.f_out + o0 + a.ldb) = fy;
            }
            if (MODE == 0 && (a.r_out || a.half_sq_part)) {
                // r = f_int - load_factor * f_ext, product rounded first like the reference (solver.py:267-269)
                double2 rx = make_double2(0.0, 0.0), ry = make_double2(0.0, 0.0);
                if (s_free[l * DIM]) rx = make_double2(__dsub_rn(fx.x, fex.x), __dsub_rn(fx.y, fex.y));
                if (DIM == 2 && s_free[l * DIM + 1]) ry = make_double2(__dsub_rn(fy.x, fey.x), __dsub_rn(fy.y, fey.y));
                if (a.r_out) {
                    *reinterpret_cast<double2*>(a.r_out + o0) = rx;
                    if (DIM == 2) *reinterpret_cast<double2*>(a.r_out + o0 + a.ldb) = ry;
                }
                sq.x += rx.x * rx.x;
                sq.y += rx.y * rx.y;
                sq.x += ry.x * ry.x;
                sq.y += ry.y * ry.y;
            }
        }
        const bool reduce = MODE == 0 && (a.half_sq_part || (STRAIN && a.max_strain_bits));
        if (reduce) {  // fold the nodes of the warp (fixed order), then one double2 per problem pair
#pragma unroll
            for (int o = LPN; o < 32; o <<= 1) {
                sq.x += __shfl_xor_sync(0xffffffffu, sq.x, o);
                sq.y += __shfl_xor_sync(0xffffffffu, sq.y, o);
                eps_abs.x = fmax(eps_abs.x, __shfl_xor_sync(0xffffffffu, eps_abs.x, o));
                eps_abs.y = fmax(eps_abs.y, __shfl_xor_sync(0xffffffffu, eps_abs.y, o));
            }
            if (sub == 0) {
                reinterpret_cast<double2*>(s_red)[warp * (CH / 2) + pj] = sq;
                reinterpret_cast<double2*>(s_red)[(kWarps + warp) * (CH / 2) + pj] = eps_abs;
            }
        }
        __syncthreads();  // stage may be refilled by the next iteration's issue; reduction inputs visible
        if (reduce && warp == 0 && lane < CH) {
            double acc = 0.0, m = 0.0;
            for (int w = 0; w < kWarps; ++w) {
                acc += s_red[w * CH + lane];
                m = fmax(m, s_red[(kWarps + w) * CH + lane]);
            }
            const int64_t bl = (int64_t)c * CH + lane;
            if (a.half_sq_part) a.half_sq_part[(int64_t)blockIdx.x * a.ldb + bl] = acc;
            if (STRAIN && a.max_strain_bits)
                atomicMax(a.max_strain_bits + bl, (unsigned long long)__double_as_longlong(m));
        }
    }
}

// out[b] = scale * sum over patches of part[patch][b]: a CTA per 32 columns, 32 strided partial sums per
// column folded in a fixed order
__global__ void __launch_bounds__(1024) patch_column_sum_kernel(const double* __restrict__ part, int64_t rows, int64_t ld,
                                                                int64_t ncol, double scale, double* __restrict__ out) {
    __shared__ double s[32][33];
    const int64_t b = (int64_t)blockIdx.x * 32 + threadIdx.x;
    double acc = 0.0;
    if (b < ncol)
        for (int64_t r = threadIdx.y; r < rows; r += 32) acc += part[r * ld + b];
    s[threadIdx.y][threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.y == 0 && b < ncol) {
        double t = 0.0;
        for (int y = 0; y < 32; ++y) t += s[y][threadIdx.x];
        out[b] = scale * t;
    }
}

template <int DIM, int MODE, bool MATB, bool STRAIN, bool BULK, int CH>
int launch_one(const PatchArgs& a, dim3 grid, size_t smem, cudaStream_t st) {
    auto kern = patch_gather_kernel<DIM, MODE, MATB, STRAIN, BULK, CH>;
    PF_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, PatchCfg<CH>::kThreads, smem, st>>>(a);
    return PF_OK;
}

template <int DIM, int CH>
int launch_dim(const PatchArgs& a, dim3 grid, size_t smem, cudaStream_t st, int mode, bool matb, bool strain, bool bulk) {
#define PF_L(M, MB, S, T) launch_one<DIM, M, MB, S, T, CH>(a, grid, smem, st)
    if (mode == 1) {
        if (matb) return bulk ? PF_L(1, true, false, true) : PF_L(1, true, false, false);
        return bulk ? PF_L(1, false, false, true) : PF_L(1, false, false, false);
    }
    if (strain) {
        if (matb) return bulk ? PF_L(0, true, true, true) : PF_L(0, true, true, false);
        return bulk ? PF_L(0, false, true, true) : PF_L(0, false, true, false);
    }
    if (matb) return bulk ? PF_L(0, true, false, true) : PF_L(0, true, false, false);
    return bulk ? PF_L(0, false, false, true) : PF_L(0, false, false, false);
#undef PF_L
}

}  // namespace

int pf_patch_gather(pf_plan* plan, const PfGatherCall& c, cudaStream_t st, int64_t* columns_done) {
    *columns_done = 0;
    static const int disabled = getenv("PF_NO_PATCH") ? atoi(getenv("PF_NO_PATCH")) : 0;
    static const int env_chunk = getenv("PF_PATCH_CHUNK") ? atoi(getenv("PF_PATCH_CHUNK")) : 0;
    static const int use_bulk = getenv("PF_PATCH_TMA") ? atoi(getenv("PF_PATCH_TMA")) : 0;
    if (disabled || !plan->patch_ok || plan->patches.empty()) return PF_OK;
    const double* x = c.mode == 0 ? c.u : c.v;
    auto aligned16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    if (c.ldb % 2 != 0 || !aligned16(x) || (c.mat_batched && (!aligned16(c.E) || !aligned16(c.A)))) return PF_OK;
    if (!aligned16(c.f_out) || !aligned16(c.r_out) || (c.fext_batched && !aligned16(c.f_ext))) return PF_OK;

    const int dim = plan->dim;
    const int stage_rows = (c.mat_batched ? 2 * plan->max_patch_elems : 0) + plan->max_patch_local * dim;
    const size_t smem32 = patch_smem_bytes<32>(stage_rows, plan->max_patch_inc, plan->max_patch_local, plan->max_patch_elems);
    const size_t smem16 = patch_smem_bytes<16>(stage_rows, plan->max_patch_inc, plan->max_patch_local, plan->max_patch_elems);
    // 16-problem chunks when two such CTAs fit on one SM (227 KB of shared memory, 1 KB reserved per CTA)
    int ch = (2 * (smem16 + 1024) <= 227 * 1024) ? 16 : 32;
    if (env_chunk == 16 || env_chunk == 32) ch = env_chunk;
    const size_t smem = ch == 16 ? smem16 : smem32;
    if (smem > 225 * 1024) return PF_OK;
    const int64_t ncol = (c.B / ch) * ch;  // full chunks only; the tail goes to the generic kernel
    if (ncol < 4 * ch) return PF_OK;       // too few chunks to pipeline: generic kernel

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
    a.fext_bmul = c.fext_batched ? 1 : 0;
    a.load_factor = c.load_factor;
    a.max_elems = plan->max_patch_elems;
    a.max_local = plan->max_patch_local;
    a.max_inc = plan->max_patch_inc;
    const int nchunk = (int)(ncol / ch);
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
    const bool strain = c.mode == 0 && c.max_strain != nullptr;
    int rc;
    if (dim == 1)
        rc = ch == 16 ? launch_dim<1, 16>(a, grid, smem, st, c.mode, c.mat_batched != 0, strain, use_bulk != 0)
                      : launch_dim<1, 32>(a, grid, smem, st, c.mode, c.mat_batched != 0, strain, use_bulk != 0);
    else
        rc = ch == 16 ? launch_dim<2, 16>(a, grid, smem, st, c.mode, c.mat_batched != 0, strain, use_bulk != 0)
                      : launch_dim<2, 32>(a, grid, smem, st, c.mode, c.mat_batched != 0, strain, use_bulk != 0);
    if (rc) return rc;
    PF_CUDA_CHECK(cudaGetLastError());
    if (c.half_sq) {
        patch_column_sum_kernel<<<(unsigned)((ncol + 31) / 32), dim3(32, 32), 0, st>>>(plan->d_work, npatch, c.ldb, ncol,
                                                                                       0.5, c.half_sq);
        PF_CUDA_CHECK(cudaGetLastError());
    }
    *columns_done = ncol;
    return PF_OK;
}
