// Peer-memory transport of the element-sharded path (one process per GPU, NVLink / NVSwitch).
//
// Every rank owns one "mailbox" in its own HBM, mapped into every other rank's address space through CUDA
// IPC.  A message is written by the SENDER's kernel straight into the receiver's mailbox with ordinary
// stores over NVLink, followed by one release store of an epoch number; the receiver's kernel spins on
// that number with acquire loads and then reads the payload out of its own HBM.  No NCCL launch, no
// staging copy, no host involvement: a halo exchange is ONE kernel (pack -> remote store -> flag -> wait
// -> unpack) and the gradient all-reduce runs INSIDE the kernel that consumes it
// (adam_theta_finish_kernel in pf_gd_large.cu).
//
// Mailbox layout (bytes):  [0, 1024)            flags: uint64 halo_flag[kPfMaxRanks], ar_flag[kPfMaxRanks]
//                          then                 double halo[2][world][halo_slot]   (parity, source rank)
//                          then                 double ar[2][world][ar_slot]
// Slots are double buffered by epoch parity.  Every exchange is symmetric (push, then wait for the same
// peer), so a sender can be at most one epoch ahead of the reader of its previous message: when rank A
// writes epoch e + 2 into a slot, it has seen rank B's epoch e + 1, which B sent after it had finished
// reading epoch e.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

constexpr int kPfMaxRanks = 16;
constexpr int64_t kPfMailboxHeader = 1024;
constexpr unsigned long long kPfPeerTimeoutNs = 60ull * 1000ull * 1000ull * 1000ull;

struct PfPeerView {
    int world = 1, rank = 0;
    int64_t halo_slot = 0, ar_slot = 0;  // doubles per (parity, source rank)
    char* box[kPfMaxRanks] = {};         // mailbox of every rank as mapped in THIS process (box[rank]: own HBM)
    unsigned long long* epochs = nullptr;  // local: [kPfMaxRanks] halo epochs per peer rank, [kPfMaxRanks] all-reduce epoch
    int* status = nullptr;                 // local: != 0 after a wait timed out (the peer died or never called)
};

#ifdef __CUDACC__
__device__ __forceinline__ unsigned long long* pf_peer_halo_flag(char* box, int src) {
    return reinterpret_cast<unsigned long long*>(box) + src;
}
__device__ __forceinline__ unsigned long long* pf_peer_ar_flag(char* box, int src) {
    return reinterpret_cast<unsigned long long*>(box) + kPfMaxRanks + src;
}
__device__ __forceinline__ double* pf_peer_halo_slot(const PfPeerView& v, char* box, int parity, int src) {
    return reinterpret_cast<double*>(box + kPfMailboxHeader) + ((int64_t)parity * v.world + src) * v.halo_slot;
}
__device__ __forceinline__ double* pf_peer_ar_slot(const PfPeerView& v, char* box, int parity, int src) {
    return reinterpret_cast<double*>(box + kPfMailboxHeader) + 2 * (int64_t)v.world * v.halo_slot +
           ((int64_t)parity * v.world + src) * v.ar_slot;
}
// publish: every store this CTA made before the preceding __syncthreads() is visible system-wide before `e`
__device__ __forceinline__ void pf_peer_signal(unsigned long long* flag, unsigned long long e) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag), "l"(e) : "memory");
}
__device__ __forceinline__ unsigned long long pf_peer_poll(const unsigned long long* flag) {
    unsigned long long x;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(x) : "l"(flag) : "memory");
    return x;
}
__device__ __forceinline__ unsigned long long pf_globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// spin until *flag >= e; gives up after kPfPeerTimeoutNs (and at once when an earlier wait already failed), so a
// lost peer surfaces as an error on the host instead of a hung GPU
__device__ __forceinline__ void pf_peer_wait(const PfPeerView& v, const unsigned long long* flag, unsigned long long e) {
    if (pf_peer_poll(flag) >= e) return;
    if (*reinterpret_cast<volatile int*>(v.status) != 0) return;
    const unsigned long long t0 = pf_globaltimer();
    while (pf_peer_poll(flag) < e) {
        if (pf_globaltimer() - t0 > kPfPeerTimeoutNs) {
            atomicExch(v.status, 1);
            return;
        }
        __nanosleep(64);
    }
}

// Sum of buf[0..n) over all ranks, in rank order 0, 1, ... (the same bits on every rank), left in buf.
// Called by ALL threads of exactly one CTA per rank; n <= v.ar_slot.
__device__ __forceinline__ void pf_peer_allreduce_block(const PfPeerView& v, double* __restrict__ buf, int n) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const unsigned long long e = v.epochs[kPfMaxRanks] + 1;
    const int parity = (int)(e & 1);
    for (int p = 0; p < v.world; ++p) {
        if (p == v.rank) continue;
        double* dst = pf_peer_ar_slot(v, v.box[p], parity, v.rank);
        for (int i = tid; i < n; i += nt) dst[i] = buf[i];
    }
    __syncthreads();
    if (tid < v.world && tid != v.rank) {
        pf_peer_signal(pf_peer_ar_flag(v.box[tid], v.rank), e);
        pf_peer_wait(v, pf_peer_ar_flag(v.box[v.rank], tid), e);
    }
    __syncthreads();
    for (int i = tid; i < n; i += nt) {
        double s = 0.0;
        for (int r = 0; r < v.world; ++r) {
            const double x = r == v.rank ? buf[i] : __ldcv(pf_peer_ar_slot(v, v.box[v.rank], parity, r) + i);
            s = r == 0 ? x : s + x;
        }
        buf[i] = s;
    }
    __syncthreads();
    if (tid == 0) v.epochs[kPfMaxRanks] = e;
}
#endif  // __CUDACC__
