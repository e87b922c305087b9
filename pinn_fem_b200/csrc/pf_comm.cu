// Collectives for the element-sharded mesh (second sharding): one process per GPU,
// every rank owns a contiguous set of nodes plus a halo, and per residual evaluation the ranks swap
// the halo rows of one vector (u, then r) with their neighbours and all-reduce one short buffer
// (dL/dtheta and three loss scalars).  The messages are a few kB -- latency bound, not bandwidth
// bound.  Two transports:
//   * peer memory (pf_peer.cuh, default when the GPUs can map each other's memory): the sending kernel
//     stores the halo rows straight into the receiver's mailbox over NVLink and raises an epoch flag;
//     one kernel per exchange, the all-reduce fused into its consumer; no NCCL launch on the data path;
//   * NCCL point-to-point inside one group call + pack / unpack kernels: messages larger than a mailbox
//     slot, or boxes without peer access.
//
// NCCL is resolved with dlopen at first use (the SONAME torch has already loaded is reused), so
// libpinnfem.so keeps loading on machines without it -- the batch-sharded path needs no collective.
#include <dlfcn.h>

#include <algorithm>
#include <cstring>

#include <mutex>
#include <vector>

#include "pf_internal.h"
#include "pf_peer.cuh"

namespace {

// the slice of the NCCL ABI this file uses (nccl.h: stable since 2.x)
typedef struct ncclComm* ncclComm_t;
typedef struct {
    char internal[128];
} ncclUniqueId;
typedef int ncclResult_t;
constexpr int kNcclFloat64 = 8, kNcclSum = 0;

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

NcclApi& nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            api.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (!api.handle) return;
        auto sym = [&](const char* n) { return dlsym(api.handle, n); };
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
        api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
        api.Send = reinterpret_cast<decltype(api.Send)>(sym("ncclSend"));
        api.Recv = reinterpret_cast<decltype(api.Recv)>(sym("ncclRecv"));
        api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
        api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
        api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.Send && api.Recv &&
                 api.GroupStart && api.GroupEnd;
    });
    return api;
}

#define PF_NCCL_CHECK(expr)                                                                              \
    do {                                                                                                 \
        ncclResult_t _r = (expr);                                                                        \
        if (_r != 0) {                                                                                   \
            pf_set_error("%s failed: %s", #expr, nccl().GetErrorString ? nccl().GetErrorString(_r) : "NCCL error"); \
            return PF_ERR_CUDA;                                                                          \
        }                                                                                                \
    } while (0)

// send[k][b] = x[dof(k)][b] for the listed local nodes (all their DOFs)
__global__ void halo_pack_kernel(const int32_t* __restrict__ nodes, int64_t n_nodes, int dim, int64_t B,
                                 const double* __restrict__ x, double* __restrict__ buf) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = n_nodes * dim * B;
    if (q >= total) return;
    const int64_t b = q % B, k = q / B;
    const int64_t node = nodes[k / dim];
    buf[q] = x[(node * dim + (k % dim)) * B + b];
}

__global__ void halo_unpack_kernel(const int32_t* __restrict__ nodes, int64_t n_nodes, int dim, int64_t B,
                                   const double* __restrict__ buf, double* __restrict__ x) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = n_nodes * dim * B;
    if (q >= total) return;
    const int64_t b = q % B, k = q / B;
    const int64_t node = nodes[k / dim];
    x[(node * dim + (k % dim)) * B + b] = buf[q];
}

// device copies of one rank's exchange lists
struct HaloLists {
    const int32_t* peers;       // [n_peers] ranks
    const int64_t* send_ptr;    // [n_peers + 1]
    const int64_t* recv_ptr;    // [n_peers + 1]
    const int32_t* send_nodes;  // local node ids
    const int32_t* recv_nodes;
};

// One CTA per neighbour: pack the rows of the send nodes straight into the neighbour's mailbox (remote
// stores over NVLink), publish the epoch, wait for the neighbour's epoch, unpack its rows out of the own
// mailbox into the halo rows of x.  Send rows (owned nodes) and halo rows are disjoint.
__global__ void __launch_bounds__(512) halo_peer_kernel(PfPeerView v, HaloLists l, int dim, int64_t B,
                                                        double* __restrict__ x) {
    const int k = blockIdx.x, peer = l.peers[k];
    const unsigned long long e = v.epochs[peer] + 1;
    const int parity = (int)(e & 1);
    {
        const int64_t s0 = l.send_ptr[k], total = (l.send_ptr[k + 1] - s0) * dim * B;
        double* dst = pf_peer_halo_slot(v, v.box[peer], parity, v.rank);
        for (int64_t q = threadIdx.x; q < total; q += blockDim.x) {
            const int64_t b = q % B, kk = q / B;
            const int64_t node = l.send_nodes[s0 + kk / dim];
            dst[q] = x[(node * dim + (kk % dim)) * B + b];
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        pf_peer_signal(pf_peer_halo_flag(v.box[peer], v.rank), e);
        pf_peer_wait(v, pf_peer_halo_flag(v.box[v.rank], peer), e);
    }
    __syncthreads();
    {
        const int64_t r0 = l.recv_ptr[k], total = (l.recv_ptr[k + 1] - r0) * dim * B;
        const double* src = pf_peer_halo_slot(v, v.box[v.rank], parity, peer);
        for (int64_t q = threadIdx.x; q < total; q += blockDim.x) {
            const int64_t b = q % B, kk = q / B;
            const int64_t node = l.recv_nodes[r0 + kk / dim];
            x[(node * dim + (kk % dim)) * B + b] = __ldcv(src + q);
        }
    }
    if (threadIdx.x == 0) v.epochs[peer] = e;
}

__global__ void __launch_bounds__(1024) allreduce_peer_kernel(PfPeerView v, double* __restrict__ buf, int n) {
    pf_peer_allreduce_block(v, buf, n);
}

}  // namespace

struct pf_comm {
    ncclComm_t comm = nullptr;
    int world = 1, rank = 0, device = 0;
    // peer-memory transport
    bool peer_on = false;
    PfPeerView view;
    char* mailbox = nullptr;
    size_t mailbox_bytes = 0;
    void* imported[kPfMaxRanks] = {};
};

struct pf_halo {
    pf_comm* comm = nullptr;
    int dim = 2;
    std::vector<int> peers;
    std::vector<int64_t> send_ptr, recv_ptr;  // [n_peers + 1] offsets (in nodes) into the lists
    int32_t* d_send_nodes = nullptr;
    int32_t* d_recv_nodes = nullptr;
    int32_t* d_peers = nullptr;      // device copies of peers / send_ptr / recv_ptr for the peer-memory kernel
    int64_t* d_send_ptr = nullptr;
    int64_t* d_recv_ptr = nullptr;
    int64_t max_msg_nodes = 0;       // largest send or receive list of one neighbour (of ANY rank once the host has reduced it)
    double* d_send = nullptr;
    double* d_recv = nullptr;
    int64_t cap_B = 0;  // buffers hold cap_B problems
};

bool pf_halo_uses_peer(const pf_halo* h, int64_t B);

extern "C" int pf_comm_available(void) { return nccl().ok ? 1 : 0; }

extern "C" int pf_comm_unique_id(unsigned char* id128) {
    PF_REQUIRE(id128 != nullptr, "id is NULL");
    PF_REQUIRE(nccl().ok, "NCCL is not available (libnccl.so.2 not found)");
    ncclUniqueId id;
    PF_NCCL_CHECK(nccl().GetUniqueId(&id));
    memcpy(id128, id.internal, 128);
    return PF_OK;
}

extern "C" int pf_comm_create(int world, int rank, const unsigned char* id128, int device, pf_comm** out) {
    PF_REQUIRE(out && id128, "pf_comm_create: NULL argument");
    PF_REQUIRE(world >= 1 && rank >= 0 && rank < world, "bad rank/world %d/%d", rank, world);
    PF_REQUIRE(nccl().ok, "NCCL is not available (libnccl.so.2 not found)");
    PF_CUDA_CHECK(cudaSetDevice(device));
    ncclUniqueId id;
    memcpy(id.internal, id128, 128);
    pf_comm* c = new pf_comm();
    c->world = world;
    c->rank = rank;
    c->device = device;
    ncclResult_t r = nccl().CommInitRank(&c->comm, world, id, rank);
    if (r != 0) {
        pf_set_error("ncclCommInitRank failed: %s", nccl().GetErrorString ? nccl().GetErrorString(r) : "NCCL error");
        delete c;
        return PF_ERR_CUDA;
    }
    *out = c;
    return PF_OK;
}

extern "C" void pf_comm_peer_detach(pf_comm* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    for (int r = 0; r < kPfMaxRanks; ++r) {
        if (c->imported[r]) cudaIpcCloseMemHandle(c->imported[r]);
        c->imported[r] = nullptr;
    }
    c->peer_on = false;
}

extern "C" void pf_comm_destroy(pf_comm* c) {
    if (!c) return;
    pf_comm_peer_detach(c);
    cudaFree(c->mailbox);
    cudaFree(c->view.epochs);
    cudaFree(c->view.status);
    if (c->comm && nccl().ok) nccl().CommDestroy(c->comm);
    delete c;
}

// ---- peer-memory transport: export the own mailbox, import everybody else's ----
extern "C" int pf_comm_peer_export(pf_comm* c, int64_t halo_slot_doubles, int64_t ar_slot_doubles,
                                   unsigned char* handle64) {
    PF_REQUIRE(c && handle64, "pf_comm_peer_export: NULL argument");
    PF_REQUIRE(c->world >= 2 && c->world <= kPfMaxRanks, "peer transport supports 2..%d ranks (world %d)", kPfMaxRanks,
               c->world);
    PF_REQUIRE(halo_slot_doubles >= 1 && ar_slot_doubles >= 1, "slot sizes must be positive");
    PF_REQUIRE(!c->mailbox, "mailbox already exported");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    PF_CUDA_CHECK(cudaSetDevice(c->device));
    c->view.world = c->world;
    c->view.rank = c->rank;
    c->view.halo_slot = halo_slot_doubles;
    c->view.ar_slot = ar_slot_doubles;
    c->mailbox_bytes = (size_t)kPfMailboxHeader +
                       2 * (size_t)c->world * (size_t)(halo_slot_doubles + ar_slot_doubles) * sizeof(double);
    PF_CUDA_CHECK(cudaMalloc((void**)&c->mailbox, c->mailbox_bytes));  // cudaMalloc, not the async pool: IPC-exportable
    PF_CUDA_CHECK(cudaMemset(c->mailbox, 0, c->mailbox_bytes));
    PF_CUDA_CHECK(cudaMalloc((void**)&c->view.epochs, 2 * kPfMaxRanks * sizeof(unsigned long long)));
    PF_CUDA_CHECK(cudaMemset(c->view.epochs, 0, 2 * kPfMaxRanks * sizeof(unsigned long long)));
    PF_CUDA_CHECK(cudaMalloc((void**)&c->view.status, sizeof(int)));
    PF_CUDA_CHECK(cudaMemset(c->view.status, 0, sizeof(int)));
    PF_CUDA_CHECK(cudaDeviceSynchronize());
    cudaIpcMemHandle_t h;
    PF_CUDA_CHECK(cudaIpcGetMemHandle(&h, c->mailbox));
    memcpy(handle64, &h, 64);
    return PF_OK;
}

// handles: [world][64] in rank order (the own entry is ignored).  Every rank must have exported first.
extern "C" int pf_comm_peer_import(pf_comm* c, const unsigned char* handles) {
    PF_REQUIRE(c && handles && c->mailbox, "pf_comm_peer_import: export first");
    PF_CUDA_CHECK(cudaSetDevice(c->device));
    for (int r = 0; r < c->world; ++r) {
        if (r == c->rank) {
            c->view.box[r] = c->mailbox;
            continue;
        }
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + (size_t)r * 64, 64);
        void* p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            pf_set_error("cudaIpcOpenMemHandle(rank %d) failed: %s", r, cudaGetErrorString(e));
            pf_comm_peer_detach(c);
            return PF_ERR_CUDA;
        }
        c->imported[r] = p;
        c->view.box[r] = static_cast<char*>(p);
    }
    c->peer_on = true;
    return PF_OK;
}

extern "C" int pf_comm_peer_enabled(const pf_comm* c) { return c && c->peer_on ? 1 : 0; }

// 0 = healthy; raises after a peer wait timed out (synchronises the stream)
extern "C" int pf_comm_peer_check(pf_comm* c, void* stream) {
    PF_REQUIRE(c, "pf_comm_peer_check: NULL communicator");
    if (!c->peer_on) return PF_OK;
    int st = 0;
    PF_CUDA_CHECK(cudaMemcpyAsync(&st, c->view.status, sizeof(int), cudaMemcpyDeviceToHost, pf_stream_of(stream)));
    PF_CUDA_CHECK(cudaStreamSynchronize(pf_stream_of(stream)));
    if (st != 0) {
        pf_set_error("peer-memory exchange timed out on rank %d: a neighbour never arrived", c->rank);
        return PF_ERR_CUDA;
    }
    return PF_OK;
}

extern "C" int pf_comm_allreduce_sum(pf_comm* c, double* buf, int64_t n, void* stream) {
    PF_REQUIRE(c && buf && n >= 0, "pf_comm_allreduce_sum: bad argument");
    if (c->world == 1 || n == 0) return PF_OK;
    if (c->peer_on && n <= c->view.ar_slot) {  // sum in rank order: bitwise identical on every rank
        allreduce_peer_kernel<<<1, 1024, 0, pf_stream_of(stream)>>>(c->view, buf, (int)n);
        PF_CUDA_CHECK(cudaGetLastError());
        return PF_OK;
    }
    PF_NCCL_CHECK(nccl().AllReduce(buf, buf, (size_t)n, kNcclFloat64, kNcclSum, c->comm, pf_stream_of(stream)));
    return PF_OK;
}

extern "C" int pf_halo_create(pf_comm* comm, int dim, int n_peers, const int32_t* peers, const int64_t* send_ptr,
                              const int32_t* send_nodes, const int64_t* recv_ptr, const int32_t* recv_nodes,
                              pf_halo** out) {
    PF_REQUIRE(comm && out, "pf_halo_create: NULL argument");
    PF_REQUIRE(dim == 1 || dim == 2, "dim must be 1 or 2");
    PF_REQUIRE(n_peers >= 0 && (n_peers == 0 || (peers && send_ptr && recv_ptr)), "pf_halo_create: bad peer lists");
    for (int k = 0; k < n_peers; ++k)
        PF_REQUIRE(peers[k] >= 0 && peers[k] < comm->world && peers[k] != comm->rank, "bad peer rank %d", peers[k]);
    PF_CUDA_CHECK(cudaSetDevice(comm->device));
    pf_halo* h = new pf_halo();
    h->comm = comm;
    h->dim = dim;
    h->peers.assign(peers, peers + n_peers);
    h->send_ptr.assign(1, 0);
    h->recv_ptr.assign(1, 0);
    if (n_peers) {
        h->send_ptr.assign(send_ptr, send_ptr + n_peers + 1);
        h->recv_ptr.assign(recv_ptr, recv_ptr + n_peers + 1);
    }
    for (int k = 0; k < n_peers; ++k)
        h->max_msg_nodes = std::max({h->max_msg_nodes, send_ptr[k + 1] - send_ptr[k], recv_ptr[k + 1] - recv_ptr[k]});
    const int64_t ns = h->send_ptr.back(), nr = h->recv_ptr.back();
    auto up = [&](const void* src, size_t bytes, void** dst) -> int {
        if (bytes == 0) return PF_OK;
        PF_CUDA_CHECK(cudaMalloc(dst, bytes));
        PF_CUDA_CHECK(cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice));
        return PF_OK;
    };
    const size_t pb = n_peers ? (size_t)(n_peers + 1) * sizeof(int64_t) : 0;
    int rc;
    if ((rc = up(send_nodes, ns * sizeof(int32_t), (void**)&h->d_send_nodes)) ||
        (rc = up(recv_nodes, nr * sizeof(int32_t), (void**)&h->d_recv_nodes)) ||
        (rc = up(h->peers.data(), n_peers * sizeof(int32_t), (void**)&h->d_peers)) ||
        (rc = up(h->send_ptr.data(), pb, (void**)&h->d_send_ptr)) ||
        (rc = up(h->recv_ptr.data(), pb, (void**)&h->d_recv_ptr))) {
        pf_halo_destroy(h);
        return rc;
    }
    *out = h;
    return PF_OK;
}

extern "C" void pf_halo_destroy(pf_halo* h) {
    if (!h) return;
    cudaFree(h->d_send_nodes);
    cudaFree(h->d_recv_nodes);
    cudaFree(h->d_peers);
    cudaFree(h->d_send_ptr);
    cudaFree(h->d_recv_ptr);
    cudaFree(h->d_send);
    cudaFree(h->d_recv);
    delete h;
}

// x dev [ndof_local][B]: rows of the send nodes go to their peers, rows of the halo nodes are overwritten
// with the owners' values.  One NCCL group; enqueued on `stream`.
extern "C" int pf_halo_exchange(pf_halo* h, double* x, int64_t B, void* stream) {
    PF_REQUIRE(h && x && B >= 1, "pf_halo_exchange: bad argument");
    if (h->comm->world == 1 || h->peers.empty()) return PF_OK;
    cudaStream_t st = pf_stream_of(stream);
    const int64_t ns = h->send_ptr.back(), nr = h->recv_ptr.back();
    const int64_t row = (int64_t)h->dim * B;  // doubles per node
    if (pf_halo_uses_peer(h, B)) {
        const HaloLists l{h->d_peers, h->d_send_ptr, h->d_recv_ptr, h->d_send_nodes, h->d_recv_nodes};
        halo_peer_kernel<<<(unsigned)h->peers.size(), 512, 0, st>>>(h->comm->view, l, h->dim, B, x);
        PF_CUDA_CHECK(cudaGetLastError());
        return PF_OK;
    }
    if (B > h->cap_B) {
        PF_CUDA_CHECK(cudaStreamSynchronize(st));
        cudaFree(h->d_send);
        cudaFree(h->d_recv);
        h->d_send = h->d_recv = nullptr;
        PF_CUDA_CHECK(cudaMalloc((void**)&h->d_send, std::max<int64_t>(ns * row, 1) * sizeof(double)));
        PF_CUDA_CHECK(cudaMalloc((void**)&h->d_recv, std::max<int64_t>(nr * row, 1) * sizeof(double)));
        h->cap_B = B;
    }
    if (ns) halo_pack_kernel<<<(unsigned)((ns * row + 255) / 256), 256, 0, st>>>(h->d_send_nodes, ns, h->dim, B, x, h->d_send);
    PF_CUDA_CHECK(cudaGetLastError());
    PF_NCCL_CHECK(nccl().GroupStart());
    for (size_t k = 0; k < h->peers.size(); ++k) {
        const int64_t s0 = h->send_ptr[k], s1 = h->send_ptr[k + 1], r0 = h->recv_ptr[k], r1 = h->recv_ptr[k + 1];
        if (s1 > s0)
            PF_NCCL_CHECK(nccl().Send(h->d_send + s0 * row, (size_t)((s1 - s0) * row), kNcclFloat64, h->peers[k],
                                      h->comm->comm, st));
        if (r1 > r0)
            PF_NCCL_CHECK(nccl().Recv(h->d_recv + r0 * row, (size_t)((r1 - r0) * row), kNcclFloat64, h->peers[k],
                                      h->comm->comm, st));
    }
    PF_NCCL_CHECK(nccl().GroupEnd());
    if (nr) halo_unpack_kernel<<<(unsigned)((nr * row + 255) / 256), 256, 0, st>>>(h->d_recv_nodes, nr, h->dim, B, h->d_recv, x);
    PF_CUDA_CHECK(cudaGetLastError());
    return PF_OK;
}

extern "C" int64_t pf_halo_max_message_nodes(const pf_halo* h) { return h ? h->max_msg_nodes : 0; }

extern "C" int pf_halo_set_max_message_nodes(pf_halo* h, int64_t global_max_nodes) {
    PF_REQUIRE(h, "pf_halo_set_max_message_nodes: NULL halo");
    PF_REQUIRE(global_max_nodes >= h->max_msg_nodes, "global maximum %lld is below this rank's own %lld",
               (long long)global_max_nodes, (long long)h->max_msg_nodes);
    h->max_msg_nodes = global_max_nodes;
    return PF_OK;
}

// accessors used by pf_gd_large.cu
int pf_comm_world(const pf_comm* c) { return c ? c->world : 1; }
bool pf_halo_uses_peer(const pf_halo* h, int64_t B) {
    return h && h->comm->peer_on && h->max_msg_nodes * h->dim * B <= h->comm->view.halo_slot;
}
bool pf_comm_peer_view(const pf_comm* c, PfPeerView* out) {
    if (!c || !c->peer_on) return false;
    *out = c->view;
    return true;
}
pf_comm* pf_halo_comm(pf_halo* h) { return h ? h->comm : nullptr; }
