// Collectives for the element-sharded mesh (second sharding): one process per GPU,
// every rank owns a contiguous set of nodes plus a halo, and per residual evaluation the ranks swap
// the halo rows of one vector (u, then r) with their neighbours and all-reduce one short buffer
// (dL/dtheta and three loss scalars).  The messages are a few kB -- latency bound, not bandwidth
// bound -- so NCCL point-to-point over NVLink/NVSwitch inside one group call is the transport; the
// pack / unpack of halo rows are small kernels on the same stream.
//
// NCCL is resolved with dlopen at first use (the SONAME torch has already loaded is reused), so
// libpinnfem.so keeps loading on machines without it -- the batch-sharded path needs no collective.
#include <dlfcn.h>

#include <cstring>

#include <mutex>
#include <vector>

#include "pf_internal.h"

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

}  // namespace

struct pf_comm {
    ncclComm_t comm = nullptr;
    int world = 1, rank = 0, device = 0;
};

struct pf_halo {
    pf_comm* comm = nullptr;
    int dim = 2;
    std::vector<int> peers;
    std::vector<int64_t> send_ptr, recv_ptr;  // [n_peers + 1] offsets (in nodes) into the lists
    int32_t* d_send_nodes = nullptr;
    int32_t* d_recv_nodes = nullptr;
    double* d_send = nullptr;
    double* d_recv = nullptr;
    int64_t cap_B = 0;  // buffers hold cap_B problems
};

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

extern "C" void pf_comm_destroy(pf_comm* c) {
    if (!c) return;
    if (c->comm && nccl().ok) nccl().CommDestroy(c->comm);
    delete c;
}

extern "C" int pf_comm_allreduce_sum(pf_comm* c, double* buf, int64_t n, void* stream) {
    PF_REQUIRE(c && buf && n >= 0, "pf_comm_allreduce_sum: bad argument");
    if (c->world == 1 || n == 0) return PF_OK;
    PF_NCCL_CHECK(nccl().AllReduce(buf, buf, (size_t)n, kNcclFloat64, kNcclSum, c->comm, pf_stream_of(stream)));
    return PF_OK;
}

extern "C" int pf_halo_create(pf_comm* comm, int dim, int n_peers, const int32_t* peers, const int64_t* send_ptr,
                              const int32_t* send_nodes, const int64_t* recv_ptr, const int32_t* recv_nodes,
                              pf_halo** out) {
    PF_REQUIRE(comm && out, "pf_halo_create: NULL argument");
    PF_REQUIRE(dim == 1 || dim == 2, "dim must be 1 or 2");
    PF_REQUIRE(n_peers >= 0 && (n_peers == 0 || (peers && send_ptr && recv_ptr)), "pf_halo_create: bad peer lists");
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
    const int64_t ns = h->send_ptr.back(), nr = h->recv_ptr.back();
    auto up = [&](const int32_t* src, int64_t n, int32_t** dst) -> int {
        if (n == 0) return PF_OK;
        PF_CUDA_CHECK(cudaMalloc((void**)dst, n * sizeof(int32_t)));
        PF_CUDA_CHECK(cudaMemcpy(*dst, src, n * sizeof(int32_t), cudaMemcpyHostToDevice));
        return PF_OK;
    };
    int rc;
    if ((rc = up(send_nodes, ns, &h->d_send_nodes)) || (rc = up(recv_nodes, nr, &h->d_recv_nodes))) {
        delete h;
        return rc;
    }
    *out = h;
    return PF_OK;
}

extern "C" void pf_halo_destroy(pf_halo* h) {
    if (!h) return;
    cudaFree(h->d_send_nodes);
    cudaFree(h->d_recv_nodes);
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

// accessors used by pf_gd_large.cu
int pf_comm_world(const pf_comm* c) { return c ? c->world : 1; }
pf_comm* pf_halo_comm(pf_halo* h) { return h ? h->comm : nullptr; }
