// Mesh plan: the integer/index work of the hot path, done once per mesh.
//
// Replaces fem/geometry.py:8-9 (element_dofs), fem/boundary.py:8-13
// (free_and_fixed_dofs) and the implicit scatter map of fem/assembly.py:71-72.
// Everything here runs on the host and is exposed through pf_plan_get_array()
// so that tests can compare it bit for bit with the oracle.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "pf_internal.h"

static thread_local std::string g_last_error;

void pf_set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
}

extern "C" const char* pf_last_error(void) { return g_last_error.c_str(); }
extern "C" int pf_version(void) { return PF_VERSION; }

extern "C" int pf_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

extern "C" int pf_plan_create(int dim, int64_t nnode, int64_t nelem, const int64_t* elements,
                              const double* nodes, const int64_t* fixed, int64_t nfixed_in, pf_plan** out) {
    PF_REQUIRE(out != nullptr, "pf_plan_create: out is NULL");
    *out = nullptr;
    PF_REQUIRE(dim == 1 || dim == 2, "dimension must be 1 or 2");
    PF_REQUIRE(nnode > 0 && nnode < (int64_t(1) << 30), "nnode out of range: %lld", (long long)nnode);
    PF_REQUIRE(nelem >= 0 && nelem < (int64_t(1) << 30), "nelem out of range: %lld", (long long)nelem);
    PF_REQUIRE(nodes != nullptr && (nelem == 0 || elements != nullptr), "nodes/elements is NULL");
    PF_REQUIRE(nfixed_in == 0 || fixed != nullptr, "fixed is NULL");

    pf_plan* p = new pf_plan();
    p->dim = dim;
    p->nnode = nnode;
    p->nelem = nelem;
    p->ndof = nnode * dim;
    p->ninc = 2 * nelem;
    p->nodes.assign(nodes, nodes + nnode * dim);

    // --- connectivity + element_dofs (fem/geometry.py:8-9; 1-D: dof == node) ---
    p->conn.resize(2 * nelem);
    p->elem_dofs.resize(nelem * 2 * dim);
    for (int64_t e = 0; e < nelem; ++e) {
        const int64_t i = elements[2 * e], j = elements[2 * e + 1];
        if (i < 0 || i >= nnode || j < 0 || j >= nnode) {
            pf_set_error("element %lld references node outside [0,%lld)", (long long)e, (long long)nnode);
            delete p;
            return PF_ERR_ARG;
        }
        p->conn[2 * e] = (int32_t)i;
        p->conn[2 * e + 1] = (int32_t)j;
        int64_t* d = &p->elem_dofs[e * 2 * dim];
        if (dim == 1) {
            d[0] = i;
            d[1] = j;
        } else {
            d[0] = 2 * i;
            d[1] = 2 * i + 1;
            d[2] = 2 * j;
            d[3] = 2 * j + 1;
        }
    }

    // --- free / fixed partition (fem/boundary.py:8-13: np.unique + mask) ---
    p->dof_free.assign(p->ndof, 1);
    for (int64_t k = 0; k < nfixed_in; ++k) {
        if (fixed[k] < 0 || fixed[k] >= p->ndof) {
            pf_set_error("fixed_dofs contain out-of-range indices");
            delete p;
            return PF_ERR_ARG;
        }
        p->dof_free[fixed[k]] = 0;
    }
    for (int64_t d = 0; d < p->ndof; ++d) (p->dof_free[d] ? p->free_dofs : p->fixed_dofs).push_back(d);
    p->nfree = (int64_t)p->free_dofs.size();
    p->nfixed = (int64_t)p->fixed_dofs.size();

    // --- geometry (fem/element.py:26-28, :59-66; centroids fem/assembly.py:59) ---
    p->l0.resize(nelem);
    p->cosv.resize(nelem);
    p->sinv.resize(nelem);
    p->centroid.resize(nelem * dim);
    for (int64_t e = 0; e < nelem; ++e) {
        const int32_t i = p->conn[2 * e], j = p->conn[2 * e + 1];
        double l0, c, s;
        if (dim == 1) {
            l0 = std::fabs(nodes[j] - nodes[i]);
            c = 1.0;
            s = 0.0;
            p->centroid[e] = (nodes[i] + nodes[j]) / 2.0;
        } else {
            const double dx = nodes[2 * j] - nodes[2 * i], dy = nodes[2 * j + 1] - nodes[2 * i + 1];
            l0 = std::sqrt(dx * dx + dy * dy);
            c = dx / l0;
            s = dy / l0;
            p->centroid[2 * e] = (nodes[2 * i] + nodes[2 * j]) / 2.0;
            p->centroid[2 * e + 1] = (nodes[2 * i + 1] + nodes[2 * j + 1]) / 2.0;
        }
        if (!(l0 > 0.0)) {
            pf_set_error("Element with zero initial length detected");
            delete p;
            return PF_ERR_GEOMETRY;
        }
        p->l0[e] = l0;
        p->cosv[e] = c;
        p->sinv[e] = s;
    }

    // --- node -> element incidence, ascending element id per node ---
    p->inc_ptr.assign(nnode + 1, 0);
    for (int64_t e = 0; e < nelem; ++e) {
        p->inc_ptr[p->conn[2 * e] + 1]++;
        p->inc_ptr[p->conn[2 * e + 1] + 1]++;
    }
    for (int64_t n = 0; n < nnode; ++n) {
        p->max_degree = std::max<int>(p->max_degree, (int)p->inc_ptr[n + 1]);
        p->inc_ptr[n + 1] += p->inc_ptr[n];
    }
    p->inc_elem.resize(p->ninc);
    p->inc_nbr.resize(p->ninc);
    {
        std::vector<int64_t> fill(p->inc_ptr.begin(), p->inc_ptr.end() - 1);
        for (int64_t e = 0; e < nelem; ++e) {
            const int32_t i = p->conn[2 * e], j = p->conn[2 * e + 1];
            p->inc_elem[fill[i]] = e;
            p->inc_nbr[fill[i]++] = j;
            p->inc_elem[fill[j]] = e;
            p->inc_nbr[fill[j]++] = i;
        }
    }

    // --- node-block CSR pattern: row n = sorted unique({n} U neighbours) ---
    p->bsr_rowptr.assign(nnode + 1, 0);
    p->bsr_colind.reserve(nnode + p->ninc);
    p->diag_slot.resize(nnode);
    p->inc_slot.resize(p->ninc);
    p->inc_first.assign(p->ninc, 0);
    std::vector<int64_t> cols;
    for (int64_t n = 0; n < nnode; ++n) {
        cols.clear();
        cols.push_back(n);
        for (int64_t k = p->inc_ptr[n]; k < p->inc_ptr[n + 1]; ++k) cols.push_back(p->inc_nbr[k]);
        std::sort(cols.begin(), cols.end());
        cols.erase(std::unique(cols.begin(), cols.end()), cols.end());
        const int64_t base = (int64_t)p->bsr_colind.size();
        p->bsr_colind.insert(p->bsr_colind.end(), cols.begin(), cols.end());
        p->bsr_rowptr[n + 1] = base + (int64_t)cols.size();
        p->diag_slot[n] = base + (std::lower_bound(cols.begin(), cols.end(), n) - cols.begin());
        for (int64_t k = p->inc_ptr[n]; k < p->inc_ptr[n + 1]; ++k) {
            const int64_t slot = base + (std::lower_bound(cols.begin(), cols.end(), p->inc_nbr[k]) - cols.begin());
            p->inc_slot[k] = slot;
            bool first = (slot != p->diag_slot[n]);  // a self-loop element adds to the diagonal
            for (int64_t q = p->inc_ptr[n]; q < k && first; ++q)
                if (p->inc_slot[q] == slot) first = false;
            p->inc_first[k] = first ? 1 : 0;
            if (!first) p->has_dup = true;
        }
    }
    p->nnzb = (int64_t)p->bsr_colind.size();

    // --- per element BSR slots of blocks (i,i) (i,j) (j,i) (j,j) ---
    p->elem_slots.resize(nelem * 4);
    for (int64_t e = 0; e < nelem; ++e) {
        const int64_t i = p->conn[2 * e], j = p->conn[2 * e + 1];
        auto find = [&](int64_t r, int64_t c) {
            const int64_t* b = &p->bsr_colind[p->bsr_rowptr[r]];
            const int64_t* en = &p->bsr_colind[0] + p->bsr_rowptr[r + 1];
            return p->bsr_rowptr[r] + (std::lower_bound(b, en, c) - b);
        };
        p->elem_slots[4 * e + 0] = find(i, i);
        p->elem_slots[4 * e + 1] = find(i, j);
        p->elem_slots[4 * e + 2] = find(j, i);
        p->elem_slots[4 * e + 3] = find(j, j);
    }

    // --- compact node patches by recursive coordinate bisection -------------------
    // Every leaf gets <= kPatchNodes nodes; the split position follows the number
    // of leaves each side must hold, so leaves are evenly filled.
    {
        std::vector<int32_t> order(nnode);
        for (int64_t n = 0; n < nnode; ++n) order[n] = (int32_t)n;
        struct Range { int64_t lo, hi, leaves; };
        std::vector<Range> stack;
        std::vector<std::pair<int64_t, int64_t>> leaves;
        stack.push_back({0, nnode, (nnode + kPatchNodes - 1) / kPatchNodes});
        const double* X = p->nodes.data();
        while (!stack.empty()) {
            Range r = stack.back();
            stack.pop_back();
            if (r.leaves <= 1) {
                leaves.push_back({r.lo, r.hi});
                continue;
            }
            int axis = 0;
            if (dim == 2) {
                double mn[2] = {1e300, 1e300}, mx[2] = {-1e300, -1e300};
                for (int64_t q = r.lo; q < r.hi; ++q)
                    for (int a = 0; a < 2; ++a) {
                        const double c = X[2 * order[q] + a];
                        mn[a] = std::min(mn[a], c);
                        mx[a] = std::max(mx[a], c);
                    }
                axis = (mx[1] - mn[1] > mx[0] - mn[0]) ? 1 : 0;
            }
            const int64_t lleaves = r.leaves / 2;
            const int64_t mid = r.lo + (r.hi - r.lo) * lleaves / r.leaves;
            auto key = [&](int32_t n) { return dim == 2 ? X[2 * n + axis] : X[n]; };
            std::nth_element(order.begin() + r.lo, order.begin() + mid, order.begin() + r.hi,
                             [&](int32_t a, int32_t b) { const double ka = key(a), kb = key(b); return ka < kb || (ka == kb && a < b); });
            stack.push_back({mid, r.hi, r.leaves - lleaves});
            stack.push_back({r.lo, mid, lleaves});
        }
        std::vector<int32_t> local_of(nnode, -1);
        std::vector<int32_t> elem_local(nelem, -1);
        std::vector<int32_t> touched_nodes, touched_elems;
        for (auto& lf : leaves) {
            if (lf.second <= lf.first) continue;
            PfPatch pt{};
            pt.node_off = (int32_t)p->patch_nodes.size();
            pt.elem_off = (int32_t)p->patch_elems.size();
            pt.inc_off = (int32_t)p->patch_inc.size();
            pt.ptr_off = (int32_t)p->patch_inc_ptr.size();
            std::vector<int32_t> owned(order.begin() + lf.first, order.begin() + lf.second);
            std::sort(owned.begin(), owned.end());
            pt.n_owned = (int32_t)owned.size();
            touched_nodes.clear();
            touched_elems.clear();
            for (int32_t n : owned) {
                local_of[n] = (int32_t)touched_nodes.size();
                touched_nodes.push_back(n);
            }
            // elements incident to owned nodes, ascending id; halo nodes in first-seen order
            for (int32_t n : owned)
                for (int64_t k = p->inc_ptr[n]; k < p->inc_ptr[n + 1]; ++k) touched_elems.push_back((int32_t)p->inc_elem[k]);
            std::sort(touched_elems.begin(), touched_elems.end());
            touched_elems.erase(std::unique(touched_elems.begin(), touched_elems.end()), touched_elems.end());
            for (size_t q = 0; q < touched_elems.size(); ++q) {
                const int32_t el = touched_elems[q];
                elem_local[el] = (int32_t)q;
                for (int e2 = 0; e2 < 2; ++e2) {
                    const int32_t n = p->conn[2 * el + e2];
                    if (local_of[n] < 0) {
                        local_of[n] = (int32_t)touched_nodes.size();
                        touched_nodes.push_back(n);
                    }
                }
            }
            pt.n_local = (int32_t)touched_nodes.size();
            pt.n_elem = (int32_t)touched_elems.size();
            p->patch_nodes.insert(p->patch_nodes.end(), touched_nodes.begin(), touched_nodes.end());
            p->patch_elems.insert(p->patch_elems.end(), touched_elems.begin(), touched_elems.end());
            int32_t cnt = 0;
            for (int32_t n : owned) {
                p->patch_inc_ptr.push_back(cnt);
                for (int64_t k = p->inc_ptr[n]; k < p->inc_ptr[n + 1]; ++k) {
                    p->patch_inc.push_back(PfPatchInc{(int16_t)elem_local[p->inc_elem[k]], (int16_t)local_of[p->inc_nbr[k]]});
                    ++cnt;
                }
            }
            p->patch_inc_ptr.push_back(cnt);  // n_owned + 1 entries per patch
            p->max_patch_elems = std::max(p->max_patch_elems, pt.n_elem);
            p->max_patch_local = std::max(p->max_patch_local, pt.n_local);
            p->max_patch_inc = std::max(p->max_patch_inc, cnt);
            for (int32_t n : touched_nodes) local_of[n] = -1;
            for (int32_t el : touched_elems) elem_local[el] = -1;
            p->patches.push_back(pt);
        }
        p->patch_ok = p->max_patch_elems < 32000 && p->max_patch_local < 32000;
    }
    *out = p;
    return PF_OK;
}

template <typename T>
static int upload(T** dst, const std::vector<T>& src) {
    const size_t bytes = std::max<size_t>(src.size(), 1) * sizeof(T);
    PF_CUDA_CHECK(cudaMalloc((void**)dst, bytes));
    if (!src.empty()) PF_CUDA_CHECK(cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
    return PF_OK;
}

extern "C" int pf_plan_upload(pf_plan* p, int device) {
    PF_REQUIRE(p != nullptr, "pf_plan_upload: plan is NULL");
    PF_REQUIRE(p->device < 0, "plan already uploaded to device %d", p->device);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        pf_set_error("no CUDA device available (%s); libpinnfem has no CPU fallback",
                     e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        return PF_ERR_NO_DEVICE;
    }
    PF_REQUIRE(device >= 0 && device < ndev, "device %d out of range (have %d)", device, ndev);
    PF_CUDA_CHECK(cudaSetDevice(device));
    pf_keep_pool_cached();
    PF_CUDA_CHECK(cudaDeviceGetAttribute(&p->sm_count, cudaDevAttrMultiProcessorCount, device));

    const int dim = p->dim;
    std::vector<int32_t> inc_ptr32(p->inc_ptr.begin(), p->inc_ptr.end());
    std::vector<PfIncidence> inc(p->ninc);
    std::vector<double4> inc_geo(p->ninc), inc_xy(p->ninc);
    for (int64_t n = 0; n < p->nnode; ++n) {
        for (int64_t k = p->inc_ptr[n]; k < p->inc_ptr[n + 1]; ++k) {
            const int64_t el = p->inc_elem[k], m = p->inc_nbr[k];
            inc[k] = PfIncidence{(int32_t)el, (int32_t)m, (int32_t)p->inc_slot[k], (int32_t)p->inc_first[k]};
            inc_geo[k] = make_double4(p->cosv[el], p->sinv[el], 1.0 / p->l0[el], p->l0[el]);
            if (dim == 2)
                inc_xy[k] = make_double4(p->nodes[2 * m], p->nodes[2 * m + 1], p->nodes[2 * n], p->nodes[2 * n + 1]);
            else
                inc_xy[k] = make_double4(p->nodes[m], 0.0, p->nodes[n], 0.0);
        }
    }
    std::vector<int2> conn(p->nelem);
    std::vector<double4> elem_geo(p->nelem), elem_xy(p->nelem);
    for (int64_t el = 0; el < p->nelem; ++el) {
        const int32_t i = p->conn[2 * el], j = p->conn[2 * el + 1];
        conn[el] = make_int2(i, j);
        elem_geo[el] = make_double4(p->cosv[el], p->sinv[el], 1.0 / p->l0[el], p->l0[el]);
        if (dim == 2)
            elem_xy[el] = make_double4(p->nodes[2 * i], p->nodes[2 * i + 1], p->nodes[2 * j], p->nodes[2 * j + 1]);
        else
            elem_xy[el] = make_double4(p->nodes[i], 0.0, p->nodes[j], 0.0);
    }
    std::vector<int32_t> diag32(p->diag_slot.begin(), p->diag_slot.end());
    std::vector<int32_t> free32(p->free_dofs.begin(), p->free_dofs.end());
    std::vector<int32_t> free_index(p->ndof, -1);
    for (int64_t k = 0; k < p->nfree; ++k) free_index[p->free_dofs[k]] = (int32_t)k;
    std::vector<int32_t> rowptr32(p->bsr_rowptr.begin(), p->bsr_rowptr.end());
    std::vector<int32_t> colind32(p->bsr_colind.begin(), p->bsr_colind.end());

    int rc;
    if ((rc = upload(&p->d_inc_ptr, inc_ptr32))) return rc;
    if ((rc = upload(&p->d_inc, inc))) return rc;
    if ((rc = upload(&p->d_inc_geo, inc_geo))) return rc;
    if ((rc = upload(&p->d_inc_xy, inc_xy))) return rc;
    {
        std::vector<int2> inc2(p->ninc);
        for (int64_t k = 0; k < p->ninc; ++k) inc2[k] = make_int2((int)p->inc_elem[k], (int)p->inc_nbr[k]);
        if ((rc = upload(&p->d_inc2, inc2))) return rc;
        if ((rc = upload(&p->d_nodes, p->nodes))) return rc;
    }
    if ((rc = upload(&p->d_diag_slot, diag32))) return rc;
    if ((rc = upload(&p->d_conn, conn))) return rc;
    if ((rc = upload(&p->d_elem_geo, elem_geo))) return rc;
    if ((rc = upload(&p->d_elem_xy, elem_xy))) return rc;
    if ((rc = upload(&p->d_centroid, p->centroid))) return rc;
    if ((rc = upload(&p->d_dof_free, p->dof_free))) return rc;
    if ((rc = upload(&p->d_free_dofs, free32))) return rc;
    if ((rc = upload(&p->d_free_index, free_index))) return rc;
    if ((rc = upload(&p->d_bsr_rowptr, rowptr32))) return rc;
    if ((rc = upload(&p->d_bsr_colind, colind32))) return rc;
    {
        // patch tables; patch_inc_ptr of patch q starts at node_off_owned_prefix(q) + q
        std::vector<double4> pgeo(p->patch_inc.size());
        size_t ptr_pos = 0;
        for (const PfPatch& pt : p->patches) {
            for (int l = 0; l < pt.n_owned; ++l) {
                const int32_t n = p->patch_nodes[pt.node_off + l];
                const int32_t k0 = p->patch_inc_ptr[ptr_pos + l], k1 = p->patch_inc_ptr[ptr_pos + l + 1];
                for (int32_t k = k0; k < k1; ++k) {
                    const int64_t el = p->inc_elem[p->inc_ptr[n] + (k - k0)];
                    pgeo[pt.inc_off + k] = make_double4(p->cosv[el], p->sinv[el], 1.0 / p->l0[el], p->l0[el]);
                }
            }
            ptr_pos += pt.n_owned + 1;
        }
        if ((rc = upload(&p->d_patches, p->patches))) return rc;
        if ((rc = upload(&p->d_patch_nodes, p->patch_nodes))) return rc;
        if ((rc = upload(&p->d_patch_elems, p->patch_elems))) return rc;
        if ((rc = upload(&p->d_patch_inc_ptr, p->patch_inc_ptr))) return rc;
        if ((rc = upload(&p->d_patch_inc, p->patch_inc))) return rc;
        if ((rc = upload(&p->d_patch_inc_geo, pgeo))) return rc;
    }
    p->device = device;
    return PF_OK;
}

int pf_plan_activate(const pf_plan* p) {
    PF_REQUIRE(p != nullptr, "plan is NULL");
    if (p->device < 0) {
        pf_set_error("plan is not uploaded to a CUDA device (call pf_plan_upload); there is no CPU fallback");
        return PF_ERR_NO_DEVICE;
    }
    PF_CUDA_CHECK(cudaSetDevice(p->device));
    return PF_OK;
}

void pf_keep_pool_cached() {
    static bool done[64] = {false};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || done[dev]) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        uint64_t keep = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    cudaGetLastError();
    done[dev] = true;
}

int pf_plan_reserve_work(pf_plan* p, size_t bytes) {
    if (bytes <= p->work_bytes) return PF_OK;
    if (p->d_work) PF_CUDA_CHECK(cudaFree(p->d_work));
    p->d_work = nullptr;
    p->work_bytes = 0;
    PF_CUDA_CHECK(cudaMalloc((void**)&p->d_work, bytes));
    p->work_bytes = bytes;
    return PF_OK;
}

extern "C" void pf_plan_destroy(pf_plan* p) {
    if (!p) return;
    if (p->device >= 0) {
        cudaSetDevice(p->device);
        void* ptrs[] = {p->d_inc2, p->d_nodes, p->d_inc_ptr, p->d_inc, p->d_inc_geo, p->d_inc_xy, p->d_diag_slot, p->d_conn,
                        p->d_elem_geo, p->d_elem_xy, p->d_centroid, p->d_dof_free, p->d_free_dofs,
                        p->d_free_index, p->d_bsr_rowptr, p->d_bsr_colind, p->d_work, p->d_patches,
                        p->d_patch_nodes, p->d_patch_elems, p->d_patch_inc_ptr, p->d_patch_inc, p->d_patch_inc_geo};
        for (void* q : ptrs)
            if (q) cudaFree(q);
        for (int s = 0; s < 3; ++s) {
            for (int a = 0; a < 4; ++a)
                if (p->d_stage[s][a]) cudaFree(p->d_stage[s][a]);
            if (p->host_streams[s]) cudaStreamDestroy(p->host_streams[s]);
        }
    }
    delete p;
}

extern "C" int64_t pf_plan_size(const pf_plan* p, int what) {
    if (!p) return -1;
    switch (what) {
        case PF_PLAN_DIM: return p->dim;
        case PF_PLAN_NNODE: return p->nnode;
        case PF_PLAN_NELEM: return p->nelem;
        case PF_PLAN_NDOF: return p->ndof;
        case PF_PLAN_NFREE: return p->nfree;
        case PF_PLAN_NFIXED: return p->nfixed;
        case PF_PLAN_NNZB: return p->nnzb;
        case PF_PLAN_NINC: return p->ninc;
        case PF_PLAN_MAX_DEGREE: return p->max_degree;
        case PF_PLAN_HAS_DUPLICATE_EDGES: return p->has_dup ? 1 : 0;
        case PF_PLAN_DEVICE: return p->device;
        default: return -1;
    }
}

static const std::vector<int64_t>* plan_array(const pf_plan* p, int which) {
    switch (which) {
        case PF_ARR_ELEM_DOFS: return &p->elem_dofs;
        case PF_ARR_FREE_DOFS: return &p->free_dofs;
        case PF_ARR_FIXED_DOFS: return &p->fixed_dofs;
        case PF_ARR_BSR_ROWPTR: return &p->bsr_rowptr;
        case PF_ARR_BSR_COLIND: return &p->bsr_colind;
        case PF_ARR_ELEM_SLOTS: return &p->elem_slots;
        case PF_ARR_INC_PTR: return &p->inc_ptr;
        case PF_ARR_INC_ELEM: return &p->inc_elem;
        case PF_ARR_INC_NBR: return &p->inc_nbr;
        case PF_ARR_INC_SLOT: return &p->inc_slot;
        case PF_ARR_DIAG_SLOT: return &p->diag_slot;
        default: return nullptr;
    }
}

extern "C" int64_t pf_plan_array_len(const pf_plan* p, int which) {
    if (!p) return -1;
    const std::vector<int64_t>* a = plan_array(p, which);
    return a ? (int64_t)a->size() : -1;
}

extern "C" int pf_plan_get_array(const pf_plan* p, int which, int64_t* dst, int64_t dst_len) {
    PF_REQUIRE(p != nullptr && dst != nullptr, "pf_plan_get_array: NULL argument");
    const std::vector<int64_t>* a = plan_array(p, which);
    PF_REQUIRE(a != nullptr, "pf_plan_get_array: unknown array id %d", which);
    PF_REQUIRE(dst_len == (int64_t)a->size(), "pf_plan_get_array: dst_len %lld != %lld", (long long)dst_len,
               (long long)a->size());
    if (!a->empty()) std::memcpy(dst, a->data(), a->size() * sizeof(int64_t));
    return PF_OK;
}

extern "C" int pf_plan_get_geometry(const pf_plan* p, int which, double* dst, int64_t dst_len) {
    PF_REQUIRE(p != nullptr && dst != nullptr, "pf_plan_get_geometry: NULL argument");
    const std::vector<double>* a = which == 0 ? &p->l0 : which == 1 ? &p->cosv : which == 2 ? &p->sinv
                                   : which == 3 ? &p->centroid : nullptr;
    PF_REQUIRE(a != nullptr, "pf_plan_get_geometry: unknown id %d", which);
    PF_REQUIRE(dst_len == (int64_t)a->size(), "pf_plan_get_geometry: dst_len mismatch");
    if (!a->empty()) std::memcpy(dst, a->data(), a->size() * sizeof(double));
    return PF_OK;
}
