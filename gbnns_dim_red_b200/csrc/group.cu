// group.cu — several GPUs of one node behind the C ABI (gbdr_group_*): the reference's only parallelism is a thread team
// over the queries of a batch (`#pragma omp parallel for`, search/search_function.h:147-152); its multi-GPU equivalents are
//   replicated  the whole index on every device, the batch cut into contiguous slices, no exchange step;
//   sharded     rows [b_i, e_i) of the base / low-dimensional base and a graph over local ids on device i, every device
//               answers every query on its shard (global ids via the id offset), and ONE exchange step: the root device
//               merges the per-shard (dist, id) lists.  Two implementations of that step, same results:
//                 GBDR_EXCHANGE_PEER  the merge kernel reads the members' result buffers where they lie, over NVLink
//                                     peer loads (gather and merge fused in one kernel, nothing staged);
//                 GBDR_EXCHANGE_NCCL  ncclAllGather of the lists on every member's stream, then the K5 merge kernel.
// One host thread per member enqueues that member's copies and kernels, so launch overhead does not add up over devices.
// NCCL is loaded at run time (dlopen of $GBDR_NCCL_LIB, else "libnccl.so.2": the one torch already mapped when the caller
// is a torch process, the system one otherwise); without it only the NCCL exchange is unavailable.
#include <dlfcn.h>
#include <nccl.h>

#include <chrono>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "index.cuh"

namespace gbdr {

namespace {

// merge of `parts` ascending (dist, id) lists per query read through a pointer table: list p of query q starts at
// ids[p] + q * k_in on whatever device owns it (peer-mapped).  Same ranking rule as merge.cu.
__global__ void merge_topk_ptrs_kernel(const uint32_t* const* __restrict__ ids, const float* const* __restrict__ dists,
                                       uint32_t parts, uint32_t n_q, uint32_t k_in, uint32_t k_out,
                                       uint32_t* __restrict__ out_ids, float* __restrict__ out_dists) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t per_q = (uint64_t)parts * k_in;
    if (t >= per_q * n_q) return;
    const uint32_t q = (uint32_t)(t / per_q);
    const uint32_t rem = (uint32_t)(t % per_q);
    const uint32_t pp = rem / k_in, j = rem % k_in;
    const size_t row = (size_t)q * k_in;
    const uint32_t id = ids[pp][row + j];
    if (id == PAD_ID) return;
    const float dist = dists[pp][row + j];
    uint32_t rank = j;
    for (uint32_t o = 0; o < parts; ++o) {
        if (o == pp) continue;
        const uint32_t* oi = ids[o] + row;
        const float* od = dists[o] + row;
        uint32_t lo = 0, hi = k_in;
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            const uint32_t xid = oi[mid];
            const float xd = od[mid];
            const bool before = xid != PAD_ID && (pair_less(xd, xid, dist, id) || (xd == dist && xid == id && o < pp));
            if (before) lo = mid + 1; else hi = mid;
        }
        rank += lo;
    }
    if (rank < k_out) {
        out_ids[(size_t)q * k_out + rank] = id;
        if (out_dists) out_dists[(size_t)q * k_out + rank] = dist;
    }
}

__global__ void fill_pad2_kernel(uint32_t* ids, float* dists, uint64_t n) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    ids[t] = PAD_ID;
    if (dists) dists[t] = __int_as_float(0x7f800000);
}

struct Nccl {
    void* lib = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool load() {
        if (lib) return true;
        // GBDR_NCCL_LIB names the library to use.  It matters in a process that also loads torch: libraries are shared by
        // soname, so whichever libnccl.so.2 is mapped first serves both, and torch needs the (newer) one it ships with —
        // the Python binding points this variable at that copy (capi.py); a C++ host gets the system library.
        const char* named = getenv("GBDR_NCCL_LIB");
        if (named && *named) lib = dlopen(named, RTLD_NOW | RTLD_GLOBAL);
        if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) return false;
        CommInitAll = (decltype(CommInitAll))dlsym(lib, "ncclCommInitAll");
        CommDestroy = (decltype(CommDestroy))dlsym(lib, "ncclCommDestroy");
        AllGather = (decltype(AllGather))dlsym(lib, "ncclAllGather");
        GroupStart = (decltype(GroupStart))dlsym(lib, "ncclGroupStart");
        GroupEnd = (decltype(GroupEnd))dlsym(lib, "ncclGroupEnd");
        GetErrorString = (decltype(GetErrorString))dlsym(lib, "ncclGetErrorString");
        if (!CommInitAll || !CommDestroy || !AllGather || !GroupStart || !GroupEnd || !GetErrorString) {
            lib = nullptr;
            return false;
        }
        return true;
    }
};
Nccl g_nccl;

}  // namespace

}  // namespace gbdr

using namespace gbdr;

struct gbdr_group {
    int n = 0, mode = GBDR_GROUP_REPLICATED, exchange = GBDR_EXCHANGE_PEER;
    std::vector<int> devs;
    std::vector<gbdr_index*> members;
    std::vector<uint64_t> row_begin;  // sharded: member i owns rows [row_begin[i], row_begin[i + 1])
    bool peer_ok = true;              // the root can map every member's memory
    bool distinct = true;             // no device listed twice (NCCL needs that)
    // one worker thread per member
    std::vector<std::thread> workers;
    std::mutex mu;
    std::condition_variable cv_go, cv_done;
    uint64_t generation = 0;
    int outstanding = 0;
    bool stop = false;
    std::function<int(int)> task;
    std::vector<int> rcs;
    std::vector<std::string> errs;
    // exchange step (sharded)
    std::vector<cudaEvent_t> done;        // member i's results are complete (recorded on its stream)
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    DevBuf ptr_table, merged_ids, merged_dists;   // on the root device
    std::vector<DevBuf> gath_ids, gath_dists;     // NCCL exchange: per member [n x n_q x k]
    std::vector<ncclComm_t> comms;
    std::vector<int32_t*> st_hops, st_dc;         // pinned staging of per-shard hops / dist_calc
    std::vector<size_t> st_cap;

    int run(const std::function<int(int)>& f) {
        {
            std::unique_lock<std::mutex> lk(mu);
            task = f;
            outstanding = n;
            ++generation;
        }
        cv_go.notify_all();
        std::unique_lock<std::mutex> lk(mu);
        cv_done.wait(lk, [&] { return outstanding == 0; });
        for (int i = 0; i < n; ++i)
            if (rcs[i] != GBDR_OK) {
                set_error("device " + std::to_string(devs[i]) + ": " + errs[i]);
                return rcs[i];
            }
        return GBDR_OK;
    }
    void worker(int i) {
        uint64_t seen = 0;
        for (;;) {
            std::function<int(int)> f;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv_go.wait(lk, [&] { return stop || generation != seen; });
                if (stop) return;
                seen = generation;
                f = task;
            }
            cudaSetDevice(devs[i]);
            const int rc = f(i);
            {
                std::unique_lock<std::mutex> lk(mu);
                rcs[i] = rc;
                errs[i] = rc ? gbdr_last_error() : "";
                if (--outstanding == 0) cv_done.notify_all();
            }
        }
    }
};

static void partition_rows(uint64_t n_items, int world, std::vector<uint64_t>& begin) {
    // contiguous, balanced: the first n_items % world members get one extra row (multigpu.partition in the Python binding)
    begin.assign(world + 1, 0);
    const uint64_t base = n_items / world, rem = n_items % world;
    for (int r = 0; r < world; ++r) begin[r + 1] = begin[r] + base + ((uint64_t)r < rem ? 1 : 0);
}

extern "C" int gbdr_group_create(const int* devices, int n_devices, int mode, gbdr_group** out) {
    if (!out) return GBDR_E_INVALID;
    *out = nullptr;
    if (!devices || n_devices < 1 || n_devices > 64 || (mode != GBDR_GROUP_REPLICATED && mode != GBDR_GROUP_SHARDED)) {
        set_error("group_create: need 1..64 devices and a valid mode");
        return GBDR_E_INVALID;
    }
    gbdr_group* g = new gbdr_group();
    for (int i = 0; i < n_devices; ++i)
        for (int j = 0; j < i; ++j)
            if (devices[i] == devices[j]) g->distinct = false;  // allowed (every entry is its own index); no NCCL then
    g->n = n_devices;
    g->mode = mode;
    g->devs.assign(devices, devices + n_devices);
    g->members.assign(n_devices, nullptr);
    g->rcs.assign(n_devices, GBDR_OK);
    g->errs.assign(n_devices, "");
    g->done.assign(n_devices, nullptr);
    g->gath_ids.resize(n_devices);
    g->gath_dists.resize(n_devices);
    g->st_hops.assign(n_devices, nullptr);
    g->st_dc.assign(n_devices, nullptr);
    g->st_cap.assign(n_devices, 0);
    int rc = GBDR_OK;
    for (int i = 0; i < n_devices && rc == GBDR_OK; ++i) {
        rc = gbdr_index_create(devices[i], &g->members[i]);
        if (rc == GBDR_OK && cudaEventCreateWithFlags(&g->done[i], cudaEventDisableTiming) != cudaSuccess) rc = GBDR_E_CUDA;
    }
    if (rc == GBDR_OK) {
        // the root maps every member's memory (NVLink peer access); without it the exchange goes through NCCL or copies
        cudaSetDevice(devices[0]);
        cudaEventCreate(&g->e0);
        cudaEventCreate(&g->e1);
        for (int i = 1; i < n_devices; ++i) {
            int can = 0;
            if (devices[i] == devices[0]) continue;
            cudaDeviceCanAccessPeer(&can, devices[0], devices[i]);
            if (!can) {
                g->peer_ok = false;
                continue;
            }
            const cudaError_t e = cudaDeviceEnablePeerAccess(devices[i], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) g->peer_ok = false;
            cudaGetLastError();
        }
        if (!g->peer_ok) g->exchange = GBDR_EXCHANGE_NCCL;
        // and the members each other's (the build's all-gather by peer copies goes device to device)
        for (int i = 1; i < n_devices; ++i) {
            cudaSetDevice(devices[i]);
            for (int j = 0; j < n_devices; ++j) {
                int can = 0;
                if (devices[j] == devices[i] || cudaDeviceCanAccessPeer(&can, devices[i], devices[j]) != cudaSuccess || !can) continue;
                cudaDeviceEnablePeerAccess(devices[j], 0);
                cudaGetLastError();
            }
        }
    }
    if (rc != GBDR_OK) {
        const std::string msg = gbdr_last_error();
        for (auto* m : g->members)
            if (m) gbdr_index_destroy(m);
        delete g;
        set_error(msg);
        return rc;
    }
    for (int i = 0; i < n_devices; ++i) g->workers.emplace_back([g, i] { g->worker(i); });
    *out = g;
    return GBDR_OK;
}

extern "C" int gbdr_group_destroy(gbdr_group* g) {
    if (!g) return GBDR_OK;
    {
        std::unique_lock<std::mutex> lk(g->mu);
        g->stop = true;
    }
    g->cv_go.notify_all();
    for (auto& t : g->workers) t.join();
    for (int i = 0; i < g->n; ++i) {
        cudaSetDevice(g->devs[i]);
        if (i < (int)g->comms.size() && g->comms[i]) g_nccl.CommDestroy(g->comms[i]);
        g->gath_ids[i].release();
        g->gath_dists[i].release();
        if (g->st_hops[i]) cudaFreeHost(g->st_hops[i]);
        if (g->st_dc[i]) cudaFreeHost(g->st_dc[i]);
        if (g->done[i]) cudaEventDestroy(g->done[i]);
    }
    cudaSetDevice(g->devs[0]);
    g->ptr_table.release();
    g->merged_ids.release();
    g->merged_dists.release();
    if (g->e0) cudaEventDestroy(g->e0);
    if (g->e1) cudaEventDestroy(g->e1);
    for (auto* m : g->members) gbdr_index_destroy(m);
    delete g;
    return GBDR_OK;
}

extern "C" int gbdr_group_size(const gbdr_group* g) { return g ? g->n : 0; }

extern "C" int gbdr_group_member(gbdr_group* g, int i, gbdr_index** out) {
    if (!g || !out || i < 0 || i >= g->n) return GBDR_E_INVALID;
    *out = g->members[i];
    return GBDR_OK;
}

static int ensure_nccl(gbdr_group* g) {
    if (!g->comms.empty()) return GBDR_OK;
    if (!g->distinct) {
        set_error("group: the NCCL exchange needs every device listed once");
        return GBDR_E_STATE;
    }
    if (!g_nccl.load()) {
        set_error("group: libnccl.so.2 could not be loaded; the NCCL exchange is unavailable (use GBDR_EXCHANGE_PEER)");
        return GBDR_E_STATE;
    }
    g->comms.assign(g->n, nullptr);
    const ncclResult_t r = g_nccl.CommInitAll(g->comms.data(), g->n, g->devs.data());
    if (r != ncclSuccess) {
        g->comms.clear();
        set_error(std::string("ncclCommInitAll: ") + g_nccl.GetErrorString(r));
        return GBDR_E_CUDA;
    }
    return GBDR_OK;
}

extern "C" int gbdr_group_set_exchange(gbdr_group* g, int exchange) {
    if (!g || (exchange != GBDR_EXCHANGE_PEER && exchange != GBDR_EXCHANGE_NCCL)) return GBDR_E_INVALID;
    if (exchange == GBDR_EXCHANGE_PEER && !g->peer_ok) {
        set_error("group_set_exchange: the root device cannot map every member's memory (no peer access)");
        return GBDR_E_STATE;
    }
    if (exchange == GBDR_EXCHANGE_NCCL)
        if (int rc = ensure_nccl(g)) return rc;
    g->exchange = exchange;
    return GBDR_OK;
}

extern "C" int gbdr_group_shard_rows(const gbdr_group* g, int i, uint64_t* begin, uint64_t* end) {
    if (!g || i < 0 || i >= g->n || g->row_begin.empty()) return GBDR_E_INVALID;
    if (begin) *begin = g->row_begin[i];
    if (end) *end = g->row_begin[i + 1];
    return GBDR_OK;
}

extern "C" int gbdr_group_set_net(gbdr_group* g, const float* l1, const float* l2, const float* l3, uint32_t d,
                                  uint32_t d_hidden, uint32_t d_hidden2, uint32_t d_low) {
    if (!g) return GBDR_E_INVALID;
    return g->run([&](int i) { return gbdr_index_set_net(g->members[i], l1, l2, l3, d, d_hidden, d_hidden2, d_low); });
}

static int set_rows(gbdr_group* g, const float* rows, uint64_t n, uint32_t d, bool low) {
    if (!g || (!rows && n)) return GBDR_E_INVALID;
    if (g->mode == GBDR_GROUP_SHARDED) {
        std::vector<uint64_t> rb;
        partition_rows(n, g->n, rb);
        if (!g->row_begin.empty() && g->row_begin != rb) {
            set_error("group: base and low-dimensional base must have the same number of rows");
            return GBDR_E_INVALID;
        }
        g->row_begin = rb;
    }
    return g->run([&](int i) -> int {
        const uint64_t b = g->mode == GBDR_GROUP_SHARDED ? g->row_begin[i] : 0;
        const uint64_t e = g->mode == GBDR_GROUP_SHARDED ? g->row_begin[i + 1] : n;
        int rc = low ? gbdr_index_set_low(g->members[i], rows + (size_t)b * d, e - b, d)
                     : gbdr_index_set_base(g->members[i], rows + (size_t)b * d, e - b, d);
        if (rc == GBDR_OK && g->mode == GBDR_GROUP_SHARDED) rc = gbdr_index_set_id_offset(g->members[i], b);
        return rc;
    });
}

extern "C" int gbdr_group_set_base(gbdr_group* g, const float* db, uint64_t n, uint32_t d) { return set_rows(g, db, n, d, false); }
extern "C" int gbdr_group_set_low(gbdr_group* g, const float* db_low, uint64_t n, uint32_t d_low) {
    return set_rows(g, db_low, n, d_low, true);
}

extern "C" int gbdr_group_set_graph(gbdr_group* g, const uint64_t* offsets, const uint32_t* edges, uint64_t n) {
    if (!g) return GBDR_E_INVALID;
    if (g->mode != GBDR_GROUP_REPLICATED) {
        set_error("group_set_graph: a sharded group takes one graph per shard, over local ids (gbdr_group_set_shard_graph)");
        return GBDR_E_STATE;
    }
    return g->run([&](int i) { return gbdr_index_set_graph(g->members[i], offsets, edges, n); });
}

extern "C" int gbdr_group_set_shard_graph(gbdr_group* g, int i, const uint64_t* offsets, const uint32_t* edges, uint64_t n_i) {
    if (!g || i < 0 || i >= g->n) return GBDR_E_INVALID;
    if (g->mode != GBDR_GROUP_SHARDED) {
        set_error("group_set_shard_graph: not a sharded group");
        return GBDR_E_STATE;
    }
    if (!g->row_begin.empty() && n_i != g->row_begin[i + 1] - g->row_begin[i]) {
        set_error("group_set_shard_graph: vertex count differs from the shard's row count");
        return GBDR_E_INVALID;
    }
    return gbdr_index_set_graph(g->members[i], offsets, edges, n_i);
}

// ---------------------------------------------------------------------------------------------------------- search
static int search_replicated(gbdr_group* g, const float* queries, const float* q_low, uint32_t n_q, uint32_t ef, uint32_t k,
                             uint32_t flags, const uint32_t* entry, uint32_t* out_ids, float* out_dists, int32_t* hops,
                             int32_t* dist_calc, double* gpu_seconds) {
    std::vector<uint64_t> qb;
    partition_rows(n_q, g->n, qb);
    std::vector<double> secs(g->n, 0.0);
    int rc = g->run([&](int i) -> int {
        gbdr_index* h = g->members[i];
        const uint64_t b = qb[i], e = qb[i + 1];
        const uint32_t dq = h->d ? h->d : h->net_d, dl = h->d_low;
        return gbdr_search(h, queries ? queries + b * dq : nullptr, q_low ? q_low + b * dl : nullptr, (uint32_t)(e - b), ef, k,
                           flags, entry + b, out_ids + b * k, out_dists ? out_dists + b * k : nullptr, hops ? hops + b : nullptr,
                           dist_calc ? dist_calc + b : nullptr, &secs[i]);
    });
    if (rc) return rc;
    if (gpu_seconds) {
        *gpu_seconds = 0;
        for (double s : secs) *gpu_seconds = std::max(*gpu_seconds, s);
    }
    return GBDR_OK;
}

static int search_sharded(gbdr_group* g, const float* queries, const float* q_low, uint32_t n_q, uint32_t ef, uint32_t k,
                          uint32_t flags, const uint32_t* entry, uint32_t* out_ids, float* out_dists, int32_t* hops,
                          int32_t* dist_calc, double* gpu_seconds) {
    const int N = g->n;
    gbdr_index* root = g->members[0];
    const bool stats = hops || dist_calc;
    uint32_t spill_before = 0;
    for (auto* m : g->members) spill_before += m->spill_min;
    if (g->exchange == GBDR_EXCHANGE_NCCL)
        if (int rc = ensure_nccl(g)) return rc;
    GBDR_CUDA(cudaSetDevice(g->devs[0]));
    GBDR_CUDA(cudaEventRecord(g->e0, root->stream));
    // phase A: every member searches every query on its shard; ids (already global) and distances stay in its HBM
    int rc = g->run([&](int i) -> int {
        gbdr_index* h = g->members[i];
        if (stats && g->st_cap[i] < n_q) {
            if (g->st_hops[i]) cudaFreeHost(g->st_hops[i]);
            if (g->st_dc[i]) cudaFreeHost(g->st_dc[i]);
            g->st_hops[i] = g->st_dc[i] = nullptr;
            GBDR_CUDA(cudaHostAlloc((void**)&g->st_hops[i], (size_t)n_q * 4, cudaHostAllocDefault));
            GBDR_CUDA(cudaHostAlloc((void**)&g->st_dc[i], (size_t)n_q * 4, cudaHostAllocDefault));
            g->st_cap[i] = n_q;
        }
        int r = search_submit_impl(h, queries, q_low, n_q, ef, k, flags, entry + (size_t)i * n_q, nullptr, nullptr,
                                   hops ? g->st_hops[i] : nullptr, dist_calc ? g->st_dc[i] : nullptr, true);
        if (r) {
            if (h->stream && !h->pending) cudaStreamSynchronize(h->stream);
            return r;
        }
        if (g->exchange == GBDR_EXCHANGE_NCCL) {
            if ((r = g->gath_ids[i].ensure((size_t)N * n_q * k * 4)) || (r = g->gath_dists[i].ensure((size_t)N * n_q * k * 4))) return r;
        }
        GBDR_CUDA(cudaEventRecord(g->done[i], h->stream));
        return GBDR_OK;
    });
    if (rc) {
        for (auto* m : g->members)
            if (m->pending) gbdr_search_wait(m, nullptr);
        return rc;
    }
    // phase B: the exchange step, on the root's stream
    GBDR_CUDA(cudaSetDevice(g->devs[0]));
    cudaStream_t rs = root->stream;
    const uint64_t nout = (uint64_t)n_q * k, total = nout * N;
    if ((rc = g->merged_ids.ensure(nout * 4 + 16)) || (rc = g->merged_dists.ensure(nout * 4 + 16))) return rc;
    if (g->exchange == GBDR_EXCHANGE_NCCL) {
        ncclResult_t r = g_nccl.GroupStart();
        for (int i = 0; i < N && r == ncclSuccess; ++i) {
            gbdr_index* h = g->members[i];
            r = g_nccl.AllGather(h->w_out_ids.p, g->gath_ids[i].p, nout, ncclUint32, g->comms[i], h->stream);
            if (r == ncclSuccess) r = g_nccl.AllGather(h->w_out_dists.p, g->gath_dists[i].p, nout, ncclFloat32, g->comms[i], h->stream);
        }
        const ncclResult_t r2 = g_nccl.GroupEnd();
        if (r != ncclSuccess || r2 != ncclSuccess) {
            set_error(std::string("ncclAllGather: ") + g_nccl.GetErrorString(r != ncclSuccess ? r : r2));
            for (auto* m : g->members)
                if (m->pending) gbdr_search_wait(m, nullptr);
            return GBDR_E_CUDA;
        }
        GBDR_CUDA(cudaSetDevice(g->devs[0]));
        if ((rc = launch_merge_topk(g->gath_ids[0].as<uint32_t>(), g->gath_dists[0].as<float>(), (uint32_t)N, n_q, k, k,
                                    g->merged_ids.as<uint32_t>(), g->merged_dists.as<float>(), rs)))
            return rc;
    } else {
        std::vector<const void*> tab(2 * N);
        for (int i = 0; i < N; ++i) {
            tab[i] = g->members[i]->w_out_ids.p;
            tab[N + i] = g->members[i]->w_out_dists.p;
            if (i) GBDR_CUDA(cudaStreamWaitEvent(rs, g->done[i], 0));
        }
        if ((rc = g->ptr_table.ensure(tab.size() * sizeof(void*)))) return rc;
        GBDR_CUDA(cudaMemcpyAsync(g->ptr_table.p, tab.data(), tab.size() * sizeof(void*), cudaMemcpyHostToDevice, rs));
        fill_pad2_kernel<<<(unsigned)((nout + 255) / 256), 256, 0, rs>>>(g->merged_ids.as<uint32_t>(), g->merged_dists.as<float>(), nout);
        GBDR_CHECK_LAUNCH();
        merge_topk_ptrs_kernel<<<(unsigned)((total + 255) / 256), 256, 0, rs>>>(
            g->ptr_table.as<const uint32_t*>(), reinterpret_cast<const float* const*>(g->ptr_table.as<const void*>() + N), (uint32_t)N,
            n_q, k, k, g->merged_ids.as<uint32_t>(), g->merged_dists.as<float>());
        GBDR_CHECK_LAUNCH();
        count_launch(2);
    }
    GBDR_CUDA(cudaMemcpyAsync(out_ids, g->merged_ids.p, nout * 4, cudaMemcpyDeviceToHost, rs));
    if (out_dists) GBDR_CUDA(cudaMemcpyAsync(out_dists, g->merged_dists.p, nout * 4, cudaMemcpyDeviceToHost, rs));
    GBDR_CUDA(cudaEventRecord(g->e1, rs));
    // the members' own waits check their status words (root last: its stream carries the merge and the download)
    int first = GBDR_OK;
    std::string first_err;
    for (int i = N - 1; i >= 0; --i) {
        cudaSetDevice(g->devs[i]);
        const int w = gbdr_search_wait(g->members[i], nullptr);
        if (w && first == GBDR_OK) {
            first = w;
            first_err = "device " + std::to_string(g->devs[i]) + ": " + gbdr_last_error();
        }
    }
    if (first) {
        set_error(first_err);
        return first;
    }
    cudaSetDevice(g->devs[0]);
    GBDR_CUDA(cudaStreamSynchronize(rs));
    {
        // a member whose overflow tables were exhausted re-ran its part with larger ones (gbdr_search_wait), after the
        // merge had read its first attempt: run the whole call again, every member now starts with the large tables
        uint32_t spill_after = 0;
        for (auto* m : g->members) spill_after += m->spill_min;
        if (spill_after != spill_before)
            return search_sharded(g, queries, q_low, n_q, ef, k, flags, entry, out_ids, out_dists, hops, dist_calc, gpu_seconds);
    }
    if (stats)
        for (uint32_t q = 0; q < n_q; ++q) {
            int32_t hs = 0, ds = 0;
            for (int i = 0; i < N; ++i) {
                if (hops) hs += g->st_hops[i][q];
                if (dist_calc) ds += g->st_dc[i][q];
            }
            if (hops) hops[q] = hs;
            if (dist_calc) dist_calc[q] = ds;
        }
    if (gpu_seconds) {
        float ms = 0;
        GBDR_CUDA(cudaEventElapsedTime(&ms, g->e0, g->e1));
        *gpu_seconds = ms * 1e-3;
    }
    return GBDR_OK;
}

extern "C" int gbdr_group_search(gbdr_group* g, const float* queries, const float* q_low, uint32_t n_q, uint32_t ef, uint32_t k,
                                 uint32_t flags, const uint32_t* entry, uint32_t* out_ids, float* out_dists, int32_t* hops,
                                 int32_t* dist_calc, double* gpu_seconds) {
    if (!g || !entry || !out_ids) {
        set_error("group_search: null pointer");
        return GBDR_E_INVALID;
    }
    if (n_q == 0) return GBDR_OK;
    if (g->mode == GBDR_GROUP_REPLICATED || g->n == 1)
        return search_replicated(g, queries, q_low, n_q, ef, k, flags, entry, out_ids, out_dists, hops, dist_calc, gpu_seconds);
    return search_sharded(g, queries, q_low, n_q, ef, k, flags, entry, out_ids, out_dists, hops, dist_calc, gpu_seconds);
}

// ---------------------------------------------------------------------------------------------------- graph build
// The graph build (Python kNN-1k, dim_red/support_func.py:374-384, then hnswlikeGD, search/prepare_graph.cpp:64-74) row-block
// sharded over the group's devices (either mode): the vector matrix goes up once, block i over device i's own PCIe link,
// and is all-gathered over NVLink (NCCL in place, or peer copies); device i then computes the kNN lists and the forward
// prune of its block from HBM; the forward lists (2M ids per row) are gathered on the first device for the reverse pass,
// the one order-dependent step (support_func.h:423-442).
extern "C" int gbdr_group_build_graph(gbdr_group* g, const float* db_low, uint64_t n, uint32_t d_low, uint32_t knn_k, uint32_t M,
                                      int reverse, int need_const_degree, uint64_t* out_offsets, uint32_t* out_edges,
                                      uint32_t* knn_out, double timings[4]) {
    if (!g || !db_low || !out_offsets || !out_edges || d_low < 4 || (d_low % 4) || knn_k == 0 || knn_k > n || M < 2) {
        set_error("group_build_graph: bad argument (d_low must be a multiple of 4, 1 <= knn_k <= n, M >= 2)");
        return GBDR_E_INVALID;
    }
    const int N = g->n;
    if (N == 1)
        return gbdr_build_graph(g->devs[0], db_low, n, d_low, knn_k, M, reverse, need_const_degree, out_offsets, out_edges, knn_out,
                                timings);
    using clk = std::chrono::steady_clock;
    const bool use_nccl = g->distinct && (g->exchange == GBDR_EXCHANGE_NCCL || !g->peer_ok);
    if (use_nccl)
        if (int rc = ensure_nccl(g)) return rc;
    const uint64_t per = (n + N - 1) / N;  // equal blocks (the last one may be short): what an in-place all-gather wants
    auto blk_b = [&](int i) { return std::min<uint64_t>(n, (uint64_t)i * per); };
    auto blk_e = [&](int i) { return std::min<uint64_t>(n, (uint64_t)(i + 1) * per); };
    const uint32_t cap = 2 * M;
    struct Dev {
        float* Y = nullptr;
        uint32_t *K = nullptr, *F = nullptr, *D = nullptr;
        cudaStream_t st = nullptr, copy_st = nullptr;
        cudaEvent_t up = nullptr, pruned = nullptr;
    };
    std::vector<Dev> dv(N);
    uint32_t *rootF = nullptr, *rootD = nullptr, *rootK = nullptr;
    auto cleanup = [&]() {
        for (int i = 0; i < N; ++i) {
            cudaSetDevice(g->devs[i]);
            if (dv[i].st) cudaStreamSynchronize(dv[i].st);
            if (dv[i].copy_st) cudaStreamSynchronize(dv[i].copy_st);
            for (void* q : {(void*)dv[i].Y, (void*)dv[i].K, (void*)dv[i].F, (void*)dv[i].D})
                if (q) cudaFree(q);
            if (dv[i].up) cudaEventDestroy(dv[i].up);
            if (dv[i].pruned) cudaEventDestroy(dv[i].pruned);
            if (dv[i].copy_st) cudaStreamDestroy(dv[i].copy_st);
            if (dv[i].st) cudaStreamDestroy(dv[i].st);
        }
        cudaSetDevice(g->devs[0]);
        for (void* q : {(void*)rootF, (void*)rootD, (void*)rootK})
            if (q) cudaFree(q);
    };
    const auto t0 = clk::now();
    // 1. allocate, upload the own block
    int rc = g->run([&](int i) -> int {
        Dev& x = dv[i];
        const uint64_t rows = blk_e(i) - blk_b(i);
        GBDR_CUDA(cudaStreamCreateWithFlags(&x.st, cudaStreamNonBlocking));
        GBDR_CUDA(cudaStreamCreateWithFlags(&x.copy_st, cudaStreamNonBlocking));
        GBDR_CUDA(cudaEventCreateWithFlags(&x.up, cudaEventDisableTiming));
        GBDR_CUDA(cudaEventCreateWithFlags(&x.pruned, cudaEventDisableTiming));
        GBDR_CUDA(cudaMalloc((void**)&x.Y, (size_t)per * N * d_low * 4 + 16));
        GBDR_CUDA(cudaMalloc((void**)&x.K, (size_t)std::max<uint64_t>(rows, 1) * knn_k * 4 + 16));
        GBDR_CUDA(cudaMalloc((void**)&x.F, (size_t)std::max<uint64_t>(rows, 1) * cap * 4 + 16));
        GBDR_CUDA(cudaMalloc((void**)&x.D, (size_t)std::max<uint64_t>(rows, 1) * 4 + 16));
        if (rows)
            GBDR_CUDA(cudaMemcpyAsync(x.Y + (size_t)blk_b(i) * d_low, db_low + (size_t)blk_b(i) * d_low, (size_t)rows * d_low * 4,
                                      cudaMemcpyHostToDevice, x.st));
        GBDR_CUDA(cudaEventRecord(x.up, x.st));
        return GBDR_OK;
    });
    // 2. all-gather of the vector matrix over NVLink
    if (rc == GBDR_OK) {
        if (use_nccl) {
            ncclResult_t r = g_nccl.GroupStart();
            for (int i = 0; i < N && r == ncclSuccess; ++i)
                r = g_nccl.AllGather(dv[i].Y + (size_t)i * per * d_low, dv[i].Y, (size_t)per * d_low, ncclFloat32, g->comms[i], dv[i].st);
            const ncclResult_t r2 = g_nccl.GroupEnd();
            if (r != ncclSuccess || r2 != ncclSuccess) {
                set_error(std::string("ncclAllGather: ") + g_nccl.GetErrorString(r != ncclSuccess ? r : r2));
                rc = GBDR_E_CUDA;
            }
        } else {
            rc = g->run([&](int i) -> int {
                for (int j = 0; j < N; ++j) {
                    if (j == i || blk_e(j) == blk_b(j)) continue;
                    GBDR_CUDA(cudaStreamWaitEvent(dv[i].st, dv[j].up, 0));
                    GBDR_CUDA(cudaMemcpyPeerAsync(dv[i].Y + (size_t)blk_b(j) * d_low, g->devs[i], dv[j].Y + (size_t)blk_b(j) * d_low,
                                                  g->devs[j], (size_t)(blk_e(j) - blk_b(j)) * d_low * 4, dv[i].st));
                }
                return GBDR_OK;
            });
        }
    }
    if (rc == GBDR_OK)
        rc = g->run([&](int i) -> int {
            GBDR_CUDA(cudaStreamSynchronize(dv[i].st));
            return GBDR_OK;
        });
    const auto t1 = clk::now();
    // 3. kNN lists of the own block
    if (rc == GBDR_OK)
        rc = g->run([&](int i) -> int {
            Dev& x = dv[i];
            const uint64_t b = blk_b(i), e = blk_e(i);
            if (e == b) return GBDR_OK;
            KnnHostSink sink;
            sink.ids = knn_out ? knn_out + (size_t)b * knn_k : nullptr;
            sink.copy_st = x.copy_st;
            std::vector<uint32_t> stale;
            int r = knn_dev_impl(g->devs[i], x.Y, b, e, x.Y, n, d_low, knn_k, x.K, nullptr, (void*)x.st, knn_out ? &sink : nullptr, &stale);
            if (r) return r;
            GBDR_CUDA(cudaStreamSynchronize(x.st));
            if (knn_out) {
                GBDR_CUDA(cudaStreamSynchronize(x.copy_st));
                if (!sink.used || (stale.size() == 1 && stale[0] == UINT32_MAX)) {
                    GBDR_CUDA(cudaMemcpyAsync(sink.ids, x.K, (size_t)(e - b) * knn_k * 4, cudaMemcpyDeviceToHost, x.copy_st));
                } else {
                    for (uint32_t row : stale)
                        GBDR_CUDA(cudaMemcpyAsync(sink.ids + (size_t)row * knn_k, x.K + (size_t)row * knn_k, (size_t)knn_k * 4,
                                                  cudaMemcpyDeviceToHost, x.copy_st));
                }
            }
            return GBDR_OK;
        });
    const auto t2 = clk::now();
    // 4. forward prune of the own block, then the blocks' forward lists (and, for the constant-degree fill, kNN lists)
    //    travel to the first device
    if (rc == GBDR_OK)
        rc = g->run([&](int i) -> int {
            Dev& x = dv[i];
            const uint64_t b = blk_b(i), e = blk_e(i);
            if (e > b) {
                int r = gbdr_gd_prune_dev(g->devs[i], x.K, knn_k, knn_k, b, e, x.Y, n, d_low, M, x.F, x.D, (void*)x.st);
                if (r) return r;
            }
            GBDR_CUDA(cudaEventRecord(x.pruned, x.st));
            return GBDR_OK;
        });
    if (rc == GBDR_OK) {
        cudaSetDevice(g->devs[0]);
        cudaStream_t rs = dv[0].st;
        cudaError_t ce = cudaMalloc((void**)&rootF, (size_t)n * cap * 4 + 16);
        if (ce == cudaSuccess) ce = cudaMalloc((void**)&rootD, (size_t)n * 4 + 16);
        if (ce == cudaSuccess && need_const_degree) ce = cudaMalloc((void**)&rootK, (size_t)n * knn_k * 4 + 16);
        for (int i = 0; i < N && ce == cudaSuccess; ++i) {
            const uint64_t b = blk_b(i), e = blk_e(i);
            if (e == b) continue;
            ce = cudaStreamWaitEvent(rs, dv[i].pruned, 0);
            if (ce == cudaSuccess)
                ce = cudaMemcpyPeerAsync(rootF + (size_t)b * cap, g->devs[0], dv[i].F, g->devs[i], (size_t)(e - b) * cap * 4, rs);
            if (ce == cudaSuccess) ce = cudaMemcpyPeerAsync(rootD + b, g->devs[0], dv[i].D, g->devs[i], (size_t)(e - b) * 4, rs);
            if (ce == cudaSuccess && need_const_degree)
                ce = cudaMemcpyPeerAsync(rootK + (size_t)b * knn_k, g->devs[0], dv[i].K, g->devs[i], (size_t)(e - b) * knn_k * 4, rs);
        }
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(rs);
        if (ce != cudaSuccess) {
            set_error(std::string("group_build_graph: gathering the forward lists: ") + cudaGetErrorString(ce));
            rc = GBDR_E_CUDA;
        }
    }
    const auto t3 = clk::now();
    // 5. reverse pass and output on the first device
    if (rc == GBDR_OK)
        rc = gd_finish(g->devs[0], rootF, rootD, n, M, reverse, need_const_degree, rootK, knn_k, knn_k, out_offsets, out_edges, dv[0].st);
    std::string err = rc ? gbdr_last_error() : "";
    cleanup();
    const auto t4 = clk::now();
    if (rc) {
        set_error(err);
        return rc;
    }
    if (timings) {
        auto sec = [](clk::time_point a, clk::time_point b) { return std::chrono::duration<double>(b - a).count(); };
        timings[0] = sec(t0, t1);
        timings[1] = sec(t1, t2);
        timings[2] = sec(t2, t3);
        timings[3] = sec(t3, t4);
    }
    return GBDR_OK;
}
