// capi.cu — the C ABI of include/gbdr.h: index object, host<->device plumbing, kernel sequencing.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "index.cuh"

namespace gbdr {

// ---- error state ----
static thread_local std::string t_error;
void set_error(const std::string& msg) { t_error = msg; }
std::atomic<uint64_t> g_launches{0};

int check_device(int device) {
    // cudaGetDeviceProperties costs milliseconds: a device that passed once is not asked again (hot entry points
    // such as gbdr_merge_topk_dev come through here on every call)
    static std::atomic<bool> passed[64];
    if (device >= 0 && device < 64 && passed[device].load(std::memory_order_relaxed)) return GBDR_OK;
    int cnt = 0;
    cudaError_t e = cudaGetDeviceCount(&cnt);
    if (e != cudaSuccess || cnt == 0) {
        cudaGetLastError();
        set_error("no CUDA device visible: this library has no CPU fallback");
        return GBDR_E_NO_DEVICE;
    }
    if (device < 0 || device >= cnt) {
        set_error("device index out of range");
        return GBDR_E_INVALID;
    }
    cudaDeviceProp prop;
    GBDR_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        set_error(std::string("device ") + prop.name + " is not sm_100-class; kernels are built for sm_100a only");
        return GBDR_E_NO_DEVICE;
    }
    if (device < 64) passed[device].store(true, std::memory_order_relaxed);
    return GBDR_OK;
}

}  // namespace gbdr

using namespace gbdr;

static int reject_view(gbdr_index* h, const char* who) {
    if (h && h->parent) {
        set_error(std::string(who) + ": a view is read-only; change the parent index");
        return GBDR_E_STATE;
    }
    return GBDR_OK;
}

// ================================================================ misc
extern "C" int gbdr_version(void) { return GBDR_VERSION; }
extern "C" const char* gbdr_last_error(void) { return t_error.c_str(); }
extern "C" uint64_t gbdr_launch_count(void) { return g_launches.load(); }

extern "C" int gbdr_device_count(int* count) {
    if (!count) return GBDR_E_INVALID;
    int c = 0;
    if (cudaGetDeviceCount(&c) != cudaSuccess) {
        cudaGetLastError();
        c = 0;
    }
    *count = c;
    return GBDR_OK;
}

// ================================================================ index
extern "C" int gbdr_index_create(int device, gbdr_index** out) {
    if (!out) return GBDR_E_INVALID;
    *out = nullptr;
    int rc = check_device(device);
    if (rc) return rc;
    GBDR_CUDA(cudaSetDevice(device));
    gbdr_index* h = new gbdr_index();
    h->device = device;
    cudaDeviceProp prop;
    GBDR_CUDA(cudaGetDeviceProperties(&prop, device));
    h->sm_count = prop.multiProcessorCount;
    GBDR_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    for (auto& e : h->ev) GBDR_CUDA(cudaEventCreate(&e));
    for (auto& q : h->ring)
        for (auto& e : q) GBDR_CUDA(cudaEventCreate(&e));
    GBDR_CUDA(cudaHostAlloc((void**)&h->h_status, 16, cudaHostAllocDefault));
    h->h_status[0] = 0;
    *out = h;
    return GBDR_OK;
}

// (re)borrow the parent's resident state; the projection plan (it owns activation workspaces) is per handle
int gbdr::sync_view(gbdr_index* v) {
    gbdr_index* p = v->parent;
    if (!p || v->epoch == p->epoch) return GBDR_OK;
    v->db.borrow(p->db); v->low.borrow(p->low); v->adj.borrow(p->adj); v->aux.borrow(p->aux);
    v->l1.borrow(p->l1); v->l2.borrow(p->l2); v->l3.borrow(p->l3);
    v->n_base = p->n_base; v->n_low = p->n_low; v->n_graph = p->n_graph; v->n_aux = p->n_aux;
    v->d = p->d; v->d_low = p->d_low; v->C = p->C; v->C_low = p->C_low;
    v->adj_stride = p->adj_stride; v->aux_stride = p->aux_stride; v->hops_bound = p->hops_bound; v->llf = p->llf;
    v->net_d = p->net_d; v->dh = p->dh; v->dh2 = p->dh2; v->net_dlow = p->net_dlow; v->has_net = p->has_net;
    v->proj_mode = p->proj_mode;
    v->id_offset = p->id_offset;
    if (v->tc_plan) {
        cudaStreamSynchronize(v->stream);
        project_tc_destroy(v->tc_plan);
        v->tc_plan = nullptr;
    }
    if (v->has_net) {
        int rc = project_tc_prepare(v->l1.as<float>(), v->l2.as<float>(), v->l3.as<float>(), v->net_d, v->dh, v->dh2,
                                    v->net_dlow, v->stream, &v->tc_plan);
        if (rc) return rc;
        GBDR_CUDA(cudaStreamSynchronize(v->stream));
    }
    v->epoch = p->epoch;
    return GBDR_OK;
}

extern "C" int gbdr_index_create_view(gbdr_index* parent, gbdr_index** out) {
    if (!out) return GBDR_E_INVALID;
    *out = nullptr;
    if (!parent) {
        set_error("create_view: null parent");
        return GBDR_E_INVALID;
    }
    if (parent->parent) parent = parent->parent;  // a view of a view is a view of the owner
    gbdr_index* v = nullptr;
    int rc = gbdr_index_create(parent->device, &v);
    if (rc) return rc;
    v->parent = parent;
    v->epoch = ~parent->epoch;
    parent->n_views++;
    if ((rc = sync_view(v))) {
        gbdr_index_destroy(v);
        return rc;
    }
    *out = v;
    return GBDR_OK;
}

extern "C" int gbdr_index_destroy(gbdr_index* h) {
    if (!h) return GBDR_OK;
    for (auto& v : h->helpers)
        if (v) {
            gbdr_index_destroy(v);
            v = nullptr;
        }
    if (h->n_views.load() > 0) {
        set_error("index_destroy: views of this index are still alive; destroy them first");
        return GBDR_E_STATE;
    }
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    if (h->parent) h->parent->n_views--;
    if (h->h_status) cudaFreeHost(h->h_status);
    if (h->graph_exec) cudaGraphExecDestroy(h->graph_exec);
    for (DevBuf* b : {&h->db, &h->low, &h->adj, &h->aux, &h->l1, &h->l2, &h->l3, &h->w_q, &h->w_qlow, &h->w_entry,
                      &h->w_low_ids, &h->w_out_ids, &h->w_out_dists, &h->w_hops, &h->w_dc, &h->w_scanned, &h->w_h1,
                      &h->w_h2, &h->w_status, &h->w_spill})
        b->release();
    if (h->tc_plan) project_tc_destroy(h->tc_plan);
    for (auto& e : h->ev)
        if (e) cudaEventDestroy(e);
    for (auto& q : h->ring)
        for (auto& e : q)
            if (e) cudaEventDestroy(e);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return GBDR_OK;
}

// upload [n x d] host rows keeping only the first (d/4)*4 dims (L2Metric ignores the tail)
static int upload_rows(gbdr_index* h, DevBuf& buf, const float* src, uint64_t n, uint32_t d) {
    const uint32_t d4 = (d / 4) * 4;
    int rc = buf.ensure((size_t)n * d4 * sizeof(float) + 16);
    if (rc) return rc;
    if (n == 0) return GBDR_OK;
    if (d4 == d) {
        GBDR_CUDA(cudaMemcpyAsync(buf.p, src, (size_t)n * d * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    } else {
        GBDR_CUDA(cudaMemcpy2DAsync(buf.p, (size_t)d4 * 4, src, (size_t)d * 4, (size_t)d4 * 4, n,
                                    cudaMemcpyHostToDevice, h->stream));
    }
    GBDR_CUDA(cudaStreamSynchronize(h->stream));
    return GBDR_OK;
}

extern "C" int gbdr_index_set_base(gbdr_index* h, const float* db, uint64_t n, uint32_t d) {
    if (!h || (!db && n) || d < 4) {
        set_error("set_base: null pointer or d < 4");
        return GBDR_E_INVALID;
    }
    if (int vrc = reject_view(h, "set_base")) return vrc;
    GBDR_CUDA(cudaSetDevice(h->device));
    int rc = upload_rows(h, h->db, db, n, d);
    if (rc) return rc;
    h->n_base = n;
    h->d = d;
    h->C = d / 4;
    h->epoch++;
    return GBDR_OK;
}

extern "C" int gbdr_index_set_low(gbdr_index* h, const float* db_low, uint64_t n, uint32_t d_low) {
    if (!h || (!db_low && n) || d_low < 4) {
        set_error("set_low: null pointer or d_low < 4");
        return GBDR_E_INVALID;
    }
    if (int vrc = reject_view(h, "set_low")) return vrc;
    GBDR_CUDA(cudaSetDevice(h->device));
    int rc = upload_rows(h, h->low, db_low, n, d_low);
    if (rc) return rc;
    h->n_low = n;
    h->d_low = d_low;
    h->C_low = d_low / 4;
    h->epoch++;
    return GBDR_OK;
}

// validate a flattened adjacency and upload it as fixed-stride padded rows
static int upload_graph(gbdr_index* h, DevBuf& buf, const char* who, const uint64_t* offsets, const uint32_t* edges,
                        uint64_t n, uint32_t* stride_out) {
    if (!offsets || (!edges && n && offsets[n] > 0)) {
        set_error(std::string(who) + ": null pointer");
        return GBDR_E_INVALID;
    }
    if (n > 0x7fffffffull) {  // the kernels keep a flag in bit 31 of list ids
        set_error(std::string(who) + ": at most 2^31 - 1 vertices per index (shard larger sets)");
        return GBDR_E_INVALID;
    }
    GBDR_CUDA(cudaSetDevice(h->device));
    uint64_t maxdeg = 0;
    for (uint64_t i = 0; i < n; ++i) {
        if (offsets[i + 1] < offsets[i]) {
            set_error(std::string(who) + ": offsets not monotone");
            return GBDR_E_INVALID;
        }
        maxdeg = std::max<uint64_t>(maxdeg, offsets[i + 1] - offsets[i]);
    }
    const uint64_t total = offsets[n];
    for (uint64_t e = 0; e < total; ++e)
        if (edges[e] >= n) {
            set_error(std::string(who) + ": edge target out of range");
            return GBDR_E_INVALID;
        }
    uint32_t stride = (uint32_t)((maxdeg + 31) / 32 * 32);
    if (stride == 0) stride = 32;
    // A neighbour listed twice in one row is skipped by the reference the second time (already visited,
    // search_function.h:25): drop repeats here, keeping first occurrences in order, so that the kernels may rely
    // on the ids of a row being distinct (their visited-set insertion resolves a whole chunk of ids at once).
    std::vector<uint32_t> padded((size_t)n * stride, GBDR_PAD_ID);
    std::vector<uint32_t> stamp(n, 0);
    for (uint64_t i = 0; i < n; ++i) {
        uint32_t* row = padded.data() + (size_t)i * stride;
        uint32_t m = 0;
        for (uint64_t e = offsets[i]; e < offsets[i + 1]; ++e) {
            const uint32_t id = edges[e];
            if (stamp[id] == (uint32_t)i + 1u) continue;
            stamp[id] = (uint32_t)i + 1u;
            row[m++] = id;
        }
    }
    int rc = buf.ensure(padded.size() * 4 + 16);
    if (rc) return rc;
    GBDR_CUDA(cudaMemcpyAsync(buf.p, padded.data(), padded.size() * 4, cudaMemcpyHostToDevice, h->stream));
    GBDR_CUDA(cudaStreamSynchronize(h->stream));
    *stride_out = stride;
    return GBDR_OK;
}

extern "C" int gbdr_index_set_graph(gbdr_index* h, const uint64_t* offsets, const uint32_t* edges, uint64_t n) {
    if (!h) {
        set_error("set_graph: null pointer");
        return GBDR_E_INVALID;
    }
    if (int vrc = reject_view(h, "set_graph")) return vrc;
    uint32_t stride = 0;
    int rc = upload_graph(h, h->adj, "set_graph", offsets, edges, n, &stride);
    if (rc) return rc;
    h->adj_stride = stride;
    h->n_graph = n;
    h->epoch++;
    return GBDR_OK;
}

extern "C" int gbdr_index_set_aux_graph(gbdr_index* h, const uint64_t* offsets, const uint32_t* edges, uint64_t n,
                                        uint32_t hops_bound, int llf) {
    if (!h) {
        set_error("set_aux_graph: null pointer");
        return GBDR_E_INVALID;
    }
    if (int vrc = reject_view(h, "set_aux_graph")) return vrc;
    if (!offsets) {  // clear
        GBDR_CUDA(cudaSetDevice(h->device));
        h->aux.release();
        h->n_aux = 0;
        h->aux_stride = 0;
        h->epoch++;
        return GBDR_OK;
    }
    uint32_t stride = 0;
    int rc = upload_graph(h, h->aux, "set_aux_graph", offsets, edges, n, &stride);
    if (rc) return rc;
    h->aux_stride = stride;
    h->n_aux = n;
    h->hops_bound = hops_bound;
    h->llf = llf ? 1u : 0u;
    h->epoch++;
    return GBDR_OK;
}

extern "C" int gbdr_index_set_net(gbdr_index* h, const float* l1, const float* l2, const float* l3, uint32_t d,
                                  uint32_t d_hidden, uint32_t d_hidden2, uint32_t d_low) {
    if (!h || !l1 || !l2 || !l3 || !d || !d_hidden || !d_hidden2 || !d_low) {
        set_error("set_net: null pointer or zero dimension");
        return GBDR_E_INVALID;
    }
    if (int vrc = reject_view(h, "set_net")) return vrc;
    GBDR_CUDA(cudaSetDevice(h->device));
    const size_t s1 = (size_t)d_hidden * (d + 1), s2 = (size_t)d_hidden2 * (d_hidden + 1),
                 s3 = (size_t)d_low * (d_hidden2 + 1);
    int rc;
    if ((rc = h->l1.ensure(s1 * 4)) || (rc = h->l2.ensure(s2 * 4)) || (rc = h->l3.ensure(s3 * 4))) return rc;
    GBDR_CUDA(cudaMemcpyAsync(h->l1.p, l1, s1 * 4, cudaMemcpyHostToDevice, h->stream));
    GBDR_CUDA(cudaMemcpyAsync(h->l2.p, l2, s2 * 4, cudaMemcpyHostToDevice, h->stream));
    GBDR_CUDA(cudaMemcpyAsync(h->l3.p, l3, s3 * 4, cudaMemcpyHostToDevice, h->stream));
    GBDR_CUDA(cudaStreamSynchronize(h->stream));
    h->net_d = d;
    h->dh = d_hidden;
    h->dh2 = d_hidden2;
    h->net_dlow = d_low;
    h->has_net = true;
    h->epoch++;
    if (h->tc_plan) {
        project_tc_destroy(h->tc_plan);
        h->tc_plan = nullptr;
    }
    rc = project_tc_prepare(h->l1.as<float>(), h->l2.as<float>(), h->l3.as<float>(), d, d_hidden, d_hidden2, d_low,
                            h->stream, &h->tc_plan);
    if (rc) return rc;
    return GBDR_OK;
}

extern "C" int gbdr_index_set_id_offset(gbdr_index* h, uint64_t off) {
    if (!h || off > 0x7fffffffull) return GBDR_E_INVALID;
    if (int vrc = reject_view(h, "set_id_offset")) return vrc;
    h->epoch++;
    h->id_offset = off;
    return GBDR_OK;
}

extern "C" int gbdr_index_set_projection_mode(gbdr_index* h, int mode) {
    if (!h || mode < 0 || mode > 2) return GBDR_E_INVALID;
    if (int vrc = reject_view(h, "set_projection_mode")) return vrc;
    h->epoch++;
    h->proj_mode = mode;
    return GBDR_OK;
}

extern "C" int gbdr_index_stream(gbdr_index* h, void** stream) {
    if (!h || !stream) return GBDR_E_INVALID;
    *stream = (void*)h->stream;
    return GBDR_OK;
}

extern "C" int gbdr_beam_plan_info(uint32_t ef, uint32_t dim, uint64_t n_vertices, int second_graph, uint32_t out[10]) {
    if (!out || ef == 0 || dim < 4) return GBDR_E_INVALID;
    BeamPlan p;
    beam_plan(ef, dim / 4, n_vertices, &p, second_graph != 0);
    out[0] = (uint32_t)p.variant;
    out[1] = p.cap;
    out[2] = p.warps_per_block;
    out[3] = p.blocks_per_sm;
    out[4] = p.smem_per_warp;
    out[5] = p.vis_bytes;
    out[6] = p.hcap;
    out[7] = p.vis_tshift ? 32u - p.vis_tshift : 0u;
    out[8] = p.vis_dbits;
    out[9] = p.blocks_per_sm * (p.smem_per_warp * p.warps_per_block + 1024u);
    return GBDR_OK;
}

extern "C" int gbdr_index_device_ptrs(gbdr_index* h, const float** d_db, const float** d_db_low,
                                      const uint32_t** d_adj, uint32_t* adj_stride) {
    if (!h) return GBDR_E_INVALID;
    if (int vrc = sync_view(h)) return vrc;
    if (d_db) *d_db = h->db.as<float>();
    if (d_db_low) *d_db_low = h->low.as<float>();
    if (d_adj) *d_adj = h->adj.as<uint32_t>();
    if (adj_stride) *adj_stride = h->adj_stride;
    return GBDR_OK;
}

// (CUDA graph of a repeated host-buffer call: see submit_body_or_graph below.  Anything else that may resize the handle's
//  workspaces drops it.)
static void drop_graph(gbdr_index* h) {
    if (h->graph_exec) cudaGraphExecDestroy(h->graph_exec);
    h->graph_exec = nullptr;
    h->graph_key_valid = false;
}
static int submit_body_or_graph(gbdr_index* h, const float* queries, const float* q_low, uint32_t n_q, uint32_t ef, uint32_t k,
                                uint32_t flags, const uint32_t* entry, uint32_t* out_ids, float* out_dists, int32_t* hops,
                                int32_t* dist_calc, bool on_device);

// ================================================================ projection
static int project_on_stream(gbdr_index* h, const float* d_q, uint32_t ldq, uint32_t n_q, float* d_out,
                             uint32_t ld_out, cudaStream_t st) {
    if (!h->has_net) {
        set_error("projection requested but no net was set (gbdr_index_set_net)");
        return GBDR_E_STATE;
    }
    if (h->proj_mode == GBDR_PROJ_FP32 || h->tc_plan == nullptr) {
        int rc;
        if ((rc = h->w_h1.ensure((size_t)n_q * h->dh * 4)) || (rc = h->w_h2.ensure((size_t)n_q * h->dh2 * 4))) return rc;
        return launch_project_fp32(d_q, ldq, n_q, h->l1.as<float>(), h->l2.as<float>(), h->l3.as<float>(), h->net_d,
                                   h->dh, h->dh2, h->net_dlow, h->w_h1.as<float>(), h->w_h2.as<float>(), d_out, ld_out,
                                   st);
    }
    return launch_project_tc(h->tc_plan, d_q, ldq, n_q, d_out, ld_out, h->proj_mode == GBDR_PROJ_TF32, st);
}

extern "C" int gbdr_project_dev(gbdr_index* h, const float* d_queries, uint32_t n_q, float* d_q_low, void* stream) {
    if (!h || !d_queries || !d_q_low) return GBDR_E_INVALID;
    GBDR_CUDA(cudaSetDevice(h->device));
    if (int vrc = sync_view(h)) return vrc;
    drop_graph(h);
    return project_on_stream(h, d_queries, h->net_d, n_q, d_q_low, h->net_dlow, (cudaStream_t)stream);
}

extern "C" int gbdr_project(gbdr_index* h, const float* queries, uint32_t n_q, float* q_low) {
    if (!h || !queries || !q_low) return GBDR_E_INVALID;
    if (int vrc = sync_view(h)) return vrc;
    if (!h->has_net) {
        set_error("projection requested but no net was set (gbdr_index_set_net)");
        return GBDR_E_STATE;
    }
    GBDR_CUDA(cudaSetDevice(h->device));
    drop_graph(h);
    int rc;
    if ((rc = h->w_q.ensure((size_t)n_q * h->net_d * 4 + 16)) || (rc = h->w_qlow.ensure((size_t)n_q * h->net_dlow * 4 + 16)))
        return rc;
    GBDR_CUDA(cudaMemcpyAsync(h->w_q.p, queries, (size_t)n_q * h->net_d * 4, cudaMemcpyHostToDevice, h->stream));
    rc = project_on_stream(h, h->w_q.as<float>(), h->net_d, n_q, h->w_qlow.as<float>(), h->net_dlow, h->stream);
    if (rc) return rc;
    GBDR_CUDA(cudaMemcpyAsync(q_low, h->w_qlow.p, (size_t)n_q * h->net_dlow * 4, cudaMemcpyDeviceToHost, h->stream));
    GBDR_CUDA(cudaStreamSynchronize(h->stream));
    return GBDR_OK;
}

// ================================================================ search
// d_q: original queries (stride ldq floats), d_qlow: low-dim queries (stride ldql) or null -> project
int gbdr::search_on_stream(gbdr_index* h, const float* d_q, uint32_t ldq, const float* d_qlow, uint32_t ldql,
                            uint32_t n_q, uint32_t ef, uint32_t k, uint32_t flags, const uint32_t* d_entry,
                            uint32_t* d_out_ids, float* d_out_dists, int32_t* d_hops, int32_t* d_dc,
                            int32_t* d_scanned, cudaStream_t st, bool timed) {
    const bool plain = flags & GBDR_SEARCH_PLAIN;
    const bool rerank = (flags & GBDR_SEARCH_RERANK) && !plain;
    if (ef == 0 || k == 0 || k > ef) {
        set_error("search: need 1 <= k <= ef");
        return GBDR_E_INVALID;
    }
    if (ef > 0x3fffffffu) return GBDR_E_INVALID;
    if (!h->adj.p) {
        set_error("search: no graph set");
        return GBDR_E_STATE;
    }
    if (plain || rerank) {
        if (!h->db.p || h->n_base != h->n_graph) {
            set_error("search: base vectors missing or size differs from the graph");
            return GBDR_E_STATE;
        }
        if (!d_q) {
            set_error("search: original-dimension queries required");
            return GBDR_E_INVALID;
        }
    }
    if (!plain && (!h->low.p || h->n_low != h->n_graph)) {
        set_error("search: low-dimensional vectors missing or size differs from the graph");
        return GBDR_E_STATE;
    }
    const bool second = flags & GBDR_SEARCH_SECOND_GRAPH;
    if (second && (!h->aux.p || h->n_aux != h->n_graph)) {
        set_error("search: GBDR_SEARCH_SECOND_GRAPH needs an auxiliary graph of the main graph's size (gbdr_index_set_aux_graph)");
        return GBDR_E_STATE;
    }
    if (n_q == 0) return GBDR_OK;
    h->timed = timed;
    int rc;
    cudaEvent_t* ev = h->ring[h->ring_pos % gbdr_index::RING];
    if (timed) GBDR_CUDA(cudaEventRecord(ev[0], st));
    // ---- projection ----
    if (!plain && !d_qlow) {
        if (!d_q) {
            set_error("search: neither q_low nor queries given");
            return GBDR_E_INVALID;
        }
        if (!h->has_net || h->net_d != h->d || h->net_dlow != h->d_low) {
            set_error("search: q_low == NULL needs a net whose d/d_low match the index");
            return GBDR_E_STATE;
        }
        if ((rc = h->w_qlow.ensure((size_t)n_q * h->net_dlow * 4 + 16))) return rc;
        if (ldq != h->net_d) {
            set_error("search: projection needs unpadded query rows");
            return GBDR_E_INVALID;
        }
        rc = project_on_stream(h, d_q, ldq, n_q, h->w_qlow.as<float>(), h->net_dlow, st);
        if (rc) return rc;
        d_qlow = h->w_qlow.as<float>();
        ldql = h->net_dlow;
    }
    if (timed) GBDR_CUDA(cudaEventRecord(ev[1], st));

    // ---- beam search ----
    BeamParams p;
    memset(&p, 0, sizeof(p));
    if (plain) {
        p.q = d_q; p.q_stride = ldq; p.db = h->db.as<float>(); p.row_stride = h->C * 4; p.C = h->C;
    } else {
        p.q = d_qlow; p.q_stride = ldql; p.db = h->low.as<float>(); p.row_stride = h->C_low * 4; p.C = h->C_low;
    }
    if (p.q_stride % 4 != 0) {
        set_error("search: query rows must be 16-byte aligned (dimension multiple of 4) in device memory");
        return GBDR_E_INVALID;
    }
    p.adj = h->adj.as<uint32_t>();
    p.adj_stride = h->adj_stride;
    p.entry = d_entry;
    p.n_q = n_q;
    p.ef = ef;
    BeamPlan plan;
    if (second) {
        p.aux_adj = h->aux.as<uint32_t>();
        p.aux_stride = h->aux_stride;
        p.hops_bound = h->hops_bound;
        p.llf = h->llf;
    }
    beam_plan(ef, p.C, h->n_graph, &plan, second);
    const uint32_t wpb = plan.warps_per_block;
    uint32_t spill_log = 11;
    while (spill_log < SPILL_LOG_MAX && (1u << spill_log) < 2u * (12u * ef + 200u)) ++spill_log;
    if (const char* e = getenv("GBDR_BEAM_SPILL_LOG")) spill_log = std::max(6, atoi(e));  // tests: start tiny, force the growth
    spill_log = std::min(std::max(spill_log, h->spill_min), SPILL_LOG_MAX);
    p.spill_cap = 1u << spill_log;
    p.spill_shift = 32 - spill_log;
    p.n_vertices = (uint32_t)h->n_graph;
    uint32_t blocks = std::min<uint32_t>((n_q + wpb - 1) / wpb, (uint32_t)h->sm_count * plan.blocks_per_sm);
    if ((rc = h->w_spill.ensure((size_t)blocks * wpb * p.spill_cap * 4))) return rc;
    if ((rc = h->w_status.ensure(64))) return rc;
    p.spill = h->w_spill.as<uint32_t>();
    p.status = h->w_status.as<uint32_t>();
    GBDR_CUDA(cudaMemsetAsync(h->w_status.p, 0, 8, st));
    p.hops = d_hops;
    p.dist_calc = d_dc;
    p.scanned = d_scanned;
    if (rerank) {
        if ((rc = h->w_low_ids.ensure((size_t)n_q * ef * 4))) return rc;
        p.k = ef;
        p.out_ids = h->w_low_ids.as<uint32_t>();
        p.out_dists = nullptr;
        p.id_offset = 0;
        p.dist_calc_bias = (int32_t)ef;  // search_function.h:164
    } else {
        p.k = k;
        p.out_ids = d_out_ids;
        p.out_dists = d_out_dists;
        p.id_offset = (uint32_t)h->id_offset;
        p.dist_calc_bias = 0;
    }
    rc = launch_beam(p, plan, blocks, st);
    if (rc) return rc;
    if (timed) GBDR_CUDA(cudaEventRecord(ev[2], st));

    // ---- re-rank ----
    if (rerank) {
        RerankParams r;
        memset(&r, 0, sizeof(r));
        r.queries = d_q; r.q_stride = ldq;
        r.db = h->db.as<float>(); r.row_stride = h->C * 4; r.C = h->C;
        r.cand = h->w_low_ids.as<uint32_t>(); r.m = ef; r.k = k; r.n_q = n_q;
        r.id_offset = (uint32_t)h->id_offset;
        r.out_ids = d_out_ids; r.out_dists = d_out_dists;
        if (ldq % 4 != 0) {
            set_error("search: query rows must be 16-byte aligned in device memory");
            return GBDR_E_INVALID;
        }
        rc = launch_rerank(r, st);
        if (rc) return rc;
    }
    if (timed) {
        GBDR_CUDA(cudaEventRecord(ev[3], st));
        h->ring_pos++;
    }
    return GBDR_OK;
}

extern "C" int gbdr_search_dev(gbdr_index* h, const float* d_queries, const float* d_q_low, uint32_t n_q, uint32_t ef,
                               uint32_t k, uint32_t flags, const uint32_t* d_entry, uint32_t* d_out_ids,
                               float* d_out_dists, int32_t* d_hops, int32_t* d_dist_calc, int32_t* d_scanned,
                               void* stream) {
    if (!h || !d_entry || !d_out_ids) return GBDR_E_INVALID;
    GBDR_CUDA(cudaSetDevice(h->device));
    if (int vrc = sync_view(h)) return vrc;
    drop_graph(h);
    if ((h->d % 4) || (!(flags & GBDR_SEARCH_PLAIN) && (h->d_low % 4))) {
        set_error("search_dev: dimensions must be multiples of 4 for device-resident queries");
        return GBDR_E_INVALID;
    }
    return search_on_stream(h, d_queries, h->d, d_q_low, h->d_low, n_q, ef, k, flags, d_entry, d_out_ids, d_out_dists,
                            d_hops, d_dist_calc, d_scanned, (cudaStream_t)stream, true);
}

// copy [n x d] host rows to device with row length (d/4)*4
static int h2d_rows(void* dst, const float* src, uint64_t n, uint32_t d, cudaStream_t st) {
    const uint32_t d4 = (d / 4) * 4;
    if (d4 == d) {
        GBDR_CUDA(cudaMemcpyAsync(dst, src, (size_t)n * d * 4, cudaMemcpyHostToDevice, st));
    } else {
        GBDR_CUDA(cudaMemcpy2DAsync(dst, (size_t)d4 * 4, src, (size_t)d * 4, (size_t)d4 * 4, n, cudaMemcpyHostToDevice, st));
    }
    return GBDR_OK;
}

int gbdr::search_submit_impl(gbdr_index* h, const float* queries, const float* q_low, uint32_t n_q, uint32_t ef,
                             uint32_t k, uint32_t flags, const uint32_t* entry, uint32_t* out_ids, float* out_dists,
                             int32_t* hops, int32_t* dist_calc, bool results_stay_on_device) {
    if (!h || !entry || (!out_ids && !results_stay_on_device)) {
        set_error("search: null pointer");
        return GBDR_E_INVALID;
    }
    if (h->pending) {
        set_error("search_submit: a call is already in flight on this handle (gbdr_search_wait it, or use a view)");
        return GBDR_E_STATE;
    }
    GBDR_CUDA(cudaSetDevice(h->device));
    if (int vrc = sync_view(h)) return vrc;
    h->h_status[0] = 0;
    if (n_q == 0) {
        GBDR_CUDA(cudaEventRecord(h->ev[4], h->stream));
        GBDR_CUDA(cudaEventRecord(h->ev[5], h->stream));
        h->pending = true;
        return GBDR_OK;
    }
    if (k == 0 || k > ef) {
        set_error("search: need 1 <= k <= ef");
        return GBDR_E_INVALID;
    }
    const bool plain = flags & GBDR_SEARCH_PLAIN;
    const bool rerank = (flags & GBDR_SEARCH_RERANK) && !plain;
    const bool need_q = plain || rerank || (!q_low);
    if (need_q && !queries) {
        set_error("search: original-dimension queries required for this mode");
        return GBDR_E_INVALID;
    }
    for (uint32_t i = 0; i < n_q; ++i)
        if (entry[i] >= h->n_graph) {
            set_error("search: entry[" + std::to_string(i) + "] = " + std::to_string(entry[i]) + " is not a vertex of the graph (" +
                      std::to_string(h->n_graph) + " vertices)");
            return GBDR_E_INVALID;
        }
    h->call = {queries, q_low, n_q, ef, k, flags, entry, out_ids, out_dists, hops, dist_calc, results_stay_on_device};
    cudaStream_t st = h->stream;
    GBDR_CUDA(cudaEventRecord(h->ev[4], st));
    int rc = submit_body_or_graph(h, queries, q_low, n_q, ef, k, flags, entry, out_ids, out_dists, hops, dist_calc,
                                  results_stay_on_device);
    if (rc) return rc;
    GBDR_CUDA(cudaEventRecord(h->ev[5], st));
    h->pending = true;
    return GBDR_OK;
}

// the stream operations of one host-buffer call: uploads, projection, walk, re-rank, downloads
static int submit_body(gbdr_index* h, const float* queries, const float* q_low, uint32_t n_q, uint32_t ef, uint32_t k,
                       uint32_t flags, const uint32_t* entry, uint32_t* out_ids, float* out_dists, int32_t* hops,
                       int32_t* dist_calc, bool timed) {
    cudaStream_t st = h->stream;
    int rc;
    const bool plain = flags & GBDR_SEARCH_PLAIN;
    const bool rerank = (flags & GBDR_SEARCH_RERANK) && !plain;
    const bool need_q = plain || rerank || (!q_low);
    const float* d_q = nullptr;
    const float* d_ql = nullptr;
    uint32_t ldq = 0, ldql = 0;
    const bool project = !plain && !q_low;
    if (need_q) {
        // projection needs the unpadded rows; the distance kernels need 16-byte aligned rows
        const uint32_t dd = h->d ? h->d : h->net_d;
        if (project && (dd % 4 != 0) && (plain || rerank)) {
            set_error("search: d % 4 != 0 with on-the-fly projection and re-rank is not supported");
            return GBDR_E_INVALID;
        }
        if ((rc = h->w_q.ensure((size_t)n_q * dd * 4 + 16))) return rc;
        if (project) {
            // (splitting this upload into chunks projected as they land was measured: the extra launches cost the
            // synchronous caller more than the overlap returns, 1.14 vs 1.00 ms per 10k-query call)
            GBDR_CUDA(cudaMemcpyAsync(h->w_q.p, queries, (size_t)n_q * dd * 4, cudaMemcpyHostToDevice, st));
            ldq = dd;
        } else {
            if ((rc = h2d_rows(h->w_q.p, queries, n_q, dd, st))) return rc;
            ldq = (dd / 4) * 4;
        }
        d_q = h->w_q.as<float>();
    }
    if (!plain && q_low) {
        if ((rc = h->w_qlow.ensure((size_t)n_q * h->d_low * 4 + 16))) return rc;
        if ((rc = h2d_rows(h->w_qlow.p, q_low, n_q, h->d_low, st))) return rc;
        d_ql = h->w_qlow.as<float>();
        ldql = h->C_low * 4;
    }
    if ((rc = h->w_entry.ensure((size_t)n_q * 4)) || (rc = h->w_out_ids.ensure((size_t)n_q * k * 4)) ||
        (rc = h->w_out_dists.ensure((size_t)n_q * k * 4)) || (rc = h->w_hops.ensure((size_t)n_q * 4)) ||
        (rc = h->w_dc.ensure((size_t)n_q * 4)) || (rc = h->w_scanned.ensure((size_t)n_q * 4)))
        return rc;
    GBDR_CUDA(cudaMemcpyAsync(h->w_entry.p, entry, (size_t)n_q * 4, cudaMemcpyHostToDevice, st));
    rc = search_on_stream(h, d_q, ldq, d_ql, ldql, n_q, ef, k, flags, h->w_entry.as<uint32_t>(),
                          h->w_out_ids.as<uint32_t>(), h->w_out_dists.as<float>(), h->w_hops.as<int32_t>(),
                          h->w_dc.as<int32_t>(), h->w_scanned.as<int32_t>(), st, timed);
    if (rc) return rc;
    if (out_ids) GBDR_CUDA(cudaMemcpyAsync(out_ids, h->w_out_ids.p, (size_t)n_q * k * 4, cudaMemcpyDeviceToHost, st));
    if (out_dists) GBDR_CUDA(cudaMemcpyAsync(out_dists, h->w_out_dists.p, (size_t)n_q * k * 4, cudaMemcpyDeviceToHost, st));
    if (hops) GBDR_CUDA(cudaMemcpyAsync(hops, h->w_hops.p, (size_t)n_q * 4, cudaMemcpyDeviceToHost, st));
    if (dist_calc) GBDR_CUDA(cudaMemcpyAsync(dist_calc, h->w_dc.p, (size_t)n_q * 4, cudaMemcpyDeviceToHost, st));
    GBDR_CUDA(cudaMemcpyAsync(h->h_status, h->w_status.p, 4, cudaMemcpyDeviceToHost, st));
    return GBDR_OK;
}

// is `p` (a host buffer of the call, or null) page-locked?  (pageable copies are staged by the runtime: not for a graph)
static bool pinned_or_null(const void* p) {
    if (!p) return true;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}

// First call of a shape: plain.  Second identical call (same buffers, same index state): captured into a graph and
// launched.  From then on: one cudaGraphLaunch.  GBDR_SEARCH_GRAPH=0 keeps every call on the plain path.
static int submit_body_or_graph(gbdr_index* h, const float* queries, const float* q_low, uint32_t n_q, uint32_t ef, uint32_t k,
                                uint32_t flags, const uint32_t* entry, uint32_t* out_ids, float* out_dists, int32_t* hops,
                                int32_t* dist_calc, bool on_device) {
    static const bool graphs = [] {
        const char* e = getenv("GBDR_SEARCH_GRAPH");
        return !(e && *e == '0');
    }();
    cudaStream_t st = h->stream;
    gbdr_index::GraphKey key = {queries, q_low, entry, out_ids, out_dists, hops, dist_calc, n_q, ef, k, flags, h->spill_min,
                                h->parent ? h->parent->epoch : h->epoch, h->proj_mode, on_device ? 1 : 0};
    const bool same = h->graph_key_valid && memcmp(&key, &h->graph_key, sizeof(key)) == 0;
    // (checked at every call, not only when capturing: an address may have been freed and handed out again as pageable memory)
    const bool pinned = graphs && same && !h->graph_off && pinned_or_null(queries) && pinned_or_null(q_low) &&
                        pinned_or_null(entry) && pinned_or_null(out_ids) && pinned_or_null(out_dists) && pinned_or_null(hops) &&
                        pinned_or_null(dist_calc);
    if (pinned && h->graph_exec) {
        GBDR_CUDA(cudaGraphLaunch(h->graph_exec, st));
        count_launch(h->graph_launches);
        return GBDR_OK;
    }
    if (!same || !pinned) drop_graph(h);
    if (pinned) {
        // (the previous, plain call of this shape sized every workspace: nothing below allocates)
        const uint64_t launches0 = g_launches.load(std::memory_order_relaxed);
        const uint64_t ring0 = h->ring_pos;
        cudaGraph_t graph = nullptr;
        bool ok = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
        if (ok) {
            const int brc = submit_body(h, queries, q_low, n_q, ef, k, flags, entry, out_ids, out_dists, hops, dist_calc, false);
            const cudaError_t ee = cudaStreamEndCapture(st, &graph);
            ok = brc == GBDR_OK && ee == cudaSuccess && graph != nullptr;
        }
        if (ok) ok = cudaGraphInstantiate(&h->graph_exec, graph, 0) == cudaSuccess;
        if (graph) cudaGraphDestroy(graph);
        h->ring_pos = ring0;
        if (ok) {
            h->graph_launches = g_launches.load(std::memory_order_relaxed) - launches0;
            GBDR_CUDA(cudaGraphLaunch(h->graph_exec, st));
            return GBDR_OK;
        }
        // capture is not available here (driver, a call inside that cannot be captured): never try again on this handle
        cudaGetLastError();
        g_launches.store(launches0, std::memory_order_relaxed);
        if (h->graph_exec) cudaGraphExecDestroy(h->graph_exec);
        h->graph_exec = nullptr;
        h->graph_off = true;
    }
    const int rc = submit_body(h, queries, q_low, n_q, ef, k, flags, entry, out_ids, out_dists, hops, dist_calc, true);
    if (rc == GBDR_OK) {
        h->graph_key = key;
        h->graph_key_valid = true;
    }
    return rc;
}

extern "C" int gbdr_search_submit(gbdr_index* h, const float* queries, const float* q_low, uint32_t n_q, uint32_t ef,
                                  uint32_t k, uint32_t flags, const uint32_t* entry, uint32_t* out_ids,
                                  float* out_dists, int32_t* hops, int32_t* dist_calc) {
    const int rc = search_submit_impl(h, queries, q_low, n_q, ef, k, flags, entry, out_ids, out_dists, hops, dist_calc, false);
    // a failed submit may have enqueued copies that read the caller's buffers: drain them before handing control back
    if (rc != GBDR_OK && h && h->stream && !h->pending) cudaStreamSynchronize(h->stream);
    return rc;
}

extern "C" int gbdr_search_wait(gbdr_index* h, double* gpu_seconds) {
    if (!h) return GBDR_E_INVALID;
    if (!h->pending) {
        set_error("search_wait: nothing in flight on this handle");
        return GBDR_E_STATE;
    }
    GBDR_CUDA(cudaSetDevice(h->device));
    h->pending = false;
    GBDR_CUDA(cudaStreamSynchronize(h->stream));
    const uint32_t status[1] = {h->h_status[0]};
    if (gpu_seconds) {
        float ms = 0;
        GBDR_CUDA(cudaEventElapsedTime(&ms, h->ev[4], h->ev[5]));
        *gpu_seconds = ms * 1e-3;
    }
    if (status[0] & BEAM_ST_WATCHDOG) {
        set_error("search: internal loop watchdog tripped (status " + std::to_string(status[0]) + "): this is a bug");
        return GBDR_E_CUDA;
    }
    if (status[0] & BEAM_ST_BAD_ENTRY) {
        set_error("search: an entry vertex id is not a vertex of the graph");
        return GBDR_E_INVALID;
    }
    if ((status[0] & BEAM_ST_VISITED_FULL) && h->spill_min < SPILL_LOG_MAX) {
        // some query outgrew the overflow tables this beam width normally needs: from now on this handle uses the
        // largest ones, and the call (its buffers are still the caller's to keep valid) runs again
        h->spill_min = SPILL_LOG_MAX;
        const gbdr_index::Call c = h->call;
        int rc = search_submit_impl(h, c.queries, c.q_low, c.n_q, c.ef, c.k, c.flags, c.entry, c.out_ids, c.out_dists, c.hops,
                                    c.dist_calc, c.on_device);
        if (rc) {
            if (h->stream && !h->pending) cudaStreamSynchronize(h->stream);
            return rc;
        }
        return gbdr_search_wait(h, gpu_seconds);
    }
    if (status[0] & (BEAM_ST_VISITED_FULL | BEAM_ST_TIE_OVERFLOW)) {
        set_error(status[0] & BEAM_ST_VISITED_FULL
                      ? "search: a query exhausted the visited-set capacity (ef too large for this build)"
                      : "search: more exact distance ties at the beam boundary than the list slack holds");
        return GBDR_E_CAPACITY;
    }
    return GBDR_OK;
}

// The blocking call can pipeline itself: the batch is cut into `depth` consecutive parts that run on the handle and on
// private views of it (own stream + workspaces, same resident data), so that the upload of part j + 1 and the download
// of part j - 1 overlap the kernels of part j.  Results are those of one call (queries are independent).  Measured on a
// B200 at the bench's operating point (run r3g: GBDR_SEARCH_SPLIT = 1 / 2 / 3 / 4 -> 11.0 / 10.7 / 10.7 / 10.7 M QPS
// through the blocking call): a 10 000-query batch is two waves of the search kernel, a part of it still pays the full
// ~0.3 ms latency of a walk, and what the split hides (0.1 ms of upload) it gives back in launches — so the default is
// one part; GBDR_SEARCH_SPLIT = 2 .. 4 turns the split on (larger batches, slower links).
static uint32_t split_depth(const gbdr_index* h, uint32_t n_q) {
    static const int forced = [] {
        const char* e = getenv("GBDR_SEARCH_SPLIT");
        return e && *e ? atoi(e) : 0;
    }();
    uint32_t depth = forced > 0 ? (uint32_t)forced : 1u;
    depth = std::min<uint32_t>(depth, gbdr_index::MAX_HELPERS + 1);
    // a part must still fill the GPU once (one wave of resident queries), or the split only adds launches
    const uint32_t min_part = (uint32_t)h->sm_count * 32u;
    while (depth > 1 && n_q / depth < min_part) --depth;
    return depth;
}

extern "C" int gbdr_search(gbdr_index* h, const float* queries, const float* q_low, uint32_t n_q, uint32_t ef,
                           uint32_t k, uint32_t flags, const uint32_t* entry, uint32_t* out_ids, float* out_dists,
                           int32_t* hops, int32_t* dist_calc, double* gpu_seconds) {
    if (!h) return GBDR_E_INVALID;
    const uint32_t depth = split_depth(h, n_q);
    if (depth <= 1) {
        int rc = gbdr_search_submit(h, queries, q_low, n_q, ef, k, flags, entry, out_ids, out_dists, hops, dist_calc);
        if (rc) return rc;
        return gbdr_search_wait(h, gpu_seconds);
    }
    if (h->pending) {
        set_error("search: a call is already in flight on this handle (gbdr_search_wait it, or use a view)");
        return GBDR_E_STATE;
    }
    if (int vrc = sync_view(h)) return vrc;
    gbdr_index* part_h[gbdr_index::MAX_HELPERS + 1] = {h};
    for (uint32_t j = 1; j < depth; ++j) {
        if (!h->helpers[j - 1]) {
            int rc = gbdr_index_create_view(h, &h->helpers[j - 1]);
            if (rc) return rc;
            h->helpers[j - 1]->internal = true;
        }
        part_h[j] = h->helpers[j - 1];
    }
    const uint32_t dq = h->d ? h->d : h->net_d, dl = h->d_low;
    const uint32_t per = ((n_q + depth - 1) / depth + 31u) & ~31u;
    int rc = GBDR_OK;
    uint32_t submitted = 0;
    for (uint32_t j = 0; j < depth && rc == GBDR_OK; ++j) {
        const uint32_t b = std::min(n_q, j * per), e = std::min(n_q, b + per);
        rc = gbdr_search_submit(part_h[j], queries ? queries + (size_t)b * dq : nullptr, q_low ? q_low + (size_t)b * dl : nullptr,
                                e - b, ef, k, flags, entry + b, out_ids + (size_t)b * k,
                                out_dists ? out_dists + (size_t)b * k : nullptr, hops ? hops + b : nullptr,
                                dist_calc ? dist_calc + b : nullptr);
        if (rc == GBDR_OK) ++submitted;
    }
    std::string first_error = rc ? gbdr_last_error() : "";
    for (uint32_t j = 0; j < submitted; ++j) {
        const int wrc = gbdr_search_wait(part_h[j], nullptr);
        if (wrc && rc == GBDR_OK) {
            rc = wrc;
            first_error = gbdr_last_error();
        }
    }
    if (rc) {
        set_error(first_error);
        return rc;
    }
    if (gpu_seconds) {  // first part's start to last part's end (events of different streams of one device)
        float ms = 0;
        GBDR_CUDA(cudaEventElapsedTime(&ms, h->ev[4], part_h[depth - 1]->ev[5]));
        *gpu_seconds = ms * 1e-3;
    }
    return GBDR_OK;
}

extern "C" int gbdr_index_status(gbdr_index* h, uint32_t* flags) {
    if (!h || !flags) return GBDR_E_INVALID;
    *flags = 0;
    if (!h->w_status.p) return GBDR_OK;
    GBDR_CUDA(cudaSetDevice(h->device));
    GBDR_CUDA(cudaDeviceSynchronize());
    GBDR_CUDA(cudaMemcpy(flags, h->w_status.p, 4, cudaMemcpyDeviceToHost));
    // a gbdr_search_dev caller that sees bit1 repeats its call: the next one gets the largest overflow tables
    if ((*flags & BEAM_ST_VISITED_FULL) && h->spill_min < SPILL_LOG_MAX) h->spill_min = SPILL_LOG_MAX;
    return GBDR_OK;
}

extern "C" int gbdr_kernel_ms(gbdr_index* h, uint32_t last_n, float* project_ms, float* search_ms,
                              float* rerank_ms) {
    if (!h) return GBDR_E_INVALID;
    if (h->ring_pos == 0) {
        set_error("no timed search on this handle yet");
        return GBDR_E_STATE;
    }
    GBDR_CUDA(cudaSetDevice(h->device));
    if (last_n == 0) last_n = 1;
    if (last_n > h->ring_pos) last_n = (uint32_t)h->ring_pos;
    if (last_n > gbdr_index::RING) last_n = gbdr_index::RING;
    double a = 0, b = 0, c = 0;
    for (uint32_t i = 0; i < last_n; ++i) {
        cudaEvent_t* ev = h->ring[(h->ring_pos - 1 - i) % gbdr_index::RING];
        GBDR_CUDA(cudaEventSynchronize(ev[3]));
        float x = 0, y = 0, z = 0;
        GBDR_CUDA(cudaEventElapsedTime(&x, ev[0], ev[1]));
        GBDR_CUDA(cudaEventElapsedTime(&y, ev[1], ev[2]));
        GBDR_CUDA(cudaEventElapsedTime(&z, ev[2], ev[3]));
        a += x; b += y; c += z;
    }
    if (project_ms) *project_ms = (float)(a / last_n);
    if (search_ms) *search_ms = (float)(b / last_n);
    if (rerank_ms) *rerank_ms = (float)(c / last_n);
    return GBDR_OK;
}

extern "C" int gbdr_last_kernel_ms(gbdr_index* h, float* project_ms, float* search_ms, float* rerank_ms) {
    return gbdr_kernel_ms(h, 1, project_ms, search_ms, rerank_ms);
}

// ================================================================ kNN build
static int knn_scan_rows(const cudaDeviceProp& prop, const float* d_Q, uint64_t q_begin, uint64_t q_end, const float* d_B,
                         uint64_t n, uint32_t d, uint32_t k, uint32_t* d_out_ids, float* d_out_dists, cudaStream_t st) {
    const uint32_t rpb = knn_rows_per_block();
    const uint64_t nblk = (q_end - q_begin + rpb - 1) / rpb;
    const uint32_t capb = knn_capb(k);
    uint32_t per_sm = capb <= 4096 ? 3 : 1;
    uint32_t grid = (uint32_t)std::min<uint64_t>(nblk, (uint64_t)prop.multiProcessorCount * per_sm);
    uint2* cand = nullptr;
    uint32_t* counter = nullptr;
    GBDR_CUDA(cudaMallocAsync((void**)&cand, (size_t)grid * rpb * capb * sizeof(uint2), st));
    GBDR_CUDA(cudaMallocAsync((void**)&counter, 4, st));
    GBDR_CUDA(cudaMemsetAsync(counter, 0, 4, st));
    int rc = launch_knn_scan(d_Q, d, q_begin, q_end, d_B, d, n, d, k, d_out_ids, d_out_dists, cand, counter, grid, st);
    cudaFreeAsync(cand, st);
    cudaFreeAsync(counter, st);
    return rc;
}

// rows `rows[i]` (relative to row0) of src -> dst[i]; results of dst-ordered rows back to their places
__global__ void knn_gather_rows_kernel(const float* __restrict__ src, uint32_t d, uint64_t row0, const uint32_t* __restrict__ rows,
                                       uint32_t m, float* __restrict__ dst) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (uint64_t)m * d) return;
    const uint32_t i = (uint32_t)(t / d), c = (uint32_t)(t % d);
    dst[t] = src[(row0 + rows[i]) * d + c];
}
template <typename T>
__global__ void knn_scatter_rows_kernel(const T* __restrict__ src, uint32_t k, const uint32_t* __restrict__ rows, uint32_t m,
                                        T* __restrict__ dst) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (uint64_t)m * k) return;
    const uint32_t i = (uint32_t)(t / k), c = (uint32_t)(t % k);
    dst[(uint64_t)rows[i] * k + c] = src[t];
}

// the exact scan for a scattered set of rows (those the tensor-core filter could not bound): gathered into a dense
// block, scanned together, results scattered back
static int knn_scan_listed_rows(const cudaDeviceProp& prop, const float* d_Q, uint64_t q_begin, const std::vector<uint32_t>& rows,
                                const float* d_B, uint64_t n, uint32_t d, uint32_t k, uint32_t* d_out_ids, float* d_out_dists,
                                cudaStream_t st) {
    const uint32_t m = (uint32_t)rows.size();
    uint32_t *d_rows = nullptr, *d_ids = nullptr;
    float *d_q = nullptr, *d_dd = nullptr;
    int rc = GBDR_OK;
    auto step = [&](cudaError_t e, const char* what) {
        if (rc == GBDR_OK && e != cudaSuccess) {
            set_error(std::string("knn redo: ") + what + ": " + cudaGetErrorString(e));
            rc = GBDR_E_CUDA;
        }
        return rc == GBDR_OK;
    };
    step(cudaMallocAsync((void**)&d_rows, (size_t)m * 4, st), "rows");
    step(cudaMallocAsync((void**)&d_q, (size_t)m * d * 4, st), "queries");
    step(cudaMallocAsync((void**)&d_ids, (size_t)m * k * 4, st), "ids");
    if (d_out_dists) step(cudaMallocAsync((void**)&d_dd, (size_t)m * k * 4, st), "dists");
    if (rc == GBDR_OK && step(cudaMemcpyAsync(d_rows, rows.data(), (size_t)m * 4, cudaMemcpyHostToDevice, st), "upload")) {
        knn_gather_rows_kernel<<<(unsigned)(((uint64_t)m * d + 255) / 256), 256, 0, st>>>(d_Q, d, q_begin, d_rows, m, d_q);
        step(cudaGetLastError(), "gather");
        count_launch();
    }
    if (rc == GBDR_OK) rc = knn_scan_rows(prop, d_q, 0, m, d_B, n, d, k, d_ids, d_dd, st);
    if (rc == GBDR_OK) {
        knn_scatter_rows_kernel<uint32_t><<<(unsigned)(((uint64_t)m * k + 255) / 256), 256, 0, st>>>(d_ids, k, d_rows, m, d_out_ids);
        if (d_out_dists)
            knn_scatter_rows_kernel<float><<<(unsigned)(((uint64_t)m * k + 255) / 256), 256, 0, st>>>(d_dd, k, d_rows, m, d_out_dists);
        step(cudaGetLastError(), "scatter");
        count_launch(d_out_dists ? 2 : 1);
    }
    cudaStreamSynchronize(st);  // rows.data() was the source of an asynchronous copy
    for (void* q : {(void*)d_rows, (void*)d_q, (void*)d_ids, (void*)d_dd})
        if (q) cudaFreeAsync(q, st);
    return rc;
}

// sink (optional): host destination streamed chunk by chunk; *stale receives the rows (relative to q_begin) whose
// device results were rewritten after their chunk was copied, or {UINT32_MAX} when everything was
int gbdr::knn_dev_impl(int device, const float* d_Q, uint64_t q_begin, uint64_t q_end, const float* d_B, uint64_t n,
                        uint32_t d, uint32_t k, uint32_t* d_out_ids, float* d_out_dists, void* stream, KnnHostSink* sink,
                        std::vector<uint32_t>* stale) {
    int rc = check_device(device);
    if (rc) return rc;
    if (!d_Q || !d_B || !d_out_ids || d < 4 || (d % 4) || k == 0 || q_end < q_begin) {
        set_error("knn_dev: bad argument (d must be a multiple of 4, k >= 1)");
        return GBDR_E_INVALID;
    }
    if (q_end == q_begin) return GBDR_OK;
    GBDR_CUDA(cudaSetDevice(device));
    {
        // The build's workspaces (2.5 GB of candidate buffers, operand images) come from the device's stream-ordered pool.
        // By default the pool hands freed memory back to the driver at the next synchronisation, so every call would pay
        // for gigabytes of fresh mappings again (0.1 - 1 s, and it varies): keep what was freed cached in the pool.
        static std::atomic<bool> pool_kept[64];
        if (device < 64 && !pool_kept[device].exchange(true)) {
            cudaMemPool_t pool = nullptr;
            uint64_t keep = UINT64_MAX;
            if (cudaDeviceGetDefaultMemPool(&pool, device) != cudaSuccess ||
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep) != cudaSuccess)
                cudaGetLastError();
        }
    }
    cudaStream_t st = (cudaStream_t)stream;
    cudaDeviceProp prop;
    GBDR_CUDA(cudaGetDeviceProperties(&prop, device));
    // tensor-core filter + exact recompute where the shape allows; the exact scan kernel otherwise and
    // for rows whose candidate set the filter could not bound (GBDR_KNN_VARIANT=scan forces the scan)
    const char* variant = getenv("GBDR_KNN_VARIANT");
    const bool force_scan = variant && !strcmp(variant, "scan");
    if (!force_scan && knn_tc_supported(q_end - q_begin, n, d, k)) {
        std::vector<uint32_t> redo;
        rc = launch_knn_tc(d_Q, d, q_begin, q_end, d_B, d, n, d, k, d_out_ids, d_out_dists, prop.multiProcessorCount, st, &redo,
                           sink);
        if (rc) return rc;
        if (redo.size() > (q_end - q_begin) / 4) {  // degenerate data (massive ties): one exact scan of the whole range
            if (stale) stale->assign(1, UINT32_MAX);
            return knn_scan_rows(prop, d_Q, q_begin, q_end, d_B, n, d, k, d_out_ids, d_out_dists, st);
        }
        if (stale) *stale = redo;
        if (redo.size() > 8)  // many scattered rows (a fraction of a percent of a large self-join): one dense scan of them
            return knn_scan_listed_rows(prop, d_Q, q_begin, redo, d_B, n, d, k, d_out_ids, d_out_dists, st);
        for (uint32_t row : redo) {
            rc = knn_scan_rows(prop, d_Q, q_begin + row, q_begin + row + 1, d_B, n, d, k, d_out_ids + (size_t)row * k,
                               d_out_dists ? d_out_dists + (size_t)row * k : nullptr, st);
            if (rc) return rc;
        }
        return GBDR_OK;
    }
    return knn_scan_rows(prop, d_Q, q_begin, q_end, d_B, n, d, k, d_out_ids, d_out_dists, st);
}

extern "C" int gbdr_knn_dev(int device, const float* d_Q, uint64_t q_begin, uint64_t q_end, const float* d_B,
                            uint64_t n, uint32_t d, uint32_t k, uint32_t* d_out_ids, float* d_out_dists,
                            void* stream) {
    return knn_dev_impl(device, d_Q, q_begin, q_end, d_B, n, d, k, d_out_ids, d_out_dists, stream, nullptr, nullptr);
}

extern "C" int gbdr_knn(int device, const float* Q, uint64_t n_q, const float* B, uint64_t n, uint32_t d, uint32_t k,
                        uint32_t* out_ids, float* out_dists, double* gpu_seconds) {
    int rc = check_device(device);
    if (rc) return rc;
    if (!Q || !B || !out_ids || d < 4 || k == 0) {
        set_error("knn: bad argument");
        return GBDR_E_INVALID;
    }
    GBDR_CUDA(cudaSetDevice(device));
    const uint32_t d4 = (d / 4) * 4;
    cudaStream_t st, copy_st;
    GBDR_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    GBDR_CUDA(cudaStreamCreateWithFlags(&copy_st, cudaStreamNonBlocking));
    float *dQ = nullptr, *dB = nullptr, *dD = nullptr;
    uint32_t* dI = nullptr;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    auto cleanup = [&]() {
        cudaStreamSynchronize(st);
        if (dQ && dQ != dB) cudaFree(dQ);
        if (dB) cudaFree(dB);
        if (dI) cudaFree(dI);
        if (dD) cudaFree(dD);
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        cudaStreamSynchronize(copy_st);
        cudaStreamDestroy(copy_st);
        cudaStreamDestroy(st);
    };
#define KNN_TRY(x)                                                                         \
    do {                                                                                   \
        cudaError_t _e = (x);                                                              \
        if (_e != cudaSuccess) {                                                           \
            set_error(std::string(#x) + ": " + cudaGetErrorString(_e));                    \
            cleanup();                                                                     \
            return GBDR_E_CUDA;                                                            \
        }                                                                                  \
    } while (0)
    KNN_TRY(cudaMalloc((void**)&dB, (size_t)n * d4 * 4 + 16));
    if (Q == B && n_q == n) {
        dQ = dB;
    } else {
        KNN_TRY(cudaMalloc((void**)&dQ, (size_t)n_q * d4 * 4 + 16));
    }
    KNN_TRY(cudaMalloc((void**)&dI, (size_t)n_q * k * 4 + 16));
    if (out_dists) KNN_TRY(cudaMalloc((void**)&dD, (size_t)n_q * k * 4 + 16));
    KNN_TRY(cudaEventRecord(e0, st));
    if ((rc = h2d_rows(dB, B, n, d, st))) { cleanup(); return rc; }
    if (dQ != dB && (rc = h2d_rows(dQ, Q, n_q, d, st))) { cleanup(); return rc; }
    // results leave for the host chunk by chunk while later chunks are computed
    KnnHostSink sink;
    sink.ids = out_ids;
    sink.dists = out_dists;
    sink.copy_st = copy_st;
    std::vector<uint32_t> stale;
    rc = knn_dev_impl(device, dQ, 0, n_q, dB, n, d4, k, dI, dD, (void*)st, &sink, &stale);
    if (rc) { cleanup(); return rc; }
    KNN_TRY(cudaStreamSynchronize(copy_st));
    if (!sink.used || (stale.size() == 1 && stale[0] == UINT32_MAX)) {
        KNN_TRY(cudaMemcpyAsync(out_ids, dI, (size_t)n_q * k * 4, cudaMemcpyDeviceToHost, st));
        if (out_dists) KNN_TRY(cudaMemcpyAsync(out_dists, dD, (size_t)n_q * k * 4, cudaMemcpyDeviceToHost, st));
    } else {
        for (uint32_t row : stale) {  // rows redone by the exact scan after their chunk had been copied
            KNN_TRY(cudaMemcpyAsync(out_ids + (size_t)row * k, dI + (size_t)row * k, (size_t)k * 4, cudaMemcpyDeviceToHost, st));
            if (out_dists)
                KNN_TRY(cudaMemcpyAsync(out_dists + (size_t)row * k, dD + (size_t)row * k, (size_t)k * 4, cudaMemcpyDeviceToHost, st));
        }
    }
    KNN_TRY(cudaEventRecord(e1, st));
    KNN_TRY(cudaStreamSynchronize(st));
    if (gpu_seconds) {
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        *gpu_seconds = ms * 1e-3;
    }
#undef KNN_TRY
    cleanup();
    return GBDR_OK;
}

extern "C" int gbdr_gd_prune(int device, const uint64_t* knn_offsets, const uint32_t* knn_edges, const float* db_low,
                             uint64_t n, uint32_t d_low, uint32_t M, int reverse, int need_const_degree,
                             uint64_t* out_offsets, uint32_t* out_edges, double* gpu_seconds) {
    int rc = check_device(device);
    if (rc) return rc;
    if (!knn_offsets || !knn_edges || !db_low || !out_offsets || !out_edges || M < 2 || d_low < 4) {
        set_error("gd_prune: bad argument");
        return GBDR_E_INVALID;
    }
    return gd_prune_device(device, knn_offsets, knn_edges, db_low, n, d_low, M, reverse, need_const_degree, out_offsets,
                           out_edges, gpu_seconds);
}

// cutKNNbyK (support_func.h:309-340): every list re-ranked by distance to its vertex and cut to the knn_size nearest
extern "C" int gbdr_knn_cut(int device, const uint64_t* knn_offsets, const uint32_t* knn_edges, const float* db,
                            uint64_t n, uint32_t d, uint32_t knn_size, uint64_t* out_offsets, uint32_t* out_edges,
                            double* gpu_seconds) {
    int rc = check_device(device);
    if (rc) return rc;
    if (!knn_offsets || !knn_edges || !db || !out_offsets || !out_edges || knn_size < 1 || d < 4) {
        set_error("knn_cut: bad argument");
        return GBDR_E_INVALID;
    }
    return gd_prune_device(device, knn_offsets, knn_edges, db, n, d, 0, 0, 0, out_offsets, out_edges, gpu_seconds, knn_size);
}

// ---- hnswlikeGD on device buffers, and the whole graph build (kNN -> prune) without leaving HBM ----
extern "C" int gbdr_gd_prune_dev(int device, const uint32_t* d_knn, uint32_t k, uint32_t kstride, uint64_t row_begin,
                                 uint64_t row_end, const float* d_db_low, uint64_t n, uint32_t d_low, uint32_t M,
                                 uint32_t* d_fwd, uint32_t* d_deg, void* stream) {
    int rc = check_device(device);
    if (rc) return rc;
    if (!d_knn || !d_db_low || !d_fwd || !d_deg || M < 2 || d_low < 4 || (d_low % 4) || k == 0 || kstride < k || row_end < row_begin ||
        row_end > n) {
        set_error("gd_prune_dev: bad argument (d_low must be a multiple of 4, kstride >= k >= 1, rows within [0, n])");
        return GBDR_E_INVALID;
    }
    GBDR_CUDA(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    uint32_t* counter = nullptr;
    GBDR_CUDA(cudaMallocAsync((void**)&counter, 4, st));
    GBDR_CUDA(cudaMemsetAsync(counter, 0, 4, st));
    int sms = 148;
    GBDR_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    rc = gd_forward_launch(d_knn, kstride, k, row_begin, row_end - row_begin, d_db_low, d_low / 4, M, d_fwd, d_deg, counter, sms, st);
    cudaFreeAsync(counter, st);
    return rc;
}

extern "C" int gbdr_gd_finish_dev(int device, uint32_t* d_fwd, uint32_t* d_deg, uint64_t n, uint32_t M, int reverse,
                                  int need_const_degree, const uint32_t* d_knn, uint32_t k, uint32_t kstride,
                                  uint64_t* out_offsets, uint32_t* out_edges, void* stream) {
    int rc = check_device(device);
    if (rc) return rc;
    if (!d_fwd || !d_deg || !out_offsets || !out_edges || M < 2 || (need_const_degree && (!d_knn || k == 0 || kstride < k))) {
        set_error("gd_finish_dev: bad argument");
        return GBDR_E_INVALID;
    }
    if (n == 0) {
        out_offsets[0] = 0;
        return GBDR_OK;
    }
    return gd_finish(device, d_fwd, d_deg, n, M, reverse, need_const_degree, d_knn, kstride, k, out_offsets, out_edges,
                     (cudaStream_t)stream);
}

extern "C" int gbdr_build_graph(int device, const float* db_low, uint64_t n, uint32_t d_low, uint32_t knn_k, uint32_t M,
                                int reverse, int need_const_degree, uint64_t* out_offsets, uint32_t* out_edges,
                                uint32_t* knn_out, double timings[4]) {
    int rc = check_device(device);
    if (rc) return rc;
    if (!db_low || !out_offsets || !out_edges || d_low < 4 || (d_low % 4) || knn_k == 0 || knn_k > n || M < 2) {
        set_error("build_graph: bad argument (d_low must be a multiple of 4, 1 <= knn_k <= n, M >= 2)");
        return GBDR_E_INVALID;
    }
    GBDR_CUDA(cudaSetDevice(device));
    using clk = std::chrono::steady_clock;
    cudaStream_t st = nullptr, copy_st = nullptr;
    float* dY = nullptr;
    uint32_t *dK = nullptr, *dF = nullptr, *dD = nullptr;
    auto cleanup = [&]() {
        if (st) cudaStreamSynchronize(st);
        if (copy_st) cudaStreamSynchronize(copy_st);
        for (void* q : {(void*)dY, (void*)dK, (void*)dF, (void*)dD})
            if (q) cudaFree(q);
        if (copy_st) cudaStreamDestroy(copy_st);
        if (st) cudaStreamDestroy(st);
    };
#define BG_TRY(x)                                                          \
    do {                                                                   \
        cudaError_t _e = (x);                                              \
        if (_e != cudaSuccess) {                                           \
            set_error(std::string(#x) + ": " + cudaGetErrorString(_e));    \
            cleanup();                                                     \
            return GBDR_E_CUDA;                                            \
        }                                                                  \
    } while (0)
    BG_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    BG_TRY(cudaStreamCreateWithFlags(&copy_st, cudaStreamNonBlocking));
    BG_TRY(cudaMalloc((void**)&dY, (size_t)n * d_low * 4 + 16));
    BG_TRY(cudaMalloc((void**)&dK, (size_t)n * knn_k * 4 + 16));
    BG_TRY(cudaMalloc((void**)&dF, (size_t)n * 2 * M * 4 + 16));
    BG_TRY(cudaMalloc((void**)&dD, (size_t)n * 4 + 16));
    auto t0 = clk::now();
    BG_TRY(cudaMemcpyAsync(dY, db_low, (size_t)n * d_low * 4, cudaMemcpyHostToDevice, st));
    BG_TRY(cudaStreamSynchronize(st));
    auto t1 = clk::now();
    rc = knn_dev_impl(device, dY, 0, n, dY, n, d_low, knn_k, dK, nullptr, (void*)st, nullptr, nullptr);
    if (rc) { cleanup(); return rc; }
    BG_TRY(cudaStreamSynchronize(st));
    auto t2 = clk::now();
    rc = gbdr_gd_prune_dev(device, dK, knn_k, knn_k, 0, n, dY, n, d_low, M, dF, dD, (void*)st);
    if (rc) { cleanup(); return rc; }
    // the kNN lists (the `_knn_1k_` file) leave over PCIe on their own stream while the prune and the reverse pass run
    // (issued after the prune launch: a pageable destination makes this call block the host, not the GPU)
    if (knn_out) BG_TRY(cudaMemcpyAsync(knn_out, dK, (size_t)n * knn_k * 4, cudaMemcpyDeviceToHost, copy_st));
    BG_TRY(cudaStreamSynchronize(st));
    auto t3 = clk::now();
    rc = gd_finish(device, dF, dD, n, M, reverse, need_const_degree, dK, knn_k, knn_k, out_offsets, out_edges, st);
    if (rc) { cleanup(); return rc; }
    BG_TRY(cudaStreamSynchronize(copy_st));
    auto t4 = clk::now();
#undef BG_TRY
    cleanup();
    if (timings) {
        auto sec = [](clk::time_point a, clk::time_point b) { return std::chrono::duration<double>(b - a).count(); };
        timings[0] = sec(t0, t1);
        timings[1] = sec(t1, t2);
        timings[2] = sec(t2, t3);
        timings[3] = sec(t3, t4);
    }
    return GBDR_OK;
}

// ================================================================ merge
extern "C" int gbdr_merge_topk_dev(int device, const uint32_t* d_in_ids, const float* d_in_dists, uint32_t parts,
                                   uint32_t n_q, uint32_t k_in, uint32_t k_out, uint32_t* d_out_ids,
                                   float* d_out_dists, void* stream) {
    int rc = check_device(device);
    if (rc) return rc;
    if (!d_in_ids || !d_in_dists || !d_out_ids || parts == 0 || k_in == 0 || k_out == 0) return GBDR_E_INVALID;
    GBDR_CUDA(cudaSetDevice(device));
    return launch_merge_topk(d_in_ids, d_in_dists, parts, n_q, k_in, k_out, d_out_ids, d_out_dists,
                             (cudaStream_t)stream);
}

// ================================================================ raw memory helpers
extern "C" int gbdr_dev_malloc(int device, size_t bytes, void** out) {
    if (!out) return GBDR_E_INVALID;
    int rc = check_device(device);
    if (rc) return rc;
    GBDR_CUDA(cudaSetDevice(device));
    GBDR_CUDA(cudaMalloc(out, bytes ? bytes : 16));
    return GBDR_OK;
}
extern "C" int gbdr_dev_free(int device, void* p) {
    if (!p) return GBDR_OK;
    GBDR_CUDA(cudaSetDevice(device));
    GBDR_CUDA(cudaFree(p));
    return GBDR_OK;
}
extern "C" int gbdr_memcpy_h2d(int device, void* dst, const void* src, size_t bytes) {
    GBDR_CUDA(cudaSetDevice(device));
    GBDR_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice));
    return GBDR_OK;
}
extern "C" int gbdr_memcpy_d2h(int device, void* dst, const void* src, size_t bytes) {
    GBDR_CUDA(cudaSetDevice(device));
    GBDR_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
    return GBDR_OK;
}
extern "C" int gbdr_host_alloc_pinned(size_t bytes, void** out) {
    if (!out) return GBDR_E_INVALID;
    int cnt = 0;
    if (cudaGetDeviceCount(&cnt) != cudaSuccess || cnt == 0) {
        cudaGetLastError();
        set_error("no CUDA device visible");
        return GBDR_E_NO_DEVICE;
    }
    GBDR_CUDA(cudaHostAlloc(out, bytes ? bytes : 16, cudaHostAllocDefault));
    return GBDR_OK;
}
extern "C" int gbdr_host_free_pinned(void* p) {
    if (!p) return GBDR_OK;
    GBDR_CUDA(cudaFreeHost(p));
    return GBDR_OK;
}
// Page-lock a caller-owned buffer in place (and undo it): what a host that cannot allocate its vectors with
// gbdr_host_alloc_pinned (std::vector, numpy) does once per buffer so that submit / wait copies are true DMA.
extern "C" int gbdr_host_register(void* p, size_t bytes) {
    if (!p || !bytes) return GBDR_E_INVALID;
    int cnt = 0;
    if (cudaGetDeviceCount(&cnt) != cudaSuccess || cnt == 0) {
        cudaGetLastError();
        set_error("no CUDA device visible");
        return GBDR_E_NO_DEVICE;
    }
    const cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterPortable);
    if (e != cudaSuccess) {
        cudaGetLastError();  // e.g. already registered, or the platform refuses: the buffer simply stays pageable
        set_error(std::string("cudaHostRegister: ") + cudaGetErrorString(e));
        return GBDR_E_CUDA;
    }
    return GBDR_OK;
}
extern "C" int gbdr_host_unregister(void* p) {
    if (!p) return GBDR_OK;
    const cudaError_t e = cudaHostUnregister(p);
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error(std::string("cudaHostUnregister: ") + cudaGetErrorString(e));
        return GBDR_E_CUDA;
    }
    return GBDR_OK;
}
extern "C" int gbdr_device_synchronize(int device) {
    GBDR_CUDA(cudaSetDevice(device));
    GBDR_CUDA(cudaDeviceSynchronize());
    return GBDR_OK;
}
