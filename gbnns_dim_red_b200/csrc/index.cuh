// index.cuh — the index object behind gbdr_index* and the internal entry points shared by capi.cu and group.cu.
#pragma once
#include <atomic>
#include <string>

#include "beam_search.cuh"
#include "common.cuh"
#include "kernels.cuh"

namespace gbdr {

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    bool borrowed = false;  // a view's alias of its parent's buffer: never freed or resized here
    void borrow(const DevBuf& o) {
        release();
        p = o.p;
        bytes = o.bytes;
        borrowed = o.p != nullptr;
    }
    int ensure(size_t need) {
        if (need <= bytes && !borrowed) return GBDR_OK;
        if (borrowed) {
            set_error("internal: resize of a borrowed buffer");
            return GBDR_E_STATE;
        }
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
        size_t want = need + need / 4;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            set_error(std::string("cudaMalloc: ") + cudaGetErrorString(e));
            return GBDR_E_NOMEM;
        }
        bytes = want;
        return GBDR_OK;
    }
    void release() {
        if (p && !borrowed) cudaFree(p);
        p = nullptr;
        bytes = 0;
        borrowed = false;
    }
    template <typename T>
    T* as() const { return reinterpret_cast<T*>(p); }
};

// device checks (cached per device) and the search sequence on a stream (capi.cu)
int check_device(int device);

}  // namespace gbdr

struct gbdr_index {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    uint64_t n_base = 0, n_low = 0, n_graph = 0;
    uint32_t d = 0, d_low = 0, C = 0, C_low = 0;
    gbdr::DevBuf db, low, adj, aux;
    uint32_t adj_stride = 0;
    uint32_t aux_stride = 0, hops_bound = 50, llf = 0;  // second graph (search_function.h:73-89)
    uint64_t n_aux = 0;
    // net (reference layout, device copies)
    gbdr::DevBuf l1, l2, l3;
    uint32_t net_d = 0, dh = 0, dh2 = 0, net_dlow = 0;
    bool has_net = false;
    int proj_mode = GBDR_PROJ_3XTF32;
    gbdr::ProjTcPlan* tc_plan = nullptr;
    uint64_t id_offset = 0;
    // workspaces
    gbdr::DevBuf w_q, w_qlow, w_entry, w_low_ids, w_out_ids, w_out_dists, w_hops, w_dc, w_scanned, w_h1, w_h2, w_status,
        w_spill;
    cudaEvent_t ev[8] = {};
    static constexpr int RING = 256;
    cudaEvent_t ring[RING][4] = {};   // per search call: start, after projection, after search, after re-rank
    uint64_t ring_pos = 0;            // number of timed calls so far
    bool timed = false;
    // views (gbdr_index_create_view): share the parent's resident arrays, own stream + workspaces
    gbdr_index* parent = nullptr;
    uint64_t epoch = 0;               // parent: bumped by every set_*; view: the parent epoch it mirrors
    std::atomic<int> n_views{0};
    // asynchronous host call (gbdr_search_submit / gbdr_search_wait)
    uint32_t* h_status = nullptr;     // pinned: status word of the call in flight
    bool pending = false;
    // per-warp HBM overflow tables of the visited set: sized for the beam width (2 x the expected visited count) and
    // grown to the maximum by the first call that exhausts them (gbdr_search_wait re-runs that call itself)
    uint32_t spill_min = 0;           // log2 of the smallest per-warp table this handle may use (0 = from ef)
    // CUDA graph of the last host-buffer call (gbdr_search_submit): a serving loop repeats one call shape on the same
    // page-locked buffers, and for small batches the ~25 stream operations of a call cost the host more than the GPU
    // needs for them.  The second identical call is captured, later ones are one cudaGraphLaunch.  Any other activity on
    // the handle (another shape, other buffers, set_*, _dev calls, a projection) drops the graph.
    struct GraphKey {
        const void *queries, *q_low, *entry, *out_ids, *out_dists, *hops, *dist_calc;
        uint32_t n_q, ef, k, flags, spill_min;
        uint64_t epoch;
        int proj_mode, on_device;
    };
    GraphKey graph_key = {};       // shape of the last host-buffer call
    bool graph_key_valid = false;
    cudaGraphExec_t graph_exec = nullptr;
    uint64_t graph_launches = 0;   // kernels inside the graph (gbdr_launch_count)
    bool graph_off = false;        // capture failed once on this handle: stay on the plain path
    struct Call {
        const float *queries, *q_low;
        uint32_t n_q, ef, k, flags;
        const uint32_t* entry;
        uint32_t* out_ids;
        float* out_dists;
        int32_t *hops, *dist_calc;
        bool on_device;
    } call = {};
    // private views the blocking gbdr_search pipelines its batch over (created on first use, destroyed with the handle)
    static constexpr int MAX_HELPERS = 3;
    gbdr_index* helpers[MAX_HELPERS] = {};
    bool internal = false;
};

static constexpr uint32_t SPILL_LOG_MAX = 16;

namespace gbdr {
// d_q: original queries (stride ldq floats), d_qlow: low-dim queries (stride ldql) or null -> project; results into
// the given device buffers; asynchronous on st
int search_on_stream(gbdr_index* h, const float* d_q, uint32_t ldq, const float* d_qlow, uint32_t ldql, uint32_t n_q,
                     uint32_t ef, uint32_t k, uint32_t flags, const uint32_t* d_entry, uint32_t* d_out_ids,
                     float* d_out_dists, int32_t* d_hops, int32_t* d_dc, int32_t* d_scanned, cudaStream_t st, bool timed);
int sync_view(gbdr_index* v);
// enqueue one search call on the handle's stream (host buffers).  results_stay_on_device: ids / dists are left in
// h->w_out_ids / h->w_out_dists for a device-side consumer (the group's merge) and out_ids may be null
int search_submit_impl(gbdr_index* h, const float* queries, const float* q_low, uint32_t n_q, uint32_t ef, uint32_t k,
                       uint32_t flags, const uint32_t* entry, uint32_t* out_ids, float* out_dists, int32_t* hops,
                       int32_t* dist_calc, bool results_stay_on_device);
// kNN of rows [q_begin, q_end) of d_Q among d_B, device buffers (tensor-core filter + exact recompute, exact scan behind it)
int knn_dev_impl(int device, const float* d_Q, uint64_t q_begin, uint64_t q_end, const float* d_B, uint64_t n, uint32_t d,
                 uint32_t k, uint32_t* d_out_ids, float* d_out_dists, void* stream, KnnHostSink* sink,
                 std::vector<uint32_t>* stale);
}  // namespace gbdr
