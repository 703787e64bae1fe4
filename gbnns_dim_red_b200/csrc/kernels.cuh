// kernels.cuh — launcher declarations shared between the kernel translation units and capi.cu.
#pragma once
#include <vector>

#include "common.cuh"

namespace gbdr {

struct RerankParams {
    const float* queries;   // [n_q x q_stride]
    uint32_t q_stride;
    const float* db;        // [n x row_stride]
    uint32_t row_stride;
    uint32_t C;             // d/4 chunks
    const uint32_t* cand;   // [n_q x m] ascending low-dim order, PAD-terminated
    uint32_t m;             // candidates per query (= ef)
    uint32_t k;             // outputs per query (<= m)
    uint32_t n_q;
    uint32_t id_offset;
    uint32_t* out_ids;      // [n_q x k]
    float* out_dists;       // [n_q x k] or null
    uint32_t smem_per_warp;
};
int launch_rerank(const RerankParams& p, cudaStream_t st);

int launch_project_fp32(const float* X, uint32_t ldx, uint32_t n_q, const float* l1, const float* l2,
                        const float* l3, uint32_t d, uint32_t dh, uint32_t dh2, uint32_t d_low, float* h1,
                        float* h2, float* out, uint32_t ld_out, cudaStream_t st);
__global__ void normalize_rows_kernel(float* __restrict__ Y, uint32_t ld, uint32_t M, uint32_t d_low);

struct ProjTcPlan;
int project_tc_prepare(const float* l1, const float* l2, const float* l3, uint32_t d, uint32_t dh, uint32_t dh2,
                       uint32_t d_low, cudaStream_t st, ProjTcPlan** out);
void project_tc_destroy(ProjTcPlan* p);
int launch_project_tc(ProjTcPlan* plan, const float* X, uint32_t ldx, uint32_t n_q, float* out, uint32_t ld_out,
                      int single_pass, cudaStream_t st);

int launch_knn_scan(const float* Q, uint32_t ldq, uint64_t q_begin, uint64_t q_end, const float* B, uint32_t ldb,
                    uint64_t n, uint32_t d, uint32_t k, uint32_t* out_ids, float* out_dists, uint2* cand,
                    uint32_t* counter, uint32_t grid, cudaStream_t st);
// Host destination of a kNN build: finished row chunks are copied out on `copy_st` while later chunks are still
// being computed (the n x k id matrix is 4 GB at 1M x 1000: its PCIe time would otherwise be added to the build).
struct KnnHostSink {
    uint32_t* ids = nullptr;    // [rows x k] host
    float* dists = nullptr;     // [rows x k] host or null
    cudaStream_t copy_st = nullptr;
    bool used = false;          // set once chunk copies were queued on copy_st
};
// tensor-core filter + exact recompute (knn_tc.cu); rows it could not bound are returned in overflow_rows
bool knn_tc_supported(uint64_t n_rows, uint64_t n, uint32_t d, uint32_t k);
int launch_knn_tc(const float* d_Q, uint32_t ldq, uint64_t q_begin, uint64_t q_end, const float* d_B, uint32_t ldb, uint64_t n,
                  uint32_t d, uint32_t k, uint32_t* d_out_ids, float* d_out_dists, int sm_count, cudaStream_t st,
                  std::vector<uint32_t>* overflow_rows, KnnHostSink* sink = nullptr);
uint32_t knn_capb(uint32_t k);
uint32_t knn_rows_per_block();

int launch_merge_topk(const uint32_t* in_ids, const float* in_dists, uint32_t parts, uint32_t n_q, uint32_t k_in,
                      uint32_t k_out, uint32_t* out_ids, float* out_dists, cudaStream_t st);

// hnswlikeGD in pieces, all on device buffers (gd_prune.cu): forward lists of a row block, id validation, and everything
// after the forward lists (reverse pass, constant-degree fill, flattened output)
int gd_forward_launch(const uint32_t* d_knn, uint32_t kstride, uint32_t klen, uint64_t row0, uint64_t rows, const float* d_db,
                      uint32_t C, uint32_t M, uint32_t* d_fwd, uint32_t* d_deg, uint32_t* d_counter, int sm_count,
                      cudaStream_t st, uint32_t cut_k = 0);
int gd_check_ids(const uint32_t* d_knn, uint64_t rows, uint32_t kstride, uint32_t klen, uint64_t n, uint32_t* d_flag,
                 cudaStream_t st);
int gd_finish(int device, uint32_t* d_fwd, uint32_t* d_deg, uint64_t n, uint32_t M, int reverse, int need_const_degree,
              const uint32_t* d_knn, uint32_t kstride, uint32_t klen, uint64_t* out_offsets, uint32_t* out_edges,
              cudaStream_t st);
int gd_prune_device(int device, const uint64_t* knn_offsets, const uint32_t* knn_edges, const float* db_low,
                    uint64_t n, uint32_t d_low, uint32_t M, int reverse, int need_const_degree,
                    uint64_t* out_offsets, uint32_t* out_edges, double* gpu_seconds, uint32_t cut_k = 0);

}  // namespace gbdr
