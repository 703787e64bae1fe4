// merge.cu — K5: k-way merge of per-shard top-k lists (no reference equivalent; the north-star's
// sharded-index mode).  Each input list is ascending by (dist,id); the rank of an
// element in the merged order is its own position plus, for every other list, the number of
// entries that precede it there (binary search), so one thread per input element suffices and the
// output is written without any inter-thread communication.
#include "kernels.cuh"

namespace gbdr {

namespace {

__global__ void merge_topk_kernel(const uint32_t* __restrict__ in_ids, const float* __restrict__ in_dists,
                                  uint32_t parts, uint32_t n_q, uint32_t k_in, uint32_t k_out,
                                  uint32_t* __restrict__ out_ids, float* __restrict__ out_dists) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t per_q = (uint64_t)parts * k_in;
    if (t >= per_q * n_q) return;
    const uint32_t q = (uint32_t)(t / per_q);
    const uint32_t rem = (uint32_t)(t % per_q);
    const uint32_t pp = rem / k_in, j = rem % k_in;
    const size_t me = ((size_t)pp * n_q + q) * k_in + j;
    const uint32_t id = in_ids[me];
    if (id == PAD_ID) return;
    const float dist = in_dists[me];
    uint32_t rank = j;
    for (uint32_t o = 0; o < parts; ++o) {
        if (o == pp) continue;
        const size_t base = ((size_t)o * n_q + q) * k_in;
        // number of valid entries of list o that order before (dist,id,pp)
        uint32_t lo = 0, hi = k_in;
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            const uint32_t oid = in_ids[base + mid];
            const float od = in_dists[base + mid];
            const bool before = oid != PAD_ID && (pair_less(od, oid, dist, id) || (od == dist && oid == id && o < pp));
            if (before) lo = mid + 1; else hi = mid;
        }
        rank += lo;
    }
    if (rank < k_out) {
        out_ids[(size_t)q * k_out + rank] = id;
        if (out_dists) out_dists[(size_t)q * k_out + rank] = dist;
    }
}

__global__ void fill_pad_kernel(uint32_t* ids, float* dists, uint64_t n) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    ids[t] = PAD_ID;
    if (dists) dists[t] = __int_as_float(0x7f800000);
}

}  // namespace

int launch_merge_topk(const uint32_t* in_ids, const float* in_dists, uint32_t parts, uint32_t n_q, uint32_t k_in,
                      uint32_t k_out, uint32_t* out_ids, float* out_dists, cudaStream_t st) {
    const uint64_t nout = (uint64_t)n_q * k_out;
    if (nout == 0) return GBDR_OK;
    fill_pad_kernel<<<(unsigned)((nout + 255) / 256), 256, 0, st>>>(out_ids, out_dists, nout);
    GBDR_CHECK_LAUNCH();
    const uint64_t total = (uint64_t)n_q * parts * k_in;
    if (total) {
        merge_topk_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(in_ids, in_dists, parts, n_q, k_in, k_out,
                                                                          out_ids, out_dists);
        GBDR_CHECK_LAUNCH();
    }
    count_launch(2);
    return GBDR_OK;
}

}  // namespace gbdr
