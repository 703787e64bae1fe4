// gd_prune.cu — K6: hnswlikeGD on the GPU (reference search/support_func.h:521-575), the graph
// pruning step of search/prepare_graph.cpp:70.
//
// Per vertex i (one warp each):
//   A. distances to all candidates of its kNN list in canonical arithmetic, drop dist <= eps
//      (:532-539), sort ascending.  The reference's std::sort compares dist only (:63-66, unstable
//      at exact ties); here ties are ordered by id (the definition DESIGN.md fixes for the CPU checker too).
//   B. greedy diversity prune (:541-558): candidate c is kept iff for every already kept a
//      dist(c,i) + eps <= dist(c,a); stops at M kept.  All kept rows live in shared memory, one
//      lane evaluates one (c,a) pair, candidates are staged 32 rows at a time with cp.async.
//   C. force-add the M/2 nearest not yet present (:559-563).
// The reverse-edge pass (addReverseEdgesForGD, :402-445) is order-dependent (vertex i sees the lists as
// modified by all i' < i).  Its membership tests turn out to depend on the forward lists only and run on the
// GPU (gd_mutual_kernel); the host keeps the one sequential piece, "is the target row full yet", in the
// reference's iteration order.  The optional pad-to-2M pass (getConstantDegreeForGD, :466-485) is per row.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <thread>

#include "kernels.cuh"

namespace gbdr {

namespace {

struct GdParams {
    const uint32_t* knn;       // [n x kstride] candidate lists of the rows of this block, klen ids each (PAD tail allowed)
    uint32_t kstride, klen;
    const float* db;           // [n_total x C*4]: all vectors
    uint32_t C;
    uint64_t row0;             // global id of the block's first row
    uint64_t n;                // rows in the block
    uint32_t M;
    uint32_t sort_cap;         // power of two >= max candidates
    uint32_t fwd_stride;       // row stride of fwd: 2M (prune) or cut_k (cut)
    uint32_t cut_k;            // > 0: cutKNNbyK mode — keep the cut_k nearest candidates of the list, no pruning
    uint32_t* fwd;             // [n x fwd_stride]
    uint32_t* deg;             // [n]
    uint32_t* counter;
    uint32_t smem_per_warp;
};

__host__ __device__ inline uint32_t gd_smem_per_warp(uint32_t C, uint32_t M, uint32_t sort_cap) {
    // sorted (dist,id) | kept rows [M x C] | stage rows [32 x C] | self row [C] | kept ids [M + M/2]
    uint32_t b = sort_cap * 8u + M * C * 16u + 32u * C * 16u + C * 16u + ((M + M / 2 + 3u) & ~3u) * 4u;
    return (b + 15u) & ~15u;
}

__device__ __forceinline__ uint32_t rot(uint32_t r, uint32_t c, uint32_t C) {
    // chunk rotation so that lanes reading chunk c of their own row hit different banks
    uint32_t x = c + (r % C);
    return x >= C ? x - C : x;
}

__global__ void __launch_bounds__(128) gd_prune_kernel(const GdParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const uint32_t C = p.C;
    unsigned char* wb = smem_raw + (size_t)warp * p.smem_per_warp;
    float* sd = reinterpret_cast<float*>(wb);                                   // [sort_cap]
    uint32_t* si = reinterpret_cast<uint32_t*>(sd + p.sort_cap);                // [sort_cap]
    float4* kept = reinterpret_cast<float4*>(si + p.sort_cap);                  // [M][C] rotated
    float4* stage = kept + (size_t)p.M * C;                                     // [32][C] rotated
    float4* self = stage + 32 * C;                                              // [C]
    uint32_t* kept_id = reinterpret_cast<uint32_t*>(self + C);                  // [M + M/2]
    const float eps = 1e-10f;  // getEps(), support_func.h:41-43
    const float INF = __int_as_float(0x7f800000);

    for (;;) {
        uint32_t vi = 0;
        if (lane == 0) vi = atomicAdd(p.counter, 1u);
        vi = __shfl_sync(FULL_MASK, vi, 0);
        if (vi >= p.n) break;
        const uint32_t* cl = p.knn + (size_t)vi * p.kstride;
        for (uint32_t c = lane; c < C; c += 32)
            self[c] = __ldg(reinterpret_cast<const float4*>(p.db + (size_t)(p.row0 + vi) * C * 4) + c);
        for (uint32_t i = lane; i < p.sort_cap; i += 32) {
            sd[i] = INF;
            si[i] = PAD_ID;
        }
        __syncwarp();

        // ---- A: distances to candidates, 32 at a time (:532-539) ----
        uint32_t m = 0;
        for (uint32_t b0 = 0; b0 < p.klen; b0 += 32) {
            const uint32_t cid = b0 + lane < p.klen ? __ldg(cl + b0 + lane) : PAD_ID;
            const unsigned vmask = __ballot_sync(FULL_MASK, cid != PAD_ID);
            if (!vmask) break;
            const uint32_t mb = __popc(vmask);  // PAD only at the tail
            // stage rows (uniform loop, shuffle inside)
            const uint32_t T = mb * C;
            for (uint32_t t0 = 0; t0 < T; t0 += 32) {
                const uint32_t t = t0 + lane;
                const uint32_t r = t / C, c = t - r * C;
                const uint32_t rid = __shfl_sync(FULL_MASK, cid, r & 31);
                if (t < T) cp_async16(stage + r * C + rot(r, c, C), p.db + (size_t)rid * C * 4 + c * 4u);
            }
            cp_async_commit();
            cp_async_wait<0>();
            __syncwarp();
            float dist = 0.f;
            bool ok = false;
            if ((uint32_t)lane < mb) {
                L2Acc acc;
                for (uint32_t c = 0; c < C; ++c) acc.add(self[c], stage[lane * C + rot(lane, c, C)]);
                dist = acc.result();
                ok = p.cut_k ? true : dist > eps;  // :535; cutKNNbyK keeps everything (:315-320)
            }
            const unsigned km = __ballot_sync(FULL_MASK, ok);
            if (ok) {
                const uint32_t pos = m + __popc(km & lanemask_lt());
                sd[pos] = dist;
                si[pos] = cid;
            }
            m += __popc(km);
            __syncwarp();
        }

        // ---- sort ascending by (dist,id) (:540): warp bitonic over sort_cap ----
        uint32_t ncap = 32;
        while (ncap < m) ncap <<= 1;  // only the occupied power-of-two prefix needs sorting
        for (uint32_t size = 2; size <= ncap; size <<= 1) {
            for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
                for (uint32_t t = lane; t < (ncap >> 1); t += 32) {
                    const uint32_t lo = 2 * t - (t & (stride - 1));
                    const uint32_t hi = lo + stride;
                    const bool up = ((lo & size) == 0);
                    const float dl = sd[lo], dh = sd[hi];
                    const uint32_t il = si[lo], ih = si[hi];
                    const bool lt = pair_less(dh, ih, dl, il);
                    if (lt == up) {
                        sd[lo] = dh; sd[hi] = dl;
                        si[lo] = ih; si[hi] = il;
                    }
                }
                __syncwarp();
            }
        }

        if (p.cut_k) {  // cutKNNbyK (support_func.h:309-340): the knn_size nearest of the list, nearest first
            const uint32_t keep = min(p.cut_k, m);
            for (uint32_t a = lane; a < keep; a += 32) p.fwd[(size_t)vi * p.fwd_stride + a] = si[a];
            if (lane == 0) p.deg[vi] = keep;
            __syncwarp();
            continue;
        }

        // ---- B: greedy prune (:541-558) ----
        uint32_t nk = 0;
        auto keep_row = [&](uint32_t slot_in_stage, uint32_t id) {
            // copy a staged row into kept[nk] (re-rotated for the kept index)
            for (uint32_t c = lane; c < C; c += 32)
                kept[nk * C + rot(nk, c, C)] = stage[slot_in_stage * C + rot(slot_in_stage, c, C)];
            if (lane == 0) kept_id[nk] = id;
            __syncwarp();
            ++nk;
        };
        if (m > 0) {
            bool done = false;
            for (uint32_t b0 = 0; b0 < m && !done; b0 += 32) {
                const uint32_t mb = min(32u, m - b0);
                const uint32_t cid = (uint32_t)lane < mb ? si[b0 + lane] : 0u;
                const uint32_t T = mb * C;
                for (uint32_t t0 = 0; t0 < T; t0 += 32) {
                    const uint32_t t = t0 + lane;
                    const uint32_t r = t / C, c = t - r * C;
                    const uint32_t rid = __shfl_sync(FULL_MASK, cid, r & 31);
                    if (t < T) cp_async16(stage + r * C + rot(r, c, C), p.db + (size_t)rid * C * 4 + c * 4u);
                }
                cp_async_commit();
                cp_async_wait<0>();
                __syncwarp();
                for (uint32_t j = 0; j < mb; ++j) {
                    const uint32_t gj = b0 + j;
                    if (gj == 0) {  // nearest is always kept (:541)
                        keep_row(0, si[0]);
                        continue;
                    }
                    const float lhs = __fadd_rn(sd[gj], eps);  // Dist(pre,i) + eps (:547)
                    bool bad = false;
                    for (uint32_t a0 = 0; a0 < nk; a0 += 32) {
                        const uint32_t a = a0 + lane;
                        bool mybad = false;
                        if (a < nk) {
                            L2Acc acc;
                            for (uint32_t c = 0; c < C; ++c)
                                acc.add(stage[j * C + rot(j, c, C)], kept[a * C + rot(a, c, C)]);
                            mybad = lhs > acc.result();
                        }
                        if (__ballot_sync(FULL_MASK, mybad)) {
                            bad = true;
                            break;
                        }
                    }
                    if (!bad) {
                        keep_row(j, si[gj]);
                        if (nk == p.M) {  // :555-557
                            done = true;
                            break;
                        }
                    }
                }
                __syncwarp();
            }
            // ---- C: force-add the M/2 nearest (:559-563) ----
            uint32_t nout = nk;
            const uint32_t edge = min(p.M / 2, m);
            for (uint32_t j = 0; j < edge; ++j) {
                const uint32_t id = si[j];
                bool found = false;
                for (uint32_t a0 = 0; a0 < nout; a0 += 32) {
                    const uint32_t a = a0 + lane;
                    if (__ballot_sync(FULL_MASK, a < nout && kept_id[a] == id)) {
                        found = true;
                        break;
                    }
                }
                if (!found) {
                    if (lane == 0) kept_id[nout] = id;
                    __syncwarp();
                    ++nout;
                }
            }
            for (uint32_t a = lane; a < nout; a += 32) p.fwd[(size_t)vi * p.fwd_stride + a] = kept_id[a];
            if (lane == 0) p.deg[vi] = nout;
        } else {
            if (lane == 0) p.deg[vi] = 0;
        }
        __syncwarp();
    }
}

// Static part of addReverseEdgesForGD (support_func.h:423-442).  Vertex i offers itself to c = fwd(i)[j] unless
// c already lists i (:430) — and whether it does can be read off the FORWARD lists alone: entries appended to a
// row by the pass are reverse edges i' with i in fwd(i'), so the walk over them never appends, and i is never
// among the entries appended to c before its own turn.  One warp per vertex: bit j of mask[i] = "c does not
// list i"; indeg[c] = number of forward lists naming c (:418-422).  What remains sequential on the host is only
// "is row c full yet" (:429) in ascending i.
__global__ void __launch_bounds__(256) gd_mutual_kernel(const uint32_t* __restrict__ fwd, const uint32_t* __restrict__ deg,
                                                        uint64_t n, uint32_t stride, uint32_t words,
                                                        unsigned long long* __restrict__ mask, uint32_t* __restrict__ indeg) {
    const int lane = threadIdx.x & 31;
    const uint64_t nwarps = (uint64_t)gridDim.x * (blockDim.x >> 5);
    for (uint64_t i = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < n; i += nwarps) {
        const uint32_t di = deg[i];
        for (uint32_t w = 0; w < words; ++w) {
            unsigned long long m = 0;
            for (uint32_t j0 = w * 64u; j0 < di && j0 < (w + 1u) * 64u; j0 += 32) {
                const uint32_t j = j0 + lane;
                bool offer = false;
                if (j < di) {
                    const uint32_t c = fwd[i * stride + j];
                    atomicAdd(&indeg[c], 1u);
                    const uint32_t dc = deg[c];
                    const uint32_t* crow = fwd + (size_t)c * stride;
                    offer = true;
                    for (uint32_t l = 0; l < dc; ++l)
                        if (crow[l] == (uint32_t)i) offer = false;
                }
                m |= (unsigned long long)__ballot_sync(FULL_MASK, offer) << (j0 - w * 64u);
            }
            if (lane == 0) mask[i * words + w] = m;
        }
    }
}

// getConstantDegreeForGD (support_func.h:466-485): rows shorter than `cap` are filled, in candidate order starting at
// the SECOND entry of the vertex's kNN list, with the candidates the row does not hold yet.  One warp per row; a chunk of
// 32 candidates is tested against the row as it stands (and against each other, for lists with repeated ids).
__global__ void __launch_bounds__(256) gd_fill_kernel(uint32_t* __restrict__ g, uint32_t* __restrict__ deg, uint64_t n, uint32_t cap,
                                                      const uint32_t* __restrict__ knn, uint32_t kstride, uint32_t klen) {
    const int lane = threadIdx.x & 31;
    const uint64_t nwarps = (uint64_t)gridDim.x * (blockDim.x >> 5);
    for (uint64_t i = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < n; i += nwarps) {
        uint32_t d = deg[i];
        uint32_t* row = g + i * cap;
        const uint32_t* cl = knn + i * kstride;
        for (uint32_t j0 = 1; j0 < klen && d < cap; j0 += 32) {
            const uint32_t j = j0 + lane;
            const uint32_t cid = j < klen ? cl[j] : PAD_ID;
            bool add = cid != PAD_ID;
            for (uint32_t l = 0; l < d && add; ++l)
                if (row[l] == cid) add = false;
            // a repeated id inside the chunk: only its first occurrence is appended
            const unsigned same = __match_any_sync(FULL_MASK, cid);
            if (add && (same & lanemask_lt())) add = false;
            const unsigned am = __ballot_sync(FULL_MASK, add);
            const uint32_t pos = d + __popc(am & lanemask_lt());
            if (add && pos < cap) row[pos] = cid;
            d = min(cap, d + (uint32_t)__popc(am));
            __syncwarp();
        }
        if (lane == 0) deg[i] = d;
    }
}

__global__ void gd_check_ids_kernel(const uint32_t* __restrict__ knn, uint64_t rows, uint32_t kstride, uint32_t klen, uint64_t n,
                                    uint32_t* __restrict__ bad) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= rows * klen) return;
    const uint32_t id = knn[(t / klen) * kstride + (t % klen)];
    if (id != PAD_ID && id >= n) *bad = 1u;
}

}  // namespace

// forward lists (steps A-C) of the block's rows on `st`; d_counter: one zeroed word
int gd_forward_launch(const uint32_t* d_knn, uint32_t kstride, uint32_t klen, uint64_t row0, uint64_t rows, const float* d_db,
                      uint32_t C, uint32_t M, uint32_t* d_fwd, uint32_t* d_deg, uint32_t* d_counter, int sm_count,
                      cudaStream_t st, uint32_t cut_k) {
    if (rows == 0) return GBDR_OK;
    uint32_t sort_cap = 32;
    while (sort_cap < klen) sort_cap <<= 1;
    GdParams p;
    memset(&p, 0, sizeof(p));
    p.smem_per_warp = gd_smem_per_warp(C, M, sort_cap);
    if (p.smem_per_warp > 220u * 1024u) {
        set_error("gd_prune: candidate lists too long for the shared-memory sort (max degree too large)");
        return GBDR_E_CAPACITY;
    }
    const uint32_t wpb = std::max<uint32_t>(1, std::min<uint32_t>(4, (200u * 1024u) / p.smem_per_warp));
    const size_t smem = (size_t)wpb * p.smem_per_warp;
    p.knn = d_knn; p.kstride = kstride; p.klen = klen; p.db = d_db; p.C = C; p.row0 = row0; p.n = rows; p.M = M;
    p.sort_cap = sort_cap; p.fwd_stride = cut_k ? cut_k : 2 * M; p.cut_k = cut_k; p.fwd = d_fwd; p.deg = d_deg; p.counter = d_counter;
    // rows are written up to their degree only: fill the rest with PAD so that the matrix is defined wherever it is copied
    GBDR_CUDA(cudaMemsetAsync(d_fwd, 0xFF, (size_t)rows * p.fwd_stride * 4, st));
    GBDR_CUDA(cudaFuncSetAttribute(gd_prune_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const uint32_t per_sm = std::max<uint32_t>(1, (uint32_t)((227u * 1024u) / (smem + 1024)));
    const uint32_t grid = (uint32_t)std::min<uint64_t>((rows + wpb - 1) / wpb, (uint64_t)sm_count * per_sm);
    gd_prune_kernel<<<grid, wpb * 32, smem, st>>>(p);
    GBDR_CHECK_LAUNCH();
    count_launch();
    return GBDR_OK;
}

// are all ids of the lists vertices?  (they index the vector matrix on the device)
int gd_check_ids(const uint32_t* d_knn, uint64_t rows, uint32_t kstride, uint32_t klen, uint64_t n, uint32_t* d_flag, cudaStream_t st) {
    const uint64_t total = rows * klen;
    if (!total) return GBDR_OK;
    GBDR_CUDA(cudaMemsetAsync(d_flag, 0, 4, st));
    gd_check_ids_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(d_knn, rows, kstride, klen, n, d_flag);
    GBDR_CHECK_LAUNCH();
    count_launch();
    uint32_t bad = 0;
    GBDR_CUDA(cudaMemcpyAsync(&bad, d_flag, 4, cudaMemcpyDeviceToHost, st));
    GBDR_CUDA(cudaStreamSynchronize(st));
    if (bad) {
        set_error("gd_prune: candidate id out of range");
        return GBDR_E_INVALID;
    }
    return GBDR_OK;
}

// Everything after the forward lists: the reverse-edge pass (static tests on the GPU, the order-dependent "row full yet"
// walk on the host), the optional constant-degree fill (GPU), and the flattened graph.  d_fwd [n x 2M] and d_deg [n] hold
// the forward lists of ALL vertices on this device and are modified; d_knn (the candidate lists of all vertices) is
// needed only when need_const_degree.
int gd_finish(int device, uint32_t* d_fwd, uint32_t* d_deg, uint64_t n, uint32_t M, int reverse, int need_const_degree,
              const uint32_t* d_knn, uint32_t kstride, uint32_t klen, uint64_t* out_offsets, uint32_t* out_edges, cudaStream_t st) {
    GBDR_CUDA(cudaSetDevice(device));
    const uint32_t cap = 2 * M;
    const uint32_t words = (M + M / 2 + 63) / 64;
    int sms = 148;
    GBDR_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    std::vector<uint32_t> g((size_t)n * cap), deg(n);
    if (reverse) {
        unsigned long long* d_mask = nullptr;
        uint32_t* d_indeg = nullptr;
        GBDR_CUDA(cudaMallocAsync((void**)&d_mask, (size_t)n * words * 8, st));
        GBDR_CUDA(cudaMallocAsync((void**)&d_indeg, (size_t)n * 4, st));
        GBDR_CUDA(cudaMemsetAsync(d_indeg, 0, (size_t)n * 4, st));
        const uint32_t mgrid = (uint32_t)std::min<uint64_t>((n + 7) / 8, (uint64_t)sms * 8);
        gd_mutual_kernel<<<mgrid, 256, 0, st>>>(d_fwd, d_deg, n, cap, words, d_mask, d_indeg);
        GBDR_CHECK_LAUNCH();
        count_launch();
        std::vector<unsigned long long> offer_mask((size_t)n * words);
        std::vector<uint32_t> indeg(n);
        GBDR_CUDA(cudaMemcpyAsync(offer_mask.data(), d_mask, (size_t)n * words * 8, cudaMemcpyDeviceToHost, st));
        GBDR_CUDA(cudaMemcpyAsync(indeg.data(), d_indeg, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
        GBDR_CUDA(cudaMemcpyAsync(g.data(), d_fwd, g.size() * 4, cudaMemcpyDeviceToHost, st));
        GBDR_CUDA(cudaMemcpyAsync(deg.data(), d_deg, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
        GBDR_CUDA(cudaStreamSynchronize(st));
        cudaFreeAsync(d_mask, st);
        cudaFreeAsync(d_indeg, st);
        // addReverseEdgesForGD, support_func.h:423-442, ascending i: only "row c not full yet" (:429) depends on the order
        for (uint64_t i = 0; i < n; ++i) {
            int thr = std::min((int)M - (int)indeg[i], (int)(M / 2));
            if (thr <= 0) continue;
            const uint32_t* row_i = g.data() + i * cap;
            for (uint32_t w = 0; w < words && thr > 0; ++w)
                for (unsigned long long m = offer_mask[i * words + w]; m; m &= m - 1) {
                    const uint32_t c = row_i[w * 64u + __builtin_ctzll(m)];
                    const uint32_t dc = deg[c];
                    if (dc < cap) {
                        g[(size_t)c * cap + dc] = (uint32_t)i;
                        deg[c] = dc + 1;
                        if (--thr <= 0) break;
                    }
                }
        }
        if (need_const_degree) {
            GBDR_CUDA(cudaMemcpyAsync(d_fwd, g.data(), g.size() * 4, cudaMemcpyHostToDevice, st));
            GBDR_CUDA(cudaMemcpyAsync(d_deg, deg.data(), (size_t)n * 4, cudaMemcpyHostToDevice, st));
        }
    }
    if (need_const_degree) {  // getConstantDegreeForGD, support_func.h:466-485
        if (!d_knn) {
            set_error("gd_finish: the constant-degree fill needs the candidate lists");
            return GBDR_E_INVALID;
        }
        const uint32_t fgrid = (uint32_t)std::min<uint64_t>((n + 7) / 8, (uint64_t)sms * 8);
        gd_fill_kernel<<<fgrid, 256, 0, st>>>(d_fwd, d_deg, n, cap, d_knn, kstride, klen);
        GBDR_CHECK_LAUNCH();
        count_launch();
    }
    if (!reverse || need_const_degree) {
        GBDR_CUDA(cudaMemcpyAsync(g.data(), d_fwd, g.size() * 4, cudaMemcpyDeviceToHost, st));
        GBDR_CUDA(cudaMemcpyAsync(deg.data(), d_deg, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
        GBDR_CUDA(cudaStreamSynchronize(st));
    }
    out_offsets[0] = 0;
    for (uint64_t i = 0; i < n; ++i) {
        memcpy(out_edges + out_offsets[i], g.data() + i * cap, (size_t)deg[i] * 4);
        out_offsets[i + 1] = out_offsets[i] + deg[i];
    }
    return GBDR_OK;
}

// host-buffer entry point (gbdr_gd_prune): upload, forward lists, finish
int gd_prune_device(int device, const uint64_t* knn_offsets, const uint32_t* knn_edges, const float* db_low,
                    uint64_t n, uint32_t d_low, uint32_t M, int reverse, int need_const_degree,
                    uint64_t* out_offsets, uint32_t* out_edges, double* gpu_seconds, uint32_t cut_k) {
    GBDR_CUDA(cudaSetDevice(device));
    if (n == 0) {
        out_offsets[0] = 0;
        return GBDR_OK;
    }
    // GBDR_GD_TIMING=1: wall time of each host stage on stderr
    const bool timing = getenv("GBDR_GD_TIMING") != nullptr;
    auto t_last = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!timing) return;
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[gd_prune] %-28s %8.3f s\n", what, std::chrono::duration<double>(now - t_last).count());
        t_last = now;
    };
    const uint32_t C = d_low / 4;
    uint64_t maxdeg = 0;
    for (uint64_t i = 0; i < n; ++i) maxdeg = std::max<uint64_t>(maxdeg, knn_offsets[i + 1] - knn_offsets[i]);
    const uint32_t klen = (uint32_t)std::max<uint64_t>(maxdeg, 1);
    // candidate lists -> [n x klen] matrix in HBM.  Fixed-length lists (the kNN-1k file: every row k ids) go up straight
    // from the caller's buffer; ragged lists are padded on the host first.
    bool uniform = true;
    const uint64_t len0 = knn_offsets[1] - knn_offsets[0];
    for (uint64_t i = 0; i < n && uniform; ++i) uniform = knn_offsets[i + 1] - knn_offsets[i] == len0;
    uniform = uniform && len0 > 0;
    std::vector<uint32_t> padded;
    if (!uniform) {
        padded.assign((size_t)n * klen, PAD_ID);
        for (uint64_t i = 0; i < n; ++i) {
            const uint64_t b = knn_offsets[i], e = knn_offsets[i + 1];
            memcpy(padded.data() + (size_t)i * klen, knn_edges + b, (size_t)(e - b) * 4);
        }
    }
    lap("pad ragged lists");
    uint32_t *d_knn = nullptr, *d_fwd = nullptr, *d_deg = nullptr, *d_misc = nullptr;
    float* d_db = nullptr;
    cudaStream_t st = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    int rc = GBDR_OK;
    auto fail = [&](cudaError_t e, const char* what) {
        set_error(std::string(what) + ": " + cudaGetErrorString(e));
        rc = GBDR_E_CUDA;
    };
#define GD_TRY(x)                          \
    if (rc == GBDR_OK) {                   \
        cudaError_t _e = (x);              \
        if (_e != cudaSuccess) fail(_e, #x); \
    }
    GD_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    GD_TRY(cudaEventCreate(&e0));
    GD_TRY(cudaEventCreate(&e1));
    GD_TRY(cudaMalloc((void**)&d_knn, (size_t)n * klen * 4));
    GD_TRY(cudaMalloc((void**)&d_db, (size_t)n * C * 16 + 16));
    const uint32_t out_stride = cut_k ? cut_k : 2 * M;
    GD_TRY(cudaMalloc((void**)&d_fwd, (size_t)n * out_stride * 4));
    GD_TRY(cudaMalloc((void**)&d_deg, (size_t)n * 4));
    GD_TRY(cudaMalloc((void**)&d_misc, 8));
    GD_TRY(cudaEventRecord(e0, st));
    GD_TRY(cudaMemcpyAsync(d_knn, uniform ? knn_edges + knn_offsets[0] : padded.data(), (size_t)n * klen * 4,
                           cudaMemcpyHostToDevice, st));
    if (rc == GBDR_OK) {
        if (d_low % 4 == 0) {
            GD_TRY(cudaMemcpyAsync(d_db, db_low, (size_t)n * d_low * 4, cudaMemcpyHostToDevice, st));
        } else {
            GD_TRY(cudaMemcpy2DAsync(d_db, (size_t)C * 16, db_low, (size_t)d_low * 4, (size_t)C * 16, n,
                                     cudaMemcpyHostToDevice, st));
        }
    }
    GD_TRY(cudaMemsetAsync(d_misc, 0, 8, st));
    int sms = 148;
    GD_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    if (rc == GBDR_OK) rc = gd_check_ids(d_knn, n, klen, klen, n, d_misc + 1, st);
    lap("upload + validate ids");
    if (rc == GBDR_OK) rc = gd_forward_launch(d_knn, klen, klen, 0, n, d_db, C, M, d_fwd, d_deg, d_misc, sms, st, cut_k);
    if (rc == GBDR_OK && !cut_k)
        rc = gd_finish(device, d_fwd, d_deg, n, M, reverse, need_const_degree, d_knn, klen, klen, out_offsets, out_edges, st);
    if (rc == GBDR_OK && cut_k) {  // the rows are the graph: download and flatten
        std::vector<uint32_t> g((size_t)n * cut_k), deg(n);
        GD_TRY(cudaMemcpyAsync(g.data(), d_fwd, g.size() * 4, cudaMemcpyDeviceToHost, st));
        GD_TRY(cudaMemcpyAsync(deg.data(), d_deg, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
        GD_TRY(cudaStreamSynchronize(st));
        if (rc == GBDR_OK) {
            out_offsets[0] = 0;
            for (uint64_t i = 0; i < n; ++i) {
                memcpy(out_edges + out_offsets[i], g.data() + i * cut_k, (size_t)deg[i] * 4);
                out_offsets[i + 1] = out_offsets[i] + deg[i];
            }
        }
    }
    GD_TRY(cudaEventRecord(e1, st));
    GD_TRY(cudaStreamSynchronize(st));
    if (rc == GBDR_OK && gpu_seconds) {
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        *gpu_seconds = ms * 1e-3;
    }
#undef GD_TRY
    if (st) cudaStreamSynchronize(st);
    for (void* q : {(void*)d_knn, (void*)d_db, (void*)d_fwd, (void*)d_deg, (void*)d_misc})
        if (q) cudaFree(q);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (st) cudaStreamDestroy(st);
    lap("prune + reverse pass + output");
    return rc;
}

}  // namespace gbdr
