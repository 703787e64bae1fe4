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
    const uint32_t* knn;       // [n x kstride] padded candidate lists (PAD tail)
    uint32_t kstride;
    const float* db;           // [n x C*4]
    uint32_t C;
    uint64_t n;
    uint32_t M;
    uint32_t sort_cap;         // power of two >= max candidates
    uint32_t fwd_stride;       // M + M/2
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
            self[c] = __ldg(reinterpret_cast<const float4*>(p.db + (size_t)vi * C * 4) + c);
        for (uint32_t i = lane; i < p.sort_cap; i += 32) {
            sd[i] = INF;
            si[i] = PAD_ID;
        }
        __syncwarp();

        // ---- A: distances to candidates, 32 at a time (:532-539) ----
        uint32_t m = 0;
        for (uint32_t b0 = 0; b0 < p.kstride; b0 += 32) {
            const uint32_t cid = __ldg(cl + b0 + lane);
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
                ok = dist > eps;  // :535
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
                                                        uint64_t n, uint32_t stride, unsigned long long* __restrict__ mask,
                                                        uint32_t* __restrict__ indeg) {
    const int lane = threadIdx.x & 31;
    const uint64_t nwarps = (uint64_t)gridDim.x * (blockDim.x >> 5);
    for (uint64_t i = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < n; i += nwarps) {
        const uint32_t di = deg[i];
        unsigned long long m = 0;
        for (uint32_t j0 = 0; j0 < di; j0 += 32) {
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
            m |= (unsigned long long)__ballot_sync(FULL_MASK, offer) << j0;
        }
        if (lane == 0) mask[i] = m;
    }
}

}  // namespace

int gd_prune_device(int device, const uint64_t* knn_offsets, const uint32_t* knn_edges, const float* db_low,
                    uint64_t n, uint32_t d_low, uint32_t M, int reverse, int need_const_degree,
                    uint64_t* out_offsets, uint32_t* out_edges, double* gpu_seconds) {
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
    uint32_t sort_cap = 32;
    while (sort_cap < maxdeg) sort_cap <<= 1;
    const uint32_t kstride = (uint32_t)((maxdeg + 31) / 32 * 32 ? (maxdeg + 31) / 32 * 32 : 32);
    GdParams p;
    memset(&p, 0, sizeof(p));
    p.smem_per_warp = gd_smem_per_warp(C, M, sort_cap);
    if (p.smem_per_warp > 220u * 1024u) {
        set_error("gd_prune: candidate lists too long for the shared-memory sort (max degree too large)");
        return GBDR_E_CAPACITY;
    }
    uint32_t wpb = std::max<uint32_t>(1, std::min<uint32_t>(4, (200u * 1024u) / p.smem_per_warp));
    const size_t smem = (size_t)wpb * p.smem_per_warp;

    // candidate lists -> [n x kstride] matrix in HBM.  Fixed-length lists (the kNN-1k file: every row k ids) go up
    // straight from the caller's buffer with a pitched copy over a PAD-filled device matrix; ragged lists are
    // padded on the host first.
    bool uniform = true;
    const uint64_t len0 = knn_offsets[1] - knn_offsets[0];
    for (uint64_t i = 0; i < n && uniform; ++i) uniform = knn_offsets[i + 1] - knn_offsets[i] == len0;
    uniform = uniform && len0 > 0;
    {
        // every id indexes db_low on the device: validate them all (a few host threads; 4 GB at 1M x 1000)
        const uint64_t total = knn_offsets[n] - knn_offsets[0];
        const uint32_t* ids = knn_edges + knn_offsets[0];
        const unsigned nt = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
        std::vector<std::thread> pool;
        std::atomic<bool> bad(false);
        for (unsigned t = 0; t < nt; ++t)
            pool.emplace_back([&, t]() {
                const uint64_t b = total * t / nt, e = total * (t + 1) / nt;
                uint32_t mx = 0;
                for (uint64_t j = b; j < e; ++j) mx = std::max(mx, ids[j]);
                if (e > b && mx >= n) bad = true;
            });
        for (auto& th : pool) th.join();
        if (bad) {
            set_error("gd_prune: candidate id out of range");
            return GBDR_E_INVALID;
        }
    }
    lap("validate ids");
    std::vector<uint32_t> padded;
    if (!uniform) {
        padded.assign((size_t)n * kstride, PAD_ID);
        for (uint64_t i = 0; i < n; ++i) {
            const uint64_t b = knn_offsets[i], e = knn_offsets[i + 1];
            memcpy(padded.data() + (size_t)i * kstride, knn_edges + b, (size_t)(e - b) * 4);
        }
    }
    // forward lists (<= M + M/2 entries) are written at the final row stride 2M, so the download IS the graph
    const uint32_t fwd_stride = 2 * M;
    const bool gpu_masks = reverse && M + M / 2 <= 64;  // one 64-bit offer mask per vertex
    uint32_t *d_knn = nullptr, *d_fwd = nullptr, *d_deg = nullptr, *d_counter = nullptr, *d_indeg = nullptr;
    unsigned long long* d_mask = nullptr;
    std::vector<unsigned long long> offer_mask(gpu_masks ? n : 0);
    std::vector<uint32_t> indeg(reverse ? n : 0, 0);
    float* d_db = nullptr;
    cudaStream_t st = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    int rc = GBDR_OK;
    std::vector<uint32_t> fwd((size_t)n * fwd_stride), deg(n);
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
    GD_TRY(cudaMalloc((void**)&d_knn, (size_t)n * kstride * 4));
    GD_TRY(cudaMalloc((void**)&d_db, (size_t)n * C * 16 + 16));
    GD_TRY(cudaMalloc((void**)&d_fwd, fwd.size() * 4));
    GD_TRY(cudaMalloc((void**)&d_deg, (size_t)n * 4));
    GD_TRY(cudaMalloc((void**)&d_counter, 4));
    if (gpu_masks) {
        GD_TRY(cudaMalloc((void**)&d_mask, (size_t)n * 8));
        GD_TRY(cudaMalloc((void**)&d_indeg, (size_t)n * 4));
        GD_TRY(cudaMemsetAsync(d_indeg, 0, (size_t)n * 4, st));
    }
    GD_TRY(cudaEventRecord(e0, st));
    if (uniform) {
        if (len0 != kstride) GD_TRY(cudaMemsetAsync(d_knn, 0xFF, (size_t)n * kstride * 4, st));  // PAD_ID = 0xFFFFFFFF
        GD_TRY(cudaMemcpy2DAsync(d_knn, (size_t)kstride * 4, knn_edges + knn_offsets[0], (size_t)len0 * 4, (size_t)len0 * 4, n,
                                 cudaMemcpyHostToDevice, st));
    } else {
        GD_TRY(cudaMemcpyAsync(d_knn, padded.data(), padded.size() * 4, cudaMemcpyHostToDevice, st));
    }
    if (rc == GBDR_OK) {
        if (d_low % 4 == 0) {
            GD_TRY(cudaMemcpyAsync(d_db, db_low, (size_t)n * d_low * 4, cudaMemcpyHostToDevice, st));
        } else {
            GD_TRY(cudaMemcpy2DAsync(d_db, (size_t)C * 16, db_low, (size_t)d_low * 4, (size_t)C * 16, n,
                                     cudaMemcpyHostToDevice, st));
        }
    }
    GD_TRY(cudaMemsetAsync(d_counter, 0, 4, st));
    if (rc == GBDR_OK) {
        p.knn = d_knn; p.kstride = kstride; p.db = d_db; p.C = C; p.n = n; p.M = M; p.sort_cap = sort_cap;
        p.fwd_stride = fwd_stride; p.fwd = d_fwd; p.deg = d_deg; p.counter = d_counter;
        GD_TRY(cudaFuncSetAttribute(gd_prune_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cudaDeviceProp prop;
        GD_TRY(cudaGetDeviceProperties(&prop, device));
        if (rc == GBDR_OK) {
            const uint32_t per_sm = std::max<uint32_t>(1, (uint32_t)((227u * 1024u) / (smem + 1024)));
            const uint32_t grid = (uint32_t)std::min<uint64_t>((n + wpb - 1) / wpb, (uint64_t)prop.multiProcessorCount * per_sm);
            gd_prune_kernel<<<grid, wpb * 32, smem, st>>>(p);
            GD_TRY(cudaGetLastError());
            count_launch();
            if (gpu_masks && rc == GBDR_OK) {
                const uint32_t mgrid = (uint32_t)std::min<uint64_t>((n + 7) / 8, (uint64_t)prop.multiProcessorCount * 8);
                gd_mutual_kernel<<<mgrid, 256, 0, st>>>(d_fwd, d_deg, n, fwd_stride, d_mask, d_indeg);
                GD_TRY(cudaGetLastError());
                count_launch();
            }
        }
    }
    if (gpu_masks) {
        GD_TRY(cudaMemcpyAsync(offer_mask.data(), d_mask, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
        GD_TRY(cudaMemcpyAsync(indeg.data(), d_indeg, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    }
    GD_TRY(cudaMemcpyAsync(fwd.data(), d_fwd, fwd.size() * 4, cudaMemcpyDeviceToHost, st));
    GD_TRY(cudaMemcpyAsync(deg.data(), d_deg, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    GD_TRY(cudaEventRecord(e1, st));
    GD_TRY(cudaStreamSynchronize(st));
    if (rc == GBDR_OK && gpu_seconds) {
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        *gpu_seconds = ms * 1e-3;
    }
#undef GD_TRY
    if (d_knn) cudaFree(d_knn);
    if (d_db) cudaFree(d_db);
    if (d_fwd) cudaFree(d_fwd);
    if (d_deg) cudaFree(d_deg);
    if (d_counter) cudaFree(d_counter);
    if (d_mask) cudaFree(d_mask);
    if (d_indeg) cudaFree(d_indeg);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (st) cudaStreamDestroy(st);
    if (rc != GBDR_OK) return rc;
    lap("upload + prune kernel + d2h");

    // ---- host: sequential reverse pass and optional constant-degree fill ----
    const uint32_t cap = 2 * M;
    std::vector<uint32_t>& g = fwd;  // rows already at stride cap
    lap("unpack forward lists");
    if (reverse && gpu_masks) {  // addReverseEdgesForGD, support_func.h:402-445, with the static tests done on the GPU
        for (uint64_t i = 0; i < n; ++i) {  // :423-442, ascending i: only "row c not full yet" (:429) depends on the order
            int thr = std::min((int)M - (int)indeg[i], (int)(M / 2));
            if (thr <= 0) continue;
            const uint32_t* row_i = g.data() + i * cap;
            for (unsigned long long m = offer_mask[i]; m; m &= m - 1) {
                const uint32_t c = row_i[__builtin_ctzll(m)];
                const uint32_t dc = deg[c];
                if (dc < cap) {
                    g[(size_t)c * cap + dc] = (uint32_t)i;
                    deg[c] = dc + 1;
                    if (--thr <= 0) break;
                }
            }
        }
    } else if (reverse) {  // forward lists longer than 64 entries: the reference's loop as it stands
        for (uint64_t i = 0; i < n; ++i)
            for (uint32_t j = 0; j < deg[i]; ++j) indeg[g[i * cap + j]]++;  // :418-422
        for (uint64_t i = 0; i < n; ++i) {  // :423-442
            const int upper = (int)M - (int)indeg[i];
            int thr = std::min(upper, (int)(M / 2));
            if (thr <= 0) continue;
            for (uint32_t j = 0; j < deg[i]; ++j) {
                const uint32_t c = g[i * cap + j];
                if (deg[c] < cap) {
                    uint32_t* row = g.data() + (size_t)c * cap;
                    if (std::find(row, row + deg[c], (uint32_t)i) == row + deg[c]) {
                        row[deg[c]++] = (uint32_t)i;
                        if (--thr <= 0) break;
                    }
                }
            }
        }
    }
    lap("reverse pass");
    if (need_const_degree) {  // getConstantDegreeForGD, support_func.h:466-485
        for (uint64_t i = 0; i < n; ++i) {
            if (deg[i] >= cap) continue;
            uint32_t* row = g.data() + i * cap;
            for (uint64_t j = knn_offsets[i] + 1; j < knn_offsets[i + 1]; ++j) {
                if (std::find(row, row + deg[i], knn_edges[j]) == row + deg[i]) {
                    row[deg[i]++] = knn_edges[j];
                    if (deg[i] == cap) break;
                }
            }
        }
    }
    out_offsets[0] = 0;
    for (uint64_t i = 0; i < n; ++i) {
        memcpy(out_edges + out_offsets[i], g.data() + i * cap, (size_t)deg[i] * 4);
        out_offsets[i + 1] = out_offsets[i] + deg[i];
    }
    lap("const degree + output");
    return GBDR_OK;
}

}  // namespace gbdr
