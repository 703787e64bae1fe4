// rerank.cu — K3: exact squared-L2 re-ranking of the low-dimensional survivors in the original
// dimension, fused gather + distance + top-k.
//
// Replaces getRealNearest (reference search/search_function.h:105-125).  The reference walks the
// low-dim heap from its worst element to its best and keeps the STRICTLY smaller exact distance
// (:117), i.e. arg-min with ties going to the worse low-dim rank.  Generalised here to top-k ordered
// by (exact dist asc, low-dim rank desc); out[0] is the reference's return value.
//
// One warp per query.  Candidates are processed 32 at a time: the warp gathers a [32 rows x 32 dims]
// tile with coalesced 16 B cp.async (8 lanes per 128 B row piece) into a swizzled shared tile,
// double-buffered along the dimension axis, and lane r accumulates row r chunk by chunk in the
// reference's summation order (common.cuh L2Acc).  HBM traffic is exactly ef*d*4 bytes per query
// plus the query itself.
#include "kernels.cuh"

namespace gbdr {

constexpr int RR_TILE_C = 8;  // float4 chunks per tile row (32 dims = 128 B)

__host__ __device__ inline uint32_t rerank_smem_per_warp(uint32_t C, uint32_t m) {
    // 2 stage tiles [32 x 8 chunks] + query [C chunks] + (dist, rank) per candidate
    uint32_t b = 2u * 32u * RR_TILE_C * 16u + C * 16u + ((m + 31u) & ~31u) * 4u;
    return (b + 15u) & ~15u;
}

__global__ void __launch_bounds__(256) rerank_kernel(const RerankParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const uint32_t qi = blockIdx.x * (blockDim.x >> 5) + warp;
    if (qi >= p.n_q) return;
    unsigned char* wb = smem_raw + (size_t)warp * p.smem_per_warp;
    float4* tile = reinterpret_cast<float4*>(wb);                       // [2][32][8]
    float4* qs = tile + 2 * 32 * RR_TILE_C;                             // [C]
    float* dist_s = reinterpret_cast<float*>(qs + p.C);                 // [m]
    const uint32_t C = p.C;

    const float4* qg = reinterpret_cast<const float4*>(p.queries + (size_t)qi * p.q_stride);
    for (uint32_t c = lane; c < C; c += 32) qs[c] = __ldg(qg + c);
    const uint32_t* cand = p.cand + (size_t)qi * p.m;

    // number of valid candidates (PAD-terminated)
    uint32_t mv = 0;
    for (uint32_t b = 0; b < p.m; b += 32) {
        uint32_t id = (b + lane < p.m) ? __ldg(cand + b + lane) : PAD_ID;
        mv += __popc(__ballot_sync(FULL_MASK, id != PAD_ID));
    }
    __syncwarp();

    const uint32_t ntile = (C + RR_TILE_C - 1) / RR_TILE_C;
    for (uint32_t b0 = 0; b0 < mv; b0 += 32) {
        const uint32_t mb = min(32u, mv - b0);
        const uint32_t myid = lane < mb ? __ldg(cand + b0 + lane) : 0u;
        auto issue = [&](uint32_t t) {
            const uint32_t cbeg = t * RR_TILE_C;
            const uint32_t cw = min((uint32_t)RR_TILE_C, C - cbeg);
            float4* dst = tile + (t & 1) * 32 * RR_TILE_C;
            // 8 lanes per row piece, 4 rows per pass; the loop is warp-uniform (shuffle inside)
            for (uint32_t r0 = 0; r0 < mb; r0 += 4) {
                const uint32_t r = r0 + (lane >> 3);
                const uint32_t c = lane & 7;
                const uint32_t rowid = __shfl_sync(FULL_MASK, myid, r & 31);
                if (r < mb && c < cw)
                    cp_async16(dst + r * RR_TILE_C + (c ^ (r & 7)),
                               p.db + (size_t)rowid * p.row_stride + (size_t)(cbeg + c) * 4u);
            }
            cp_async_commit();
        };
        L2Acc acc;
        issue(0);
        for (uint32_t t = 0; t < ntile; ++t) {
            if (t + 1 < ntile) {
                issue(t + 1);
                cp_async_wait<1>();
            } else {
                cp_async_wait<0>();
            }
            __syncwarp();
            if (lane < mb) {
                const float4* row = tile + (t & 1) * 32 * RR_TILE_C + lane * RR_TILE_C;
                const uint32_t cbeg = t * RR_TILE_C;
                const uint32_t cw = min((uint32_t)RR_TILE_C, C - cbeg);
                // reference argument order: Dist(point_i, point_q) (search_function.h:110,116)
                for (uint32_t c = 0; c < cw; ++c) acc.add(row[c ^ (lane & 7)], qs[cbeg + c]);
            }
            __syncwarp();
        }
        if (lane < mb) dist_s[b0 + lane] = acc.result();
    }
    __syncwarp();

    // top-k by (dist asc, low-dim rank desc)
    if (p.k == 1) {
        float best = __int_as_float(0x7f800000);
        int brank = -1;
        for (uint32_t j = lane; j < mv; j += 32) {
            float dj = dist_s[j];
            if (dj < best || (dj == best && (int)j > brank)) {
                best = dj;
                brank = (int)j;
            }
        }
        for (int o = 16; o; o >>= 1) {
            float ob = __shfl_xor_sync(FULL_MASK, best, o);
            int orank = __shfl_xor_sync(FULL_MASK, brank, o);
            if (orank >= 0 && (brank < 0 || ob < best || (ob == best && orank > brank))) {
                best = ob;
                brank = orank;
            }
        }
        if (lane == 0) {
            p.out_ids[qi] = brank >= 0 ? __ldg(cand + brank) + p.id_offset : PAD_ID;
            if (p.out_dists) p.out_dists[qi] = brank >= 0 ? best : __int_as_float(0x7f800000);
        }
    } else {
        for (uint32_t j = lane; j < p.k; j += 32) {
            if (j >= mv) {
                p.out_ids[(size_t)qi * p.k + j] = PAD_ID;
                if (p.out_dists) p.out_dists[(size_t)qi * p.k + j] = __int_as_float(0x7f800000);
            }
        }
        for (uint32_t j = lane; j < mv; j += 32) {
            const float dj = dist_s[j];
            uint32_t rank = 0;
            for (uint32_t i = 0; i < mv; ++i) {
                const float di = dist_s[i];
                rank += (di < dj || (di == dj && i > j)) ? 1u : 0u;
            }
            if (rank < p.k) {
                p.out_ids[(size_t)qi * p.k + rank] = __ldg(cand + j) + p.id_offset;
                if (p.out_dists) p.out_dists[(size_t)qi * p.k + rank] = dj;
            }
        }
    }
}

int launch_rerank(const RerankParams& p_in, cudaStream_t st) {
    RerankParams p = p_in;
    p.smem_per_warp = rerank_smem_per_warp(p.C, p.m);
    uint32_t wpb = 8;
    while (wpb > 1 && (size_t)wpb * p.smem_per_warp > 200u * 1024u) wpb >>= 1;
    const size_t smem = (size_t)wpb * p.smem_per_warp;
    if (smem > 227u * 1024u) {
        set_error("rerank: per-query shared memory exceeds 227 KB (d or ef too large)");
        return GBDR_E_CAPACITY;
    }
    GBDR_CUDA(cudaFuncSetAttribute(rerank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const uint32_t blocks = (p.n_q + wpb - 1) / wpb;
    rerank_kernel<<<blocks, wpb * 32, smem, st>>>(p);
    GBDR_CHECK_LAUNCH();
    count_launch();
    return GBDR_OK;
}

}  // namespace gbdr
