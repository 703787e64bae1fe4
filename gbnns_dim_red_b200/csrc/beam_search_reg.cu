// beam_search_reg.cu — K2, register-resident variant for ef + slack <= 256.
//
// Same semantics as beam_search.cu (reference search/search_function.h:15-102, see the header
// comment there) but the sorted result/candidate list lives in registers: lane l holds the R
// consecutive entries [l*R, l*R+R).  One sequential insertion (makeStep's accept rule, :31-36) is
//   R ballots (rank of the new element) + one shuffle-up of each lane's last entry + 2R selects,
// ~25 instructions for R = 2 instead of the ~150 of the shared-memory list, which the first ncu
// profile showed to be >40 % of all issued instructions (profiles/r1a_*).
#include "beam_reglist.cuh"

namespace gbdr {

namespace {

template <int R, int C_T>
__global__ void __launch_bounds__(256, 2) beam_search_reg_kernel(const BeamParams p, uint32_t* __restrict__ counter) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const uint32_t C = C_T ? (uint32_t)C_T : p.C;
    const BeamLayout Lo = beam_layout(C, 0, p.hcap);
    unsigned char* wbase = smem_raw + (size_t)warp * p.smem_per_warp;
    float* stage = reinterpret_cast<float*>(wbase + Lo.stage_off);
    float* qs = reinterpret_cast<float*>(wbase + Lo.q_off);
    uint32_t* nbr = reinterpret_cast<uint32_t*>(wbase + Lo.nbr_off);
    uint32_t* vis = reinterpret_cast<uint32_t*>(wbase + Lo.vis_off);
    const uint32_t gwarp = blockIdx.x * (blockDim.x >> 5) + warp;
    uint32_t* spill = p.spill + (size_t)gwarp * p.spill_cap;
    const int ef = (int)p.ef;
    constexpr int CAP = 32 * R;
    const float INF = __int_as_float(0x7f800000);
    uint32_t status_acc = 0;

    for (;;) {
        uint32_t qi = 0;
        if (lane == 0) qi = atomicAdd(counter, 1u);
        qi = __shfl_sync(FULL_MASK, qi, 0);
        if (qi >= p.n_q) break;

        // ---- per-query init ----
        {
            uint4 fill = make_uint4(PAD_ID, PAD_ID, PAD_ID, PAD_ID);
            for (uint32_t i = lane; i < p.hcap / 4; i += 32) reinterpret_cast<uint4*>(vis)[i] = fill;
        }
        const float* qg = p.q + (size_t)qi * p.q_stride;
        for (uint32_t c = lane; c < C; c += 32)
            reinterpret_cast<float4*>(qs)[c] = __ldg(reinterpret_cast<const float4*>(qg) + c);
        __syncwarp();
        float4 qreg[C_T ? C_T : 1];
        if (C_T) {
#pragma unroll
            for (int c = 0; c < (C_T ? C_T : 1); ++c) qreg[c] = reinterpret_cast<const float4*>(qs)[c];
        }

        RegList<R> L;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            L.d[r] = INF;
            L.i[r] = PAD_ID;
        }
        int size = 0;
        float worst = INF;  // dist of entry ef-1, valid when size >= ef
        int hops = 0, dist_calc = 1, scanned = 0;  // dist_calc starts at 1 (search_function.h:52)
        uint32_t vcount = 0, scount = 0;
        bool spill_ready = false, failed = false;

        // distances of nbr[b0 .. b0+mb) -> lane r holds the distance of row r
        auto batch_dist = [&](int b0, int mb) -> float {
            if (C_T == 8) {
                // 8 lanes per 128-byte row, 4 rows per pass
                const uint32_t c = lane & 7;
                for (int r0 = 0; r0 < mb; r0 += 4) {
                    const int r = r0 + (lane >> 3);
                    if (r < mb)
                        cp_async16(stage + ((size_t)r * 8 + (c ^ (r & 7u))) * 4u,
                                   p.db + (size_t)nbr[b0 + r] * p.row_stride + c * 4u);
                }
            } else {
                const uint32_t T = (uint32_t)mb * C;
                for (uint32_t t = lane; t < T; t += 32) {
                    uint32_t r = C_T ? t / (uint32_t)(C_T ? C_T : 1) : t / C;
                    uint32_t c = t - r * C;
                    cp_async16(stage + ((size_t)r * C + swz<C_T>(r, c, C)) * 4u,
                               p.db + (size_t)nbr[b0 + r] * p.row_stride + c * 4u);
                }
            }
            cp_async_commit();
            cp_async_wait<0>();
            __syncwarp();
            L2Acc acc;
            if (lane < mb) {
                const float4* row = reinterpret_cast<const float4*>(stage) + (size_t)lane * C;
                if (C_T) {
#pragma unroll
                    for (int c = 0; c < (C_T ? C_T : 1); ++c) acc.add(qreg[c], row[swz<C_T>(lane, c, C)]);
                } else {
                    for (uint32_t c = 0; c < C; ++c)
                        acc.add(reinterpret_cast<const float4*>(qs)[c], row[swz<C_T>(lane, c, C)]);
                }
            }
            __syncwarp();
            return acc.result();
        };

        // ---- entry point (search_function.h:56-64) ----
        {
            uint32_t e = __ldg(p.entry + qi);
            if (e >= p.n_vertices) {  // not a vertex: the query fails (PAD results) instead of reading out of bounds
                e = 0;
                failed = true;
                status_acc |= BEAM_ST_BAD_ENTRY;
            }
            if (lane == 0) {
                nbr[0] = e;
                vis[(e * 0x9E3779B1u) >> p.hshift] = e;
            }
            __syncwarp();
            float d0 = batch_dist(0, 1);
            d0 = __shfl_sync(FULL_MASK, d0, 0);
            list_insert_reg<R>(L, size, d0, e, lane);
            if (size >= ef) worst = list_get_d<R>(L, ef - 1);
            vcount = 1;
        }

        // ---- main loop (search_function.h:65-91) ----
        while (!failed) {
            // best un-expanded entry = top of candidateSet
            int best = 0x7fffffff;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const bool u = (lane * R + r < size) && !(L.i[r] & EXPANDED);
                const unsigned m = __ballot_sync(FULL_MASK, u);
                if (m) best = min(best, (__ffs(m) - 1) * R + r);
            }
            if (best == 0x7fffffff) break;  // candidateSet empty, or its best is worse than worst (:65,:67)
            int csel = best;
            if (best + 1 < size) {
                // ties on dist: the reference pops the largest id first (max-heap of (-dist,id))
                const float dsel = list_get_d<R>(L, best);
                for (int j = best + 1; j < size; ++j) {
                    if (list_get_d<R>(L, j) != dsel) break;
                    if (!(list_get_i<R>(L, j) & EXPANDED)) csel = j;
                }
            }
            const uint32_t node = list_get_i<R>(L, csel) & ID_MASK;
#pragma unroll
            for (int r = 0; r < R; ++r)
                if (lane * R + r == csel) L.i[r] |= EXPANDED;

            // ---- makeStep over the adjacency row, 64 ids at a time (:23-39) ----
            const uint32_t* arow = p.adj + (size_t)node * p.adj_stride;
            for (uint32_t cb = 0; cb < p.adj_stride; cb += 64) {
                const uint32_t a0 = __ldg(arow + cb + lane);
                const uint32_t a1 = (cb + 32 < p.adj_stride) ? __ldg(arow + cb + 32 + lane) : PAD_ID;
                const unsigned v0 = __ballot_sync(FULL_MASK, a0 != PAD_ID);
                const unsigned v1 = __ballot_sync(FULL_MASK, a1 != PAD_ID);
                scanned += __popc(v0) + __popc(v1);
                if ((v0 | v1) == 0) break;

                const bool smem_open = vcount + 64 <= p.hlimit;
                if (!smem_open) {
                    if (!spill_ready) {
                        for (uint32_t i = lane; i < p.spill_cap; i += 32) spill[i] = PAD_ID;
                        __syncwarp();
                        spill_ready = true;
                        status_acc |= BEAM_ST_SPILLED;
                    }
                    if (scount + 64 > (p.spill_cap >> 1) + (p.spill_cap >> 2)) {
                        failed = true;
                        status_acc |= BEAM_ST_VISITED_FULL;
                        break;
                    }
                }
                bool n0 = false, n1 = false;
                if (a0 != PAD_ID) n0 = visit(vis, p.hcap, p.hshift, smem_open, spill, p.spill_cap, p.spill_shift, a0);
                __syncwarp();
                if (v1) {
                    if (a1 != PAD_ID) n1 = visit(vis, p.hcap, p.hshift, smem_open, spill, p.spill_cap, p.spill_shift, a1);
                    __syncwarp();
                }
                const unsigned m0 = __ballot_sync(FULL_MASK, n0);
                const unsigned m1 = __ballot_sync(FULL_MASK, n1);
                const int c0 = __popc(m0), mtot = c0 + __popc(m1);
                if (smem_open) vcount += mtot; else scount += mtot;
                if (n0) nbr[__popc(m0 & lanemask_lt())] = a0;
                if (n1) nbr[c0 + __popc(m1 & lanemask_lt())] = a1;
                __syncwarp();
                dist_calc += mtot;  // :29

                for (int b0 = 0; b0 < mtot; b0 += 32) {
                    const int mb = min(32, mtot - b0);
                    const float dist = batch_dist(b0, mb);
                    const uint32_t myid = lane < mb ? nbr[b0 + lane] : 0u;
                    // pre-filter with the worst at the start of the batch (worst never increases)
                    unsigned am = __ballot_sync(FULL_MASK, lane < mb && (size < ef || worst > dist));
                    while (am) {
                        const int src = __ffs(am) - 1;
                        am &= am - 1;
                        const float x = __shfl_sync(FULL_MASK, dist, src);
                        const uint32_t xid = __shfl_sync(FULL_MASK, myid, src);
                        if (size >= ef && !(worst > x)) continue;  // :31
                        list_insert_reg<R>(L, size, x, xid, lane);  // :32-34
                        if (size >= ef) {
                            worst = list_get_d<R>(L, ef - 1);
                            if (size > ef) {
                                // :35-36 eviction; boundary ties (dist == new worst) stay in the slack
                                int keep = 0;
#pragma unroll
                                for (int r = 0; r < R; ++r) {
                                    const int e = lane * R + r;
                                    keep += __popc(__ballot_sync(FULL_MASK, e >= ef && e < size && L.d[r] == worst));
                                }
                                size = ef + keep;
                                if (size >= CAP) {
                                    failed = true;
                                    status_acc |= BEAM_ST_TIE_OVERFLOW;
                                }
                            }
                        }
                    }
                }
                if (failed) break;
                if (v1 != FULL_MASK) break;  // row ended inside this chunk
            }
            if (failed) break;
            ++hops;  // :90
        }

        // ---- emit the k best (:96-100) ----
        const int nres = min(min(size, ef), (int)p.k);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int e = lane * R + r;
            if (e < (int)p.k) {
                const bool ok = e < nres && !failed;
                p.out_ids[(size_t)qi * p.k + e] = ok ? (L.i[r] & ID_MASK) + p.id_offset : PAD_ID;
                if (p.out_dists) p.out_dists[(size_t)qi * p.k + e] = ok ? L.d[r] : INF;
            }
        }
        if (lane == 0) {
            if (p.hops) p.hops[qi] = hops;
            if (p.dist_calc) p.dist_calc[qi] = dist_calc + p.dist_calc_bias;
            if (p.scanned) p.scanned[qi] = scanned;
        }
        __syncwarp();
    }
    if (status_acc && lane == 0) atomicOr(p.status, status_acc);
}

template <int R, int C_T>
int launch_rt(const BeamParams& p, uint32_t wpb, uint32_t blocks, uint32_t* counter, cudaStream_t st) {
    const size_t smem = (size_t)p.smem_per_warp * wpb;
    GBDR_CUDA(cudaFuncSetAttribute(beam_search_reg_kernel<R, C_T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    beam_search_reg_kernel<R, C_T><<<blocks, wpb * 32, smem, st>>>(p, counter);
    GBDR_CHECK_LAUNCH();
    count_launch();
    return GBDR_OK;
}

template <int R>
int launch_r(const BeamParams& p, uint32_t wpb, uint32_t blocks, uint32_t* counter, cudaStream_t st) {
    switch (p.C) {
        case 4: return launch_rt<R, 4>(p, wpb, blocks, counter, st);
        case 8: return launch_rt<R, 8>(p, wpb, blocks, counter, st);
        case 16: return launch_rt<R, 16>(p, wpb, blocks, counter, st);
        default: return launch_rt<R, 0>(p, wpb, blocks, counter, st);
    }
}

}  // namespace

// p.cap selects the variant: 32, 64, 128 or 256 list slots held in registers
int launch_beam_search_reg(const BeamParams& p, uint32_t wpb, uint32_t blocks, cudaStream_t st) {
    uint32_t* counter = p.status + 1;
    switch (p.cap) {
        case 32: return launch_r<1>(p, wpb, blocks, counter, st);
        case 64: return launch_r<2>(p, wpb, blocks, counter, st);
        case 128: return launch_r<4>(p, wpb, blocks, counter, st);
        case 256: return launch_r<8>(p, wpb, blocks, counter, st);
        default:
            set_error("beam_search_reg: unsupported list capacity");
            return GBDR_E_INVALID;
    }
}

}  // namespace gbdr
