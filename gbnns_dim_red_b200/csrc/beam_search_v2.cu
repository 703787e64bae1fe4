// beam_search_v2.cu — K2, second-generation register-list kernel for d_low in {16,32,48,64}.
//
// Same results as beam_search.cu / beam_search_reg.cu (reference search/search_function.h:15-102,
// bit-exact ids, distances, hops and dist_calc) with the per-hop instruction count cut ~2.5x and the
// per-warp footprint cut so that 24 instead of 16 warps (queries) are resident per SM.  The first ncu
// capture (profiles/r1b_*) showed the previous kernel issue-bound on bookkeeping, not HBM-bound:
// ~1000 warp instructions per hop, ~45 % of them in the one-at-a-time sorted insertion.
//
//   * Batched insertion.  All candidates of a hop that pass makeStep's accept test against the
//     worst distance at the start of the hop are merged into the sorted register list in one step:
//     every candidate gets its rank among list entries (ballot/popc) and among the other candidates,
//     every list entry its shift, then everything moves through a 512-byte shared scratch.  This is
//     exact because the result of the reference's sequential insert/evict sequence depends only on
//     the set of (dist,id) pairs unless two distances compare equal; any equality
//     seen while ranking, a tie across the ef boundary, or slack already in use makes the hop fall
//     back to the sequential path (same code as beam_search_reg.cu), so tie semantics are unchanged.
//   * Two lanes per row.  The four lane-strided partial sums of L2Metric::Dist are independent
//     chains, so lane 2r accumulates (s0,s1) and lane 2r+1 (s2,s3) of row r over all chunks in
//     order; two shuffles bring (s2,s3) over for the reference's final ((s0+s1)+s2)+s3.  Half the
//     FP instructions per row, still the exact bit pattern.
//   * Speculative adjacency prefetch.  The adjacency row of the node most likely to be expanded next
//     (second-best unexpanded entry, or the best new candidate as soon as its distance is known) is
//     loaded into registers while the current hop is still ranking/merging, which takes one of the
//     two dependent DRAM round trips per hop off the critical path.  A wrong guess costs one
//     128-byte read and nothing else: results never depend on it.
//   * Row gather by the TMA engine.  Each lane issues ONE cp.async.bulk (UBLKCP) for the whole
//     16*C-byte row of its candidate; completion is signalled on a per-warp mbarrier.  The second ncu
//     capture (profiles/r1c_*) showed ~130 of ~940 instructions per hop spent computing cp.async
//     addresses.  Rows land 16 bytes apart-padded so the two-lanes-per-row LDS.64 pattern is
//     bank-conflict free without a software swizzle.
//   * Visited set without atomics.  Ids of one adjacency chunk are distinct, so claiming an empty
//     slot is store / __syncwarp / read-back: the lane that reads its own id back owns the slot, a
//     loser keeps probing.  Replaces a divergent atomicCAS loop (~150 instructions per hop).
//   * The (dist,id) list is mirrored in shared memory (the merge scratch), so list ranks of all
//     candidates come from one lane-parallel binary search and broadcasts are single LDS.
//   * Footprint: 16-row stage, query half-row in registers, visited table of any size (multiply-high
//     slot mapping instead of a power-of-two mask) -> 9.4 KB and <= 80 registers per warp.
#include "beam_reglist.cuh"

namespace gbdr {

namespace {

struct V2Layout {
    uint32_t stage_off, q_off, nbr_off, scr_off, bar_off, vis_off, total;
};
__host__ __device__ inline V2Layout v2_layout(uint32_t C, uint32_t cap, uint32_t hcap) {
    V2Layout L;
    uint32_t o = 0;
    L.stage_off = o; o += 16u * (C * 16u + 16u);  // rows padded by 16 B (bank spread)
    L.q_off = o;     o += C * 16u;
    L.nbr_off = o;   o += 64u * 4u;
    L.scr_off = o;   o += cap * 8u;
    L.bar_off = o;   o += 16u;
    L.vis_off = o;   o += hcap * 4u;
    L.total = (o + 15u) & ~15u;
    return L;
}

// The shared visited table is an array of 4-slot buckets filled front to back; an id hashes to one
// bucket and overflows to the next one only when that bucket is full.  One LDS.128 tests a bucket.
__device__ __forceinline__ uint32_t bucket_of(uint32_t id, uint32_t nbuckets) {
    return __umulhi(id * 0x9E3779B1u, nbuckets);
}

// slow path once the shared table is closed to inserts: look the id up there, then test-and-set in
// the per-warp global overflow table.  true when `id` was not visited before.
__device__ __forceinline__ bool visit_spill(const uint32_t* vis, uint32_t nbuckets, uint32_t* spill,
                                            uint32_t spill_cap, uint32_t spill_shift, uint32_t id) {
    uint32_t g = bucket_of(id, nbuckets);
    for (;;) {
        const uint4 cur = reinterpret_cast<const uint4*>(vis)[g];
        if (cur.x == id || cur.y == id || cur.z == id || cur.w == id) return false;
        if (cur.w == PAD_ID) break;  // bucket not full: the id never overflowed past it
        g = g + 1 == nbuckets ? 0u : g + 1;
    }
    const uint32_t smask = spill_cap - 1;
    uint32_t slot = (id * 0x85EBCA6Bu) >> spill_shift;
    for (;;) {
        const uint32_t old = atomicCAS(&spill[slot], PAD_ID, id);
        if (old == PAD_ID) return true;
        if (old == id) return false;
        slot = (slot + 1) & smask;
    }
}

// ---- mbarrier + bulk-copy PTX ----
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

template <int C_T>
struct RowGeom {
    static constexpr uint32_t ROW_BYTES = C_T * 16u;
    static constexpr uint32_t PITCH = ROW_BYTES + 16u;  // bytes between staged rows
};

// rows ids[0..mb) (mb <= 16) -> stage: one bulk copy per row, issued by lane r; all lanes wait
template <int C_T>
__device__ __forceinline__ void gather16(uint32_t stage_s, uint32_t bar_s, uint32_t& parity, const uint32_t* ids,
                                         int mb, const float* db, uint32_t row_stride, int lane) {
    if (lane == 0) mbar_expect_tx(bar_s, (uint32_t)mb * RowGeom<C_T>::ROW_BYTES);
    if (lane < mb)
        bulk_g2s(stage_s + lane * RowGeom<C_T>::PITCH, db + (size_t)ids[lane] * row_stride, RowGeom<C_T>::ROW_BYTES,
                 bar_s);
    mbar_wait(bar_s, parity);
    parity ^= 1u;
}

// packed f32x2 arithmetic (sm_100): both elements individually rounded to nearest-even, and the
// explicit .rn keeps ptxas from contracting mul+add into an fma, so each element sees exactly the
// reference's sub / mul / add sequence.
__device__ __forceinline__ uint64_t f2_sub(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint64_t f2_mul(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
// canonical squared L2 of the staged rows against the query; the distance of row r is returned in
// lane 2r (odd lanes hold garbage).  qh[c] = packed (q[4c+2h], q[4c+2h+1]) with h = lane & 1.
template <int C_T>
__device__ __forceinline__ float dist16(const unsigned char* stage, const uint64_t (&qh)[C_T], int mb, int lane) {
    const int r = lane >> 1, h = lane & 1;
    float sa = 0.f, sb = 0.f;
    if (r < mb) {
        const uint64_t* row = reinterpret_cast<const uint64_t*>(stage + (size_t)r * RowGeom<C_T>::PITCH) + h;
#pragma unroll
        for (int c = 0; c < C_T; ++c) {
            // packed subtract and square, scalar accumulate: ptxas contracts mul.rn.f32x2 + add.rn.f32x2
            // into FFMA2 (one rounding) even with explicit .rn, which would break bit-exactness
            const uint64_t e = f2_sub(qh[c], row[c * 2]);
            const uint64_t sq = f2_mul(e, e);
            sa = __fadd_rn(sa, __uint_as_float((uint32_t)sq));
            sb = __fadd_rn(sb, __uint_as_float((uint32_t)(sq >> 32)));
        }
    }
    const float t2 = __shfl_down_sync(FULL_MASK, sa, 1), t3 = __shfl_down_sync(FULL_MASK, sb, 1);
    __syncwarp();  // the stage may be overwritten by the next gather
    return __fadd_rn(__fadd_rn(__fadd_rn(sa, sb), t2), t3);
}

// exact visited test-and-set for one adjacency chunk (ids distinct across lanes, PAD_ID = none) on
// the shared table, without atomics: claiming an empty slot is store / __syncwarp / read-back, the
// lane that reads its own id back owns the slot, a loser retries the bucket.  Warp-uniform; returns
// true when `id` was not visited before.
__device__ __forceinline__ bool visit_chunk(uint32_t* vis, uint32_t nbuckets, uint32_t id) {
    bool pending = id != PAD_ID, isnew = false;
    uint32_t g = bucket_of(id, nbuckets);
    while (__any_sync(FULL_MASK, pending)) {
        uint4 cur = make_uint4(0u, 0u, 0u, 0u);
        if (pending) cur = reinterpret_cast<const uint4*>(vis)[g];
        const bool found = (cur.x == id) | (cur.y == id) | (cur.z == id) | (cur.w == id);
        const uint32_t e = cur.x == PAD_ID ? 0u : cur.y == PAD_ID ? 1u : cur.z == PAD_ID ? 2u : cur.w == PAD_ID ? 3u : 4u;
        if (found) pending = false;  // already visited
        const bool claim = pending && e < 4u;
        if (claim) vis[g * 4u + e] = id;  // several lanes may race for one slot
        __syncwarp();
        if (claim) {
            if (vis[g * 4u + e] == id) {  // the id read back owns the slot
                isnew = true;
                pending = false;
            }                             // else: lost the race, the bucket has other free slots: retry it
        } else if (pending) {
            g = g + 1 == nbuckets ? 0u : g + 1;  // bucket full
        }
    }
    return isnew;
}

__device__ __forceinline__ void prefetch_l2(const void* ptr) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
}

// Merge the candidates flagged in `am` (one per lane: cdist, cid) into the sorted list.  Requires
// size <= ef and scr[0..size) to mirror the list.  Returns false, leaving list and mirror untouched,
// when an exact distance tie is involved (the caller then applies the sequential rules).
template <int R>
__device__ __forceinline__ bool merge_batch(RegList<R>& L, int& size, float& worst, const int ef, const unsigned am,
                                            const float cdist, const uint32_t cid, uint2* scr, const int lane) {
    const float INF = __int_as_float(0x7f800000);
    const bool mine = (am >> lane) & 1u;
    // rank among list entries: lower bound of cdist in scr[0..size) (lane-parallel binary search)
    int lo = 0;
    {
        int hi = size;
#pragma unroll 1
        for (int step = 32 * R; step > 0; step >>= 1) {  // CAP = 32R >= size: log2(CAP)+1 probes suffice
            const int mid = (lo + hi) >> 1;
            const bool go = lo < hi && __uint_as_float(scr[mid].x) < cdist;
            if (lo < hi) {
                if (go) lo = mid + 1; else hi = mid;
            }
        }
    }
    bool eq = mine && lo < size && __uint_as_float(scr[lo].x) == cdist;
    // rank among the other candidates, and the shift of every list entry
    int sh[R];
#pragma unroll
    for (int r = 0; r < R; ++r) sh[r] = 0;
    int cr = 0;
    unsigned m = am;
    while (m) {
        const int src = __ffs(m) - 1;
        m &= m - 1;
        const float x = __shfl_sync(FULL_MASK, cdist, src);
#pragma unroll
        for (int r = 0; r < R; ++r) sh[r] += (x < L.d[r]) ? 1 : 0;  // entries beyond `size` hold +inf
        cr += (x < cdist) ? 1 : 0;
        eq |= mine && lane != src && x == cdist;
    }
    if (__any_sync(FULL_MASK, eq)) return false;
    int nsize = size + __popc(am);
    // a tie across the ef boundary can only involve two old entries now (candidates are tie-free):
    // check it on the registers' view before anything is written
    __syncwarp();
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int e = lane * R + r;
        const int np = e + sh[r];
        if (e < size && np <= ef) scr[np] = make_uint2(__float_as_uint(L.d[r]), L.i[r]);
    }
    if (mine) {
        const int np = lo + cr;
        if (np <= ef) scr[np] = make_uint2(__float_as_uint(cdist), cid);
    }
    __syncwarp();
    bool tie = false;
    if (nsize > ef) {
        tie = scr[ef - 1].x == scr[ef].x;  // the sequential rules decide such a tie
        nsize = ef;
    }
    if (tie) {
        // undo: restore the mirror from the registers
        __syncwarp();
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int e = lane * R + r;
            if (e < size) scr[e] = make_uint2(__float_as_uint(L.d[r]), L.i[r]);
        }
        __syncwarp();
        return false;
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int e = lane * R + r;
        uint2 v = make_uint2(__float_as_uint(INF), PAD_ID);
        if (e < nsize) v = scr[e];
        L.d[r] = __uint_as_float(v.x);
        L.i[r] = v.y;
    }
    size = nsize;
    if (size >= ef) worst = __uint_as_float(scr[ef - 1].x);
    return true;
}

template <int R, int C_T>
__global__ void __launch_bounds__(256, (R <= 2 ? 3 : 2))
    beam_search_v2_kernel(const BeamParams p, uint32_t* __restrict__ counter) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    constexpr int CAP = 32 * R;
    const V2Layout Lo = v2_layout(C_T, CAP, p.hcap);
    unsigned char* wbase = smem_raw + (size_t)warp * p.smem_per_warp;
    unsigned char* stage = wbase + Lo.stage_off;
    float* qs = reinterpret_cast<float*>(wbase + Lo.q_off);
    uint32_t* nbr = reinterpret_cast<uint32_t*>(wbase + Lo.nbr_off);
    uint2* scr = reinterpret_cast<uint2*>(wbase + Lo.scr_off);
    uint32_t* vis = reinterpret_cast<uint32_t*>(wbase + Lo.vis_off);
    const uint32_t stage_s = (uint32_t)__cvta_generic_to_shared(stage);
    const uint32_t bar_s = (uint32_t)__cvta_generic_to_shared(wbase + Lo.bar_off);
    uint32_t parity = 0;
    const uint32_t gwarp = blockIdx.x * (blockDim.x >> 5) + warp;
    uint32_t* spill = p.spill + (size_t)gwarp * p.spill_cap;
    const int ef = (int)p.ef;
    const float INF = __int_as_float(0x7f800000);
    uint32_t status_acc = 0;
    if (lane == 0) mbar_init(bar_s, 1);
    __syncwarp();

    for (;;) {
        uint32_t qi = 0;
        if (lane == 0) qi = atomicAdd(counter, 1u);
        qi = __shfl_sync(FULL_MASK, qi, 0);
        if (qi >= p.n_q) break;

        // ---- per-query init ----
        {
            const uint4 fill = make_uint4(PAD_ID, PAD_ID, PAD_ID, PAD_ID);
            for (uint32_t i = lane; i < p.hcap / 4; i += 32) reinterpret_cast<uint4*>(vis)[i] = fill;
        }
        const float* qg = p.q + (size_t)qi * p.q_stride;
        if (lane < C_T) reinterpret_cast<float4*>(qs)[lane] = __ldg(reinterpret_cast<const float4*>(qg) + lane);
        __syncwarp();
        uint64_t qh[C_T];
#pragma unroll
        for (int c = 0; c < C_T; ++c) qh[c] = reinterpret_cast<const uint64_t*>(qs)[c * 2 + (lane & 1)];

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

        // ---- entry point (search_function.h:56-64) ----
        {
            const uint32_t e = __ldg(p.entry + qi);
            if (lane == 0) {
                nbr[0] = e;
                vis[bucket_of(e, p.hcap / 4u) * 4u] = e;
            }
            __syncwarp();
            gather16<C_T>(stage_s, bar_s, parity, nbr, 1, p.db, p.row_stride, lane);
            float d0 = dist16<C_T>(stage, qh, 1, lane);
            d0 = __shfl_sync(FULL_MASK, d0, 0);
            if (lane == 0) {
                L.d[0] = d0;
                L.i[0] = e;
                scr[0] = make_uint2(__float_as_uint(d0), e);
            }
            __syncwarp();
            size = 1;
            if (ef == 1) worst = d0;
            vcount = 1;
        }

        uint32_t pnode = PAD_ID, pa0 = PAD_ID, pa1 = PAD_ID;  // speculatively loaded adjacency row

        // ---- main loop (search_function.h:65-91) ----
        for (;;) {
            // best (and second best) un-expanded entries: the top of candidateSet and its successor
            int best = 0x7fffffff, second = 0x7fffffff;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const bool u = (lane * R + r < size) && !(L.i[r] & EXPANDED);
                const unsigned m = __ballot_sync(FULL_MASK, u);
                if (m) {
                    const int c1 = (__ffs(m) - 1) * R + r;
                    const unsigned m2 = m & (m - 1);
                    const int c2 = m2 ? (__ffs(m2) - 1) * R + r : 0x7fffffff;
                    if (c1 < best) {
                        second = min(best, c2);
                        best = c1;
                    } else {
                        second = min(second, c1);
                    }
                }
            }
            if (best == 0x7fffffff) break;  // candidateSet empty, or its best is worse than worst (:65,:67)
            int csel = best;
            if (best + 1 < size && scr[best + 1].x == scr[best].x) {
                // ties on dist: the reference pops the largest id first (max-heap of (-dist,id))
                const float dsel = list_get_d<R>(L, best);
                for (int j = best + 1; j < size; ++j) {
                    if (list_get_d<R>(L, j) != dsel) break;
                    if (!(list_get_i<R>(L, j) & EXPANDED)) csel = j;
                }
            }
            const uint32_t node = scr[csel].y & ID_MASK;
#pragma unroll
            for (int r = 0; r < R; ++r)
                if (lane * R + r == csel) L.i[r] |= EXPANDED;

            // adjacency row of `node`: from the speculative load when the guess was right
            const uint32_t* arow = p.adj + (size_t)node * p.adj_stride;
            uint32_t a0, a1;
            if (node == pnode) {
                a0 = pa0;
                a1 = pa1;
            } else {
                a0 = __ldg(arow + lane);
                a1 = (32 < p.adj_stride) ? __ldg(arow + 32 + lane) : PAD_ID;
            }
            // guess the next node: the runner-up of the current list (refined below once the new
            // candidates' distances are known)
            float pdist = INF;
            pnode = PAD_ID;
            if (second != 0x7fffffff && csel == best) {
                const uint2 sv = scr[second];
                pnode = sv.y & ID_MASK;
                pdist = __uint_as_float(sv.x);
                const uint32_t* prow = p.adj + (size_t)pnode * p.adj_stride;
                pa0 = __ldg(prow + lane);
                pa1 = (32 < p.adj_stride) ? __ldg(prow + 32 + lane) : PAD_ID;
            }

            // ---- makeStep over the adjacency row, 64 ids at a time (:23-39) ----
            for (uint32_t cb = 0; cb < p.adj_stride; cb += 64) {
                if (cb) {
                    a0 = __ldg(arow + cb + lane);
                    a1 = (cb + 32 < p.adj_stride) ? __ldg(arow + cb + 32 + lane) : PAD_ID;
                }
                const unsigned v0 = __ballot_sync(FULL_MASK, a0 != PAD_ID);
                const unsigned v1 = __ballot_sync(FULL_MASK, a1 != PAD_ID);
                scanned += __popc(v0) + __popc(v1);
                if ((v0 | v1) == 0) break;

                const bool smem_open = vcount + 64 <= p.hlimit;
                bool n0 = false, n1 = false;
                if (smem_open) {
                    n0 = visit_chunk(vis, p.hcap / 4u, a0);
                    if (v1) n1 = visit_chunk(vis, p.hcap / 4u, a1);
                } else {
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
                    if (a0 != PAD_ID) n0 = visit_spill(vis, p.hcap / 4u, spill, p.spill_cap, p.spill_shift, a0);
                    __syncwarp();
                    if (a1 != PAD_ID) n1 = visit_spill(vis, p.hcap / 4u, spill, p.spill_cap, p.spill_shift, a1);
                    __syncwarp();
                }
                const unsigned m0 = __ballot_sync(FULL_MASK, n0);
                const unsigned m1 = __ballot_sync(FULL_MASK, n1);
                const int c0 = __popc(m0), mtot = c0 + __popc(m1);
                if (smem_open) vcount += mtot; else scount += mtot;
                if (n0) nbr[__popc(m0 & lanemask_lt())] = a0;
                if (n1) nbr[c0 + __popc(m1 & lanemask_lt())] = a1;
                // whichever of these is expanded next, its adjacency row will be waiting in L2
                // (spends idle HBM bandwidth to take a DRAM round trip off the per-hop critical path)
                if (n0) prefetch_l2(p.adj + (size_t)a0 * p.adj_stride);
                if (n1) prefetch_l2(p.adj + (size_t)a1 * p.adj_stride);
                __syncwarp();
                dist_calc += mtot;  // :29

                for (int b0 = 0; b0 < mtot; b0 += 32) {
                    const int mb = min(32, mtot - b0);
                    // rows b0..b0+15 -> even lanes, rows b0+16..b0+31 -> odd lanes
                    gather16<C_T>(stage_s, bar_s, parity, nbr + b0, min(16, mb), p.db, p.row_stride, lane);
                    float cdist = dist16<C_T>(stage, qh, min(16, mb), lane);
                    if (mb > 16) {
                        gather16<C_T>(stage_s, bar_s, parity, nbr + b0 + 16, mb - 16, p.db, p.row_stride, lane);
                        const float d1 = dist16<C_T>(stage, qh, mb - 16, lane);
                        const float d1u = __shfl_up_sync(FULL_MASK, d1, 1);
                        if (lane & 1) cdist = d1u;
                    }
                    const int rr = (lane >> 1) + ((lane & 1) << 4);  // adjacency-order row of this lane
                    const bool have = rr < mb;
                    const uint32_t cid = have ? nbr[b0 + rr] : 0u;
                    // accept test against the worst at the start of the batch (worst never increases)
                    const bool pre = have && (size < ef || worst > cdist);
                    const unsigned am = __ballot_sync(FULL_MASK, pre);
                    if (!am) continue;

                    // refine the guess: a new candidate closer than the runner-up will be expanded next
                    {
                        const uint32_t key = pre ? __float_as_uint(cdist) : 0xffffffffu;
                        const uint32_t kmin = __reduce_min_sync(FULL_MASK, key);
                        if (pnode == PAD_ID || __uint_as_float(kmin) < pdist) {
                            const int who = __ffs(__ballot_sync(FULL_MASK, key == kmin)) - 1;
                            pnode = __shfl_sync(FULL_MASK, cid, who);
                            pdist = __uint_as_float(kmin);
                            const uint32_t* prow = p.adj + (size_t)pnode * p.adj_stride;
                            pa0 = __ldg(prow + lane);
                            pa1 = (32 < p.adj_stride) ? __ldg(prow + 32 + lane) : PAD_ID;
                        }
                    }

                    if (size <= ef && merge_batch<R>(L, size, worst, ef, am, cdist, cid, scr, lane)) continue;

                    // ---- exact-tie fallback: the reference's sequential accept/evict (:31-36) ----
                    for (int row = 0; row < mb; ++row) {
                        const int src = row < 16 ? 2 * row : 2 * (row - 16) + 1;
                        if (!((am >> src) & 1u)) continue;
                        const float x = __shfl_sync(FULL_MASK, cdist, src);
                        const uint32_t xid = __shfl_sync(FULL_MASK, cid, src);
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
                    // restore the invariants merge_batch relies on: +inf beyond size, mirror == list
                    __syncwarp();
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const int e = lane * R + r;
                        if (e >= size) {
                            L.d[r] = INF;
                            L.i[r] = PAD_ID;
                        }
                        scr[e] = make_uint2(__float_as_uint(L.d[r]), L.i[r]);
                    }
                    __syncwarp();
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
    GBDR_CUDA(cudaFuncSetAttribute(beam_search_v2_kernel<R, C_T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    beam_search_v2_kernel<R, C_T><<<blocks, wpb * 32, smem, st>>>(p, counter);
    GBDR_CHECK_LAUNCH();
    count_launch();
    return GBDR_OK;
}

template <int R>
int launch_r(const BeamParams& p, uint32_t wpb, uint32_t blocks, uint32_t* counter, cudaStream_t st) {
    switch (p.C) {
        case 4: return launch_rt<R, 4>(p, wpb, blocks, counter, st);
        case 8: return launch_rt<R, 8>(p, wpb, blocks, counter, st);
        case 12: return launch_rt<R, 12>(p, wpb, blocks, counter, st);
        case 16: return launch_rt<R, 16>(p, wpb, blocks, counter, st);
        default:
            set_error("beam_search_v2: unsupported row width");
            return GBDR_E_INVALID;
    }
}

}  // namespace

bool beam_v2_supports(uint32_t C) { return C == 4 || C == 8 || C == 12 || C == 16; }

uint32_t beam_v2_smem_per_warp(uint32_t C, uint32_t cap, uint32_t hcap) { return v2_layout(C, cap, hcap).total; }

int launch_beam_search_v2(const BeamParams& p, uint32_t wpb, uint32_t blocks, cudaStream_t st) {
    uint32_t* counter = p.status + 1;
    switch (p.cap) {
        case 32: return launch_r<1>(p, wpb, blocks, counter, st);
        case 64: return launch_r<2>(p, wpb, blocks, counter, st);
        case 128: return launch_r<4>(p, wpb, blocks, counter, st);
        case 256: return launch_r<8>(p, wpb, blocks, counter, st);
        default:
            set_error("beam_search_v2: unsupported list capacity");
            return GBDR_E_INVALID;
    }
}

}  // namespace gbdr
