// beam_search_v3.cu — K2, third generation: TWO queries per warp (one per half-warp of 16 lanes).
//
// Same results as beam_search_v2.cuh (reference search/search_function.h:15-102; ids, distances, hops and dist_calc
// bit-exact against the oracle), same per-query shared-memory footprint, same exact visited set — but every warp
// instruction now advances two independent walks.  The ncu captures of the v2 kernel (profiles/r2a_*, r3b_*) showed it
// issue-bound, not HBM-bound: ~615 warp instructions per hop at 72 % issue-slot utilisation, most of them bookkeeping
// whose cost does not depend on how many lanes take part (ballots, find-first-set, address arithmetic, loop control,
// one binary search per lane, one bulk copy per row).  A hop touches ~28 adjacency ids and ~15 new rows, so 16 lanes are
// enough to cover it in one or two passes, and the second half of the warp can run another query through the same
// instruction stream:
//   * every collective (ballot, shuffle, redux, syncwarp) is issued with the half's own member mask, so the two halves
//     are independent walks that merely happen to execute together; data-dependent branches (second pass over > 16 new
//     rows, exact-tie fallback, spill, query hand-over) diverge and reconverge, and one __syncwarp() per hop pulls the
//     halves back into lock step.  Nothing is shared between the halves but the instruction stream;
//   * list entry e lives in half-lane e & 15, register e >> 4; ballots over a register give 16-bit position masks that
//     are glued into 64-bit words, so "best / runner-up un-expanded entry" are two 64-bit find-first-set operations;
//   * one lane per gathered row: 16 rows per pass, each lane walks its whole row with LDS.128 (conflict-free at the
//     144-/80-byte row pitch) and keeps all four partial sums of L2Metric::Dist itself — no shuffles in the distance;
//   * adjacency rows are read 32 ids at a time, two consecutive ids per lane (one 8-byte load).
// Launch shape: CTAs of W warps = 2W query slots; shared memory per SLOT is the v2 layout (p.smem_per_warp is per slot).
#include "beam_search_v2.cuh"

namespace gbdr {

namespace {

constexpr int V3_NONE = 0x7fffffff;

__device__ __forceinline__ uint2 ldg_u2(const uint32_t* p) { return __ldg(reinterpret_cast<const uint2*>(p)); }

// Vis16 insertion with a shared-memory atomic on the bucket's fill count (vis16_visit_chunk_atomic of the v2 kernel),
// without the trailing warp barrier: the caller synchronises its half once per chunk.
__device__ __forceinline__ bool v3_visit16(uint32_t* vis, const VisCtx& c, uint32_t id, bool& exhausted) {
    bool isnew = false;
    if (id != PAD_ID) {
        uint32_t g, entry0;
        Vis16::locate(c, id, g, entry0);
        const uint32_t maxdisp = (1u << c.dbits) - 1u;
        for (uint32_t disp = 0;; ++disp) {
            const uint4 cur = reinterpret_cast<const uint4*>(vis)[g];
            const uint32_t entry = entry0 | disp;
            if (Vis16::found_in(cur, entry)) break;
            if ((cur.x & 0xFFFFu) < Vis16::SLOTS) {
                const uint32_t slot = atomicAdd(&vis[g * 4u], 1u) & 0xFFFFu;
                if (slot < Vis16::SLOTS) {
                    reinterpret_cast<uint16_t*>(vis)[g * 8u + 1u + slot] = (uint16_t)entry;
                    isnew = true;
                    break;
                }
            }
            if (disp == maxdisp) {
                exhausted = true;  // window full: this id lives in the global table
                break;
            }
            g = Vis16::next(c, g);
        }
    }
    return isnew;
}

// ---- collectives of a half, issued warp-wide ----
// A *_sync intrinsic whose member mask differs between the halves compiles to a loop over the distinct masks (the first
// build of this kernel spent 12 % of its instructions there and ran the halves one after the other).  So the kernel keeps
// its control flow warp-uniform around every collective — loops run while ANY half has work, halves without work are
// predicated off — and uses full-mask votes, width-16 shuffles and packed reductions.
__device__ __forceinline__ unsigned half_ballot(bool pred, int hs) { return (__ballot_sync(FULL_MASK, pred) >> hs) & 0xFFFFu; }
__device__ __forceinline__ uint32_t half_min(uint32_t key, int hs) {
    const uint32_t m0 = __reduce_min_sync(FULL_MASK, hs ? 0xffffffffu : key);
    const uint32_t m1 = __reduce_min_sync(FULL_MASK, hs ? key : 0xffffffffu);
    return hs ? m1 : m0;
}
__device__ __forceinline__ uint32_t half_or(uint32_t v, int hs) {
    const uint32_t m0 = __reduce_or_sync(FULL_MASK, hs ? 0u : v);
    const uint32_t m1 = __reduce_or_sync(FULL_MASK, hs ? v : 0u);
    return hs ? m1 : m0;
}

// rows ids[0..mb) (mb <= 16) -> the half's stage: one bulk copy per row, issued by half-lane r; the half waits
// (mb == 0: nothing to fetch for this half in this round)
template <int C_T>
__device__ __forceinline__ void v3_gather(uint32_t stage_s, uint32_t bar_s, uint32_t& parity, const uint32_t* ids, int mb,
                                          const float* db, uint32_t row_stride, int hl, uint32_t& status_acc) {
    // the tail of the stage doubles as the merge scratch (generic-proxy stores): order them before the async-proxy
    // writes of the copies below
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (mb > 0) {
        if (hl == 0) mbar_expect_tx(bar_s, (uint32_t)mb * RowGeom<C_T>::ROW_BYTES);
        if (hl < mb)
            bulk_g2s(stage_s + hl * RowGeom<C_T>::PITCH, db + (size_t)ids[hl] * row_stride, RowGeom<C_T>::ROW_BYTES, bar_s);
        if (!mbar_wait(bar_s, parity)) status_acc |= BEAM_ST_WATCHDOG | 0x400u;
        parity ^= 1u;
    }
}

// canonical squared L2 (search/support_func.h:107-128) of staged row `hl` against the query row kept in the stage pads:
// the four lane-strided sums of the reference in one thread, chunks in order, ((s0+s1)+s2)+s3
template <int C_T>
__device__ __forceinline__ float v3_dist(const unsigned char* stage, const unsigned char* qs, int mb, int hl) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    if (hl < mb) {
        const unsigned char* row = stage + (size_t)hl * RowGeom<C_T>::PITCH;
#pragma unroll
        for (int c = 0; c < C_T; ++c) {
            const ulonglong2 x = *reinterpret_cast<const ulonglong2*>(row + c * 16);
            const ulonglong2 q = *reinterpret_cast<const ulonglong2*>(q_chunk<C_T>(stage, qs, c));
            // packed subtract and square, scalar accumulate (ptxas would contract a packed add into FFMA2)
            const uint64_t e01 = f2_sub(q.x, x.x), e23 = f2_sub(q.y, x.y);
            const uint64_t p01 = f2_mul(e01, e01), p23 = f2_mul(e23, e23);
            s0 = __fadd_rn(s0, __uint_as_float((uint32_t)p01));
            s1 = __fadd_rn(s1, __uint_as_float((uint32_t)(p01 >> 32)));
            s2 = __fadd_rn(s2, __uint_as_float((uint32_t)p23));
            s3 = __fadd_rn(s3, __uint_as_float((uint32_t)(p23 >> 32)));
        }
    }
    return __fadd_rn(__fadd_rn(__fadd_rn(s0, s1), s2), s3);
}

// Merge the candidates flagged in `am` (16-bit, one per half-lane: cdist, cid) into the half's sorted list; `go_in` says
// whether this half takes part (the other half may be the only one with candidates).  Same contract as merge_batch of
// the v2 kernel: requires size <= ef and the mirror to hold (+inf, PAD) behind `size`; returns false with registers and
// mirror untouched when an exact distance tie is involved (or the half did not take part).  Called by the whole warp.
template <int R4>
__device__ __forceinline__ bool v3_merge(float (&Ld)[R4], uint32_t (&Li)[R4], int& size, float& worst, const int ef,
                                         const bool go_in, const unsigned am, const float cdist, const uint32_t cid, uint2* scr,
                                         float* candf, uint2* cs, const int hl, const int hs) {
    constexpr int CAP = 16 * R4;
    constexpr int NW32 = (CAP + 31) / 32;
    const float INF = __int_as_float(0x7f800000);
    const unsigned lt = (1u << hl) - 1u;
    const bool mine = go_in && ((am >> hl) & 1u);
    const int na = go_in ? __popc(am) : 0;
    if (mine) candf[__popc(am & lt)] = cdist;
    if (go_in && hl < 4) candf[na + hl] = INF;
    // rank among the list entries (branch-free lower bound on the mirror; +inf behind `size`, size < CAP)
    constexpr int P2 = CAP & (CAP - 1) ? (CAP >= 256 ? 256 : CAP >= 128 ? 128 : CAP >= 64 ? 64 : 32) : CAP;
    int lo = 0;
    if (P2 != CAP && __uint_as_float(scr[P2 - 1].x) < cdist) lo = CAP - P2;
#pragma unroll
    for (int step = P2 / 2; step > 0; step >>= 1)
        if (__uint_as_float(scr[lo + step - 1].x) < cdist) lo += step;
    const bool eq_list = mine && __uint_as_float(scr[lo].x) == cdist;
    __syncwarp();
    // rank among the other candidates
    int cr = 0;
    for (int j = 0; j < na; j += 4) {
        const float4 x = *reinterpret_cast<const float4*>(candf + j);
        cr += (x.x < cdist ? 1 : 0) + (x.y < cdist ? 1 : 0) + (x.z < cdist ? 1 : 0) + (x.w < cdist ? 1 : 0);
    }
    const int np = lo + cr;  // final position
    const bool keep = mine && np < CAP;
    unsigned P[NW32];
    int nbits = 0;
#pragma unroll
    for (int w = 0; w < NW32; ++w) {
        P[w] = half_or((keep && (np >> 5) == w) ? 1u << (np & 31) : 0u, hs);
        nbits += __popc(P[w]);
    }
    const int nkeep = __popc(half_ballot(keep, hs));
    // two candidates on one position = equal distances; equal to a list entry = same
    const bool bad = half_ballot(eq_list, hs) != 0u || nbits != nkeep;
    bool go = go_in && !bad;
    __syncwarp();  // candf (aliases cs) is dead from here
    if (go && keep) cs[cr] = make_uint2(__float_as_uint(cdist), cid);
    __syncwarp();
    // gather: slot j takes candidate #popc(P below j) or old entry j - popc(P below j)
    uint2 nv[R4];
    int below = 0;
#pragma unroll
    for (int r = 0; r < R4; ++r) {
        const unsigned pr = (P[r >> 1] >> (16 * (r & 1))) & 0xFFFFu;
        const int cnt = below + __popc(pr & lt);
        const bool is_c = (pr >> hl) & 1u;
        nv[r] = make_uint2(__float_as_uint(Ld[r]), Li[r]);
        if (go) nv[r] = is_c ? cs[cnt] : scr[r * 16 + hl - cnt];
        below += __popc(pr);
    }
    __syncwarp();
    if (go) {
#pragma unroll
        for (int r = 0; r < R4; ++r) scr[r * 16 + hl] = nv[r];
    }
    __syncwarp();
    int nsize = size + na;
    float nworst = worst;
    bool trunc = false;
    if (go) {
        if (nsize > ef) {
            const uint32_t wl = scr[ef - 1].x, wn = scr[ef].x;
            if (wl == wn) {
                go = false;  // a tie across the ef boundary: the sequential rules decide it.  Undo below.
            } else {
                nsize = ef;
                nworst = __uint_as_float(wl);
                trunc = true;
            }
        } else if (nsize == ef) {
            nworst = __uint_as_float(scr[ef - 1].x);
        }
    }
    __syncwarp();
    const bool undo = go_in && !bad && !go;
    if (undo) {
#pragma unroll
        for (int r = 0; r < R4; ++r) scr[r * 16 + hl] = make_uint2(__float_as_uint(Ld[r]), Li[r]);
    }
    if (trunc) {
#pragma unroll
        for (int r = 0; r < R4; ++r)
            if (r * 16 + hl >= ef) {
                nv[r] = make_uint2(__float_as_uint(INF), PAD_ID);
                scr[r * 16 + hl] = nv[r];
            }
    }
    __syncwarp();
    if (go) {
#pragma unroll
        for (int r = 0; r < R4; ++r) {
            Ld[r] = __uint_as_float(nv[r].x);
            Li[r] = nv[r].y;
        }
        size = nsize;
        worst = nworst;
    }
    return go;
}

// register budget of the pair kernel: 16 warps per SM (4 per scheduler) at <= 128 registers
constexpr int V3_MAX_THREADS = 512;

template <int R4, int C_T>
__global__ void __launch_bounds__(V3_MAX_THREADS, 1) beam_search_v3_kernel(const BeamParams p, uint32_t* __restrict__ counter) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int hl = lane & 15;              // lane within the half
    const int hs = lane & 16;              // position of the half's bits in a warp-wide ballot
    const unsigned lt = (1u << hl) - 1u;   // half-lanes below this one
    constexpr int CAP = 16 * R4;
    constexpr int NW = (CAP + 63) / 64;
    using V = Vis16;
    const V2Layout Lo = v2_layout(C_T, CAP, p.vis_bytes);
    const uint32_t slot_in_cta = (uint32_t)warp * 2u + (uint32_t)(hs >> 4);
    unsigned char* wbase = smem_raw + (size_t)slot_in_cta * p.smem_per_warp;
    unsigned char* stage = wbase + Lo.stage_off;
    unsigned char* qs = wbase + Lo.q_off;
    uint32_t* nbr = reinterpret_cast<uint32_t*>(wbase + Lo.nbr_off);
    uint2* scr = reinterpret_cast<uint2*>(wbase + Lo.scr_off);
    uint2* cs = reinterpret_cast<uint2*>(wbase + Lo.cs_off);
    uint32_t* vis = reinterpret_cast<uint32_t*>(wbase + Lo.vis_off);
    const uint32_t stage_s = (uint32_t)__cvta_generic_to_shared(stage);
    const uint32_t bar_s = (uint32_t)__cvta_generic_to_shared(wbase + Lo.bar_off);
    uint32_t parity = 0;
    const uint32_t gslot = blockIdx.x * (blockDim.x >> 4) + slot_in_cta;
    uint32_t* spill = p.spill + (size_t)gslot * p.spill_cap;
    const int ef = (int)p.ef;
    const float INF = __int_as_float(0x7f800000);
    VisCtx vc;
    vc.nbuckets = p.vis_bytes / 16u;
    vc.hshift = p.vis_hshift;
    vc.tshift = p.vis_tshift;
    vc.dbits = p.vis_dbits;
    vc.spill = spill;
    vc.spill_cap = p.spill_cap;
    vc.spill_shift = p.spill_shift;
    uint32_t status_acc = 0;
    if (hl == 0) mbar_init(bar_s, 1);
    __syncwarp();

    // per-half walk state (uniform inside a half)
    bool live = false, drained = false;
    uint32_t qi = 0;
    float Ld[R4];
    uint32_t Li[R4];
#pragma unroll
    for (int r = 0; r < R4; ++r) {
        Ld[r] = INF;
        Li[r] = PAD_ID;
    }
    int size = 0, hops = 0, dist_calc = 0, scanned = 0;
    float worst = INF;
    uint32_t vcount = 0, scount = 0;
    bool spill_ready = false, failed = false;
    uint32_t pnode = PAD_ID;
    uint2 pa = make_uint2(PAD_ID, PAD_ID);  // speculatively loaded adjacency row (first 32 ids)
    bool pf_due = false;

    for (;;) {
        __syncwarp();
        if (__all_sync(FULL_MASK, drained)) break;

        // ---- best (and second best) un-expanded entries of the half's list ----
        const bool walking = live && !failed;
        int best = V3_NONE, second = V3_NONE;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            uint64_t U = 0;
#pragma unroll
            for (int r = 4 * w; r < 4 * w + 4 && r < R4; ++r)
                U |= (uint64_t)half_ballot(walking && (int)Li[r] >= 0, hs) << (16 * (r & 3));
            if (U && second == V3_NONE) {
                const int c1 = w * 64 + __ffsll((long long)U) - 1;
                const uint64_t U2 = U & (U - 1);
                if (best == V3_NONE) {
                    best = c1;
                    if (U2) second = w * 64 + __ffsll((long long)U2) - 1;
                } else {
                    second = c1;
                }
            }
        }

        // ---- hand-over: a half whose walk is over (candidateSet empty :65,:67, or failed) emits and takes the next query ----
        const bool need_new = best == V3_NONE && !drained;
        if (__any_sync(FULL_MASK, need_new)) {
            if (need_new && live) {
                const int nres = min(min(size, ef), (int)p.k);
#pragma unroll
                for (int r = 0; r < R4; ++r) {  // (k <= ef < CAP)
                    const int e = r * 16 + hl;
                    if (e < (int)p.k) {
                        const bool ok = e < nres && !failed;
                        p.out_ids[(size_t)qi * p.k + e] = ok ? (Li[r] & ID_MASK) + p.id_offset : PAD_ID;
                        if (p.out_dists) p.out_dists[(size_t)qi * p.k + e] = ok ? Ld[r] : INF;
                    }
                }
                if (hl == 0) {
                    if (p.hops) p.hops[qi] = hops;
                    if (p.dist_calc) p.dist_calc[qi] = dist_calc + p.dist_calc_bias;
                    if (p.scanned) p.scanned[qi] = scanned;
                }
                live = false;
            }
            uint32_t nq = 0;
            if (need_new && hl == 0) nq = atomicAdd(counter, 1u);
            nq = __shfl_sync(FULL_MASK, nq, 0, 16);
            const bool start = need_new && nq < p.n_q;
            if (need_new && !start) drained = true;
            uint32_t e = 0;
            if (start) {
                qi = nq;
                // ---- per-query init ----
                const uint4 fill = make_uint4(0xFFFF0000u, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
                for (uint32_t i = hl; i < vc.nbuckets; i += 16) reinterpret_cast<uint4*>(vis)[i] = fill;
                const float* qg = p.q + (size_t)qi * p.q_stride;
                if (hl < C_T)
                    *reinterpret_cast<float4*>(const_cast<unsigned char*>(q_chunk<C_T>(stage, qs, hl))) =
                        __ldg(reinterpret_cast<const float4*>(qg) + hl);
#pragma unroll
                for (int r = 0; r < R4; ++r) {
                    Ld[r] = INF;
                    Li[r] = PAD_ID;
                    scr[r * 16 + hl] = make_uint2(__float_as_uint(INF), PAD_ID);
                }
                size = 0;
                worst = INF;
                hops = 0;
                dist_calc = 1;  // search_function.h:52
                scanned = 0;
                vcount = 0;
                scount = 0;
                spill_ready = false;
                failed = false;
                pnode = PAD_ID;
                pf_due = false;
                // ---- entry point (search_function.h:56-64) ----
                e = __ldg(p.entry + qi);
                if (e >= p.n_vertices) {  // not a vertex: the query fails (PAD results) instead of reading out of bounds
                    e = 0;
                    failed = true;
                    status_acc |= BEAM_ST_BAD_ENTRY;
                }
                if (hl == 0) nbr[0] = e;
            }
            __syncwarp();
            if (start && hl == 0) V::insert_first(vis, vc, e);
            __syncwarp();
            v3_gather<C_T>(stage_s, bar_s, parity, nbr, start ? 1 : 0, p.db, p.row_stride, hl, status_acc);
            float d0 = v3_dist<C_T>(stage, qs, start ? 1 : 0, hl);
            d0 = __shfl_sync(FULL_MASK, d0, 0, 16);
            __syncwarp();
            if (start) {
                if (hl == 0) {
                    Ld[0] = d0;
                    Li[0] = e;
                    scr[0] = make_uint2(__float_as_uint(d0), e);
                }
                size = 1;
                if (ef == 1) worst = d0;
                vcount = 1;
                live = true;
                best = failed ? V3_NONE : 0;  // the entry is the list's only (un-expanded) entry
                second = V3_NONE;
            }
            __syncwarp();
        }

        // ================= one hop of this half's walk (search_function.h:65-91) =================
        const bool hop = live && !failed && best != V3_NONE;
        int csel = best;
        uint32_t node = 0;
        float pdist = INF;
        uint32_t pguess = PAD_ID;
        if (hop) {
            if (best + 1 < size && scr[best + 1].x == scr[best].x) {
                // ties on dist: the reference pops the largest id first (max-heap of (-dist,id))
                const uint32_t dsel = scr[best].x;
                for (int j = best + 1; j < size; ++j) {
                    const uint2 v = scr[j];
                    if (v.x != dsel) break;
                    if (!(v.y & EXPANDED)) csel = j;
                }
            }
            node = scr[csel].y & ID_MASK;
            // guess the next node: the runner-up of the current list (refined below once the new candidates' distances are known)
            if (second != V3_NONE && csel == best) {
                const uint2 sv = scr[second];
                pguess = sv.y & ID_MASK;
                pdist = __uint_as_float(sv.x);
            }
        }
        __syncwarp();
        if (hop && hl == (csel & 15)) {
#pragma unroll
            for (int r = 0; r < R4; ++r)
                if (r == (csel >> 4)) {
                    Li[r] |= EXPANDED;
                    scr[csel].y = Li[r];
                }
        }

        // adjacency row of `node`, 32 ids at a time, ids 2*hl and 2*hl + 1 in this lane; the first chunk comes from the
        // speculative load when the guess was right
        const uint32_t* arow = p.adj + (size_t)node * p.adj_stride;
        uint2 a = make_uint2(PAD_ID, PAD_ID);
        if (hop) {
            if (node == pnode) a = pa;
            else a = ldg_u2(arow + 2 * hl);
            pnode = pguess;
            pf_due = false;
            if (pnode != PAD_ID) {
                pa = ldg_u2(p.adj + (size_t)pnode * p.adj_stride + 2 * hl);
                pf_due = (p.pf_rows & 1u) != 0u;
            }
        }

        // ---- makeStep over the adjacency row (:23-39) ----
        bool row_open = hop;  // this half still has chunks of its row to scan
        for (uint32_t cb = 0;; cb += 32) {
            bool act = row_open && cb < p.adj_stride;
            if (!__any_sync(FULL_MASK, act)) break;
            if (cb) a = act ? ldg_u2(arow + cb + 2 * hl) : make_uint2(PAD_ID, PAD_ID);
            const unsigned v0 = half_ballot(act && a.x != PAD_ID, hs);
            const unsigned v1 = half_ballot(act && a.y != PAD_ID, hs);
            scanned += __popc(v0) + __popc(v1);
            if ((v0 | v1) == 0) act = false;  // the row ended at the chunk boundary (or this half sits the round out)
            if (v1 != 0xFFFFu) row_open = false;  // the row ends inside this chunk

            const bool smem_open = vcount + 32 <= p.hlimit;
            bool n0 = false, n1 = false, x0 = false, x1 = false;
            if (act && smem_open) {
                n0 = v3_visit16(vis, vc, a.x, x0);
                n1 = v3_visit16(vis, vc, a.y, x1);
            }
            __syncwarp();
            // ids the shared table cannot take (table closed, or their probe window is full) are tracked exactly in the
            // per-slot global table
            const bool to_spill = act && (smem_open ? (x0 | x1) : true);
            const unsigned sb = __ballot_sync(FULL_MASK, to_spill);
            uint32_t snew = 0;
            if (sb) {  // (rare; warp-uniform)
                const bool mine = ((sb >> hs) & 0xFFFFu) != 0u;
                if (mine && !spill_ready) {
                    for (uint32_t i = hl; i < p.spill_cap; i += 16) spill[i] = PAD_ID;
                    spill_ready = true;
                    status_acc |= BEAM_ST_SPILLED;
                }
                __syncwarp();
                if (mine && scount + 32 > (p.spill_cap >> 1) + (p.spill_cap >> 2)) {
                    failed = true;
                    status_acc |= BEAM_ST_VISITED_FULL;
                }
                const bool go = mine && !failed;
                if (go) {
                    if (smem_open) {
                        if (x0) n0 = spill_test_and_set(vc, a.x);
                    } else if (a.x != PAD_ID) {
                        n0 = visit_spill<V>(vis, vc, a.x);
                    }
                }
                __syncwarp();
                if (go) {
                    if (smem_open) {
                        if (x1) n1 = spill_test_and_set(vc, a.y);
                    } else if (a.y != PAD_ID) {
                        n1 = visit_spill<V>(vis, vc, a.y);
                    }
                }
                __syncwarp();
                // (votes are warp-wide: never under a condition that differs between the halves)
                snew = __popc(half_ballot(go && smem_open && x0 && n0, hs)) + __popc(half_ballot(go && smem_open && x1 && n1, hs));
                if (failed) {
                    n0 = n1 = false;
                    act = false;
                    row_open = false;
                }
            }
            const unsigned m0 = half_ballot(n0, hs);
            const unsigned m1 = half_ballot(n1, hs);
            const int mtot = __popc(m0) + __popc(m1);
            if (act) {
                if (smem_open) {
                    vcount += mtot - snew;
                    scount += snew;
                } else {
                    scount += mtot;
                }
            }
            // compact the new ids in adjacency order (id 2*hl before id 2*hl + 1)
            {
                const int pos0 = __popc(m0 & lt) + __popc(m1 & lt);
                if (n0) nbr[pos0] = a.x;
                if (n1) nbr[pos0 + (n0 ? 1 : 0)] = a.y;
            }
            if (!(p.pf_rows & 2u)) {
                if (n0) prefetch_l2(p.adj + (size_t)a.x * p.adj_stride);
                if (n1) prefetch_l2(p.adj + (size_t)a.y * p.adj_stride);
            }
            __syncwarp();
            dist_calc += mtot;  // :29
            // the guessed next node's adjacency row (requested at the top of the hop) has arrived by now: pull the vectors
            // it names into L2, so that the next hop's gather is an L2 hit
            if (act && pf_due) {
                prefetch_rows<C_T>(p.db, p.row_stride, pa.x, pa.y);
                pf_due = false;
            }

            for (int b0 = 0;; b0 += 16) {
                const bool bact = b0 < mtot && !failed;
                if (!__any_sync(FULL_MASK, bact)) break;
                const int mb = bact ? min(16, mtot - b0) : 0;
                v3_gather<C_T>(stage_s, bar_s, parity, nbr + b0, mb, p.db, p.row_stride, hl, status_acc);
                const float cdist = v3_dist<C_T>(stage, qs, mb, hl);
                __syncwarp();  // the stage may be overwritten by the merge scratch / the next gather
                const bool have = hl < mb;
                const uint32_t cid = have ? nbr[b0 + hl] : 0u;
                // accept test against the worst at the start of the batch (worst never increases)
                const bool pre = have && (size < ef || worst > cdist);
                const unsigned amw = __ballot_sync(FULL_MASK, pre);
                if (!amw) continue;
                const unsigned am = (amw >> hs) & 0xFFFFu;
                // only an accepted candidate can ever be expanded: its adjacency row will be waiting in L2
                if ((p.pf_rows & 2u) && pre) prefetch_l2(p.adj + (size_t)cid * p.adj_stride);

                // refine the guess: a new candidate closer than the runner-up will be expanded next
                {
                    const uint32_t key = pre ? __float_as_uint(cdist) : 0xffffffffu;
                    const uint32_t kmin = half_min(key, hs);
                    const int who = __ffs(half_ballot(key == kmin, hs)) - 1;  // (some lane of the half always matches)
                    const uint32_t cand = __shfl_sync(FULL_MASK, cid, who, 16);
                    if (am && (pnode == PAD_ID || __uint_as_float(kmin) < pdist)) {
                        pnode = cand;
                        pdist = __uint_as_float(kmin);
                        pa = ldg_u2(p.adj + (size_t)pnode * p.adj_stride + 2 * hl);
                        pf_due = (p.pf_rows & 1u) != 0u;
                    }
                }

                const bool merged = v3_merge<R4>(Ld, Li, size, worst, ef, am != 0u && size <= ef, am, cdist, cid, scr,
                                                 reinterpret_cast<float*>(cs), cs, hl, hs);
                const bool fb = am != 0u && !merged;
                if (!__any_sync(FULL_MASK, fb)) continue;

                // ---- exact-tie fallback: the reference's sequential accept/evict (:31-36), candidates in adjacency order.
                //      Rare; the whole warp walks the 16 rows, the half that needs it does the work ----
                for (int row = 0; row < 16; ++row) {
                    const float x = __shfl_sync(FULL_MASK, cdist, row, 16);
                    const uint32_t xid = __shfl_sync(FULL_MASK, cid, row, 16);
                    // :31 (and nothing more once the walk has failed)
                    const bool ins = fb && !failed && ((am >> row) & 1u) && !(size >= ef && !(worst > x));
                    // :32-34 sorted insert by (dist,id): shift the tail through the mirror
                    int pos = 0;
#pragma unroll
                    for (int r = 0; r < R4; ++r) {
                        const bool less = (r * 16 + hl < size) && pair_less(Ld[r], Li[r] & ID_MASK, x, xid);
                        pos += __popc(half_ballot(less, hs));
                    }
                    __syncwarp();
                    if (ins) {
#pragma unroll
                        for (int r = 0; r < R4; ++r) {
                            const int e = r * 16 + hl;
                            if (e >= pos && e < size && e + 1 < CAP) scr[e + 1] = make_uint2(__float_as_uint(Ld[r]), Li[r]);
                        }
                        if (hl == 0 && pos < CAP) scr[pos] = make_uint2(__float_as_uint(x), xid);
                    }
                    __syncwarp();
                    if (ins) {
                        size = size < CAP ? size + 1 : CAP;
#pragma unroll
                        for (int r = 0; r < R4; ++r) {
                            const uint2 v = scr[r * 16 + hl];
                            Ld[r] = __uint_as_float(v.x);
                            Li[r] = v.y;
                        }
                        if (size >= ef) worst = __uint_as_float(scr[ef - 1].x);
                    }
                    // :35-36 eviction; boundary ties (dist == new worst) stay in the slack
                    int keepn = 0;
#pragma unroll
                    for (int r = 0; r < R4; ++r) {
                        const int e = r * 16 + hl;
                        keepn += __popc(half_ballot(e >= ef && e < size && Ld[r] == worst, hs));
                    }
                    if (ins && size > ef) {
                        size = ef + keepn;
                        if (size >= CAP) {
                            failed = true;
                            status_acc |= BEAM_ST_TIE_OVERFLOW;
                        }
                    }
                }
                // restore the invariants v3_merge relies on: (+inf, PAD) behind size, mirror == list
                __syncwarp();
                if (fb) {
#pragma unroll
                    for (int r = 0; r < R4; ++r) {
                        const int e = r * 16 + hl;
                        if (e >= size) {
                            Ld[r] = INF;
                            Li[r] = PAD_ID;
                        }
                        scr[e] = make_uint2(__float_as_uint(Ld[r]), Li[r]);
                    }
                }
                __syncwarp();
            }
            if (failed) row_open = false;
        }
        if (hop && !failed) {
            if (pf_due) {  // guess refined during this hop: its adjacency row was requested before the merge
                prefetch_rows<C_T>(p.db, p.row_stride, pa.x, pa.y);
                pf_due = false;
            }
            ++hops;  // :90
            if (hops > dist_calc || (status_acc & BEAM_ST_WATCHDOG)) {  // every hop expands a distinct evaluated vertex
                status_acc |= BEAM_ST_WATCHDOG | 0x200u;
                failed = true;
            }
        }
    }
    if (status_acc && hl == 0) atomicOr(p.status, status_acc);
}

template <int R4, int C_T>
int launch_v3_rt(const BeamParams& p, uint32_t wpb, uint32_t blocks, uint32_t* counter, cudaStream_t st) {
    if (wpb * 32u > (uint32_t)V3_MAX_THREADS) {
        set_error("beam_search_v3: too many warps per CTA");
        return GBDR_E_INVALID;
    }
    const size_t smem = (size_t)p.smem_per_warp * wpb * 2u;
    GBDR_CUDA(cudaFuncSetAttribute(beam_search_v3_kernel<R4, C_T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    {  // the grid is persistent: never launch more CTAs than are resident
        static thread_local uint32_t seen_wpb = 0, seen_smem = 0, seen_dev = ~0u, seen_blocks = 0;
        int dev = 0;
        GBDR_CUDA(cudaGetDevice(&dev));
        if (seen_wpb != wpb || seen_smem != (uint32_t)smem || seen_dev != (uint32_t)dev) {
            int per_sm = 0, sms = 0;
            GBDR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, beam_search_v3_kernel<R4, C_T>, (int)(wpb * 32), smem));
            GBDR_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
            seen_wpb = wpb;
            seen_smem = (uint32_t)smem;
            seen_dev = (uint32_t)dev;
            seen_blocks = (uint32_t)std::max(1, per_sm * sms);
        }
        blocks = std::min(blocks, seen_blocks);
    }
    beam_search_v3_kernel<R4, C_T><<<blocks, wpb * 32, smem, st>>>(p, counter);
    GBDR_CHECK_LAUNCH();
    count_launch();
    return GBDR_OK;
}

template <int R4>
int launch_v3_r(const BeamParams& p, uint32_t wpb, uint32_t blocks, cudaStream_t st) {
    uint32_t* counter = p.status + 1;
    switch (p.C) {
        case 4: return launch_v3_rt<R4, 4>(p, wpb, blocks, counter, st);
        case 8: return launch_v3_rt<R4, 8>(p, wpb, blocks, counter, st);
        default:
            set_error("beam_search_v3: unsupported row width");
            return GBDR_E_INVALID;
    }
}

}  // namespace

bool beam_v3_supports(uint32_t C, uint32_t cap) { return (C == 4 || C == 8) && (cap == 32 || cap == 64 || cap == 96 || cap == 128); }

// `wpb` warps per CTA = 2 * wpb query slots; p.smem_per_warp is the shared memory of ONE slot
int launch_beam_search_v3(const BeamParams& p, uint32_t wpb, uint32_t blocks, cudaStream_t st) {
    if (!p.vis_tshift) {
        set_error("beam_search_v3: needs the 16-bit-tag visited table");
        return GBDR_E_INVALID;
    }
    switch (p.cap) {
        case 32: return launch_v3_r<2>(p, wpb, blocks, st);
        case 64: return launch_v3_r<4>(p, wpb, blocks, st);
        case 96: return launch_v3_r<6>(p, wpb, blocks, st);
        case 128: return launch_v3_r<8>(p, wpb, blocks, st);
        default:
            set_error("beam_search_v3: unsupported list capacity");
            return GBDR_E_INVALID;
    }
}

}  // namespace gbdr
