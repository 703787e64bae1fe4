// beam_search.cu — K2: greedy best-first beam search over a flat proximity graph, one warp per query.
//
// Reproduces getOneSearchResults + makeStep (reference search/search_function.h:15-102) for a single
// entry point, with or without the second ("long link") graph of :73-89, including its tie rules:
//   * topResults  = max-heap of (dist,id)  -> here: one list sorted ascending by (dist,id); the
//                   logical heap is the first `ef` entries, eviction removes the last of them.
//   * candidateSet = max-heap of (-dist,id) -> here: the un-expanded entries of the same list.  An
//                   accepted element that is later evicted has dist >= worst; with dist > worst the
//                   reference's loop breaks on it (:67) so it is dropped, with dist == worst it is
//                   still expanded by the reference, so such boundary ties are kept in the slack
//                   slots behind position ef-1.
//   * makeStep accepts iff `worst > dist || size < ef` (:31) evaluated sequentially in adjacency
//     order; the same order is used here (pre-filtered with the worst at the start of the hop, which
//     can only reject elements the sequential rule rejects too, because worst never increases).
//   * visited_list_pool.h's exact epoch-stamped array is replaced by an exact open-addressing hash in
//     shared memory that spills to a per-warp global table instead of ever dropping an id.
//
// Memory behaviour: per hop one 2x128 B adjacency read, then one 16 B x C coalesced row gather per
// unvisited neighbour (cp.async straight into a swizzled shared-memory tile, 8 lanes per 128 B row),
// then each lane owns one row and walks its chunks in order so the distance has the reference's
// summation order (common.cuh L2Acc) without any shuffle reduction.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "beam_search.cuh"

namespace gbdr {

namespace {

// insert (x,xid) into the ascending list; returns nothing, updates size / first_unexp
__device__ __forceinline__ void list_insert(float* rd, uint32_t* rid, int& size, const int cap, float x,
                                            uint32_t xid, int lane, int& first_unexp) {
    int pos = 0;
    for (int base = 0; base < size; base += 32) {
        int i = base + lane;
        bool less = false;
        if (i < size) less = pair_less(rd[i], rid[i] & ID_MASK, x, xid);
        pos += __popc(__ballot_sync(FULL_MASK, less));
    }
    const int last = size < cap ? size : cap - 1;  // entries [pos,last) move up by one
    if (last > pos) {
        for (int base = (last - 1) & ~31; base >= (pos & ~31); base -= 32) {
            int i = base + lane;
            bool mv = (i >= pos && i < last);
            float dv = 0.f;
            uint32_t iv = 0;
            if (mv) {
                dv = rd[i];
                iv = rid[i];
            }
            __syncwarp();
            if (mv) {
                rd[i + 1] = dv;
                rid[i + 1] = iv;
            }
            __syncwarp();
        }
    }
    if (lane == 0) {
        rd[pos] = x;
        rid[pos] = xid;
    }
    __syncwarp();
    size = size < cap ? size + 1 : cap;
    if (pos < first_unexp) first_unexp = pos;
}

template <int C_T>
__global__ void __launch_bounds__(256) beam_search_kernel(const BeamParams p, uint32_t* __restrict__ counter) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const uint32_t C = C_T ? (uint32_t)C_T : p.C;
    const BeamLayout L = beam_layout(C, p.cap, p.hcap);
    unsigned char* wbase = smem_raw + (size_t)warp * p.smem_per_warp;
    float* stage = reinterpret_cast<float*>(wbase + L.stage_off);
    float* qs = reinterpret_cast<float*>(wbase + L.q_off);
    float* rd = reinterpret_cast<float*>(wbase + L.rd_off);
    uint32_t* rid = reinterpret_cast<uint32_t*>(wbase + L.rid_off);
    uint32_t* nbr = reinterpret_cast<uint32_t*>(wbase + L.nbr_off);
    uint32_t* vis = reinterpret_cast<uint32_t*>(wbase + L.vis_off);
    const uint32_t gwarp = blockIdx.x * (blockDim.x >> 5) + warp;
    uint32_t* spill = p.spill + (size_t)gwarp * p.spill_cap;
    const int ef = (int)p.ef;
    const int cap = (int)p.cap;
    uint32_t status_acc = 0;

    for (;;) {
        uint32_t qi = 0;
        if (lane == 0) qi = atomicAdd(counter, 1u);
        qi = __shfl_sync(FULL_MASK, qi, 0);
        if (qi >= p.n_q) break;

        // ---- per-query init ----
        for (uint32_t i = lane; i < p.hcap; i += 32) vis[i] = PAD_ID;
        const float* qg = p.q + (size_t)qi * p.q_stride;
        for (uint32_t c = lane; c < C; c += 32)
            reinterpret_cast<float4*>(qs)[c] = __ldg(reinterpret_cast<const float4*>(qg) + c);
        __syncwarp();
        float4 qreg[C_T ? C_T : 1];
        if (C_T) {
#pragma unroll
            for (int c = 0; c < (C_T ? C_T : 1); ++c) qreg[c] = reinterpret_cast<const float4*>(qs)[c];
        }

        int size = 0, first_unexp = 0;
        int hops = 0, dist_calc = 1, scanned = 0;  // dist_calc starts at 1 (search_function.h:52)
        uint32_t vcount = 0, scount = 0;
        bool spill_ready = false, failed = false;

        // distances of nbr[b0 .. b0+mb) -> lane r holds the distance of row r
        auto batch_dist = [&](int b0, int mb) -> float {
            const uint32_t T = (uint32_t)mb * C;
            for (uint32_t t = lane; t < T; t += 32) {
                uint32_t r = C_T ? t / (uint32_t)(C_T ? C_T : 1) : t / C;
                uint32_t c = t - r * C;
                const float* src = p.db + (size_t)nbr[b0 + r] * p.row_stride + c * 4u;
                cp_async16(stage + ((size_t)r * C + swz<C_T>(r, c, C)) * 4u, src);
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
            list_insert(rd, rid, size, cap, d0, e, lane, first_unexp);
            vcount = 1;
        }

        // ---- main loop (search_function.h:65-91) ----
        while (!failed) {
            // best un-expanded entry = top of candidateSet
            int pfirst = -1;
            for (int base = first_unexp & ~31; base < size; base += 32) {
                int i = base + lane;
                bool u = (i >= first_unexp && i < size && !(rid[i] & EXPANDED));
                unsigned m = __ballot_sync(FULL_MASK, u);
                if (m) {
                    pfirst = base + __ffs(m) - 1;
                    break;
                }
            }
            if (pfirst < 0) break;  // candidateSet empty, or its best is worse than worst (:65,:67)
            int csel = pfirst;
            {
                // ties on dist: the reference pops the largest id first (max-heap of (-dist,id))
                float dsel = rd[pfirst];
                for (int j = pfirst + 1; j < size && rd[j] == dsel; ++j)
                    if (!(rid[j] & EXPANDED)) csel = j;
            }
            const uint32_t node = rid[csel] & ID_MASK;
            __syncwarp();
            if (lane == 0) rid[csel] |= EXPANDED;
            __syncwarp();
            first_unexp = (csel == pfirst) ? pfirst + 1 : pfirst;

            // ---- makeStep over one adjacency row, 64 ids at a time (:23-39); true when a neighbour
            // was accepted (`found`, :33) ----
            auto make_step = [&](const uint32_t* arow, const uint32_t stride) -> bool {
            bool found = false;
            for (uint32_t cb = 0; cb < stride; cb += 64) {
                uint32_t a0 = __ldg(arow + cb + lane);
                uint32_t a1 = (cb + 32 < stride) ? __ldg(arow + cb + 32 + lane) : PAD_ID;
                const unsigned v0 = __ballot_sync(FULL_MASK, a0 != PAD_ID);
                const unsigned v1 = __ballot_sync(FULL_MASK, a1 != PAD_ID);
                scanned += __popc(v0) + __popc(v1);
                if ((v0 | v1) == 0) break;

                // visited test-and-set
                bool smem_open = vcount + 64 <= p.hlimit;
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
                if (a1 != PAD_ID) n1 = visit(vis, p.hcap, p.hshift, smem_open, spill, p.spill_cap, p.spill_shift, a1);
                __syncwarp();
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
                    // pre-filter with the worst at the start of the batch
                    const bool full0 = size >= ef;
                    const float worst0 = full0 ? rd[ef - 1] : 0.f;
                    unsigned am = __ballot_sync(FULL_MASK, lane < mb && (!full0 || worst0 > dist));
                    while (am) {
                        const int src = __ffs(am) - 1;
                        am &= am - 1;
                        const float x = __shfl_sync(FULL_MASK, dist, src);
                        const uint32_t xid = __shfl_sync(FULL_MASK, myid, src);
                        if (size >= ef && !(rd[ef - 1] > x)) continue;  // :31
                        list_insert(rd, rid, size, cap, x, xid, lane, first_unexp);  // :32-34
                        found = true;
                        if (size > ef) {
                            // :35-36 eviction; keep boundary ties (dist == new worst) in the slack
                            const float w = rd[ef - 1];
                            int keep = 0;
                            for (int base = ef; base < size; base += 32) {
                                int i = base + lane;
                                keep += __popc(__ballot_sync(FULL_MASK, i < size && rd[i] == w));
                            }
                            size = ef + keep;
                            if (size >= cap) {
                                failed = true;
                                status_acc |= BEAM_ST_TIE_OVERFLOW;
                            }
                        }
                    }
                    __syncwarp();
                }
                if (failed) break;
                if (v1 != FULL_MASK) break;  // row ended inside this chunk
            }
            return found;
            };
            bool aux_found = false;
            if (p.aux_adj && (uint32_t)hops < p.hops_bound)  // :73
                aux_found = make_step(p.aux_adj + (size_t)node * p.aux_stride, p.aux_stride);
            if (!failed && !(aux_found && p.llf))  // :82 (always taken without a second graph)
                make_step(p.adj + (size_t)node * p.adj_stride, p.adj_stride);
            if (failed) break;
            ++hops;  // :90
        }

        // ---- emit the k best (:96-100) ----
        const int nres = min(min(size, ef), (int)p.k);
        for (int j = lane; j < (int)p.k; j += 32) {
            const bool ok = j < nres && !failed;
            p.out_ids[(size_t)qi * p.k + j] = ok ? (rid[j] & ID_MASK) + p.id_offset : PAD_ID;
            if (p.out_dists) p.out_dists[(size_t)qi * p.k + j] = ok ? rd[j] : __int_as_float(0x7f800000);
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

}  // namespace

static uint32_t env_u32(const char* name, uint32_t dflt) {
    const char* s = getenv(name);
    return s && *s ? (uint32_t)strtoul(s, nullptr, 10) : dflt;
}

void beam_plan(uint32_t ef, uint32_t C, uint64_t n, BeamPlan* plan, bool second_graph) {
    // list capacity: ef + >= 8 slack slots for boundary ties; the register kernels hold 32 ... 512 slots (the v2 kernel in
    // steps of 32 up to 192, then 256, 320, 384, 512: V2_CAPS; the sequential register kernel powers of two up to 256)
    uint32_t cp = (ef + 8 + 31) & ~31u;
    int variant = BEAM_SMEM_LIST;
    if (beam_v2_supports(C)) {
        for (uint32_t c : V2_CAPS)
            if (ef + 8 <= c) {
                cp = c;
                variant = BEAM_V2;
                break;
            }
    } else {
        for (uint32_t c = 32; c <= 256; c <<= 1)
            if (ef + 8 <= c) {
                cp = c;
                variant = BEAM_REG_LIST;
                break;
            }
    }
    const char* force = getenv("GBDR_BEAM_VARIANT");
    if (second_graph || (force && !strcmp(force, "smem"))) {
        variant = BEAM_SMEM_LIST;
        cp = (ef + 8 + 31) & ~31u;
    } else if (force && !strcmp(force, "reg") && variant == BEAM_V2) {
        uint32_t c2 = 32;
        while (c2 < ef + 8) c2 <<= 1;
        if (c2 <= 256) {
            variant = BEAM_REG_LIST;
            cp = c2;
        } else {  // the sequential register kernel stops at 256 slots
            variant = BEAM_SMEM_LIST;
            cp = (ef + 8 + 31) & ~31u;
        }
    }
    // expected visited ~ 12*ef + 200; keep the shared table below ~60 % at the mean
    const uint32_t want = (uint32_t)((12.0 * ef + 200.0) / 0.6);
    const uint32_t force_h = env_u32("GBDR_BEAM_HCAP", 0);
    const uint32_t force_w = env_u32("GBDR_BEAM_WPB", 0);
    plan->variant = variant;
    plan->cap = cp;
    plan->vis_bytes = plan->vis_hshift = plan->vis_tshift = plan->vis_dbits = 0;
    plan->dense = 0;
    if (variant == BEAM_V2) {
        // Two exact visited-set formats (beam_search_v2.cuh): (a) 16-bit tags, 7 entries per 16-byte bucket and the query
        // row in shared memory, for ids that split into (bucket, <= 14-bit tag); (b) 32-bit ids, 4 per bucket, the query
        // half-row in registers.  Either way the table holds the mean visited count (12 ef + 200) at <= 75 %
        // load (it closes at 7/8 and diverts to HBM, exactly, beyond), and the CTA shape is the one that keeps the most
        // warps resident under the register budget of the list capacity (v2_shape in beam_search.cuh), 228 KB of shared
        // memory per SM with 1 KB reserved per CTA, and 32 CTAs per SM; whatever shared memory is left goes to the table.
        // CTAs of 4k warps spread evenly over the four schedulers' register partitions.
        const uint32_t fixed = beam_v2_smem_per_warp(C, cp, 0);
        const uint32_t vis16 = env_u32("GBDR_BEAM_VIS16", 2);  // 0 = never, 1/2 = when the shape allows
        const uint32_t force_nb = env_u32("GBDR_BEAM_VIS16_LOGNB", 0);  // tests shrink the table to force spills
        const uint32_t force_b = env_u32("GBDR_BEAM_BPS", 0);           // tuning: cap the resident CTAs per SM
        // the dense build (56 registers, 2 x 17 warps) of lists of <= 64 slots: 0 never, 1 (default) wherever its table
        // still holds the expected visited count: 34 instead of 32 resident warps, 0.594 vs 0.616 ms at SIFT-1M / ef 53
        // (profiles/r3a_*).  A forced CTA shape (tests, tuning) uses the regular build.
        // (2 x 18 warps, all that 56 registers allow, leave the table 201 buckets: 0.618 vs 0.600 ms, run r3k)
        // (the dense build has the default tuning flags compiled in: any other GBDR_BEAM_PF_ROWS runs the regular build)
        const uint32_t dense_mode = (force_w || force_b || env_u32("GBDR_BEAM_PF_ROWS", 7) != 7u) ? 0u : env_u32("GBDR_BEAM_DENSE", 1);
        uint32_t b = 1;
        while (b < 32 && (1ull << b) < n) ++b;
        // (Sizing the table from the MEASURED visited count of earlier launches, and running it at 88-100 % instead of 75 %
        // load at the mean, were both tried: no operating point moved by more than noise, run r3j — resident warps are
        // bounded by registers wherever the table would have shrunk.)
        const uint32_t mean_visited = 12u * ef + 200u;
        struct Pick {
            uint32_t w = 0, b = 0, nb = 0, dense = 0;
        };
        auto search = [&](bool tags, uint32_t nb_min, uint32_t nb_max) {
            Pick best;
            for (uint32_t dense = 0; dense <= 1; ++dense) {
                if (dense && (!tags || cp > 64 || dense_mode == 0)) continue;
                const V2Shape sh = v2_shape((int)(cp / 32), tags, dense != 0);
                const uint32_t reg_warps = (uint32_t)sh.reg_warps, max_wpb = (uint32_t)sh.threads / 32u;
                for (uint32_t wp = 1; wp <= max_wpb; ++wp) {
                    if (force_w && wp != std::min<uint32_t>(force_w, max_wpb)) continue;
                    if (dense && wp != 17) continue;
                    uint32_t bp_hi = std::min<uint32_t>(reg_warps / wp, 32u);
                    if (force_b) bp_hi = std::min(bp_hi, force_b);
                    for (uint32_t bp = bp_hi; bp >= 1; --bp) {
                        if (!dense && !force_w && bp > 1 && (wp & 3u)) continue;
                        const uint32_t cta = std::min<uint32_t>((228u * 1024u) / bp - 1024u, 227u * 1024u);
                        const uint32_t per_warp = (cta / wp) & ~15u;
                        if (per_warp < fixed + 16u * nb_min) continue;
                        const uint32_t nb = std::min<uint32_t>((per_warp - fixed) / 16u, nb_max);
                        // more resident warps first; then CTAs close to 8 warps (a CTA frees its shared memory only when
                        // its last warp is done, so small CTAs let the next batch's CTAs in earlier); then the larger table
                        const uint32_t off8 = wp > 8 ? wp - 8 : 8 - wp, boff8 = best.w > 8 ? best.w - 8 : 8 - best.w;
                        if (wp * bp > best.w * best.b ||
                            (wp * bp == best.w * best.b && (off8 < boff8 || (off8 == boff8 && nb >= best.nb)))) {
                            best.w = wp;
                            best.b = bp;
                            best.nb = nb;
                            best.dense = dense;
                        }
                        break;  // smaller bp only lowers the residency of this wp
                    }
                }
            }
            return best;
        };
        // (b) 32-bit ids, 4 per bucket
        uint32_t nb32 = std::max<uint32_t>(16u, (mean_visited + 2u) / 3u);  // mean visited at 75 % of 4 per bucket
        if (force_h >= 64) nb32 = (force_h & ~63u) / 4u;
        Pick best32 = search(false, nb32, force_h >= 64 ? nb32 : 4096u);
        if (vis16 && !force_h) {
            // (a) enough buckets for the expected visited count AND for the tag to fit: b - floor(log2 buckets) <= 14, i.e.
            // at least 2^(b-14) buckets (a 12.5 M-vertex shard: 1024 buckets = 16 KB — half of what the same number of
            // entries costs in 32-bit slots, which pays for beams whose table would be that large anyway, ef >~ 300, and
            // costs resident warps for small beams: whichever format keeps more warps resident wins, ties go to the tags)
            const uint32_t nb_tag = b > 14 ? 1u << (b - 14) : 1u;
            const uint32_t nb_min = force_nb ? (1u << std::min<uint32_t>(std::max<uint32_t>(force_nb, 2), 12))
                                             : std::max<uint32_t>(std::max<uint32_t>(64u, nb_tag), (mean_visited * 4u + 20u) / 21u);
            const Pick best = search(true, nb_min, force_nb ? nb_min : 4096u);
            uint32_t flog = 0;
            while ((2u << flog) <= best.nb) ++flog;  // floor(log2 nb)
            // b == flog would mean 0-bit tags (a shift by 32 in the kernel)
            if (best.nb && b > flog && b - flog <= 14 && (vis16 == 1 || force_nb || best.w * best.b >= best32.w * best32.b)) {
                plan->vis_bytes = 16u * best.nb;
                plan->vis_hshift = 32u - b;
                plan->vis_tshift = (32u - b) + flog;
                plan->vis_dbits = std::min<uint32_t>(2u, 15u - (b - flog));
                plan->hcap = 7u * best.nb;
                plan->smem_per_warp = beam_v2_smem_per_warp(C, cp, plan->vis_bytes);
                plan->warps_per_block = best.w;
                plan->blocks_per_sm = best.b;
                plan->dense = best.dense;
                return;
            }
        }
        Pick best = best32;
        if (!best.nb) {  // does not fit even with one warp per SM: smallest shape, table as large as the CTA allows
            best.w = best.b = 1;
            best.nb = std::max<uint32_t>(16u, std::min<uint32_t>(nb32, (227u * 1024u - fixed) / 16u));
        }
        plan->hcap = 4u * best.nb;
        plan->vis_bytes = 16u * best.nb;
        plan->smem_per_warp = beam_v2_smem_per_warp(C, cp, plan->vis_bytes);
        plan->warps_per_block = best.w;
        plan->blocks_per_sm = best.b;
        return;
    }
    const bool reg = variant == BEAM_REG_LIST;
    uint32_t hc = 1024;
    while (hc < want && hc < 16384) hc <<= 1;
    BeamLayout L = beam_layout(C, reg ? 0 : cp, hc);
    const uint32_t budget = 200 * 1024;  // per CTA
    while (L.total > budget && hc > 1024) {
        hc >>= 1;
        L = beam_layout(C, reg ? 0 : cp, hc);
    }
    if (force_h >= 64 && (force_h & (force_h - 1)) == 0) {
        hc = force_h;
        L = beam_layout(C, reg ? 0 : cp, hc);
    }
    uint32_t wpb = budget / L.total;
    if (wpb > 8) wpb = 8;
    if (wpb < 1) wpb = 1;
    if (force_w) wpb = force_w;
    plan->hcap = hc;
    plan->warps_per_block = wpb;
    plan->smem_per_warp = L.total;
    plan->blocks_per_sm = std::max<uint32_t>(1, std::min<uint32_t>(reg ? 2 : 32, (227u * 1024u) / (L.total * wpb + 1024u)));
}

int launch_beam(BeamParams& p, const BeamPlan& plan, uint32_t blocks, cudaStream_t st) {
    p.cap = plan.cap;
    p.hcap = plan.hcap;
    p.smem_per_warp = plan.smem_per_warp;
    p.vis_bytes = plan.vis_bytes;
    p.vis_hshift = plan.vis_hshift;
    p.vis_tshift = plan.vis_tshift;
    p.vis_dbits = plan.vis_dbits;
    // 4-slot buckets stay cheap to probe well past the load a one-slot table tolerates
    p.hlimit = plan.variant == BEAM_V2 ? plan.hcap - plan.hcap / 8 : plan.hcap / 2 + plan.hcap / 4;
    // bit 0: +3.4 % at SIFT-1M/ef 53 (L2 hit 26 -> 42 %, profiles/r1n_*); bit 1: adjacency prefetch of accepted candidates
    // only, -2 %; bit 2: atomic visited-set insertion, -6 % (0.669 -> 0.655 -> 0.616 ms, profiles/r3a_*)
    p.pf_rows = env_u32("GBDR_BEAM_PF_ROWS", 7);
    p.hshift = 0;
    if (plan.variant != BEAM_V2) p.hshift = 32 - __builtin_ctz(plan.hcap);
    switch (plan.variant) {
        case BEAM_V2: return launch_beam_search_v2(p, plan.warps_per_block, blocks, plan.dense != 0, st);
        case BEAM_REG_LIST: return launch_beam_search_reg(p, plan.warps_per_block, blocks, st);
        default: return launch_beam_search(p, plan.warps_per_block, blocks, st);
    }
}

template <int C_T>
static int launch_t(const BeamParams& p, uint32_t wpb, uint32_t blocks, uint32_t* counter, cudaStream_t st) {
    const size_t smem = (size_t)p.smem_per_warp * wpb;
    GBDR_CUDA(cudaFuncSetAttribute(beam_search_kernel<C_T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    beam_search_kernel<C_T><<<blocks, wpb * 32, smem, st>>>(p, counter);
    GBDR_CHECK_LAUNCH();
    count_launch();
    return GBDR_OK;
}

int launch_beam_search(const BeamParams& p, uint32_t wpb, uint32_t blocks, cudaStream_t st) {
    // p.status doubles as {status word, work counter}: counter is the word after it
    uint32_t* counter = p.status + 1;
    switch (p.C) {
        case 4: return launch_t<4>(p, wpb, blocks, counter, st);
        case 8: return launch_t<8>(p, wpb, blocks, counter, st);
        case 16: return launch_t<16>(p, wpb, blocks, counter, st);
        default: return launch_t<0>(p, wpb, blocks, counter, st);
    }
}

}  // namespace gbdr
