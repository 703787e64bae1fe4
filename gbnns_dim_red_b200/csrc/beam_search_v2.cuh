// beam_search_v2.cuh — K2, second-generation register-list kernel for d_low in {16,32,48,64}.
// (Kernel template; beam_search_v2_*.cu instantiate it per list capacity so the translation units build in parallel.)
//
// Same results as beam_search.cu / beam_search_reg.cu (reference search/search_function.h:15-102,
// bit-exact ids, distances, hops and dist_calc) with the per-hop instruction count cut ~2.5x and the
// per-warp footprint cut so that 32 instead of 16 warps (queries) are resident per SM.  The first ncu
// capture (profiles/r1b_*) showed the previous kernel issue-bound on bookkeeping, not HBM-bound:
// ~1000 warp instructions per hop, ~45 % of them in the one-at-a-time sorted insertion.
//
//   * Batched insertion.  All candidates of a hop that pass makeStep's accept test against the
//     worst distance at the start of the hop are merged into the sorted list in one step.  Every
//     candidate gets its rank among the list entries (branch-free binary search on the shared-memory
//     mirror of the list) and among the other candidates (their distances are compacted into shared
//     memory and read back four per LDS.128); rank sum = final position.  The positions form a bit
//     mask P, and output slot j of the merged list is candidate number popc(P below j) when bit j is
//     set, else old entry j - popc(P below j): a pure gather, no per-entry shift counting.  (The third
//     ncu capture, profiles/r1g_*, showed 27 % of all instructions in the previous all-pairs
//     shuffle loop that computed those shifts.)  This is exact because the result of the reference's
//     sequential insert/evict sequence depends only on the set of (dist,id) pairs unless two
//     distances compare equal; any equality seen while ranking (candidate vs list
//     entry, two candidates landing on one position), a tie across the ef boundary, or slack
//     already in use makes the hop fall back to the sequential path, so tie semantics are unchanged.
//   * Entry e of the list lives in lane e & 31, register e >> 5 ("r-major"), so ballots over a
//     register give contiguous position masks, mirror accesses are conflict-free and the best /
//     second-best un-expanded entries are two find-first-set operations.  The mirror carries the
//     "expanded" flags and is authoritative: every list update is "write the mirror, reload the
//     registers".
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
//   * Visited set without atomics or retries.  Ids of one adjacency chunk are distinct; lanes that
//     hash to the same 4-slot bucket are found with match.any and take consecutive free slots by
//     rank, so a chunk is resolved in one pass unless a bucket overflows into its successor.
//   * The (dist,id) list is mirrored in shared memory (the merge scratch), so list ranks of all
//     candidates come from one lane-parallel binary search and broadcasts are single LDS.
//   * Footprint: 16-row stage whose row pads hold the query row and the mbarrier and whose tail doubles as the
//     merge scratch, a 16-bit-tag visited table of any bucket count -> 7 KB and 64 registers per warp at
//     d_low = 32, ef <= 56 (4 CTAs x 8 warps fill the SM's shared memory and register file exactly); with
//     32-bit visited slots the query half-row lives in registers instead (<= 80 registers).
//   * Speculation that never changes results: the guessed next node's adjacency row is loaded a hop early, and the
//     vectors it names are prefetched into L2.
#pragma once
#include <algorithm>

#include "beam_search.cuh"

namespace gbdr {

namespace {

struct V2Layout {
    uint32_t stage_off, q_off, nbr_off, scr_off, cs_off, bar_off, vis_off, total;
};
__host__ __device__ inline V2Layout v2_layout(uint32_t C, uint32_t cap, uint32_t vis_bytes) {
    V2Layout L;
    uint32_t o = 0;
    L.stage_off = o; o += 16u * (C * 16u + 16u);  // rows padded by 16 B (bank spread)
    // the query row: chunk c sits in the 16-byte pad behind staged row c (the bulk copies never touch the
    // pads), the mbarrier in the pad behind row C; the last two pads lie under the merge scratch.  Rows
    // wider than 12 chunks keep separate regions.
    const bool in_pads = C <= 12u;
    L.q_off = o;     o += in_pads ? 0u : C * 16u;
    L.nbr_off = o;   o += 64u * 4u;
    L.scr_off = o;   o += cap * 8u;               // the list mirror
    // merge scratch: first the compacted candidate distances (32 floats + 4 of padding), then, once
    // every lane has its rank, the candidates in rank order (32 pairs).  It overlays the last 256 bytes of
    // the row stage: a merge starts after the last dist16 of its batch and ends before the next gather.
    L.cs_off = L.stage_off + 16u * (C * 16u + 16u) - 256u;
    if (in_pads) {
        L.bar_off = C * (C * 16u + 16u) + C * 16u;
    } else {
        L.bar_off = o; o += 16u;
    }
    L.vis_off = o;   o += vis_bytes;
    L.total = (o + 15u) & ~15u;
    return L;
}

// ---- exact visited set in shared memory (replaces search/visited_list_pool.h) ----
// Two table formats behind one interface.  Both are arrays of 16-byte buckets filled front to back; an
// id hashes to one bucket and overflows to the next one only when that bucket is full, so one LDS.128
// tests a bucket and "not full" proves the id never went further.  Insertion of an adjacency chunk
// (ids distinct across lanes) needs no atomics: lanes that want a slot in the same bucket are grouped
// with match.any and take the bucket's free slots in lane order.
struct VisCtx {
    uint32_t nbuckets;  // any count
    uint32_t hshift;    // Vis16: 32 - b with 2^b >= number of vertices
    uint32_t tshift;    // Vis16: right shift that turns the low product word into the tag (see Vis16::locate)
    uint32_t dbits;     // Vis16: bits of the stored entry that record how many buckets it was displaced
    uint32_t maxdisp;   // Vis16: (1 << dbits) - 1
    // per-warp global overflow table (exact fallback, any id width)
    uint32_t* spill;
    uint32_t spill_cap, spill_shift;
};

// test-and-set in the per-warp global table; true when `id` was not there
__device__ __forceinline__ bool spill_test_and_set(const VisCtx& c, uint32_t id) {
    const uint32_t smask = c.spill_cap - 1;
    uint32_t slot = (id * 0x85EBCA6Bu) >> c.spill_shift;
    for (;;) {
        const uint32_t old = atomicCAS(&c.spill[slot], PAD_ID, id);
        if (old == PAD_ID) return true;
        if (old == id) return false;
        slot = (slot + 1) & smask;
    }
}

// 4 x 32-bit ids per bucket, PAD_ID = empty.  Any number of vertices.
struct Vis32 {
    static constexpr uint32_t SLOTS = 4;
    __device__ static __forceinline__ uint32_t bucket_of(const VisCtx& c, uint32_t id) {
        return __umulhi(id * 0x9E3779B1u, c.nbuckets);
    }
    __device__ static __forceinline__ void clear(uint32_t* vis, const VisCtx& c, int lane) {
        const uint4 fill = make_uint4(PAD_ID, PAD_ID, PAD_ID, PAD_ID);
        for (uint32_t i = lane; i < c.nbuckets; i += 32) reinterpret_cast<uint4*>(vis)[i] = fill;
    }
    __device__ static __forceinline__ void insert_first(uint32_t* vis, const VisCtx& c, uint32_t id) {
        vis[bucket_of(c, id) * 4u] = id;
    }
    // table closed to inserts: is `id` in it?
    __device__ static __forceinline__ bool contains(const uint32_t* vis, const VisCtx& c, uint32_t id) {
        uint32_t g = bucket_of(c, id);
        for (uint32_t guard = 0; guard <= c.nbuckets; ++guard) {  // the closed table keeps a non-full bucket
            const uint4 cur = reinterpret_cast<const uint4*>(vis)[g];
            if (cur.x == id || cur.y == id || cur.z == id || cur.w == id) return true;
            if (cur.w == PAD_ID) break;  // bucket not full: the id never overflowed past it
            g = g + 1 == c.nbuckets ? 0u : g + 1;
        }
        return false;
    }
    // warp-uniform test-and-set of one chunk; true when `id` was not visited before.  `exhausted`: the id
    // is not in the table and could not be placed (never happens with full-width slots).
    __device__ static __forceinline__ bool visit_chunk(uint32_t* vis, const VisCtx& c, uint32_t id, uint32_t& status_acc,
                                                       bool& exhausted) {
        exhausted = false;
        bool pending = id != PAD_ID, isnew = false;
        uint32_t g = bucket_of(c, id);
        unsigned act = __ballot_sync(FULL_MASK, pending);
        uint32_t guard = 0;
        while (act) {
            if (++guard > c.nbuckets + 1u) {  // the table always has a bucket with a free slot
                status_acc |= BEAM_ST_WATCHDOG | 0x100u;
                break;
            }
            if (pending) {
                const uint4 cur = reinterpret_cast<const uint4*>(vis)[g];
                const bool found = (cur.x == id) | (cur.y == id) | (cur.z == id) | (cur.w == id);
                // buckets fill front to back: the first PAD slot is the fill count
                const uint32_t e = cur.x == PAD_ID ? 0u : cur.y == PAD_ID ? 1u : cur.z == PAD_ID ? 2u : cur.w == PAD_ID ? 3u : 4u;
                const unsigned same = __match_any_sync(act, g);
                const unsigned want = __ballot_sync(act, !found);
                __syncwarp(act);  // every lane has read its bucket before any lane writes one
                if (found) {
                    pending = false;  // already visited
                } else {
                    const uint32_t slot = e + __popc(same & want & lanemask_lt());
                    if (slot < 4u) {
                        vis[g * 4u + slot] = id;
                        isnew = true;
                        pending = false;
                    } else {
                        g = g + 1 == c.nbuckets ? 0u : g + 1;  // bucket full
                    }
                }
            }
            __syncwarp();
            act = __ballot_sync(FULL_MASK, pending);
        }
        return isnew;
    }
};

// Half the bytes per entry.  The id is scrambled inside its own b-bit range (an odd multiplier is a bijection
// there) and moved to the top of a 32-bit word H; the 64-bit product H * nbuckets then splits into the home bucket
// (high word: floor(H * nbuckets / 2^32), any bucket count) and, from the low word, a tag: two ids of one bucket
// have low words at least nbuckets * 2^(32-b) apart, so shifting by tshift = (32 - b) + floor(log2 nbuckets) keeps
// them distinct in b - floor(log2 nbuckets) <= 14 bits.  The tag, together with the number of buckets the entry
// was displaced from home (0 .. 2^dbits - 1), is stored as a 16-bit word, so a stored word still identifies the
// id exactly.  Bucket = [count | 7 entries]; 0xFFFF = empty (never a valid entry: tag width + dbits <= 15).  An id
// whose whole probe window is full goes to the global overflow table instead (`exhausted`), and every later
// lookup of it retraces the same full window.
struct Vis16 {
    static constexpr uint32_t SLOTS = 7;
    // home bucket and stored entry (before the displacement bits are added)
    __device__ static __forceinline__ void locate(const VisCtx& c, uint32_t id, uint32_t& g, uint32_t& entry0) {
        const uint32_t H = (id * 0x9E3779B1u) << c.hshift;
        const uint64_t prod = (uint64_t)H * c.nbuckets;
        g = (uint32_t)(prod >> 32);
        entry0 = ((uint32_t)prod >> c.tshift) << c.dbits;
    }
    __device__ static __forceinline__ uint32_t next(const VisCtx& c, uint32_t g) { return g + 1u == c.nbuckets ? 0u : g + 1u; }
    __device__ static __forceinline__ void clear(uint32_t* vis, const VisCtx& c, int lane) {
        const uint4 fill = make_uint4(0xFFFF0000u, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
        for (uint32_t i = lane; i < c.nbuckets; i += 32) reinterpret_cast<uint4*>(vis)[i] = fill;
    }
    __device__ static __forceinline__ void insert_first(uint32_t* vis, const VisCtx& c, uint32_t id) {
        uint32_t g, entry0;
        locate(c, id, g, entry0);
        vis[g * 4u] = (entry0 << 16) | 1u;
    }
    // nonzero iff some 16-bit half of x is zero (the classic has-zero test; flags above the lowest zero half
    // may be spurious, the "any" answer is exact)
    __device__ static __forceinline__ uint32_t any_zero_half(uint32_t x) { return (x - 0x00010001u) & ~x & 0x80008000u; }
    __device__ static __forceinline__ bool found_in(const uint4& cur, uint32_t entry) {
        const uint32_t pat = entry | (entry << 16);
        // word x = [count | entry 0]: only its upper half is an entry
        return ((cur.x >> 16) == entry) |
               ((any_zero_half(cur.y ^ pat) | any_zero_half(cur.z ^ pat) | any_zero_half(cur.w ^ pat)) != 0u);
    }
    // table closed to inserts: is `id` in it?  (false also when its probe window is full: the caller
    // then consults the global table, which is where such an id would be)
    __device__ static __forceinline__ bool contains(const uint32_t* vis, const VisCtx& c, uint32_t id) {
        uint32_t g, entry0;
        locate(c, id, g, entry0);
        for (uint32_t disp = 0; disp < (1u << c.dbits); ++disp) {
            const uint4 cur = reinterpret_cast<const uint4*>(vis)[g];
            if (found_in(cur, entry0 | disp)) return true;
            if ((cur.x & 0xFFFFu) < SLOTS) break;  // bucket not full: the id never went past it
            g = next(c, g);
        }
        return false;
    }
    __device__ static __forceinline__ bool visit_chunk(uint32_t* vis, const VisCtx& c, uint32_t id, uint32_t& status_acc,
                                                       bool& exhausted) {
        exhausted = false;
        bool pending = id != PAD_ID, isnew = false;
        uint32_t g, entry0, disp = 0;
        locate(c, id, g, entry0);
        const uint32_t maxdisp = (1u << c.dbits) - 1u;
        unsigned act = __ballot_sync(FULL_MASK, pending);
        uint32_t guard = 0;
        while (act) {
            if (++guard > maxdisp + 2u) {
                status_acc |= BEAM_ST_WATCHDOG | 0x100u;
                break;
            }
            if (pending) {
                const uint4 cur = reinterpret_cast<const uint4*>(vis)[g];
                const bool found = found_in(cur, entry0 | disp);
                const uint32_t e = cur.x & 0xFFFFu;  // fill count
                const unsigned same = __match_any_sync(act, g);
                const unsigned want = __ballot_sync(act, !found);
                __syncwarp(act);  // every lane has read its bucket before any lane writes one
                if (found) {
                    pending = false;
                } else {
                    const unsigned grp = same & want;
                    const uint32_t rank = __popc(grp & lanemask_lt());
                    uint16_t* b16 = reinterpret_cast<uint16_t*>(vis) + g * 8u;
                    if (rank == 0u && e < SLOTS) b16[0] = (uint16_t)min(e + (uint32_t)__popc(grp), SLOTS);
                    if (e + rank < SLOTS) {
                        b16[1u + e + rank] = (uint16_t)(entry0 | disp);
                        isnew = true;
                        pending = false;
                    } else if (disp == maxdisp) {
                        exhausted = true;  // window full: this id lives in the global table
                        pending = false;
                    } else {
                        g = next(c, g);  // bucket full
                        ++disp;
                    }
                }
            }
            __syncwarp();
            act = __ballot_sync(FULL_MASK, pending);
        }
        return isnew;
    }
};

// Vis16 insertion with a shared-memory atomic on the bucket's fill count instead of match.any grouping: every lane
// walks its own probe sequence (ids of a chunk are distinct, so a lane can never be looking for an entry another lane
// is writing), claims slot `count++` of the first bucket that is not full, and moves on when the claim comes back >= 7
// (other lanes of the chunk filled the bucket first: it IS full now, so the "not full = never overflowed" rule of the
// lookups still holds; counts above 7 just read as full).  compute-sanitizer's racecheck reports the bucket read (LDS.128)
// against another lane's 16-bit entry store as a hazard: that concurrency is the protocol — the reader is never looking for
// the entry being written, and a stale fill count only sends it to the atomic, whose return value decides
// (profiles/r4s_sanitizer.txt).  GBDR_BEAM_PF_ROWS=3 selects the match.any insertion, which has no concurrent writers.
__device__ __forceinline__ bool vis16_visit_chunk_atomic(uint32_t* vis, const VisCtx& c, uint32_t id, bool& exhausted) {
    exhausted = false;
    bool isnew = false;
    if (id != PAD_ID) {
        uint32_t g, entry0;
        Vis16::locate(c, id, g, entry0);
        const uint32_t maxdisp = c.maxdisp;
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
    __syncwarp();
    return isnew;
}

// The same for 32-bit slots: compare-and-swap on the first empty slot; a lost race moves on to the next slot (slots
// still fill front to back: a lane only tries slot s + 1 after it has seen slot s taken).
__device__ __forceinline__ bool vis32_visit_chunk_atomic(uint32_t* vis, const VisCtx& c, uint32_t id) {
    bool isnew = false;
    if (id != PAD_ID) {
        uint32_t g = Vis32::bucket_of(c, id);
        for (uint32_t guard = 0; guard <= c.nbuckets && !isnew; ++guard) {  // the table always keeps a free slot
            const uint4 cur = reinterpret_cast<const uint4*>(vis)[g];
            if (cur.x == id || cur.y == id || cur.z == id || cur.w == id) break;
            uint32_t s = cur.x == PAD_ID ? 0u : cur.y == PAD_ID ? 1u : cur.z == PAD_ID ? 2u : cur.w == PAD_ID ? 3u : 4u;
            for (; s < 4u; ++s)
                if (atomicCAS(&vis[g * 4u + s], PAD_ID, id) == PAD_ID) {
                    isnew = true;
                    break;
                }
            g = g + 1 == c.nbuckets ? 0u : g + 1;  // bucket full
        }
    }
    __syncwarp();
    return isnew;
}

// slow path once the shared table is closed to inserts: look the id up there, then test-and-set in
// the per-warp global overflow table.  true when `id` was not visited before.
template <class V>
__device__ __forceinline__ bool visit_spill(const uint32_t* vis, const VisCtx& c, uint32_t id) {
    if (V::contains(vis, c, id)) return false;
    return spill_test_and_set(c, id);
}

// ---- mbarrier + bulk-copy PTX ----
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// bounded spin: a bulk copy that never lands (it cannot, short of a bug) trips the watchdog instead
// of hanging the device
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
    for (uint32_t spin = 0; spin < (1u << 24); ++spin) {
        uint32_t ok;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (ok) return true;
    }
    return false;
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
    static constexpr bool Q_IN_PADS = C_T <= 12;        // see v2_layout
};
// 16-byte chunk c of the query row
template <int C_T>
__device__ __forceinline__ const unsigned char* q_chunk(const unsigned char* stage, const unsigned char* qsep, int c) {
    return RowGeom<C_T>::Q_IN_PADS ? stage + (size_t)c * RowGeom<C_T>::PITCH + RowGeom<C_T>::ROW_BYTES
                                   : qsep + (size_t)c * 16u;
}

// rows ids[0..mb) (mb <= 16) -> stage: one bulk copy per row, issued by lane r; all lanes wait
template <int C_T>
__device__ __forceinline__ void gather16(uint32_t stage_s, uint32_t bar_s, uint32_t& parity, const uint32_t* ids,
                                         int mb, const float* db, uint32_t row_stride, int lane, uint32_t& status_acc) {
    if (lane == 0) mbar_expect_tx(bar_s, (uint32_t)mb * RowGeom<C_T>::ROW_BYTES);
    // the tail of the stage doubles as the merge scratch (generic-proxy stores): order them before the
    // async-proxy writes of the copies below
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (lane < mb)
        bulk_g2s(stage_s + lane * RowGeom<C_T>::PITCH, db + (size_t)ids[lane] * row_stride, RowGeom<C_T>::ROW_BYTES,
                 bar_s);
    if (!mbar_wait(bar_s, parity)) status_acc |= BEAM_ST_WATCHDOG | 0x400u;
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
// lane 2r (odd lanes hold garbage).  qh[c] = packed (q[4c+2h], q[4c+2h+1]) with h = lane & 1, either
// held in registers or re-read from the query row in shared memory (qs) when registers are short.
template <int C_T, bool Q_REG>
__device__ __forceinline__ float dist16(const unsigned char* stage, const uint64_t (&qh)[Q_REG ? C_T : 1],
                                        const unsigned char* qs, int mb, int lane) {
    const int r = lane >> 1, h = lane & 1;
    float sa = 0.f, sb = 0.f;
    if (r < mb) {
        const uint64_t* row = reinterpret_cast<const uint64_t*>(stage + (size_t)r * RowGeom<C_T>::PITCH) + h;
#pragma unroll
        for (int c = 0; c < C_T; ++c) {
            // packed subtract and square, scalar accumulate: ptxas contracts mul.rn.f32x2 + add.rn.f32x2
            // into FFMA2 (one rounding) even with explicit .rn, which would break bit-exactness
            const uint64_t qc = Q_REG ? qh[Q_REG ? c : 0] : reinterpret_cast<const uint64_t*>(q_chunk<C_T>(stage, qs, c))[h];
            const uint64_t e = f2_sub(qc, row[c * 2]);
            const uint64_t sq = f2_mul(e, e);
            sa = __fadd_rn(sa, __uint_as_float((uint32_t)sq));
            sb = __fadd_rn(sb, __uint_as_float((uint32_t)(sq >> 32)));
        }
    }
    const float t2 = __shfl_down_sync(FULL_MASK, sa, 1), t3 = __shfl_down_sync(FULL_MASK, sb, 1);
    __syncwarp();  // the stage may be overwritten by the next gather
    return __fadd_rn(__fadd_rn(__fadd_rn(sa, sb), t2), t3);
}

__device__ __forceinline__ void prefetch_l2(const void* ptr) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
}

// L2 prefetch of the vectors named by one (speculatively loaded) adjacency row; PAD lanes skip
template <int C_T>
__device__ __forceinline__ void prefetch_rows(const float* db, uint32_t row_stride, uint32_t a0, uint32_t a1) {
    if (a0 != PAD_ID) {
        const float* r = db + (size_t)a0 * row_stride;
        prefetch_l2(r);
        if (C_T > 8) prefetch_l2(r + 32);
    }
    if (a1 != PAD_ID) {
        const float* r = db + (size_t)a1 * row_stride;
        prefetch_l2(r);
        if (C_T > 8) prefetch_l2(r + 32);
    }
}

// Merge the candidates flagged in `am` (one per lane: cdist, cid) into the sorted list.  Requires
// size <= ef, scr[0..CAP) to mirror the list with (+inf, PAD) behind `size`.  Returns false, leaving
// registers and mirror untouched, when an exact distance tie is involved (the caller then applies the
// sequential rules).  candf (36 floats) may alias cs (32 pairs): it is dead before cs is written.
template <int R>
__device__ __forceinline__ bool merge_batch(float (&Ld)[R], uint32_t (&Li)[R], int& size, float& worst, const int ef,
                                            const unsigned am, const float cdist, const uint32_t cid, uint2* scr,
                                            float* candf, uint2* cs, const int lane) {
    constexpr int CAP = 32 * R;
    const float INF = __int_as_float(0x7f800000);
    const bool mine = (am >> lane) & 1u;
    const int na = __popc(am);
    // compacted candidate distances, padded with +inf to a multiple of four
    if (mine) candf[__popc(am & lanemask_lt())] = cdist;
    if (lane < 4) candf[na + lane] = INF;
    // rank among list entries: number of entries < cdist (branch-free lower bound; the mirror holds
    // +inf behind `size`, and size < CAP)
    // (CAP need not be a power of two: with P2 the largest power of two <= CAP, one probe at P2 - 1 decides between
    // the windows [0, P2) and [CAP - P2, CAP), both P2 wide)
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
    unsigned P[R];
    int nbits = 0;
#pragma unroll
    for (int w = 0; w < R; ++w) {
        P[w] = __reduce_or_sync(FULL_MASK, (keep && (np >> 5) == w) ? 1u << (np & 31) : 0u);
        nbits += __popc(P[w]);
    }
    // two candidates on one position = equal distances; equal to a list entry = same
    const int nkeep = __popc(__ballot_sync(FULL_MASK, keep));  // (not inside a short-circuit: every lane votes)
    if (__any_sync(FULL_MASK, eq_list) || nbits != nkeep) return false;
    if (keep) cs[cr] = make_uint2(__float_as_uint(cdist), cid);
    __syncwarp();
    // gather: slot j takes candidate #popc(P below j) or old entry j - popc(P below j)
    uint2 nv[R];
    int below = 0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int cnt = below + __popc(P[r] & lanemask_lt());
        const bool is_c = (P[r] >> lane) & 1u;
        nv[r] = is_c ? cs[cnt] : scr[r * 32 + lane - cnt];
        below += __popc(P[r]);
    }
    __syncwarp();
#pragma unroll
    for (int r = 0; r < R; ++r) scr[r * 32 + lane] = nv[r];
    __syncwarp();
    int nsize = size + na;
    if (nsize > ef) {
        const uint32_t wl = scr[ef - 1].x, wn = scr[ef].x;
        if (wl == wn) {
            // a tie across the ef boundary: the sequential rules decide it.  Undo.
            __syncwarp();
#pragma unroll
            for (int r = 0; r < R; ++r) scr[r * 32 + lane] = make_uint2(__float_as_uint(Ld[r]), Li[r]);
            __syncwarp();
            return false;
        }
        nsize = ef;
        worst = __uint_as_float(wl);
        __syncwarp();
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (r * 32 + lane >= ef) {
                nv[r] = make_uint2(__float_as_uint(INF), PAD_ID);
                scr[r * 32 + lane] = nv[r];
            }
        __syncwarp();
    } else if (nsize == ef) {
        worst = __uint_as_float(scr[ef - 1].x);
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        Ld[r] = __uint_as_float(nv[r].x);
        Li[r] = nv[r].y;
    }
    size = nsize;
    return true;
}

// Register budgets: v2_regs / v2_shape in beam_search.cuh (the host plan sizes CTAs from the same table).  16-bit tags
// keep the query in shared memory -> 64 registers for lists of <= 64 slots (4 x 8 warps per SM) and 8 more per further
// 32 slots; DENSE is the same <= 64-slot kernel squeezed into 56 registers for 2 x 17 warps per SM.
template <int R, int C_T, class V, bool DENSE>
__global__ void __launch_bounds__(v2_shape(R, V::SLOTS == 7, DENSE).threads, v2_shape(R, V::SLOTS == 7, DENSE).min_blocks)
    beam_search_v2_kernel(const BeamParams p, uint32_t* __restrict__ counter) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    constexpr int CAP = 32 * R;
    constexpr int NONE = 0x7fffffff;
    // 16-bit tags: the query row stays in shared memory (the stage pads), which is what fits lists of <= 64 slots into 64 registers
    constexpr bool Q_REG = V::SLOTS != 7;
    const V2Layout Lo = v2_layout(C_T, CAP, p.vis_bytes);
    unsigned char* wbase = smem_raw + (size_t)warp * p.smem_per_warp;
    unsigned char* stage = wbase + Lo.stage_off;
    unsigned char* qs = wbase + Lo.q_off;  // separate query row (rows wider than 12 chunks only)
    uint32_t* nbr = reinterpret_cast<uint32_t*>(wbase + Lo.nbr_off);
    uint2* scr = reinterpret_cast<uint2*>(wbase + Lo.scr_off);
    uint2* cs = reinterpret_cast<uint2*>(wbase + Lo.cs_off);
    uint32_t* vis = reinterpret_cast<uint32_t*>(wbase + Lo.vis_off);
    const uint32_t stage_s = (uint32_t)__cvta_generic_to_shared(stage);
    const uint32_t bar_s = (uint32_t)__cvta_generic_to_shared(wbase + Lo.bar_off);
    uint32_t parity = 0;
    const uint32_t gwarp = blockIdx.x * (blockDim.x >> 5) + warp;
    uint32_t* spill = p.spill + (size_t)gwarp * p.spill_cap;
    const int ef = (int)p.ef;
    const float INF = __int_as_float(0x7f800000);
    VisCtx vc;
    vc.nbuckets = p.vis_bytes / 16u;
    vc.hshift = p.vis_hshift;
    vc.tshift = p.vis_tshift;
    vc.dbits = p.vis_dbits;
    vc.maxdisp = (1u << p.vis_dbits) - 1u;
    vc.spill = spill;
    vc.spill_cap = p.spill_cap;
    vc.spill_shift = p.spill_shift;
    uint32_t status_acc = 0;
    // the shared visited table takes a whole 64-id chunk while vcount + 64 <= hlimit
    const uint32_t open_limit = p.hlimit >= 64u ? p.hlimit - 64u : 0u;
    const bool never_open = p.hlimit < 64u;
    // rows of the searched matrix are exactly C_T chunks long (launch_r checks it), so row addresses are shifts
    constexpr uint32_t ROW_STRIDE = C_T * 4u;
    // tuning flags (BeamParams::pf_rows): the dense build is launched only with the default set, so that its tests fold
    // away (~15 instructions per hop)
    const uint32_t pf_flags = DENSE ? 7u : p.pf_rows;
    if (lane == 0) mbar_init(bar_s, 1);
    __syncwarp();

    for (;;) {
        uint32_t qi = 0;
        if (lane == 0) qi = atomicAdd(counter, 1u);
        qi = __shfl_sync(FULL_MASK, qi, 0);
        if (qi >= p.n_q) break;

        // ---- per-query init ----
        V::clear(vis, vc, lane);
        const float* qg = p.q + (size_t)qi * p.q_stride;
        if (lane < C_T)
            *reinterpret_cast<float4*>(const_cast<unsigned char*>(q_chunk<C_T>(stage, qs, lane))) =
                __ldg(reinterpret_cast<const float4*>(qg) + lane);
        // the list: entry e in lane e & 31, register e >> 5; (+inf, PAD) behind `size`
        float Ld[R];
        uint32_t Li[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            Ld[r] = INF;
            Li[r] = PAD_ID;
            scr[r * 32 + lane] = make_uint2(__float_as_uint(INF), PAD_ID);
        }
        __syncwarp();
        uint64_t qh[Q_REG ? C_T : 1];
        if (Q_REG) {
#pragma unroll
            for (int c = 0; c < (Q_REG ? C_T : 1); ++c)
                qh[c] = reinterpret_cast<const uint64_t*>(q_chunk<C_T>(stage, qs, c))[lane & 1];
        } else {
            qh[0] = 0;
        }

        int size = 0;
        float worst = INF;  // dist of entry ef-1, valid when size >= ef
        int hops = 0, dist_calc = 1, scanned = 0;  // dist_calc starts at 1 (search_function.h:52)
        uint32_t vcount = 0, scount = 0;
        bool spill_ready = false, failed = false;

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
                V::insert_first(vis, vc, e);
            }
            __syncwarp();
            gather16<C_T>(stage_s, bar_s, parity, nbr, 1, p.db, ROW_STRIDE, lane, status_acc);
            float d0 = dist16<C_T, Q_REG>(stage, qh, qs, 1, lane);
            d0 = __shfl_sync(FULL_MASK, d0, 0);
            if (lane == 0) {
                Ld[0] = d0;
                Li[0] = e;
                scr[0] = make_uint2(__float_as_uint(d0), e);
            }
            __syncwarp();
            size = 1;
            if (ef == 1) worst = d0;
            vcount = 1;
        }

        uint32_t pnode = PAD_ID, pa0 = PAD_ID, pa1 = PAD_ID;  // speculatively loaded adjacency row
        bool pf_due = false;  // the rows that adjacency row names have not been prefetched yet

        // ---- main loop (search_function.h:65-91) ----
        while (!failed) {
            // best (and second best) un-expanded entries: the top of candidateSet and its successor.
            // Entries behind `size` carry PAD_ID, whose MSB reads as "expanded".
            int best = NONE, second = NONE;
            if constexpr (R <= 2) {
                // "un-expanded" flags of the (at most two) list registers: best and runner-up are find-first-set operations
                // on 32-bit words (a 64-bit ffs costs twice the instructions)
                const unsigned m0 = __ballot_sync(FULL_MASK, (int)Li[0] >= 0);
                const unsigned m1 = R == 2 ? __ballot_sync(FULL_MASK, (int)Li[R - 1] >= 0) : 0u;
                const unsigned m0b = m0 & (m0 - 1u);                  // m0 without its lowest bit
                const unsigned w1 = m0 ? m0 : m1;                     // word holding the best
                const unsigned w2 = m0b ? m0b : (m0 ? m1 : (m1 & (m1 - 1u)));   // word holding the runner-up
                if (w1) best = __ffs(w1) - 1 + (m0 ? 0 : 32);
                if (w2) second = __ffs(w2) - 1 + (m0b ? 0 : 32);
            } else {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const unsigned m = __ballot_sync(FULL_MASK, (int)Li[r] >= 0);
                    if (m && second == NONE) {
                        const int c1 = r * 32 + __ffs(m) - 1;
                        const unsigned m2 = m & (m - 1);
                        if (best == NONE) {
                            best = c1;
                            if (m2) second = r * 32 + __ffs(m2) - 1;
                        } else {
                            second = c1;
                        }
                    }
                }
            }
            if (best == NONE) break;  // candidateSet empty, or its best is worse than worst (:65,:67)
            int csel = best;
            if (best + 1 < size && scr[best + 1].x == scr[best].x) {
                // ties on dist: the reference pops the largest id first (max-heap of (-dist,id))
                const uint32_t dsel = scr[best].x;
                for (int j = best + 1; j < size; ++j) {
                    const uint2 v = scr[j];
                    if (v.x != dsel) break;
                    if (!(v.y & EXPANDED)) csel = j;
                }
            }
            const uint32_t node = scr[csel].y & ID_MASK;
            // guess the next node: the runner-up of the current list (refined below once the new
            // candidates' distances are known)
            float pdist = INF;
            uint32_t pguess = PAD_ID;
            if (second != NONE && csel == best) {
                const uint2 sv = scr[second];
                pguess = sv.y & ID_MASK;
                pdist = __uint_as_float(sv.x);
            }
            __syncwarp();
            if (lane == (csel & 31)) {
#pragma unroll
                for (int r = 0; r < R; ++r)
                    if (r == (csel >> 5)) {
                        Li[r] |= EXPANDED;
                        scr[csel].y = Li[r];
                    }
            }

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
            pnode = pguess;
            pf_due = false;
            if (pnode != PAD_ID) {
                const uint32_t* prow = p.adj + (size_t)pnode * p.adj_stride;
                pa0 = __ldg(prow + lane);
                pa1 = (32 < p.adj_stride) ? __ldg(prow + 32 + lane) : PAD_ID;
                pf_due = (pf_flags & 1u) != 0u;
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

                const bool smem_open = vcount <= open_limit && !never_open;
                bool n0 = false, n1 = false, x0 = false, x1 = false;
                if (smem_open) {
                    if (V::SLOTS == 7 && (pf_flags & 4u)) {
                        n0 = vis16_visit_chunk_atomic(vis, vc, a0, x0);
                        if (v1) n1 = vis16_visit_chunk_atomic(vis, vc, a1, x1);
                    } else if (V::SLOTS == 4 && (pf_flags & 4u)) {
                        n0 = vis32_visit_chunk_atomic(vis, vc, a0);
                        if (v1) n1 = vis32_visit_chunk_atomic(vis, vc, a1);
                    } else {
                        n0 = V::visit_chunk(vis, vc, a0, status_acc, x0);
                        if (v1) n1 = V::visit_chunk(vis, vc, a1, status_acc, x1);
                    }
                }
                // ids the shared table cannot take (table closed, or their probe window is full) are
                // tracked exactly in the per-warp global table
                const unsigned xm = smem_open ? __ballot_sync(FULL_MASK, x0 | x1) : FULL_MASK;
                uint32_t snew = 0;
                if (xm) {
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
                    if (smem_open) {
                        if (x0) n0 = spill_test_and_set(vc, a0);
                        __syncwarp();
                        if (x1) n1 = spill_test_and_set(vc, a1);
                        __syncwarp();
                        snew = __popc(__ballot_sync(FULL_MASK, x0 && n0)) + __popc(__ballot_sync(FULL_MASK, x1 && n1));
                    } else {
                        if (a0 != PAD_ID) n0 = visit_spill<V>(vis, vc, a0);
                        __syncwarp();
                        if (a1 != PAD_ID) n1 = visit_spill<V>(vis, vc, a1);
                        __syncwarp();
                    }
                }
                const unsigned m0 = __ballot_sync(FULL_MASK, n0);
                const unsigned m1 = __ballot_sync(FULL_MASK, n1);
                const int c0 = __popc(m0), mtot = c0 + __popc(m1);
                if (smem_open) {
                    vcount += mtot - snew;
                    scount += snew;
                } else {
                    scount += mtot;
                }
                if (n0) nbr[__popc(m0 & lanemask_lt())] = a0;
                if (n1) nbr[c0 + __popc(m1 & lanemask_lt())] = a1;
                // whichever of these is expanded next, its adjacency row will be waiting in L2
                // (spends idle HBM bandwidth to take a DRAM round trip off the per-hop critical path)
                if (!(pf_flags & 2u)) {
                    if (n0) prefetch_l2(p.adj + (size_t)a0 * p.adj_stride);
                    if (n1) prefetch_l2(p.adj + (size_t)a1 * p.adj_stride);
                }
                __syncwarp();
                dist_calc += mtot;  // :29
                // the guessed next node's adjacency row (requested at the top of the hop) has arrived by
                // now: pull the vectors it names into L2, so the next hop's gather is an L2 hit
                if (pf_due) {
                    prefetch_rows<C_T>(p.db, ROW_STRIDE, pa0, pa1);
                    pf_due = false;
                }

                for (int b0 = 0; b0 < mtot; b0 += 32) {
                    const int mb = min(32, mtot - b0);
                    // rows b0..b0+15 -> even lanes, rows b0+16..b0+31 -> odd lanes
                    gather16<C_T>(stage_s, bar_s, parity, nbr + b0, min(16, mb), p.db, ROW_STRIDE, lane, status_acc);
                    float cdist = dist16<C_T, Q_REG>(stage, qh, qs, min(16, mb), lane);
                    if (mb > 16) {
                        gather16<C_T>(stage_s, bar_s, parity, nbr + b0 + 16, mb - 16, p.db, ROW_STRIDE, lane, status_acc);
                        const float d1 = dist16<C_T, Q_REG>(stage, qh, qs, mb - 16, lane);
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
                    // only an accepted candidate can ever be expanded: its adjacency row will be waiting in L2
                    if ((pf_flags & 2u) && pre) prefetch_l2(p.adj + (size_t)cid * p.adj_stride);

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
                            pf_due = (pf_flags & 1u) != 0u;
                        }
                    }

                    if (size <= ef && merge_batch<R>(Ld, Li, size, worst, ef, am, cdist, cid, scr, reinterpret_cast<float*>(cs), cs, lane)) continue;

                    // ---- exact-tie fallback: the reference's sequential accept/evict (:31-36) ----
                    for (int row = 0; row < mb; ++row) {
                        const int src = row < 16 ? 2 * row : 2 * (row - 16) + 1;
                        if (failed) break;
                        if (!((am >> src) & 1u)) continue;
                        const float x = __shfl_sync(FULL_MASK, cdist, src);
                        const uint32_t xid = __shfl_sync(FULL_MASK, cid, src);
                        if (size >= ef && !(worst > x)) continue;  // :31
                        // :32-34 sorted insert by (dist,id): shift the tail through the mirror
                        int pos = 0;
#pragma unroll
                        for (int r = 0; r < R; ++r) {
                            const bool less = (r * 32 + lane < size) && pair_less(Ld[r], Li[r] & ID_MASK, x, xid);
                            pos += __popc(__ballot_sync(FULL_MASK, less));
                        }
                        __syncwarp();
#pragma unroll
                        for (int r = 0; r < R; ++r) {
                            const int e = r * 32 + lane;
                            if (e >= pos && e < size && e + 1 < CAP) scr[e + 1] = make_uint2(__float_as_uint(Ld[r]), Li[r]);
                        }
                        if (lane == 0 && pos < CAP) scr[pos] = make_uint2(__float_as_uint(x), xid);
                        __syncwarp();
                        size = size < CAP ? size + 1 : CAP;
#pragma unroll
                        for (int r = 0; r < R; ++r) {
                            const uint2 v = scr[r * 32 + lane];
                            Ld[r] = __uint_as_float(v.x);
                            Li[r] = v.y;
                        }
                        if (size >= ef) {
                            worst = __uint_as_float(scr[ef - 1].x);
                            if (size > ef) {
                                // :35-36 eviction; boundary ties (dist == new worst) stay in the slack
                                int keep = 0;
#pragma unroll
                                for (int r = 0; r < R; ++r) {
                                    const int e = r * 32 + lane;
                                    keep += __popc(__ballot_sync(FULL_MASK, e >= ef && e < size && Ld[r] == worst));
                                }
                                size = ef + keep;
                                if (size >= CAP) {
                                    failed = true;
                                    status_acc |= BEAM_ST_TIE_OVERFLOW;
                                }
                            }
                        }
                    }
                    // restore the invariants merge_batch relies on: (+inf, PAD) behind size, mirror == list
                    __syncwarp();
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const int e = r * 32 + lane;
                        if (e >= size) {
                            Ld[r] = INF;
                            Li[r] = PAD_ID;
                        }
                        scr[e] = make_uint2(__float_as_uint(Ld[r]), Li[r]);
                    }
                    __syncwarp();
                }
                if (failed) break;
                if (v1 != FULL_MASK) break;  // row ended inside this chunk
            }
            if (failed) break;
            if (pf_due) {  // guess refined during this hop: its adjacency row was requested before the merge
                prefetch_rows<C_T>(p.db, ROW_STRIDE, pa0, pa1);
                pf_due = false;
            }
            ++hops;  // :90
            if (hops > dist_calc || (status_acc & BEAM_ST_WATCHDOG)) {  // every hop expands a distinct evaluated vertex
                status_acc |= BEAM_ST_WATCHDOG | 0x200u;
                failed = true;
                break;
            }
        }

        // ---- emit the k best (:96-100) ----
        const int nres = min(min(size, ef), (int)p.k);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int e = r * 32 + lane;
            if (e < (int)p.k) {
                const bool ok = e < nres && !failed;
                p.out_ids[(size_t)qi * p.k + e] = ok ? (Li[r] & ID_MASK) + p.id_offset : PAD_ID;
                if (p.out_dists) p.out_dists[(size_t)qi * p.k + e] = ok ? Ld[r] : INF;
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

template <int R, int C_T, class V, bool DENSE>
int launch_rtv(const BeamParams& p, uint32_t wpb, uint32_t blocks, uint32_t* counter, cudaStream_t st) {
    if (wpb * 32u > (uint32_t)v2_shape(R, V::SLOTS == 7, DENSE).threads) {
        set_error("beam_search_v2: too many warps per CTA for this list capacity / visited format");
        return GBDR_E_INVALID;
    }
    const size_t smem = (size_t)p.smem_per_warp * wpb;
    GBDR_CUDA(cudaFuncSetAttribute(beam_search_v2_kernel<R, C_T, V, DENSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // the grid is persistent (queries are handed out by an atomic counter): never launch more CTAs than are resident
    {
        static thread_local uint32_t seen_wpb = 0, seen_smem = 0, seen_dev = ~0u, seen_blocks = 0;
        int dev = 0;
        GBDR_CUDA(cudaGetDevice(&dev));
        if (seen_wpb != wpb || seen_smem != (uint32_t)smem || seen_dev != (uint32_t)dev) {
            int per_sm = 0, sms = 0;
            GBDR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, beam_search_v2_kernel<R, C_T, V, DENSE>,
                                                                    (int)(wpb * 32), smem));
            GBDR_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
            seen_wpb = wpb;
            seen_smem = (uint32_t)smem;
            seen_dev = (uint32_t)dev;
            seen_blocks = (uint32_t)std::max(1, per_sm * sms);
        }
        blocks = std::min(blocks, seen_blocks);
    }
    beam_search_v2_kernel<R, C_T, V, DENSE><<<blocks, wpb * 32, smem, st>>>(p, counter);
    GBDR_CHECK_LAUNCH();
    count_launch();
    return GBDR_OK;
}

template <int R, int C_T>
int launch_rt(const BeamParams& p, uint32_t wpb, uint32_t blocks, bool dense, uint32_t* counter, cudaStream_t st) {
    if (p.vis_tshift) {
        if constexpr (R <= 2) {
            if (dense) return launch_rtv<R, C_T, Vis16, true>(p, wpb, blocks, counter, st);
        }
        return launch_rtv<R, C_T, Vis16, false>(p, wpb, blocks, counter, st);
    }
    return launch_rtv<R, C_T, Vis32, false>(p, wpb, blocks, counter, st);
}

template <int R>
int launch_r(const BeamParams& p, uint32_t wpb, uint32_t blocks, bool dense, cudaStream_t st) {
    uint32_t* counter = p.status + 1;
    if (p.row_stride != p.C * 4u) {
        set_error("beam_search_v2: rows of the searched matrix must be exactly C chunks long");
        return GBDR_E_INVALID;
    }
    switch (p.C) {
        case 4: return launch_rt<R, 4>(p, wpb, blocks, dense, counter, st);
        case 8: return launch_rt<R, 8>(p, wpb, blocks, dense, counter, st);
        case 12: return launch_rt<R, 12>(p, wpb, blocks, dense, counter, st);
        case 16: return launch_rt<R, 16>(p, wpb, blocks, dense, counter, st);
        default:
            set_error("beam_search_v2: unsupported row width");
            return GBDR_E_INVALID;
    }
}

}  // namespace

// one per translation unit: the list capacities 32 * RA and 32 * RB
#define GBDR_V2_INSTANTIATE(NAME, RA, RB)                                                                          \
    int NAME(const BeamParams& p, uint32_t wpb, uint32_t blocks, bool dense, cudaStream_t st) {                    \
        return p.cap == 32u * (RA) ? launch_r<RA>(p, wpb, blocks, dense, st) : launch_r<RB>(p, wpb, blocks, dense, st); \
    }

}  // namespace gbdr
