// beam_search.cuh — parameters and shared-memory layout of the greedy beam-search kernel (K2).
#pragma once
#include "common.cuh"

namespace gbdr {

struct BeamParams {
    // inputs
    const float* q;          // [n_q_total x q_stride] queries in the searched space, 16B-aligned rows
    uint32_t q_stride;       // floats between query rows
    const float* db;         // [n x row_stride] searched vectors, 16B-aligned rows
    uint32_t row_stride;     // floats between rows
    uint32_t C;              // float4 chunks per row that take part in the distance (= d/4)
    const uint32_t* adj;     // [n x adj_stride] padded adjacency, GBDR_PAD_ID tail
    uint32_t adj_stride;     // multiple of 32
    const uint32_t* aux_adj; // second graph (search_function.h:73-80), same layout; null = single-graph search
    uint32_t aux_stride;
    uint32_t hops_bound;     // the second graph is scanned while hops < hops_bound (:73)
    uint32_t llf;            // "long links first": skip the main row when the auxiliary row produced a candidate (:82)
    const uint32_t* entry;   // [n_q_total]
    uint32_t n_q;            // queries in this launch
    uint32_t ef, k;
    uint32_t cap;            // result-list capacity (ef + tie slack, multiple of 32)
    uint32_t hcap;           // shared-memory visited table slots (power of two)
    uint32_t hshift;         // 32 - log2(hcap)
    uint32_t hlimit;         // inserts after which the shared table is closed
    uint32_t vis_bytes;      // beam_search_v2: bytes of the shared visited table (16-byte buckets)
    uint32_t vis_hshift;     // beam_search_v2, 16-bit tags: 32 - b, 2^b >= vertices; else 0
    uint32_t vis_tshift;     // beam_search_v2, 16-bit tags: (32 - b) + floor(log2 buckets); 0 = 32-bit slots
    uint32_t vis_dbits;      // beam_search_v2, 16-bit tags: displacement bits stored with the tag
    uint32_t* spill;         // [grid_warps x spill_cap] global overflow visited tables
    uint32_t spill_cap;      // power of two
    uint32_t spill_shift;    // 32 - log2(spill_cap)
    uint32_t id_offset;      // added to emitted ids (sharded indexes); 0 when feeding the re-rank
    uint32_t smem_per_warp;  // bytes
    uint32_t pf_rows;        // beam_search_v2 tuning flags (results never depend on them): bit 0 = L2-prefetch the vectors
                             // of the guessed next node's neighbours; bit 1 = L2-prefetch the adjacency rows of accepted
                             // candidates only (clear: of every newly visited vertex); bit 2 = visited-set insertion with
                             // shared-memory atomics (clear: match.any grouping)
    uint32_t n_vertices;     // entry ids >= n_vertices fail their query (PAD results, BEAM_ST_BAD_ENTRY)
    // outputs
    uint32_t* out_ids;       // [n_q_total x k]
    float* out_dists;        // [n_q_total x k] or null
    int32_t* hops;           // or null
    int32_t* dist_calc;      // or null
    int32_t* scanned;        // or null
    int32_t dist_calc_bias;  // added to dist_calc (the +ef of search_function.h:164)
    uint32_t* status;        // single word: OR of per-query failure flags
};

constexpr uint32_t BEAM_ST_SPILLED = 1u;        // informational: some query used the global table
constexpr uint32_t BEAM_ST_VISITED_FULL = 2u;   // failure: global visited table exhausted
constexpr uint32_t BEAM_ST_TIE_OVERFLOW = 4u;   // failure: more boundary ties than list slack
constexpr uint32_t BEAM_ST_BAD_ENTRY = 16u;     // failure: an entry vertex id is not a vertex of the graph
constexpr uint32_t BEAM_ST_WATCHDOG = 8u;       // failure: a loop ran past its proven bound (a bug, never a hang);
                                                // bits 8.. name the loop

struct BeamLayout {
    uint32_t stage_off, q_off, rd_off, rid_off, nbr_off, vis_off, total;
};
__host__ __device__ inline BeamLayout beam_layout(uint32_t C, uint32_t cap, uint32_t hcap) {
    BeamLayout L;
    uint32_t o = 0;
    L.stage_off = o; o += 32u * C * 16u;
    L.q_off = o;     o += C * 16u;
    L.rd_off = o;    o += cap * 4u;
    L.rid_off = o;   o += cap * 4u;
    L.nbr_off = o;   o += 64u * 4u;
    L.vis_off = o;   o += hcap * 4u;
    L.total = (o + 15u) & ~15u;
    return L;
}

#ifdef __CUDACC__
// chunk swizzle: row r, chunk c of C -> physical chunk slot inside the row (bank-conflict-free
// LDS.128 when every lane of a quarter-warp reads chunk c of its own row)
template <int C_T>
__device__ __forceinline__ uint32_t swz(uint32_t r, uint32_t c, uint32_t C) {
    if (C_T == 4) return c ^ ((r >> 1) & 3u);
    if (C_T >= 8 && (C_T & 7) == 0) return c ^ (r & 7u);
    // generic: rotate
    uint32_t x = c + r % C;
    return x >= C ? x - C : x;
}

// exact visited test-and-set; returns true when `id` was not visited before
__device__ __forceinline__ bool visit(uint32_t* vis, uint32_t hcap, uint32_t hshift, bool smem_open,
                                      uint32_t* spill, uint32_t spill_cap, uint32_t spill_shift,
                                      uint32_t id) {
    uint32_t slot = (id * 0x9E3779B1u) >> hshift;
    const uint32_t hmask = hcap - 1;
    for (;;) {
        uint32_t cur = vis[slot];
        if (cur == id) return false;
        if (cur == PAD_ID) {
            if (!smem_open) break;
            uint32_t old = atomicCAS(&vis[slot], PAD_ID, id);
            if (old == PAD_ID) return true;
            if (old == id) return false;
        }
        slot = (slot + 1) & hmask;
    }
    // shared table closed and id not in it: global overflow table
    const uint32_t smask = spill_cap - 1;
    slot = (id * 0x85EBCA6Bu) >> spill_shift;
    for (;;) {
        uint32_t old = atomicCAS(&spill[slot], PAD_ID, id);
        if (old == PAD_ID) return true;
        if (old == id) return false;
        slot = (slot + 1) & smask;
    }
}

#endif  // __CUDACC__

// ---- beam_search_v2: list capacities and their launch bounds, shared by the kernel templates and the host plan ----
// The list holds 32*R entries, R registers per lane.  An SM's register file is four 16 K-register partitions, one per
// scheduler, so the useful budgets are those of a whole number w of warps per scheduler: 64 registers (w = 8, 32 warps
// per SM), 72 (7), 80 (6), 96 (5), 128 (4), 168 (3).  v2_regs() names the budget each variant is compiled for (what
// ptxas needs without spilling more than a few words); everything else follows from it.  (Tighter budgets were tried:
// 96 registers for the 256- and 320-slot lists spill 52-96 bytes in the merge and LOSE 10-30 % despite 20 instead of 16
// resident warps, run r3i.)
struct V2Shape {
    int threads;     // __launch_bounds__ max threads per CTA
    int min_blocks;  // __launch_bounds__ min resident CTAs
    int reg_warps;   // resident warps per SM the register file allows
};
__host__ __device__ constexpr int v2_regs(int R, bool vis16, bool dense) {
    return dense ? 56
           : vis16 ? (R <= 2 ? 64 : R == 3 ? 72 : R <= 5 ? 80 : R == 6 ? 96 : R <= 12 ? 128 : 168)
                   : (R <= 2 ? 80 : R <= 5 ? 96 : R <= 8 ? 128 : 168);
}
__host__ __device__ constexpr V2Shape v2_shape(int R, bool vis16, bool dense) {
    // dense: 2 CTAs x 17 warps = 34 resident warps, so that a 10 000-query batch is two full waves on 148 SMs
    // (5032 slots) instead of 2.11 waves of 4736
    const int regs = v2_regs(R, vis16, dense);
    // (one CTA of all the resident warps as the bound: any split into smaller CTAs then fits as well)
    return regs == 56    ? V2Shape{544, 2, 34}
           : regs == 64  ? V2Shape{1024, 1, 32}
           : regs == 72  ? V2Shape{896, 1, 28}
           : regs == 80  ? V2Shape{768, 1, 24}
           : regs == 96  ? V2Shape{640, 1, 20}
           : regs == 128 ? V2Shape{512, 1, 16}
                         : V2Shape{384, 1, 12};
}
constexpr uint32_t V2_CAPS[] = {32, 64, 96, 128, 160, 192, 256, 320, 384, 512};

// ---- host side: variant selection and launch (beam_search.cu) ----
enum BeamVariant { BEAM_SMEM_LIST = 0, BEAM_REG_LIST = 1, BEAM_V2 = 2 };
struct BeamPlan {
    int variant;
    uint32_t cap;             // result-list capacity
    uint32_t hcap;            // shared visited-table slots
    uint32_t vis_bytes;       // v2: table bytes
    uint32_t vis_hshift, vis_tshift, vis_dbits;  // v2: 16-bit tag format (tshift == 0: 32-bit slots)
    uint32_t warps_per_block;
    uint32_t blocks_per_sm;
    uint32_t smem_per_warp;
    uint32_t dense;           // v2: the 34-warps-per-SM build of the <= 64-slot kernels (56 registers)
};
// picks kernel variant, list capacity, visited-table size/format and launch geometry for (ef, C = d/4)
// on an index of n vertices.  Environment overrides (tests, tuning): GBDR_BEAM_VARIANT = smem | reg | v2,
// GBDR_BEAM_HCAP, GBDR_BEAM_WPB, GBDR_BEAM_VIS16 = 0 | 1.
// second_graph: the two-adjacency mode runs in the shared-memory-list kernel only
void beam_plan(uint32_t ef, uint32_t C, uint64_t n, BeamPlan* plan, bool second_graph = false);
// fills p.cap/hcap/hshift/hlimit/smem_per_warp from the plan and launches `blocks` CTAs
int launch_beam(BeamParams& p, const BeamPlan& plan, uint32_t blocks, cudaStream_t stream);

int launch_beam_search(const BeamParams& p, uint32_t warps_per_block, uint32_t blocks, cudaStream_t stream);
// register-resident list variant (beam_search_reg.cu), p.cap in {32,64,128,256}
int launch_beam_search_reg(const BeamParams& p, uint32_t warps_per_block, uint32_t blocks, cudaStream_t stream);
// batched-merge variant (beam_search_v2.cuh, instantiated per list capacity in beam_search_v2_*.cu), C in {4,8,12,16}
int launch_beam_search_v2(const BeamParams& p, uint32_t warps_per_block, uint32_t blocks, bool dense, cudaStream_t stream);
bool beam_v2_supports(uint32_t C);
uint32_t beam_v2_smem_per_warp(uint32_t C, uint32_t cap, uint32_t vis_bytes);

}  // namespace gbdr
