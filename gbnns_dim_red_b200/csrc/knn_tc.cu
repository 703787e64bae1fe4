// knn_tc.cu — K4, tensor-core path: brute-force kNN as a tiled distance GEMM fused with per-row
// top-k selection.  Replaces get_nearestneighbors / get_nearestneighbors_partly (reference
// dim_red/support_func.py:20-74, 374-384), the producer of `<ds>_knn_1k_<name>.ivecs`
// (search/prepare_graph.cpp:66).  Results are EXACT: the tensor cores only filter.
//
//   dist2(q, b) = |q|^2 + |b|^2 - 2 q.b,   q.b on tcgen05 (kind::tf32, fp32 accumulate in TMEM)
//
// Inputs are rounded to tf32 once (cvt.rna), so |approx - exact| <= eps = eps_rel * |q| * max|b| with
// eps_rel covering the two input roundings (2^-10 on the dot, x2 in the distance), the tensor core's
// accumulation and a 2x safety factor.  If tau is the k-th smallest APPROXIMATE distance of a row,
// every true top-k member has approx <= tau + 2 eps, so the kernel keeps, per row, all candidates
// under a running tau + 2 eps, and the final pass recomputes the survivors' distances in the canonical
// fp32 arithmetic of the C++ side (common.cuh L2Acc, reference search/support_func.h:107-128) and
// sorts them by (dist, id).  The n x n matrix is never materialised.
//
// Three phases per chunk of query rows (profiles/r1i_*: selection inside the GEMM kernel, with only the
// four filter warps resident, was bound by exposed HBM latency and ran 10x slower than the GEMM):
//   0. thresholds   the same tensor-core kernel over a pseudo-random SAMPLE of S base rows; every row counts
//                   its sample distances in a log-scale histogram in shared memory (16 bins per octave,
//                   relative to 4 max|b|^2) and takes the upper edge of the bin where the count reaches s
//                   as its threshold thr.  s is chosen so that, for the Poisson count of sample points
//                   inside the true k-NN ball, P(thr < tau) ~ 3e-7.  (A per-row sorted list, the first
//                   version, serialised 32 divergent insertion sorts per warp: 50 ms per chunk.)
//   1. scan         tensor-core GEMM over ALL base rows with the thresholds fixed: one FFMA + compare
//                   per element, survivors (approx <= thr + 2 eps) appended to the row's buffer in HBM.
//   2. select       warp per row at full occupancy: VERIFIES count(approx <= thr) >= k (which makes the
//                   buffer a proven superset of the top-k), bisection for tau, exact recompute of the
//                   entries <= tau + 2 eps, bitonic sort by (dist, id), emit.
// Rows that fail the verification, or whose buffer overflowed, are listed for the exact scan kernel, so
// the statistical step can only cost time, never correctness.
//
// The GEMM kernel is one persistent CTA per SM; a CTA takes 128 query rows at a time:
//   warp 0      TMA producer: the row block's operand image once, then 256-column base tiles
//               (cp.async.bulk of pre-swizzled 128-byte K-block images) through a 2-3 stage mbarrier ring
//   warp 1      one lane issues tcgen05.mma (M=128, N=256, K=8 per instruction) into one of TWO
//               256-column TMEM accumulators, so tile t+1 is multiplied while tile t is filtered
//   warps 2-17  filter: thread = (row, 64-column slice); tcgen05.ld 32 columns at a time, one FFMA and one
//               FMNMX per element, the per-element compare only where the running minimum passes.  Sixteen
//               warps because one warp per scheduler left the filter latency-bound at 5 us per tile against
//               0.8 us of MMA (profiles/r1i_knn_tc_*); survivors are appended with one atomicAdd per hit.
//               Phase 0 keeps thread = row (warps 2-5 only): its per-row sorted list has a single writer.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "kernels.cuh"

namespace gbdr {

namespace {

constexpr uint32_t KBLK = 32;            // floats per K block (128 bytes)
constexpr uint32_t MT = 128;             // query rows per block
constexpr uint32_t NB = 256;             // base rows per tile
constexpr uint32_t A_IMG = MT * 128;     // bytes per A K-block image
constexpr uint32_t B_IMG = NB * 128;     // bytes per B K-block image
constexpr uint32_t CAP = 4096;           // candidate slots per row
constexpr uint32_t SEG = CAP / 4;        // ... one segment per 64-column slice of the tiles (single writer, no atomics)
constexpr uint32_t SORT_CAP = 2048;      // survivors sorted exactly per row
constexpr uint32_t KMAX_TC = 1024;
constexpr uint32_t HBINS = 224;          // phase 0: log-scale histogram bins per row (14 octaves x 16)
constexpr uint32_t HIST_BYTES = HBINS / 2 * MT * 4;   // two 16-bit counters per word, word (bin/2) of row r at [bin/2][r]
constexpr uint32_t HKEY0 = (127u - 14u) << 4;         // float bits >> 19 of 2^-14: first bin
constexpr uint32_t SMEM_BUDGET = 227u * 1024u;
constexpr uint32_t CHUNK_BLOCKS = 4;     // row blocks per CTA per chunk
constexpr uint32_t FILTER_WARPS = 16;    // 4 TMEM lane quadrants x 4 column slices
constexpr uint32_t KNN_THREADS = (2 + FILTER_WARPS) * 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
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
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024u >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;  // SWIZZLE_128B
    return d;
}

// X [n x ld] fp32 -> tf32-rounded images [tiles][KB][rows_per_tile][128 B] (128-byte swizzle), exact
// squared norms (+inf for padding rows when pad_inf) and the maximum norm.
__global__ void knn_pack_kernel(const float* __restrict__ X, uint32_t ld, uint64_t row0, uint64_t n, uint32_t d, uint32_t KB,
                                uint64_t gather_mul, uint64_t gather_mod,
                                uint32_t rows_per_tile, uint64_t rows_padded, uint8_t* __restrict__ img,
                                uint8_t* __restrict__ img_lo, float* __restrict__ norms, int pad_inf,
                                uint32_t* __restrict__ max_norm_bits) {
    const uint64_t row = (uint64_t)blockIdx.x * blockDim.y + threadIdx.y;  // 8 lanes (chunks) x KB per row
    if (row >= rows_padded) return;
    const uint32_t c = threadIdx.x & 7u;
    float ss = 0.f;
    for (uint32_t kb = threadIdx.x >> 3; kb < KB; kb += blockDim.x >> 3) {
        uint32_t t[4], tl[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t k = kb * KBLK + c * 4u + i;
            // gather_mod != 0: packed row j is source row (j * gather_mul) % gather_mod (a pseudo-random sample)
            const uint64_t src = gather_mod ? (row * gather_mul) % gather_mod : row0 + row;
            const float v = (row < n && k < d) ? __ldg(X + (size_t)src * ld + k) : 0.f;
            ss = fmaf(v, v, ss);
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t[i]) : "f"(v));
            const float rem = __fsub_rn(v, __uint_as_float(t[i]));
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(tl[i]) : "f"(rem));
        }
        const uint32_t r = (uint32_t)(row % rows_per_tile);
        const uint64_t tile = row / rows_per_tile;
        const size_t off = ((size_t)tile * KB + kb) * ((size_t)rows_per_tile * 128u) + r * 128u + ((c ^ (r & 7u)) << 4);
        *reinterpret_cast<uint4*>(img + off) = make_uint4(t[0], t[1], t[2], t[3]);
        if (img_lo) *reinterpret_cast<uint4*>(img_lo + off) = make_uint4(tl[0], tl[1], tl[2], tl[3]);
    }
    // reduce ss over the blockDim.x lanes of this row (blockDim.x is 8 or 32, a power of two <= 32)
    for (int o = blockDim.x >> 1; o; o >>= 1) ss += __shfl_xor_sync(FULL_MASK, ss, o, 32);
    if (threadIdx.x == 0) {
        const bool valid = row < n;
        norms[row] = valid ? ss : (pad_inf ? __int_as_float(0x7f800000) : 0.f);
        if (valid && max_norm_bits) atomicMax(max_norm_bits, __float_as_uint(ss));
    }
}

struct KnnTcParams {
    const uint8_t* q_img;     // [q_blocks][KB][A_IMG]
    const uint8_t* b_img;     // [b_tiles][KB][B_IMG]
    const uint8_t* q_lo;      // low-order tf32 parts (3xTF32), or null: single-pass tf32
    const uint8_t* b_lo;
    const float* qn;          // [q_blocks*MT] squared norms of the query rows
    const float* bn;          // [b_tiles*NB] squared norms of the base rows, +inf padding
    const uint32_t* bmax_bits;  // max squared base norm (float bits)
    uint32_t KB;
    uint32_t stages;
    uint32_t q_blocks, b_tiles;
    float eps_rel;
    // phase 0 (sample): s-th smallest approximate distance per row -> thr
    uint32_t s;               // 0 = phase 1
    float* thr;               // [q_blocks*MT]  written in phase 0, read in phase 1
    // phase 1 (scan): candidates with approx <= thr + margin
    float* cand_d;            // [q_blocks*MT][CAP]
    uint32_t* cand_i;
    uint32_t* cand_n;         // [q_blocks*MT][4] survivors per segment (> SEG: that segment overflowed)
};

// A row's survivors live in four segments of its buffer (one per column slice); logical index i of the
// concatenation -> physical slot.
struct SegView {
    uint32_t c1, c2, c3, cnt;   // prefix counts of segments 0, 0-1, 0-2 and the total
    __device__ __forceinline__ uint32_t phys(uint32_t i) const {
        const uint32_t seg = (i >= c1 ? 1u : 0u) + (i >= c2 ? 1u : 0u) + (i >= c3 ? 1u : 0u);
        const uint32_t start = seg == 0 ? 0u : seg == 1 ? c1 : seg == 2 ? c2 : c3;
        return seg * SEG + (i - start);
    }
};

// number of entries of d[0..cnt) that are <= t (warp-cooperative)
__device__ __forceinline__ uint32_t count_le(const float* d, uint32_t cnt, float t, int lane) {
    uint32_t c = 0;
    for (uint32_t i = lane; i < cnt; i += 32) c += (d[i] <= t) ? 1u : 0u;
    return __reduce_add_sync(FULL_MASK, c);
}

// a threshold t with count(<= t) >= k, tightened by bisection until the count is within k + 64 (or the
// interval is exhausted: ties).  Requires cnt >= k.  Warp-cooperative; result uniform.
__device__ __forceinline__ float select_threshold(const float* d, uint32_t cnt, uint32_t k, int lane) {
    float lo = __int_as_float(0x7f800000), hi = -lo;
    for (uint32_t i = lane; i < cnt; i += 32) {
        const float x = d[i];
        lo = fminf(lo, x);
        hi = fmaxf(hi, x);
    }
    for (int o = 16; o; o >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(FULL_MASK, lo, o));
        hi = fmaxf(hi, __shfl_xor_sync(FULL_MASK, hi, o));
    }
    if (cnt <= k + 64) return hi;
    for (int it = 0; it < 40; ++it) {
        const float mid = lo + 0.5f * (hi - lo);
        if (!(mid > lo) || !(mid < hi)) break;
        const uint32_t c = count_le(d, cnt, mid, lane);
        if (c >= k) {
            hi = mid;
            if (c <= k + 64) break;
        } else {
            lo = mid;
        }
    }
    return hi;
}

// keep only the entries with d <= t (in-place, order-preserving); returns the new count
__device__ __forceinline__ uint32_t filter_le(float* d, uint32_t* id, uint32_t cnt, float t, int lane) {
    uint32_t out = 0;
    for (uint32_t base = 0; base < cnt; base += 32) {
        const uint32_t i = base + lane;
        float v = 0.f;
        uint32_t vi = 0;
        bool keep = false;
        if (i < cnt) {
            v = d[i];
            vi = id[i];
            keep = v <= t;
        }
        const unsigned m = __ballot_sync(FULL_MASK, keep);
        __syncwarp();
        if (keep) {
            const uint32_t pos = out + __popc(m & lanemask_lt());
            d[pos] = v;
            id[pos] = vi;
        }
        out += __popc(m);
        __syncwarp();
    }
    return out;
}

__device__ __forceinline__ void bitonic_sort_warp(float* sd, uint32_t* si, uint32_t n, int lane) {
    for (uint32_t size = 2; size <= n; size <<= 1) {
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            for (uint32_t t = lane; t < (n >> 1); t += 32) {
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
}

__global__ void __launch_bounds__(KNN_THREADS, 1) knn_tc_kernel(const KnnTcParams p) {
    extern __shared__ uint8_t smem_dyn[];
    const uint32_t warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_dyn + (base - smem_u32(smem_dyn));
    const uint32_t nparts = p.q_lo ? 2u : 1u;                    // hi (+ lo) images per operand
    const uint32_t a_bytes = p.KB * A_IMG * nparts;              // [hi KB blocks][lo KB blocks]
    const uint32_t st_bytes = B_IMG * nparts;                    // per stage: [hi][lo]
    const uint32_t a_s = base;
    const uint32_t b_s = a_s + a_bytes;
    const uint32_t hist_off = a_bytes + p.stages * st_bytes;                        // phase 0 only
    uint32_t* hist = reinterpret_cast<uint32_t*>(base_ptr + hist_off);
    const uint32_t bars = base + hist_off + (p.s ? HIST_BYTES : 0u);
    const uint32_t b_full0 = bars, b_empty0 = bars + 8u * p.stages;
    const uint32_t a_full = bars + 16u * p.stages, a_empty = a_full + 8u;
    const uint32_t t_full0 = a_empty + 8u, t_empty0 = t_full0 + 16u;
    const uint32_t tptr = t_empty0 + 16u;

    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < p.stages; ++s) {
            mbar_init(b_full0 + 8u * s, 1);
            mbar_init(b_empty0 + 8u * s, 1);
        }
        mbar_init(a_full, 1);
        mbar_init(a_empty, 1);
        for (uint32_t i = 0; i < 2; ++i) {
            mbar_init(t_full0 + 8u * i, 1);
            mbar_init(t_empty0 + 8u * i, FILTER_WARPS);   // one arrival per filter warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tptr), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tptr));

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            uint32_t g = 0, bi = 0;
            for (uint32_t blk = blockIdx.x; blk < p.q_blocks; blk += gridDim.x, ++bi) {
                if (bi > 0) mbar_wait(a_empty, (bi - 1) & 1u);
                mbar_expect_tx(a_full, a_bytes);
                for (uint32_t kb = 0; kb < p.KB; ++kb) {
                    bulk_g2s(a_s + kb * A_IMG, p.q_img + ((size_t)blk * p.KB + kb) * A_IMG, A_IMG, a_full);
                    if (p.q_lo)
                        bulk_g2s(a_s + (p.KB + kb) * A_IMG, p.q_lo + ((size_t)blk * p.KB + kb) * A_IMG, A_IMG, a_full);
                }
                for (uint32_t tile = 0; tile < p.b_tiles; ++tile)
                    for (uint32_t kb = 0; kb < p.KB; ++kb, ++g) {
                        const uint32_t s = g % p.stages, it = g / p.stages;
                        if (it > 0) mbar_wait(b_empty0 + 8u * s, (it - 1) & 1u);
                        mbar_expect_tx(b_full0 + 8u * s, st_bytes);
                        bulk_g2s(b_s + s * st_bytes, p.b_img + ((size_t)tile * p.KB + kb) * B_IMG, B_IMG, b_full0 + 8u * s);
                        if (p.b_lo)
                            bulk_g2s(b_s + s * st_bytes + B_IMG, p.b_lo + ((size_t)tile * p.KB + kb) * B_IMG, B_IMG,
                                     b_full0 + 8u * s);
                    }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((NB >> 3) << 17) | ((MT >> 4) << 24);
            uint32_t g = 0, bi = 0, T = 0;
            for (uint32_t blk = blockIdx.x; blk < p.q_blocks; blk += gridDim.x, ++bi) {
                mbar_wait(a_full, bi & 1u);
                for (uint32_t tile = 0; tile < p.b_tiles; ++tile, ++T) {
                    const uint32_t buf = T & 1u;
                    if (T >= 2) mbar_wait(t_empty0 + 8u * buf, ((T >> 1) - 1) & 1u);
                    tc_fence_after();
                    for (uint32_t kb = 0; kb < p.KB; ++kb, ++g) {
                        const uint32_t s = g % p.stages, it = g / p.stages;
                        mbar_wait(b_full0 + 8u * s, it & 1u);
                        tc_fence_after();
#pragma unroll
                        for (uint32_t kk = 0; kk < KBLK / 8u; ++kk) {
                            const uint32_t ah = a_s + kb * A_IMG + kk * 32u, bh = b_s + s * st_bytes + kk * 32u;
                            umma_tf32(tmem_base + buf * NB, make_desc(ah), make_desc(bh), idesc, (kb | kk) ? 1u : 0u);
                            if (p.q_lo) {   // 3xTF32: + lo*hi + hi*lo
                                umma_tf32(tmem_base + buf * NB, make_desc(ah + p.KB * A_IMG), make_desc(bh), idesc, 1u);
                                umma_tf32(tmem_base + buf * NB, make_desc(ah), make_desc(bh + B_IMG), idesc, 1u);
                            }
                        }
                        umma_commit(b_empty0 + 8u * s);
                    }
                    umma_commit(t_full0 + 8u * buf);
                }
                umma_commit(a_empty);
            }
        }
    } else {
        // ===== filter: warps 2..17 =====
        const uint32_t quad = warp & 3u;                 // TMEM lane quadrant this warp may read
        const uint32_t slice = (warp - 2u) >> 2;         // 64-column slice of the tile (phase 1)
        const uint32_t r = quad * 32u + lane;
        const float bmax = sqrtf(__uint_as_float(__ldg(p.bmax_bits)));
        const float INF = __int_as_float(0x7f800000);
        if (p.s) {
            // ---- phase 0: histogram of the sample distances of every row -> threshold ----
            const uint32_t ftid = threadIdx.x - 64u;   // 0..511 among the filter threads
            // distances are binned relative to 4 max|b|^2 (>= any distance between base rows)
            const float scale = 4.f * bmax * bmax + 1e-30f, inv_scale = 1.f / scale;
            uint32_t T = 0;
            for (uint32_t blk = blockIdx.x; blk < p.q_blocks; blk += gridDim.x) {
                for (uint32_t i = ftid; i < HIST_BYTES / 4u; i += FILTER_WARPS * 32u) hist[i] = 0u;
                asm volatile("bar.sync 1, %0;" ::"r"(FILTER_WARPS * 32u) : "memory");
                const uint64_t grow = (uint64_t)blk * MT + r;
                const float qn = __ldg(p.qn + grow);
                for (uint32_t tile = 0; tile < p.b_tiles; ++tile, ++T) {
                    const uint32_t buf = T & 1u;
                    mbar_wait(t_full0 + 8u * buf, (T >> 1) & 1u);
                    tc_fence_after();
                    const uint32_t trow = tmem_base + ((quad * 32u) << 16) + buf * NB + slice * 64u;
                    const float4* bn4 = reinterpret_cast<const float4*>(p.bn + (size_t)tile * NB + slice * 64u);
#pragma unroll
                    for (uint32_t j = 0; j < 2; ++j) {
                        uint32_t v[32];
                        tmem_ld32(trow + j * 32u, v);
#pragma unroll
                        for (uint32_t c = 0; c < 8; ++c) {
                            const float4 b4 = __ldg(bn4 + j * 8u + c);
                            const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                            for (uint32_t e = 0; e < 4; ++e) {
                                // approximate distance relative to the scale; padding columns (+inf norm) and
                                // anything >= the scale fall off the top and are not counted
                                const float dr = fmaxf((fmaf(-2.f, __uint_as_float(v[c * 4 + e]), bb[e]) + qn) * inv_scale, 0.f);
                                const uint32_t key = __float_as_uint(dr) >> 19;
                                const uint32_t bin = key > HKEY0 ? key - HKEY0 : 0u;
                                if (bin < HBINS) atomicAdd(&hist[(bin >> 1) * MT + r], 1u << ((bin & 1u) * 16u));
                            }
                        }
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(t_empty0 + 8u * buf);
                }
                asm volatile("bar.sync 1, %0;" ::"r"(FILTER_WARPS * 32u) : "memory");
                if (slice == 0) {
                    // upper edge of the first bin where the running count reaches s
                    uint32_t cum = 0, bin = HBINS;
                    for (uint32_t w = 0; w < HBINS / 2u; ++w) {
                        const uint32_t c2 = hist[w * MT + r];
                        cum += c2 & 0xFFFFu;
                        if (cum >= p.s) { bin = 2u * w; break; }
                        cum += c2 >> 16;
                        if (cum >= p.s) { bin = 2u * w + 1u; break; }
                    }
                    // (a counter that wrapped past 65535 can only misplace thr: phase 2 verifies it)
                    p.thr[grow] = bin < HBINS ? __uint_as_float((bin + 1u + HKEY0) << 19) * scale : 3.0e38f;
                }
                asm volatile("bar.sync 1, %0;" ::"r"(FILTER_WARPS * 32u) : "memory");
            }
        } else {
            // ---- phase 1: fixed threshold, survivors appended to this thread's segment of the row's buffer ----
            uint32_t T = 0;
            for (uint32_t blk = blockIdx.x; blk < p.q_blocks; blk += gridDim.x) {
                const uint64_t grow = (uint64_t)blk * MT + r;
                const float qn = __ldg(p.qn + grow);
                const float margin = 2.f * p.eps_rel * sqrtf(qn) * bmax + 1e-30f;
                // append iff (bn - 2 dot) <= thr, i.e. approx <= threshold + margin
                const float thr = p.thr[grow] + margin - qn;
                float* my_d = p.cand_d + (size_t)grow * CAP + slice * SEG;
                uint32_t* my_i = p.cand_i + (size_t)grow * CAP + slice * SEG;
                uint32_t cnt = 0;
                for (uint32_t tile = 0; tile < p.b_tiles; ++tile, ++T) {
                    const uint32_t buf = T & 1u;
                    mbar_wait(t_full0 + 8u * buf, (T >> 1) & 1u);
                    tc_fence_after();
                    const uint32_t trow = tmem_base + ((quad * 32u) << 16) + buf * NB + slice * 64u;
                    const float4* bn4 = reinterpret_cast<const float4*>(p.bn + (size_t)tile * NB + slice * 64u);
#pragma unroll
                    for (uint32_t j = 0; j < 2; ++j) {
                        uint32_t v[32];
                        tmem_ld32(trow + j * 32u, v);
                        float t[32], gm[8];
#pragma unroll
                        for (uint32_t c = 0; c < 8; ++c) {
                            const float4 b4 = __ldg(bn4 + j * 8u + c);
                            t[c * 4 + 0] = fmaf(-2.f, __uint_as_float(v[c * 4 + 0]), b4.x);
                            t[c * 4 + 1] = fmaf(-2.f, __uint_as_float(v[c * 4 + 1]), b4.y);
                            t[c * 4 + 2] = fmaf(-2.f, __uint_as_float(v[c * 4 + 2]), b4.z);
                            t[c * 4 + 3] = fmaf(-2.f, __uint_as_float(v[c * 4 + 3]), b4.w);
                            gm[c] = fminf(fminf(t[c * 4 + 0], t[c * 4 + 1]), fminf(t[c * 4 + 2], t[c * 4 + 3]));
                        }
                        const float m = fminf(fminf(fminf(gm[0], gm[1]), fminf(gm[2], gm[3])),
                                              fminf(fminf(gm[4], gm[5]), fminf(gm[6], gm[7])));
                        // survivors are rare (a row keeps ~2 of every 1000 columns) but some lane of the warp has
                        // one in most chunks: the slow path only revisits the groups of four whose minimum passes
                        if (!__any_sync(FULL_MASK, m <= thr)) continue;
#pragma unroll
                        for (uint32_t c = 0; c < 8; ++c) {
                            if (gm[c] <= thr) {
#pragma unroll
                                for (uint32_t e = 0; e < 4; ++e) {
                                    if (t[c * 4 + e] <= thr) {
                                        if (cnt < SEG) {
                                            my_d[cnt] = t[c * 4 + e] + qn;
                                            my_i[cnt] = tile * NB + slice * 64u + j * 32u + c * 4u + e;
                                        }
                                        ++cnt;
                                    }
                                }
                            }
                        }
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(t_empty0 + 8u * buf);
                }
                p.cand_n[grow * 4u + slice] = cnt;   // > SEG: this segment overflowed
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---------------------------------------------------------------- phase 2: exact selection
struct KnnSelParams {
    float* cand_d;            // [rows][CAP] approximate distances (reused as id scratch once staged)
    const uint32_t* cand_i;
    const uint32_t* cand_n;   // [rows][4] per segment
    const float* thr;         // [rows] the fixed threshold phase 1 used (approx space)
    const float* qn;          // [rows]
    const uint32_t* bmax_bits;
    float eps_rel;
    const float* Q;           // row i of this chunk = Q + (q_first + i) * ldq
    uint32_t ldq;
    uint64_t q_first;
    const float* B;
    uint32_t ldb;
    uint32_t C;               // d/4
    uint32_t k;
    uint32_t rows;            // rows of this chunk
    uint32_t* out_ids;        // [rows x k] (already offset to the chunk)
    float* out_dists;         // or null
    uint64_t out_row0;        // row index (within the call) of the chunk's first row, for the redo list
    uint32_t* overflow;       // [0] = count, [1..] = rows for the exact scan kernel
    uint32_t overflow_cap;
};

__global__ void __launch_bounds__(128) knn_select_kernel(const KnnSelParams p) {
    extern __shared__ __align__(16) unsigned char sel_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* sd = reinterpret_cast<float*>(sel_smem) + (size_t)warp * SORT_CAP * 2;
    uint32_t* si = reinterpret_cast<uint32_t*>(sd + SORT_CAP);
    const float INF = __int_as_float(0x7f800000);
    const float bmax = sqrtf(__uint_as_float(__ldg(p.bmax_bits)));
    for (uint32_t row = blockIdx.x * (blockDim.x >> 5) + warp; row < p.rows; row += gridDim.x * (blockDim.x >> 5)) {
        const uint4 n4 = *reinterpret_cast<const uint4*>(p.cand_n + (size_t)row * 4u);
        bool redo = n4.x > SEG || n4.y > SEG || n4.z > SEG || n4.w > SEG;
        SegView sv;
        sv.c1 = n4.x; sv.c2 = n4.x + n4.y; sv.c3 = sv.c2 + n4.z; sv.cnt = sv.c3 + n4.w;
        const uint32_t rc = redo ? 0u : sv.cnt;
        float* rd = p.cand_d + (size_t)row * CAP;
        const uint32_t* ri = p.cand_i + (size_t)row * CAP;
        const float qn = p.qn[row];
        const float margin = 2.f * p.eps_rel * sqrtf(qn) * bmax + 1e-30f;
        uint32_t m = 0;
        if (!redo) {
            // stage the row's approximate distances, segments concatenated, in this warp's shared memory (the
            // bisection re-reads them a dozen times); the same words become sd/si once they are consumed
            float* buf = sd;   // CAP floats = sd + si
            for (uint32_t i = lane; i < rc; i += 32) buf[i] = rd[sv.phys(i)];
            __syncwarp();
            // the buffer holds every element with approx <= thr + margin.  If at least k of them are <= thr,
            // the k-th smallest approximate distance tau is <= thr and every true top-k member
            // (approx <= tau + margin) is in the buffer.
            const uint32_t kk = min(p.k, rc);
            if (count_le(buf, rc, p.thr[row], lane) < p.k) {
                redo = true;   // the sample threshold was too tight for this row (or fewer than k points exist)
            } else {
                const float tau = select_threshold(buf, rc, kk, lane);
                const float cut = fminf(tau, p.thr[row]) + margin;
                const float4* qrow = reinterpret_cast<const float4*>(p.Q + (size_t)(p.q_first + row) * p.ldq);
                uint32_t* kept_ids = reinterpret_cast<uint32_t*>(rd);   // the global copy of the distances is dead: scratch
                for (uint32_t b0 = 0; b0 < rc; b0 += 32) {
                    const uint32_t i = b0 + lane;
                    const bool keep = i < rc && buf[i] <= cut;
                    const unsigned km = __ballot_sync(FULL_MASK, keep);
                    const uint32_t pos = m + __popc(km & lanemask_lt());
                    float ed = 0.f;
                    uint32_t id = 0;
                    if (keep && pos < SORT_CAP) {
                        id = ri[sv.phys(i)];
                        const float4* brow = reinterpret_cast<const float4*>(p.B + (size_t)id * p.ldb);
                        L2Acc acc;
                        for (uint32_t c = 0; c < p.C; ++c) acc.add(__ldg(qrow + c), __ldg(brow + c));
                        ed = acc.result();
                    }
                    __syncwarp();   // every lane has read buf[b0..b0+32): words pos <= i may be overwritten
                    if (keep && pos < SORT_CAP) {
                        sd[pos] = ed;
                        kept_ids[pos] = id;
                    }
                    m += __popc(km);
                }
                if (m > SORT_CAP) redo = true;   // massive ties around tau
                __syncwarp();
                if (!redo)
                    for (uint32_t i = lane; i < m; i += 32) si[i] = kept_ids[i];
            }
        }
        if (redo) {
            if (lane == 0) {
                const uint32_t slot = atomicAdd(p.overflow, 1u);
                if (slot < p.overflow_cap) p.overflow[1 + slot] = (uint32_t)(p.out_row0 + row);
            }
            continue;
        }
        uint32_t ns = 32;
        while (ns < m) ns <<= 1;
        for (uint32_t i = m + lane; i < ns; i += 32) {
            sd[i] = INF;
            si[i] = PAD_ID;
        }
        __syncwarp();
        bitonic_sort_warp(sd, si, ns, lane);
        for (uint32_t i = lane; i < p.k; i += 32) {
            const bool ok = i < m;
            p.out_ids[(size_t)row * p.k + i] = ok ? si[i] : PAD_ID;
            if (p.out_dists) p.out_dists[(size_t)row * p.k + i] = ok ? sd[i] : INF;
        }
        __syncwarp();
    }
}

}  // namespace

bool knn_tc_supported(uint64_t n_rows, uint64_t n, uint32_t d, uint32_t k) {
    return (d % 4 == 0) && d <= 128 && k <= KMAX_TC && n >= 32768 && n_rows >= 1 && n < (1ull << 32) - NB;
}

static uint64_t gcd_u64(uint64_t a, uint64_t b) {
    while (b) {
        const uint64_t t = a % b;
        a = b;
        b = t;
    }
    return a;
}

// Rows [q_begin, q_end) of d_Q against all of d_B.  Device pointers; work is queued on `st`, which is
// synchronised before returning.  `overflow_rows` receives rows the caller must redo with the exact scan.
int launch_knn_tc(const float* d_Q, uint32_t ldq, uint64_t q_begin, uint64_t q_end, const float* d_B, uint32_t ldb, uint64_t n,
                  uint32_t d, uint32_t k, uint32_t* d_out_ids, float* d_out_dists, int sm_count, cudaStream_t st,
                  std::vector<uint32_t>* overflow_rows, KnnHostSink* sink) {
    const uint64_t n_rows = q_end - q_begin;
    const uint32_t KB = (d + KBLK - 1) / KBLK;
    const uint32_t b_tiles = (uint32_t)((n + NB - 1) / NB);
    // 3xTF32 (hi*hi + lo*hi + hi*lo) wherever both operand halves fit in shared memory: the single-pass
    // bound (~4e-3 |q||b|) is too loose for unit-norm embeddings whose k-NN radius^2 is ~1e-2
    const bool three = KB == 1;
    // sample size S and list length s: with lambda = k S / n sample points expected inside the true k-NN
    // ball, thr (the s-th smallest sample distance) is below tau only if >= s of them fall inside:
    // s = lambda + 5 sqrt(lambda) + 6 puts that beyond a 5-sigma Poisson tail; the expected number of
    // survivors in phase 1 is s n / S (relative spread 1/sqrt(s)), kept under CAP / 1.6.
    // The histogram rounds thr up to a bin edge (<= 1/16 in distance, i.e. a few tens of percent in count for
    // intrinsic dimensions around 8), hence the extra 1.25.
    // For small k / n the sample must also be large enough for the s-th sample distance to mean something: with
    // lambda well below 1 the threshold is the ~6th smallest of the sample whatever k is, the buffers fill to ~2000
    // entries per row and a fraction of a percent of the rows overflows (at 12.5 M rows and k = 33 that was 62 000 rows
    // for the exact scan).  lambda >= 4 costs a sample of 4 n / k rows (an eighth of the scan at k = 33) and cuts the
    // expected survivors to ~25 n / S.
    uint64_t S = std::min<uint64_t>(n, std::max<uint64_t>(65536, (4 * n + k - 1) / k));
    uint32_t s_len = 0;
    for (;;) {
        const double lambda = (double)k * (double)S / (double)n;
        s_len = (uint32_t)(lambda + 5.0 * std::sqrt(lambda) + 6.0);
        const double expect = 1.25 * (double)s_len * (double)n / (double)S;
        if ((s_len <= 60000 && expect <= CAP / 1.6) || S >= n) break;
        S = std::min<uint64_t>(n, S * 2);
    }
    if (s_len > 60000 || 1.25 * (double)s_len * (double)n / (double)S > CAP / 1.6) {
        // k too large relative to CAP for the sampling scheme: let the caller use the exact scan
        overflow_rows->clear();
        for (uint64_t r = 0; r < n_rows; ++r) overflow_rows->push_back((uint32_t)r);
        return GBDR_OK;
    }
    S = (S + NB - 1) / NB * NB;
    if (S > n) S = n;
    const uint32_t s_tiles = (uint32_t)((S + NB - 1) / NB);
    uint64_t gmul = 2654435761ull % n;
    if (gmul < 2) gmul = 1;
    while (gcd_u64(gmul, n) != 1) ++gmul;   // (j * gmul) % n is then a permutation of the rows

    const uint32_t chunk_blocks = (uint32_t)sm_count * CHUNK_BLOCKS;
    const uint64_t chunk_rows = (uint64_t)chunk_blocks * MT;
    const uint32_t overflow_cap = 1u << 20;
    uint8_t *q_img = nullptr, *b_img = nullptr, *q_lo = nullptr, *b_lo = nullptr, *s_img = nullptr, *s_lo = nullptr;
    float *qn = nullptr, *bn = nullptr, *sn = nullptr, *cand_d = nullptr, *thr = nullptr;
    uint32_t *cand_i = nullptr, *cand_n = nullptr, *misc = nullptr;
    auto release = [&]() {
        for (void* ptr : {(void*)q_img, (void*)b_img, (void*)q_lo, (void*)b_lo, (void*)s_img, (void*)s_lo, (void*)qn, (void*)bn,
                          (void*)sn, (void*)cand_d, (void*)thr, (void*)cand_i, (void*)cand_n, (void*)misc})
            if (ptr) cudaFreeAsync(ptr, st);
    };
#define KTC_TRY(x)                                                                       \
    do {                                                                                 \
        cudaError_t _e = (x);                                                            \
        if (_e != cudaSuccess) {                                                         \
            set_error(std::string(#x) + ": " + cudaGetErrorString(_e));                  \
            release();                                                                   \
            return GBDR_E_CUDA;                                                          \
        }                                                                                \
    } while (0)
    KTC_TRY(cudaMallocAsync((void**)&q_img, (size_t)chunk_blocks * KB * A_IMG, st));
    KTC_TRY(cudaMallocAsync((void**)&b_img, (size_t)b_tiles * KB * B_IMG, st));
    KTC_TRY(cudaMallocAsync((void**)&s_img, (size_t)s_tiles * KB * B_IMG, st));
    if (three) {
        KTC_TRY(cudaMallocAsync((void**)&q_lo, (size_t)chunk_blocks * KB * A_IMG, st));
        KTC_TRY(cudaMallocAsync((void**)&b_lo, (size_t)b_tiles * KB * B_IMG, st));
        KTC_TRY(cudaMallocAsync((void**)&s_lo, (size_t)s_tiles * KB * B_IMG, st));
    }
    KTC_TRY(cudaMallocAsync((void**)&qn, chunk_rows * 4, st));
    KTC_TRY(cudaMallocAsync((void**)&thr, chunk_rows * 4, st));
    KTC_TRY(cudaMallocAsync((void**)&cand_n, chunk_rows * 4 * 4, st));
    KTC_TRY(cudaMallocAsync((void**)&bn, (size_t)b_tiles * NB * 4, st));
    KTC_TRY(cudaMallocAsync((void**)&sn, (size_t)s_tiles * NB * 4, st));
    KTC_TRY(cudaMallocAsync((void**)&cand_d, chunk_rows * CAP * 4, st));
    KTC_TRY(cudaMallocAsync((void**)&cand_i, chunk_rows * CAP * 4, st));
    KTC_TRY(cudaMallocAsync((void**)&misc, (size_t)(2 + overflow_cap) * 4, st));
    KTC_TRY(cudaMemsetAsync(misc, 0, 8, st));
    const dim3 pblk(KB >= 4 ? 32 : 8, KB >= 4 ? 8 : 32);
    {
        const uint64_t br = (uint64_t)b_tiles * NB, sr = (uint64_t)s_tiles * NB;
        knn_pack_kernel<<<(unsigned)((br + pblk.y - 1) / pblk.y), pblk, 0, st>>>(d_B, ldb, 0, n, d, KB, 0, 0, NB, br, b_img, b_lo, bn,
                                                                                1, misc);
        KTC_TRY(cudaGetLastError());
        knn_pack_kernel<<<(unsigned)((sr + pblk.y - 1) / pblk.y), pblk, 0, st>>>(d_B, ldb, 0, S, d, KB, gmul, n, NB, sr, s_img, s_lo,
                                                                                sn, 1, nullptr);
        KTC_TRY(cudaGetLastError());
        count_launch(2);
    }
    // |approx - exact| <= eps_rel |q| max|b|.  Dot error: single pass 2^-10 (two roundings of 2^-11), 3xTF32
    // 3*2^-22 (dropped lo*lo and second-order residuals); tensor-core accumulation K*2^-23 (truncating adds);
    // doubled in the distance, plus the norms' roundings, x2 safety.
    const float dot_err = three ? 7.2e-7f : 9.8e-4f;
    const float eps_rel = 2.f * (2.f * (dot_err + (float)(KB * KBLK) * 1.2e-7f) + 4e-7f);
    // operand pipeline: as many TMA stages as shared memory holds (the 2-stage version left each 64 KB tile's
    // load latency, ~2.7 us, exposed: 3.7 us per tile against 1 us of MMA); phase 0 gives 56 KB to the histograms
    const uint32_t a_bytes = KB * A_IMG * (three ? 2u : 1u), st_bytes = B_IMG * (three ? 2u : 1u);
    const uint32_t fixed_bytes = a_bytes + 16u * 8u + 96u + 1024u;
    const uint32_t stages1 = std::min<uint32_t>(8u, (SMEM_BUDGET - fixed_bytes) / st_bytes);
    const uint32_t stages0 = std::min<uint32_t>(8u, (SMEM_BUDGET - fixed_bytes - HIST_BYTES) / st_bytes);
    if (stages0 < 2 || stages1 < 2) {
        set_error("knn_tc: row too wide for the shared-memory operand pipeline");
        release();
        return GBDR_E_INVALID;
    }
    const size_t smem1 = (size_t)fixed_bytes + (size_t)stages1 * st_bytes;
    const size_t smem0 = (size_t)fixed_bytes + (size_t)stages0 * st_bytes + HIST_BYTES;
    KTC_TRY(cudaFuncSetAttribute(knn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max(smem0, smem1)));
    const size_t sel_smem = 4u * SORT_CAP * 8u;
    KTC_TRY(cudaFuncSetAttribute(knn_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sel_smem));

    std::vector<cudaEvent_t> chunk_done;
    for (uint64_t c0 = 0; c0 < n_rows; c0 += chunk_rows) {
        const uint64_t rows = std::min<uint64_t>(chunk_rows, n_rows - c0);
        const uint32_t q_blocks = (uint32_t)((rows + MT - 1) / MT);
        const uint32_t grid = std::min<uint32_t>(q_blocks, (uint32_t)sm_count);
        const uint64_t qr = (uint64_t)q_blocks * MT;
        knn_pack_kernel<<<(unsigned)((qr + pblk.y - 1) / pblk.y), pblk, 0, st>>>(d_Q, ldq, q_begin + c0, rows, d, KB, 0, 0, MT, qr,
                                                                                q_img, q_lo, qn, 0, nullptr);
        KTC_TRY(cudaGetLastError());
        KnnTcParams p;
        memset(&p, 0, sizeof(p));
        p.q_img = q_img; p.q_lo = q_lo; p.qn = qn; p.bmax_bits = misc; p.KB = KB;
        p.q_blocks = q_blocks; p.eps_rel = eps_rel; p.thr = thr; p.cand_d = cand_d; p.cand_i = cand_i; p.cand_n = cand_n;
        // phase 0: thresholds from the sample
        p.b_img = s_img; p.b_lo = s_lo; p.bn = sn; p.b_tiles = s_tiles; p.s = s_len; p.stages = stages0;
        knn_tc_kernel<<<grid, KNN_THREADS, smem0, st>>>(p);
        KTC_TRY(cudaGetLastError());
        // phase 1: fixed-threshold scan of the whole base
        p.b_img = b_img; p.b_lo = b_lo; p.bn = bn; p.b_tiles = b_tiles; p.s = 0; p.stages = stages1;
        knn_tc_kernel<<<grid, KNN_THREADS, smem1, st>>>(p);
        KTC_TRY(cudaGetLastError());
        // phase 2: exact selection
        KnnSelParams q;
        memset(&q, 0, sizeof(q));
        q.cand_d = cand_d; q.cand_i = cand_i; q.cand_n = cand_n; q.thr = thr; q.qn = qn; q.bmax_bits = misc; q.eps_rel = eps_rel;
        q.Q = d_Q; q.ldq = ldq; q.q_first = q_begin + c0; q.B = d_B; q.ldb = ldb; q.C = d / 4; q.k = k; q.rows = (uint32_t)rows;
        q.out_ids = d_out_ids + c0 * k; q.out_dists = d_out_dists ? d_out_dists + c0 * k : nullptr; q.out_row0 = c0;
        q.overflow = misc + 1; q.overflow_cap = overflow_cap;
        const uint32_t sgrid = (uint32_t)std::min<uint64_t>((rows + 3) / 4, (uint64_t)sm_count * 3);
        knn_select_kernel<<<sgrid, 128, sel_smem, st>>>(q);
        KTC_TRY(cudaGetLastError());
        count_launch(4);
        if (sink && sink->ids && sink->copy_st) {
            cudaEvent_t ev;
            KTC_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            chunk_done.push_back(ev);
            KTC_TRY(cudaEventRecord(ev, st));
        }
    }
    // every kernel is queued: now stream the finished chunks to the host behind them (a pageable destination makes
    // cudaMemcpyAsync block this thread, which no longer holds up the GPU)
    for (size_t ci = 0; ci < chunk_done.size(); ++ci) {
        const uint64_t c0 = ci * chunk_rows, rows = std::min<uint64_t>(chunk_rows, n_rows - c0);
        cudaStreamWaitEvent(sink->copy_st, chunk_done[ci], 0);
        cudaMemcpyAsync(sink->ids + c0 * k, d_out_ids + c0 * k, rows * k * 4, cudaMemcpyDeviceToHost, sink->copy_st);
        if (sink->dists && d_out_dists)
            cudaMemcpyAsync(sink->dists + c0 * k, d_out_dists + c0 * k, rows * k * 4, cudaMemcpyDeviceToHost, sink->copy_st);
        sink->used = true;
    }
    for (cudaEvent_t ev : chunk_done) cudaEventDestroy(ev);
    uint32_t novf = 0;
    KTC_TRY(cudaMemcpyAsync(&novf, misc + 1, 4, cudaMemcpyDeviceToHost, st));
    KTC_TRY(cudaStreamSynchronize(st));
    overflow_rows->clear();
    if (novf) {
        if (novf > overflow_cap) {
            set_error("knn_tc: too many rows with unbounded candidate sets");
            release();
            return GBDR_E_CAPACITY;
        }
        overflow_rows->resize(novf);
        KTC_TRY(cudaMemcpyAsync(overflow_rows->data(), misc + 2, (size_t)novf * 4, cudaMemcpyDeviceToHost, st));
        KTC_TRY(cudaStreamSynchronize(st));
    }
#undef KTC_TRY
    release();
    return GBDR_OK;
}

}  // namespace gbdr
