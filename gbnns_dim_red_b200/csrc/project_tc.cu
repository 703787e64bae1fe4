// project_tc.cu — K1, tensor-core mode: the query projection
//   y = normalize(W3 relu(W2 relu(W1 x + b1) + b2) + b3)
// (reference search/support_func.h:624-658, applied per query at search/search_function.h:354-355) as
// three batched GEMMs on the 5th-generation tensor cores.
//
// Each layer is one launch of linear_tc_kernel: a CTA owns a [128 x N_T] output tile whose fp32
// accumulator lives in TMEM (tcgen05.alloc), a producer warp streams 128-byte-wide K blocks of both
// operands into shared memory with the TMA engine (cp.async.bulk on mbarriers), one elected thread
// issues tcgen05.mma.kind::tf32 and releases the stages with tcgen05.commit, and four epilogue warps
// read the accumulator back with tcgen05.ld, add the bias, apply ReLU and write the NEXT layer's
// operand (or, for the last layer, the L2-normalised rows).
//
// Precision (include/gbdr.h GBDR_PROJ_*):
//   3xTF32 (default)  every fp32 value v is split as hi = rna_tf32(v), lo = rna_tf32(v - hi); the
//                     product is accumulated as hi*hi + lo*hi + hi*lo in fp32 (three MMAs per K step).
//                     |error| on the unit-norm outputs <= 2e-6, the same class as fp32 reassociation.
//   TF32              hi*hi only, |relative error| <= 2e-3.
//
// Operand images.  tcgen05.mma reads K-major operands from shared memory in the canonical
// 128-byte-swizzle layout: row r of a K block at byte r*128, its 16-byte chunk c stored at chunk
// position c ^ (r & 7).  Both operands are kept in global memory already in that form, one contiguous
// image per (tile, K block), so a stage is filled by two to four linear bulk copies: weights are packed
// once in gbdr_index_set_net, activations by pack_rows_kernel (layer 1) or by the previous layer's
// epilogue (layers 2 and 3).
#include <cstring>

#include "kernels.cuh"

namespace gbdr {

namespace {

constexpr uint32_t KBLK = 32;            // floats per K block (128 bytes: one swizzle atom)
constexpr uint32_t MT = 128;             // rows per CTA tile (UMMA M)
constexpr uint32_t A_IMG = MT * 128;     // bytes of one A K-block image (hi or lo)
constexpr uint32_t NT_MAX = 256;

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
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
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// D[tmem] (+)= A[smem] * B[smem]^T, kind::tf32, issued by one thread
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// mbarrier arrive once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns of the accumulator -> 32 registers per thread
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

// fp32 -> (hi, lo) tf32 pair, both exactly representable in tf32 so the tensor core's own
// conversion of the 32-bit containers is a no-op
__device__ __forceinline__ void split_tf32(float v, uint32_t& hi, uint32_t& lo) {
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(v));
    const float rem = __fsub_rn(v, __uint_as_float(hi));
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(rem));
}

// byte offset of element (row r, float k) inside one swizzled K-block image
__device__ __forceinline__ uint32_t img_off(uint32_t r, uint32_t k) {
    return r * 128u + ((((k >> 2) ^ (r & 7u)) << 4) | ((k & 3u) << 2));
}

// shared-memory matrix descriptor: K-major, 128-byte swizzle, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);       // start address
    d |= (uint64_t)1 << 16;                         // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024u >> 4) << 32;              // stride byte offset
    d |= (uint64_t)1 << 46;                         // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                         // SWIZZLE_128B
    return d;
}

// ---------------------------------------------------------------- packing kernels
// X [n x ldx] row-major fp32 -> hi/lo images [m_tiles][KB][128 rows][128 B]; rows >= n and columns
// >= K are zero.  One thread per (row, 16-byte chunk).
__global__ void pack_rows_kernel(const float* __restrict__ X, uint32_t ldx, uint32_t n, uint32_t K, uint32_t KB,
                                 uint32_t rows_padded, uint8_t* __restrict__ out_hi, uint8_t* __restrict__ out_lo,
                                 int want_lo) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t chunks_per_row = KB * 8u;
    if (t >= (uint64_t)rows_padded * chunks_per_row) return;
    const uint32_t row = (uint32_t)(t / chunks_per_row), ch = (uint32_t)(t % chunks_per_row);
    const uint32_t kb = ch >> 3, c = ch & 7u;
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t k = kb * KBLK + c * 4u + i;
        const float v = (row < n && k < K) ? __ldg(X + (size_t)row * ldx + k) : 0.f;
        split_tf32(v, hi[i], lo[i]);
    }
    const uint32_t r = row & (MT - 1), mt = row / MT;
    const size_t off = ((size_t)mt * KB + kb) * A_IMG + r * 128u + ((c ^ (r & 7u)) << 4);
    *reinterpret_cast<uint4*>(out_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    if (want_lo) *reinterpret_cast<uint4*>(out_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

// W [N x (K+1)] (reference layout, bias last) -> hi/lo images [n_tiles][KB][N_T rows][128 B] + bias[n_tiles*N_T]
__global__ void pack_weights_kernel(const float* __restrict__ W, uint32_t N, uint32_t K, uint32_t KB, uint32_t NT,
                                    uint32_t n_tiles, uint8_t* __restrict__ out_hi, uint8_t* __restrict__ out_lo,
                                    float* __restrict__ bias) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t chunks_per_row = KB * 8u;
    const uint32_t rows = NT * n_tiles;
    if (t >= (uint64_t)rows * chunks_per_row) return;
    const uint32_t row = (uint32_t)(t / chunks_per_row), ch = (uint32_t)(t % chunks_per_row);
    const uint32_t kb = ch >> 3, c = ch & 7u;
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t k = kb * KBLK + c * 4u + i;
        const float v = (row < N && k < K) ? __ldg(W + (size_t)row * (K + 1) + k) : 0.f;
        split_tf32(v, hi[i], lo[i]);
    }
    const uint32_t r = row % NT, nt = row / NT;
    const size_t off = ((size_t)nt * KB + kb) * ((size_t)NT * 128u) + r * 128u + ((c ^ (r & 7u)) << 4);
    *reinterpret_cast<uint4*>(out_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(out_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    if (ch == 0) bias[row] = row < N ? __ldg(W + (size_t)row * (K + 1) + K) : 0.f;
}

// ---------------------------------------------------------------- the GEMM
struct LinearTcParams {
    const uint8_t* a_hi;   // [m_tiles][KB][A_IMG]
    const uint8_t* a_lo;
    const uint8_t* b_hi;   // [n_tiles][KB][NT*128]
    const uint8_t* b_lo;
    const float* bias;     // [n_tiles*NT]
    uint32_t KB;           // K blocks of this layer
    uint32_t NT;           // width of the tile this CTA computes (multiple of 32, <= 256)
    uint32_t NT_img;       // width of the weight-image tiles (NT * n_sub): small batches split an image tile into
    uint32_t n_sub;        // n_sub column slices of NT rows each, one CTA per slice, to put more SMs to work
    uint32_t stages;
    uint32_t terms;        // 3 = 3xTF32, 1 = TF32
    uint32_t tmem_cols;    // power of two >= max(32, NT)
    // hidden layer: write the next layer's operand images
    uint8_t* next_hi;      // [m_tiles][KB_next][A_IMG] or null
    uint8_t* next_lo;
    uint32_t KB_next;
    int relu;
    // last layer: normalised rows
    float* out;            // [n_rows x ld_out] or null
    uint32_t ld_out;
    uint32_t n_rows;       // valid rows
    uint32_t d_low;        // valid output columns (last layer)
    uint32_t pdl;          // launched with programmatic stream serialization behind the previous layer: the activations
                           // (a_hi / a_lo) are that layer's output and may only be read after griddepcontrol.wait
};

__global__ void __launch_bounds__(192, 1) linear_tc_kernel(const LinearTcParams p) {
    extern __shared__ uint8_t smem_dyn[];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t mt = blockIdx.x, nt = blockIdx.y / p.n_sub, sub = blockIdx.y % p.n_sub;
    const uint32_t col0 = nt * p.NT_img + sub * p.NT;   // first output column of this CTA
    // a slice of NT rows (a multiple of 8) of a 128-byte-swizzled image tile is itself a valid image
    const uint32_t b_img = p.NT * 128u, b_img_full = p.NT_img * 128u, b_sub = sub * b_img;
    const uint32_t stage_bytes = (A_IMG + b_img) * (p.terms == 3 ? 2u : 1u);
    const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;   // swizzle-128B images need 1024-byte alignment
    const uint32_t bars = base + p.stages * stage_bytes;            // full[stages], empty[stages], tmem_full, tmem_ptr
    const uint32_t full0 = bars, empty0 = bars + 8u * p.stages, tfull = bars + 16u * p.stages, tptr = tfull + 8u;

    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < p.stages; ++s) {
            mbar_init(full0 + 8u * s, 1);
            mbar_init(empty0 + 8u * s, 1);
        }
        mbar_init(tfull, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tptr, p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tptr));
    // Programmatic dependent launch: the next layer's CTAs may start now, on the SMs this grid leaves idle (79 CTAs of a
    // 10 000-query batch on 148 SMs), set up their barriers and TMEM and stream their first weight tiles in; they block in
    // griddepcontrol.wait until this grid has completed before they touch its output.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            const uint8_t* a_hi = p.a_hi + (size_t)mt * p.KB * A_IMG;
            const uint8_t* a_lo = p.a_lo ? p.a_lo + (size_t)mt * p.KB * A_IMG : nullptr;
            const uint8_t* b_hi = p.b_hi + (size_t)nt * p.KB * b_img_full + b_sub;
            const uint8_t* b_lo = p.b_lo + (size_t)nt * p.KB * b_img_full + b_sub;
            for (uint32_t kb = 0; kb < p.KB; ++kb) {
                const uint32_t s = kb % p.stages, it = kb / p.stages;
                if (it > 0) mbar_wait(empty0 + 8u * s, (it - 1) & 1u);
                const uint32_t dst = base + s * stage_bytes, fb = full0 + 8u * s;
                mbar_expect_tx(fb, stage_bytes);
                // weights first: they do not depend on the previous layer
                bulk_g2s(dst + A_IMG, b_hi + (size_t)kb * b_img_full, b_img, fb);
                if (p.terms == 3) bulk_g2s(dst + 2u * A_IMG + b_img, b_lo + (size_t)kb * b_img_full, b_img, fb);
                if (kb == 0 && p.pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
                bulk_g2s(dst, a_hi + (size_t)kb * A_IMG, A_IMG, fb);
                if (p.terms == 3) bulk_g2s(dst + A_IMG + b_img, a_lo + (size_t)kb * A_IMG, A_IMG, fb);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            // instruction descriptor: D=f32, A=B=tf32, both K-major, N>>3 at bit 17, M>>4 at bit 24
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((p.NT >> 3) << 17) | ((MT >> 4) << 24);
            for (uint32_t kb = 0; kb < p.KB; ++kb) {
                const uint32_t s = kb % p.stages, it = kb / p.stages;
                mbar_wait(full0 + 8u * s, it & 1u);
                tc_fence_after();
                const uint32_t sa_hi = base + s * stage_bytes, sb_hi = sa_hi + A_IMG;
                const uint32_t sa_lo = sb_hi + b_img, sb_lo = sa_lo + A_IMG;
#pragma unroll
                for (uint32_t kk = 0; kk < KBLK / 8u; ++kk) {  // UMMA_K = 8 tf32 = 32 bytes inside the swizzle atom
                    const uint32_t ko = kk * 32u;
                    umma_tf32(tmem_base, make_desc(sa_hi + ko), make_desc(sb_hi + ko), idesc, (kb | kk) ? 1u : 0u);
                    if (p.terms == 3) {
                        umma_tf32(tmem_base, make_desc(sa_lo + ko), make_desc(sb_hi + ko), idesc, 1u);
                        umma_tf32(tmem_base, make_desc(sa_hi + ko), make_desc(sb_lo + ko), idesc, 1u);
                    }
                }
                umma_commit(empty0 + 8u * s);   // stage free once these MMAs have read it
            }
            umma_commit(tfull);                 // accumulator complete
        }
    } else {
        // ===== epilogue: warps 2..5, TMEM lane quadrant = warp % 4 =====
        const uint32_t quad = warp & 3u;
        const uint32_t r = quad * 32u + lane;           // row inside the tile
        const uint32_t grow = mt * MT + r;              // global row
        mbar_wait(tfull, 0);
        tc_fence_after();
        const uint32_t trow = tmem_base + ((quad * 32u) << 16);
        const float* bias = p.bias + col0;
        if (p.next_hi) {
            for (uint32_t j = 0; j < p.NT / 32u; ++j) {
                const uint32_t kbn = col0 / 32u + j;   // K block of the next layer
                uint32_t v[32];
                tmem_ld32(trow + j * 32u, v);
                if (kbn >= p.KB_next) continue;
                uint8_t* dh = p.next_hi + ((size_t)mt * p.KB_next + kbn) * A_IMG + r * 128u;
                uint8_t* dl = p.next_lo ? p.next_lo + ((size_t)mt * p.KB_next + kbn) * A_IMG + r * 128u : nullptr;
#pragma unroll
                for (uint32_t c = 0; c < 8; ++c) {
                    uint32_t hi[4], lo[4];
#pragma unroll
                    for (uint32_t i = 0; i < 4; ++i) {
                        float x = __fadd_rn(__uint_as_float(v[c * 4 + i]), __ldg(bias + j * 32u + c * 4u + i));
                        if (p.relu && x < 0.f) x = 0.f;
                        split_tf32(x, hi[i], lo[i]);
                    }
                    const uint32_t off = (c ^ (r & 7u)) << 4;
                    *reinterpret_cast<uint4*>(dh + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                    if (dl) *reinterpret_cast<uint4*>(dl + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                }
            }
        } else {
            // last layer: y = acc + bias; norm = sqrt(L2Metric(y, 0)) over floor(d_low/4)*4 dims in the
            // reference's lane-strided order (normalizeVector, support_func.h:636-642); y /= norm
            L2Acc acc;
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            const uint32_t nchunk = (p.d_low + 31u) / 32u;
            for (uint32_t j = 0; j < nchunk; ++j) {
                uint32_t v[32];
                tmem_ld32(trow + j * 32u, v);
#pragma unroll
                for (uint32_t c = 0; c < 8; ++c) {
                    const uint32_t col = j * 32u + c * 4u;
                    if (col + 4u <= p.d_low) {
                        float4 y;
                        y.x = __fadd_rn(__uint_as_float(v[c * 4 + 0]), __ldg(bias + col + 0));
                        y.y = __fadd_rn(__uint_as_float(v[c * 4 + 1]), __ldg(bias + col + 1));
                        y.z = __fadd_rn(__uint_as_float(v[c * 4 + 2]), __ldg(bias + col + 2));
                        y.w = __fadd_rn(__uint_as_float(v[c * 4 + 3]), __ldg(bias + col + 3));
                        acc.add(y, z);
                    }
                }
            }
            const float norm = __fsqrt_rn(acc.result());
            for (uint32_t j = 0; j < nchunk; ++j) {
                uint32_t v[32];
                tmem_ld32(trow + j * 32u, v);
                if (grow < p.n_rows) {
#pragma unroll
                    for (uint32_t i = 0; i < 32; ++i) {
                        const uint32_t col = j * 32u + i;
                        if (col < p.d_low)
                            p.out[(size_t)grow * p.ld_out + col] =
                                __fdiv_rn(__fadd_rn(__uint_as_float(v[i]), __ldg(bias + col)), norm);
                    }
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, p.tmem_cols);
    }
}

struct LayerPlan {
    uint32_t N = 0, K = 0, KB = 0, NT = 0, n_tiles = 0;
    uint8_t *b_hi = nullptr, *b_lo = nullptr;
    float* bias = nullptr;
};

}  // namespace

struct ProjTcPlan {
    LayerPlan layer[3];
    uint32_t d = 0, dh = 0, dh2 = 0, d_low = 0;
    uint32_t sm_count = 148;
    // activation images, grown on demand
    uint8_t* act[3][2] = {{nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}};  // [layer input][hi/lo]
    uint32_t act_tiles = 0;
};

static uint32_t round_up(uint32_t x, uint32_t m) { return (x + m - 1) / m * m; }

int project_tc_prepare(const float* l1, const float* l2, const float* l3, uint32_t d, uint32_t dh, uint32_t dh2,
                       uint32_t d_low, cudaStream_t st, ProjTcPlan** out) {
    *out = nullptr;
    if (d_low > NT_MAX) return GBDR_OK;  // the last layer's tile must hold a whole row; fp32 path otherwise
    ProjTcPlan* P = new ProjTcPlan();
    P->d = d; P->dh = dh; P->dh2 = dh2; P->d_low = d_low;
    {
        int dev = 0, sms = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms > 0)
            P->sm_count = (uint32_t)sms;
    }
    const float* W[3] = {l1, l2, l3};
    const uint32_t N[3] = {dh, dh2, d_low}, K[3] = {d, dh, dh2};
    for (int i = 0; i < 3; ++i) {
        LayerPlan& L = P->layer[i];
        L.N = N[i]; L.K = K[i];
        L.KB = (K[i] + KBLK - 1) / KBLK;
        L.n_tiles = (round_up(N[i], 32) + NT_MAX - 1) / NT_MAX;
        L.NT = round_up((N[i] + L.n_tiles - 1) / L.n_tiles, 32);
        const size_t img = (size_t)L.n_tiles * L.KB * L.NT * 128u;
        cudaError_t e;
        if ((e = cudaMalloc((void**)&L.b_hi, img)) != cudaSuccess || (e = cudaMalloc((void**)&L.b_lo, img)) != cudaSuccess ||
            (e = cudaMalloc((void**)&L.bias, (size_t)L.n_tiles * L.NT * 4)) != cudaSuccess) {
            set_error(std::string("project_tc_prepare: ") + cudaGetErrorString(e));
            project_tc_destroy(P);
            return GBDR_E_NOMEM;
        }
        const uint64_t threads = (uint64_t)L.n_tiles * L.NT * L.KB * 8u;
        pack_weights_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(W[i], L.N, L.K, L.KB, L.NT, L.n_tiles, L.b_hi,
                                                                              L.b_lo, L.bias);
        cudaError_t le = cudaGetLastError();
        if (le != cudaSuccess) {
            set_error(std::string("pack_weights_kernel: ") + cudaGetErrorString(le));
            project_tc_destroy(P);
            return GBDR_E_CUDA;
        }
        count_launch();
    }
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
        set_error(std::string("project_tc_prepare: ") + cudaGetErrorString(e));
        project_tc_destroy(P);
        return GBDR_E_CUDA;
    }
    *out = P;
    return GBDR_OK;
}

void project_tc_destroy(ProjTcPlan* P) {
    if (!P) return;
    for (auto& L : P->layer) {
        if (L.b_hi) cudaFree(L.b_hi);
        if (L.b_lo) cudaFree(L.b_lo);
        if (L.bias) cudaFree(L.bias);
    }
    for (auto& a : P->act)
        for (auto& b : a)
            if (b) cudaFree(b);
    delete P;
}

static int ensure_act(ProjTcPlan* P, uint32_t m_tiles) {
    if (m_tiles <= P->act_tiles) return GBDR_OK;
    const uint32_t want = m_tiles + m_tiles / 4 + 1;
    for (int i = 0; i < 3; ++i)
        for (int h = 0; h < 2; ++h) {
            if (P->act[i][h]) cudaFree(P->act[i][h]);
            P->act[i][h] = nullptr;
            cudaError_t e = cudaMalloc((void**)&P->act[i][h], (size_t)want * P->layer[i].KB * A_IMG);
            if (e != cudaSuccess) {
                P->act_tiles = 0;
                set_error(std::string("projection workspace: ") + cudaGetErrorString(e));
                return GBDR_E_NOMEM;
            }
        }
    P->act_tiles = want;
    return GBDR_OK;
}

int launch_project_tc(ProjTcPlan* P, const float* X, uint32_t ldx, uint32_t n_q, float* out, uint32_t ld_out,
                      int single_pass, cudaStream_t st) {
    if (n_q == 0) return GBDR_OK;
    const uint32_t m_tiles = (n_q + MT - 1) / MT;
    int rc = ensure_act(P, m_tiles);
    if (rc) return rc;
    const uint32_t terms = single_pass ? 1u : 3u;
    {
        const LayerPlan& L = P->layer[0];
        const uint64_t threads = (uint64_t)m_tiles * MT * L.KB * 8u;
        pack_rows_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(X, ldx, n_q, L.K, L.KB, m_tiles * MT, P->act[0][0],
                                                                           P->act[0][1], terms == 3);
        GBDR_CHECK_LAUNCH();
        count_launch();
    }
    for (int i = 0; i < 3; ++i) {
        const LayerPlan& L = P->layer[i];
        LinearTcParams p;
        memset(&p, 0, sizeof(p));
        p.a_hi = P->act[i][0];
        p.a_lo = terms == 3 ? P->act[i][1] : nullptr;
        p.b_hi = L.b_hi; p.b_lo = L.b_lo; p.bias = L.bias;
        // hidden layers of a small batch: slice the image tiles (>= 64 columns per CTA) while the grid stays within one
        // wave, so that e.g. 1000 GIST queries x 1024 hidden units run on 128 CTAs instead of 32
        uint32_t n_sub = 1;
        if (i < 2)
            for (uint32_t c : {4u, 2u})
                if (L.NT % (32u * c) == 0 && L.NT / c >= 64u && m_tiles * L.n_tiles * c <= P->sm_count) {
                    n_sub = c;
                    break;
                }
        p.KB = L.KB; p.NT = L.NT / n_sub; p.NT_img = L.NT; p.n_sub = n_sub; p.terms = terms;
        const uint32_t stage_bytes = (A_IMG + p.NT * 128u) * (terms == 3 ? 2u : 1u);
        p.stages = std::max<uint32_t>(1, std::min<uint32_t>(std::min<uint32_t>(4, L.KB), (200u * 1024u) / stage_bytes));
        p.tmem_cols = 32;
        while (p.tmem_cols < p.NT) p.tmem_cols <<= 1;
        if (i < 2) {
            p.next_hi = P->act[i + 1][0];
            p.next_lo = terms == 3 ? P->act[i + 1][1] : nullptr;
            p.KB_next = P->layer[i + 1].KB;
            p.relu = 1;
        } else {
            p.out = out; p.ld_out = ld_out; p.d_low = P->d_low;
        }
        p.n_rows = n_q;
        const size_t smem = (size_t)p.stages * stage_bytes + 1024 + 16 * p.stages + 32;
        GBDR_CUDA(cudaFuncSetAttribute(linear_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        // layers 2 and 3 are launched behind their predecessor with programmatic stream serialization (see the kernel)
        static const bool use_pdl = [] {
            const char* e = getenv("GBDR_PROJ_PDL");
            return !(e && *e == '0');
        }();
        p.pdl = (i > 0 && use_pdl) ? 1u : 0u;
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(m_tiles, L.n_tiles * n_sub);
        cfg.blockDim = dim3(192);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = p.pdl ? 1 : 0;
        GBDR_CUDA(cudaLaunchKernelEx(&cfg, linear_tc_kernel, p));
        count_launch();
    }
    return GBDR_OK;
}

}  // namespace gbdr
