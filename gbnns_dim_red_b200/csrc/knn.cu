// knn.cu — K4 (exact CUDA-core path): brute-force k nearest neighbours with fused per-row top-k
// selection.  Replaces get_nearestneighbors / get_nearestneighbors_partly (reference
// dim_red/support_func.py:20-74, 374-384) which feed prepare_graph.cpp's input file
// `<ds>_knn_1k_<name>.ivecs` (search/prepare_graph.cpp:66) and the ground-truth files.
//
// Distances are the canonical direct-difference fp32 form of the C++ side (common.cuh L2Acc,
// reference search/support_func.h:107-128) so results are ordered exactly by (dist, id); the n x n
// distance matrix is never materialised:
//   * one thread per query row (128 rows per CTA), base rows streamed through shared memory in
//     tiles (broadcast LDS), partial sums carried across 32-dim chunks in registers;
//   * every row keeps a running threshold tau (its current k-th distance); survivors are appended to
//     a per-row candidate buffer; a full buffer is compacted by a CTA-wide bitonic sort that also
//     tightens tau.  Expected appends per row ~ CAPB + k ln(n/CAPB).
#include "kernels.cuh"

namespace gbdr {

namespace {

constexpr int KQB = 128;  // query rows per CTA (= threads)
constexpr int KBT = 16;   // base rows per tile
constexpr int KKC = 8;    // float4 chunks per k-step (32 dims)
constexpr int KQS = KKC * 4 + 4;  // padded query row stride in floats (conflict-free LDS.128)

struct KnnParams {
    const float* Q;       // [.. x ldq]
    uint32_t ldq;
    uint64_t q_begin, q_end;
    const float* B;       // [n x ldb]
    uint32_t ldb;
    uint64_t n;
    uint32_t C;           // d/4
    uint32_t k;
    uint32_t capb;        // candidate buffer slots per row (power of two, >= 2k)
    uint2* cand;          // [grid*KQB x capb] (dist bits, id)
    uint32_t* out_ids;    // [(q_end-q_begin) x k]
    float* out_dists;     // or null
    uint32_t* counter;    // row-block work counter
};

__device__ __forceinline__ void bitonic_sort_pairs(float* sd, uint32_t* si, uint32_t n, int tid) {
    for (uint32_t size = 2; size <= n; size <<= 1) {
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            for (uint32_t t = tid; t < (n >> 1); t += KQB) {
                const uint32_t lo = 2 * t - (t & (stride - 1));
                const uint32_t hi = lo + stride;
                const bool up = ((lo & size) == 0);
                const float dl = sd[lo], dh = sd[hi];
                const uint32_t il = si[lo], ih = si[hi];
                const bool lt = pair_less(dh, ih, dl, il);  // hi < lo
                if (lt == up) {
                    sd[lo] = dh; sd[hi] = dl;
                    si[lo] = ih; si[hi] = il;
                }
            }
            __syncthreads();
        }
    }
}

__global__ void __launch_bounds__(KQB) knn_scan_kernel(const KnnParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* qs = reinterpret_cast<float*>(smem_raw);                 // [KQB][KQS]
    float4* bs = reinterpret_cast<float4*>(qs + KQB * KQS);        // [2][KBT][KKC]
    float* sd = reinterpret_cast<float*>(bs + 2 * KBT * KKC);      // [capb]
    uint32_t* si = reinterpret_cast<uint32_t*>(sd + p.capb);       // [capb]
    uint32_t* cnt_s = si + p.capb;                                 // [KQB]
    float* tau_s = reinterpret_cast<float*>(cnt_s + KQB);          // [KQB]
    __shared__ uint32_t blk_s;

    const int tid = threadIdx.x;
    const uint32_t C = p.C;
    const uint32_t nkc = (C + KKC - 1) / KKC;
    const uint64_t nrows = p.q_end - p.q_begin;
    const uint32_t nblocks = (uint32_t)((nrows + KQB - 1) / KQB);
    uint2* mycand = p.cand + ((size_t)blockIdx.x * KQB + tid) * p.capb;
    const float INF = __int_as_float(0x7f800000);

    // compaction of row r (CTA-wide): keep the k best, tighten tau; when `final`, emit the result
    auto compact_row = [&](uint32_t r, uint64_t out_row, bool final) {
        const uint32_t c = cnt_s[r];
        uint2* rc = p.cand + ((size_t)blockIdx.x * KQB + r) * p.capb;
        for (uint32_t i = tid; i < p.capb; i += KQB) {
            if (i < c) {
                uint2 v = rc[i];
                sd[i] = __uint_as_float(v.x);
                si[i] = v.y;
            } else {
                sd[i] = INF;
                si[i] = PAD_ID;
            }
        }
        __syncthreads();
        bitonic_sort_pairs(sd, si, p.capb, tid);
        const uint32_t keep = c < p.k ? c : p.k;
        if (!final) {
            for (uint32_t i = tid; i < keep; i += KQB) rc[i] = make_uint2(__float_as_uint(sd[i]), si[i]);
            if (tid == 0) {
                cnt_s[r] = keep;
                if (c >= p.k) tau_s[r] = sd[p.k - 1];
            }
        } else {
            for (uint32_t i = tid; i < p.k; i += KQB) {
                p.out_ids[out_row * p.k + i] = i < keep ? si[i] : PAD_ID;
                if (p.out_dists) p.out_dists[out_row * p.k + i] = i < keep ? sd[i] : INF;
            }
        }
        __syncthreads();
    };

    for (;;) {
        if (tid == 0) blk_s = atomicAdd(p.counter, 1u);
        __syncthreads();
        const uint32_t blk = blk_s;
        __syncthreads();
        if (blk >= nblocks) break;
        const uint64_t row0 = p.q_begin + (uint64_t)blk * KQB;
        const uint64_t myrow = row0 + tid;
        const bool rowok = myrow < p.q_end;
        const float* qrow = p.Q + (size_t)(rowok ? myrow : p.q_begin) * p.ldq;
        cnt_s[tid] = 0;
        tau_s[tid] = INF;
        __syncthreads();

        float4 qreg[KKC];
        if (nkc == 1) {
#pragma unroll
            for (int c = 0; c < KKC; ++c)
                qreg[c] = (uint32_t)c < C ? __ldg(reinterpret_cast<const float4*>(qrow) + c) : make_float4(0, 0, 0, 0);
        }

        const uint64_t ntiles = (p.n + KBT - 1) / KBT;
        uint32_t cnt = 0;
        float tau = INF;
        for (uint64_t tile = 0; tile < ntiles; ++tile) {
            L2Acc acc[KBT];
            for (uint32_t kc = 0; kc < nkc; ++kc) {
                const uint32_t cbeg = kc * KKC;
                const uint32_t cw = min((uint32_t)KKC, C - cbeg);
                float4* bt = bs + ((tile * nkc + kc) & 1) * KBT * KKC;
                // stage B chunk: KBT rows x KKC chunks = 128 float4, one per thread
                {
                    const uint32_t r = tid / KKC, c = tid % KKC;
                    const uint64_t g = tile * KBT + r;
                    if (g < p.n && c < cw)
                        cp_async16(bt + r * KKC + c, p.B + (size_t)g * p.ldb + (size_t)(cbeg + c) * 4u);
                    else
                        bt[r * KKC + c] = make_float4(0, 0, 0, 0);
                }
                if (nkc > 1) {
                    // stage Q chunk [KQB x KKC]
                    for (uint32_t i = tid; i < KQB * KKC; i += KQB) {
                        const uint32_t r = i / KKC, c = i % KKC;
                        const uint64_t g = row0 + r;
                        float4* dst = reinterpret_cast<float4*>(qs + r * KQS) + c;
                        if (g < p.q_end && c < cw)
                            cp_async16(dst, p.Q + (size_t)g * p.ldq + (size_t)(cbeg + c) * 4u);
                        else
                            *dst = make_float4(0, 0, 0, 0);
                    }
                }
                cp_async_commit();
                cp_async_wait<0>();
                __syncthreads();
                if (nkc > 1) {
#pragma unroll
                    for (int c = 0; c < KKC; ++c) qreg[c] = reinterpret_cast<const float4*>(qs + tid * KQS)[c];
                }
                // zero-padded chunks add (0-0)^2 = +0 and leave the canonical sums unchanged
#pragma unroll
                for (int j = 0; j < KBT; ++j) {
#pragma unroll
                    for (int c = 0; c < KKC; ++c) acc[j].add(qreg[c], bt[j * KKC + c]);
                }
                if (nkc > 1) __syncthreads();  // qs is single-buffered
            }
            // test against tau and append
            if (rowok) {
#pragma unroll
                for (int j = 0; j < KBT; ++j) {
                    const uint64_t g = tile * KBT + j;
                    const float dist = acc[j].result();
                    if (g < p.n && dist <= tau) {
                        mycand[cnt] = make_uint2(__float_as_uint(dist), (uint32_t)g);
                        ++cnt;
                    }
                }
            }
            const bool need = cnt + KBT > p.capb;
            if (__syncthreads_or(need)) {
                cnt_s[tid] = cnt;
                __syncthreads();
                for (uint32_t r = 0; r < KQB; ++r) {
                    if (cnt_s[r] + KBT > p.capb) compact_row(r, 0, false);
                }
                cnt = cnt_s[tid];
                tau = tau_s[tid];
            }
        }
        // final selection
        cnt_s[tid] = cnt;
        __syncthreads();
        for (uint32_t r = 0; r < KQB; ++r) {
            const uint64_t g = row0 + r;
            if (g < p.q_end) compact_row(r, g - p.q_begin, true);
        }
        __syncthreads();
    }
}

}  // namespace

uint32_t knn_capb(uint32_t k) {
    uint32_t c = 512;
    while (c < 4 * k) c <<= 1;
    return c;
}

// workspace: `cand` must hold grid*KQB*capb uint2, `counter` one zeroed uint32
int launch_knn_scan(const float* Q, uint32_t ldq, uint64_t q_begin, uint64_t q_end, const float* B,
                    uint32_t ldb, uint64_t n, uint32_t d, uint32_t k, uint32_t* out_ids, float* out_dists,
                    uint2* cand, uint32_t* counter, uint32_t grid, cudaStream_t st) {
    KnnParams p;
    p.Q = Q; p.ldq = ldq; p.q_begin = q_begin; p.q_end = q_end;
    p.B = B; p.ldb = ldb; p.n = n; p.C = d / 4; p.k = k;
    p.capb = knn_capb(k);
    p.cand = cand; p.out_ids = out_ids; p.out_dists = out_dists; p.counter = counter;
    const size_t smem = (size_t)KQB * KQS * 4 + 2u * KBT * KKC * 16 + (size_t)p.capb * 8 + KQB * 8;
    if (smem > 227u * 1024u) {
        set_error("knn: k too large for the shared-memory selection buffer");
        return GBDR_E_CAPACITY;
    }
    GBDR_CUDA(cudaFuncSetAttribute(knn_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    knn_scan_kernel<<<grid, KQB, smem, st>>>(p);
    GBDR_CHECK_LAUNCH();
    count_launch();
    return GBDR_OK;
}

uint32_t knn_rows_per_block() { return KQB; }

}  // namespace gbdr
