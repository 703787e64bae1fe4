// project.cu — K1 (CUDA-core fp32 mode): batched query projection
//   y = normalize(W3 relu(W2 relu(W1 x + b1) + b2) + b3)
// Replaces the per-query GetLowQueryFromNet / computeNetLayer / normalizeVector
// (reference search/support_func.h:624-658) with one batched GEMM per layer.  Weights are consumed in
// the reference's on-disk layout [out][in+1] (bias in the last column) without repacking.
// This is the GBDR_PROJ_FP32 mode; the default tensor-core mode lives in project_tc.cu.
#include "kernels.cuh"

namespace gbdr {

namespace {

constexpr int PT = 64;   // output tile (rows and cols)
constexpr int PK = 16;   // k tile

// Y[M x N] = act(X[M x K] * W[N x (K+1)]^T + W[:,K])
__global__ void __launch_bounds__(256) linear_kernel(const float* __restrict__ X, uint32_t ldx,
                                                     const float* __restrict__ W, uint32_t K,
                                                     float* __restrict__ Y, uint32_t ldy, uint32_t M,
                                                     uint32_t N, int relu) {
    __shared__ float xs[PK][PT + 4];
    __shared__ float ws[PK][PT + 4];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const uint32_t m0 = blockIdx.y * PT, n0 = blockIdx.x * PT;
    const uint32_t ldw = K + 1;
    float acc[4][4] = {};
    for (uint32_t k0 = 0; k0 < K; k0 += PK) {
        for (int i = threadIdx.x; i < PT * PK; i += 256) {
            const int r = i / PK, c = i % PK;
            const uint32_t gm = m0 + r, gn = n0 + r, gk = k0 + c;
            xs[c][r] = (gm < M && gk < K) ? __ldg(X + (size_t)gm * ldx + gk) : 0.f;
            ws[c][r] = (gn < N && gk < K) ? __ldg(W + (size_t)gn * ldw + gk) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < PK; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                a[i] = xs[kk][ty * 4 + i];
                b[i] = ws[kk][tx * 4 + i];
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t gm = m0 + ty * 4 + i;
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t gn = n0 + tx * 4 + j;
            if (gn >= N) continue;
            float v = acc[i][j] + __ldg(W + (size_t)gn * ldw + K);
            if (relu && v < 0.f) v = 0.f;
            Y[(size_t)gm * ldy + gn] = v;
        }
    }
}

}  // namespace

// normalizeVector (support_func.h:636-642): norm via L2Metric against a zero vector, i.e. the
// canonical 4-lane sum of squares over floor(d_low/4)*4 dims, sqrt, divide.  One thread per row.
__global__ void normalize_rows_kernel(float* __restrict__ Y, uint32_t ld, uint32_t M, uint32_t d_low) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= M) return;
    float* y = Y + (size_t)r * ld;
    L2Acc acc;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (uint32_t c = 0; c < (d_low >> 2); ++c)
        acc.add(make_float4(y[4 * c], y[4 * c + 1], y[4 * c + 2], y[4 * c + 3]), z);
    const float norm = __fsqrt_rn(acc.result());
    for (uint32_t i = 0; i < d_low; ++i) y[i] = __fdiv_rn(y[i], norm);
}

int launch_project_fp32(const float* X, uint32_t ldx, uint32_t n_q, const float* l1, const float* l2,
                        const float* l3, uint32_t d, uint32_t dh, uint32_t dh2, uint32_t d_low, float* h1,
                        float* h2, float* out, uint32_t ld_out, cudaStream_t st) {
    if (n_q == 0) return GBDR_OK;
    dim3 blk(256);
    linear_kernel<<<dim3((dh + PT - 1) / PT, (n_q + PT - 1) / PT), blk, 0, st>>>(X, ldx, l1, d, h1, dh, n_q, dh, 1);
    GBDR_CHECK_LAUNCH();
    linear_kernel<<<dim3((dh2 + PT - 1) / PT, (n_q + PT - 1) / PT), blk, 0, st>>>(h1, dh, l2, dh, h2, dh2, n_q, dh2, 1);
    GBDR_CHECK_LAUNCH();
    linear_kernel<<<dim3((d_low + PT - 1) / PT, (n_q + PT - 1) / PT), blk, 0, st>>>(h2, dh2, l3, dh2, out, ld_out, n_q, d_low, 0);
    GBDR_CHECK_LAUNCH();
    normalize_rows_kernel<<<(n_q + 127) / 128, 128, 0, st>>>(out, ld_out, n_q, d_low);
    GBDR_CHECK_LAUNCH();
    count_launch(4);
    return GBDR_OK;
}

}  // namespace gbdr
