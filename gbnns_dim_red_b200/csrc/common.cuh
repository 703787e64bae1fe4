// common.cuh — shared device helpers for the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <string>

#include "../../include/gbdr.h"

namespace gbdr {

constexpr unsigned FULL_MASK = 0xffffffffu;
constexpr uint32_t PAD_ID = GBDR_PAD_ID;
constexpr uint32_t ID_MASK = 0x7fffffffu;   // result-list ids; MSB = "expanded" flag
constexpr uint32_t EXPANDED = 0x80000000u;

// ---- error plumbing (host) ----
void set_error(const std::string& msg);
extern std::atomic<uint64_t> g_launches;
inline void count_launch(uint64_t n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define GBDR_CUDA(expr)                                                                       \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            ::gbdr::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" +     \
                              __FILE__ + ":" + std::to_string(__LINE__) + ")");               \
            return GBDR_E_CUDA;                                                               \
        }                                                                                     \
    } while (0)

#define GBDR_CHECK_LAUNCH()                                                                   \
    do {                                                                                      \
        cudaError_t _e = cudaGetLastError();                                                  \
        if (_e != cudaSuccess) {                                                              \
            ::gbdr::set_error(std::string("kernel launch: ") + cudaGetErrorString(_e) + " (" + \
                              __FILE__ + ":" + std::to_string(__LINE__) + ")");               \
            return GBDR_E_CUDA;                                                               \
        }                                                                                     \
    } while (0)

// ---- canonical squared-L2 arithmetic ----
// The reference's L2Metric::Dist (search/support_func.h:107-128): four lane-strided partial sums,
// each `sum = sum + (a-b)*(a-b)` individually rounded, then ((T0+T1)+T2)+T3.  The _rn intrinsics
// are never contracted into FMAs, so a thread that walks a row's float4 chunks in order produces
// the bit pattern of the reference's strict-IEEE build.
struct L2Acc {
    float s0, s1, s2, s3;
    __device__ __forceinline__ L2Acc() : s0(0.f), s1(0.f), s2(0.f), s3(0.f) {}
    __device__ __forceinline__ void add(const float4& a, const float4& b) {
        float e0 = __fsub_rn(a.x, b.x), e1 = __fsub_rn(a.y, b.y), e2 = __fsub_rn(a.z, b.z),
              e3 = __fsub_rn(a.w, b.w);
        s0 = __fadd_rn(s0, __fmul_rn(e0, e0));
        s1 = __fadd_rn(s1, __fmul_rn(e1, e1));
        s2 = __fadd_rn(s2, __fmul_rn(e2, e2));
        s3 = __fadd_rn(s3, __fmul_rn(e3, e3));
    }
    __device__ __forceinline__ float result() const {
        return __fadd_rn(__fadd_rn(__fadd_rn(s0, s1), s2), s3);
    }
};

// ---- small PTX wrappers ----
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}
__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// (dist,id) lexicographic order used by every result list: std::pair<float,int> comparison of the
// reference's priority queues (search/search_function.h:50,55).
__device__ __forceinline__ bool pair_less(float da, uint32_t ia, float db, uint32_t ib) {
    return da < db || (da == db && ia < ib);
}

}  // namespace gbdr
