// beam_reglist.cuh — register-resident sorted (dist,id) list shared by the register beam-search kernels:
// lane l holds the R consecutive entries [l*R, l*R+R).
#pragma once
#include "beam_search.cuh"

namespace gbdr {

template <int R>
struct RegList {
    float d[R];
    uint32_t i[R];
};

template <int R>
__device__ __forceinline__ float list_get_d(const RegList<R>& L, int e) {
    float v = L.d[0];
#pragma unroll
    for (int r = 1; r < R; ++r)
        if ((e & (R - 1)) == r) v = L.d[r];
    return __shfl_sync(FULL_MASK, v, e / R);
}
template <int R>
__device__ __forceinline__ uint32_t list_get_i(const RegList<R>& L, int e) {
    uint32_t v = L.i[0];
#pragma unroll
    for (int r = 1; r < R; ++r)
        if ((e & (R - 1)) == r) v = L.i[r];
    return __shfl_sync(FULL_MASK, v, e / R);
}

// sorted insert of (x, xid); entries at index >= pos move up by one, the entry at CAP-1 falls off
template <int R>
__device__ __forceinline__ void list_insert_reg(RegList<R>& L, int& size, float x, uint32_t xid, int lane) {
    int pos = 0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const bool less = (lane * R + r < size) && pair_less(L.d[r], L.i[r] & ID_MASK, x, xid);
        pos += __popc(__ballot_sync(FULL_MASK, less));
    }
    const float pd = __shfl_up_sync(FULL_MASK, L.d[R - 1], 1);
    const uint32_t pi = __shfl_up_sync(FULL_MASK, L.i[R - 1], 1);
#pragma unroll
    for (int r = R - 1; r >= 0; --r) {
        const int e = lane * R + r;
        if (e > pos) {
            L.d[r] = r > 0 ? L.d[r - 1] : pd;
            L.i[r] = r > 0 ? L.i[r - 1] : pi;
        } else if (e == pos) {
            L.d[r] = x;
            L.i[r] = xid;
        }
    }
    size = size < 32 * R ? size + 1 : 32 * R;
}

}  // namespace gbdr
