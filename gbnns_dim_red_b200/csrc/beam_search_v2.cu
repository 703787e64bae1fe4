// beam_search_v2.cu — host dispatch of the K2 batched-merge kernel (template in beam_search_v2.cuh; one translation
// unit per pair of list capacities, beam_search_v2_a.cu ... _e.cu, so that they compile in parallel).
#include "beam_search_v2.cuh"

namespace gbdr {

int launch_beam_search_v2_a(const BeamParams&, uint32_t, uint32_t, bool, cudaStream_t);  // 32, 64 slots
int launch_beam_search_v2_b(const BeamParams&, uint32_t, uint32_t, bool, cudaStream_t);  // 96, 128
int launch_beam_search_v2_c(const BeamParams&, uint32_t, uint32_t, bool, cudaStream_t);  // 160, 192
int launch_beam_search_v2_d(const BeamParams&, uint32_t, uint32_t, bool, cudaStream_t);  // 256, 320
int launch_beam_search_v2_e(const BeamParams&, uint32_t, uint32_t, bool, cudaStream_t);  // 384, 512

bool beam_v2_supports(uint32_t C) { return C == 4 || C == 8 || C == 12 || C == 16; }

uint32_t beam_v2_smem_per_warp(uint32_t C, uint32_t cap, uint32_t vis_bytes) { return v2_layout(C, cap, vis_bytes).total; }

int launch_beam_search_v2(const BeamParams& p, uint32_t wpb, uint32_t blocks, bool dense, cudaStream_t st) {
    switch (p.cap) {
        case 32: case 64: return launch_beam_search_v2_a(p, wpb, blocks, dense, st);
        case 96: case 128: return launch_beam_search_v2_b(p, wpb, blocks, dense, st);
        case 160: case 192: return launch_beam_search_v2_c(p, wpb, blocks, dense, st);
        case 256: case 320: return launch_beam_search_v2_d(p, wpb, blocks, dense, st);
        case 384: case 512: return launch_beam_search_v2_e(p, wpb, blocks, dense, st);
        default:
            set_error("beam_search_v2: unsupported list capacity");
            return GBDR_E_INVALID;
    }
}

}  // namespace gbdr
