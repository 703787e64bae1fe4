// beam_search_v2_a.cu — instantiates the K2 kernel template (beam_search_v2.cuh) for lists of 32 and 64 slots.
#include "beam_search_v2.cuh"

namespace gbdr {
GBDR_V2_INSTANTIATE(launch_beam_search_v2_a, 1, 2)
}  // namespace gbdr
