// beam_search_v2_b.cu — instantiates the K2 kernel template (beam_search_v2.cuh) for lists of 96 and 128 slots.
#include "beam_search_v2.cuh"

namespace gbdr {
GBDR_V2_INSTANTIATE(launch_beam_search_v2_b, 3, 4)
}  // namespace gbdr
