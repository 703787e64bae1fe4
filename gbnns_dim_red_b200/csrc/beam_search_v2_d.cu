// beam_search_v2_d.cu — instantiates the K2 kernel template (beam_search_v2.cuh) for lists of 256 and 320 slots.
#include "beam_search_v2.cuh"

namespace gbdr {
GBDR_V2_INSTANTIATE(launch_beam_search_v2_d, 8, 10)
}  // namespace gbdr
