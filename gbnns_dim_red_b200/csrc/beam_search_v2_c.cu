// beam_search_v2_c.cu — instantiates the K2 kernel template (beam_search_v2.cuh) for lists of 160 and 192 slots.
#include "beam_search_v2.cuh"

namespace gbdr {
GBDR_V2_INSTANTIATE(launch_beam_search_v2_c, 5, 6)
}  // namespace gbdr
