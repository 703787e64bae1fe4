// beam_search_v2_e.cu — instantiates the K2 kernel template (beam_search_v2.cuh) for lists of 384 and 512 slots.
#include "beam_search_v2.cuh"

namespace gbdr {
GBDR_V2_INSTANTIATE(launch_beam_search_v2_e, 12, 16)
}  // namespace gbdr
