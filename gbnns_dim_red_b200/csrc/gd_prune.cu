// gd_prune.cu — placeholder
#include "kernels.cuh"
namespace gbdr {
int gd_prune_device(int, const uint64_t*, const uint32_t*, const float*, uint64_t, uint32_t, uint32_t, int, int,
                    uint64_t*, uint32_t*, double*) { set_error("gd_prune not built"); return GBDR_E_STATE; }
}
