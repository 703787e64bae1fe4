// project_tc.cu — K1 tensor-core mode (placeholder until the tcgen05 kernel lands).
#include "kernels.cuh"
namespace gbdr {
struct ProjTcPlan { int unused; };
int project_tc_prepare(const float*, const float*, const float*, uint32_t, uint32_t, uint32_t, uint32_t, cudaStream_t,
                       ProjTcPlan** out) { *out = nullptr; return GBDR_OK; }
void project_tc_destroy(ProjTcPlan* p) { delete p; }
int launch_project_tc(ProjTcPlan*, const float*, uint32_t, uint32_t, float*, uint32_t, int, cudaStream_t) {
    set_error("tensor-core projection not built"); return GBDR_E_STATE; }
}
