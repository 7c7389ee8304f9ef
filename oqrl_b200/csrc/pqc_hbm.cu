// placeholder until the HBM family lands
#include "common.cuh"
namespace oqrl {
int hbm_qvalues(const oqrl_spec &sp, const float *, int, const float *, int, float *, float *, void *, size_t, int, cudaStream_t)
{ set_error("n_qubits=%d: HBM family not built yet", sp.n_qubits); return OQRL_ERR_UNSUPPORTED; }
size_t hbm_workspace_bytes(const oqrl_spec &, int) { return 0; }
double hbm_flops_per_eval(const oqrl_spec &) { return 0; }
double hbm_sweeps_per_eval(const oqrl_spec &) { return 0; }
}
