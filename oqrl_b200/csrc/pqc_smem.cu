// placeholder until the shared-memory family lands
#include "common.cuh"
namespace oqrl {
int smem_fitness(const oqrl_spec &sp, const float *, int, const TransitionView &, float, float *, double *, int, cudaStream_t)
{ set_error("n_qubits=%d: shared-memory family not built yet", sp.n_qubits); return OQRL_ERR_UNSUPPORTED; }
int smem_qvalues(const oqrl_spec &sp, const float *, int, const float *, int, float *, cudaStream_t)
{ set_error("n_qubits=%d: shared-memory family not built yet", sp.n_qubits); return OQRL_ERR_UNSUPPORTED; }
double smem_flops_per_eval(const oqrl_spec &) { return 0; }
}
