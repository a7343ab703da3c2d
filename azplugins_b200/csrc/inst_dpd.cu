// Explicit instantiations for DPDPairEvaluatorGeneralWeight: the thermostatted kernel
// (reference src/PotentialPairDPDThermoGPUKernel.cu.inc:21-24) and the conservative-only
// isotropic kernel behind "PotentialPairConservativeGeneralWeight"
// (reference src/export_PotentialPairDPDThermo.cc.inc:33-35).
#include "evaluators/eval_dpd.cuh"
#include "launch.cuh"

namespace azp
    {
template cudaError_t launch_dpd<DPDPairEvaluatorGeneralWeight<float>, float>(const azp_pair_args*, const void*, cudaStream_t);
template cudaError_t launch_dpd<DPDPairEvaluatorGeneralWeight<double>, double>(const azp_pair_args*, const void*, cudaStream_t);
template cudaError_t launch_pair<DPDPairEvaluatorGeneralWeight<float>, float>(const azp_pair_args*, const void*, cudaStream_t);
template cudaError_t launch_pair<DPDPairEvaluatorGeneralWeight<double>, double>(const azp_pair_args*, const void*, cudaStream_t);
    } // namespace azp
