// Explicit instantiations for AnisoPairEvaluatorTwoPatchMorse -- the counterpart of reference
// src/AnisoPotentialPairGPUKernel.cu.inc:21-25.
#include "evaluators/eval_morse.cuh"
#include "launch.cuh"

namespace azp
    {
template cudaError_t launch_aniso<AnisoPairEvaluatorTwoPatchMorse<float>, float>(const azp_pair_args*, const void*, cudaStream_t);
template cudaError_t launch_aniso<AnisoPairEvaluatorTwoPatchMorse<double>, double>(const azp_pair_args*, const void*, cudaStream_t);
    } // namespace azp
