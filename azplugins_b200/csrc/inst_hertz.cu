// Explicit instantiations of the isotropic launch layer for PairEvaluatorHertz, fp32 and fp64 -- the
// counterpart of reference src/PotentialPairGPUKernel.cu.inc:25-28 (one .cu per evaluator,
// src/CMakeLists.txt:88-104).
#include "evaluators/eval_hertz.cuh"
#include "launch.cuh"

namespace azp
    {
template cudaError_t launch_pair<PairEvaluatorHertz<float>, float>(const azp_pair_args*, const void*, cudaStream_t);
template cudaError_t launch_pair<PairEvaluatorHertz<double>, double>(const azp_pair_args*, const void*, cudaStream_t);
    } // namespace azp
