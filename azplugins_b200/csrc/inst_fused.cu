// Explicit instantiations of the fused two-potential pass for the pair
// the reference documents on one neighbour list: Colloid + Hertz (reference src/pair.py:66-76).
#include "evaluators/eval_colloid.cuh"
#include "evaluators/eval_hertz.cuh"
#include "launch.cuh"

namespace azp
    {
template cudaError_t launch_pair_fused<PairEvaluatorColloid<float>, PairEvaluatorHertz<float>, float>(const azp_pair_args*, const void*, const azp_pair_args*, const void*, cudaStream_t);
template cudaError_t launch_pair_fused<PairEvaluatorColloid<double>, PairEvaluatorHertz<double>, double>(const azp_pair_args*, const void*, const azp_pair_args*, const void*, cudaStream_t);
    } // namespace azp
