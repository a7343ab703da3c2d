// TEST INFRASTRUCTURE. Not part of the product path.
//
// The reference's OWN evaluator headers (/root/reference/src/PairEvaluator*.h,
// DPDPairEvaluatorGeneralWeight.h, AnisoPairEvaluatorTwoPatchMorse.h), compiled in place by nvcc
// for sm_100a (-D__HIPCC__ -DNVCC, oracle/hoomd_stub for the three HOOMD headers they include),
// riding the PRODUCT's kernels (csrc/pair_kernels.cuh, launch.cuh) through the contract adapters
// of csrc/evaluators/eval_base.cuh. This is what a HOOMD build does with the reference's
// *.cu.inc stubs (src/PotentialPairGPUKernel.cu.inc:25-28 etc.): instantiate the driver template
// per evaluator class. Built by oracle/Makefile (target `contract`) into oracle/_ref/, only where
// the reference tree is present; tests/test_gpu_contract.py compares it with the hand-written
// evaluators and with the CPU oracle.
#include "hoomd/HOOMDMath.h"
#include "hoomd/RandomNumbers.h"
#include "hoomd/VectorMath.h"

#include "AnisoPairEvaluatorTwoPatchMorse.h"
#include "DPDPairEvaluatorGeneralWeight.h"
#include "PairEvaluatorColloid.h"
#include "PairEvaluatorExpandedYukawa.h"
#include "PairEvaluatorHertz.h"
#include "PairEvaluatorPerturbedLennardJones.h"

#include "evaluators/eval_base.cuh"
#include "launch.cuh"

namespace azp
    {
// the long-row scratch of the launch layer (the product defines it in capi.cu)
cudaError_t long_row_scratch(cudaStream_t stream, LongRowScratch& out)
    {
    static LongRowScratch s;
    if (!s.queue)
        {
        unsigned int* mem = nullptr;
        cudaError_t err = cudaMalloc(reinterpret_cast<void**>(&mem), sizeof(unsigned int) * (kLongRowCapacity + 4));
        if (err != cudaSuccess)
            return err;
        s.count = mem;
        s.queue = mem + 4;
        }
    out = s;
    return cudaSuccess;
    }
    } // namespace azp

using hoomd::Scalar;
namespace azref = hoomd::azplugins::detail;

template<class E> using Iso = azp::ContractEvaluator<E, Scalar>;
using Aniso = azp::ContractAnisoEvaluator<azref::AnisoPairEvaluatorTwoPatchMorse, Scalar, hoomd::Scalar3, hoomd::Scalar4>;

extern "C"
    {
    int contract_scalar_bits(void)
        {
        return 8 * (int)sizeof(Scalar);
        }
    int contract_param_size(int ev)
        {
        switch (ev)
            {
        case AZP_EV_PERTURBED_LENNARD_JONES:
            return (int)sizeof(azref::PairEvaluatorPerturbedLennardJones::param_type);
        case AZP_EV_EXPANDED_YUKAWA:
            return (int)sizeof(azref::PairEvaluatorExpandedYukawa::param_type);
        case AZP_EV_COLLOID:
            return (int)sizeof(azref::PairEvaluatorColloid::param_type);
        case AZP_EV_HERTZ:
            return (int)sizeof(azref::PairEvaluatorHertz::param_type);
        case AZP_EV_DPD_GENERAL_WEIGHT:
            return (int)sizeof(azref::DPDPairEvaluatorGeneralWeight::param_type);
        case AZP_EV_TWO_PATCH_MORSE:
            return (int)sizeof(azref::AnisoPairEvaluatorTwoPatchMorse::param_type);
        default:
            return -1;
            }
        }
    // family: 0 gpu_compute_pair_forces, 1 gpu_compute_dpd_forces, 2 gpu_compute_pair_aniso_forces
    int contract_forces(int family, int ev, const azp_pair_args* a, const void* d_params, void* stream)
        {
        cudaStream_t st = (cudaStream_t)stream;
        if (family == 0)
            {
            switch (ev)
                {
            case AZP_EV_PERTURBED_LENNARD_JONES:
                return (int)azp::launch_pair<Iso<azref::PairEvaluatorPerturbedLennardJones>, Scalar>(a, d_params, st);
            case AZP_EV_EXPANDED_YUKAWA:
                return (int)azp::launch_pair<Iso<azref::PairEvaluatorExpandedYukawa>, Scalar>(a, d_params, st);
            case AZP_EV_COLLOID:
                return (int)azp::launch_pair<Iso<azref::PairEvaluatorColloid>, Scalar>(a, d_params, st);
            case AZP_EV_HERTZ:
                return (int)azp::launch_pair<Iso<azref::PairEvaluatorHertz>, Scalar>(a, d_params, st);
            case AZP_EV_DPD_GENERAL_WEIGHT:
                return (int)azp::launch_pair<Iso<azref::DPDPairEvaluatorGeneralWeight>, Scalar>(a, d_params, st);
            default:
                return (int)cudaErrorInvalidValue;
                }
            }
        if (family == 1 && ev == AZP_EV_DPD_GENERAL_WEIGHT)
            return (int)azp::launch_dpd<Iso<azref::DPDPairEvaluatorGeneralWeight>, Scalar>(a, d_params, st);
        if (family == 2 && ev == AZP_EV_TWO_PATCH_MORSE)
            return (int)azp::launch_aniso<Aniso, Scalar>(a, d_params, st);
        return (int)cudaErrorInvalidValue;
        }
    }
