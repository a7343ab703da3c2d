// capi.cu -- extern "C" boundary of libazp_b200.so (declared in include/azp_b200.h).
//
// Thin dispatch from (evaluator id, precision) to the explicitly instantiated launch layers in
// inst_*.cu, plus host-side param_type packing and the launch autotuner. Never throws; every
// entry point returns a cudaError_t as int.
#include "../../include/azp_b200.h"

#include "evaluators/eval_colloid.cuh"
#include "evaluators/eval_dpd.cuh"
#include "evaluators/eval_hertz.cuh"
#include "evaluators/eval_morse.cuh"
#include "evaluators/eval_plj.cuh"
#include "evaluators/eval_yukawa.cuh"
#include "launch.cuh"

#include <cstring>

namespace azp
    {
// instantiated in inst_*.cu
#define AZP_EXTERN_PAIR(E)                                                                         \
    extern template cudaError_t launch_pair<E<float>, float>(const azp_pair_args*, const void*,    \
                                                             cudaStream_t);                        \
    extern template cudaError_t launch_pair<E<double>, double>(const azp_pair_args*, const void*,  \
                                                               cudaStream_t);
AZP_EXTERN_PAIR(PairEvaluatorPerturbedLennardJones)
AZP_EXTERN_PAIR(PairEvaluatorExpandedYukawa)
AZP_EXTERN_PAIR(PairEvaluatorColloid)
AZP_EXTERN_PAIR(PairEvaluatorHertz)
AZP_EXTERN_PAIR(DPDPairEvaluatorGeneralWeight)
#undef AZP_EXTERN_PAIR
extern template cudaError_t launch_dpd<DPDPairEvaluatorGeneralWeight<float>, float>(const azp_pair_args*, const void*, cudaStream_t);
extern template cudaError_t launch_dpd<DPDPairEvaluatorGeneralWeight<double>, double>(const azp_pair_args*, const void*, cudaStream_t);
extern template cudaError_t launch_aniso<AnisoPairEvaluatorTwoPatchMorse<float>, float>(const azp_pair_args*, const void*, cudaStream_t);
extern template cudaError_t launch_aniso<AnisoPairEvaluatorTwoPatchMorse<double>, double>(const azp_pair_args*, const void*, cudaStream_t);

extern template cudaError_t launch_pair_fused<PairEvaluatorColloid<float>, PairEvaluatorHertz<float>, float>(const azp_pair_args*, const void*, const azp_pair_args*, const void*, cudaStream_t);
extern template cudaError_t launch_pair_fused<PairEvaluatorColloid<double>, PairEvaluatorHertz<double>, double>(const azp_pair_args*, const void*, const azp_pair_args*, const void*, cudaStream_t);

template<class S> static cudaError_t dispatch_pair(int ev, const azp_pair_args* a, const void* p, cudaStream_t st)
    {
    switch (ev)
        {
    case AZP_EV_PERTURBED_LENNARD_JONES:
        return launch_pair<PairEvaluatorPerturbedLennardJones<S>, S>(a, p, st);
    case AZP_EV_EXPANDED_YUKAWA:
        return launch_pair<PairEvaluatorExpandedYukawa<S>, S>(a, p, st);
    case AZP_EV_COLLOID:
        return launch_pair<PairEvaluatorColloid<S>, S>(a, p, st);
    case AZP_EV_HERTZ:
        return launch_pair<PairEvaluatorHertz<S>, S>(a, p, st);
    case AZP_EV_DPD_GENERAL_WEIGHT:
        return launch_pair<DPDPairEvaluatorGeneralWeight<S>, S>(a, p, st);
    default:
        return cudaErrorInvalidValue;
        }
    }

template<class S> static int param_size(int ev)
    {
    switch (ev)
        {
    case AZP_EV_PERTURBED_LENNARD_JONES:
        return (int)sizeof(typename PairEvaluatorPerturbedLennardJones<S>::param_type);
    case AZP_EV_EXPANDED_YUKAWA:
        return (int)sizeof(typename PairEvaluatorExpandedYukawa<S>::param_type);
    case AZP_EV_COLLOID:
        return (int)sizeof(typename PairEvaluatorColloid<S>::param_type);
    case AZP_EV_HERTZ:
        return (int)sizeof(typename PairEvaluatorHertz<S>::param_type);
    case AZP_EV_DPD_GENERAL_WEIGHT:
        return (int)sizeof(typename DPDPairEvaluatorGeneralWeight<S>::param_type);
    case AZP_EV_TWO_PATCH_MORSE:
        return (int)sizeof(typename AnisoPairEvaluatorTwoPatchMorse<S>::param_type);
    default:
        return -1;
        }
    }

template<class E> static void pack_one(const double* f, void* out)
    {
    typename E::param_type p;
    std::memset(&p, 0, sizeof(p)); // deterministic padding bytes
    E::pack(f, &p);
    std::memcpy(out, &p, sizeof(p));
    }
template<class E> static void unpack_one(const void* in, double* f)
    {
    typename E::param_type p;
    std::memcpy(&p, in, sizeof(p));
    E::unpack(&p, f);
    }

template<class S> static int pack(int ev, const double* f, void* out)
    {
    switch (ev)
        {
    case AZP_EV_PERTURBED_LENNARD_JONES:
        pack_one<PairEvaluatorPerturbedLennardJones<S>>(f, out);
        return 0;
    case AZP_EV_EXPANDED_YUKAWA:
        pack_one<PairEvaluatorExpandedYukawa<S>>(f, out);
        return 0;
    case AZP_EV_COLLOID:
        pack_one<PairEvaluatorColloid<S>>(f, out);
        return 0;
    case AZP_EV_HERTZ:
        pack_one<PairEvaluatorHertz<S>>(f, out);
        return 0;
    case AZP_EV_DPD_GENERAL_WEIGHT:
        pack_one<DPDPairEvaluatorGeneralWeight<S>>(f, out);
        return 0;
    case AZP_EV_TWO_PATCH_MORSE:
        pack_one<AnisoPairEvaluatorTwoPatchMorse<S>>(f, out);
        return 0;
    default:
        return 1;
        }
    }
template<class S> static int unpack(int ev, const void* in, double* f)
    {
    switch (ev)
        {
    case AZP_EV_PERTURBED_LENNARD_JONES:
        unpack_one<PairEvaluatorPerturbedLennardJones<S>>(in, f);
        return 0;
    case AZP_EV_EXPANDED_YUKAWA:
        unpack_one<PairEvaluatorExpandedYukawa<S>>(in, f);
        return 0;
    case AZP_EV_COLLOID:
        unpack_one<PairEvaluatorColloid<S>>(in, f);
        return 0;
    case AZP_EV_HERTZ:
        unpack_one<PairEvaluatorHertz<S>>(in, f);
        return 0;
    case AZP_EV_DPD_GENERAL_WEIGHT:
        unpack_one<DPDPairEvaluatorGeneralWeight<S>>(in, f);
        return 0;
    case AZP_EV_TWO_PATCH_MORSE:
        unpack_one<AnisoPairEvaluatorTwoPatchMorse<S>>(in, f);
        return 0;
    default:
        return 1;
        }
    }
    } // namespace azp

namespace azp
    {
cudaError_t long_row_scratch(cudaStream_t stream, LongRowScratch& out)
    {
    static std::mutex lock;
    static std::map<std::pair<int, cudaStream_t>, LongRowScratch> table;
    int dev = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err != cudaSuccess)
        return err;
    std::lock_guard<std::mutex> guard(lock);
    LongRowScratch& s = table[std::make_pair(dev, stream)];
    if (!s.queue)
        {
        unsigned int* mem = nullptr;
        err = cudaMalloc(reinterpret_cast<void**>(&mem), sizeof(unsigned int) * (kLongRowCapacity + 4));
        if (err != cudaSuccess)
            return err;
        s.count = mem;
        s.queue = mem + 4;
        }
    out = s;
    return cudaSuccess;
    }
    } // namespace azp

using namespace azp;

static int run_family(int family, int ev, int bits, const azp_pair_args* a, const void* p, cudaStream_t st)
    {
    if (bits == 32)
        {
        if (family == 0)
            return (int)dispatch_pair<float>(ev, a, p, st);
        if (family == 1)
            return ev == AZP_EV_DPD_GENERAL_WEIGHT ? (int)launch_dpd<DPDPairEvaluatorGeneralWeight<float>, float>(a, p, st) : (int)cudaErrorInvalidValue;
        if (family == 2)
            return ev == AZP_EV_TWO_PATCH_MORSE ? (int)launch_aniso<AnisoPairEvaluatorTwoPatchMorse<float>, float>(a, p, st) : (int)cudaErrorInvalidValue;
        }
    else if (bits == 64)
        {
        if (family == 0)
            return (int)dispatch_pair<double>(ev, a, p, st);
        if (family == 1)
            return ev == AZP_EV_DPD_GENERAL_WEIGHT ? (int)launch_dpd<DPDPairEvaluatorGeneralWeight<double>, double>(a, p, st) : (int)cudaErrorInvalidValue;
        if (family == 2)
            return ev == AZP_EV_TWO_PATCH_MORSE ? (int)launch_aniso<AnisoPairEvaluatorTwoPatchMorse<double>, double>(a, p, st) : (int)cudaErrorInvalidValue;
        }
    return (int)cudaErrorInvalidValue;
    }

extern "C"
    {
    int azp_abi_version(void)
        {
        return AZP_B200_ABI_VERSION;
        }

    const char* azp_error_string(int code)
        {
        return cudaGetErrorString((cudaError_t)code);
        }

    const char* azp_evaluator_name(int ev)
        {
        switch (ev)
            {
        case AZP_EV_PERTURBED_LENNARD_JONES:
            return PairEvaluatorPerturbedLennardJones<float>::getName();
        case AZP_EV_EXPANDED_YUKAWA:
            return PairEvaluatorExpandedYukawa<float>::getName();
        case AZP_EV_COLLOID:
            return PairEvaluatorColloid<float>::getName();
        case AZP_EV_HERTZ:
            return PairEvaluatorHertz<float>::getName();
        case AZP_EV_DPD_GENERAL_WEIGHT:
            return DPDPairEvaluatorGeneralWeight<float>::getName();
        case AZP_EV_TWO_PATCH_MORSE:
            return AnisoPairEvaluatorTwoPatchMorse<float>::getName();
        default:
            return "";
            }
        }

    int azp_param_num_fields(int ev)
        {
        switch (ev)
            {
        case AZP_EV_PERTURBED_LENNARD_JONES:
            return PairEvaluatorPerturbedLennardJones<float>::num_fields;
        case AZP_EV_EXPANDED_YUKAWA:
            return PairEvaluatorExpandedYukawa<float>::num_fields;
        case AZP_EV_COLLOID:
            return PairEvaluatorColloid<float>::num_fields;
        case AZP_EV_HERTZ:
            return PairEvaluatorHertz<float>::num_fields;
        case AZP_EV_DPD_GENERAL_WEIGHT:
            return DPDPairEvaluatorGeneralWeight<float>::num_fields;
        case AZP_EV_TWO_PATCH_MORSE:
            return AnisoPairEvaluatorTwoPatchMorse<float>::num_fields;
        default:
            return -1;
            }
        }

    int azp_param_size(int ev, int bits)
        {
        if (bits == 32)
            return param_size<float>(ev);
        if (bits == 64)
            return param_size<double>(ev);
        return -1;
        }

    int azp_param_pack(int ev, int bits, const double* fields, void* out)
        {
        if (!fields || !out)
            return (int)cudaErrorInvalidValue;
        int rc = 1;
        if (bits == 32)
            rc = pack<float>(ev, fields, out);
        else if (bits == 64)
            rc = pack<double>(ev, fields, out);
        return rc == 0 ? 0 : (int)cudaErrorInvalidValue;
        }

    int azp_param_unpack(int ev, int bits, const void* in, double* fields)
        {
        if (!fields || !in)
            return (int)cudaErrorInvalidValue;
        int rc = 1;
        if (bits == 32)
            rc = unpack<float>(ev, in, fields);
        else if (bits == 64)
            rc = unpack<double>(ev, in, fields);
        return rc == 0 ? 0 : (int)cudaErrorInvalidValue;
        }

    int azp_pair_forces_f32(int ev, const azp_pair_args* a, const void* p, void* stream)
        {
        return run_family(0, ev, 32, a, p, (cudaStream_t)stream);
        }
    int azp_pair_forces_f64(int ev, const azp_pair_args* a, const void* p, void* stream)
        {
        return run_family(0, ev, 64, a, p, (cudaStream_t)stream);
        }
    int azp_pair_forces_fused_f32(int ev_a, const azp_pair_args* a, const void* pa, int ev_b, const azp_pair_args* b, const void* pb, void* stream)
        {
        if (ev_a == AZP_EV_COLLOID && ev_b == AZP_EV_HERTZ)
            return (int)launch_pair_fused<PairEvaluatorColloid<float>, PairEvaluatorHertz<float>, float>(a, pa, b, pb, (cudaStream_t)stream);
        return (int)cudaErrorNotSupported;
        }
    int azp_pair_forces_fused_f64(int ev_a, const azp_pair_args* a, const void* pa, int ev_b, const azp_pair_args* b, const void* pb, void* stream)
        {
        if (ev_a == AZP_EV_COLLOID && ev_b == AZP_EV_HERTZ)
            return (int)launch_pair_fused<PairEvaluatorColloid<double>, PairEvaluatorHertz<double>, double>(a, pa, b, pb, (cudaStream_t)stream);
        return (int)cudaErrorNotSupported;
        }
    int azp_dpd_forces_f32(int ev, const azp_pair_args* a, const void* p, void* stream)
        {
        return run_family(1, ev, 32, a, p, (cudaStream_t)stream);
        }
    int azp_dpd_forces_f64(int ev, const azp_pair_args* a, const void* p, void* stream)
        {
        return run_family(1, ev, 64, a, p, (cudaStream_t)stream);
        }
    // d_shape_params: TwoPatchMorse's shape_type is empty (reference
    // src/AnisoPairEvaluator.h:65-85), so the pointer is accepted and ignored.
    int azp_aniso_forces_f32(int ev, const azp_pair_args* a, const void* p, const void*, void* stream)
        {
        return run_family(2, ev, 32, a, p, (cudaStream_t)stream);
        }
    int azp_aniso_forces_f64(int ev, const azp_pair_args* a, const void* p, const void*, void* stream)
        {
        return run_family(2, ev, 64, a, p, (cudaStream_t)stream);
        }

    int azp_autotune(int family, int ev, int bits, const azp_pair_args* a, const void* p, void* stream, uint32_t* best_block, uint32_t* best_tpp, float* best_ms)
        {
        if (!a || !best_block || !best_tpp)
            return (int)cudaErrorInvalidValue;
        cudaStream_t st = (cudaStream_t)stream;
        cudaEvent_t e0, e1;
        cudaError_t err = cudaEventCreate(&e0);
        if (err != cudaSuccess)
            return (int)err;
        err = cudaEventCreate(&e1);
        if (err != cudaSuccess)
            {
            cudaEventDestroy(e0);
            return (int)err;
            }
        // multiples of 32 that divide the register file differently (e.g. 160 threads x 79
        // registers keeps 25 warps per SM resident where 128 keeps 24)
        const uint32_t blocks[] = {64, 96, 128, 160, 192, 256, 512};
        float best = -1.0f;
        uint32_t bb = 128, bt = 8;
        int rc = 0;
        for (uint32_t bi = 0; bi < sizeof(blocks) / sizeof(blocks[0]) && rc == 0; ++bi)
            for (uint32_t tpp = 1; tpp <= 32 && rc == 0; tpp <<= 1)
                {
                azp_pair_args t = *a;
                t.block_size = blocks[bi];
                t.threads_per_particle = tpp;
                if (bits == 64 && blocks[bi] > 256)
                    continue; // fp64 kernels are built for blocks of at most 256 threads
                rc = run_family(family, ev, bits, &t, p, st); // warm-up
                if (rc != 0)
                    break;
                const int reps = 3;
                cudaEventRecord(e0, st);
                for (int r = 0; r < reps && rc == 0; ++r)
                    rc = run_family(family, ev, bits, &t, p, st);
                cudaEventRecord(e1, st);
                err = cudaEventSynchronize(e1);
                if (err != cudaSuccess)
                    rc = (int)err;
                float ms = 0;
                cudaEventElapsedTime(&ms, e0, e1);
                ms /= reps;
                if (rc == 0 && (best < 0 || ms < best))
                    {
                    best = ms;
                    bb = blocks[bi];
                    bt = tpp;
                    }
                }
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        *best_block = bb;
        *best_tpp = bt;
        if (best_ms)
            *best_ms = best;
        return rc;
        }

    double azp_dpd_alpha(int bits, uint32_t seed, uint32_t tag_i, uint32_t tag_j, uint64_t timestep)
        {
        const uint32_t ts = (uint32_t)(timestep & 0xffffffffull);
        if (bits == 32)
            return (double)dpd_uniform_pm1<float>(seed, tag_i, tag_j, ts);
        return dpd_uniform_pm1<double>(seed, tag_i, tag_j, ts);
        }

    void azp_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
        {
        const Philox4 u = philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1]);
        for (int i = 0; i < 4; ++i)
            out[i] = u.v[i];
        }
    } // extern "C"
