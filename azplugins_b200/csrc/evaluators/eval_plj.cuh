// eval_plj.cuh -- perturbed Lennard-Jones (WCA core + lambda-scaled attractive tail).
// Behaviour: reference src/PairEvaluatorPerturbedLennardJones.h:25-66 (param_type),
// :96-104 (derived constants), :117-155 (force/energy/shift).
#ifndef AZP_EVAL_PLJ_CUH_
#define AZP_EVAL_PLJ_CUH_

#include "eval_base.cuh"
#include <cmath>

namespace azp
    {
template<class S> class PairEvaluatorPerturbedLennardJones : public PairEvaluatorBase<S>
    {
    public:
    static constexpr int evaluator_id = 0;
    // same field order, size and alignment as the reference struct (16 B fp32 / 32 B fp64)
    struct alignas(4 * sizeof(S)) param_type : public PairParametersBase
        {
        S sigma_6;
        S epsilon_x_4;
        S attraction_scale_factor;
        S rwcasq;
        };

    struct alignas(16) cache_type
        {
        S lj1;    // 4 eps sigma^12
        S lj2;    // 4 eps sigma^6
        S lambda;
        S rwcasq;
        S wca_shift; // eps (1 - lambda)
        S e_cut;     // energy at r_cut (0 unless shifting)
        };

    AZP_HD static cache_type make_cache(const param_type& p, S rcutsq, bool energy_shift)
        {
        cache_type c;
        c.lj1 = p.epsilon_x_4 * p.sigma_6 * p.sigma_6;
        c.lj2 = p.epsilon_x_4 * p.sigma_6;
        c.lambda = p.attraction_scale_factor;
        c.rwcasq = p.rwcasq;
        c.wca_shift = p.epsilon_x_4 * (S(1.0) - c.lambda) / S(4.0);
        c.e_cut = S(0);
        if (energy_shift)
            {
            const S c2 = S(1.0) / rcutsq;
            const S c6 = c2 * c2 * c2;
            S e = c6 * (c.lj1 * c6 - c.lj2);
            if (rcutsq < c.rwcasq)
                e += c.wca_shift;
            else
                e *= c.lambda;
            c.e_cut = e;
            }
        return c;
        }

    AZP_D PairEvaluatorPerturbedLennardJones(S _rsq, S _rcutsq, const cache_type& _c)
        : PairEvaluatorBase<S>(_rsq, _rcutsq), c(_c)
        {
        }

    // see IsoFamily::pair: skip the evaluator when no lane of the warp is inside the cutoff
    static constexpr bool kWarpVote = false;

    AZP_HD static bool disabled(const cache_type& c)
        {
        return c.lj1 == S(0);
        }

    AZP_D bool evalForceAndEnergy(S& force_divr, S& pair_eng, bool energy_shift)
        {
        if (this->rsq < this->rcutsq && !disabled(c))
            {
            evalPair(force_divr, pair_eng, energy_shift);
            return true;
            }
        return false;
        }

    AZP_D void evalPair(S& force_divr, S& pair_eng, bool)
        {
            {
            const S r2inv = fast::rcp(this->rsq);
            const S r6inv = r2inv * r2inv * r2inv;
            S f = r2inv * r6inv * (S(12.0) * c.lj1 * r6inv - S(6.0) * c.lj2);
            S e = r6inv * (c.lj1 * r6inv - c.lj2);
            if (this->rsq < c.rwcasq)
                e += c.wca_shift;
            else
                {
                f *= c.lambda;
                e *= c.lambda;
                }
            force_divr = f;
            pair_eng = e - c.e_cut;
            }
        }

    static const char* getName()
        {
        return "PerturbedLennardJones";
        }

    // host: fields {epsilon, sigma, attraction_scale_factor} -> param_type with the reference
    // constructor's roundings (narrow first, derive in Scalar; 2^(1/3) sigma^2 formed in double)
    static void pack(const double* f, param_type* p)
        {
        const S eps = S(f[0]), sig = S(f[1]);
        const S s2 = sig * sig;
        const S s4 = s2 * s2;
        p->sigma_6 = s2 * s4;
        p->epsilon_x_4 = S(4.0) * eps;
        p->attraction_scale_factor = S(f[2]);
        p->rwcasq = S(std::pow(2.0, 1.0 / 3.0) * double(s2));
        }
    static void unpack(const param_type* p, double* f)
        {
        f[0] = double(p->epsilon_x_4 / S(4.));
        f[1] = std::pow(double(p->sigma_6), 1. / 6.);
        f[2] = double(p->attraction_scale_factor);
        }
    static constexpr int num_fields = 3;

    private:
    const cache_type& c;
    };
    } // namespace azp
#endif
