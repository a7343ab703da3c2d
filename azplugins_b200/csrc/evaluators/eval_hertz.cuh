// eval_hertz.cuh -- Hertz soft repulsion U = eps (1 - r/r_cut)^(5/2).
// Behaviour: reference src/PairEvaluatorHertz.h:23-46 (param_type), :93-109. The energy shift
// flag is ignored there as well: the potential is already zero at r_cut.
#ifndef AZP_EVAL_HERTZ_CUH_
#define AZP_EVAL_HERTZ_CUH_

#include "eval_base.cuh"

namespace azp
    {
template<class S> class PairEvaluatorHertz : public PairEvaluatorBase<S>
    {
    public:
    static constexpr int evaluator_id = 3;
    struct alignas(sizeof(S)) param_type : public PairParametersBase
        {
        S epsilon;
        };

    struct alignas(16) cache_type
        {
        S epsilon;
        S rcut_inv; // 1 / sqrt(rcutsq), hoisted per type pair
        };

    AZP_HD static cache_type make_cache(const param_type& p, S rcutsq, bool)
        {
        cache_type c;
        c.epsilon = p.epsilon;
        c.rcut_inv = S(1.0) / ::sqrt(rcutsq);
        return c;
        }

    AZP_D PairEvaluatorHertz(S _rsq, S _rcutsq, const cache_type& _c)
        : PairEvaluatorBase<S>(_rsq, _rcutsq), c(_c)
        {
        }

    // see IsoFamily::pair: skip the evaluator when no lane of the warp is inside the cutoff
    static constexpr bool kWarpVote = true;

    AZP_HD static bool disabled(const cache_type& c)
        {
        return c.epsilon == S(0);
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
            const S rinv = fast::rsqrt(this->rsq);
            const S r = this->rsq * rinv;
            // r < r_cut was tested on the squares; guard the last-ulp case before the sqrt
            const S x = fmax(S(1.0) - r * c.rcut_inv, S(0));
            const S e32 = c.epsilon * x * fast::sqrt(x);
            force_divr = S(2.5) * e32 * rinv * c.rcut_inv;
            pair_eng = e32 * x;
            }
        }

    static const char* getName()
        {
        return "hertz";
        }
    static void pack(const double* f, param_type* p)
        {
        p->epsilon = S(f[0]);
        }
    static void unpack(const param_type* p, double* f)
        {
        f[0] = double(p->epsilon);
        }
    static constexpr int num_fields = 1;

    private:
    const cache_type& c;
    };
    } // namespace azp
#endif
