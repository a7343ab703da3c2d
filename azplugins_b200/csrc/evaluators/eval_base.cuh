// eval_base.cuh -- the evaluator contract of the B200 pair kernels.
//
// The reference contract (src/PairEvaluator.h:39-140, src/AnisoPairEvaluator.h:39-215) is:
//   param_type (POD per type pair), ctor (rsq, rcutsq, const param_type&),
//   static needsCharge(), setCharge(), evalForceAndEnergy(force_divr&, pair_eng&, energy_shift),
//   LRC integrals (0), getName().
// The kernels here keep that call sequence and add ONE hook that exists for the hardware:
//   cache_type + make_cache(param, rcutsq, energy_shift)
// is evaluated once per type pair when a CTA stages its shared-memory table, so everything that
// depends only on the type pair (lj1/lj2, -kappa*log2(e), 1/r_cut, the energy shift at r_cut ...)
// leaves the per-neighbour instruction stream. The per-pair arithmetic is the reference's, with
// the loop-invariant parts hoisted. ContractEvaluator<E> below adapts any evaluator that only
// follows the reference contract (cache_type = param_type).
//
// Kernel fast path. The kernels test the cutoff themselves (they need rsq < rcutsq to skip the
// whole evaluation and the accumulation), so each evaluator also exposes
//   static disabled(cache)   -- the "potential scaled to zero" clause of the reference test
//                               (e.g. epsilon == 0); staged as an effective r_cut^2 of 0 so the
//                               single compare rsq < rcutsq covers both clauses;
//   evalPair(force_divr, pair_eng, energy_shift) -- the body of evalForceAndEnergy without the
//                               test. evalForceAndEnergy == test + evalPair.
//
// "Zero on reject": like HOOMD's GPU driver, the kernels ignore the bool and rely on force_divr /
// pair_eng being left untouched (0) when a pair is rejected.
#ifndef AZP_EVAL_BASE_CUH_
#define AZP_EVAL_BASE_CUH_

#include "../azp_core.cuh"

namespace azp
    {
// Mirrors detail::PairParameters' staging hooks (no-ops: every param_type here is a POD that is
// copied into shared memory as a whole).
struct PairParametersBase
    {
    AZP_D void load_shared(char*&, unsigned int&) { }
    AZP_HD void allocate_shared(char*&, unsigned int&) const { }
    void set_memory_hint() const { }
    };

template<class S> class PairEvaluatorBase
    {
    public:
    AZP_D PairEvaluatorBase(S _rsq, S _rcutsq) : rsq(_rsq), rcutsq(_rcutsq) { }
    AZP_HD static bool needsCharge()
        {
        return false;
        }
    AZP_D void setCharge(S, S) { }
    AZP_D S evalPressureLRCIntegral()
        {
        return S(0);
        }
    AZP_D S evalEnergyLRCIntegral()
        {
        return S(0);
        }

    protected:
    S rsq;
    S rcutsq;
    };

// Adapter: ride the kernels with an evaluator that implements only the reference contract.
template<class E, class S> class ContractEvaluator
    {
    public:
    typedef typename E::param_type param_type;
    typedef typename E::param_type cache_type;
    static constexpr int evaluator_id = -1;
    AZP_HD static cache_type make_cache(const param_type& p, S, bool)
        {
        return p;
        }
    AZP_D ContractEvaluator(S rsq, S rcutsq, const cache_type& c) : m_eval(rsq, rcutsq, c) { }
    AZP_HD static bool needsCharge()
        {
        return E::needsCharge();
        }
    AZP_D void setCharge(S qi, S qj)
        {
        m_eval.setCharge(qi, qj);
        }
    AZP_D bool evalForceAndEnergy(S& force_divr, S& pair_eng, bool energy_shift)
        {
        return m_eval.evalForceAndEnergy(force_divr, pair_eng, energy_shift);
        }
    // kernel-side entry points (see "kernel fast path" below): a contract-only evaluator keeps
    // its own cutoff / zero-parameter tests
    static constexpr bool kWarpVote = true;
    AZP_HD static bool disabled(const cache_type&)
        {
        return false;
        }
    AZP_D void evalPair(S& force_divr, S& pair_eng, bool energy_shift)
        {
        m_eval.evalForceAndEnergy(force_divr, pair_eng, energy_shift);
        }

    private:
    E m_eval;
    };
    } // namespace azp
#endif
