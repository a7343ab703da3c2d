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
// the loop-invariant parts hoisted. ContractEvaluator<E> / ContractAnisoEvaluator<E> below adapt
// any evaluator that only follows the reference contract (cache_type = param_type + shift flag).
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

// ---------------------------------------------------------------------------------------------
// Adapters: ride the kernels with an evaluator class that implements ONLY the reference contract
// (reference src/PairEvaluator.h:67-140, src/DPDPairEvaluatorGeneralWeight.h:121-255,
// src/AnisoPairEvaluator.h:97-215) -- e.g. the reference's own evaluator headers compiled by nvcc,
// or a user's new potential. A maintainer instantiates
//     launch_pair <ContractEvaluator<MyEvaluator, Scalar>, Scalar>(args, d_params, stream)
//     launch_dpd  <ContractEvaluator<MyDPDEvaluator, Scalar>, Scalar>(...)
//     launch_aniso<ContractAnisoEvaluator<MyAnisoEvaluator, Scalar, Scalar3, Scalar4>, Scalar>(...)
// in one .cu file (csrc/launch.cuh) and gets every kernel variant. The adapter keeps the
// evaluator's own cutoff and zero-parameter tests (evalForceAndEnergy is called as is), stages
// param_type per type pair as the cache, and carries the energy_shift flag of the type pair
// (shift mode, or xplor with r_on > r_cut) next to it. tests/contract builds the reference's six
// evaluators this way and compares them with the hand-written ones (tests/test_gpu_contract.py).
// ---------------------------------------------------------------------------------------------
template<class P> struct ContractCache
    {
    P p;
    int energy_shift;
    };

template<class E, class S> class ContractEvaluator
    {
    public:
    typedef typename E::param_type param_type;
    typedef ContractCache<param_type> cache_type;
    static constexpr int evaluator_id = -1;
    AZP_HD static cache_type make_cache(const param_type& p, S, bool energy_shift)
        {
        cache_type c;
        c.p = p;
        c.energy_shift = energy_shift ? 1 : 0;
        return c;
        }
    // DPD thermostat family: dt and kT reach the evaluator through setDeltaT / setT
    AZP_HD static cache_type make_cache_thermo(const param_type& p, S rcutsq, S, S)
        {
        return make_cache(p, rcutsq, false);
        }
    AZP_D ContractEvaluator(S rsq, S rcutsq, const cache_type& c)
        : m_eval(rsq, rcutsq, c.p), m_shift(c.energy_shift != 0)
        {
        }
    AZP_D static bool needsCharge()
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
    // kernel-side entry points (see "kernel fast path" above)
    static constexpr bool kWarpVote = true;
    AZP_HD static bool disabled(const cache_type&)
        {
        return false;
        }
    AZP_D void evalPair(S& force_divr, S& pair_eng, bool)
        {
        m_eval.evalForceAndEnergy(force_divr, pair_eng, m_shift);
        }
    // DPD thermostat contract (instantiated only by launch_dpd)
    AZP_D void set_seed_ij_timestep(uint16_t seed, unsigned int i, unsigned int j, unsigned int timestep)
        {
        m_eval.set_seed_ij_timestep(seed, i, j, timestep);
        }
    AZP_D void setDeltaT(S dt)
        {
        m_eval.setDeltaT(dt);
        }
    AZP_D void setRDotV(S dot)
        {
        m_eval.setRDotV(dot);
        }
    AZP_D void setT(S T)
        {
        m_eval.setT(T);
        }
    AZP_D void evalThermoPair(S& force_divr, S& force_divr_cons, S& pair_eng, bool)
        {
        m_eval.evalForceEnergyThermo(force_divr, force_divr_cons, pair_eng, m_shift);
        }

    private:
    E m_eval;
    bool m_shift;
    };

// Anisotropic contract: ctor (Scalar3& dr, Scalar4& quat_i, Scalar4& quat_j, rcutsq, params),
// evaluate(force, pair_eng, energy_shift, torque_i, torque_j). V3 / V4 are the evaluator's own
// Scalar3 / Scalar4 aggregates ({x, y, z} / {x, y, z, w}).
template<class E, class S, class V3, class V4> class ContractAnisoEvaluator
    {
    public:
    typedef typename E::param_type param_type;
    typedef typename E::shape_type shape_type;
    typedef ContractCache<param_type> cache_type;
    static constexpr int evaluator_id = -1;
    AZP_HD static cache_type make_cache(const param_type& p, S, bool energy_shift)
        {
        cache_type c;
        c.p = p;
        c.energy_shift = energy_shift ? 1 : 0;
        return c;
        }
    AZP_D ContractAnisoEvaluator(const Vec3<S>& dr, const Vec4<S>& qi, const Vec4<S>& qj, S rcutsq, const cache_type& c)
        : m_dr {dr.x, dr.y, dr.z}, m_qi {qi.x, qi.y, qi.z, qi.w}, m_qj {qj.x, qj.y, qj.z, qj.w},
          m_rcutsq(rcutsq), m_c(c)
        {
        }
    // kernel hook (see eval_morse.cuh): a contract evaluator keeps the quaternion of particle i
    typedef Vec4<S> row_type;
    AZP_D static row_type make_row(const Vec4<S>& quat_i)
        {
        return quat_i;
        }
    struct FromRow
        {
        };
    AZP_D ContractAnisoEvaluator(FromRow, const Vec3<S>& dr, const row_type& qi, const Vec4<S>& qj, S rcutsq, const cache_type& c)
        : ContractAnisoEvaluator(dr, qi, qj, rcutsq, c)
        {
        }
    AZP_HD static bool disabled(const cache_type&)
        {
        return false;
        }
    AZP_D void evaluatePair(S, Vec3<S>& force, S& pair_eng, bool, Vec3<S>& torque_i, Vec3<S>& torque_j)
        {
        E eval(m_dr, m_qi, m_qj, m_rcutsq, m_c.p);
        V3 f {S(0), S(0), S(0)}, ti {S(0), S(0), S(0)}, tj {S(0), S(0), S(0)};
        eval.evaluate(f, pair_eng, m_c.energy_shift != 0, ti, tj);
        force = Vec3<S> {f.x, f.y, f.z};
        torque_i = Vec3<S> {ti.x, ti.y, ti.z};
        torque_j = Vec3<S> {tj.x, tj.y, tj.z};
        }

    private:
    V3 m_dr;
    V4 m_qi, m_qj;
    S m_rcutsq;
    const cache_type& m_c;
    };
    } // namespace azp
#endif
