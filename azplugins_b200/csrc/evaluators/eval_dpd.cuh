// eval_dpd.cuh -- DPD with the generalised dissipative/random weight w(r) = (1 - r/r_cut)^s.
// Behaviour: reference src/DPDPairEvaluatorGeneralWeight.h:32-62 (param_type), :121-155
// (thermostat inputs), :165-186 (conservative part), :198-255 (thermostatted force);
// RNG keying in azp_philox.cuh.
//
// Hoisted per type pair: 1/r_cut, r_cut, s/2 and the random-force amplitude
// rsqrt(dt / (6 gamma kT)) (0 when gamma*kT == 0, which the kT=0 reference tests rely on).
// Deviation, documented in DESIGN.md: the weight base (1 - r/r_cut) is clamped at 0. In exact
// arithmetic it is positive for every accepted pair; the clamp only removes the NaN the
// reference produces when r*(1/r_cut) rounds above 1 in the last ulp below the cutoff.
#ifndef AZP_EVAL_DPD_CUH_
#define AZP_EVAL_DPD_CUH_

#include "../azp_philox.cuh"
#include "eval_base.cuh"

namespace azp
    {
template<class S> class DPDPairEvaluatorGeneralWeight : public PairEvaluatorBase<S>
    {
    public:
    static constexpr int evaluator_id = 4;
    struct alignas(4 * sizeof(S)) param_type : public PairParametersBase
        {
        S A;
        S gamma;
        S s;
        };

    struct alignas(16) cache_type
        {
        S A;
        S gamma;
        S half_s;
        S rcut;
        S rcut_inv;
        S noise; // rsqrt(dt / (kT gamma 6))
        };

    // conservative-only use (PotentialPairConservativeGeneralWeight)
    AZP_HD static cache_type make_cache(const param_type& p, S rcutsq, bool)
        {
        return make_cache_thermo(p, rcutsq, S(0), S(0));
        }
    AZP_HD static cache_type make_cache_thermo(const param_type& p, S rcutsq, S deltaT, S T)
        {
        cache_type c;
        c.A = p.A;
        c.gamma = p.gamma;
        c.half_s = S(0.5) * p.s;
        c.rcut_inv = S(1.0) / ::sqrt(rcutsq);
        c.rcut = S(1.0) / c.rcut_inv;
        c.noise = S(1.0) / ::sqrt(deltaT / (T * p.gamma * S(6.0)));
        return c;
        }

    AZP_D DPDPairEvaluatorGeneralWeight(S _rsq, S _rcutsq, const cache_type& _c)
        : PairEvaluatorBase<S>(_rsq, _rcutsq), c(_c), m_seed(0), m_i(0), m_j(0), m_timestep(0),
          m_dot(0)
        {
        }

    AZP_D void set_seed_ij_timestep(uint16_t seed, unsigned int i, unsigned int j, unsigned int timestep)
        {
        m_seed = seed;
        m_i = i;
        m_j = j;
        m_timestep = timestep;
        }
    // dt and kT enter through make_cache_thermo; kept for contract compatibility
    AZP_D void setDeltaT(S) { }
    AZP_D void setT(S) { }
    AZP_D void setRDotV(S dot)
        {
        m_dot = dot;
        }

    // see IsoFamily::pair: skip the evaluator when no lane of the warp is inside the cutoff
    static constexpr bool kWarpVote = false;

    AZP_HD static bool disabled(const cache_type&)
        {
        return false;
        }

    AZP_D bool evalForceAndEnergy(S& force_divr, S& pair_eng, bool energy_shift)
        {
        if (this->rsq < this->rcutsq)
            {
            evalPair(force_divr, pair_eng, energy_shift);
            return true;
            }
        return false;
        }

    AZP_D void evalPair(S& force_divr, S& pair_eng, bool)
        {
        const S rinv = fast::rsqrt(this->rsq);
        const S r = this->rsq * rinv;
        force_divr = c.A * (rinv - c.rcut_inv);
        pair_eng = c.A * (c.rcut - r) - S(0.5) * c.A * c.rcut_inv * (this->rcutsq - this->rsq);
        }

    AZP_D bool evalForceEnergyThermo(S& force_divr, S& force_divr_cons, S& pair_eng, bool energy_shift)
        {
        if (this->rsq < this->rcutsq)
            {
            evalThermoPair(force_divr, force_divr_cons, pair_eng, energy_shift);
            return true;
            }
        return false;
        }

    AZP_D void evalThermoPair(S& force_divr, S& force_divr_cons, S& pair_eng, bool)
        {
        S r, rinv, base;
        if (c.half_s < S(1.0))
            {
            // s < 2: the weight (1 - r/r_cut)^(s/2) has an infinite slope at the cutoff -- a pair
            // one ulp of r below r_cut gets (6e-8)^(1/4) = 0.016, not ~0 -- so r and the base are
            // rounded exactly where the reference's host code rounds them (:203-204, :242):
            // rinv = 1 / sqrt(rsq), r = 1 / rinv, 1 - r * rcutinv without an FMA. The branch is
            // uniform per type pair.
            ref::r_and_rinv(this->rsq, r, rinv);
            base = fmax(ref::sub(S(1.0), ref::mul(r, c.rcut_inv)), S(0));
            }
        else
            {
            rinv = fast::rsqrt(this->rsq);
            r = this->rsq * rinv;
            base = fmax(S(1.0) - r * c.rcut_inv, S(0));
            }
        const S alpha = dpd_uniform_pm1<S>(m_seed, m_i, m_j, m_timestep);
        const S fc = c.A * (rinv - c.rcut_inv);
        force_divr_cons = fc;
        const S wR = fast::pow(base, c.half_s) * rinv;
        S f = fc - c.gamma * wR * wR * m_dot;
        f += c.noise * wR * alpha;
        force_divr = f;
        pair_eng = c.A * (c.rcut - r) - S(0.5) * c.A * c.rcut_inv * (this->rcutsq - this->rsq);
        }

    static const char* getName()
        {
        return "dpd_gen";
        }
    static void pack(const double* f, param_type* p)
        {
        p->A = S(f[0]);
        p->gamma = S(f[1]);
        p->s = S(f[2]);
        }
    static void unpack(const param_type* p, double* f)
        {
        f[0] = double(p->A);
        f[1] = double(p->gamma);
        f[2] = double(p->s);
        }
    static constexpr int num_fields = 3;

    private:
    const cache_type& c;
    uint16_t m_seed;
    unsigned int m_i, m_j, m_timestep;
    S m_dot;
    };
    } // namespace azp
#endif
