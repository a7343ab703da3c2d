// eval_yukawa.cuh -- expanded (shifted-origin) Yukawa: U = eps exp(-kappa (r - delta)) / (r - delta).
// Behaviour: reference src/PairEvaluatorExpandedYukawa.h:23-53 (param_type), :92-113.
#ifndef AZP_EVAL_YUKAWA_CUH_
#define AZP_EVAL_YUKAWA_CUH_

#include "eval_base.cuh"

namespace azp
    {
template<class S> class PairEvaluatorExpandedYukawa : public PairEvaluatorBase<S>
    {
    public:
    static constexpr int evaluator_id = 1;
    struct alignas(4 * sizeof(S)) param_type : public PairParametersBase
        {
        S epsilon;
        S kappa;
        S delta;
        };

    struct alignas(16) cache_type
        {
        S epsilon;
        S kappa;
        S delta;
        S e_cut; // eps exp(-kappa (r_cut - delta)) / (r_cut - delta) when shifting, else 0
        S neg_kappa_scaled; // -kappa * log2(e) (fp32: argument of ex2) or -kappa (fp64)
        };

    AZP_HD static cache_type make_cache(const param_type& p, S rcutsq, bool energy_shift)
        {
        cache_type c;
        c.epsilon = p.epsilon;
        c.kappa = p.kappa;
        c.delta = p.delta;
        c.neg_kappa_scaled = fast::exp_scale(-p.kappa);
        c.e_cut = S(0);
        if (energy_shift)
            {
            const S rcut = ::sqrt(rcutsq);
            const S rcd = rcut - p.delta;
            c.e_cut = p.epsilon * ::exp(-p.kappa * rcd) / rcd;
            }
        return c;
        }

    AZP_D PairEvaluatorExpandedYukawa(S _rsq, S _rcutsq, const cache_type& _c)
        : PairEvaluatorBase<S>(_rsq, _rcutsq), c(_c)
        {
        }

    // see IsoFamily::pair: skip the evaluator when no lane of the warp is inside the cutoff
    static constexpr bool kWarpVote = false;

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
            const S rd = r - c.delta;
            const S rdinv = fast::rcp(rd);
            const S e = c.epsilon * fast::exp_prescaled(c.neg_kappa_scaled * rd) * rdinv;
            force_divr = e * kappa_plus(rdinv) * rinv;
            pair_eng = e - c.e_cut;
            }
        }

    // kappa + x. fp32: kappa is recovered from the staged -kappa log2(e) inside the FMA
    // (kappa' = fl(-kappa log2 e) * -ln 2, within 2 ulp of kappa: 1e-7 of the force), so a
    // two-type row selects one constant less per neighbour
    AZP_D float kappa_plus(float x) const
        {
#ifdef AZP_YUKAWA_KEEP_KAPPA
        return c.kappa + x;
#else
        return __fmaf_rn(c.neg_kappa_scaled, -0.6931471805599453f, x);
#endif
        }
    AZP_D double kappa_plus(double x) const
        {
        return c.kappa + x;
        }

    static const char* getName()
        {
        return "ExpandedYukawa";
        }
    static void pack(const double* f, param_type* p)
        {
        p->epsilon = S(f[0]);
        p->kappa = S(f[1]);
        p->delta = S(f[2]);
        }
    static void unpack(const param_type* p, double* f)
        {
        f[0] = double(p->epsilon);
        f[1] = double(p->kappa);
        f[2] = double(p->delta);
        }
    static constexpr int num_fields = 3;

    private:
    const cache_type& c;
    };
#ifdef AZP_YUKAWA_REG_CAP
// A/B: cap the one-lane kernel of this evaluator at 72 / 80 registers (pair_kernels.cuh, OneLaneCap)
template<class E> struct IsoTraits;
template<class S> struct IsoTraits<PairEvaluatorExpandedYukawa<S>>
    {
    static constexpr int pipe = 2;
    static constexpr bool register_tables = true;
    static constexpr int one_lane_cap = AZP_YUKAWA_REG_CAP;
    };
#endif
    } // namespace azp
#endif
