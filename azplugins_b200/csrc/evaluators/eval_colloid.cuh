// eval_colloid.cuh -- integrated Lennard-Jones ("colloid") potential with three couplings chosen
// by the radii of the type pair: point-point, sphere-point, sphere-sphere (Everaers-Ejtehadi).
// Behaviour: reference src/PairEvaluatorColloid.h:23-57 (param_type), :101-113 (solvent-solvent),
// :125-152 (colloid-solvent), :164-220 (colloid-colloid), :233-269 (dispatch, shift).
//
// The coupling is a property of the type pair, so it is resolved once in make_cache (with the
// energy at r_cut) instead of by three radius comparisons per neighbour. The sphere-sphere
// branch suffers catastrophic cancellation in fp32 (differences of k^-7 terms); it is rare
// (colloid-colloid contacts) and is therefore evaluated with IEEE division/sqrt/log rather than
// the SFU approximations, which keeps it within rounding of the reference's fp32 CPU result.
#ifndef AZP_EVAL_COLLOID_CUH_
#define AZP_EVAL_COLLOID_CUH_

#include "eval_base.cuh"
#include <cmath>

namespace azp
    {
template<class S> class PairEvaluatorColloid : public PairEvaluatorBase<S>
    {
    public:
    static constexpr int evaluator_id = 2;
    struct alignas(4 * sizeof(S)) param_type : public PairParametersBase
        {
        S A;
        S a_1;
        S a_2;
        S sigma_3;
        };

    enum Coupling
        {
        SolventSolvent = 0,
        ColloidSolvent = 1,
        ColloidColloid = 2
        };

    // k and asq are what the two light forms need of (A, sigma, a): hoisted per type pair with
    // the reference's own operation order, so the per-pair results keep their bits --
    //   point-point:  k = A sigma^6 / 36               (reference :105, an IEEE division per pair)
    //   sphere-point: k = sigma^3 A a a^2, asq = a^2   (reference :127-135), asq3 = a^2 / 3 (:148)
    struct alignas(16) cache_type
        {
        S A;
        S sigma_3;
        S sigma_6;
        S ai;
        S aj;
        S e_cut;
        int coupling;
        S k;
        S asq;
        S asq3;
        };

    // ---- the three couplings; `force` selects whether force_divr is produced -------------
    template<bool force> AZP_HD static S solventSolvent(const cache_type& c, S rsq, S& fdr)
        {
#ifdef __CUDA_ARCH__
        const S r2inv = fast::rcp(rsq);
#else
        const S r2inv = S(1.0) / rsq;
#endif
        const S r6inv = r2inv * r2inv * r2inv;
        const S c1 = c.k;
        if (force)
            fdr = S(6.0) * c1 * r2inv * r6inv * (S(2.0) * c.sigma_6 * r6inv - S(1.0));
        return c1 * r6inv * (c.sigma_6 * r6inv - S(1.0));
        }

    template<bool force> AZP_HD static S colloidSolvent(const cache_type& c, S rsq, S& fdr)
        {
        const S asq = c.asq;
        const S d = asq - rsq;
        const S r4 = rsq * rsq;
        const S d3 = d * d * d;
        const S d6 = d3 * d3;
#ifdef __CUDA_ARCH__
        const S d3inv = fast::rcp(d3);
        const S d6inv = d3inv * d3inv;
        const S dinv = fast::rcp(d);
#else
        const S d3inv = S(1.0) / d3;
        const S d6inv = S(1.0) / d6;
        const S dinv = S(1.0) / d;
#endif
        const S fR = c.k * d3inv;
        if (force)
            {
            fdr = S(4.0 / 15.0) * fR
                  * (S(2.0) * (asq + rsq) * (asq * (S(5.0) * asq + S(22.0) * rsq) + S(5.0) * r4)
                         * c.sigma_6 * d6inv
                     - S(5.0))
                  * dinv;
            }
        return S(2.0 / 9.0) * fR
               * (S(1.0)
                  - (asq * (asq * (c.asq3 + S(3.0) * rsq) + S(4.2) * r4) + rsq * r4)
                        * c.sigma_6 * d6inv);
        }

    AZP_HD static S inv7(S x)
        {
        const S xi = S(1.0) / x;
        S g = xi * xi;
        g *= g * g;
        g *= xi;
        return g;
        }

    template<bool force> AZP_HD static S colloidColloid(const cache_type& c, S rsq, S& fdr)
        {
        const S r = ::sqrt(rsq);
        const S prod = c.ai * c.aj, sum = c.ai + c.aj, diff = c.ai - c.aj;
        const S sum_p = sum + r, sum_m = sum - r, diff_p = diff + r, diff_m = diff - r;
        const S inv_sum = S(1.0) / (sum_p * sum_m);
        const S inv_diff = S(1.0) / (diff_p * diff_m);
        S w_sp = inv7(sum_p), w_sm = inv7(sum_m), w_dp = inv7(diff_p), w_dm = inv7(diff_m);
        const S e_sp = ((sum_p + S(5.0) * sum) * sum_p + S(30.0) * prod) * w_sp;
        const S e_sm = ((sum_m + S(5.0) * sum) * sum_m + S(30.0) * prod) * w_sm;
        const S e_dp = ((diff_p + S(5.0) * diff) * diff_p - S(30.0) * prod) * w_dp;
        const S e_dm = ((diff_m + S(5.0) * diff) * diff_m - S(30.0) * prod) * w_dm;
        w_sp *= S(42.0) * prod / sum_p + S(6.0) * sum + sum_p;
        w_sm *= S(42.0) * prod / sum_m + S(6.0) * sum + sum_m;
        w_dp *= S(-42.0) * prod / diff_p + S(6.0) * diff + diff_p;
        w_dm *= S(-42.0) * prod / diff_m + S(6.0) * diff + diff_m;
        const S fR = c.A * c.sigma_6 / r / S(37800.0);
        S eng = fR * (e_sp - e_sm - e_dp + e_dm);
        if (force)
            {
            const S dUR = eng / r + S(5.0) * fR * (w_sp + w_sm - w_dp - w_dm);
            const S dUA = -c.A / S(3.0) * r
                          * ((S(2.0) * prod * inv_sum + S(1.0)) * inv_sum + (S(2.0) * prod * inv_diff - S(1.0)) * inv_diff);
            fdr = (dUR + dUA) / r;
            }
        eng += c.A / S(6.0) * (S(2.0) * prod * (inv_sum + inv_diff) - ::log(inv_diff / inv_sum));
        return eng;
        }

    AZP_HD static cache_type make_cache(const param_type& p, S rcutsq, bool energy_shift)
        {
        cache_type c;
        c.A = p.A;
        c.sigma_3 = p.sigma_3;
        c.sigma_6 = p.sigma_3 * p.sigma_3;
        c.ai = p.a_1;
        c.aj = p.a_2;
        if (p.a_1 == S(0) && p.a_2 == S(0))
            c.coupling = SolventSolvent;
        else if (p.a_1 != S(0) && p.a_2 != S(0))
            c.coupling = ColloidColloid;
        else
            c.coupling = ColloidSolvent;
        c.k = S(0), c.asq = S(0), c.asq3 = S(0);
        if (c.coupling == SolventSolvent)
            c.k = c.A * c.sigma_6 / S(36.0);
        else if (c.coupling == ColloidSolvent)
            {
            const S a = (c.ai > c.aj) ? c.ai : c.aj;
            c.asq = a * a;
            c.asq3 = c.asq / S(3.0);
            c.k = c.sigma_3 * c.A * a * c.asq;
            }
        c.e_cut = S(0);
        if (energy_shift && p.A != S(0))
            {
            S unused = S(0);
            if (c.coupling == SolventSolvent)
                c.e_cut = solventSolvent<false>(c, rcutsq, unused);
            else if (c.coupling == ColloidColloid)
                c.e_cut = colloidColloid<false>(c, rcutsq, unused);
            else
                c.e_cut = colloidSolvent<false>(c, rcutsq, unused);
            }
        return c;
        }

    AZP_D PairEvaluatorColloid(S _rsq, S _rcutsq, const cache_type& _c)
        : PairEvaluatorBase<S>(_rsq, _rcutsq), c(_c)
        {
        }

    // see IsoFamily::pair: skip the evaluator when no lane of the warp is inside the cutoff
    static constexpr bool kWarpVote = true;
    // rare-form deferral (pair_kernels.cuh, FormSplit): the three couplings are three forms,
    // cheapest first
#ifndef AZP_COLLOID_NO_SPLIT
    static constexpr bool kSplitForms = true;
#endif
    static constexpr int kHeavyForm = ColloidColloid; // never evaluated inside the neighbour loop
    AZP_HD static int form(const cache_type& c)
        {
        return c.coupling;
        }
    // evalPair for the forms below kHeavyForm
    AZP_D void evalPairLight(S& force_divr, S& pair_eng)
        {
        S e;
        if (c.coupling == SolventSolvent)
            e = solventSolvent<true>(c, this->rsq, force_divr);
        else
            e = colloidSolvent<true>(c, this->rsq, force_divr);
        pair_eng = e - c.e_cut;
        }

    AZP_HD static bool disabled(const cache_type& c)
        {
        return c.A == S(0);
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
            S e;
            if (c.coupling == SolventSolvent)
                e = solventSolvent<true>(c, this->rsq, force_divr);
            else if (c.coupling == ColloidSolvent)
                e = colloidSolvent<true>(c, this->rsq, force_divr);
            else
                e = colloidColloid<true>(c, this->rsq, force_divr);
            pair_eng = e - c.e_cut;
            }
        }

    static const char* getName()
        {
        return "colloid";
        }
    // fields {A, a_1, a_2, sigma}
    static void pack(const double* f, param_type* p)
        {
        p->A = S(f[0]);
        p->a_1 = S(f[1]);
        p->a_2 = S(f[2]);
        const S sigma = S(f[3]);
        p->sigma_3 = sigma * sigma * sigma;
        }
    static void unpack(const param_type* p, double* f)
        {
        f[0] = double(p->A);
        f[1] = double(p->a_1);
        f[2] = double(p->a_2);
        f[3] = double(std::cbrt(p->sigma_3));
        }
    static constexpr int num_fields = 4;

    private:
    const cache_type& c;
    };
// Tuning traits of the isotropic family for this evaluator. AZP_COLLOID_CAP=96 caps the one-lane
// kernel at 96 registers (five 128-thread CTAs per SM instead of four at its natural 106; 12
// bytes of spills outside the neighbour loop). Measured on C3 (profiles/ab/r02d_ab_c3_cap96.json):
// main pass 1.312 -> 1.266 ms in the tuner's back-to-back timing, but the step (Colloid and
// Hertz launches alternating) 1.438 -> 1.448 ms -- no gain, so the default is uncapped.
#ifndef AZP_COLLOID_CAP
#define AZP_COLLOID_CAP 0
#endif
template<class E> struct IsoTraits;
template<class S> struct IsoTraits<PairEvaluatorColloid<S>>
    {
#ifdef AZP_COLLOID_PIPE0
    static constexpr int pipe = 0;
#else
    static constexpr int pipe = 2;
#endif
#ifdef AZP_COLLOID_SMEM
    static constexpr bool register_tables = false;
#else
    static constexpr bool register_tables = true;
#endif
    static constexpr int one_lane_cap = AZP_COLLOID_CAP;
    };
    } // namespace azp
#endif
