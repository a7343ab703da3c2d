// eval_morse.cuh -- two-patch Morse: U = U_Morse(r) Omega(rhat.n_i) Omega(rhat.n_j), forces and
// torques. Behaviour: reference src/AnisoPairEvaluatorTwoPatchMorse.h:32-69 (param_type),
// :127-216 (evaluate); contract of src/AnisoPairEvaluator.h:97-215 (no charge/shape/tags,
// implementsEnergyShift).
//
// The patch director n = rotate(q, (1,0,0)) is written out for the unit x axis:
//   n = (s^2 - |v|^2 + 2 x^2, 2 (s z + x y), 2 (x z - s y))        q = (s, x, y, z)
// n_i depends only on the row, so the kernel's compiler hoists it out of the neighbour loop.
// U_Morse(r_cut) for the energy shift is hoisted per type pair.
//
// Rounding: the potential is stiff (M_r = 0.03, omega = 20), so the fp32 build rounds r, rhat, n,
// gamma and Omega exactly where the reference's host code rounds them (see evaluatePair); the
// products after Omega are not amplified and use contracted FMAs.
#ifndef AZP_EVAL_MORSE_CUH_
#define AZP_EVAL_MORSE_CUH_

#include "eval_base.cuh"

namespace azp
    {
struct AnisoShapeParametersEmpty
    {
    AZP_D void load_shared(char*&, unsigned int&) { }
    AZP_HD void allocate_shared(char*&, unsigned int&) const { }
    void set_memory_hint() const { }
    };

template<class S> AZP_D Vec3<S> patch_director(const Vec4<S>& q)
    {
    // Scalar4 (x,y,z,w) carries the quaternion (s, v.x, v.y, v.z). HOOMD's
    // rotate(q, b) = (s^2 - |v|^2) b + 2 s (v x b) + 2 (v . b) v for b = (1, 0, 0), rounding
    // where the library routine rounds (the terms multiplied by the zeros of b drop out exactly)
    const S s = q.x, a = q.y, b = q.z, c = q.w;
    const S c0 = ref::sub(ref::mul(s, s), ref::dot3(a, b, c, a, b, c));
    const S c1 = S(2) * s;
    const S c2 = S(2) * a;
    Vec3<S> n;
    n.x = ref::add(c0, ref::mul(c2, a));
    n.y = ref::add(ref::mul(c1, c), ref::mul(c2, b));
    n.z = ref::add(ref::mul(c1, -b), ref::mul(c2, c));
    return n;
    }

template<class S> class AnisoPairEvaluatorTwoPatchMorse
    {
    public:
    static constexpr int evaluator_id = 5;
    // 5 Scalars + bool, no alignment attribute: 24 B (fp32) / 48 B (fp64) like the reference
    struct param_type : public PairParametersBase
        {
        S M_d;
        S M_rinv;
        S r_eq;
        S omega;
        S alpha;
        bool repulsion;
        };
    typedef AnisoShapeParametersEmpty shape_type;

    struct alignas(16) cache_type
        {
        S M_d;
        S M_rinv;
        S r_eq;
        S omega;
        S alpha;
        S U_cut; // U_Morse(r_cut) when shifting, else 0
        int repulsion;
        };

    AZP_HD static cache_type make_cache(const param_type& p, S rcutsq, bool energy_shift)
        {
        cache_type c;
        c.M_d = p.M_d;
        c.M_rinv = p.M_rinv;
        c.r_eq = p.r_eq;
        c.omega = p.omega;
        c.alpha = p.alpha;
        c.repulsion = p.repulsion ? 1 : 0;
        c.U_cut = S(0);
        if (energy_shift)
            {
            const S rcut = ::sqrt(rcutsq);
            const S e = ::exp(-(rcut - p.r_eq) * p.M_rinv);
            const S om = S(1.0) - e;
            c.U_cut = p.M_d * (om * om - S(1.0));
            }
        return c;
        }

    AZP_D AnisoPairEvaluatorTwoPatchMorse(const Vec3<S>& _dr,
                                          const Vec4<S>& _quat_i,
                                          const Vec4<S>& _quat_j,
                                          S _rcutsq,
                                          const cache_type& _c)
        : dr(_dr), rcutsq(_rcutsq), ni(patch_director(_quat_i)), quat_j(_quat_j), c(_c)
        {
        }

    // Kernel hook: what the evaluator needs of particle i's orientation, computed once per row
    // (the patch director; the kernels keep it in registers instead of the quaternion).
    typedef Vec3<S> row_type;
    AZP_D static row_type make_row(const Vec4<S>& quat_i)
        {
        return patch_director(quat_i);
        }
    struct FromRow
        {
        };
    AZP_D AnisoPairEvaluatorTwoPatchMorse(FromRow,
                                          const Vec3<S>& _dr,
                                          const row_type& _ni,
                                          const Vec4<S>& _quat_j,
                                          S _rcutsq,
                                          const cache_type& _c)
        : dr(_dr), rcutsq(_rcutsq), ni(_ni), quat_j(_quat_j), c(_c)
        {
        }

    AZP_HD static bool needsCharge()
        {
        return false;
        }
    AZP_D void setCharge(S, S) { }
    AZP_HD static bool needsShape()
        {
        return false;
        }
    AZP_D void setShape(const shape_type*, const shape_type*) { }
    AZP_HD static bool needsTags()
        {
        return false;
        }
    AZP_D void setTags(unsigned int, unsigned int) { }
    AZP_HD static constexpr bool implementsEnergyShift()
        {
        return true;
        }

    AZP_HD static bool disabled(const cache_type&)
        {
        return false;
        }

    AZP_D bool evaluate(Vec3<S>& force, S& pair_eng, bool energy_shift, Vec3<S>& torque_i, Vec3<S>& torque_j)
        {
        const S rsq = ref::dot3(dr.x, dr.y, dr.z, dr.x, dr.y, dr.z);
        if (rsq > rcutsq) // the reference rejects only strictly-greater (:135)
            return false;
        evaluatePair(rsq, force, pair_eng, energy_shift, torque_i, torque_j);
        return true;
        }

    // body of evaluate() for a pair already known to satisfy rsq <= rcutsq
    AZP_D void evaluatePair(S rsq, Vec3<S>& force, S& pair_eng, bool, Vec3<S>& torque_i, Vec3<S>& torque_j)
        {
        // The well exp(-(r - r_eq) / M_r) (M_r = 0.03: one ulp of r is 3e-6 of U, 6e-6 of its
        // repulsive wall) amplifies the fp32 rounding of r beyond the parity budget, and the
        // patch switch Omega(gamma) (omega = 20) that of gamma. The CPU reference's own fp32
        // result is farther from the exact one than the budget, so "more accurate" does not
        // help: r and gamma are rounded exactly where the reference's host code rounds them
        // (azp_core.cuh, namespace ref): rsq = dot(dr, dr) without FMA, rinv = 1 / sqrt(rsq),
        // r = 1 / rinv (:137-139), rhat = dr * rinv, n = rotate(q, ex), gamma = dot(rhat, n),
        // and the radial exponential is the math library's expf. (The argument rsq of this
        // function was accumulated with FMAs for the cutoff test.) Everything after that --
        // Omega's exponential and reciprocal, the force and torque assembly -- is not
        // amplified and uses the SFU approximations and contracted FMAs
        // (AZP_MORSE_FULL_REF: library exp and IEEE reciprocal there too, +18 % time on C5, kept
        // for A/B measurements).
        S r, rinv;
        ref::r_and_rinv(ref::dot3(dr.x, dr.y, dr.z, dr.x, dr.y, dr.z), r, rinv);
        const Vec3<S> u {ref::mul(dr.x, rinv), ref::mul(dr.y, rinv), ref::mul(dr.z, rinv)};
        const Vec3<S> nj = patch_director(quat_j);

        S UM = -c.M_d;
        S dUM = S(0);
        if (r > c.r_eq || c.repulsion)
            {
            const S me = ref::exp(ref::mul(-ref::sub(r, c.r_eq), c.M_rinv));
            const S om = ref::sub(S(1.0), me);
            UM = ref::mul(c.M_d, ref::sub(ref::mul(om, om), S(1.0)));
            dUM = S(2.0) * c.M_d * c.M_rinv * me * om;
            }
        const S gi = ref::dot3(u.x, u.y, u.z, ni.x, ni.y, ni.z);
        const S gj = ref::dot3(u.x, u.y, u.z, nj.x, nj.y, nj.z);
#ifdef AZP_MORSE_FULL_REF
        const S gie = ref::exp(ref::mul(-c.omega, ref::sub(ref::mul(gi, gi), c.alpha)));
        const S Oi = ref::rcp(ref::add(S(1.0), gie));
        const S gje = ref::exp(ref::mul(-c.omega, ref::sub(ref::mul(gj, gj), c.alpha)));
        const S Oj = ref::rcp(ref::add(S(1.0), gje));
#else
        const S gie = fast::exp(ref::mul(-c.omega, ref::sub(ref::mul(gi, gi), c.alpha)));
        const S Oi = fast::rcp(S(1.0) + gie);
        const S gje = fast::exp(ref::mul(-c.omega, ref::sub(ref::mul(gj, gj), c.alpha)));
        const S Oj = fast::rcp(S(1.0) + gje);
#endif

        const S OiOj = Oi * Oj;
        const S dU_dr = dUM * OiOj;
        const S dU_dgi = (S(2.0) * c.omega * gi * gie * Oi * Oi) * UM * Oj;
        const S dU_dgj = (S(2.0) * c.omega * gj * gje * Oj * Oj) * UM * Oi;

        // u x n (the torque direction) and the in-plane director
        // n_perp = -u x (u x n) = n - (u.n) u = n - gamma u   (|u| = 1)
        const Vec3<S> ci {u.y * ni.z - u.z * ni.y, u.z * ni.x - u.x * ni.z, u.x * ni.y - u.y * ni.x};
        const Vec3<S> cj {u.y * nj.z - u.z * nj.y, u.z * nj.x - u.x * nj.z, u.x * nj.y - u.y * nj.x};
        const Vec3<S> pi {fma(-gi, u.x, ni.x), fma(-gi, u.y, ni.y), fma(-gi, u.z, ni.z)};
        const Vec3<S> pj {fma(-gj, u.x, nj.x), fma(-gj, u.y, nj.y), fma(-gj, u.z, nj.z)};

        const S ai = rinv * dU_dgi, aj = rinv * dU_dgj;
        force.x = -(dU_dr * u.x + (ai * pi.x + aj * pj.x));
        force.y = -(dU_dr * u.y + (ai * pi.y + aj * pj.y));
        force.z = -(dU_dr * u.z + (ai * pi.z + aj * pj.z));
        torque_i = Vec3<S> {dU_dgi * ci.x, dU_dgi * ci.y, dU_dgi * ci.z};
        torque_j = Vec3<S> {dU_dgj * cj.x, dU_dgj * cj.y, dU_dgj * cj.z};
        pair_eng = (UM - c.U_cut) * OiOj;
        }

    static const char* getName()
        {
        return "TwoPatchMorse";
        }
    static const char* getShapeParamName()
        {
        return "";
        }
    // fields {M_d, M_r, r_eq, omega, alpha, repulsion}
    static void pack(const double* f, param_type* p)
        {
        p->M_d = S(f[0]);
        p->M_rinv = S(1.0) / S(f[1]);
        p->r_eq = S(f[2]);
        p->omega = S(f[3]);
        p->alpha = S(f[4]);
        p->repulsion = (f[5] != 0.0);
        }
    static void unpack(const param_type* p, double* f)
        {
        f[0] = double(p->M_d);
        f[1] = double(S(1.0) / p->M_rinv);
        f[2] = double(p->r_eq);
        f[3] = double(p->omega);
        f[4] = double(p->alpha);
        f[5] = p->repulsion ? 1.0 : 0.0;
        }
    static constexpr int num_fields = 6;

    private:
    Vec3<S> dr;
    S rcutsq;
    Vec3<S> ni;
    Vec4<S> quat_j;
    const cache_type& c;
    };
    } // namespace azp
#endif
