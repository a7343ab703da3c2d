// TEST INFRASTRUCTURE (oracle). Not part of the product path: only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may build or call this.
//
// CPU restatement ("port") of the per-pair arithmetic of the six azplugins evaluators on the
// pair-force hot path. Each function cites the reference lines it follows. The restatement is
// validated bit-for-bit against the reference headers themselves (oracle/_ref, built from
// /root/reference/src in place) by tests/test_oracle.py, and against the reference's own
// known-answer vectors (reference src/pytest/test_pair.py:23-306, test_pair_aniso.py:22-110).
//
// S is the HOOMD `Scalar` (float for HOOMD_LONGREAL_SIZE=32, double for 64). Literal types are
// deliberately kept as in the reference where C++ promotion changes the fp32 result (Colloid).
#ifndef AZP_ORACLE_PORT_EVALUATORS_H_
#define AZP_ORACLE_PORT_EVALUATORS_H_

#include <cmath>
#include <cstdint>
#include <math.h>

namespace azp_oracle
    {
// ---- HOOMD fast:: on the host: plain libm in S precision --------------
template<class S> inline S h_rsqrt(S x)
    {
    return S(1.0) / std::sqrt(x);
    }

// ---- Philox4x32-10 + HOOMD Seed/Counter packing --------------------------
inline void philox4x32_10(const uint32_t c_in[4], const uint32_t k_in[2], uint32_t out[4])
    {
    uint32_t c0 = c_in[0], c1 = c_in[1], c2 = c_in[2], c3 = c_in[3];
    uint32_t k0 = k_in[0], k1 = k_in[1];
    for (int r = 0; r < 10; ++r)
        {
        const uint64_t a = 0xD2511F53ull * c0;
        const uint64_t b = 0xCD9E8D57ull * c2;
        const uint32_t n0 = uint32_t(b >> 32) ^ c1 ^ k0;
        const uint32_t n2 = uint32_t(a >> 32) ^ c3 ^ k1;
        c1 = uint32_t(b);
        c3 = uint32_t(a);
        c0 = n0;
        c2 = n2;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
        }
    out[0] = c0;
    out[1] = c1;
    out[2] = c2;
    out[3] = c3;
    }

// First uniform(-1,1) draw of RandomGenerator(Seed(200, timestep32, seed16), Counter(lo, hi)),
// reference src/DPDPairEvaluatorGeneralWeight.h:213-233, src/RNGIdentifiers.h:20-23.
template<class S> inline S dpd_alpha(uint16_t seed, uint32_t tag_i, uint32_t tag_j, uint32_t timestep)
    {
    const uint32_t lo = tag_i > tag_j ? tag_j : tag_i;
    const uint32_t hi = tag_i > tag_j ? tag_i : tag_j;
    // the reference truncates the step to `unsigned int` (DPDPairEvaluatorGeneralWeight.h:121-122)
    // so bits 32..39 of the 64-bit step never reach key word 0.
    const uint32_t key[2] = {(200u << 24) | (uint32_t(seed) << 8), timestep};
    const uint32_t ctr[4] = {0u, 0u, hi, lo}; // Counter(a=lo, b=hi) -> {d<<16, c, b, a}
    uint32_t u[4];
    philox4x32_10(ctr, key, u);
    S canonical;
    if (sizeof(S) == 4)
        canonical = S(float(u[0]) * 2.3283064365386963e-10f + 1.1641532182693481e-10f);
    else
        canonical = S(double((uint64_t(u[0]) << 32) | u[1]) * 5.421010862427522e-20
                      + 2.710505431213761e-20);
    return S(-1) + S(2) * canonical;
    }

// ---- PerturbedLennardJones --------------------------------------------------------------------
template<class S> struct alignas(4 * sizeof(S)) PLJParams
    {
    S sigma_6, epsilon_x_4, lambda, rwcasq;
    };

// fields = {epsilon, sigma, attraction_scale_factor}; reference
// src/PairEvaluatorPerturbedLennardJones.h:33-45 (all narrowing happens before the products;
// 2^(1/3) is evaluated in double and multiplied into Scalar sigma^2 in double, then narrowed).
template<class S> inline void pack_plj(const double* f, PLJParams<S>* p)
    {
    const S eps = S(f[0]), sig = S(f[1]);
    const S s2 = sig * sig;
    const S s4 = s2 * s2;
    p->sigma_6 = s2 * s4;
    p->epsilon_x_4 = S(4.0) * eps;
    p->lambda = S(f[2]);
    p->rwcasq = S(std::pow(double(S(2.)), 1. / 3.) * s2);
    }

// reference src/PairEvaluatorPerturbedLennardJones.h:96-104 (ctor), :117-155
template<class S>
inline bool eval_plj(S rsq, S rcutsq, const PLJParams<S>& p, bool shift, S& fdr, S& eng)
    {
    const S lj1 = p.epsilon_x_4 * p.sigma_6 * p.sigma_6;
    const S lj2 = p.epsilon_x_4 * p.sigma_6;
    const S lam = p.lambda;
    const S wca = p.epsilon_x_4 * (S(1.0) - lam) / S(4.0);
    if (!(rsq < rcutsq && lj1 != 0))
        return false;
    const S i2 = S(1.0) / rsq;
    const S i6 = i2 * i2 * i2;
    fdr = i2 * i6 * (S(12.0) * lj1 * i6 - S(6.0) * lj2);
    eng = i6 * (lj1 * i6 - lj2);
    if (rsq < p.rwcasq)
        eng += wca;
    else
        {
        fdr *= lam;
        eng *= lam;
        }
    if (shift)
        {
        const S c2 = S(1.0) / rcutsq;
        const S c6 = c2 * c2 * c2;
        S es = c6 * (lj1 * c6 - lj2);
        if (rcutsq < p.rwcasq)
            es += wca;
        else
            es *= lam;
        eng -= es;
        }
    return true;
    }

// ---- ExpandedYukawa -----------------------------------------------------------------------------
template<class S> struct alignas(4 * sizeof(S)) YukawaParams
    {
    S epsilon, kappa, delta;
    };
template<class S> inline void pack_yukawa(const double* f, YukawaParams<S>* p)
    {
    p->epsilon = S(f[0]); // reference src/PairEvaluatorExpandedYukawa.h:28-33
    p->kappa = S(f[1]);
    p->delta = S(f[2]);
    }
// reference src/PairEvaluatorExpandedYukawa.h:92-113
template<class S>
inline bool eval_yukawa(S rsq, S rcutsq, const YukawaParams<S>& p, bool shift, S& fdr, S& eng)
    {
    if (!(rsq < rcutsq && p.epsilon != S(0)))
        return false;
    const S r = std::sqrt(rsq);
    const S rd = r - p.delta;
    const S rdi = S(1.0) / rd;
    eng = p.epsilon * std::exp(-p.kappa * rd) * rdi;
    fdr = eng * (p.kappa + rdi) / r;
    if (shift)
        {
        const S rc = std::sqrt(rcutsq);
        const S rcd = rc - p.delta;
        eng -= p.epsilon * std::exp(-p.kappa * rcd) / rcd;
        }
    return true;
    }

// ---- Hertz --------------------------------------------------------------------------------------
template<class S> struct alignas(sizeof(S)) HertzParams
    {
    S epsilon;
    };
template<class S> inline void pack_hertz(const double* f, HertzParams<S>* p)
    {
    p->epsilon = S(f[0]); // reference src/PairEvaluatorHertz.h:28-31
    }
// reference src/PairEvaluatorHertz.h:93-109 (energy_shift ignored: U(r_cut)=0 already)
template<class S>
inline bool eval_hertz(S rsq, S rcutsq, const HertzParams<S>& p, bool, S& fdr, S& eng)
    {
    if (!(rsq < rcutsq && p.epsilon != S(0)))
        return false;
    const S r = std::sqrt(rsq);
    const S rc = std::sqrt(rcutsq);
    const S x = S(1.0) - (r / rc);
    const S e32 = p.epsilon * x * std::sqrt(x);
    fdr = S(2.5) * e32 / (r * rc);
    eng = e32 * x;
    return true;
    }

// ---- Colloid ------------------------------------------------------------------------------------
template<class S> struct alignas(4 * sizeof(S)) ColloidParams
    {
    S A, a_1, a_2, sigma_3;
    };
// fields = {A, a_1, a_2, sigma}; reference src/PairEvaluatorColloid.h:29-36
template<class S> inline void pack_colloid(const double* f, ColloidParams<S>* p)
    {
    p->A = S(f[0]);
    p->a_1 = S(f[1]);
    p->a_2 = S(f[2]);
    const S sig = S(f[3]);
    p->sigma_3 = sig * sig * sig;
    }

// solvent-solvent, reference src/PairEvaluatorColloid.h:101-113
template<class S, bool force> inline S colloid_ss(S A, S s6, S rsq, S& fdr)
    {
    const S i2 = S(1.0) / rsq;
    const S i6 = i2 * i2 * i2;
    const S c1 = A * s6 / S(36.0);
    if (force)
        fdr = S(6.0) * c1 * i2 * i6 * (S(2.0) * s6 * i6 - S(1.0));
    return c1 * i6 * (s6 * i6 - S(1.0));
    }

// colloid-solvent, reference src/PairEvaluatorColloid.h:125-152. The bare `2.0` / `5.0` double
// literals (:140-141) promote that sub-expression to double in an fp32 build; kept as written.
template<class S, bool force> inline S colloid_cs(S A, S s3, S s6, S ai, S aj, S rsq, S& fdr)
    {
    const S a = (ai > aj) ? ai : aj;
    const S asq = a * a;
    const S d = asq - rsq;
    const S r4 = rsq * rsq;
    const S d3 = d * d * d;
    const S d6 = d3 * d3;
    const S fR = s3 * A * a * asq / d3;
    if (force)
        {
        fdr = S(4.0 / 15.0) * fR
              * (2.0 * (asq + rsq) * (asq * (S(5.0) * asq + S(22.0) * rsq) + 5.0 * r4) * s6 / d6
                 - S(5.0))
              / d;
        }
    return S(2.0 / 9.0) * fR
           * (S(1.0)
              - (asq * (asq * (asq / S(3.0) + S(3.0) * rsq) + S(4.2) * r4) + rsq * r4) * s6 / d6);
    }

// 1/x^7 built the way the reference does (x^-2, cubed by two multiplies, times x^-1), :177-195
template<class S> inline S inv7(S x)
    {
    const S xi = S(1.0) / x;
    S g = xi * xi;
    g *= g * g;
    g *= xi;
    return g;
    }

// colloid-colloid (Everaers-Ejtehadi), reference src/PairEvaluatorColloid.h:164-220
template<class S, bool force> inline S colloid_cc(S A, S s6, S ai, S aj, S rsq, S& fdr)
    {
    const S r = std::sqrt(rsq);
    const S k0 = ai * aj, k1 = ai + aj, k2 = ai - aj;
    const S k3 = k1 + r, k4 = k1 - r, k5 = k2 + r, k6 = k2 - r;
    const S k7 = S(1.0) / (k3 * k4);
    const S k8 = S(1.0) / (k5 * k6);
    S g0 = inv7(k3), g1 = inv7(k4), g2 = inv7(k5), g3 = inv7(k6);

    const S h0 = ((k3 + S(5.0) * k1) * k3 + S(30.0) * k0) * g0;
    const S h1 = ((k4 + S(5.0) * k1) * k4 + S(30.0) * k0) * g1;
    const S h2 = ((k5 + S(5.0) * k2) * k5 - S(30.0) * k0) * g2;
    const S h3 = ((k6 + S(5.0) * k2) * k6 - S(30.0) * k0) * g3;

    g0 *= S(42.0) * k0 / k3 + S(6.0) * k1 + k3;
    g1 *= S(42.0) * k0 / k4 + S(6.0) * k1 + k4;
    g2 *= S(-42.0) * k0 / k5 + S(6.0) * k2 + k5;
    g3 *= S(-42.0) * k0 / k6 + S(6.0) * k2 + k6;

    const S fR = A * s6 / r / S(37800.0);
    S eng = fR * (h0 - h1 - h2 + h3);
    if (force)
        {
        const S dUR = eng / r + S(5.0) * fR * (g0 + g1 - g2 - g3);
        const S dUA = -A / S(3.0) * r
                      * ((S(2.0) * k0 * k7 + S(1.0)) * k7 + (S(2.0) * k0 * k8 - S(1.0)) * S(k8));
        fdr = (dUR + dUA) / r;
        }
    eng += A / S(6.0) * (S(2.0) * k0 * (k7 + k8) - std::log(k8 / k7));
    return eng;
    }

// dispatch on the radii, reference src/PairEvaluatorColloid.h:233-269
template<class S>
inline bool eval_colloid(S rsq, S rcutsq, const ColloidParams<S>& p, bool shift, S& fdr, S& eng)
    {
    if (!(rsq < rcutsq && p.A != S(0)))
        return false;
    const S s3 = p.sigma_3, s6 = s3 * s3;
    S unused = fdr;
    if (p.a_1 == S(0) && p.a_2 == S(0))
        {
        eng = colloid_ss<S, true>(p.A, s6, rsq, fdr);
        if (shift)
            eng -= colloid_ss<S, false>(p.A, s6, rcutsq, unused);
        }
    else if (p.a_1 != S(0) && p.a_2 != S(0))
        {
        eng = colloid_cc<S, true>(p.A, s6, p.a_1, p.a_2, rsq, fdr);
        if (shift)
            eng -= colloid_cc<S, false>(p.A, s6, p.a_1, p.a_2, rcutsq, unused);
        }
    else
        {
        eng = colloid_cs<S, true>(p.A, s3, s6, p.a_1, p.a_2, rsq, fdr);
        if (shift)
            eng -= colloid_cs<S, false>(p.A, s3, s6, p.a_1, p.a_2, rcutsq, unused);
        }
    return true;
    }

// ---- DPD general weight ------------------------------------------------------------------------
template<class S> struct alignas(4 * sizeof(S)) DPDParams
    {
    S A, gamma, s;
    };
template<class S> inline void pack_dpd(const double* f, DPDParams<S>* p)
    {
    p->A = S(f[0]); // reference src/DPDPairEvaluatorGeneralWeight.h:37-42
    p->gamma = S(f[1]);
    p->s = S(f[2]);
    }
// conservative part, reference src/DPDPairEvaluatorGeneralWeight.h:165-186
template<class S>
inline bool eval_dpd_conservative(S rsq, S rcutsq, const DPDParams<S>& p, bool, S& fdr, S& eng)
    {
    if (!(rsq < rcutsq))
        return false;
    const S ri = h_rsqrt(rsq);
    const S r = S(1.0) / ri;
    const S rci = h_rsqrt(rcutsq);
    const S rc = S(1.0) / rci;
    fdr = p.A * (ri - rci);
    eng = p.A * (rc - r) - S(0.5) * p.A * rci * (rcutsq - rsq);
    return true;
    }
// thermostatted, reference src/DPDPairEvaluatorGeneralWeight.h:198-255
template<class S>
inline bool eval_dpd_thermo(S rsq,
                            S rcutsq,
                            const DPDParams<S>& p,
                            uint16_t seed,
                            uint32_t tag_i,
                            uint32_t tag_j,
                            uint32_t timestep,
                            S deltaT,
                            S rdotv,
                            S T,
                            S& fdr,
                            S& fdr_cons,
                            S& eng)
    {
    if (!(rsq < rcutsq))
        return false;
    const S ri = h_rsqrt(rsq);
    const S r = S(1.0) / ri;
    const S rci = h_rsqrt(rcutsq);
    const S rc = S(1.0) / rci;
    const S alpha = dpd_alpha<S>(seed, tag_i, tag_j, timestep);
    fdr = p.A * (ri - rci);
    fdr_cons = fdr;
    const S wR = std::pow(S(1.) - r * rci, S(0.5) * p.s) * ri;
    fdr -= p.gamma * wR * wR * rdotv;
    fdr += h_rsqrt(deltaT / (T * p.gamma * S(6.0))) * wR * alpha;
    eng = p.A * (rc - r) - S(0.5) * p.A * rci * (rcutsq - rsq);
    return true;
    }

// ---- TwoPatchMorse ------------------------------------------------------------------------------
// No alignment attribute in the reference (src/AnisoPairEvaluatorTwoPatchMorse.h:63-69):
// 5 Scalars + bool, sizeof 24 (fp32) / 48 (fp64).
template<class S> struct MorseParams
    {
    S M_d, M_rinv, r_eq, omega, alpha;
    bool repulsion;
    };
// fields = {M_d, M_r, r_eq, omega, alpha, repulsion}; reference
// src/AnisoPairEvaluatorTwoPatchMorse.h:40-48
template<class S> inline void pack_morse(const double* f, MorseParams<S>* p)
    {
    p->M_d = S(f[0]);
    p->M_rinv = S(1.0) / S(f[1]);
    p->r_eq = S(f[2]);
    p->omega = S(f[3]);
    p->alpha = S(f[4]);
    p->repulsion = (f[5] != 0.0);
    }

template<class S> struct V3
    {
    S x, y, z;
    };
template<class S> inline V3<S> v_cross(const V3<S>& a, const V3<S>& b)
    {
    return V3<S> {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
    }
template<class S> inline S v_dot(const V3<S>& a, const V3<S>& b)
    {
    return a.x * b.x + a.y * b.y + a.z * b.z;
    }
// HOOMD rotate(q, v) with q = (s, u), applied to v = (1,0,0) in full
// generality so the rounding sequence equals the library routine's.
template<class S> inline V3<S> q_rotate(const S q[4], const V3<S>& b)
    {
    const S s = q[0];
    const V3<S> u {q[1], q[2], q[3]};
    const S c0 = s * s - v_dot(u, u);
    const V3<S> ub = v_cross(u, b);
    const S c1 = S(2) * s;
    const S c2 = S(2) * v_dot(u, b);
    return V3<S> {c0 * b.x + c1 * ub.x + c2 * u.x,
                  c0 * b.y + c1 * ub.y + c2 * u.y,
                  c0 * b.z + c1 * ub.z + c2 * u.z};
    }

// reference src/AnisoPairEvaluatorTwoPatchMorse.h:127-216. quat arrays are (s, x, y, z).
template<class S>
inline bool eval_morse(const V3<S>& dr,
                       const S qi[4],
                       const S qj[4],
                       S rcutsq,
                       const MorseParams<S>& p,
                       bool shift,
                       V3<S>& force,
                       S& eng,
                       V3<S>& tq_i,
                       V3<S>& tq_j)
    {
    const S rsq = dr.x * dr.x + dr.y * dr.y + dr.z * dr.z;
    if (rsq > rcutsq) // note: only strictly-greater is rejected (:135)
        return false;
    const S ri = h_rsqrt(rsq);
    const S r = S(1.) / ri;
    const V3<S> rh {dr.x * ri, dr.y * ri, dr.z * ri};
    const V3<S> ex {S(1.0), S(0), S(0)};
    const V3<S> ni = q_rotate(qi, ex);
    const V3<S> nj = q_rotate(qj, ex);

    S UM = S(-1.0) * p.M_d;
    S dUM = S(0.0);
    if (r > p.r_eq || p.repulsion)
        {
        const S me = std::exp(-(r - p.r_eq) * p.M_rinv);
        const S om = S(1.0) - me;
        UM = p.M_d * (om * om - S(1.0));
        dUM = S(2.0) * p.M_d * p.M_rinv * me * om;
        }
    const S gi = v_dot(rh, ni);
    const S gie = std::exp(-p.omega * (gi * gi - p.alpha));
    const S Oi = S(1.0) / (S(1.0) + gie);
    const S gj = v_dot(rh, nj);
    const S gje = std::exp(-p.omega * (gj * gj - p.alpha));
    const S Oj = S(1.0) / (S(1.0) + gje);

    S e = S(0.0);
    e += UM * Oi * Oj;
    const S dU_dr = dUM * Oi * Oj;
    const S dOi = S(2.0) * p.omega * gi * gie * Oi * Oi;
    const S dOj = S(2.0) * p.omega * gj * gje * Oj * Oj;
    const S dU_dgi = dOi * UM * Oj;
    const S dU_dgj = dOj * UM * Oi;

    const V3<S> nrh {-rh.x, -rh.y, -rh.z};
    const V3<S> rxi = v_cross(rh, ni);
    const V3<S> rxj = v_cross(rh, nj);
    const V3<S> nip = v_cross(nrh, rxi);
    const V3<S> njp = v_cross(nrh, rxj);

    // f = -dU_dr * rh - ri * (dU_dgi * nip + dU_dgj * njp), evaluated left to right (:189)
    const S m = -dU_dr;
    force.x = m * rh.x - ri * (dU_dgi * nip.x + dU_dgj * njp.x);
    force.y = m * rh.y - ri * (dU_dgi * nip.y + dU_dgj * njp.y);
    force.z = m * rh.z - ri * (dU_dgi * nip.z + dU_dgj * njp.z);
    tq_i = V3<S> {dU_dgi * rxi.x, dU_dgi * rxi.y, dU_dgi * rxi.z};
    tq_j = V3<S> {dU_dgj * rxj.x, dU_dgj * rxj.y, dU_dgj * rxj.z};

    if (shift)
        {
        const S rc = std::sqrt(rcutsq);
        const S mes = std::exp(-(rc - p.r_eq) * p.M_rinv);
        const S oms = S(1.0) - mes;
        const S UMs = p.M_d * (oms * oms - S(1.0));
        e -= UMs * Oi * Oj;
        }
    eng = e;
    return true;
    }
    } // namespace azp_oracle

#endif
