// TEST INFRASTRUCTURE (oracle). Not part of the product path: only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may build, load or call this library.
//
// One translation unit, four libraries (see oracle/Makefile):
//   liboracle_port_f32.so / _f64.so        : restated evaluators (port_evaluators.h)
//   _ref/liboracle_ref_f32.so / _f64.so    : -DORACLE_USE_REFERENCE, the reference's OWN evaluator
//                                            headers #included in place from /root/reference/src
//                                            (never copied) against oracle/hoomd_stub
// Both are driven by the same restated HOOMD host loops (driver_loops.h). All four export the same
// C symbols; oracle/oracle.py picks one by (kind, precision).
#include <cmath>
#include <cstring>

#include "driver_loops.h"
#include "nlist_cpu.h"

#ifdef ORACLE_USE_REFERENCE
#include "WallEvaluatorColloid.h"
#include "WallEvaluatorLJ93.h"
#include "PlanarBarrierEvaluator.h"
#include "SphericalBarrierEvaluator.h"
#include "AnisoPairEvaluatorTwoPatchMorse.h"
#include "DPDPairEvaluatorGeneralWeight.h"
#include "PairEvaluatorColloid.h"
#include "PairEvaluatorExpandedYukawa.h"
#include "PairEvaluatorHertz.h"
#include "PairEvaluatorPerturbedLennardJones.h"
typedef hoomd::Scalar S;
namespace ref = hoomd::azplugins::detail;
#else
#include "port_evaluators.h"
#ifndef HOOMD_LONGREAL_SIZE
#define HOOMD_LONGREAL_SIZE 64
#endif
#if HOOMD_LONGREAL_SIZE == 32
typedef float S;
#else
typedef double S;
#endif
#endif

using namespace azp_oracle;

// evaluator ids shared with include/azp_b200.h
enum
    {
    EV_PLJ = 0,
    EV_YUKAWA = 1,
    EV_COLLOID = 2,
    EV_HERTZ = 3,
    EV_DPD = 4,
    EV_MORSE = 5
    };

#ifdef ORACLE_USE_REFERENCE
template<class E> struct RefIso
    {
    typedef typename E::param_type param_type;
    static bool eval(S rsq, S rcutsq, const param_type& p, bool shift, S& fdr, S& eng)
        {
        E e(rsq, rcutsq, p);
        return e.evalForceAndEnergy(fdr, eng, shift);
        }
    };
typedef RefIso<ref::PairEvaluatorPerturbedLennardJones> AdPLJ;
typedef RefIso<ref::PairEvaluatorExpandedYukawa> AdYukawa;
typedef RefIso<ref::PairEvaluatorColloid> AdColloid;
typedef RefIso<ref::PairEvaluatorHertz> AdHertz;
struct AdDPD : public RefIso<ref::DPDPairEvaluatorGeneralWeight>
    {
    static bool eval_thermo(S rsq,
                            S rcutsq,
                            const param_type& p,
                            uint16_t seed,
                            unsigned int tag_i,
                            unsigned int tag_j,
                            unsigned int timestep,
                            S dt,
                            S rdotv,
                            S T,
                            S& fdr,
                            S& fdr_cons,
                            S& eng)
        {
        ref::DPDPairEvaluatorGeneralWeight e(rsq, rcutsq, p);
        e.set_seed_ij_timestep(seed, tag_i, tag_j, timestep);
        e.setDeltaT(dt);
        e.setRDotV(rdotv);
        e.setT(T);
        return e.evalForceEnergyThermo(fdr, fdr_cons, eng, false);
        }
    };
struct AdMorse
    {
    typedef ref::AnisoPairEvaluatorTwoPatchMorse::param_type param_type;
    static bool eval(const S dr[3],
                     const S qi[4],
                     const S qj[4],
                     S rcutsq,
                     const param_type& p,
                     bool shift,
                     S f[3],
                     S& eng,
                     S ti[3],
                     S tj[3])
        {
        hoomd::Scalar3 d = hoomd::make_scalar3(dr[0], dr[1], dr[2]);
        hoomd::Scalar4 a = hoomd::make_scalar4(qi[0], qi[1], qi[2], qi[3]);
        hoomd::Scalar4 b = hoomd::make_scalar4(qj[0], qj[1], qj[2], qj[3]);
        hoomd::Scalar3 F = hoomd::make_scalar3(0, 0, 0), Ti = F, Tj = F;
        ref::AnisoPairEvaluatorTwoPatchMorse e(d, a, b, rcutsq, p);
        const bool ok = e.evaluate(F, eng, shift, Ti, Tj);
        if (ok)
            {
            f[0] = F.x, f[1] = F.y, f[2] = F.z;
            ti[0] = Ti.x, ti[1] = Ti.y, ti[2] = Ti.z;
            tj[0] = Tj.x, tj[1] = Tj.y, tj[2] = Tj.z;
            }
        return ok;
        }
    };
#else
struct AdPLJ
    {
    typedef PLJParams<S> param_type;
    static bool eval(S rsq, S rcutsq, const param_type& p, bool shift, S& fdr, S& eng)
        {
        return eval_plj(rsq, rcutsq, p, shift, fdr, eng);
        }
    };
struct AdYukawa
    {
    typedef YukawaParams<S> param_type;
    static bool eval(S rsq, S rcutsq, const param_type& p, bool shift, S& fdr, S& eng)
        {
        return eval_yukawa(rsq, rcutsq, p, shift, fdr, eng);
        }
    };
struct AdColloid
    {
    typedef ColloidParams<S> param_type;
    static bool eval(S rsq, S rcutsq, const param_type& p, bool shift, S& fdr, S& eng)
        {
        return eval_colloid(rsq, rcutsq, p, shift, fdr, eng);
        }
    };
struct AdHertz
    {
    typedef HertzParams<S> param_type;
    static bool eval(S rsq, S rcutsq, const param_type& p, bool shift, S& fdr, S& eng)
        {
        return eval_hertz(rsq, rcutsq, p, shift, fdr, eng);
        }
    };
struct AdDPD
    {
    typedef DPDParams<S> param_type;
    static bool eval(S rsq, S rcutsq, const param_type& p, bool shift, S& fdr, S& eng)
        {
        return eval_dpd_conservative(rsq, rcutsq, p, shift, fdr, eng);
        }
    static bool eval_thermo(S rsq,
                            S rcutsq,
                            const param_type& p,
                            uint16_t seed,
                            unsigned int tag_i,
                            unsigned int tag_j,
                            unsigned int timestep,
                            S dt,
                            S rdotv,
                            S T,
                            S& fdr,
                            S& fdr_cons,
                            S& eng)
        {
        return eval_dpd_thermo(rsq, rcutsq, p, seed, tag_i, tag_j, timestep, dt, rdotv, T, fdr,
                               fdr_cons, eng);
        }
    };
struct AdMorse
    {
    typedef MorseParams<S> param_type;
    static bool eval(const S dr[3],
                     const S qi[4],
                     const S qj[4],
                     S rcutsq,
                     const param_type& p,
                     bool shift,
                     S f[3],
                     S& eng,
                     S ti[3],
                     S tj[3])
        {
        V3<S> d {dr[0], dr[1], dr[2]}, F {0, 0, 0}, Ti {0, 0, 0}, Tj {0, 0, 0};
        const bool ok = eval_morse(d, qi, qj, rcutsq, p, shift, F, eng, Ti, Tj);
        if (ok)
            {
            f[0] = F.x, f[1] = F.y, f[2] = F.z;
            ti[0] = Ti.x, ti[1] = Ti.y, ti[2] = Ti.z;
            tj[0] = Tj.x, tj[1] = Tj.y, tj[2] = Tj.z;
            }
        return ok;
        }
    };
#endif

// Plain-C view of PairArgs (doubles for scalars so one ctypes struct serves both precisions).
struct OracleArgs
    {
    uint32_t N;
    uint32_t ntypes;
    const void* pos;
    const uint32_t* n_neigh;
    const uint32_t* nlist;
    const uint64_t* head_list;
    double L[3];
    double tilt[3]; // xy, xz, yz
    int32_t periodic[3];
    int32_t shift_mode;
    int32_t compute_virial;
    int32_t half_list;
    int32_t rint_image;
    int32_t nthreads;
    const void* rcutsq;
    const void* ronsq;
    void* force;
    void* virial;
    uint64_t virial_pitch;
    const void* vel;
    const uint32_t* tag;
    uint32_t seed;
    uint64_t timestep;
    double deltaT;
    double T;
    const void* orientation;
    void* torque;
    };

static PairArgs<S> convert(const OracleArgs* o)
    {
    PairArgs<S> a;
    a.N = o->N;
    a.pos = static_cast<const S*>(o->pos);
    a.n_neigh = o->n_neigh;
    a.nlist = o->nlist;
    a.head_list = o->head_list;
    for (int d = 0; d < 3; ++d)
        {
        a.box.L[d] = S(o->L[d]);
        a.box.Linv[d] = S(1.0) / a.box.L[d];
        a.box.periodic[d] = o->periodic[d];
        }
    a.box.xy = S(o->tilt[0]);
    a.box.xz = S(o->tilt[1]);
    a.box.yz = S(o->tilt[2]);
    a.ntypes = o->ntypes;
    a.rcutsq = static_cast<const S*>(o->rcutsq);
    a.ronsq = static_cast<const S*>(o->ronsq);
    a.shift_mode = o->shift_mode;
    a.compute_virial = o->compute_virial;
    a.half_list = o->half_list;
    a.rint_image = o->rint_image;
    a.force = static_cast<S*>(o->force);
    a.virial = static_cast<S*>(o->virial);
    a.virial_pitch = o->virial_pitch;
    a.vel = static_cast<const S*>(o->vel);
    a.tag = o->tag;
    a.seed = uint16_t(o->seed);
    a.timestep = o->timestep;
    a.deltaT = S(o->deltaT);
    a.T = S(o->T);
    a.orientation = static_cast<const S*>(o->orientation);
    a.torque = static_cast<S*>(o->torque);
    a.nthreads = o->nthreads;
    return a;
    }

// ---- wall potentials -----------------------------------------------------------------------------
// Evaluators: the reference's own headers (ref build) or the restatement below (port build).
// The wall loop restates HOOMD's EvaluatorWalls / WallData (not in the reference tree): see
// azplugins_b200/csrc/wall_kernels.cu for the semantics kept.
#ifdef ORACLE_USE_REFERENCE
template<class T> struct WallLJ93
    {
    ref::WallParametersLJ93 p;
    WallLJ93(const T* q)
        {
        p.sigma_3 = q[0];
        p.A = q[1];
        }
    bool eval(T rsq, T rcutsq, T& fdr, T& eng) const
        {
        ref::WallEvaluatorLJ93 e(rsq, rcutsq, p);
        return e.evalForceAndEnergy(fdr, eng, true);
        }
    };
template<class T> struct WallColloid
    {
    ref::WallParametersColloid p;
    WallColloid(const T* q)
        {
        p.c_1 = q[0];
        p.c_2 = q[1];
        p.a = q[2];
        }
    bool eval(T rsq, T rcutsq, T& fdr, T& eng) const
        {
        ref::WallEvaluatorColloid e(rsq, rcutsq, p);
        return e.evalForceAndEnergy(fdr, eng, true);
        }
    };
#else
// restated from src/WallEvaluatorLJ93.h:93-132 and src/WallEvaluatorColloid.h:104-175
template<class T> struct WallLJ93
    {
    T lj1, lj2;
    WallLJ93(const T* q)
        {
        lj1 = (T(2.0) / T(15.0)) * q[1] * q[0] * q[0] * q[0];
        lj2 = q[1] * q[0];
        }
    bool eval(T rsq, T rcutsq, T& fdr, T& eng) const
        {
        if (!(rsq < rcutsq && lj1 != 0))
            return false;
        T r2inv = T(1.0) / rsq;
        T r3inv = r2inv * std::sqrt(r2inv);
        T r6inv = r3inv * r3inv;
        fdr = r2inv * r3inv * (T(9.0) * lj1 * r6inv - T(3.0) * lj2);
        eng = r3inv * (lj1 * r6inv - lj2);
        T rcut2inv = T(1.0) / rcutsq;
        T rcut3inv = rcut2inv * std::sqrt(rcut2inv);
        T rcut6inv = rcut3inv * rcut3inv;
        eng -= rcut3inv * (lj1 * rcut6inv - lj2);
        return true;
        }
    };
template<class T> struct WallColloid
    {
    T c_1, c_2, a;
    WallColloid(const T* q) : c_1(q[0]), c_2(q[1]), a(q[2]) { }
    template<bool force> T potential(T& fdr, T rsq) const
        {
        T r = std::sqrt(rsq);
        T arinv = a / r;
        T rma = T(1.0) / (r - a);
        T rpa = T(1.0) / (r + a);
        T r2a2 = rma * rpa;
        T rma2 = rma * rma;
        T rma6 = rma2 * rma2 * rma2;
        T rpa2 = rpa * rpa;
        T rpa6 = rpa2 * rpa2 * rpa2;
        if (force)
            {
            T arinv8 = T(8.0) * arinv;
            fdr = T(6.0) * c_1 * ((arinv8 - T(1.0)) * rma2 * rma6 + (arinv8 + T(1.0)) * rpa2 * rpa6);
            fdr -= c_2 * (T(4.0) * a * a * arinv * r2a2 * r2a2);
            }
        T a7 = T(7.0) * a;
        T energy = c_1 * ((a7 - r) * rma * rma6 + (a7 + r) * rpa * rpa6);
        energy -= c_2 * (T(2.0) * a * r * r2a2 + std::log(rpa / rma));
        return energy;
        }
    bool eval(T rsq, T rcutsq, T& fdr, T& eng) const
        {
        if (!(rsq < rcutsq && c_1 != 0 && a > 0))
            return false;
        eng = potential<true>(fdr, rsq);
        T unused;
        eng -= potential<false>(unused, rcutsq);
        return true;
        }
    };
#endif

template<class T, class E>
static void wall_loop(uint32_t N, const T* pos, uint32_t nparam, const T* params, uint32_t ns, const double* sph, uint32_t nc, const double* cyl, uint32_t np, const double* pla, T* force, T* virial, uint64_t pitch)
    {
    auto side = [](T d, T r, bool inside, bool open)
    { return open ? ((d < r && inside) || (d > r && !inside)) : ((d <= r && inside) || (d >= r && !inside)); };
    for (uint32_t i = 0; i < N; ++i)
        {
        const T x = pos[4 * i], y = pos[4 * i + 1], z = pos[4 * i + 2];
        uint32_t type;
        if (sizeof(T) == 4)
            memcpy(&type, &pos[4 * i + 3], 4);
        else
            {
            uint64_t t64;
            memcpy(&t64, &pos[4 * i + 3], 8);
            type = (uint32_t)t64;
            }
        const T* q = params + (size_t)nparam * type;
        const E ev(q);
        const T rcutsq = q[nparam - 2];
        T fx = 0, fy = 0, fz = 0, energy = 0;
        const T rextrap = q[nparam - 1];
        const bool extrap = rextrap > 0;
        auto add_wall = [&](T dx, T dy, T dz)
        {
            const T rx = -dx, ry = -dy, rz = -dz;
            const T rsq = rx * rx + ry * ry + rz * rz;
            T fdr = 0, eng = 0;
            if (ev.eval(rsq, rcutsq, fdr, eng))
                {
                if (!std::isfinite(fdr))
                    fdr = 0, eng = 0;
                fx += rx * fdr, fy += ry * fdr, fz += rz * fdr;
                energy += eng;
                }
        };
        // HOOMD's extrapolated mode (restated): see wall_kernels.cu add_wall_extrap
        auto add_wall_extrap = [&](bool in_active, T dx, T dy, T dz, T nx, T ny, T nz)
        {
            const T rextrapsq = rextrap * rextrap;
            const T rsq = dx * dx + dy * dy + dz * dz;
            if (in_active && rsq >= rextrapsq)
                {
                add_wall(dx, dy, dz);
                return;
                }
            T r = std::sqrt(rsq);
            if (rsq == 0)
                {
                in_active = true;
                dx = nx, dy = ny, dz = nz;
                }
            else
                {
                const T rinv = T(1.0) / r;
                dx *= rinv, dy *= rinv, dz *= rinv;
                }
            r = in_active ? rextrap - r : rextrap + r;
            const T scale = in_active ? rextrap : -rextrap;
            dx *= scale, dy *= scale, dz *= scale;
            T fdr = 0, eng = 0;
            if (ev.eval(rextrapsq, rcutsq, fdr, eng))
                {
                eng = eng + fdr * rextrap * r;
                energy += eng;
                fx += -dx * fdr, fy += -dy * fdr, fz += -dz * fdr;
                }
        };
        for (uint32_t k = 0; k < ns; ++k)
            {
            const double* w = sph + 6 * k;
            const T r = T(w[0]), tx = x - T(w[1]), ty = y - T(w[2]), tz = z - T(w[3]);
            const bool inside = w[4] != 0, open = w[5] != 0;
            const T rxyz = std::sqrt(tx * tx + ty * ty + tz * tz);
            if (rxyz > 0)
                {
                const bool act = side(rxyz, r, inside, open);
                const T s = r / rxyz - T(1.0);
                if (extrap)
                    {
                    const T inv = (inside ? T(1) : T(-1)) / rxyz;
                    add_wall_extrap(act, s * tx, s * ty, s * tz, tx * inv, ty * inv, tz * inv);
                    }
                else if (act)
                    add_wall(s * tx, s * ty, s * tz);
                }
            else if (extrap)
                add_wall_extrap(inside, r, 0, 0, inside ? T(1) : T(-1), 0, 0);
            else if (inside)
                add_wall(r, 0, 0);
            }
        for (uint32_t k = 0; k < nc; ++k)
            {
            const double* w = cyl + 9 * k;
            const T r = T(w[0]), tx = x - T(w[1]), ty = y - T(w[2]), tz = z - T(w[3]);
            const T ax = T(w[4]), ay = T(w[5]), az = T(w[6]);
            const bool inside = w[7] != 0, open = w[8] != 0;
            const T along = tx * ax + ty * ay + tz * az;
            const T qx = tx - along * ax, qy = ty - along * ay, qz = tz - along * az;
            const T rxy = std::sqrt(qx * qx + qy * qy + qz * qz);
            if (rxy > 0)
                {
                const bool act = side(rxy, r, inside, open);
                const T s = r / rxy - T(1.0);
                if (extrap)
                    {
                    const T inv = (inside ? T(1) : T(-1)) / rxy;
                    add_wall_extrap(act, s * qx, s * qy, s * qz, qx * inv, qy * inv, qz * inv);
                    }
                else if (act)
                    add_wall(s * qx, s * qy, s * qz);
                }
            else if (inside)
                {
                T ux = 1, uy = 0, uz = 0;
                if (std::fabs(ax) > T(0.9))
                    ux = 0, uy = 1;
                const T d = ux * ax + uy * ay + uz * az;
                ux -= d * ax, uy -= d * ay, uz -= d * az;
                const T n = r / std::sqrt(ux * ux + uy * uy + uz * uz);
                if (extrap)
                    add_wall_extrap(true, n * ux, n * uy, n * uz, 0, 0, 0);
                else
                    add_wall(n * ux, n * uy, n * uz);
                }
            }
        for (uint32_t k = 0; k < np; ++k)
            {
            const double* w = pla + 7 * k;
            const T ox = T(w[0]), oy = T(w[1]), oz = T(w[2]), nx = T(w[3]), ny = T(w[4]), nz = T(w[5]);
            const bool open = w[6] != 0;
            const T d = (nx * x + ny * y + nz * z) - (nx * ox + ny * oy + nz * oz);
            const bool act = open ? (d > 0) : (d >= 0);
            if (extrap)
                add_wall_extrap(act, -d * nx, -d * ny, -d * nz, -nx, -ny, -nz);
            else if (act)
                add_wall(-d * nx, -d * ny, -d * nz);
            }
        force[4 * i] = fx, force[4 * i + 1] = fy, force[4 * i + 2] = fz, force[4 * i + 3] = energy;
        if (virial)
            {
            virial[0 * pitch + i] = fx * x;
            virial[1 * pitch + i] = fx * y;
            virial[2 * pitch + i] = fx * z;
            virial[3 * pitch + i] = fy * y;
            virial[4 * pitch + i] = fy * z;
            virial[5 * pitch + i] = fz * z;
            }
        }
    }

// ---- external harmonic barrier -------------------------------------------------------------------
#ifdef ORACLE_USE_REFERENCE
template<class T>
static int barrier_loop(int geometry, T location, uint32_t N, const T* pos, uint32_t ntypes, const T* params, const double* L, const double* tilt, const int32_t* periodic, T* force, T* virial, uint64_t pitch, bool check_valid)
    {
    hoomd::BoxDim box((T)L[0], (T)L[1], (T)L[2], (T)tilt[0], (T)tilt[1], (T)tilt[2]);
    box.setPeriodic(periodic[0], periodic[1], periodic[2]);
    auto run = [&](auto evaluator) -> int
    {
        if (check_valid && !evaluator.valid(box))
            return 1;
        for (uint32_t i = 0; i < N; ++i)
            {
            hoomd::Scalar3 p = hoomd::make_scalar3(pos[4 * i], pos[4 * i + 1], pos[4 * i + 2]);
            uint32_t type;
            if (sizeof(T) == 4)
                memcpy(&type, &pos[4 * i + 3], 4);
            else
                {
                uint64_t t64;
                memcpy(&t64, &pos[4 * i + 3], 8);
                type = (uint32_t)t64;
                }
            hoomd::int3 img = hoomd::make_int3(0, 0, 0);
            box.wrap(p, img);
            const hoomd::Scalar4 f = evaluator(p, params[2 * type], params[2 * type + 1]);
            force[4 * i] = f.x, force[4 * i + 1] = f.y, force[4 * i + 2] = f.z, force[4 * i + 3] = f.w;
            }
        if (virial)
            for (int r = 0; r < 6; ++r)
                for (uint32_t i = 0; i < N; ++i)
                    virial[r * pitch + i] = T(0);
        return 0;
    };
    if (geometry == 0)
        return run(hoomd::azplugins::PlanarBarrierEvaluator(location));
    return run(hoomd::azplugins::SphericalBarrierEvaluator(location));
    }
#else
// own restatement of src/PlanarBarrierEvaluator.h:37-51,54-59, src/SphericalBarrierEvaluator.h:36-53,
// 56-62 and of HOOMD's BoxDim::wrap / makeCoordinates / getNearestPlaneDistance
template<class T>
static int barrier_loop(int geometry, T location, uint32_t N, const T* pos, uint32_t ntypes, const T* params, const double* L, const double* tilt, const int32_t* periodic, T* force, T* virial, uint64_t pitch, bool check_valid)
    {
    const T Lx = T(L[0]), Ly = T(L[1]), Lz = T(L[2]);
    const T xy = T(tilt[0]), xz = T(tilt[1]), yz = T(tilt[2]);
    const T lox = -Lx / T(2.0), loy = -Ly / T(2.0), loz = -Lz / T(2.0);
    const T hix = lox + Lx, hiy = loy + Ly, hiz = loz + Lz;
    if (check_valid)
        {
        if (geometry == 0)
            {
            const T lo = loy + yz * loz, hi = hiy + yz * hiz;
            if (!(location >= lo && location < hi))
                return 1;
            }
        else
            {
            const T term = xy * yz - xz;
            const T dx = Lx / std::sqrt(T(1.0) + xy * xy + term * term);
            const T dy = Ly / std::sqrt(T(1.0) + yz * yz);
            const T two_R = T(2.0) * location;
            if (!(location >= T(0.0) && dx >= two_R && dy >= two_R && Lz >= two_R))
                return 1;
            }
        }
    for (uint32_t i = 0; i < N; ++i)
        {
        T x = pos[4 * i], y = pos[4 * i + 1], z = pos[4 * i + 2];
        uint32_t type;
        if (sizeof(T) == 4)
            memcpy(&type, &pos[4 * i + 3], 4);
        else
            {
            uint64_t t64;
            memcpy(&t64, &pos[4 * i + 3], 8);
            type = (uint32_t)t64;
            }
        if (periodic[2])
            {
            if (z >= hiz)
                z -= Lz, y -= Lz * yz, x -= Lz * xz;
            else if (z < loz)
                z += Lz, y += Lz * yz, x += Lz * xz;
            }
        if (periodic[1])
            {
            const T ty = yz * z;
            if (y >= hiy + ty)
                y -= Ly, x -= Ly * xy;
            else if (y < loy + ty)
                y += Ly, x += Ly * xy;
            }
        if (periodic[0])
            {
            const T tx = (xz - xy * yz) * z + xy * y;
            if (x >= hix + tx)
                x -= Lx;
            else if (x < lox + tx)
                x += Lx;
            }
        const T k = params[2 * type], offset = params[2 * type + 1];
        T fx = 0, fy = 0, fz = 0, e = 0;
        if (geometry == 0)
            {
            const T dy = y - (location + offset);
            if (dy > T(0.0))
                {
                const T f = -k * dy;
                fy = f;
                e = T(-0.5) * f * dy;
                }
            }
        else
            {
            const T r = std::sqrt(x * x + y * y + z * z);
            const T dr = r - (location + offset);
            if (dr > T(0.0))
                {
                const T k_dr = k * dr;
                const T c = -(k_dr / r);
                fx = c * x, fy = c * y, fz = c * z;
                e = T(0.5) * k_dr * dr;
                }
            }
        force[4 * i] = fx, force[4 * i + 1] = fy, force[4 * i + 2] = fz, force[4 * i + 3] = e;
        }
    if (virial)
        for (int r = 0; r < 6; ++r)
            for (uint32_t i = 0; i < N; ++i)
                virial[r * pitch + i] = T(0);
    return 0;
    }
#endif

extern "C"
    {
    int oracle_scalar_size()
        {
        return int(sizeof(S));
        }

    int oracle_is_reference()
        {
#ifdef ORACLE_USE_REFERENCE
        return 1;
#else
        return 0;
#endif
        }

    int oracle_max_threads()
        {
#ifdef _OPENMP
        return omp_get_max_threads();
#else
        return 1;
#endif
        }

    int oracle_param_size(int ev)
        {
        switch (ev)
            {
        case EV_PLJ:
            return int(sizeof(AdPLJ::param_type));
        case EV_YUKAWA:
            return int(sizeof(AdYukawa::param_type));
        case EV_COLLOID:
            return int(sizeof(AdColloid::param_type));
        case EV_HERTZ:
            return int(sizeof(AdHertz::param_type));
        case EV_DPD:
            return int(sizeof(AdDPD::param_type));
        case EV_MORSE:
            return int(sizeof(AdMorse::param_type));
            }
        return -1;
        }

    // isotropic PotentialPair loop; ev = EV_DPD gives PotentialPairConservative<GeneralWeight>
    int oracle_pair_forces(int ev, const OracleArgs* o, const void* params)
        {
        PairArgs<S> a = convert(o);
        switch (ev)
            {
        case EV_PLJ:
            iso_loop<S, AdPLJ>(a, params);
            return 0;
        case EV_YUKAWA:
            iso_loop<S, AdYukawa>(a, params);
            return 0;
        case EV_COLLOID:
            iso_loop<S, AdColloid>(a, params);
            return 0;
        case EV_HERTZ:
            iso_loop<S, AdHertz>(a, params);
            return 0;
        case EV_DPD:
            iso_loop<S, AdDPD>(a, params);
            return 0;
            }
        return 1;
        }

    int oracle_dpd_forces(int ev, const OracleArgs* o, const void* params)
        {
        if (ev != EV_DPD)
            return 1;
        PairArgs<S> a = convert(o);
        dpd_loop<S, AdDPD>(a, params);
        return 0;
        }

    int oracle_aniso_forces(int ev, const OracleArgs* o, const void* params)
        {
        if (ev != EV_MORSE)
            return 1;
        PairArgs<S> a = convert(o);
        aniso_loop<S, AdMorse>(a, params);
        return 0;
        }

    // single-pair probes for the known-answer tests; out = {evaluated, force_divr, pair_eng}
    int oracle_eval_pair(int ev, const void* param, double rsq, double rcutsq, int shift, double* out)
        {
        S fdr = 0, eng = 0;
        bool ok = false;
        switch (ev)
            {
        case EV_PLJ:
            ok = AdPLJ::eval(S(rsq), S(rcutsq), *static_cast<const AdPLJ::param_type*>(param),
                             shift != 0, fdr, eng);
            break;
        case EV_YUKAWA:
            ok = AdYukawa::eval(S(rsq), S(rcutsq),
                                *static_cast<const AdYukawa::param_type*>(param), shift != 0, fdr,
                                eng);
            break;
        case EV_COLLOID:
            ok = AdColloid::eval(S(rsq), S(rcutsq),
                                 *static_cast<const AdColloid::param_type*>(param), shift != 0,
                                 fdr, eng);
            break;
        case EV_HERTZ:
            ok = AdHertz::eval(S(rsq), S(rcutsq), *static_cast<const AdHertz::param_type*>(param),
                               shift != 0, fdr, eng);
            break;
        case EV_DPD:
            ok = AdDPD::eval(S(rsq), S(rcutsq), *static_cast<const AdDPD::param_type*>(param),
                             shift != 0, fdr, eng);
            break;
        default:
            return 1;
            }
        out[0] = ok ? 1.0 : 0.0;
        out[1] = double(fdr);
        out[2] = double(eng);
        return 0;
        }

    // out = {evaluated, fdr, fdr_cons, eng}
    int oracle_eval_dpd_thermo(const void* param,
                               double rsq,
                               double rcutsq,
                               uint32_t seed,
                               uint32_t tag_i,
                               uint32_t tag_j,
                               uint64_t timestep,
                               double dt,
                               double rdotv,
                               double T,
                               double* out)
        {
        S fdr = 0, fc = 0, eng = 0;
        const bool ok = AdDPD::eval_thermo(S(rsq), S(rcutsq),
                                           *static_cast<const AdDPD::param_type*>(param),
                                           uint16_t(seed), tag_i, tag_j, (unsigned int)timestep,
                                           S(dt), S(rdotv), S(T), fdr, fc, eng);
        out[0] = ok ? 1.0 : 0.0;
        out[1] = double(fdr);
        out[2] = double(fc);
        out[3] = double(eng);
        return 0;
        }

    // out = {evaluated, fx, fy, fz, eng, tix, tiy, tiz, tjx, tjy, tjz}
    int oracle_eval_aniso(const void* param,
                          const double* dr,
                          const double* qi,
                          const double* qj,
                          double rcutsq,
                          int shift,
                          double* out)
        {
        S d[3] = {S(dr[0]), S(dr[1]), S(dr[2])};
        S a[4] = {S(qi[0]), S(qi[1]), S(qi[2]), S(qi[3])};
        S b[4] = {S(qj[0]), S(qj[1]), S(qj[2]), S(qj[3])};
        S f[3] = {0, 0, 0}, ti[3] = {0, 0, 0}, tj[3] = {0, 0, 0}, eng = 0;
        const bool ok = AdMorse::eval(d, a, b, S(rcutsq),
                                      *static_cast<const AdMorse::param_type*>(param), shift != 0,
                                      f, eng, ti, tj);
        out[0] = ok ? 1.0 : 0.0;
        for (int c = 0; c < 3; ++c)
            {
            out[1 + c] = double(f[c]);
            out[5 + c] = double(ti[c]);
            out[8 + c] = double(tj[c]);
            }
        out[4] = double(eng);
        return 0;
        }

    // uniform(-1,1) the DPD evaluator draws for (seed, tags, timestep): isolates the RNG packing
    double oracle_dpd_alpha(uint32_t seed, uint32_t tag_i, uint32_t tag_j, uint64_t timestep)
        {
#ifdef ORACLE_USE_REFERENCE
        unsigned int lo = tag_i > tag_j ? tag_j : tag_i, hi = tag_i > tag_j ? tag_i : tag_j;
        hoomd::RandomGenerator rng(
            hoomd::Seed(ref::RNGIdentifier::DPDEvaluatorGeneralWeight, (unsigned int)timestep,
                        uint16_t(seed)),
            hoomd::Counter(lo, hi));
        return double(hoomd::UniformDistribution<S>(-1, 1)(rng));
#else
        return double(dpd_alpha<S>(uint16_t(seed), tag_i, tag_j, (unsigned int)timestep));
#endif
        }

    void oracle_philox(const uint32_t* ctr, const uint32_t* key, uint32_t* out)
        {
#ifdef ORACLE_USE_REFERENCE
        hoomd::detail::philox_u4 c;
        hoomd::detail::philox_u2 k;
        for (int i = 0; i < 4; ++i)
            c.v[i] = ctr[i];
        k.v[0] = key[0], k.v[1] = key[1];
        c = hoomd::detail::philox4x32_10(c, k);
        for (int i = 0; i < 4; ++i)
            out[i] = c.v[i];
#else
        philox4x32_10(ctr, key, out);
#endif
        }

    void oracle_min_image(const double* L, const double* tilt, const int32_t* periodic, int rint_mode,
                          double* v)
        {
        Box<S> b;
        for (int d = 0; d < 3; ++d)
            {
            b.L[d] = S(L[d]);
            b.Linv[d] = S(1.0) / b.L[d];
            b.periodic[d] = periodic[d];
            }
        b.xy = S(tilt[0]), b.xz = S(tilt[1]), b.yz = S(tilt[2]);
        S x = S(v[0]), y = S(v[1]), z = S(v[2]);
        if (rint_mode)
            min_image_rint(b, x, y, z);
        else
            min_image_host(b, x, y, z);
        v[0] = x, v[1] = y, v[2] = z;
        }

    // Wall potentials (evaluator 0 = Colloid {c_1, c_2, a, rcutsq, rextrap}, 1 = LJ93 {sigma_3, A,
    // rcutsq, rextrap} per type); walls as double arrays: spheres [r, o3, inside, open],
    // cylinders [r, o3, axis3 (unit), inside, open], planes [o3, n3 (unit), open]
    int oracle_wall(int evaluator,
                    uint32_t N,
                    const void* pos,
                    const void* params,
                    uint32_t ns,
                    const double* sph,
                    uint32_t nc,
                    const double* cyl,
                    uint32_t np,
                    const double* pla,
                    void* force,
                    void* virial,
                    uint64_t pitch)
        {
        if (evaluator == 0)
            wall_loop<S, WallColloid<S>>(N, static_cast<const S*>(pos), 5, static_cast<const S*>(params), ns, sph, nc, cyl, np, pla, static_cast<S*>(force), static_cast<S*>(virial), pitch);
        else if (evaluator == 1)
            wall_loop<S, WallLJ93<S>>(N, static_cast<const S*>(pos), 4, static_cast<const S*>(params), ns, sph, nc, cyl, np, pla, static_cast<S*>(force), static_cast<S*>(virial), pitch);
        else
            return 1;
        return 0;
        }

    // External harmonic barrier: restatement of HarmonicBarrier<Evaluator>::computeForces
    // (reference src/HarmonicBarrier.h:149-175): per particle wrap into the global box, evaluate,
    // overwrite force; the virial is zeroed. geometry 0 = planar (y = location), 1 = spherical.
    // Returns 0, or 1 when the barrier position is invalid (reference :126-130 throws).
    int oracle_barrier(int geometry,
                       double location,
                       uint32_t N,
                       const void* pos_,
                       uint32_t ntypes,
                       const void* params_,
                       const double* L,
                       const double* tilt,
                       const int32_t* periodic,
                       void* force_,
                       void* virial_,
                       uint64_t virial_pitch,
                       int check_valid)
        {
        const S* pos = static_cast<const S*>(pos_);
        const S* params = static_cast<const S*>(params_); // {k, offset} per type
        S* force = static_cast<S*>(force_);
        S* virial = static_cast<S*>(virial_);
        return barrier_loop<S>(geometry, S(location), N, pos, ntypes, params, L, tilt, periodic, force, virial, virial_pitch, check_valid != 0);
        }

    // HOOMD-layout neighbour list on the CPU (nlist_cpu.h); pass 0 counts, pass 1 fills
    int oracle_nlist(int pass,
                     uint32_t N,
                     const void* pos,
                     const double* L,
                     const double* tilt,
                     const int32_t* periodic,
                     uint32_t ntypes,
                     const void* rlistsq,
                     int half,
                     uint32_t* n_neigh,
                     const uint64_t* head_list,
                     uint32_t* nlist,
                     int nthreads,
                     uint32_t n_rows)
        {
        Box<S> b;
        for (int d = 0; d < 3; ++d)
            {
            b.L[d] = S(L[d]);
            b.Linv[d] = S(1.0) / b.L[d];
            b.periodic[d] = periodic[d];
            }
        b.xy = S(tilt[0]), b.xz = S(tilt[1]), b.yz = S(tilt[2]);
        nlist_pass<S>(pass, N, static_cast<const S*>(pos), b, ntypes,
                      static_cast<const S*>(rlistsq), half, n_neigh, head_list, nlist, nthreads, n_rows);
        return 0;
        }

#ifndef ORACLE_USE_REFERENCE
    // restated param_type(dict) constructors: fields in the order documented per evaluator in
    // port_evaluators.h (the order of include/azp_b200.h's azp_param_pack)
    int oracle_pack_params(int ev, const double* f, void* out)
        {
        switch (ev)
            {
        case EV_PLJ:
            pack_plj<S>(f, static_cast<PLJParams<S>*>(out));
            return 0;
        case EV_YUKAWA:
            pack_yukawa<S>(f, static_cast<YukawaParams<S>*>(out));
            return 0;
        case EV_COLLOID:
            pack_colloid<S>(f, static_cast<ColloidParams<S>*>(out));
            return 0;
        case EV_HERTZ:
            pack_hertz<S>(f, static_cast<HertzParams<S>*>(out));
            return 0;
        case EV_DPD:
            pack_dpd<S>(f, static_cast<DPDParams<S>*>(out));
            return 0;
        case EV_MORSE:
            pack_morse<S>(f, static_cast<MorseParams<S>*>(out));
            return 0;
            }
        return 1;
        }
#endif
    } // extern "C"

#ifdef ORACLE_USE_REFERENCE
// param_type(pybind11::dict) / asDict() / toPython() of the reference itself. Call through
// ctypes.PyDLL (GIL held). `out` must be zero-initialised by the caller (padding bytes).
#include <new>
template<class P> static int pack_with_ctor(PyObject* d, void* out)
    {
    try
        {
        pybind11::dict v = pybind11::reinterpret_borrow<pybind11::dict>(d);
        new (out) P(v, false);
        return 0;
        }
    catch (const std::exception& e)
        {
        PyErr_SetString(PyExc_RuntimeError, e.what());
        return 1;
        }
    }
extern "C"
    {
    int oracle_ref_pack_params(int ev, PyObject* d, void* out)
        {
        switch (ev)
            {
        case EV_PLJ:
            return pack_with_ctor<AdPLJ::param_type>(d, out);
        case EV_YUKAWA:
            return pack_with_ctor<AdYukawa::param_type>(d, out);
        case EV_COLLOID:
            return pack_with_ctor<AdColloid::param_type>(d, out);
        case EV_HERTZ:
            return pack_with_ctor<AdHertz::param_type>(d, out);
        case EV_DPD:
            return pack_with_ctor<AdDPD::param_type>(d, out);
        case EV_MORSE:
            return pack_with_ctor<AdMorse::param_type>(d, out);
            }
        return 1;
        }

    PyObject* oracle_ref_unpack_params(int ev, void* in)
        {
        pybind11::object o;
        switch (ev)
            {
        case EV_PLJ:
            o = static_cast<AdPLJ::param_type*>(in)->asDict();
            break;
        case EV_YUKAWA:
            o = static_cast<AdYukawa::param_type*>(in)->asDict();
            break;
        case EV_COLLOID:
            o = static_cast<AdColloid::param_type*>(in)->asDict();
            break;
        case EV_HERTZ:
            o = static_cast<AdHertz::param_type*>(in)->asDict();
            break;
        case EV_DPD:
            o = static_cast<AdDPD::param_type*>(in)->asDict();
            break;
        case EV_MORSE:
            o = static_cast<AdMorse::param_type*>(in)->toPython();
            break;
        default:
            o = pybind11::none();
            }
        return o.release().ptr();
        }

    // getName() strings of the reference evaluators (pair-class naming)
    PyObject* oracle_ref_name(int ev)
        {
        std::string s;
        switch (ev)
            {
        case EV_PLJ:
            s = ref::PairEvaluatorPerturbedLennardJones::getName();
            break;
        case EV_YUKAWA:
            s = ref::PairEvaluatorExpandedYukawa::getName();
            break;
        case EV_COLLOID:
            s = ref::PairEvaluatorColloid::getName();
            break;
        case EV_HERTZ:
            s = ref::PairEvaluatorHertz::getName();
            break;
        case EV_DPD:
            s = ref::DPDPairEvaluatorGeneralWeight::getName();
            break;
        case EV_MORSE:
            s = ref::AnisoPairEvaluatorTwoPatchMorse::getName();
            break;
            }
        return pybind11::str(s).release().ptr();
        }
    }
#endif
