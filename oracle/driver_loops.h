// TEST INFRASTRUCTURE (oracle). Not part of the product path.
//
// CPU restatement of the three HOOMD-blue v7.0.1 host force loops the azplugins pair classes
// instantiate (reference src/export_PotentialPair.cc.inc:13,28,
// src/export_PotentialPairDPDThermo.cc.inc:23-24, src/export_AnisoPotentialPair.cc.inc:9,24):
//   md::PotentialPair<E>::computeForces            -> iso_loop
//   md::PotentialPairDPDThermo<E>::computeForces   -> dpd_loop
//   md::AnisoPotentialPair<E>::computeForces       -> aniso_loop
// HOOMD's sources are NOT in /root/reference (un-vendored dependency, pinned v7.0.1), so these
// follow the published algorithm as summarised in SURVEY.md Appendix A.3/A.4/A.6/A.7. The loops
// are templated on an adapter so the very same loop drives (a) the reference's own evaluator
// classes compiled in place (oracle_main.cc with -DORACLE_USE_REFERENCE) and (b) the restated
// evaluators of port_evaluators.h.
//
// Full list (what the GPU classes use): each row owns its outputs, OpenMP over i is allowed.
// Half list (what HOOMD's CPU classes use): Newton's third law to j < N, single thread.
// half_list == 2, "domains" (the CPU baseline of bench.py): HOOMD's CPU parallelism is MPI domain
// decomposition -- every rank runs the half-list loop over its own particles plus ghosts, so a
// pair inside a domain is evaluated once and a pair across a domain face once on EACH side.
// Restated with one OpenMP thread per domain over a FULL list: thread d owns the contiguous rows
// [N d / D, N (d+1) / D) (a compact blob of the Morton-sorted particles); a neighbour j inside
// the domain is taken only when j > i and receives the reaction, a neighbour outside is a
// "ghost": evaluated, nothing written to it. No two threads write the same particle.
#ifndef AZP_ORACLE_DRIVER_LOOPS_H_
#define AZP_ORACLE_DRIVER_LOOPS_H_

#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace azp_oracle
    {
template<class S> struct Box
    {
    S L[3];
    S Linv[3];
    S xy, xz, yz;
    int periodic[3];
    };

// HOOMD BoxDim::minImage, host branch: compare against +-L/2 (boxes are
// centred, lo = -L/2, hi = L/2); z wraps once, y and x wrap by an integer multiple so that the
// shifts induced by the tilt factors are undone too.
template<class S> inline void min_image_host(const Box<S>& b, S& x, S& y, S& z)
    {
    const S hx = b.L[0] / S(2), hy = b.L[1] / S(2), hz = b.L[2] / S(2);
    if (b.periodic[2])
        {
        if (z >= hz)
            {
            z -= b.L[2];
            y -= b.L[2] * b.yz;
            x -= b.L[2] * b.xz;
            }
        else if (z < -hz)
            {
            z += b.L[2];
            y += b.L[2] * b.yz;
            x += b.L[2] * b.xz;
            }
        }
    if (b.periodic[1])
        {
        if (y >= hy)
            {
            const int i = int(y * b.Linv[1] + S(0.5));
            y -= S(i) * b.L[1];
            x -= S(i) * b.L[1] * b.xy;
            }
        else if (y < -hy)
            {
            const int i = int(-y * b.Linv[1] + S(0.5));
            y += S(i) * b.L[1];
            x += S(i) * b.L[1] * b.xy;
            }
        }
    if (b.periodic[0])
        {
        if (x >= hx)
            {
            const int i = int(x * b.Linv[0] + S(0.5));
            x -= S(i) * b.L[0];
            }
        else if (x < -hx)
            {
            const int i = int(-x * b.Linv[0] + S(0.5));
            x += S(i) * b.L[0];
            }
        }
    }

// HOOMD BoxDim::minImage, device branch: img = rint(w * Linv), z then y then x.
template<class S> inline void min_image_rint(const Box<S>& b, S& x, S& y, S& z)
    {
    if (b.periodic[2])
        {
        const S img = std::rint(z * b.Linv[2]);
        z -= b.L[2] * img;
        y -= b.L[2] * b.yz * img;
        x -= b.L[2] * b.xz * img;
        }
    if (b.periodic[1])
        {
        const S img = std::rint(y * b.Linv[1]);
        y -= b.L[1] * img;
        x -= b.L[1] * b.xy * img;
        }
    if (b.periodic[0])
        {
        const S img = std::rint(x * b.Linv[0]);
        x -= b.L[0] * img;
        }
    }

template<class S> struct PairArgs
    {
    unsigned int N;             // local particles (rows)
    const S* pos;               // Scalar4[N + ghosts]: x, y, z, type bit-cast in w
    const unsigned int* n_neigh;
    const unsigned int* nlist;
    const uint64_t* head_list;
    Box<S> box;
    unsigned int ntypes;
    const S* rcutsq;            // [ntypes^2], Index2D(i,j) = j*ntypes + i
    const S* ronsq;             // [ntypes^2] (iso only)
    int shift_mode;             // 0 none, 1 shift, 2 xplor
    int compute_virial;
    int half_list;              // 1: apply third law to j < N (HOOMD CPU storage mode)
    int rint_image;             // 1: device-style minImage
    S* force;                   // Scalar4[N]: fx, fy, fz, energy
    S* virial;                  // [6 * virial_pitch]
    size_t virial_pitch;
    // DPD
    const S* vel;               // Scalar4: vx, vy, vz, mass
    const unsigned int* tag;
    uint16_t seed;
    uint64_t timestep;
    S deltaT, T;
    // aniso
    const S* orientation;       // Scalar4 (s, x, y, z)
    S* torque;                  // Scalar4[N]
    int nthreads;
    };

template<class S> inline unsigned int type_of(const S* pos4)
    {
    unsigned int t;
    if (sizeof(S) == 4)
        std::memcpy(&t, pos4 + 3, 4);
    else
        {
        // __scalar_as_int on a double build takes the low word pair semantics of
        // __double_as_longlong narrowed to int
        int64_t tt;
        std::memcpy(&tt, pos4 + 3, 8);
        t = (unsigned int)tt;
        }
    return t;
    }

template<class S> inline void zero_outputs(const PairArgs<S>& a)
    {
    std::memset(a.force, 0, sizeof(S) * 4 * a.N);
    if (a.torque)
        std::memset(a.torque, 0, sizeof(S) * 4 * a.N);
    if (a.virial)
        std::memset(a.virial, 0, sizeof(S) * 6 * a.virial_pitch);
    }

// ---------------------------------------------------------------------------------------------
// PotentialPair<E>::computeForces  
// Ad: param_type, static bool eval(rsq, rcutsq, const param_type&, bool shift, S& fdr, S& eng)
// ---------------------------------------------------------------------------------------------
template<class S, class Ad> void iso_loop(const PairArgs<S>& a, const void* params_v)
    {
    typedef typename Ad::param_type P;
    const P* params = static_cast<const P*>(params_v);
    zero_outputs(a);
    const int nthreads = a.half_list == 1 ? 1 : (a.nthreads > 0 ? a.nthreads : 1);
    (void)nthreads;
    auto row = [&](const unsigned int i, const unsigned int dlo, const unsigned int dhi)
        {
        const S* pi = a.pos + 4 * (size_t)i;
        const unsigned int ti = type_of(pi);
        S fx = 0, fy = 0, fz = 0, pe = 0;
        S w[6] = {0, 0, 0, 0, 0, 0};
        const uint64_t head = a.head_list[i];
        const unsigned int nn = a.n_neigh[i];
        for (unsigned int k = 0; k < nn; ++k)
            {
            const unsigned int j = a.nlist[head + k];
            bool react = a.half_list == 1 && j < a.N; // third law to the neighbour
            if (a.half_list == 2)
                {
                react = j >= dlo && j < dhi;
                if (react && j < i)
                    continue; // pair inside the domain: taken from the lower index only
                }
            const S* pj = a.pos + 4 * (size_t)j;
            const unsigned int tj = type_of(pj);
            S dx = pi[0] - pj[0], dy = pi[1] - pj[1], dz = pi[2] - pj[2];
            if (a.rint_image)
                min_image_rint(a.box, dx, dy, dz);
            else
                min_image_host(a.box, dx, dy, dz);
            const S rsq = dx * dx + dy * dy + dz * dz;
            const unsigned int tp = tj * a.ntypes + ti;
            const S rcutsq = a.rcutsq[tp];
            S ronsq = S(0);
            if (a.shift_mode == 2)
                ronsq = a.ronsq[tp];
            bool energy_shift = false;
            if (a.shift_mode == 1)
                energy_shift = true;
            else if (a.shift_mode == 2 && ronsq > rcutsq)
                energy_shift = true;

            S fdr = S(0), eng = S(0);
            const bool evaluated = Ad::eval(rsq, rcutsq, params[tp], energy_shift, fdr, eng);
            if (!evaluated)
                continue;
            if (a.shift_mode == 2 && rsq >= ronsq && rsq < rcutsq)
                {
                const S old_eng = eng, old_fdr = fdr;
                const S dr2 = rcutsq - ronsq;
                const S denom_inv = S(1.0) / (dr2 * dr2 * dr2);
                const S m = rsq - rcutsq;
                const S s = m * m * (rcutsq + S(2.0) * rsq - S(3.0) * ronsq) * denom_inv;
                const S ds = S(12.0) * (rsq - ronsq) * m * denom_inv;
                eng = old_eng * s;
                fdr = s * old_fdr - ds * old_eng;
                }
            S vw[6] = {0, 0, 0, 0, 0, 0};
            if (a.compute_virial)
                {
                const S h = S(0.5) * fdr;
                vw[0] = h * dx * dx;
                vw[1] = h * dx * dy;
                vw[2] = h * dx * dz;
                vw[3] = h * dy * dy;
                vw[4] = h * dy * dz;
                vw[5] = h * dz * dz;
                for (int c = 0; c < 6; ++c)
                    w[c] += vw[c];
                }
            fx += dx * fdr;
            fy += dy * fdr;
            fz += dz * fdr;
            pe += eng * S(0.5);
            if (react)
                {
                S* fj = a.force + 4 * (size_t)j;
                fj[0] -= dx * fdr;
                fj[1] -= dy * fdr;
                fj[2] -= dz * fdr;
                fj[3] += eng * S(0.5);
                if (a.compute_virial)
                    for (int c = 0; c < 6; ++c)
                        a.virial[c * a.virial_pitch + j] += vw[c];
                }
            }
        S* fi = a.force + 4 * (size_t)i;
        fi[0] += fx;
        fi[1] += fy;
        fi[2] += fz;
        fi[3] += pe;
        if (a.compute_virial)
            for (int c = 0; c < 6; ++c)
                a.virial[c * a.virial_pitch + i] += w[c];
        };
    if (a.half_list == 2)
        {
        const int nd = a.nthreads > 0 ? a.nthreads : 1;
#pragma omp parallel for schedule(static, 1) num_threads(nd)
        for (int d = 0; d < nd; ++d)
            {
            const unsigned int dlo = (unsigned int)((unsigned long long)a.N * d / nd);
            const unsigned int dhi = (unsigned int)((unsigned long long)a.N * (d + 1) / nd);
            for (unsigned int i = dlo; i < dhi; ++i)
                row(i, dlo, dhi);
            }
        return;
        }
#pragma omp parallel for schedule(dynamic, 256) num_threads(nthreads)
    for (long long ii = 0; ii < (long long)a.N; ++ii)
        row((unsigned int)ii, 0u, 0u);
    }

// ---------------------------------------------------------------------------------------------
// PotentialPairDPDThermo<E>::computeForces  
// Ad: param_type, static bool eval_thermo(rsq, rcutsq, P, seed, tag_i, tag_j, timestep, dt,
//                                         rdotv, T, S& fdr, S& fdr_cons, S& eng)
// ---------------------------------------------------------------------------------------------
template<class S, class Ad> void dpd_loop(const PairArgs<S>& a, const void* params_v)
    {
    typedef typename Ad::param_type P;
    const P* params = static_cast<const P*>(params_v);
    zero_outputs(a);
    const int nthreads = a.half_list == 1 ? 1 : (a.nthreads > 0 ? a.nthreads : 1);
    (void)nthreads;
    auto row = [&](const unsigned int i, const unsigned int dlo, const unsigned int dhi)
        {
        const S* pi = a.pos + 4 * (size_t)i;
        const S* vi = a.vel + 4 * (size_t)i;
        const unsigned int ti = type_of(pi);
        S fx = 0, fy = 0, fz = 0, pe = 0;
        S w[6] = {0, 0, 0, 0, 0, 0};
        const uint64_t head = a.head_list[i];
        const unsigned int nn = a.n_neigh[i];
        for (unsigned int k = 0; k < nn; ++k)
            {
            const unsigned int j = a.nlist[head + k];
            bool react = a.half_list == 1 && j < a.N; // third law to the neighbour
            if (a.half_list == 2)
                {
                react = j >= dlo && j < dhi;
                if (react && j < i)
                    continue; // pair inside the domain: taken from the lower index only
                }
            const S* pj = a.pos + 4 * (size_t)j;
            const S* vj = a.vel + 4 * (size_t)j;
            const unsigned int tj = type_of(pj);
            S dx = pi[0] - pj[0], dy = pi[1] - pj[1], dz = pi[2] - pj[2];
            if (a.rint_image)
                min_image_rint(a.box, dx, dy, dz);
            else
                min_image_host(a.box, dx, dy, dz);
            const S dvx = vi[0] - vj[0], dvy = vi[1] - vj[1], dvz = vi[2] - vj[2];
            const S rsq = dx * dx + dy * dy + dz * dz;
            const S rdotv = dx * dvx + dy * dvy + dz * dvz;
            const unsigned int tp = tj * a.ntypes + ti;
            const S rcutsq = a.rcutsq[tp];
            S fdr = S(0), fdr_cons = S(0), eng = S(0);
            const bool evaluated = Ad::eval_thermo(rsq,
                                                   rcutsq,
                                                   params[tp],
                                                   a.seed,
                                                   a.tag[i],
                                                   a.tag[j],
                                                   (unsigned int)a.timestep,
                                                   a.deltaT,
                                                   rdotv,
                                                   a.T,
                                                   fdr,
                                                   fdr_cons,
                                                   eng);
            if (!evaluated)
                continue;
            S vw[6] = {0, 0, 0, 0, 0, 0};
            if (a.compute_virial)
                {
                const S h = S(0.5) * fdr_cons; // conservative part only
                vw[0] = h * dx * dx;
                vw[1] = h * dx * dy;
                vw[2] = h * dx * dz;
                vw[3] = h * dy * dy;
                vw[4] = h * dy * dz;
                vw[5] = h * dz * dz;
                for (int c = 0; c < 6; ++c)
                    w[c] += vw[c];
                }
            fx += dx * fdr;
            fy += dy * fdr;
            fz += dz * fdr;
            pe += eng * S(0.5);
            if (react)
                {
                S* fj = a.force + 4 * (size_t)j;
                fj[0] -= dx * fdr;
                fj[1] -= dy * fdr;
                fj[2] -= dz * fdr;
                fj[3] += eng * S(0.5);
                if (a.compute_virial)
                    for (int c = 0; c < 6; ++c)
                        a.virial[c * a.virial_pitch + j] += vw[c];
                }
            }
        S* fi = a.force + 4 * (size_t)i;
        fi[0] += fx;
        fi[1] += fy;
        fi[2] += fz;
        fi[3] += pe;
        if (a.compute_virial)
            for (int c = 0; c < 6; ++c)
                a.virial[c * a.virial_pitch + i] += w[c];
        };
    if (a.half_list == 2)
        {
        const int nd = a.nthreads > 0 ? a.nthreads : 1;
#pragma omp parallel for schedule(static, 1) num_threads(nd)
        for (int d = 0; d < nd; ++d)
            {
            const unsigned int dlo = (unsigned int)((unsigned long long)a.N * d / nd);
            const unsigned int dhi = (unsigned int)((unsigned long long)a.N * (d + 1) / nd);
            for (unsigned int i = dlo; i < dhi; ++i)
                row(i, dlo, dhi);
            }
        return;
        }
#pragma omp parallel for schedule(dynamic, 256) num_threads(nthreads)
    for (long long ii = 0; ii < (long long)a.N; ++ii)
        row((unsigned int)ii, 0u, 0u);
    }

// ---------------------------------------------------------------------------------------------
// AnisoPotentialPair<E>::computeForces  
// Ad: param_type, static bool eval(dr[3], qi[4], qj[4], rcutsq, P, shift, f[3], eng, ti[3], tj[3])
// ---------------------------------------------------------------------------------------------
template<class S, class Ad> void aniso_loop(const PairArgs<S>& a, const void* params_v)
    {
    typedef typename Ad::param_type P;
    const P* params = static_cast<const P*>(params_v);
    zero_outputs(a);
    const int nthreads = a.half_list == 1 ? 1 : (a.nthreads > 0 ? a.nthreads : 1);
    (void)nthreads;
    auto row = [&](const unsigned int i, const unsigned int dlo, const unsigned int dhi)
        {
        const S* pi = a.pos + 4 * (size_t)i;
        const S* qi = a.orientation + 4 * (size_t)i;
        const unsigned int ti = type_of(pi);
        S fx = 0, fy = 0, fz = 0, pe = 0, tx = 0, ty = 0, tz = 0;
        S w[6] = {0, 0, 0, 0, 0, 0};
        const uint64_t head = a.head_list[i];
        const unsigned int nn = a.n_neigh[i];
        for (unsigned int k = 0; k < nn; ++k)
            {
            const unsigned int j = a.nlist[head + k];
            bool react = a.half_list == 1 && j < a.N; // third law to the neighbour
            if (a.half_list == 2)
                {
                react = j >= dlo && j < dhi;
                if (react && j < i)
                    continue; // pair inside the domain: taken from the lower index only
                }
            const S* pj = a.pos + 4 * (size_t)j;
            const S* qj = a.orientation + 4 * (size_t)j;
            const unsigned int tj = type_of(pj);
            S dr[3] = {pi[0] - pj[0], pi[1] - pj[1], pi[2] - pj[2]};
            if (a.rint_image)
                min_image_rint(a.box, dr[0], dr[1], dr[2]);
            else
                min_image_host(a.box, dr[0], dr[1], dr[2]);
            const unsigned int tp = tj * a.ntypes + ti;
            const S rcutsq = a.rcutsq[tp];
            const bool energy_shift = (a.shift_mode == 1);
            S f[3] = {0, 0, 0}, t_i[3] = {0, 0, 0}, t_j[3] = {0, 0, 0};
            S eng = S(0);
            const bool evaluated
                = Ad::eval(dr, qi, qj, rcutsq, params[tp], energy_shift, f, eng, t_i, t_j);
            if (!evaluated)
                continue;
            S vw[6] = {0, 0, 0, 0, 0, 0};
            if (a.compute_virial)
                {
                vw[0] = S(0.5) * dr[0] * f[0];
                vw[1] = S(0.5) * dr[1] * f[0];
                vw[2] = S(0.5) * dr[2] * f[0];
                vw[3] = S(0.5) * dr[1] * f[1];
                vw[4] = S(0.5) * dr[2] * f[1];
                vw[5] = S(0.5) * dr[2] * f[2];
                for (int c = 0; c < 6; ++c)
                    w[c] += vw[c];
                }
            fx += f[0];
            fy += f[1];
            fz += f[2];
            tx += t_i[0];
            ty += t_i[1];
            tz += t_i[2];
            pe += eng * S(0.5);
            if (react)
                {
                S* fj = a.force + 4 * (size_t)j;
                S* tqj = a.torque + 4 * (size_t)j;
                fj[0] -= f[0];
                fj[1] -= f[1];
                fj[2] -= f[2];
                fj[3] += eng * S(0.5);
                tqj[0] += t_j[0];
                tqj[1] += t_j[1];
                tqj[2] += t_j[2];
                if (a.compute_virial)
                    for (int c = 0; c < 6; ++c)
                        a.virial[c * a.virial_pitch + j] += vw[c];
                }
            }
        S* fi = a.force + 4 * (size_t)i;
        S* tqi = a.torque + 4 * (size_t)i;
        fi[0] += fx;
        fi[1] += fy;
        fi[2] += fz;
        fi[3] += pe;
        tqi[0] += tx;
        tqi[1] += ty;
        tqi[2] += tz;
        if (a.compute_virial)
            for (int c = 0; c < 6; ++c)
                a.virial[c * a.virial_pitch + i] += w[c];
        };
    if (a.half_list == 2)
        {
        const int nd = a.nthreads > 0 ? a.nthreads : 1;
#pragma omp parallel for schedule(static, 1) num_threads(nd)
        for (int d = 0; d < nd; ++d)
            {
            const unsigned int dlo = (unsigned int)((unsigned long long)a.N * d / nd);
            const unsigned int dhi = (unsigned int)((unsigned long long)a.N * (d + 1) / nd);
            for (unsigned int i = dlo; i < dhi; ++i)
                row(i, dlo, dhi);
            }
        return;
        }
#pragma omp parallel for schedule(dynamic, 256) num_threads(nthreads)
    for (long long ii = 0; ii < (long long)a.N; ++ii)
        row((unsigned int)ii, 0u, 0u);
    }
    } // namespace azp_oracle

#endif
