// pair_kernels.cuh -- the sm_100a neighbour-list pair-force kernels.
//
// Replaces the kernel bodies HOOMD-blue supplies to azplugins through
//   gpu_compute_pair_forces<E>        (reference src/PotentialPairGPUKernel.cu.inc:25-28)
//   gpu_compute_dpd_forces<E>         (reference src/PotentialPairDPDThermoGPUKernel.cu.inc:21-24)
//   gpu_compute_pair_aniso_forces<E>  (reference src/AnisoPotentialPairGPUKernel.cu.inc:21-25)
// with a from-scratch design for B200 (SURVEY.md 3.2-3.4 give the per-pair sequence kept here).
//
// One kernel skeleton (row_kernel) drives three "families" (IsoFamily, DpdFamily, AnisoFamily)
// that differ in what they gather per neighbour and what they accumulate:
//
//   * one row (particle) per group of `tpp` consecutive lanes, tpp a runtime power of two <= 32;
//   * the row is consumed as aligned uint4 vectors of four neighbour indices: one 16-byte
//     coalesced nlist load, then four independent 16-byte position gathers in flight per lane
//     (ld.global.nc, L1-resident for spatially sorted particles) before any math. The first and
//     last partial vectors of a row go through a guarded scalar path, so rows may start at any
//     head_list offset and nothing outside [head, head + n_neigh) is ever read;
//   * per type-pair constants (Evaluator::cache_type: parameters, derived constants, energy at
//     r_cut) and the effective r_cut^2 are built once per CTA in shared memory; a pair whose
//     potential is switched off gets r_cut^2 = 0, so ONE compare per neighbour implements the
//     reference's "rsq < rcutsq && parameter != 0" test. Single-type systems (NT1) keep the
//     constants in registers;
//   * minimum image costs 3 full-rate instructions per axis (magic-number rint, no FRND) and is
//     skipped for a whole warp when every row of the warp is farther than the largest cutoff
//     from all periodic faces: for such rows wrapping can only change pairs that fail the
//     cutoff test either way, so the result is bit-identical;
//   * forces accumulate with explicit FMAs (identical bits with or without the virial);
//     force/energy/virial/torque are reduced over the tpp lanes with xor-shuffles and written by
//     lane 0 as one 16-byte store (+ one for the torque, + 6 virial scalars).
// There is no tensor-core work here: the path is a gather-bound stencil.
#ifndef AZP_PAIR_KERNELS_CUH_
#define AZP_PAIR_KERNELS_CUH_

#include "azp_core.cuh"
#include "azp_philox.cuh"

namespace azp
    {
template<class S> struct KernelArgs
    {
    S* force;
    S* virial;
    S* torque;
    size_t virial_pitch;
    const S* pos;
    const S* vel;
    const S* orientation;
    const unsigned int* tag;
    const unsigned int* n_neigh;
    const unsigned int* nlist;
    const uint64_t* head_list;
    const S* rcutsq;
    const S* ronsq;
    const unsigned int* row_ids;
    BoxDim<S> box;
    unsigned int N;
    unsigned int ntypes;
    unsigned int shift_mode;
    unsigned int row_offset;
    unsigned int n_row_ids;
    unsigned int seed;
    unsigned int timestep;
    S deltaT;
    S T;
    };

constexpr unsigned int kMaxBlock = 512;

template<class S> AZP_D S shfl_xor(S v, unsigned int o)
    {
    return __shfl_xor_sync(0xffffffffu, v, o);
    }

// xplor smoothing coefficients per type pair
template<class S> struct XplorEntry
    {
    S ronsq;
    S denom_inv; // 1 / (rcutsq - ronsq)^3
    };

// ---------------------------------------------------------------------------------------------
// Per-row geometry shared by the families: particle i, its type, and the displacement rule.
// ---------------------------------------------------------------------------------------------
template<class S> struct RowGeometry
    {
    Vec4<S> pi;
    unsigned int ti;
    bool skip_wrap;
    S Lx, Ly, Lz, ix, iy, iz;

    AZP_D void displacement(const BoxDim<S>& b, const Vec4<S>& pj, S& dx, S& dy, S& dz) const
        {
        dx = pi.x - pj.x;
        dy = pi.y - pj.y;
        dz = pi.z - pj.z;
        if (!skip_wrap)
            {
            if (b.flags == 2)
                min_image_ortho(Lx, Ly, Lz, ix, iy, iz, dx, dy, dz);
            else
                min_image_general(b, dx, dy, dz);
            }
        }
    };

// Virial accumulators (compiled out when VIRIAL is false).
template<class S> struct Virial6
    {
    S xx = S(0), xy = S(0), xz = S(0), yy = S(0), yz = S(0), zz = S(0);
    AZP_D void reduce(unsigned int o)
        {
        xx += shfl_xor(xx, o);
        xy += shfl_xor(xy, o);
        xz += shfl_xor(xz, o);
        yy += shfl_xor(yy, o);
        yz += shfl_xor(yz, o);
        zz += shfl_xor(zz, o);
        }
    AZP_D void store(S* virial, size_t pitch, unsigned int row) const
        {
        S* v = virial + row;
        v[0] = S(0.5) * xx;
        v[pitch] = S(0.5) * xy;
        v[2 * pitch] = S(0.5) * xz;
        v[3 * pitch] = S(0.5) * yy;
        v[4 * pitch] = S(0.5) * yz;
        v[5 * pitch] = S(0.5) * zz;
        }
    };

// Shared-memory type-pair table common to the families: cache[ntp], rcutsq_eff[ntp], rc_max.
// The table is addressed through the `extern __shared__` array itself (never through a stored
// generic pointer), so every access compiles to an LDS with a known address space.
extern __shared__ __align__(16) unsigned char azp_smem[];

template<class E, class S> struct PairTable
    {
    typedef typename E::cache_type Cache;
    unsigned int rcutsq_off; // byte offset of rcutsq[] (the cache array starts at 0)

    AZP_HD static size_t bytes(size_t ntp)
        {
        return ntp * sizeof(Cache) + (ntp + 1) * sizeof(S);
        }
    AZP_D void carve(unsigned int ntp)
        {
        rcutsq_off = ntp * (unsigned int)sizeof(Cache);
        }
    AZP_D Cache& cache(unsigned int t) const
        {
        return reinterpret_cast<Cache*>(azp_smem)[t];
        }
    // effective r_cut^2: 0 for pairs whose potential is switched off
    AZP_D S& rcutsq(unsigned int t) const
        {
        return reinterpret_cast<S*>(azp_smem + rcutsq_off)[t];
        }
    AZP_D unsigned int end_off(unsigned int ntp) const
        {
        return rcutsq_off + (ntp + 1) * (unsigned int)sizeof(S);
        }
    // sqrt(max rcutsq), for the interior-warp test; stored after rcutsq[ntp - 1]
    AZP_D void finish(unsigned int ntp)
        {
        if (threadIdx.x == 0)
            {
            S m = S(0);
            for (unsigned int t = 0; t < ntp; ++t)
                m = fmax(m, rcutsq(t));
            rcutsq(ntp) = ::sqrt(m);
            }
        }
    };

// =============================================================================================
// Isotropic family: F_i = sum dx * force_divr, E_i = 1/2 sum U, W_i = 1/2 sum dx_a dx_b force_divr
// =============================================================================================
template<class E_, class S_, bool XPLOR, bool VIRIAL, bool NT1_> struct IsoFamily
    {
    typedef E_ E;
    typedef S_ S;
    typedef typename E::cache_type Cache;
    static constexpr bool NT1 = NT1_;

    PairTable<E, S> tab;
    unsigned int xplor_off;
    Cache c0;
    S rc0;
    S fx = S(0), fy = S(0), fz = S(0), pe = S(0);
    Virial6<S> w;

    AZP_HD static size_t smem_bytes(size_t ntp)
        {
        return PairTable<E, S>::bytes(ntp) + (XPLOR ? ntp * sizeof(XplorEntry<S>) : 0) + 32;
        }

    AZP_D XplorEntry<S>& xplor(unsigned int t) const
        {
        return reinterpret_cast<XplorEntry<S>*>(azp_smem + xplor_off)[t];
        }

    AZP_D void stage(const KernelArgs<S>& a, const typename E::param_type* params, unsigned int ntp)
        {
        tab.carve(ntp);
        xplor_off = (tab.end_off(ntp) + 15u) & ~15u;
        for (unsigned int t = threadIdx.x; t < ntp; t += blockDim.x)
            {
            const S rc = a.rcutsq[t];
            S ron = S(0);
            if (XPLOR)
                ron = a.ronsq[t];
            const bool energy_shift = (a.shift_mode == 1) || (XPLOR && ron > rc);
            const Cache c = E::make_cache(params[t], rc, energy_shift);
            tab.cache(t) = c;
            tab.rcutsq(t) = E::disabled(c) ? S(0) : rc;
            if (XPLOR)
                {
                const S d = rc - ron;
                xplor(t).ronsq = ron;
                xplor(t).denom_inv = S(1.0) / (d * d * d);
                }
            }
        __syncthreads();
        tab.finish(ntp);
        __syncthreads();
        c0 = tab.cache(0);
        rc0 = tab.rcutsq(0);
        }

    AZP_D void begin_row(const KernelArgs<S>&, unsigned int) { }

    template<class C>
    AZP_D void accept(const C& c, unsigned int tp, S rsq, S rcutsq, S dx, S dy, S dz)
        {
        S force_divr = S(0), pair_eng = S(0);
        E eval(rsq, rcutsq, c);
        eval.evalPair(force_divr, pair_eng, false);
        if (XPLOR)
            {
            const XplorEntry<S> x = xplor(tp);
            if (rsq >= x.ronsq)
                {
                const S m = rsq - rcutsq;
                const S s = m * m * (rcutsq + S(2.0) * rsq - S(3.0) * x.ronsq) * x.denom_inv;
                const S ds = S(12.0) * (rsq - x.ronsq) * m * x.denom_inv;
                const S old_eng = pair_eng;
                pair_eng = old_eng * s;
                force_divr = s * force_divr - ds * old_eng;
                }
            }
        fx = fma(dx, force_divr, fx);
        fy = fma(dy, force_divr, fy);
        fz = fma(dz, force_divr, fz);
        pe += pair_eng;
        if (VIRIAL)
            {
            const S vx = dx * force_divr, vy = dy * force_divr, vz = dz * force_divr;
            w.xx = fma(dx, vx, w.xx);
            w.xy = fma(dx, vy, w.xy);
            w.xz = fma(dx, vz, w.xz);
            w.yy = fma(dy, vy, w.yy);
            w.yz = fma(dy, vz, w.yz);
            w.zz = fma(dz, vz, w.zz);
            }
        }

    AZP_D void pair(const KernelArgs<S>& a, const RowGeometry<S>& g, unsigned int, const Vec4<S>& pj)
        {
        S dx, dy, dz;
        g.displacement(a.box, pj, dx, dy, dz);
        const S rsq = fma(dz, dz, fma(dy, dy, dx * dx));
        if (NT1)
            {
            if (rsq < rc0)
                accept(c0, 0u, rsq, rc0, dx, dy, dz);
            }
        else
            {
            const unsigned int tp = index2d(a.ntypes, g.ti, scalar_as_uint(pj.w));
            const S rcutsq = tab.rcutsq(tp);
            if (rsq < rcutsq)
                accept(tab.cache(tp), tp, rsq, rcutsq, dx, dy, dz);
            }
        }

    AZP_D void finish(const KernelArgs<S>& a, unsigned int row, bool writer, unsigned int tpp)
        {
        for (unsigned int o = tpp >> 1; o > 0; o >>= 1)
            {
            fx += shfl_xor(fx, o);
            fy += shfl_xor(fy, o);
            fz += shfl_xor(fz, o);
            pe += shfl_xor(pe, o);
            if (VIRIAL)
                w.reduce(o);
            }
        if (writer)
            {
            store4(a.force, row, fx, fy, fz, S(0.5) * pe);
            if (VIRIAL)
                w.store(a.virial, a.virial_pitch, row);
            }
        }
    };

// =============================================================================================
// DPD thermostat family: also gathers vel_j and tag_j; force from force_divr (conservative +
// drag + random), virial from the conservative part only.
// =============================================================================================
template<class E_, class S_, bool VIRIAL, bool NT1_> struct DpdFamily
    {
    typedef E_ E;
    typedef S_ S;
    typedef typename E::cache_type Cache;
    static constexpr bool NT1 = NT1_;

    PairTable<E, S> tab;
    Cache c0;
    S rc0;
    Vec4<S> vi;
    unsigned int tag_i;
    S fx = S(0), fy = S(0), fz = S(0), pe = S(0);
    Virial6<S> w;

    AZP_HD static size_t smem_bytes(size_t ntp)
        {
        return PairTable<E, S>::bytes(ntp) + 16;
        }

    AZP_D void stage(const KernelArgs<S>& a, const typename E::param_type* params, unsigned int ntp)
        {
        tab.carve(ntp);
        for (unsigned int t = threadIdx.x; t < ntp; t += blockDim.x)
            {
            const S rc = a.rcutsq[t];
            tab.cache(t) = E::make_cache_thermo(params[t], rc, a.deltaT, a.T);
            tab.rcutsq(t) = rc;
            }
        __syncthreads();
        tab.finish(ntp);
        __syncthreads();
        c0 = tab.cache(0);
        rc0 = tab.rcutsq(0);
        }

    AZP_D void begin_row(const KernelArgs<S>& a, unsigned int i)
        {
        vi = load4(a.vel, i);
        tag_i = __ldg(a.tag + i);
        }

    template<class C>
    AZP_D void accept(const KernelArgs<S>& a, const C& c, unsigned int j, S rsq, S rcutsq, S dx, S dy, S dz)
        {
        // velocity and tag are only needed for accepted pairs (about a third of the list at
        // buffer 0.4), so they are gathered behind the cutoff test
        const Vec4<S> vj = load4(a.vel, j);
        const unsigned int tag_j = __ldg(a.tag + j);
        const S rdotv = dx * (vi.x - vj.x) + dy * (vi.y - vj.y) + dz * (vi.z - vj.z);
        S force_divr = S(0), force_divr_cons = S(0), pair_eng = S(0);
        E eval(rsq, rcutsq, c);
        eval.set_seed_ij_timestep((uint16_t)a.seed, tag_i, tag_j, a.timestep);
        eval.setDeltaT(a.deltaT);
        eval.setRDotV(rdotv);
        eval.setT(a.T);
        eval.evalThermoPair(force_divr, force_divr_cons, pair_eng, false);
        fx = fma(dx, force_divr, fx);
        fy = fma(dy, force_divr, fy);
        fz = fma(dz, force_divr, fz);
        pe += pair_eng;
        if (VIRIAL)
            {
            const S vx = dx * force_divr_cons, vy = dy * force_divr_cons, vz = dz * force_divr_cons;
            w.xx = fma(dx, vx, w.xx);
            w.xy = fma(dx, vy, w.xy);
            w.xz = fma(dx, vz, w.xz);
            w.yy = fma(dy, vy, w.yy);
            w.yz = fma(dy, vz, w.yz);
            w.zz = fma(dz, vz, w.zz);
            }
        }

    AZP_D void pair(const KernelArgs<S>& a, const RowGeometry<S>& g, unsigned int j, const Vec4<S>& pj)
        {
        S dx, dy, dz;
        g.displacement(a.box, pj, dx, dy, dz);
        const S rsq = fma(dz, dz, fma(dy, dy, dx * dx));
        if (NT1)
            {
            if (rsq < rc0)
                accept(a, c0, j, rsq, rc0, dx, dy, dz);
            }
        else
            {
            const unsigned int tp = index2d(a.ntypes, g.ti, scalar_as_uint(pj.w));
            const S rcutsq = tab.rcutsq(tp);
            if (rsq < rcutsq)
                accept(a, tab.cache(tp), j, rsq, rcutsq, dx, dy, dz);
            }
        }

    AZP_D void finish(const KernelArgs<S>& a, unsigned int row, bool writer, unsigned int tpp)
        {
        for (unsigned int o = tpp >> 1; o > 0; o >>= 1)
            {
            fx += shfl_xor(fx, o);
            fy += shfl_xor(fy, o);
            fz += shfl_xor(fz, o);
            pe += shfl_xor(pe, o);
            if (VIRIAL)
                w.reduce(o);
            }
        if (writer)
            {
            store4(a.force, row, fx, fy, fz, S(0.5) * pe);
            if (VIRIAL)
                w.store(a.virial, a.virial_pitch, row);
            }
        }
    };

// =============================================================================================
// Anisotropic family: gathers orientation_j; vector force, torque on i, energy,
// virial 1/2 dx_a F_b.
// =============================================================================================
template<class E_, class S_, bool VIRIAL, bool NT1_> struct AnisoFamily
    {
    typedef E_ E;
    typedef S_ S;
    typedef typename E::cache_type Cache;
    static constexpr bool NT1 = NT1_;

    PairTable<E, S> tab;
    Cache c0;
    S rc0;
    Vec4<S> qi;
    S fx = S(0), fy = S(0), fz = S(0), pe = S(0);
    S tx = S(0), ty = S(0), tz = S(0);
    Virial6<S> w;

    AZP_HD static size_t smem_bytes(size_t ntp)
        {
        return PairTable<E, S>::bytes(ntp) + 16;
        }

    AZP_D void stage(const KernelArgs<S>& a, const typename E::param_type* params, unsigned int ntp)
        {
        tab.carve(ntp);
        for (unsigned int t = threadIdx.x; t < ntp; t += blockDim.x)
            {
            const S rc = a.rcutsq[t];
            tab.cache(t) = E::make_cache(params[t], rc, a.shift_mode == 1);
            tab.rcutsq(t) = rc;
            }
        __syncthreads();
        tab.finish(ntp);
        __syncthreads();
        c0 = tab.cache(0);
        rc0 = tab.rcutsq(0);
        }

    AZP_D void begin_row(const KernelArgs<S>& a, unsigned int i)
        {
        qi = load4(a.orientation, i);
        }

    template<class C>
    AZP_D void accept(const KernelArgs<S>& a, const C& c, unsigned int j, S rsq, S rcutsq, const Vec3<S>& dr)
        {
        // the orientation gather and the evaluator body are skipped for the (about half)
        // rejected entries
        const Vec4<S> qj = load4(a.orientation, j);
        Vec3<S> force {S(0), S(0), S(0)}, torque_i {S(0), S(0), S(0)}, torque_j {S(0), S(0), S(0)};
        S pair_eng = S(0);
        E eval(dr, qi, qj, rcutsq, c);
        eval.evaluatePair(rsq, force, pair_eng, false, torque_i, torque_j);
        fx += force.x;
        fy += force.y;
        fz += force.z;
        tx += torque_i.x;
        ty += torque_i.y;
        tz += torque_i.z;
        pe += pair_eng;
        if (VIRIAL)
            {
            w.xx = fma(dr.x, force.x, w.xx);
            w.xy = fma(dr.y, force.x, w.xy);
            w.xz = fma(dr.z, force.x, w.xz);
            w.yy = fma(dr.y, force.y, w.yy);
            w.yz = fma(dr.z, force.y, w.yz);
            w.zz = fma(dr.z, force.z, w.zz);
            }
        }

    AZP_D void pair(const KernelArgs<S>& a, const RowGeometry<S>& g, unsigned int j, const Vec4<S>& pj)
        {
        Vec3<S> dr;
        g.displacement(a.box, pj, dr.x, dr.y, dr.z);
        const S rsq = fma(dr.z, dr.z, fma(dr.y, dr.y, dr.x * dr.x));
        // the reference evaluator rejects only rsq > rcutsq (strictly), so accept rsq <= rcutsq
        if (NT1)
            {
            if (rsq <= rc0)
                accept(a, c0, j, rsq, rc0, dr);
            }
        else
            {
            const unsigned int tp = index2d(a.ntypes, g.ti, scalar_as_uint(pj.w));
            const S rcutsq = tab.rcutsq(tp);
            if (rsq <= rcutsq)
                accept(a, tab.cache(tp), j, rsq, rcutsq, dr);
            }
        }

    AZP_D void finish(const KernelArgs<S>& a, unsigned int row, bool writer, unsigned int tpp)
        {
        for (unsigned int o = tpp >> 1; o > 0; o >>= 1)
            {
            fx += shfl_xor(fx, o);
            fy += shfl_xor(fy, o);
            fz += shfl_xor(fz, o);
            pe += shfl_xor(pe, o);
            tx += shfl_xor(tx, o);
            ty += shfl_xor(ty, o);
            tz += shfl_xor(tz, o);
            if (VIRIAL)
                w.reduce(o);
            }
        if (writer)
            {
            store4(a.force, row, fx, fy, fz, S(0.5) * pe);
            store4(a.torque, row, tx, ty, tz, S(0));
            if (VIRIAL)
                w.store(a.virial, a.virial_pitch, row);
            }
        }
    };

// =============================================================================================
// The kernel skeleton
// =============================================================================================
template<class Fam>
__global__ void __launch_bounds__(kMaxBlock)
    row_kernel(const __grid_constant__ KernelArgs<typename Fam::S> a,
               const typename Fam::E::param_type* __restrict__ params,
               const unsigned int tpp_log2)
    {
    typedef typename Fam::S S;
    const unsigned int ntp = Fam::NT1 ? 1u : a.ntypes * a.ntypes;
    Fam fam;
    fam.stage(a, params, ntp);

    // ---- which row, which lane of the row's group ----------------------------------------
    const unsigned int tpp = 1u << tpp_log2;
    const unsigned int gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned int slot = gtid >> tpp_log2;
    const unsigned int lane = gtid & (tpp - 1u);
    const unsigned int nslots = a.row_ids ? a.n_row_ids : a.N;
    const bool active = slot < nslots;
    unsigned int row = 0, n = 0;
    uint64_t head = 0;
    if (active)
        {
        row = a.row_ids ? __ldg(a.row_ids + slot) : slot;
        n = __ldg(a.n_neigh + row);
        head = __ldg(a.head_list + row);
        }
    const unsigned int i = active ? row + a.row_offset : 0u;

    RowGeometry<S> g;
    g.pi = load4(a.pos, i);
    g.ti = scalar_as_uint(g.pi.w);
    g.Lx = a.box.L[0], g.Ly = a.box.L[1], g.Lz = a.box.L[2];
    g.ix = a.box.Linv[0], g.iy = a.box.Linv[1], g.iz = a.box.Linv[2];
        {
        // the warp may skip the minimum-image wrap when the box is orthorhombic and fully
        // periodic and every active row of the warp is farther than rc_max from all faces
        const S rc_max = fam.tab.rcutsq(ntp);
        const S m = S(0.49999);
        const bool inside = (fabs(g.pi.x) + rc_max < m * g.Lx) && (fabs(g.pi.y) + rc_max < m * g.Ly)
                            && (fabs(g.pi.z) + rc_max < m * g.Lz);
        g.skip_wrap = (a.box.flags == 2) && __all_sync(0xffffffffu, inside || !active);
        }
    fam.begin_row(a, i);

    // ---- the row as aligned uint4 vectors of neighbour indices ------------------------------
    // `pre` = entries between the 16-byte boundary below the row start and the row start.
    const unsigned int* rowp = a.nlist + head;
    const unsigned int pre = (unsigned int)((reinterpret_cast<uintptr_t>(rowp) >> 2) & 3u);
    const unsigned int* base = rowp - pre;
    const unsigned int end = pre + n; // valid entries of `base` are [pre, end)
    const uint4* base4 = reinterpret_cast<const uint4*>(base);
    for (unsigned int v = lane; 4u * v < end; v += tpp)
        {
        const unsigned int e = 4u * v;
        if (e >= pre && e + 4u <= end)
            {
            const uint4 j = __ldg(base4 + v);
            const Vec4<S> p0 = load4(a.pos, j.x);
            const Vec4<S> p1 = load4(a.pos, j.y);
            const Vec4<S> p2 = load4(a.pos, j.z);
            const Vec4<S> p3 = load4(a.pos, j.w);
            fam.pair(a, g, j.x, p0);
            fam.pair(a, g, j.y, p1);
            fam.pair(a, g, j.z, p2);
            fam.pair(a, g, j.w, p3);
            }
        else
            {
            // partial vector at either end of the row: guarded scalar loads
            for (unsigned int q = 0; q < 4u; ++q)
                {
                const unsigned int idx = e + q;
                if (idx >= pre && idx < end)
                    {
                    const unsigned int j = __ldg(base + idx);
                    const Vec4<S> pj = load4(a.pos, j);
                    fam.pair(a, g, j, pj);
                    }
                }
            }
        }

    fam.finish(a, row, active && lane == 0, tpp);
    }
    } // namespace azp

#endif
