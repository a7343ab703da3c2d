// pair_kernels.cuh -- the sm_100a neighbour-list pair-force kernels and their launch layer.
//
// Replaces the kernel bodies HOOMD-blue supplies to azplugins through
//   gpu_compute_pair_forces<E>        (reference src/PotentialPairGPUKernel.cu.inc:25-28)
//   gpu_compute_dpd_forces<E>         (reference src/PotentialPairDPDThermoGPUKernel.cu.inc:21-24)
//   gpu_compute_pair_aniso_forces<E>  (reference src/AnisoPotentialPairGPUKernel.cu.inc:21-25)
// with a from-scratch design for B200 (SURVEY.md 3.2-3.4 give the per-pair sequence kept here):
//
//   * one row (particle) per group of `tpp` consecutive lanes, tpp a runtime power of two <= 32:
//     the lanes stride the row, so nlist reads are coalesced and the float4 position gathers of
//     a group hit a handful of 128-byte lines (spatially sorted particles);
//   * per type-pair constants (Evaluator::cache_type: parameters, derived constants, energy at
//     r_cut, r_cut^2, xplor coefficients) are built once per CTA and staged in shared memory;
//     single-type systems (the NT1 instantiation) keep them in registers;
//   * two neighbours in flight per lane (index + position loads issued before the math) to cover
//     L2/L1 gather latency; rejected and out-of-range slots are folded into the cutoff test by
//     giving them rsq = +inf;
//   * minimum image costs 3 full-rate instructions per axis (magic-number rint, no FRND), and is
//     skipped for a whole warp when every row of the warp is farther than the largest cutoff from
//     all periodic faces -- for such rows wrapping can only change pairs that fail the cutoff
//     test either way, so the result is bit-identical;
//   * force/energy/virial/torque are reduced over the tpp lanes with xor-shuffles and written by
//     lane 0 as one float4 (+ one float4 torque, + 6 virial scalars).
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

template<class S> AZP_D S infinity();
template<> AZP_D float infinity<float>()
    {
    return __int_as_float(0x7f800000);
    }
template<> AZP_D double infinity<double>()
    {
    return __longlong_as_double(0x7ff0000000000000ll);
    }

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

// Row bookkeeping shared by the three kernels.
template<class S> struct RowInfo
    {
    unsigned int row;   // output / nlist row
    unsigned int i;     // global particle index
    unsigned int n;     // neighbours in the row
    unsigned int lane;  // lane within the tpp group
    uint64_t head;
    bool active;
    };

template<class S> AZP_D RowInfo<S> locate_row(const KernelArgs<S>& a, unsigned int tpp_log2)
    {
    RowInfo<S> r;
    const unsigned int gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned int slot = gtid >> tpp_log2;
    r.lane = gtid & ((1u << tpp_log2) - 1u);
    const unsigned int nslots = a.row_ids ? a.n_row_ids : a.N;
    r.active = slot < nslots;
    r.row = 0;
    r.n = 0;
    r.head = 0;
    if (r.active)
        {
        r.row = a.row_ids ? __ldg(a.row_ids + slot) : slot;
        r.n = __ldg(a.n_neigh + r.row);
        r.head = __ldg(a.head_list + r.row);
        }
    r.i = r.row + a.row_offset;
    return r;
    }

// true when the warp may skip the minimum-image wrap: fully periodic orthorhombic box and every
// active row of the warp farther than rc_max from all faces (see file header).
template<class S>
AZP_D bool warp_is_interior(const BoxDim<S>& b, const Vec4<S>& pi, S rc_max, bool active)
    {
    const S m = S(0.49999);
    const bool inside = (fabs(pi.x) + rc_max < m * b.L[0]) && (fabs(pi.y) + rc_max < m * b.L[1])
                        && (fabs(pi.z) + rc_max < m * b.L[2]);
    return (b.flags == 2) && __all_sync(0xffffffffu, inside || !active);
    }

template<class S> AZP_D void displacement(const BoxDim<S>& b, bool skip_wrap, const Vec4<S>& pi, const Vec4<S>& pj, S& dx, S& dy, S& dz)
    {
    dx = pi.x - pj.x;
    dy = pi.y - pj.y;
    dz = pi.z - pj.z;
    if (!skip_wrap)
        {
        if (b.flags == 2)
            min_image_ortho(b.L[0], b.L[1], b.L[2], b.Linv[0], b.Linv[1], b.Linv[2], dx, dy, dz);
        else
            min_image_general(b, dx, dy, dz);
        }
    }

// ---------------------------------------------------------------------------------------------
// Isotropic kernel: F_i = sum dx * force_divr, E_i = 1/2 sum U, W_i = 1/2 sum dx_a dx_b force_divr
// ---------------------------------------------------------------------------------------------
template<class E, class S, bool XPLOR, bool VIRIAL, bool NT1>
__global__ void __launch_bounds__(kMaxBlock)
    pair_force_kernel(const __grid_constant__ KernelArgs<S> a,
                      const typename E::param_type* __restrict__ params,
                      const unsigned int tpp_log2)
    {
    typedef typename E::cache_type Cache;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const unsigned int ntp = NT1 ? 1u : a.ntypes * a.ntypes;
    Cache* s_cache = reinterpret_cast<Cache*>(smem_raw);
    S* s_rcutsq = reinterpret_cast<S*>(s_cache + ntp);
    XplorEntry<S>* s_xplor = reinterpret_cast<XplorEntry<S>*>(s_rcutsq + ntp + 1);

    for (unsigned int t = threadIdx.x; t < ntp; t += blockDim.x)
        {
        const S rc = a.rcutsq[t];
        S ron = S(0);
        if (XPLOR)
            ron = a.ronsq[t];
        const bool energy_shift = (a.shift_mode == 1) || (XPLOR && ron > rc);
        s_cache[t] = E::make_cache(params[t], rc, energy_shift);
        s_rcutsq[t] = rc;
        if (XPLOR)
            {
            const S d = rc - ron;
            s_xplor[t].ronsq = ron;
            s_xplor[t].denom_inv = S(1.0) / (d * d * d);
            }
        }
    if (threadIdx.x == 0)
        {
        S m = S(0);
        for (unsigned int t = 0; t < ntp; ++t)
            m = fmax(m, a.rcutsq[t]);
        s_rcutsq[ntp] = ::sqrt(m);
        }
    __syncthreads();

    const RowInfo<S> r = locate_row(a, tpp_log2);
    const unsigned int tpp = 1u << tpp_log2;
    const Vec4<S> pi = load4(a.pos, r.active ? r.i : 0u);
    const unsigned int ti = scalar_as_uint(pi.w);
    const bool skip_wrap = warp_is_interior(a.box, pi, s_rcutsq[ntp], r.active);

    S fx = S(0), fy = S(0), fz = S(0), pe = S(0);
    S w0 = S(0), w1 = S(0), w2 = S(0), w3 = S(0), w4 = S(0), w5 = S(0);

    // registers for the single-type case
    const Cache c0 = s_cache[0];
    const S rc0 = s_rcutsq[0];

    auto pair = [&](const Vec4<S>& pj, bool valid)
    {
        S dx, dy, dz;
        displacement(a.box, skip_wrap, pi, pj, dx, dy, dz);
        S rsq = dx * dx + dy * dy + dz * dz;
        if (!valid)
            rsq = infinity<S>();
        unsigned int tp = 0;
        if (!NT1)
            tp = index2d(a.ntypes, ti, scalar_as_uint(pj.w));
        const S rcutsq = NT1 ? rc0 : s_rcutsq[tp];
        if (rsq < rcutsq)
            {
            S force_divr = S(0), pair_eng = S(0);
            const Cache* cp = NT1 ? &c0 : &s_cache[tp];
            E eval(rsq, rcutsq, *cp);
            eval.evalForceAndEnergy(force_divr, pair_eng, false);
            if (XPLOR)
                {
                const XplorEntry<S> x = s_xplor[tp];
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
            // force by explicit FMA in both variants, so toggling the virial does not change
            // a single bit of the forces
            fx = fma(dx, force_divr, fx);
            fy = fma(dy, force_divr, fy);
            fz = fma(dz, force_divr, fz);
            pe += pair_eng;
            if (VIRIAL)
                {
                const S vx = dx * force_divr, vy = dy * force_divr, vz = dz * force_divr;
                w0 += dx * vx;
                w1 += dx * vy;
                w2 += dx * vz;
                w3 += dy * vy;
                w4 += dy * vz;
                w5 += dz * vz;
                }
            }
    };

    const unsigned int* __restrict__ row = a.nlist + r.head;
    for (unsigned int k = r.lane; k < r.n; k += 2 * tpp)
        {
        const unsigned int k1 = k + tpp;
        const bool has1 = k1 < r.n;
        const unsigned int j0 = __ldg(row + k);
        const unsigned int j1 = has1 ? __ldg(row + k1) : j0;
        const Vec4<S> p0 = load4(a.pos, j0);
        const Vec4<S> p1 = load4(a.pos, j1);
        pair(p0, true);
        pair(p1, has1);
        }

    for (unsigned int o = tpp >> 1; o > 0; o >>= 1)
        {
        fx += shfl_xor(fx, o);
        fy += shfl_xor(fy, o);
        fz += shfl_xor(fz, o);
        pe += shfl_xor(pe, o);
        if (VIRIAL)
            {
            w0 += shfl_xor(w0, o);
            w1 += shfl_xor(w1, o);
            w2 += shfl_xor(w2, o);
            w3 += shfl_xor(w3, o);
            w4 += shfl_xor(w4, o);
            w5 += shfl_xor(w5, o);
            }
        }

    if (r.active && r.lane == 0)
        {
        store4(a.force, r.row, fx, fy, fz, S(0.5) * pe);
        if (VIRIAL)
            {
            S* v = a.virial + r.row;
            const size_t p = a.virial_pitch;
            v[0] = S(0.5) * w0;
            v[p] = S(0.5) * w1;
            v[2 * p] = S(0.5) * w2;
            v[3 * p] = S(0.5) * w3;
            v[4 * p] = S(0.5) * w4;
            v[5 * p] = S(0.5) * w5;
            }
        }
    }

// ---------------------------------------------------------------------------------------------
// DPD thermostat kernel: also gathers vel_j and tag_j; force from force_divr (conservative +
// drag + random), virial from the conservative part only.
// ---------------------------------------------------------------------------------------------
template<class E, class S, bool VIRIAL, bool NT1>
__global__ void __launch_bounds__(kMaxBlock)
    dpd_force_kernel(const __grid_constant__ KernelArgs<S> a,
                     const typename E::param_type* __restrict__ params,
                     const unsigned int tpp_log2)
    {
    typedef typename E::cache_type Cache;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const unsigned int ntp = NT1 ? 1u : a.ntypes * a.ntypes;
    Cache* s_cache = reinterpret_cast<Cache*>(smem_raw);
    S* s_rcutsq = reinterpret_cast<S*>(s_cache + ntp);

    for (unsigned int t = threadIdx.x; t < ntp; t += blockDim.x)
        {
        const S rc = a.rcutsq[t];
        s_cache[t] = E::make_cache_thermo(params[t], rc, a.deltaT, a.T);
        s_rcutsq[t] = rc;
        }
    if (threadIdx.x == 0)
        {
        S m = S(0);
        for (unsigned int t = 0; t < ntp; ++t)
            m = fmax(m, a.rcutsq[t]);
        s_rcutsq[ntp] = ::sqrt(m);
        }
    __syncthreads();

    const RowInfo<S> r = locate_row(a, tpp_log2);
    const unsigned int tpp = 1u << tpp_log2;
    const unsigned int isafe = r.active ? r.i : 0u;
    const Vec4<S> pi = load4(a.pos, isafe);
    const Vec4<S> vi = load4(a.vel, isafe);
    const unsigned int tag_i = __ldg(a.tag + isafe);
    const unsigned int ti = scalar_as_uint(pi.w);
    const bool skip_wrap = warp_is_interior(a.box, pi, s_rcutsq[ntp], r.active);

    S fx = S(0), fy = S(0), fz = S(0), pe = S(0);
    S w0 = S(0), w1 = S(0), w2 = S(0), w3 = S(0), w4 = S(0), w5 = S(0);
    const Cache c0 = s_cache[0];
    const S rc0 = s_rcutsq[0];

    auto pair = [&](unsigned int j, const Vec4<S>& pj, bool valid)
    {
        S dx, dy, dz;
        displacement(a.box, skip_wrap, pi, pj, dx, dy, dz);
        S rsq = dx * dx + dy * dy + dz * dz;
        if (!valid)
            rsq = infinity<S>();
        unsigned int tp = 0;
        if (!NT1)
            tp = index2d(a.ntypes, ti, scalar_as_uint(pj.w));
        const S rcutsq = NT1 ? rc0 : s_rcutsq[tp];
        if (rsq < rcutsq)
            {
            // velocity and tag are only needed for accepted pairs (about a third of the list at
            // buffer 0.4), so they are gathered behind the cutoff test
            const Vec4<S> vj = load4(a.vel, j);
            const unsigned int tag_j = __ldg(a.tag + j);
            const S rdotv = dx * (vi.x - vj.x) + dy * (vi.y - vj.y) + dz * (vi.z - vj.z);
            S force_divr = S(0), force_divr_cons = S(0), pair_eng = S(0);
            const Cache* cp = NT1 ? &c0 : &s_cache[tp];
            E eval(rsq, rcutsq, *cp);
            eval.set_seed_ij_timestep((uint16_t)a.seed, tag_i, tag_j, a.timestep);
            eval.setDeltaT(a.deltaT);
            eval.setRDotV(rdotv);
            eval.setT(a.T);
            eval.evalForceEnergyThermo(force_divr, force_divr_cons, pair_eng, false);
            fx = fma(dx, force_divr, fx);
            fy = fma(dy, force_divr, fy);
            fz = fma(dz, force_divr, fz);
            pe += pair_eng;
            if (VIRIAL)
                {
                const S vx = dx * force_divr_cons, vy = dy * force_divr_cons,
                        vz = dz * force_divr_cons;
                w0 += dx * vx;
                w1 += dx * vy;
                w2 += dx * vz;
                w3 += dy * vy;
                w4 += dy * vz;
                w5 += dz * vz;
                }
            }
    };

    const unsigned int* __restrict__ row = a.nlist + r.head;
    for (unsigned int k = r.lane; k < r.n; k += 2 * tpp)
        {
        const unsigned int k1 = k + tpp;
        const bool has1 = k1 < r.n;
        const unsigned int j0 = __ldg(row + k);
        const unsigned int j1 = has1 ? __ldg(row + k1) : j0;
        const Vec4<S> p0 = load4(a.pos, j0);
        const Vec4<S> p1 = load4(a.pos, j1);
        pair(j0, p0, true);
        pair(j1, p1, has1);
        }

    for (unsigned int o = tpp >> 1; o > 0; o >>= 1)
        {
        fx += shfl_xor(fx, o);
        fy += shfl_xor(fy, o);
        fz += shfl_xor(fz, o);
        pe += shfl_xor(pe, o);
        if (VIRIAL)
            {
            w0 += shfl_xor(w0, o);
            w1 += shfl_xor(w1, o);
            w2 += shfl_xor(w2, o);
            w3 += shfl_xor(w3, o);
            w4 += shfl_xor(w4, o);
            w5 += shfl_xor(w5, o);
            }
        }

    if (r.active && r.lane == 0)
        {
        store4(a.force, r.row, fx, fy, fz, S(0.5) * pe);
        if (VIRIAL)
            {
            S* v = a.virial + r.row;
            const size_t p = a.virial_pitch;
            v[0] = S(0.5) * w0;
            v[p] = S(0.5) * w1;
            v[2 * p] = S(0.5) * w2;
            v[3 * p] = S(0.5) * w3;
            v[4 * p] = S(0.5) * w4;
            v[5 * p] = S(0.5) * w5;
            }
        }
    }

// ---------------------------------------------------------------------------------------------
// Anisotropic kernel: gathers orientation_j; vector force, torque on i, energy,
// virial 1/2 dx_a F_b.
// ---------------------------------------------------------------------------------------------
template<class E, class S, bool VIRIAL, bool NT1>
__global__ void __launch_bounds__(kMaxBlock)
    aniso_force_kernel(const __grid_constant__ KernelArgs<S> a,
                       const typename E::param_type* __restrict__ params,
                       const unsigned int tpp_log2)
    {
    typedef typename E::cache_type Cache;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const unsigned int ntp = NT1 ? 1u : a.ntypes * a.ntypes;
    Cache* s_cache = reinterpret_cast<Cache*>(smem_raw);
    S* s_rcutsq = reinterpret_cast<S*>(s_cache + ntp);

    for (unsigned int t = threadIdx.x; t < ntp; t += blockDim.x)
        {
        const S rc = a.rcutsq[t];
        s_cache[t] = E::make_cache(params[t], rc, a.shift_mode == 1);
        s_rcutsq[t] = rc;
        }
    if (threadIdx.x == 0)
        {
        S m = S(0);
        for (unsigned int t = 0; t < ntp; ++t)
            m = fmax(m, a.rcutsq[t]);
        s_rcutsq[ntp] = ::sqrt(m);
        }
    __syncthreads();

    const RowInfo<S> r = locate_row(a, tpp_log2);
    const unsigned int tpp = 1u << tpp_log2;
    const unsigned int isafe = r.active ? r.i : 0u;
    const Vec4<S> pi = load4(a.pos, isafe);
    const Vec4<S> qi = load4(a.orientation, isafe);
    const unsigned int ti = scalar_as_uint(pi.w);
    const bool skip_wrap = warp_is_interior(a.box, pi, s_rcutsq[ntp], r.active);

    S fx = S(0), fy = S(0), fz = S(0), pe = S(0);
    S tx = S(0), ty = S(0), tz = S(0);
    S w0 = S(0), w1 = S(0), w2 = S(0), w3 = S(0), w4 = S(0), w5 = S(0);
    const Cache c0 = s_cache[0];
    const S rc0 = s_rcutsq[0];

    auto pair = [&](unsigned int j, const Vec4<S>& pj, bool valid)
    {
        Vec3<S> dr;
        displacement(a.box, skip_wrap, pi, pj, dr.x, dr.y, dr.z);
        const S rsq = dr.x * dr.x + dr.y * dr.y + dr.z * dr.z;
        unsigned int tp = 0;
        if (!NT1)
            tp = index2d(a.ntypes, ti, scalar_as_uint(pj.w));
        const S rcutsq = NT1 ? rc0 : s_rcutsq[tp];
        // the evaluator accepts rsq <= rcutsq; test here so that the orientation gather and the
        // evaluator body are skipped for the (about half) rejected entries
        if (valid && !(rsq > rcutsq))
            {
            const Vec4<S> qj = load4(a.orientation, j);
            Vec3<S> force {S(0), S(0), S(0)}, torque_i {S(0), S(0), S(0)},
                torque_j {S(0), S(0), S(0)};
            S pair_eng = S(0);
            const Cache* cp = NT1 ? &c0 : &s_cache[tp];
            E eval(dr, qi, qj, rcutsq, *cp);
            eval.evaluate(force, pair_eng, false, torque_i, torque_j);
            fx += force.x;
            fy += force.y;
            fz += force.z;
            tx += torque_i.x;
            ty += torque_i.y;
            tz += torque_i.z;
            pe += pair_eng;
            if (VIRIAL)
                {
                w0 += dr.x * force.x;
                w1 += dr.y * force.x;
                w2 += dr.z * force.x;
                w3 += dr.y * force.y;
                w4 += dr.z * force.y;
                w5 += dr.z * force.z;
                }
            }
    };

    const unsigned int* __restrict__ row = a.nlist + r.head;
    for (unsigned int k = r.lane; k < r.n; k += 2 * tpp)
        {
        const unsigned int k1 = k + tpp;
        const bool has1 = k1 < r.n;
        const unsigned int j0 = __ldg(row + k);
        const unsigned int j1 = has1 ? __ldg(row + k1) : j0;
        const Vec4<S> p0 = load4(a.pos, j0);
        const Vec4<S> p1 = load4(a.pos, j1);
        pair(j0, p0, true);
        pair(j1, p1, has1);
        }

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
            {
            w0 += shfl_xor(w0, o);
            w1 += shfl_xor(w1, o);
            w2 += shfl_xor(w2, o);
            w3 += shfl_xor(w3, o);
            w4 += shfl_xor(w4, o);
            w5 += shfl_xor(w5, o);
            }
        }

    if (r.active && r.lane == 0)
        {
        store4(a.force, r.row, fx, fy, fz, S(0.5) * pe);
        store4(a.torque, r.row, tx, ty, tz, S(0));
        if (VIRIAL)
            {
            S* v = a.virial + r.row;
            const size_t p = a.virial_pitch;
            v[0] = S(0.5) * w0;
            v[p] = S(0.5) * w1;
            v[2 * p] = S(0.5) * w2;
            v[3 * p] = S(0.5) * w3;
            v[4 * p] = S(0.5) * w4;
            v[5 * p] = S(0.5) * w5;
            }
        }
    }
    } // namespace azp

#endif
