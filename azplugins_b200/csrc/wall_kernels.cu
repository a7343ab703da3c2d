// wall_kernels.cu -- wall potentials wall.Colloid / wall.LJ93.
//
// Replaces hoomd::md::kernel::gpu_compute_potential_external_forces<EvaluatorWalls<E>> as
// instantiated by reference src/PotentialExternalWallGPUKernel.cu.inc:13-34 for
// E = WallEvaluatorColloid (src/WallEvaluatorColloid.h:91-203) and WallEvaluatorLJ93
// (src/WallEvaluatorLJ93.h:81-150). The evaluator arithmetic is in the reference tree and is
// repeated here operation for operation (IEEE, no FMA contraction); the wall loop
// (hoomd/md/EvaluatorWalls.h, hoomd/md/WallData.h of HOOMD-blue v7.0.1) is NOT in the tree and is
// restated from its published behaviour:
//   * for every sphere, cylinder and plane wall: drv = vector from the particle to the nearest
//     point of the wall, in_active_space = the particle is on the side the wall confines;
//   * in the active space the evaluator sees rsq = |drv|^2 with energy_shift = true, and
//     F += -drv * force_divr, U += pair_eng (a non-finite force_divr counts as zero);
//   * virial = F_a * pos_b (xx, xy, xz, yy, yz, zz);
//   * r_extrap > 0: closer to the wall than r_extrap, or on its wrong side, the potential is
//     evaluated at r_extrap and continued linearly (add_wall_extrap). The reference's own tests
//     never use this mode: restated from HOOMD's published behaviour, parity-unpinned.
// One-body streaming kernel, HBM-bound like the harmonic barrier (barrier_kernels.cu): 16 B in,
// 16 + 24 B out per particle; the wall list and the per-type parameters sit in shared memory.
#include "../../include/azp_b200.h"
#include "azp_core.cuh"

#include <math.h>

namespace azp
    {
namespace wall
    {
AZP_D float mul(float a, float b) { return __fmul_rn(a, b); }
AZP_D double mul(double a, double b) { return __dmul_rn(a, b); }
AZP_D float add(float a, float b) { return __fadd_rn(a, b); }
AZP_D double add(double a, double b) { return __dadd_rn(a, b); }
AZP_D float sub(float a, float b) { return __fadd_rn(a, -b); }
AZP_D double sub(double a, double b) { return __dadd_rn(a, -b); }
AZP_D float div(float a, float b) { return __fdiv_rn(a, b); }
AZP_D double div(double a, double b) { return __ddiv_rn(a, b); }
AZP_D float root(float a) { return __fsqrt_rn(a); }
AZP_D double root(double a) { return __dsqrt_rn(a); }
AZP_D float ln(float a) { return logf(a); }
AZP_D double ln(double a) { return ::log(a); }

// reference src/WallEvaluatorLJ93.h:93-132 (param_type {sigma_3, A})
template<class S> struct LJ93
    {
    struct param_type
        {
        S sigma_3, A;
        };
    S lj1, lj2;
    AZP_D explicit LJ93(const param_type& p)
        {
        lj1 = mul(mul(mul(mul(div(S(2.0), S(15.0)), p.A), p.sigma_3), p.sigma_3), p.sigma_3);
        lj2 = mul(p.A, p.sigma_3);
        }
    AZP_D bool eval(S rsq, S rcutsq, S& force_divr, S& energy) const
        {
        if (!(rsq < rcutsq && lj1 != S(0)))
            return false;
        const S r2inv = div(S(1.0), rsq);
        const S r3inv = mul(r2inv, root(r2inv));
        const S r6inv = mul(r3inv, r3inv);
        force_divr = mul(mul(r2inv, r3inv), sub(mul(mul(S(9.0), lj1), r6inv), mul(S(3.0), lj2)));
        energy = mul(r3inv, sub(mul(lj1, r6inv), lj2));
        // walls always shift (EvaluatorWalls passes energy_shift = true)
        const S rcut2inv = div(S(1.0), rcutsq);
        const S rcut3inv = mul(rcut2inv, root(rcut2inv));
        const S rcut6inv = mul(rcut3inv, rcut3inv);
        energy = sub(energy, mul(rcut3inv, sub(mul(lj1, rcut6inv), lj2)));
        return true;
        }
    };

// reference src/WallEvaluatorColloid.h:104-175 (param_type {c_1, c_2, a})
template<class S> struct Colloid
    {
    struct param_type
        {
        S c_1, c_2, a;
        };
    S c_1, c_2, a;
    AZP_D explicit Colloid(const param_type& p) : c_1(p.c_1), c_2(p.c_2), a(p.a) { }
    template<bool FORCE> AZP_D S potential(S& force_divr, S rsq) const
        {
        const S r = root(rsq);
        const S a_over_r = div(a, r);
        const S near_inv = div(S(1.0), sub(r, a));
        const S far_inv = div(S(1.0), add(r, a));
        const S both_inv = mul(near_inv, far_inv);
        const S near_inv2 = mul(near_inv, near_inv);
        const S near_inv6 = mul(mul(near_inv2, near_inv2), near_inv2);
        const S far_inv2 = mul(far_inv, far_inv);
        const S far_inv6 = mul(mul(far_inv2, far_inv2), far_inv2);
        if (FORCE)
            {
            const S a_over_r_x8 = mul(S(8.0), a_over_r);
            force_divr = mul(mul(S(6.0), c_1),
                             add(mul(mul(sub(a_over_r_x8, S(1.0)), near_inv2), near_inv6),
                                 mul(mul(add(a_over_r_x8, S(1.0)), far_inv2), far_inv6)));
            force_divr = sub(force_divr,
                             mul(c_2, mul(mul(mul(mul(mul(S(4.0), a), a), a_over_r), both_inv), both_inv)));
            }
        const S seven_a = mul(S(7.0), a);
        S energy = mul(c_1, add(mul(mul(sub(seven_a, r), near_inv), near_inv6),
                                mul(mul(add(seven_a, r), far_inv), far_inv6)));
        energy = sub(energy, mul(c_2, add(mul(mul(mul(S(2.0), a), r), both_inv),
                                          ln(div(far_inv, near_inv)))));
        return energy;
        }
    AZP_D bool eval(S rsq, S rcutsq, S& force_divr, S& energy) const
        {
        if (!(rsq < rcutsq && c_1 != S(0) && a > S(0)))
            return false;
        energy = potential<true>(force_divr, rsq);
        S unused;
        energy = sub(energy, potential<false>(unused, rcutsq));
        return true;
        }
    };

// per-type parameters: HOOMD's EvaluatorWalls<E>::param_type {E::param_type params; rcutsq; rextrap}
template<class S, class E> struct TypeParams
    {
    typename E::param_type params;
    S rcutsq;
    S rextrap;
    };

// flattened wall list (HOOMD wall_type): counts, then spheres, cylinders, planes
template<class S> struct Walls
    {
    unsigned int n_spheres, n_cylinders, n_planes, _pad;
    struct Sphere
        {
        S r, ox, oy, oz;
        int inside, open;
        } spheres[AZP_MAX_SPHERE_WALLS];
    struct Cylinder
        {
        S r, ox, oy, oz, ax, ay, az; // axis normalised by the host
        int inside, open;
        } cylinders[AZP_MAX_CYLINDER_WALLS];
    struct Plane
        {
        S ox, oy, oz, nx, ny, nz; // normal normalised by the host
        int open, _pad;
        } planes[AZP_MAX_PLANE_WALLS];
    };

template<class S> AZP_D bool active_side(S dist, S r, int inside, int open)
    {
    if (open)
        return (dist < r && inside) || (dist > r && !inside);
    return (dist <= r && inside) || (dist >= r && !inside);
    }

template<class S, class E> AZP_D void add_wall(const E& ev, S rcutsq, S dx, S dy, S dz, S& fx, S& fy, S& fz, S& energy)
    {
    // callEvaluator: dr = -drv, rsq = dr.dr
    const S rx = -dx, ry = -dy, rz = -dz;
    const S rsq = add(add(mul(rx, rx), mul(ry, ry)), mul(rz, rz));
    S force_divr = S(0), pair_eng = S(0);
    if (ev.eval(rsq, rcutsq, force_divr, pair_eng))
        {
        if (!isfinite(force_divr))
            {
            force_divr = S(0);
            pair_eng = S(0);
            }
        fx = add(fx, mul(rx, force_divr));
        fy = add(fy, mul(ry, force_divr));
        fz = add(fz, mul(rz, force_divr));
        energy = add(energy, pair_eng);
        }
    }

// One wall in HOOMD's extrapolated mode (r_extrap > 0): beyond r_extrap from the wall, in the
// active space, the potential as is; closer -- or on the wrong side of the wall -- the potential
// evaluated AT r_extrap and continued linearly with the force it has there:
//   U = U(r_e) + F(r_e) (r_e -/+ r),  F = F(r_e) along the wall normal, pointing into the active space.
// (nx, ny, nz) is the unit direction used when the particle sits exactly on the wall.
template<class S, class E>
AZP_D void add_wall_extrap(const E& ev, S rcutsq, S rextrap, bool in_active, S dx, S dy, S dz, S nx, S ny, S nz, S& fx, S& fy, S& fz, S& energy)
    {
    const S rextrapsq = mul(rextrap, rextrap);
    const S rsq = add(add(mul(dx, dx), mul(dy, dy)), mul(dz, dz));
    if (in_active && rsq >= rextrapsq)
        {
        add_wall(ev, rcutsq, dx, dy, dz, fx, fy, fz, energy);
        return;
        }
    S r = root(rsq);
    if (rsq == S(0))
        {
        in_active = true;
        dx = nx, dy = ny, dz = nz;
        }
    else
        {
        const S rinv = div(S(1.0), r);
        dx = mul(dx, rinv), dy = mul(dy, rinv), dz = mul(dz, rinv);
        }
    r = in_active ? sub(rextrap, r) : add(rextrap, r);
    const S scale = in_active ? rextrap : -rextrap;
    dx = mul(dx, scale), dy = mul(dy, scale), dz = mul(dz, scale);
    const S rx = -dx, ry = -dy, rz = -dz;
    S force_divr = S(0), pair_eng = S(0);
    if (ev.eval(rextrapsq, rcutsq, force_divr, pair_eng))
        {
        pair_eng = add(pair_eng, mul(mul(force_divr, rextrap), r));
        energy = add(energy, pair_eng);
        fx = add(fx, mul(rx, force_divr));
        fy = add(fy, mul(ry, force_divr));
        fz = add(fz, mul(rz, force_divr));
        }
    }

template<class S, class E>
__global__ void __launch_bounds__(256) wall_kernel(S* __restrict__ force,
                                                    S* __restrict__ virial,
                                                    const size_t virial_pitch,
                                                    const S* __restrict__ pos,
                                                    const TypeParams<S, E>* __restrict__ params,
                                                    const Walls<S>* __restrict__ d_walls,
                                                    const unsigned int N,
                                                    const unsigned int ntypes)
    {
    extern __shared__ __align__(16) unsigned char wall_smem[];
    Walls<S>* walls = reinterpret_cast<Walls<S>*>(wall_smem);
    TypeParams<S, E>* s_params = reinterpret_cast<TypeParams<S, E>*>(wall_smem + ((sizeof(Walls<S>) + 15) & ~size_t(15)));
        {
        const unsigned int* src = reinterpret_cast<const unsigned int*>(d_walls);
        unsigned int* dst = reinterpret_cast<unsigned int*>(walls);
        for (unsigned int w = threadIdx.x; w < sizeof(Walls<S>) / 4; w += blockDim.x)
            dst[w] = src[w];
        for (unsigned int t = threadIdx.x; t < ntypes; t += blockDim.x)
            s_params[t] = params[t];
        }
    __syncthreads();

    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N)
        return;
    const Vec4<S> p = load4(pos, i);
    const unsigned int type = scalar_as_uint(p.w);
    const TypeParams<S, E> tp = s_params[type < ntypes ? type : 0u];
    const E ev(tp.params);
    const bool extrap = tp.rextrap > S(0);
    S fx = S(0), fy = S(0), fz = S(0), energy = S(0);

    for (unsigned int k = 0; k < walls->n_spheres; ++k)
        {
        const typename Walls<S>::Sphere w = walls->spheres[k];
        const S tx = sub(p.x, w.ox), ty = sub(p.y, w.oy), tz = sub(p.z, w.oz);
        const S rxyz = root(add(add(mul(tx, tx), mul(ty, ty)), mul(tz, tz)));
        S dx, dy, dz;
        bool in_active;
        if (rxyz > S(0))
            {
            in_active = active_side(rxyz, w.r, w.inside, w.open);
            const S s = sub(div(w.r, rxyz), S(1.0));
            dx = mul(s, tx), dy = mul(s, ty), dz = mul(s, tz);
            }
        else
            {
            in_active = w.inside != 0;
            dx = w.r, dy = S(0), dz = S(0);
            }
        if (extrap)
            {
            // on-wall direction = from the particle towards the wall as seen from the active
            // space: radially outwards for an inside wall, inwards otherwise
            const S sgn = w.inside ? S(1) : S(-1);
            const S inv = rxyz > S(0) ? div(sgn, rxyz) : S(0);
            add_wall_extrap(ev, tp.rcutsq, tp.rextrap, in_active, dx, dy, dz, rxyz > S(0) ? mul(tx, inv) : sgn, mul(ty, inv), mul(tz, inv), fx, fy, fz, energy);
            }
        else if (in_active)
            add_wall(ev, tp.rcutsq, dx, dy, dz, fx, fy, fz, energy);
        }
    for (unsigned int k = 0; k < walls->n_cylinders; ++k)
        {
        const typename Walls<S>::Cylinder w = walls->cylinders[k];
        const S tx = sub(p.x, w.ox), ty = sub(p.y, w.oy), tz = sub(p.z, w.oz);
        const S along = add(add(mul(tx, w.ax), mul(ty, w.ay)), mul(tz, w.az));
        // component of t perpendicular to the axis
        const S qx = sub(tx, mul(along, w.ax)), qy = sub(ty, mul(along, w.ay)), qz = sub(tz, mul(along, w.az));
        const S rxy = root(add(add(mul(qx, qx), mul(qy, qy)), mul(qz, qz)));
        if (rxy > S(0))
            {
            const bool in_active = active_side(rxy, w.r, w.inside, w.open);
            const S s = sub(div(w.r, rxy), S(1.0));
            if (extrap)
                {
                const S inv = div(w.inside ? S(1) : S(-1), rxy);
                add_wall_extrap(ev, tp.rcutsq, tp.rextrap, in_active, mul(s, qx), mul(s, qy), mul(s, qz), mul(qx, inv), mul(qy, inv), mul(qz, inv), fx, fy, fz, energy);
                }
            else if (in_active)
                add_wall(ev, tp.rcutsq, mul(s, qx), mul(s, qy), mul(s, qz), fx, fy, fz, energy);
            }
        else if (w.inside)
            {
            // on the axis: any radial direction; take one perpendicular to the axis
            S ux = S(1), uy = S(0), uz = S(0);
            if (fabs(w.ax) > S(0.9))
                ux = S(0), uy = S(1);
            const S d = add(add(mul(ux, w.ax), mul(uy, w.ay)), mul(uz, w.az));
            ux = sub(ux, mul(d, w.ax)), uy = sub(uy, mul(d, w.ay)), uz = sub(uz, mul(d, w.az));
            const S n = div(w.r, root(add(add(mul(ux, ux), mul(uy, uy)), mul(uz, uz))));
            if (extrap)
                add_wall_extrap(ev, tp.rcutsq, tp.rextrap, true, mul(n, ux), mul(n, uy), mul(n, uz), S(0), S(0), S(0), fx, fy, fz, energy);
            else
                add_wall(ev, tp.rcutsq, mul(n, ux), mul(n, uy), mul(n, uz), fx, fy, fz, energy);
            }
        }
    for (unsigned int k = 0; k < walls->n_planes; ++k)
        {
        const typename Walls<S>::Plane w = walls->planes[k];
        const S d = sub(add(add(mul(w.nx, p.x), mul(w.ny, p.y)), mul(w.nz, p.z)),
                        add(add(mul(w.nx, w.ox), mul(w.ny, w.oy)), mul(w.nz, w.oz)));
        const bool in_active = w.open ? (d > S(0)) : (d >= S(0));
        if (extrap)
            add_wall_extrap(ev, tp.rcutsq, tp.rextrap, in_active, mul(-d, w.nx), mul(-d, w.ny), mul(-d, w.nz), -w.nx, -w.ny, -w.nz, fx, fy, fz, energy);
        else if (in_active)
            add_wall(ev, tp.rcutsq, mul(-d, w.nx), mul(-d, w.ny), mul(-d, w.nz), fx, fy, fz, energy);
        }

    store4(force, i, fx, fy, fz, energy);
    if (virial)
        {
        S* v = virial + i;
        v[0] = mul(fx, p.x);
        v[virial_pitch] = mul(fx, p.y);
        v[2 * virial_pitch] = mul(fx, p.z);
        v[3 * virial_pitch] = mul(fy, p.y);
        v[4 * virial_pitch] = mul(fy, p.z);
        v[5 * virial_pitch] = mul(fz, p.z);
        }
    }

template<class S, class E> static int launch(const azp_wall_args* a, cudaStream_t stream)
    {
    unsigned int block = a->block_size ? a->block_size : 256u;
    if (block % 32u != 0 || block > 256u)
        return (int)cudaErrorInvalidValue;
    const unsigned int grid = (a->N + block - 1u) / block;
    const size_t smem = ((sizeof(Walls<S>) + 15) & ~size_t(15)) + sizeof(TypeParams<S, E>) * a->ntypes;
    if (smem > 48u * 1024u)
        return (int)cudaErrorInvalidValue;
    wall_kernel<S, E><<<grid, block, smem, stream>>>(static_cast<S*>(a->d_force), static_cast<S*>(a->d_virial), (size_t)a->virial_pitch,
                                                     static_cast<const S*>(a->d_pos), static_cast<const TypeParams<S, E>*>(a->d_params),
                                                     static_cast<const Walls<S>*>(a->d_walls), a->N, a->ntypes);
    return (int)cudaGetLastError();
    }

template<class S> static int dispatch(int evaluator, const azp_wall_args* a, cudaStream_t stream)
    {
    if (!a)
        return (int)cudaErrorInvalidValue;
    if (a->N == 0)
        return 0;
    if (!a->d_force || !a->d_pos || !a->d_params || !a->d_walls || a->ntypes == 0)
        return (int)cudaErrorInvalidValue;
    if (evaluator == AZP_WALL_COLLOID)
        return launch<S, Colloid<S>>(a, stream);
    if (evaluator == AZP_WALL_LJ93)
        return launch<S, LJ93<S>>(a, stream);
    return (int)cudaErrorInvalidValue;
    }
    } // namespace wall
    } // namespace azp

extern "C"
    {
    int azp_wall_forces_f32(int evaluator, const azp_wall_args* args, void* stream)
        {
        return azp::wall::dispatch<float>(evaluator, args, (cudaStream_t)stream);
        }
    int azp_wall_forces_f64(int evaluator, const azp_wall_args* args, void* stream)
        {
        return azp::wall::dispatch<double>(evaluator, args, (cudaStream_t)stream);
        }
    int azp_wall_param_size(int evaluator, int scalar_bits)
        {
        const bool f32 = scalar_bits == 32;
        if (evaluator == AZP_WALL_COLLOID)
            return f32 ? (int)sizeof(azp::wall::TypeParams<float, azp::wall::Colloid<float>>)
                       : (int)sizeof(azp::wall::TypeParams<double, azp::wall::Colloid<double>>);
        if (evaluator == AZP_WALL_LJ93)
            return f32 ? (int)sizeof(azp::wall::TypeParams<float, azp::wall::LJ93<float>>)
                       : (int)sizeof(azp::wall::TypeParams<double, azp::wall::LJ93<double>>);
        return -1;
        }
    int azp_walls_size(int scalar_bits)
        {
        return scalar_bits == 32 ? (int)sizeof(azp::wall::Walls<float>) : (int)sizeof(azp::wall::Walls<double>);
        }
    }
