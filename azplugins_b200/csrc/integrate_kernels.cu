// integrate_kernels.cu -- velocity-Verlet (NVE) steps either side of the force path
// (SURVEY.md 8(f) rank 4), so that the configurations run as real MD loops.
//
// HOOMD's md.methods.ConstantVolume (TwoStepConstantVolume, HOOMD-blue v7.0.1) is not in the
// reference tree; the reference only uses it (src/pytest/test_pair.py:325-327,
// test_external.py:31-36). Restated as the standard two-step scheme:
//   step one:  v += 1/2 a dt;  x += v dt;  wrap x into the box, counting images
//   (forces at the new positions)
//   step two:  F = sum of the force computes' forces (HOOMD's net force);  a = F / m;  v += 1/2 a dt
// Both are one-body streaming kernels (HBM-bound); the arithmetic is IEEE round-to-nearest
// without FMA contraction so that a plain numpy restatement (tests/test_md.py) is bit-exact.
// Parity with HOOMD itself is unpinned (no HOOMD here); the tests check the restatement and the
// conservation laws.
#include "../../include/azp_b200.h"
#include "azp_core.cuh"
#include "azp_philox.cuh"

namespace azp
    {
namespace md
    {
AZP_D float mul(float a, float b) { return __fmul_rn(a, b); }
AZP_D double mul(double a, double b) { return __dmul_rn(a, b); }
AZP_D float add(float a, float b) { return __fadd_rn(a, b); }
AZP_D double add(double a, double b) { return __dadd_rn(a, b); }
AZP_D float div(float a, float b) { return __fdiv_rn(a, b); }
AZP_D double div(double a, double b) { return __ddiv_rn(a, b); }
AZP_D float root(float a) { return __fsqrt_rn(a); }
AZP_D double root(double a) { return __dsqrt_rn(a); }

template<class S> struct Box
    {
    S lo[3], hi[3], L[3];
    int periodic[3];
    };

template<class S>
__global__ void __launch_bounds__(256) nve_step_one(S* pos, S* vel, const S* __restrict__ accel, int* __restrict__ image, const Box<S> box, const S dt, const unsigned int N)
    {
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N)
        return;
    // pos and vel are also written by this kernel: plain (coherent) loads, not ld.global.nc
    Vec4<S> p = plain_load4(pos, i);
    Vec4<S> v = plain_load4(vel, i);
    const Vec4<S> a = load4(accel, i);
    const S half_dt = mul(S(0.5), dt);
    v.x = add(v.x, mul(a.x, half_dt));
    v.y = add(v.y, mul(a.y, half_dt));
    v.z = add(v.z, mul(a.z, half_dt));
    S x[3] = {add(p.x, mul(v.x, dt)), add(p.y, mul(v.y, dt)), add(p.z, mul(v.z, dt))};
    int img[3] = {0, 0, 0};
    if (image)
        {
        img[0] = image[3 * i], img[1] = image[3 * i + 1], img[2] = image[3 * i + 2];
        }
#pragma unroll
    for (int d = 0; d < 3; ++d)
        {
        if (box.periodic[d])
            {
            if (x[d] >= box.hi[d])
                {
                x[d] = add(x[d], -box.L[d]);
                ++img[d];
                }
            else if (x[d] < box.lo[d])
                {
                x[d] = add(x[d], box.L[d]);
                --img[d];
                }
            }
        }
    store4(pos, i, x[0], x[1], x[2], p.w);
    store4(vel, i, v.x, v.y, v.z, v.w);
    if (image)
        {
        image[3 * i] = img[0], image[3 * i + 1] = img[1], image[3 * i + 2] = img[2];
        }
    }

template<class S> struct ForceList
    {
    const S* f[AZP_MD_MAX_FORCES];
    unsigned int n;
    };

template<class S>
__global__ void __launch_bounds__(256) nve_step_two(S* vel, S* __restrict__ accel, S* __restrict__ net_force, const ForceList<S> forces, const S dt, const unsigned int N)
    {
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N)
        return;
    S fx = S(0), fy = S(0), fz = S(0), e = S(0);
    for (unsigned int k = 0; k < forces.n; ++k)
        {
        const Vec4<S> f = load4(forces.f[k], i);
        fx = add(fx, f.x), fy = add(fy, f.y), fz = add(fz, f.z), e = add(e, f.w);
        }
    Vec4<S> v = plain_load4(vel, i); // vel is written below: no ld.global.nc
    const S minv = div(S(1.0), v.w);
    const S ax = mul(fx, minv), ay = mul(fy, minv), az = mul(fz, minv);
    const S half_dt = mul(S(0.5), dt);
    v.x = add(v.x, mul(ax, half_dt));
    v.y = add(v.y, mul(ay, half_dt));
    v.z = add(v.z, mul(az, half_dt));
    store4(vel, i, v.x, v.y, v.z, v.w);
    store4(accel, i, ax, ay, az, S(0));
    if (net_force)
        store4(net_force, i, fx, fy, fz, e);
    }

// Langevin step two (HOOMD md.methods.Langevin, TwoStepLangevin::integrateStepTwo; not in the
// reference tree -- BASELINE.json configs[0] runs the PerturbedLennardJones fluid under it):
//   F_bd = coeff (r_x, r_y, r_z) - gamma v,  coeff = sqrt(6 gamma kT / dt),  r ~ Uniform(-1, 1)
//   a = (F + F_bd) / m;  v += 1/2 a dt
// gamma per particle type; three draws per particle from
//   RandomGenerator(Seed(rng_id, timestep, seed16), Counter(tag))     
// i.e. Philox4x32-10 with key {id<<24 | seed<<8 | (timestep>>32)&0xff, timestep&0xffffffff} and
// counter {draw, 0, 0, tag}, draw = 0, 1, 2. The net-force array keeps the sum of the force
// computes only (as HOOMD's does).
template<class S> struct LangevinParams
    {
    const unsigned int* tag;
    const S* gamma;
    const S* pos; // type id in pos.w
    unsigned int ntypes;
    uint32_t key0, key1;
    S kT;
    int noiseless;
    };

template<class S> AZP_D S langevin_draw(uint32_t draw, uint32_t tag, uint32_t k0, uint32_t k1)
    {
    const Philox4 u = philox4x32_10(draw, 0u, 0u, tag, k0, k1);
    // UniformDistribution(-1, 1): a + (b - a) * canonical
    return add(S(-1), mul(S(2), u01_from(u, S())));
    }

template<class S>
__global__ void __launch_bounds__(256) langevin_step_two(S* vel, S* __restrict__ accel, S* __restrict__ net_force, const ForceList<S> forces, const LangevinParams<S> lp, const S dt, const unsigned int N)
    {
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N)
        return;
    S fx = S(0), fy = S(0), fz = S(0), e = S(0);
    for (unsigned int k = 0; k < forces.n; ++k)
        {
        const Vec4<S> f = load4(forces.f[k], i);
        fx = add(fx, f.x), fy = add(fy, f.y), fz = add(fz, f.z), e = add(e, f.w);
        }
    Vec4<S> v = plain_load4(vel, i);
    unsigned int type = scalar_as_uint(load4(lp.pos, i).w);
    type = type < lp.ntypes ? type : lp.ntypes - 1u;
    const S gamma = lp.gamma[type];
    const unsigned int tag = lp.tag[i];
    const S rx = langevin_draw<S>(0u, tag, lp.key0, lp.key1);
    const S ry = langevin_draw<S>(1u, tag, lp.key0, lp.key1);
    const S rz = langevin_draw<S>(2u, tag, lp.key0, lp.key1);
    const S coeff = lp.noiseless ? S(0) : root(div(mul(mul(S(6), gamma), lp.kT), dt));
    const S bx = add(mul(rx, coeff), -mul(gamma, v.x));
    const S by = add(mul(ry, coeff), -mul(gamma, v.y));
    const S bz = add(mul(rz, coeff), -mul(gamma, v.z));
    const S minv = div(S(1.0), v.w);
    const S ax = mul(add(fx, bx), minv), ay = mul(add(fy, by), minv), az = mul(add(fz, bz), minv);
    const S half_dt = mul(S(0.5), dt);
    v.x = add(v.x, mul(ax, half_dt));
    v.y = add(v.y, mul(ay, half_dt));
    v.z = add(v.z, mul(az, half_dt));
    store4(vel, i, v.x, v.y, v.z, v.w);
    store4(accel, i, ax, ay, az, S(0));
    if (net_force)
        store4(net_force, i, fx, fy, fz, e);
    }

template<class S> static int step_one(const azp_md_args* a, cudaStream_t stream)
    {
    if (!a)
        return (int)cudaErrorInvalidValue;
    if (a->N == 0)
        return 0;
    if (!a->d_pos || !a->d_vel || !a->d_accel || a->box.tilt[0] != 0 || a->box.tilt[1] != 0 || a->box.tilt[2] != 0)
        return (int)cudaErrorInvalidValue; // orthorhombic boxes
    Box<S> box;
    for (int d = 0; d < 3; ++d)
        {
        box.L[d] = S(a->box.L[d]);
        box.lo[d] = -box.L[d] / S(2.0);
        box.hi[d] = box.lo[d] + box.L[d];
        box.periodic[d] = a->box.periodic[d];
        }
    const unsigned int block = 256;
    nve_step_one<S><<<(a->N + block - 1) / block, block, 0, stream>>>(static_cast<S*>(a->d_pos), static_cast<S*>(a->d_vel), static_cast<const S*>(a->d_accel), a->d_image, box, S(a->dt), a->N);
    return (int)cudaGetLastError();
    }

template<class S> static int step_two(const azp_md_args* a, cudaStream_t stream)
    {
    if (!a)
        return (int)cudaErrorInvalidValue;
    if (a->N == 0)
        return 0;
    if (!a->d_vel || !a->d_accel || a->n_forces > AZP_MD_MAX_FORCES)
        return (int)cudaErrorInvalidValue;
    ForceList<S> fl;
    fl.n = a->n_forces;
    for (unsigned int k = 0; k < AZP_MD_MAX_FORCES; ++k)
        {
        fl.f[k] = k < a->n_forces ? static_cast<const S*>(a->d_forces[k]) : nullptr;
        if (k < a->n_forces && !fl.f[k])
            return (int)cudaErrorInvalidValue;
        }
    const unsigned int block = 256;
    nve_step_two<S><<<(a->N + block - 1) / block, block, 0, stream>>>(static_cast<S*>(a->d_vel), static_cast<S*>(a->d_accel), static_cast<S*>(a->d_net_force), fl, S(a->dt), a->N);
    return (int)cudaGetLastError();
    }
template<class S> static int langevin_two(const azp_md_args* a, const azp_langevin_args* l, cudaStream_t stream)
    {
    if (!a || !l)
        return (int)cudaErrorInvalidValue;
    if (a->N == 0)
        return 0;
    if (!a->d_vel || !a->d_accel || !a->d_pos || a->n_forces > AZP_MD_MAX_FORCES || !l->d_tag || !l->d_gamma || l->ntypes == 0
        || !(a->dt > 0) || l->kT < 0)
        return (int)cudaErrorInvalidValue;
    ForceList<S> fl;
    fl.n = a->n_forces;
    for (unsigned int k = 0; k < AZP_MD_MAX_FORCES; ++k)
        {
        fl.f[k] = k < a->n_forces ? static_cast<const S*>(a->d_forces[k]) : nullptr;
        if (k < a->n_forces && !fl.f[k])
            return (int)cudaErrorInvalidValue;
        }
    LangevinParams<S> lp;
    lp.tag = l->d_tag;
    lp.gamma = static_cast<const S*>(l->d_gamma);
    lp.pos = static_cast<const S*>(a->d_pos);
    lp.ntypes = l->ntypes;
    lp.key0 = ((l->rng_id & 0xffu) << 24) | ((l->seed & 0xffffu) << 8) | (uint32_t)((l->timestep >> 32) & 0xffull);
    lp.key1 = (uint32_t)(l->timestep & 0xffffffffull);
    lp.kT = S(l->kT);
    lp.noiseless = l->noiseless ? 1 : 0;
    const unsigned int block = 256;
    langevin_step_two<S><<<(a->N + block - 1) / block, block, 0, stream>>>(static_cast<S*>(a->d_vel), static_cast<S*>(a->d_accel), static_cast<S*>(a->d_net_force), fl, lp, S(a->dt), a->N);
    return (int)cudaGetLastError();
    }
    } // namespace md
    } // namespace azp

extern "C"
    {
    int azp_nve_step_one_f32(const azp_md_args* a, void* stream) { return azp::md::step_one<float>(a, (cudaStream_t)stream); }
    int azp_nve_step_one_f64(const azp_md_args* a, void* stream) { return azp::md::step_one<double>(a, (cudaStream_t)stream); }
    int azp_nve_step_two_f32(const azp_md_args* a, void* stream) { return azp::md::step_two<float>(a, (cudaStream_t)stream); }
    int azp_nve_step_two_f64(const azp_md_args* a, void* stream) { return azp::md::step_two<double>(a, (cudaStream_t)stream); }
    int azp_langevin_step_two_f32(const azp_md_args* a, const azp_langevin_args* l, void* stream) { return azp::md::langevin_two<float>(a, l, (cudaStream_t)stream); }
    int azp_langevin_step_two_f64(const azp_md_args* a, const azp_langevin_args* l, void* stream) { return azp::md::langevin_two<double>(a, l, (cudaStream_t)stream); }
    }
