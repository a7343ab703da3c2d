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

template<class S> struct Box
    {
    S lo[3], hi[3], L[3];
    int periodic[3];
    };

template<class S>
__global__ void __launch_bounds__(256) nve_step_one(S* __restrict__ pos, S* __restrict__ vel, const S* __restrict__ accel, int* __restrict__ image, const Box<S> box, const S dt, const unsigned int N)
    {
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N)
        return;
    Vec4<S> p = load4(pos, i);
    Vec4<S> v = load4(vel, i);
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
__global__ void __launch_bounds__(256) nve_step_two(S* __restrict__ vel, S* __restrict__ accel, S* __restrict__ net_force, const ForceList<S> forces, const S dt, const unsigned int N)
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
    Vec4<S> v = load4(vel, i);
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
    } // namespace md
    } // namespace azp

extern "C"
    {
    int azp_nve_step_one_f32(const azp_md_args* a, void* stream) { return azp::md::step_one<float>(a, (cudaStream_t)stream); }
    int azp_nve_step_one_f64(const azp_md_args* a, void* stream) { return azp::md::step_one<double>(a, (cudaStream_t)stream); }
    int azp_nve_step_two_f32(const azp_md_args* a, void* stream) { return azp::md::step_two<float>(a, (cudaStream_t)stream); }
    int azp_nve_step_two_f64(const azp_md_args* a, void* stream) { return azp::md::step_two<double>(a, (cudaStream_t)stream); }
    }
