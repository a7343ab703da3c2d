// barrier_kernels.cu -- external harmonic barriers: the one-body force of
// reference src/HarmonicBarrierGPU.cuh:49-132 (kernel::compute_harmonic_barrier<Evaluator>) with
// the evaluators of src/PlanarBarrierEvaluator.h:37-51 and src/SphericalBarrierEvaluator.h:36-53.
//
// A streaming kernel, HBM-bound: per particle it reads 16 bytes (position + type), and writes the
// 16-byte force/energy and the six virial scalars, which the reference zeroes with a separate
// cudaMemset (:129) -- fused here into the same pass, so every byte crosses HBM once:
// 4S (pos) + 4S (force) + 6S (virial zero) = 56 bytes per particle in fp32.
// One thread handles four consecutive particles: four 16-byte loads in flight, four 16-byte force
// stores, and one 16-byte store per virial row (rows are pitch-aligned).
//
// The arithmetic is the reference's operation for operation in IEEE round-to-nearest
// (__fmul_rn/__fadd_rn/... so that nvcc does not contract into FMAs, IEEE sqrt and division), so
// the result is bit-identical to the reference's CPU class HarmonicBarrier<Evaluator>
// (src/HarmonicBarrier.h:149-175) in both precisions; being HBM-bound, the exact ops are free.
#include "../../include/azp_b200.h"
#include "azp_core.cuh"

namespace azp
    {
// exact IEEE helpers (no FMA contraction)
AZP_D float mul_rn(float a, float b) { return __fmul_rn(a, b); }
AZP_D double mul_rn(double a, double b) { return __dmul_rn(a, b); }
AZP_D float add_rn(float a, float b) { return __fadd_rn(a, b); }
AZP_D double add_rn(double a, double b) { return __dadd_rn(a, b); }
AZP_D float div_rn(float a, float b) { return __fdiv_rn(a, b); }
AZP_D double div_rn(double a, double b) { return __ddiv_rn(a, b); }
AZP_D float sqrt_rn(float a) { return __fsqrt_rn(a); }
AZP_D double sqrt_rn(double a) { return __dsqrt_rn(a); }

// HOOMD BoxDim::wrap (v7.0.1, not in the reference tree; restated): one box image per axis, z then
// y then x, the tilt factors shifting the lower axes. The image counters are not needed here.
template<class S> struct WrapBox
    {
    S lo[3], hi[3], L[3];
    S xy, xz, yz;
    int periodic[3];
    };
template<class S> AZP_D void wrap_into_box(const WrapBox<S>& b, S& x, S& y, S& z)
    {
    if (b.periodic[2])
        {
        if (z >= b.hi[2])
            {
            z = add_rn(z, -b.L[2]);
            y = add_rn(y, -mul_rn(b.L[2], b.yz));
            x = add_rn(x, -mul_rn(b.L[2], b.xz));
            }
        else if (z < b.lo[2])
            {
            z = add_rn(z, b.L[2]);
            y = add_rn(y, mul_rn(b.L[2], b.yz));
            x = add_rn(x, mul_rn(b.L[2], b.xz));
            }
        }
    if (b.periodic[1])
        {
        const S tilt_y = mul_rn(b.yz, z);
        if (y >= add_rn(b.hi[1], tilt_y))
            {
            y = add_rn(y, -b.L[1]);
            x = add_rn(x, -mul_rn(b.L[1], b.xy));
            }
        else if (y < add_rn(b.lo[1], tilt_y))
            {
            y = add_rn(y, b.L[1]);
            x = add_rn(x, mul_rn(b.L[1], b.xy));
            }
        }
    if (b.periodic[0])
        {
        const S tilt_x = add_rn(mul_rn(add_rn(b.xz, -mul_rn(b.xy, b.yz)), z), mul_rn(b.xy, y));
        if (x >= add_rn(b.hi[0], tilt_x))
            x = add_rn(x, -b.L[0]);
        else if (x < add_rn(b.lo[0], tilt_x))
            x = add_rn(x, b.L[0]);
        }
    }

// reference src/PlanarBarrierEvaluator.h:37-51
template<class S> struct PlanarBarrier
    {
    S H;
    AZP_D void operator()(S x, S y, S z, S k, S offset, S& fx, S& fy, S& fz, S& e) const
        {
        fx = fy = fz = e = S(0);
        const S dy = add_rn(y, -add_rn(H, offset));
        if (dy <= S(0))
            return;
        const S f = mul_rn(-k, dy);
        fy = f;
        e = mul_rn(mul_rn(S(-0.5), f), dy);
        }
    };

// reference src/SphericalBarrierEvaluator.h:36-53
template<class S> struct SphericalBarrier
    {
    S R;
    AZP_D void operator()(S x, S y, S z, S k, S offset, S& fx, S& fy, S& fz, S& e) const
        {
        fx = fy = fz = e = S(0);
        const S r = sqrt_rn(add_rn(add_rn(mul_rn(x, x), mul_rn(y, y)), mul_rn(z, z)));
        const S dr = add_rn(r, -add_rn(R, offset));
        if (dr <= S(0))
            return;
        const S k_dr = mul_rn(k, dr);
        const S c = -div_rn(k_dr, r);
        fx = mul_rn(c, x);
        fy = mul_rn(c, y);
        fz = mul_rn(c, z);
        e = mul_rn(mul_rn(S(0.5), k_dr), dr);
        }
    };

template<class S> struct Scalar2T
    {
    S x, y;
    };

constexpr unsigned int kBarrierMaxTypesShared = 4096;

template<class S, class Evaluator>
__global__ void __launch_bounds__(256) barrier_kernel(S* __restrict__ force,
                                                       S* __restrict__ virial,
                                                       const size_t virial_pitch,
                                                       const S* __restrict__ pos,
                                                       const Scalar2T<S>* __restrict__ params,
                                                       const WrapBox<S> box,
                                                       const Evaluator evaluator,
                                                       const unsigned int N,
                                                       const unsigned int ntypes)
    {
    extern __shared__ __align__(16) unsigned char barrier_smem[];
    Scalar2T<S>* s_params = reinterpret_cast<Scalar2T<S>*>(barrier_smem);
    for (unsigned int t = threadIdx.x; t < ntypes; t += blockDim.x)
        s_params[t] = params[t];
    __syncthreads();

    // four consecutive particles per thread
    const unsigned int first = 4u * (blockIdx.x * blockDim.x + threadIdx.x);
    if (first >= N)
        return;
    const unsigned int count = min(4u, N - first);
    Vec4<S> p[4];
#pragma unroll
    for (unsigned int q = 0; q < 4; ++q)
        if (q < count)
            p[q] = load4(pos, first + q);
#pragma unroll
    for (unsigned int q = 0; q < 4; ++q)
        {
        if (q < count)
            {
            S x = p[q].x, y = p[q].y, z = p[q].z;
            const unsigned int type = scalar_as_uint(p[q].w);
            const Scalar2T<S> kp = s_params[type < ntypes ? type : 0u];
            wrap_into_box(box, x, y, z);
            S fx, fy, fz, e;
            evaluator(x, y, z, kp.x, kp.y, fx, fy, fz, e);
            store4(force, first + q, fx, fy, fz, e);
            }
        }
    // the barrier contributes no virial (reference :129 zeroes the array)
    if (virial)
        {
        const bool vec_ok = (count == 4u) && ((virial_pitch & 3u) == 0u)
                            && ((reinterpret_cast<uintptr_t>(virial) & (4u * sizeof(S) - 1u)) == 0u);
#pragma unroll
        for (unsigned int row = 0; row < 6; ++row)
            {
            S* v = virial + row * virial_pitch + first;
            if (vec_ok)
                store4(v, 0u, S(0), S(0), S(0), S(0));
            else
                for (unsigned int q = 0; q < count; ++q)
                    v[q] = S(0);
            }
        }
    }

template<class S> static int launch_barrier(const azp_barrier_args* a, cudaStream_t stream)
    {
    if (!a)
        return (int)cudaErrorInvalidValue;
    if (a->N == 0)
        return 0;
    if (!a->d_force || !a->d_pos || !a->d_params || a->ntypes == 0 || a->ntypes > kBarrierMaxTypesShared)
        return (int)cudaErrorInvalidValue;
    if (a->geometry != AZP_BARRIER_PLANAR && a->geometry != AZP_BARRIER_SPHERICAL)
        return (int)cudaErrorInvalidValue;
    unsigned int block = a->block_size ? a->block_size : 256u;
    if (block % 32u != 0 || block > 256u)
        return (int)cudaErrorInvalidValue;
    WrapBox<S> box;
    for (int d = 0; d < 3; ++d)
        {
        box.L[d] = S(a->box.L[d]);
        // HOOMD: lo = -L/2, hi = lo + L (in Scalar precision)
        box.lo[d] = -box.L[d] / S(2.0);
        box.hi[d] = box.lo[d] + box.L[d];
        box.periodic[d] = a->box.periodic[d];
        }
    box.xy = S(a->box.tilt[0]);
    box.xz = S(a->box.tilt[1]);
    box.yz = S(a->box.tilt[2]);
    const unsigned int threads = (a->N + 3u) / 4u;
    const unsigned int grid = (threads + block - 1u) / block;
    const size_t smem = sizeof(Scalar2T<S>) * a->ntypes;
    S* force = static_cast<S*>(a->d_force);
    S* virial = static_cast<S*>(a->d_virial);
    const S* pos = static_cast<const S*>(a->d_pos);
    const Scalar2T<S>* params = static_cast<const Scalar2T<S>*>(a->d_params);
    if (a->geometry == AZP_BARRIER_PLANAR)
        {
        PlanarBarrier<S> ev {S(a->location)};
        barrier_kernel<S, PlanarBarrier<S>><<<grid, block, smem, stream>>>(force, virial, (size_t)a->virial_pitch, pos, params, box, ev, a->N, a->ntypes);
        }
    else
        {
        SphericalBarrier<S> ev {S(a->location)};
        barrier_kernel<S, SphericalBarrier<S>><<<grid, block, smem, stream>>>(force, virial, (size_t)a->virial_pitch, pos, params, box, ev, a->N, a->ntypes);
        }
    return (int)cudaGetLastError();
    }
    } // namespace azp

extern "C"
    {
    int azp_harmonic_barrier_f32(const azp_barrier_args* args, void* stream)
        {
        return azp::launch_barrier<float>(args, (cudaStream_t)stream);
        }
    int azp_harmonic_barrier_f64(const azp_barrier_args* args, void* stream)
        {
        return azp::launch_barrier<double>(args, (cudaStream_t)stream);
        }

    // BarrierEvaluator::valid (reference src/PlanarBarrierEvaluator.h:54-59,
    // src/SphericalBarrierEvaluator.h:56-62), evaluated in Scalar precision on the host
    int azp_harmonic_barrier_valid(int geometry, int scalar_bits, double location, const azp_box* box)
        {
        if (!box)
            return 0;
        const double Lx = box->L[0], Ly = box->L[1], Lz = box->L[2];
        const double xy = box->tilt[0], xz = box->tilt[1], yz = box->tilt[2];
        if (geometry == AZP_BARRIER_PLANAR)
            {
            // makeCoordinates(0,0,0).y = lo.y, makeCoordinates(1,1,1).y = lo.y + L.y + yz * L.z ... at f=(1,1,1):
            // v = lo + f * L; v.y += yz * v.z
            double lo_y, hi_y;
            if (scalar_bits == 32)
                {
                const float ly = (float)Ly, lz = (float)Lz, t = (float)yz;
                const float loz = -lz / 2.0f, hiz = loz + lz;
                const float loy = -ly / 2.0f, hiy = loy + ly;
                lo_y = loy + t * loz;
                hi_y = hiy + t * hiz;
                return ((float)location >= (float)lo_y && (float)location < (float)hi_y) ? 1 : 0;
                }
            const double loz = -Lz / 2.0, hiz = loz + Lz;
            const double loy = -Ly / 2.0, hiy = loy + Ly;
            lo_y = loy + yz * loz;
            hi_y = hiy + yz * hiz;
            return (location >= lo_y && location < hi_y) ? 1 : 0;
            }
        if (geometry == AZP_BARRIER_SPHERICAL)
            {
            // BoxDim::getNearestPlaneDistance (HOOMD, restated): distances between opposite faces
            const double term = xy * yz - xz;
            const double dx = Lx / sqrt(1.0 + xy * xy + term * term);
            const double dy = Ly / sqrt(1.0 + yz * yz);
            const double dz = Lz;
            const double two_R = 2.0 * location;
            return (location >= 0.0 && dx >= two_R && dy >= two_R && dz >= two_R) ? 1 : 0;
            }
        return 0;
        }
    }
