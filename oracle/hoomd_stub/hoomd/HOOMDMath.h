// TEST INFRASTRUCTURE (oracle). Not part of the product path.
//
// Stand-in (host, and device when compiled by nvcc: tests/contract) for the parts of HOOMD-blue v7.0.1 `hoomd/HOOMDMath.h` that the azplugins
// evaluator headers use (`Scalar`, `Scalar2/3/4`, `make_scalar*`, `dot`, `fast::`). HOOMD is not
// installed in this container and is not vendored by the reference, so this file restates the
// published behaviour; it is written from scratch.
//
// It lets `/root/reference/src/PairEvaluator*.h` etc. compile *unmodified, in place* for the
// oracle/_ref build (see oracle/Makefile). Precision is selected with HOOMD_LONGREAL_SIZE (32|64),
// the same macro the reference headers switch their alignment on
// (reference src/PairEvaluatorColloid.h:53-57).
#ifndef AZP_ORACLE_STUB_HOOMDMATH_H_
#define AZP_ORACLE_STUB_HOOMDMATH_H_

#include <cmath>
#include <cstdint>
#include <math.h> // HOOMD brings in math.h, so unqualified sqrt/log/pow see the float overloads
#include <stdexcept>
#include <stdint.h>

#ifndef HOOMD_LONGREAL_SIZE
#define HOOMD_LONGREAL_SIZE 64
#endif

// Under nvcc every function here is also a device function: the reference's evaluator headers
// then compile for the GPU (-D__HIPCC__ -DNVCC) and ride the product kernels through the
// ContractEvaluator adapters (tests/contract/contract_lib.cu).
#ifdef __CUDACC__
#define AZP_STUB_HD __host__ __device__
#else
#define AZP_STUB_HD
#endif

namespace hoomd
    {
#if HOOMD_LONGREAL_SIZE == 32
typedef float Scalar;
#else
typedef double Scalar;
#endif

struct Scalar2
    {
    Scalar x, y;
    };
struct Scalar3
    {
    Scalar x, y, z;
    };
struct Scalar4
    {
    Scalar x, y, z, w;
    };

AZP_STUB_HD inline Scalar2 make_scalar2(Scalar x, Scalar y)
    {
    return Scalar2 {x, y};
    }
AZP_STUB_HD inline Scalar3 make_scalar3(Scalar x, Scalar y, Scalar z)
    {
    return Scalar3 {x, y, z};
    }
AZP_STUB_HD inline Scalar4 make_scalar4(Scalar x, Scalar y, Scalar z, Scalar w)
    {
    return Scalar4 {x, y, z, w};
    }

AZP_STUB_HD inline Scalar dot(const Scalar3& a, const Scalar3& b)
    {
    return a.x * b.x + a.y * b.y + a.z * b.z;
    }

// Host mappings of HOOMD's fast:: namespace (Appendix A.5): on the host these are the plain libm
// calls in Scalar precision; rsqrt is 1/sqrt.
namespace fast
    {
AZP_STUB_HD inline float sqrt(float x)
    {
    return ::sqrtf(x);
    }
AZP_STUB_HD inline double sqrt(double x)
    {
    return ::sqrt(x);
    }
AZP_STUB_HD inline float rsqrt(float x)
    {
#ifdef __CUDA_ARCH__
    return ::rsqrtf(x); // HOOMD's device mapping (Appendix A.5)
#else
    return 1.0f / ::sqrtf(x);
#endif
    }
AZP_STUB_HD inline double rsqrt(double x)
    {
#ifdef __CUDA_ARCH__
    return ::rsqrt(x);
#else
    return 1.0 / ::sqrt(x);
#endif
    }
AZP_STUB_HD inline float exp(float x)
    {
#ifdef __CUDA_ARCH__
    return ::__expf(x);
#else
    return ::expf(x);
#endif
    }
AZP_STUB_HD inline double exp(double x)
    {
    return ::exp(x);
    }
AZP_STUB_HD inline float pow(float x, float y)
    {
#ifdef __CUDA_ARCH__
    return ::__powf(x, y);
#else
    return ::powf(x, y);
#endif
    }
AZP_STUB_HD inline double pow(double x, double y)
    {
    return ::pow(x, y);
    }
    } // namespace fast
    } // namespace hoomd

#endif
