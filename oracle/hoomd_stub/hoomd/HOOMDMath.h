// TEST INFRASTRUCTURE (oracle). Not part of the product path.
//
// Host-only stand-in for the parts of HOOMD-blue v7.0.1 `hoomd/HOOMDMath.h` that the azplugins
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

inline Scalar2 make_scalar2(Scalar x, Scalar y)
    {
    return Scalar2 {x, y};
    }
inline Scalar3 make_scalar3(Scalar x, Scalar y, Scalar z)
    {
    return Scalar3 {x, y, z};
    }
inline Scalar4 make_scalar4(Scalar x, Scalar y, Scalar z, Scalar w)
    {
    return Scalar4 {x, y, z, w};
    }

inline Scalar dot(const Scalar3& a, const Scalar3& b)
    {
    return a.x * b.x + a.y * b.y + a.z * b.z;
    }

// Host mappings of HOOMD's fast:: namespace (Appendix A.5): on the host these are the plain libm
// calls in Scalar precision; rsqrt is 1/sqrt.
namespace fast
    {
inline float sqrt(float x)
    {
    return ::sqrtf(x);
    }
inline double sqrt(double x)
    {
    return ::sqrt(x);
    }
inline float rsqrt(float x)
    {
    return 1.0f / ::sqrtf(x);
    }
inline double rsqrt(double x)
    {
    return 1.0 / ::sqrt(x);
    }
inline float exp(float x)
    {
    return ::expf(x);
    }
inline double exp(double x)
    {
    return ::exp(x);
    }
inline float pow(float x, float y)
    {
    return ::powf(x, y);
    }
inline double pow(double x, double y)
    {
    return ::pow(x, y);
    }
    } // namespace fast
    } // namespace hoomd

#endif
