// TEST INFRASTRUCTURE (oracle). Not part of the product path.
//
// Stand-in (host + device under nvcc) for the parts of HOOMD-blue v7.0.1 `hoomd/VectorMath.h` used by
// reference src/AnisoPairEvaluatorTwoPatchMorse.h:141-146,186-192,210-212: vec3<Real>,
// quat<Real>, rotate, cross, dot, the arithmetic operators and vec_to_scalar3. Written from
// scratch from the documented algebra:
//   orientation Scalar4 (x,y,z,w) holds the quaternion (s, v.x, v.y, v.z)
//   rotate(q, a) = (s^2 - |v|^2) a + 2 s (v x a) + 2 (v . a) v
#ifndef AZP_ORACLE_STUB_VECTORMATH_H_
#define AZP_ORACLE_STUB_VECTORMATH_H_

#include "HOOMDMath.h"

namespace hoomd
    {
template<class Real> struct vec3
    {
    AZP_STUB_HD vec3() : x(0), y(0), z(0) { }
    AZP_STUB_HD vec3(Real _x, Real _y, Real _z) : x(_x), y(_y), z(_z) { }
    AZP_STUB_HD explicit vec3(const Scalar3& a) : x(Real(a.x)), y(Real(a.y)), z(Real(a.z)) { }
    Real x, y, z;
    };

template<class Real> AZP_STUB_HD inline vec3<Real> operator+(const vec3<Real>& a, const vec3<Real>& b)
    {
    return vec3<Real>(a.x + b.x, a.y + b.y, a.z + b.z);
    }
template<class Real> AZP_STUB_HD inline vec3<Real> operator-(const vec3<Real>& a, const vec3<Real>& b)
    {
    return vec3<Real>(a.x - b.x, a.y - b.y, a.z - b.z);
    }
template<class Real> AZP_STUB_HD inline vec3<Real> operator-(const vec3<Real>& a)
    {
    return vec3<Real>(-a.x, -a.y, -a.z);
    }
template<class Real> AZP_STUB_HD inline vec3<Real> operator*(const vec3<Real>& a, const Real& b)
    {
    return vec3<Real>(a.x * b, a.y * b, a.z * b);
    }
template<class Real> AZP_STUB_HD inline vec3<Real> operator*(const Real& b, const vec3<Real>& a)
    {
    return vec3<Real>(a.x * b, a.y * b, a.z * b);
    }
template<class Real> AZP_STUB_HD inline Real dot(const vec3<Real>& a, const vec3<Real>& b)
    {
    return a.x * b.x + a.y * b.y + a.z * b.z;
    }
template<class Real> AZP_STUB_HD inline vec3<Real> cross(const vec3<Real>& a, const vec3<Real>& b)
    {
    return vec3<Real>(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
    }

template<class Real> struct quat
    {
    AZP_STUB_HD quat() : s(1), v(0, 0, 0) { }
    AZP_STUB_HD quat(Real _s, const vec3<Real>& _v) : s(_s), v(_v) { }
    AZP_STUB_HD explicit quat(const Scalar4& a) : s(Real(a.x)), v(Real(a.y), Real(a.z), Real(a.w)) { }
    Real s;
    vec3<Real> v;
    };

template<class Real> AZP_STUB_HD inline vec3<Real> rotate(const quat<Real>& a, const vec3<Real>& b)
    {
    return (a.s * a.s - dot(a.v, a.v)) * b + (Real(2) * a.s) * cross(a.v, b)
           + (Real(2) * dot(a.v, b)) * a.v;
    }

template<class Real> AZP_STUB_HD inline Scalar3 vec_to_scalar3(const vec3<Real>& a)
    {
    return make_scalar3(Scalar(a.x), Scalar(a.y), Scalar(a.z));
    }
    } // namespace hoomd

#endif
