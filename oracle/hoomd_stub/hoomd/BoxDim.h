// TEST INFRASTRUCTURE (oracle). Not part of the product path.
//
// Host-only stand-in for the part of HOOMD-blue v7.0.1 `hoomd/BoxDim.h` that the reference's
// barrier evaluators and HarmonicBarrier::computeForces use (src/PlanarBarrierEvaluator.h:54-59,
// src/SphericalBarrierEvaluator.h:56-62, src/HarmonicBarrier.h:160-170): a centred triclinic box
// with makeCoordinates, getNearestPlaneDistance and wrap. HOOMD is not in the reference tree, so
// this restates the published behaviour from scratch (recalled; the reference's own tests only
// exercise cubic boxes, so the triclinic branches are parity-unpinned).
#ifndef AZP_ORACLE_STUB_BOXDIM_H_
#define AZP_ORACLE_STUB_BOXDIM_H_

#include "HOOMDMath.h"

namespace hoomd
    {
struct int3
    {
    int x, y, z;
    };
struct char3
    {
    signed char x, y, z;
    };
inline int3 make_int3(int x, int y, int z)
    {
    return int3 {x, y, z};
    }
inline char3 make_char3(signed char x, signed char y, signed char z)
    {
    return char3 {x, y, z};
    }
inline Scalar3 operator*(Scalar a, const Scalar3& v)
    {
    return Scalar3 {a * v.x, a * v.y, a * v.z};
    }

class BoxDim
    {
    public:
    BoxDim(Scalar Lx, Scalar Ly, Scalar Lz, Scalar xy = 0, Scalar xz = 0, Scalar yz = 0)
        : m_xy(xy), m_xz(xz), m_yz(yz)
        {
        m_L = make_scalar3(Lx, Ly, Lz);
        m_lo = make_scalar3(-Lx / Scalar(2.0), -Ly / Scalar(2.0), -Lz / Scalar(2.0));
        m_hi = make_scalar3(m_lo.x + Lx, m_lo.y + Ly, m_lo.z + Lz);
        m_periodic[0] = m_periodic[1] = m_periodic[2] = 1;
        }
    void setPeriodic(int px, int py, int pz)
        {
        m_periodic[0] = px, m_periodic[1] = py, m_periodic[2] = pz;
        }
    Scalar3 getL() const
        {
        return m_L;
        }
    Scalar3 makeCoordinates(const Scalar3& f) const
        {
        Scalar3 v = make_scalar3(m_lo.x + f.x * m_L.x, m_lo.y + f.y * m_L.y, m_lo.z + f.z * m_L.z);
        v.x += m_xy * v.y + m_xz * v.z;
        v.y += m_yz * v.z;
        return v;
        }
    Scalar3 getNearestPlaneDistance() const
        {
        const Scalar term = m_xy * m_yz - m_xz;
        Scalar3 d;
        d.x = m_L.x / fast::sqrt(Scalar(1.0) + m_xy * m_xy + term * term);
        d.y = m_L.y / fast::sqrt(Scalar(1.0) + m_yz * m_yz);
        d.z = m_L.z;
        return d;
        }
    // one image per axis: z, then y, then x; the tilt factors shift the lower axes
    void wrap(Scalar3& w, int3& img, char3 flags = make_char3(0, 0, 0)) const
        {
        if (m_periodic[2])
            {
            if (((w.z >= m_hi.z) && !flags.z) || flags.z == 1)
                {
                w.z -= m_L.z;
                w.y -= m_L.z * m_yz;
                w.x -= m_L.z * m_xz;
                img.z++;
                }
            else if (((w.z < m_lo.z) && !flags.z) || flags.z == -1)
                {
                w.z += m_L.z;
                w.y += m_L.z * m_yz;
                w.x += m_L.z * m_xz;
                img.z--;
                }
            }
        if (m_periodic[1])
            {
            const Scalar tilt_y = m_yz * w.z;
            if (((w.y >= m_hi.y + tilt_y) && !flags.y) || flags.y == 1)
                {
                w.y -= m_L.y;
                w.x -= m_L.y * m_xy;
                img.y++;
                }
            else if (((w.y < m_lo.y + tilt_y) && !flags.y) || flags.y == -1)
                {
                w.y += m_L.y;
                w.x += m_L.y * m_xy;
                img.y--;
                }
            }
        if (m_periodic[0])
            {
            const Scalar tilt_x = (m_xz - m_xy * m_yz) * w.z + m_xy * w.y;
            if (((w.x >= m_hi.x + tilt_x) && !flags.x) || flags.x == 1)
                {
                w.x -= m_L.x;
                img.x++;
                }
            else if (((w.x < m_lo.x + tilt_x) && !flags.x) || flags.x == -1)
                {
                w.x += m_L.x;
                img.x--;
                }
            }
        }

    private:
    Scalar3 m_lo, m_hi, m_L;
    Scalar m_xy, m_xz, m_yz;
    int m_periodic[3];
    };
    } // namespace hoomd

#endif
