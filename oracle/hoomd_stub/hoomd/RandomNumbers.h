// TEST INFRASTRUCTURE (oracle). Not part of the product path.
//
// Stand-in (host + device under nvcc) for the parts of HOOMD-blue v7.0.1 `hoomd/RandomNumbers.h` (+ the Random123
// Philox4x32-10 engine HOOMD vendors) used by reference
// src/DPDPairEvaluatorGeneralWeight.h:226-233. Neither HOOMD nor Random123 is in the reference
// tree, so this restates their published behaviour from scratch:
//   * Philox4x32-10: multipliers 0xD2511F53 / 0xCD9E8D57, Weyl 0x9E3779B9 / 0xBB67AE85
//     (known-answer vectors: tests/test_philox.py);
//   * Seed(id:u8, timestep:u64, seed:u16) -> key = { id<<24 | seed<<8 | (timestep>>32)&0xff,
//                                                    timestep & 0xffffffff };
//   * Counter(a,b,c,d:u16) -> ctr = { d<<16, c, b, a };  each draw bumps ctr[0];
//   * canonical float  = u01<float>(v[0]),  double = u01<double>(v[0]<<32 | v[1]),
//     u01(x) = x * 2^-W + 2^-(W+1);
//   * UniformDistribution(a,b)(rng) = a + (b-a) * canonical.
// PARITY UNPINNED: no reference test fixes the DPD random stream bit-for-bit
// (reference src/pytest/test_pair_dpd.py:13-46 is statistical), so this packing is the
// definition the GPU path is held to.
#ifndef AZP_ORACLE_STUB_RANDOMNUMBERS_H_
#define AZP_ORACLE_STUB_RANDOMNUMBERS_H_

#include "HOOMDMath.h"
#include <cstdint>

namespace hoomd
    {
namespace detail
    {
struct philox_u4
    {
    uint32_t v[4];
    };
struct philox_u2
    {
    uint32_t v[2];
    };

AZP_STUB_HD inline philox_u4 philox4x32_10(philox_u4 ctr, philox_u2 key)
    {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    const uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    for (int round = 0; round < 10; ++round)
        {
        const uint64_t p0 = uint64_t(M0) * ctr.v[0];
        const uint64_t p1 = uint64_t(M1) * ctr.v[2];
        philox_u4 nxt;
        nxt.v[0] = uint32_t(p1 >> 32) ^ ctr.v[1] ^ key.v[0];
        nxt.v[1] = uint32_t(p1);
        nxt.v[2] = uint32_t(p0 >> 32) ^ ctr.v[3] ^ key.v[1];
        nxt.v[3] = uint32_t(p0);
        ctr = nxt;
        key.v[0] += W0;
        key.v[1] += W1;
        }
    return ctr;
    }
    } // namespace detail

class Seed
    {
    public:
    AZP_STUB_HD Seed(uint8_t id, uint64_t timestep, uint16_t seed)
        {
        m_key.v[0] = (uint32_t(id) << 24) | (uint32_t(seed) << 8)
                     | uint32_t((timestep & 0x000000ff00000000ull) >> 32);
        m_key.v[1] = uint32_t(timestep & 0x00000000ffffffffull);
        }
    AZP_STUB_HD const detail::philox_u2& getKey() const
        {
        return m_key;
        }

    private:
    detail::philox_u2 m_key;
    };

class Counter
    {
    public:
    AZP_STUB_HD Counter(uint32_t a = 0, uint32_t b = 0, uint32_t c = 0, uint16_t d = 0)
        {
        m_ctr.v[0] = uint32_t(d) << 16;
        m_ctr.v[1] = c;
        m_ctr.v[2] = b;
        m_ctr.v[3] = a;
        }
    AZP_STUB_HD const detail::philox_u4& getCounter() const
        {
        return m_ctr;
        }

    private:
    detail::philox_u4 m_ctr;
    };

class RandomGenerator
    {
    public:
    AZP_STUB_HD RandomGenerator(const Seed& seed, const Counter& counter)
        : m_key(seed.getKey()), m_ctr(counter.getCounter())
        {
        }
    AZP_STUB_HD detail::philox_u4 operator()()
        {
        detail::philox_u4 u = detail::philox4x32_10(m_ctr, m_key);
        m_ctr.v[0] += 1;
        return u;
        }

    private:
    detail::philox_u2 m_key;
    detail::philox_u4 m_ctr;
    };

namespace detail
    {
AZP_STUB_HD inline uint32_t generate_u32(RandomGenerator& rng)
    {
    return rng().v[0];
    }
AZP_STUB_HD inline uint64_t generate_u64(RandomGenerator& rng)
    {
    philox_u4 u = rng();
    return (uint64_t(u.v[0]) << 32) | u.v[1];
    }
template<class Real> AZP_STUB_HD inline Real generate_canonical(RandomGenerator& rng);
template<> AZP_STUB_HD inline float generate_canonical<float>(RandomGenerator& rng)
    {
    const float factor = 1.0f / (4294967295.0f + 1.0f); // 2^-32
    const float halffactor = 0.5f * factor;
    return float(generate_u32(rng)) * factor + halffactor;
    }
template<> AZP_STUB_HD inline double generate_canonical<double>(RandomGenerator& rng)
    {
    const double factor = 1.0 / (18446744073709551615.0 + 1.0); // 2^-64
    const double halffactor = 0.5 * factor;
    return double(generate_u64(rng)) * factor + halffactor;
    }
    } // namespace detail

template<class Real> class UniformDistribution
    {
    public:
    AZP_STUB_HD UniformDistribution(Real a = Real(0), Real b = Real(1)) : m_a(a), m_width(b - a) { }
    AZP_STUB_HD Real operator()(RandomGenerator& rng)
        {
        return m_a + m_width * detail::generate_canonical<Real>(rng);
        }

    private:
    Real m_a, m_width;
    };
    } // namespace hoomd

#endif
