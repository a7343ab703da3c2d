// azp_philox.cuh -- counter-based RNG of the DPD thermostat.
//
// Philox4x32-10 (Salmon et al., SC'11; the Random123 engine HOOMD-blue vendors) and the
// HOOMD v7.0.1 RandomGenerator keying, written from scratch. One draw per
// pair, keyed exactly like reference src/DPDPairEvaluatorGeneralWeight.h:213-233:
//   Seed(RNGIdentifier::DPDEvaluatorGeneralWeight = 200 (src/RNGIdentifiers.h:23),
//        timestep (32-bit, :121-122), seed16)  -> key  = { 200<<24 | seed<<8 | 0, timestep }
//   Counter(min(tag_i,tag_j), max(tag_i,tag_j)) -> ctr  = { 0, 0, max, min }
//   alpha = -1 + 2 * u01(first draw)
// Symmetric in (i, j), so both directions of a pair in a full neighbour list see the same alpha.
#ifndef AZP_PHILOX_CUH_
#define AZP_PHILOX_CUH_

#include "azp_core.cuh"

namespace azp
    {
struct Philox4
    {
    uint32_t v[4];
    };

AZP_HD void mulhilo32(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo)
    {
#ifdef __CUDA_ARCH__
    lo = a * b;
    hi = __umulhi(a, b);
#else
    const uint64_t p = (uint64_t)a * b;
    lo = (uint32_t)p;
    hi = (uint32_t)(p >> 32);
#endif
    }

AZP_HD Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1)
    {
#pragma unroll
    for (int r = 0; r < 10; ++r)
        {
        uint32_t hi0, lo0, hi1, lo1;
        mulhilo32(0xD2511F53u, c0, hi0, lo0);
        mulhilo32(0xCD9E8D57u, c2, hi1, lo1);
        c0 = hi1 ^ c1 ^ k0;
        c2 = hi0 ^ c3 ^ k1;
        c1 = lo1;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
        }
    Philox4 out;
    out.v[0] = c0;
    out.v[1] = c1;
    out.v[2] = c2;
    out.v[3] = c3;
    return out;
    }

// Random123 u01: x * 2^-W + 2^-(W+1), W = 32 (float, from v[0]) or 64 (double, v[0]<<32 | v[1])
AZP_HD float u01_from(const Philox4& u, float)
    {
    return (float)u.v[0] * 2.3283064365386963e-10f + 1.1641532182693481e-10f;
    }
AZP_HD double u01_from(const Philox4& u, double)
    {
    const uint64_t x = ((uint64_t)u.v[0] << 32) | u.v[1];
    return (double)x * 5.421010862427522e-20 + 2.710505431213761e-20;
    }

template<class S> AZP_HD S dpd_uniform_pm1(uint32_t seed16, uint32_t tag_i, uint32_t tag_j, uint32_t timestep32)
    {
    const uint32_t lo = tag_i > tag_j ? tag_j : tag_i;
    const uint32_t hi = tag_i > tag_j ? tag_i : tag_j;
    const uint32_t k0 = (200u << 24) | ((seed16 & 0xffffu) << 8);
    const Philox4 u = philox4x32_10(0u, 0u, hi, lo, k0, timestep32);
    return S(-1) + S(2) * u01_from(u, S());
    }
    } // namespace azp

#endif
