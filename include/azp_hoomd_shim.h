// azp_hoomd_shim.h -- C++ shim that gives libazp_b200.so the symbols HOOMD-blue's host classes
// call, so the B200 kernels drop in behind the reference's own boundary.
//
// azplugins instantiates three HOOMD driver templates, one per evaluator
// (reference src/PotentialPairGPUKernel.cu.inc:25-28,
//  src/PotentialPairDPDThermoGPUKernel.cu.inc:21-24, src/AnisoPotentialPairGPUKernel.cu.inc:21-25):
//
//   hipError_t hoomd::md::kernel::gpu_compute_pair_forces<E>(const pair_args_t&, const E::param_type*)
//   hipError_t hoomd::md::kernel::gpu_compute_dpd_forces<E>(const dpd_pair_args_t&, const E::param_type*)
//   hipError_t hoomd::md::kernel::gpu_compute_pair_aniso_forces<E>(const a_pair_args_t&,
//                                   const E::param_type*, const E::shape_type*)
//
// This header re-declares those templates as thin inline forwarders to the C ABI
// (include/azp_b200.h). In a HOOMD build it is included INSTEAD of hoomd/md/PotentialPairGPU.cuh
// etc. by the three *.cu.inc stubs; the argument structs then come from
// HOOMD's own headers. Stand-alone (AZP_SHIM_STANDALONE, used by tests/test_shim.py) it carries
// field-for-field stand-ins of those structs as recalled in SURVEY.md 8(b), so the forwarding code
// can be compiled and exercised without HOOMD.
//
// The evaluator -> id mapping is a trait; a maintainer adds one line per evaluator class.
#ifndef AZP_HOOMD_SHIM_H_
#define AZP_HOOMD_SHIM_H_

#include "azp_b200.h"

#include <cstddef>
#include <cstdint>

#ifdef AZP_SHIM_STANDALONE
// ---- stand-ins for HOOMD types (HOOMDMath.h / BoxDim.h / PotentialPairGPU.cuh) --------------
typedef int hipError_t;
struct hipDeviceProp_t
    {
    int major, minor;
    };
namespace hoomd
    {
#if defined(HOOMD_LONGREAL_SIZE) && HOOMD_LONGREAL_SIZE == 32
typedef float Scalar;
#else
typedef double Scalar;
#endif
struct Scalar3
    {
    Scalar x, y, z;
    };
struct Scalar4
    {
    Scalar x, y, z, w;
    };
struct uchar3
    {
    unsigned char x, y, z;
    };
class BoxDim
    {
    public:
    BoxDim(Scalar Lx, Scalar Ly, Scalar Lz, Scalar xy = 0, Scalar xz = 0, Scalar yz = 0)
        : m_L {Lx, Ly, Lz}, m_xy(xy), m_xz(xz), m_yz(yz), m_periodic {1, 1, 1}
        {
        }
    Scalar3 getL() const
        {
        return m_L;
        }
    Scalar getTiltFactorXY() const
        {
        return m_xy;
        }
    Scalar getTiltFactorXZ() const
        {
        return m_xz;
        }
    Scalar getTiltFactorYZ() const
        {
        return m_yz;
        }
    uchar3 getPeriodic() const
        {
        return m_periodic;
        }
    void setPeriodic(uchar3 p)
        {
        m_periodic = p;
        }

    private:
    Scalar3 m_L;
    Scalar m_xy, m_xz, m_yz;
    uchar3 m_periodic;
    };
namespace md
    {
namespace kernel
    {
struct pair_args_t
    {
    Scalar4* d_force;
    Scalar* d_virial;
    size_t virial_pitch;
    unsigned int N;
    unsigned int n_max;
    const Scalar4* d_pos;
    const Scalar* d_charge;
    BoxDim box;
    const unsigned int* d_n_neigh;
    const unsigned int* d_nlist;
    const size_t* d_head_list;
    const Scalar* d_rcutsq;
    const Scalar* d_ronsq;
    size_t size_neigh_list;
    unsigned int ntypes;
    unsigned int block_size;
    unsigned int shift_mode;
    unsigned int compute_virial;
    unsigned int threads_per_particle;
    const hipDeviceProp_t& devprop;
    };
struct dpd_pair_args_t
    {
    Scalar4* d_force;
    Scalar* d_virial;
    size_t virial_pitch;
    unsigned int N;
    unsigned int n_max;
    const Scalar4* d_pos;
    const Scalar4* d_vel;
    const unsigned int* d_tag;
    BoxDim box;
    const unsigned int* d_n_neigh;
    const unsigned int* d_nlist;
    const size_t* d_head_list;
    const Scalar* d_rcutsq;
    size_t size_neigh_list;
    unsigned int ntypes;
    unsigned int block_size;
    uint16_t seed;
    uint64_t timestep;
    Scalar deltaT;
    Scalar T;
    unsigned int shift_mode;
    unsigned int compute_virial;
    unsigned int threads_per_particle;
    const hipDeviceProp_t& devprop;
    };
struct a_pair_args_t
    {
    Scalar4* d_force;
    Scalar4* d_torque;
    Scalar* d_virial;
    size_t virial_pitch;
    unsigned int N;
    unsigned int n_max;
    const Scalar4* d_pos;
    const Scalar* d_charge;
    const Scalar4* d_orientation;
    const unsigned int* d_tag;
    BoxDim box;
    const unsigned int* d_n_neigh;
    const unsigned int* d_nlist;
    const size_t* d_head_list;
    const Scalar* d_rcutsq;
    unsigned int ntypes;
    unsigned int block_size;
    unsigned int shift_mode;
    unsigned int compute_virial;
    unsigned int threads_per_particle;
    const hipDeviceProp_t& devprop;
    };
    } // namespace kernel
    } // namespace md
    } // namespace hoomd
#endif // AZP_SHIM_STANDALONE

namespace azp_shim
    {
// evaluator class -> azp_evaluator id; specialise once per evaluator (see bottom of this file)
template<class Evaluator> struct evaluator_id;

template<class Box> inline azp_box flatten_box(const Box& box)
    {
    azp_box b;
    const auto L = box.getL();
    b.L[0] = L.x, b.L[1] = L.y, b.L[2] = L.z;
    b.tilt[0] = box.getTiltFactorXY();
    b.tilt[1] = box.getTiltFactorXZ();
    b.tilt[2] = box.getTiltFactorYZ();
    const auto p = box.getPeriodic();
    b.periodic[0] = p.x, b.periodic[1] = p.y, b.periodic[2] = p.z;
    b._pad = 0;
    return b;
    }

template<class Args> inline azp_pair_args common_args(const Args& a)
    {
    static_assert(sizeof(size_t) == sizeof(uint64_t), "head_list entries are 64-bit");
    azp_pair_args o = {};
    o.d_force = a.d_force;
    o.d_virial = a.d_virial;
    o.virial_pitch = a.virial_pitch;
    o.d_pos = a.d_pos;
    o.d_n_neigh = a.d_n_neigh;
    o.d_nlist = a.d_nlist;
    o.d_head_list = reinterpret_cast<const uint64_t*>(a.d_head_list);
    o.d_rcutsq = a.d_rcutsq;
    o.box = flatten_box(a.box);
    o.N = a.N;
    o.n_max = a.n_max;
    o.ntypes = a.ntypes;
    o.shift_mode = a.shift_mode;
    o.compute_virial = a.compute_virial;
    o.block_size = a.block_size;
    o.threads_per_particle = a.threads_per_particle;
    return o;
    }
    } // namespace azp_shim

namespace hoomd
    {
namespace md
    {
namespace kernel
    {
//! Drop-in for HOOMD's isotropic pair-force driver
template<class evaluator>
inline hipError_t gpu_compute_pair_forces(const pair_args_t& pair_args,
                                          const typename evaluator::param_type* d_params)
    {
    azp_pair_args a = azp_shim::common_args(pair_args);
    a.d_ronsq = pair_args.d_ronsq;
    a.size_neigh_list = pair_args.size_neigh_list;
    const int id = azp_shim::evaluator_id<evaluator>::value;
    const int rc = sizeof(Scalar) == 4 ? azp_pair_forces_f32(id, &a, d_params, nullptr)
                                       : azp_pair_forces_f64(id, &a, d_params, nullptr);
    return static_cast<hipError_t>(rc);
    }

//! Drop-in for HOOMD's DPD thermostat driver
template<class evaluator>
inline hipError_t gpu_compute_dpd_forces(const dpd_pair_args_t& args,
                                         const typename evaluator::param_type* d_params)
    {
    azp_pair_args a = azp_shim::common_args(args);
    a.size_neigh_list = args.size_neigh_list;
    a.d_vel = args.d_vel;
    a.d_tag = args.d_tag;
    a.seed = args.seed;
    a.timestep = args.timestep;
    a.deltaT = args.deltaT;
    a.T = args.T;
    const int id = azp_shim::evaluator_id<evaluator>::value;
    const int rc = sizeof(Scalar) == 4 ? azp_dpd_forces_f32(id, &a, d_params, nullptr)
                                       : azp_dpd_forces_f64(id, &a, d_params, nullptr);
    return static_cast<hipError_t>(rc);
    }

//! Drop-in for HOOMD's anisotropic pair-force driver
template<class evaluator>
inline hipError_t gpu_compute_pair_aniso_forces(const a_pair_args_t& pair_args,
                                                const typename evaluator::param_type* d_params,
                                                const typename evaluator::shape_type* d_shape_params)
    {
    azp_pair_args a = azp_shim::common_args(pair_args);
    a.d_torque = pair_args.d_torque;
    a.d_orientation = pair_args.d_orientation;
    a.d_tag = pair_args.d_tag;
    const int id = azp_shim::evaluator_id<evaluator>::value;
    const int rc = sizeof(Scalar) == 4
                       ? azp_aniso_forces_f32(id, &a, d_params, d_shape_params, nullptr)
                       : azp_aniso_forces_f64(id, &a, d_params, d_shape_params, nullptr);
    return static_cast<hipError_t>(rc);
    }
    } // namespace kernel
    } // namespace md
    } // namespace hoomd

// One line per reference evaluator class (forward-declared; the definitions stay in the
// reference's own headers, src/PairEvaluator*.h etc.).
namespace hoomd
    {
namespace azplugins
    {
namespace detail
    {
class PairEvaluatorPerturbedLennardJones;
class PairEvaluatorExpandedYukawa;
class PairEvaluatorColloid;
class PairEvaluatorHertz;
class DPDPairEvaluatorGeneralWeight;
class AnisoPairEvaluatorTwoPatchMorse;
    } // namespace detail
    } // namespace azplugins
    } // namespace hoomd

#define AZP_SHIM_EVALUATOR(cls, id)                                    \
    template<> struct azp_shim::evaluator_id<hoomd::azplugins::detail::cls> \
        {                                                              \
        static constexpr int value = id;                               \
        }
AZP_SHIM_EVALUATOR(PairEvaluatorPerturbedLennardJones, AZP_EV_PERTURBED_LENNARD_JONES);
AZP_SHIM_EVALUATOR(PairEvaluatorExpandedYukawa, AZP_EV_EXPANDED_YUKAWA);
AZP_SHIM_EVALUATOR(PairEvaluatorColloid, AZP_EV_COLLOID);
AZP_SHIM_EVALUATOR(PairEvaluatorHertz, AZP_EV_HERTZ);
AZP_SHIM_EVALUATOR(DPDPairEvaluatorGeneralWeight, AZP_EV_DPD_GENERAL_WEIGHT);
AZP_SHIM_EVALUATOR(AnisoPairEvaluatorTwoPatchMorse, AZP_EV_TWO_PATCH_MORSE);

#endif // AZP_HOOMD_SHIM_H_
