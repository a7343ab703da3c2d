// Test driver for include/azp_hoomd_shim.h: plays the role of a HOOMD host class
// (PotentialPairGPU<E>::computeForces) calling the driver template azplugins
// instantiates. The evaluator classes here are minimal stand-ins carrying only `param_type`
// (what the shim needs); in a real build they are the reference's own headers.
#define AZP_SHIM_STANDALONE
#define HOOMD_LONGREAL_SIZE 64
#include "azp_hoomd_shim.h"

#include <cstdio>
#include <cuda_runtime_api.h>
#include <vector>

namespace hoomd
    {
namespace azplugins
    {
namespace detail
    {
class PairEvaluatorHertz
    {
    public:
    struct param_type
        {
        Scalar epsilon;
        };
    };
class DPDPairEvaluatorGeneralWeight
    {
    public:
    // reference src/DPDPairEvaluatorGeneralWeight.h:32-62: {A, gamma, s}, aligned to 4 Scalars
    struct alignas(4 * sizeof(Scalar)) param_type
        {
        Scalar A, gamma, s;
        };
    };
class AnisoPairEvaluatorTwoPatchMorse
    {
    public:
    struct param_type
        {
        Scalar M_d, M_rinv, r_eq, omega, alpha;
        bool repulsion;
        };
    struct shape_type
        {
        };
    };
    } // namespace detail
    } // namespace azplugins
    } // namespace hoomd

using namespace hoomd;

template<class T> static T* upload(const std::vector<T>& v)
    {
    T* d = nullptr;
    cudaMalloc(reinterpret_cast<void**>(&d), sizeof(T) * v.size());
    cudaMemcpy(d, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice);
    return d;
    }

int main(int argc, char** argv)
    {
    if (argc > 1 && argv[1][0] == 'l') // "link": prove the symbols resolve, no device work
        {
        std::printf("linked abi %d\n", azp_abi_version());
        return 0;
        }
    hipDeviceProp_t prop {10, 0};
    int worst = 0;
        {
        // reference KAT (src/pytest/test_pair.py:177-186): Hertz epsilon=2, r_cut=1.5, d=1.05
        //   -> U = 0.0985, F = 0.5477, through gpu_compute_pair_forces<E>
        const double d = 1.05;
        std::vector<Scalar4> pos = {{-d / 2, 0, 0, 0}, {d / 2, 0, 0, 0}};
        std::vector<unsigned int> n_neigh = {1, 1}, nlist = {1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        std::vector<size_t> head = {0, 8};
        std::vector<Scalar> rcutsq = {1.5 * 1.5}, ronsq = {0.0};
        std::vector<azplugins::detail::PairEvaluatorHertz::param_type> params = {{2.0}};
        Scalar4* d_force = nullptr;
        Scalar* d_virial = nullptr;
        cudaMalloc(reinterpret_cast<void**>(&d_force), sizeof(Scalar4) * 2);
        cudaMalloc(reinterpret_cast<void**>(&d_virial), sizeof(Scalar) * 12);
        md::kernel::pair_args_t args {d_force, d_virial, 2, 2, 0, upload(pos), nullptr,
                                      BoxDim(7.98, 7.98, 7.98), upload(n_neigh), upload(nlist),
                                      upload(head), upload(rcutsq), upload(ronsq), nlist.size(), 1, 128,
                                      0, 1, 1, prop};
        hipError_t rc = md::kernel::gpu_compute_pair_forces<azplugins::detail::PairEvaluatorHertz>(
            args, upload(params));
        cudaDeviceSynchronize();
        std::vector<Scalar4> f(2);
        cudaMemcpy(f.data(), d_force, sizeof(Scalar4) * 2, cudaMemcpyDeviceToHost);
        std::printf("hertz rc %d f0 %.6f %.6f %.6f e0 %.6f f1 %.6f e1 %.6f\n", rc, f[0].x, f[0].y, f[0].z,
                    f[0].w, f[1].x, f[1].w);
        worst |= rc;
        }
        {
        // reference KAT (src/pytest/test_pair.py:76-85): DPDGeneralWeight A=2, gamma=4.5, s=0.5,
        // r_cut=1, d=0.5, kT=0 (no random force), particles at rest -> U = 0.25, F = 1.0,
        // through gpu_compute_dpd_forces<E>
        const double d = 0.5;
        std::vector<Scalar4> pos = {{-d / 2, 0, 0, 0}, {d / 2, 0, 0, 0}};
        std::vector<Scalar4> vel = {{0, 0, 0, 1}, {0, 0, 0, 1}};
        std::vector<unsigned int> tag = {0, 1};
        std::vector<unsigned int> n_neigh = {1, 1}, nlist = {1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        std::vector<size_t> head = {0, 8};
        std::vector<Scalar> rcutsq = {1.0};
        std::vector<azplugins::detail::DPDPairEvaluatorGeneralWeight::param_type> params(1);
        params[0].A = 2.0, params[0].gamma = 4.5, params[0].s = 0.5;
        Scalar4* d_force = nullptr;
        Scalar* d_virial = nullptr;
        cudaMalloc(reinterpret_cast<void**>(&d_force), sizeof(Scalar4) * 2);
        cudaMalloc(reinterpret_cast<void**>(&d_virial), sizeof(Scalar) * 12);
        md::kernel::dpd_pair_args_t args {d_force, d_virial, 2, 2, 0, upload(pos), upload(vel),
                                          upload(tag), BoxDim(7.98, 7.98, 7.98), upload(n_neigh),
                                          upload(nlist), upload(head), upload(rcutsq), nlist.size(),
                                          1, 128, 42, 0, 0.005, 0.0, 0, 1, 1, prop};
        hipError_t rc = md::kernel::gpu_compute_dpd_forces<azplugins::detail::DPDPairEvaluatorGeneralWeight>(
            args, upload(params));
        cudaDeviceSynchronize();
        std::vector<Scalar4> f(2);
        cudaMemcpy(f.data(), d_force, sizeof(Scalar4) * 2, cudaMemcpyDeviceToHost);
        std::printf("dpd rc %d f0 %.6f %.6f %.6f e0 %.6f f1 %.6f e1 %.6f\n", rc, f[0].x, f[0].y, f[0].z,
                    f[0].w, f[1].x, f[1].w);
        worst |= rc;
        }
        {
        // reference KAT (src/pytest/test_pair_aniso.py:22-40,116-171): TwoPatchMorse M_d=1.8341,
        // M_r=0.0302, r_eq=1.0043, omega=5, alpha=0.4, no repulsion, r_cut=1.6, positions
        // -/+(0.5, 0.10, 0.15), identity orientations -> U = -0.41134,
        // F_0 = (11.75766, 2.46991, 3.70487), T_0 = T_1 = (0, -0.08879, 0.05919), through
        // gpu_compute_pair_aniso_forces<E>
        std::vector<Scalar4> pos = {{-0.5, -0.10, -0.15, 0}, {0.5, 0.10, 0.15, 0}};
        std::vector<Scalar4> quat = {{1, 0, 0, 0}, {1, 0, 0, 0}};
        std::vector<unsigned int> tag = {0, 1};
        std::vector<unsigned int> n_neigh = {1, 1}, nlist = {1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        std::vector<size_t> head = {0, 8};
        std::vector<Scalar> rcutsq = {1.6 * 1.6};
        std::vector<azplugins::detail::AnisoPairEvaluatorTwoPatchMorse::param_type> params(1);
        params[0].M_d = 1.8341, params[0].M_rinv = 1.0 / 0.0302, params[0].r_eq = 1.0043;
        params[0].omega = 5.0, params[0].alpha = 0.40, params[0].repulsion = false;
        Scalar4 *d_force = nullptr, *d_torque = nullptr;
        Scalar* d_virial = nullptr;
        cudaMalloc(reinterpret_cast<void**>(&d_force), sizeof(Scalar4) * 2);
        cudaMalloc(reinterpret_cast<void**>(&d_torque), sizeof(Scalar4) * 2);
        cudaMalloc(reinterpret_cast<void**>(&d_virial), sizeof(Scalar) * 12);
        md::kernel::a_pair_args_t args {d_force, d_torque, d_virial, 2, 2, 0, upload(pos), nullptr,
                                        upload(quat), upload(tag), BoxDim(7.98, 7.98, 7.98),
                                        upload(n_neigh), upload(nlist), upload(head), upload(rcutsq),
                                        1, 128, 0, 1, 1, prop};
        hipError_t rc = md::kernel::gpu_compute_pair_aniso_forces<azplugins::detail::AnisoPairEvaluatorTwoPatchMorse>(
            args, upload(params), nullptr);
        cudaDeviceSynchronize();
        std::vector<Scalar4> f(2), t(2);
        cudaMemcpy(f.data(), d_force, sizeof(Scalar4) * 2, cudaMemcpyDeviceToHost);
        cudaMemcpy(t.data(), d_torque, sizeof(Scalar4) * 2, cudaMemcpyDeviceToHost);
        std::printf("morse rc %d f0 %.6f %.6f %.6f e0 %.6f f1 %.6f e1 %.6f t0 %.6f %.6f %.6f t1 %.6f %.6f %.6f\n",
                    rc, f[0].x, f[0].y, f[0].z, f[0].w, f[1].x, f[1].w, t[0].x, t[0].y, t[0].z, t[1].x,
                    t[1].y, t[1].z);
        worst |= rc;
        }
    return worst;
    }
