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
    // reference KAT (src/pytest/test_pair.py:177-186): Hertz epsilon=2, r_cut=1.5, d=1.05
    //   -> U = 0.0985, F = 0.5477
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
    hipDeviceProp_t prop {10, 0};
    md::kernel::pair_args_t args {d_force, d_virial, 2, 2, 0, upload(pos), nullptr,
                                  BoxDim(7.98, 7.98, 7.98), upload(n_neigh), upload(nlist),
                                  upload(head), upload(rcutsq), upload(ronsq), nlist.size(), 1, 128,
                                  0, 1, 1, prop};
    hipError_t rc = md::kernel::gpu_compute_pair_forces<azplugins::detail::PairEvaluatorHertz>(
        args, upload(params));
    cudaDeviceSynchronize();
    std::vector<Scalar4> f(2);
    cudaMemcpy(f.data(), d_force, sizeof(Scalar4) * 2, cudaMemcpyDeviceToHost);
    std::printf("rc %d f0 %.6f %.6f %.6f e0 %.6f f1 %.6f e1 %.6f\n", rc, f[0].x, f[0].y, f[0].z,
                f[0].w, f[1].x, f[1].w);
    return rc;
    }
