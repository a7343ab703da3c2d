// Microbenchmark for the "evaluate j > i only, scatter -F to j with red.global.add.v4.f32"
// question: how many 16-byte vector reductions per
// second does the B200 L2 sustain with the access pattern of a spatially sorted neighbour list?
// C2 has 61 M pairs per step (half of 122 M list entries); halving the evaluations would save
// ~0.11 ms of the 0.285 ms kernel, so the scatter has to sustain >> 61 M / 0.11 ms = 550 G
// reductions/s to pay (more with the virial: 6 more scalar reductions per pair).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 redg_v4.cu -o redg_v4 && ./redg_v4
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned int hash(unsigned int x)
    {
    x ^= x >> 16;
    x *= 0x7feb352du;
    x ^= x >> 15;
    x *= 0x846ca68bu;
    x ^= x >> 16;
    return x;
    }

// MODE 0: red.global.add.v4.f32 per "pair"; 1: four scalar red.global.add.f32; 2: no scatter
// (the compute-only floor of this loop); window = index distance of a neighbour (locality)
template<int MODE> __global__ void scatter(float4* force, unsigned int N, unsigned int K, unsigned int window)
    {
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N)
        return;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (unsigned int k = 0; k < K; ++k)
        {
        const unsigned int h = hash(i * 131u + k);
        unsigned int j = i + 1u + (h % window);
        if (j >= N)
            j -= N;
        const float f = __uint_as_float(0x3f800000u | (h & 0xffffu)) - 1.5f;
        const float4 v = make_float4(f, -f, 0.5f * f, 1.0f);
        acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
        if (MODE == 0)
            atomicAdd(force + j, make_float4(-v.x, -v.y, -v.z, v.w));
        else if (MODE == 1)
            {
            float* p = reinterpret_cast<float*>(force + j);
            atomicAdd(p, -v.x);
            atomicAdd(p + 1, -v.y);
            atomicAdd(p + 2, -v.z);
            atomicAdd(p + 3, v.w);
            }
        }
    if (MODE == 2)
        force[i] = acc;
    else
        atomicAdd(force + i, acc);
    }

template<int MODE> float run(float4* d, unsigned int N, unsigned int K, unsigned int window)
    {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep)
        {
        cudaMemsetAsync(d, 0, sizeof(float4) * N);
        cudaEventRecord(e0);
        scatter<MODE><<<(N + 127) / 128, 128>>>(d, N, K, window);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        best = ms < best ? ms : best;
        }
    return best;
    }

int main()
    {
    const unsigned int N = 1000000, K = 61;
    float4* d;
    cudaMalloc(&d, sizeof(float4) * N);
    const unsigned int windows[] = {256, 4096, 65536, 1000000};
    std::printf("N = %u rows, K = %u scattered pairs per row (C2: 61 M pairs per step)\n", N, K);
    for (unsigned int w : windows)
        {
        const float t0 = run<0>(d, N, K, w), t1 = run<1>(d, N, K, w), t2 = run<2>(d, N, K, w);
        std::printf("window %7u: red.v4.f32 %.3f ms (%.0f G red/s) | 4 x red.f32 %.3f ms | no scatter %.3f ms\n", w,
                    t0, 1e-6 * N * K / t0, t1, t2);
        }
    return 0;
    }
