// L1 / shared-memory 16-byte gather throughput under controlled lane patterns.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>
// idx table: per (iteration slot, thread) 16B-element index within a window of W elements
template<int MODE> __global__ void __launch_bounds__(256) k(const float4* __restrict__ data, const unsigned* __restrict__ idx, int nslots, int iters, int window, float* out)
{
  extern __shared__ float4 sm[];
  const unsigned t = threadIdx.x;
  const float4* base = data + (size_t)blockIdx.x * window;  // each CTA its own window (L1-resident after first touch)
  if (MODE == 1) { for (int i = t; i < window; i += blockDim.x) sm[i] = base[i]; __syncthreads(); }
  float4 acc = make_float4(0,0,0,0);
  for (int it = 0; it < iters; ++it) {
    #pragma unroll 4
    for (int s = 0; s < nslots; ++s) {
      const unsigned j = __ldg(idx + s * 256 + t);
      float4 v;
      if (MODE == 0) v = __ldg(base + j);
      else if (MODE == 1) v = sm[j];
      else { const float2* p = reinterpret_cast<const float2*>(base + j); float2 a = __ldg(p); v = make_float4(a.x, a.y, 0, 0); }
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  out[blockIdx.x * blockDim.x + t] = acc.x + acc.y + acc.z + acc.w;
}
int main() {
  const int window = 2048;            // 32 KB of float4 per CTA
  const int nslots = 64, iters = 200;
  const int grid = 148 * 2;           // 2 CTAs of 256 threads per SM -> 64 KB smem/L1 footprint per SM
  float4* data; cudaMalloc(&data, (size_t)grid * window * 16); cudaMemset(data, 0, (size_t)grid * window * 16);
  unsigned* idx; cudaMalloc(&idx, nslots * 256 * 4);
  float* out; cudaMalloc(&out, grid * 256 * 4);
  cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, window * 16);
  const char* names[] = {"coalesced", "random16B", "pairs(32B)", "quads(64B)", "octets(128B)", "random-line,same-offset", "2 lines/quarter", "4 lines/quarter, spread", "stride 96B (gap 6)"};
  for (int pat = 0; pat < 9; ++pat) {
    std::vector<unsigned> h(nslots * 256);
    srand(1234);
    for (int s = 0; s < nslots; ++s) for (int w = 0; w < 8; ++w) {
      unsigned* q = &h[s * 256 + w * 32];
      for (int l = 0; l < 32; ++l) {
        unsigned v = 0;
        switch (pat) {
          case 0: v = ((rand() % (window / 32)) * 32); v = (l == 0) ? v : q[0] + l; break;
          case 1: v = rand() % window; break;
          case 2: v = (l % 2 == 0) ? (rand() % (window / 2)) * 2 : q[l - 1] + 1; break;
          case 3: v = (l % 4 == 0) ? (rand() % (window / 4)) * 4 : q[l - 1] + 1; break;
          case 4: v = (l % 8 == 0) ? (rand() % (window / 8)) * 8 : q[l - 1] + 1; break;
          case 5: v = (rand() % (window / 8)) * 8 + (l % 8); break;
          case 6: { if (l % 8 == 0) { q[l] = (rand() % (window / 8)) * 8; v = q[l]; } else if (l % 8 == 4) { v = (rand() % (window / 8)) * 8; } else v = q[l - 1] + 1; } break;
          case 7: { if (l % 2 == 0) v = (rand() % (window / 8)) * 8 + (rand() % 4) * 2; else v = q[l - 1] + 1; } break;
          case 8: v = (l % 8 == 0) ? (rand() % (window - 64)) : q[l - 1] + 6; break;
        }
        q[l] = v % window;
      }
    }
    cudaMemcpy(idx, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    for (int mode = 0; mode < 3; ++mode) {
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      float best = 1e9;
      for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        if (mode == 0) k<0><<<grid, 256, 0>>>(data, idx, nslots, iters, window, out);
        else if (mode == 1) k<1><<<grid, 256, window * 16>>>(data, idx, nslots, iters, window, out);
        else k<2><<<grid, 256, 0>>>(data, idx, nslots, iters, window, out);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
      }
      // warp-requests per SM = 2 CTAs * 8 warps * nslots * iters
      double req = 2.0 * 8 * nslots * iters;
      printf("%-26s %-8s %.3f ms  %.2f cycles/request/SM\n", names[pat], mode == 0 ? "LDG.128" : mode == 1 ? "LDS.128" : "LDG.64", best, best * 1e-3 * 1.965e9 / req);
    }
  }
  cudaError_t e = cudaDeviceSynchronize(); printf("%s\n", cudaGetErrorString(e));
  return 0;
}
