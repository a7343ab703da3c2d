#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 ffma2(u64 a,u64 b,u64 c){u64 d; asm volatile("fma.rn.f32x2 %0,%1,%2,%3;":"=l"(d):"l"(a),"l"(b),"l"(c)); return d;}
__device__ __forceinline__ float ffma(float a,float b,float c){float d; asm volatile("fma.rn.f32 %0,%1,%2,%3;":"=f"(d):"f"(a),"f"(b),"f"(c)); return d;}
__device__ __forceinline__ unsigned lop(unsigned a,unsigned b,unsigned c){unsigned d; asm volatile("lop3.b32 %0,%1,%2,%3,0x96;":"=r"(d):"r"(a),"r"(b),"r"(c)); return d;}
__device__ __forceinline__ float mufu(float a){float d; asm volatile("ex2.approx.ftz.f32 %0,%1;":"=f"(d):"f"(a)); return d;}
__device__ __forceinline__ float fsel(float a,float b,unsigned p){float d; asm volatile("{.reg .pred q; setp.ne.u32 q,%3,0; selp.f32 %0,%1,%2,q;}":"=f"(d):"f"(a),"f"(b),"r"(p)); return d;}
template<int NF,int NF2,int NL,int NM> __global__ void k(float* out,int iters,float s){
  float f[16]; u64 g[16]; unsigned l[16]; float m[8];
  for(int i=0;i<16;i++){f[i]=threadIdx.x*0.001f+i; g[i]=((u64)__float_as_uint(f[i])<<32)|__float_as_uint(f[i]*0.5f); l[i]=threadIdx.x*7+i;}
  for(int i=0;i<8;i++) m[i]=threadIdx.x*1e-3f;
  u64 s2=((u64)__float_as_uint(s)<<32)|__float_as_uint(s); u64 c2=((u64)__float_as_uint(0.5f)<<32)|__float_as_uint(0.25f);
  for(int it=0;it<iters;++it){
    #pragma unroll
    for(int i=0;i<16;i++){
      if(i<NF) f[i]=ffma(f[i],s,0.5f);
      if(i<NF2) g[i]=ffma2(g[i],s2,c2);
      if(i<NL) l[i]=lop(l[i],l[(i+1)&15],0x5a5a5a5a);
      if(i<NM) m[i&7]=mufu(m[i&7]);
    }
  }
  float r=0; for(int i=0;i<16;i++) r+=f[i]+__uint_as_float((unsigned)g[i])+__uint_as_float((unsigned)(g[i]>>32))+__uint_as_float(l[i]); for(int i=0;i<8;i++) r+=m[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=r;
}
template<int NF,int NF2,int NL,int NM> void run(const char* name,float* out){
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int iters=20000;
  float best=1e9;
  for(int rep=0;rep<3;rep++){
    cudaEventRecord(e0);
    k<NF,NF2,NL,NM><<<148*4,256>>>(out,iters,0.999f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms,e0,e1); if(ms<best) best=ms;
  }
  // 8 warps per SMSP; cycles per iteration per warp-slot
  double cyc = best*1e-3*1.965e9/iters/8.0;
  printf("%-28s %.3f ms  -> %.2f cycles/iter/warp (NF=%d NF2=%d NL=%d NM=%d; instr=%d)\n",name,best,cyc,NF,NF2,NL,NM,NF+NF2+NL+NM);
}
int main(){
  float* out; cudaMalloc(&out,148*4*256*4);
  run<16,0,0,0>("16 FFMA",out);
  run<0,8,0,0>("8 FFMA2",out);
  run<0,16,0,0>("16 FFMA2",out);
  run<0,0,16,0>("16 LOP3",out);
  run<16,0,16,0>("16 FFMA + 16 LOP3",out);
  run<0,8,16,0>("8 FFMA2 + 16 LOP3",out);
  run<0,16,16,0>("16 FFMA2 + 16 LOP3",out);
  run<8,8,0,0>("8 FFMA + 8 FFMA2",out);
  run<16,0,0,4>("16 FFMA + 4 MUFU",out);
  run<0,8,0,4>("8 FFMA2 + 4 MUFU",out);
  run<0,0,0,4>("4 MUFU",out);
  run<0,0,0,8>("8 MUFU",out);
  run<0,8,8,4>("8 FFMA2 + 8 LOP3 + 4 MUFU",out);
  run<16,0,8,4>("16 FFMA + 8 LOP3 + 4 MUFU",out);
  return 0;
}
