#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float ffma3(float a,float b,float c){float d; asm volatile("fma.rn.f32 %0,%1,%2,%3;":"=f"(d):"f"(a),"f"(b),"f"(c)); return d;}
__device__ __forceinline__ float fmul2r(float a,float b){float d; asm volatile("mul.rn.f32 %0,%1,%2;":"=f"(d):"f"(a),"f"(b)); return d;}
__device__ __forceinline__ float fadd2r(float a,float b){float d; asm volatile("add.rn.f32 %0,%1,%2;":"=f"(d):"f"(a),"f"(b)); return d;}
__device__ __forceinline__ unsigned lop(unsigned a,unsigned b,unsigned c){unsigned d; asm volatile("lop3.b32 %0,%1,%2,%3,0x96;":"=r"(d):"r"(a),"r"(b),"r"(c)); return d;}
__device__ __forceinline__ float fsel(float a,float b,float p){float d; asm volatile("{.reg .pred q; setp.gt.f32 q,%3,0f00000000; selp.f32 %0,%1,%2,q;}":"=f"(d):"f"(a),"f"(b),"f"(p)); return d;}
__device__ __forceinline__ float mufu(float a){float d; asm volatile("ex2.approx.ftz.f32 %0,%1;":"=f"(d):"f"(a)); return d;}
// MODE bits: which op per slot
template<int NF3,int NFI,int NM2,int NA2,int NL,int NS,int NX> __global__ void k(float* out,int iters,float s,float s2){
  float f[16], g[16], h[16], m[8]; unsigned l[16];
  for(int i=0;i<16;i++){f[i]=threadIdx.x*0.001f+i; g[i]=s+i*1e-3f; h[i]=s2+i*1e-4f; l[i]=threadIdx.x*7+i;}
  for(int i=0;i<8;i++) m[i]=threadIdx.x*1e-3f;
  for(int it=0;it<iters;++it){
    #pragma unroll
    for(int i=0;i<16;i++){
      if(i<NF3) f[i]=ffma3(f[i],g[i],h[i]);          // 3 distinct registers
      if(i<NFI) f[i]=ffma3(f[i],s,0.5f);             // reg, reg(uniform), imm
      if(i<NM2) g[i]=fmul2r(g[i],h[i]);
      if(i<NA2) h[i]=fadd2r(h[i],g[(i+1)&15]);
      if(i<NL) l[i]=lop(l[i],l[(i+1)&15],0x5a5a5a5a);
      if(i<NS) g[i]=fsel(g[i],h[i],f[i]);            // FSETP + FSEL
      if(i<NX) m[i&7]=mufu(m[i&7]);
    }
  }
  float r=0; for(int i=0;i<16;i++) r+=f[i]+g[i]+h[i]+__uint_as_float(l[i]); for(int i=0;i<8;i++) r+=m[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=r;
}
template<int NF3,int NFI,int NM2,int NA2,int NL,int NS,int NX> void run(const char* name,float* out){
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int iters=20000; float best=1e9;
  for(int rep=0;rep<3;rep++){
    cudaEventRecord(e0);
    k<NF3,NFI,NM2,NA2,NL,NS,NX><<<148*4,256>>>(out,iters,0.999f,0.5f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms,e0,e1); if(ms<best) best=ms;
  }
  double cyc = best*1e-3*1.965e9/iters/8.0;
  printf("%-44s %.3f ms -> %.2f cycles/iter/warp\n",name,best,cyc);
}
int main(){
  float* out; cudaMalloc(&out,148*4*256*4);
  run<16,0,0,0,0,0,0>("16 FFMA 3-reg",out);
  run<0,16,0,0,0,0,0>("16 FFMA reg,reg,imm",out);
  run<0,0,16,0,0,0,0>("16 FMUL 2-reg",out);
  run<0,0,0,16,0,0,0>("16 FADD 2-reg",out);
  run<8,0,8,0,0,0,0>("8 FFMA3 + 8 FMUL",out);
  run<16,0,0,0,16,0,0>("16 FFMA3 + 16 LOP3",out);
  run<16,0,0,0,8,0,0>("16 FFMA3 + 8 LOP3",out);
  run<0,0,0,0,0,16,0>("16 (FSETP+FSEL)",out);
  run<16,0,0,0,0,8,0>("16 FFMA3 + 8 (FSETP+FSEL)",out);
  run<8,0,8,8,0,8,0>("8 FFMA3+8 FMUL+8 FADD+8(FSETP+FSEL)",out);
  run<8,0,8,8,0,8,3>("8 FFMA3+8 FMUL+8 FADD+8(FSETP+FSEL)+3 MUFU",out);
  run<14,0,10,7,0,8,3>("kernel-like mix 14 F3+10 M+7 A+8 SEL+3 MUFU",out);
  return 0;
}
