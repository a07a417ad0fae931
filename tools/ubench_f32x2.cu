// Micro-benchmark: scalar FADD/FFMA vs packed FADD2/FFMA2 issue throughput on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_f32x2 ubench_f32x2.cu ; run on a B200.
#include <cuda_runtime.h>
#include <cstdio>
typedef unsigned long long u64;
__device__ __forceinline__ u64 mk(float a, float b){ u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void un(u64 a, float& x, float& y){ asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a)); }
__device__ __forceinline__ u64 add2(u64 a, u64 b){ u64 r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c){ u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float adds(float a, float b){ float r; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float fmas(float a, float b, float c){ float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
constexpr int ITERS = 4096, ILP = 8;
template <int MODE> __global__ void k(float* out, float seed){
  float a[2*ILP]; u64 p[ILP];
  for (int i=0;i<2*ILP;++i) a[i]=seed+i+threadIdx.x;
  for (int i=0;i<ILP;++i) p[i]=mk(a[2*i],a[2*i+1]);
  const u64 c2=mk(seed,seed*0.5f);
  for (int it=0; it<ITERS; ++it){
    if (MODE==0) { for (int i=0;i<2*ILP;++i) a[i]=adds(a[i],seed); }            // 16 scalar FADD = 16 lane-ops
    if (MODE==1) { for (int i=0;i<ILP;++i) p[i]=add2(p[i],c2); }                 // 8 FADD2 = 16 lane-ops
    if (MODE==2) { for (int i=0;i<2*ILP;++i) a[i]=fmas(a[i],seed,seed); }
    if (MODE==3) { for (int i=0;i<ILP;++i) p[i]=fma2(p[i],c2,c2); }
  }
  float s=0; for (int i=0;i<2*ILP;++i) s+=a[i]; for (int i=0;i<ILP;++i){ float x,y; un(p[i],x,y); s+=x+y; }
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template <int MODE> float run(float* d){
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148*8,256>>>(d,1.0f); cudaDeviceSynchronize();
  cudaEventRecord(e0); k<MODE><<<148*8,256>>>(d,1.0f); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms,e0,e1); return ms;
}
int main(){ float* d; cudaMalloc(&d, 148*8*256*4);
  const double laneops = 148.0*8*256*ITERS*16;
  float t0=run<0>(d), t1=run<1>(d), t2=run<2>(d), t3=run<3>(d);
  printf("FADD  %.3f ms  %.1f Tlane-op/s\nFADD2 %.3f ms  %.1f Tlane-op/s\nFFMA  %.3f ms  %.1f\nFFMA2 %.3f ms  %.1f\n", t0, laneops/t0/1e9, t1, laneops/t1/1e9, t2, laneops/t2/1e9, t3, laneops/t3/1e9);
  return 0; }
