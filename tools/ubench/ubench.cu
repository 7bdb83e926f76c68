// microbench: DMMA.8x8x4 vs DFMA peak throughput on B200
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("err %s line %d\n", cudaGetErrorString(e), __LINE__); return 1;}}while(0)
template<int NACC>
__global__ void __launch_bounds__(256) k_dmma(double* out, int iters, double s){
  double c[NACC][2]; double a = s*threadIdx.x, b = s*(threadIdx.x+1);
  #pragma unroll
  for(int i=0;i<NACC;i++){c[i][0]=0;c[i][1]=0;}
  for(int it=0; it<iters; ++it){
    #pragma unroll
    for(int i=0;i<NACC;i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double r=0;
  #pragma unroll
  for(int i=0;i<NACC;i++) r+=c[i][0]+c[i][1];
  out[blockIdx.x*blockDim.x+threadIdx.x]=r;
}
template<int NACC>
__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double s){
  double c[NACC]; double a = s*threadIdx.x, b = s*(threadIdx.x+1);
  #pragma unroll
  for(int i=0;i<NACC;i++){c[i]=i;}
  for(int it=0; it<iters; ++it){
    #pragma unroll
    for(int i=0;i<NACC;i++) c[i] = fma(c[i], a, b);
  }
  double r=0;
  #pragma unroll
  for(int i=0;i<NACC;i++) r+=c[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=r;
}
int main(){
  double* out; CK(cudaMalloc(&out, 148*32*1024*8));
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int iters=20000;
  for(int warps=4; warps<=32; warps*=2){
    for(int ctas=1; ctas<=2; ctas++){
      int threads = warps*32/ctas; if(threads>1024||threads<32) continue;
      int grid=148*ctas;
      float ms;
      k_dmma<16><<<grid,threads>>>(out, 100, 1e-9); CK(cudaDeviceSynchronize());
      cudaEventRecord(e0); k_dmma<16><<<grid,threads>>>(out, iters, 1e-9); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
      cudaEventElapsedTime(&ms,e0,e1);
      double fl = (double)grid*(threads/32)*iters*16*512.0;
      printf("DMMA warps/SM=%d ctas/SM=%d: %.2f TFLOP/s (%.3f ms)\n", warps, ctas, fl/ms/1e9, ms);
      k_dfma<16><<<grid,threads>>>(out, 100, 1e-9); CK(cudaDeviceSynchronize());
      cudaEventRecord(e0); k_dfma<16><<<grid,threads>>>(out, iters, 1e-9); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
      cudaEventElapsedTime(&ms,e0,e1);
      fl = (double)grid*threads*iters*16*2.0;
      printf("DFMA warps/SM=%d ctas/SM=%d: %.2f TFLOP/s (%.3f ms)\n", warps, ctas, fl/ms/1e9, ms);
    }
  }
  // sustained 2 s DMMA
  {
    int threads=512, grid=148; float ms;
    cudaEventRecord(e0); for(int r=0;r<20;r++) k_dmma<16><<<grid,threads>>>(out, 200000, 1e-9); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    cudaEventElapsedTime(&ms,e0,e1);
    double fl = 20.0*grid*(threads/32)*200000.0*16*512.0;
    printf("DMMA sustained: %.2f TFLOP/s (%.1f ms)\n", fl/ms/1e9, ms);
  }
  return 0;
}
