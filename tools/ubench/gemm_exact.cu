// a-block of the structured chain at its exact size (D = 120) against the padded layout in use (D = 128, kvalid = 120)
#include "../opentn_b200/csrc/gemm_dmma.cuh"
#include <vector>
#include <cstdlib>
using namespace otn;
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1);}}while(0)
template<int BM,int BN,int BK,int WM,int WN,int ST,bool TA,bool TB>
void run(GemmParams p, const char* tag){
  int tiles=(p.rows/BM)*(p.D/BN);
  cudaError_t e = launch_dmma<BM,BN,BK,WM,WN,ST,TA,TB>(p,tiles,0);
  if(e!=cudaSuccess){printf("%s: launch failed %s\n",tag,cudaGetErrorString(e)); cudaGetLastError(); return;}
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int reps=20;
  cudaEventRecord(e0); for(int r=0;r<reps;r++) launch_dmma<BM,BN,BK,WM,WN,ST,TA,TB>(p,tiles,0); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms,e0,e1); ms/=reps;
  printf("D=%d %s tile %dx%dx%d warp %dx%d st%d ta%d tb%d dual%d: %.3f ms (useful 2*120^3: %.2f TFLOP/s)\n",p.D,tag,BM,BN,BK,WM,WN,ST,TA,TB,p.A2?1:0,ms,2.0*p.batch*120.0*120*120*(p.A2?2:1)/ms/1e9);
}
int main(){
  int B=1024;
  for(int D : {128,120}){
    size_t n=(size_t)D*D;
    std::vector<double> h(B*n); srand(1); for(auto&x:h) x=(rand()/(double)RAND_MAX)-0.5;
    double *dA,*dB,*dC,*dA2,*dB2; for(auto pp:{&dA,&dB,&dC,&dA2,&dB2}) CK(cudaMalloc(pp,B*n*8+4096));
    for(auto pp:{dA,dB,dA2,dB2}) CK(cudaMemcpy(pp,h.data(),B*n*8,cudaMemcpyHostToDevice));
    GemmParams p{}; p.A1=dA;p.B1=dB;p.C=dC;p.sA1=p.sB1=p.sC=p.sA2=p.sB2=n;p.D=D;p.rows=D;p.batch=B;p.epi=EPI_NONE;
    for(int dual=0; dual<2; dual++){
      p.A2=dual?dA2:nullptr; p.B2=dual?dB2:nullptr;
      if(D==128){
        p.kvalid=120;
        run<64,64,16,32,32,3,false,false>(p,"padded");
        run<64,64,16,32,32,3,false,true>(p,"padded");
        run<64,64,16,32,32,3,true,false>(p,"padded");
      } else {
        p.kvalid=0;
        run<120,120,24,40,40,3,false,false>(p,"exact");
        run<120,120,8,40,40,4,false,false>(p,"exact");
        run<120,120,24,40,40,2,false,false>(p,"exact");
        run<120,40,24,40,40,3,false,false>(p,"exact");
        run<40,120,24,40,40,3,false,false>(p,"exact");
        run<120,40,8,40,40,4,false,false>(p,"exact");
        run<40,40,24,40,40,3,false,false>(p,"exact");
        run<120,120,24,40,40,3,false,true>(p,"exact");
        run<120,120,24,40,40,3,true,false>(p,"exact");
        run<120,120,24,24,40,3,false,false>(p,"exact15w");
        run<120,120,24,40,24,3,false,false>(p,"exact15w");
      }
    }
  }
  return 0;
}
