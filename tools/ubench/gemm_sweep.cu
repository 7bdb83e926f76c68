// tile sweep for the two block sizes of the structured chain (D = 144, 128), batch 1024, random data
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
  double fl=2.0*p.batch*(double)p.D*p.D*p.D*(p.A2?2:1);
  printf("D=%d %s tile %dx%dx%d warp %dx%d st%d ta%d tb%d dual%d: %.3f ms %.2f TFLOP/s\n",p.D,tag,BM,BN,BK,WM,WN,ST,TA,TB,p.A2?1:0,ms,fl/ms/1e9);
}
int main(){
  int B=1024;
  for(int D : {144,128}){
    size_t n=(size_t)D*D;
    std::vector<double> h(B*n); srand(1); for(auto&x:h) x=(rand()/(double)RAND_MAX)-0.5;
    double *dA,*dB,*dC,*dA2,*dB2; for(auto pp:{&dA,&dB,&dC,&dA2,&dB2}) CK(cudaMalloc(pp,B*n*8));
    for(auto pp:{dA,dB,dA2,dB2}) CK(cudaMemcpy(pp,h.data(),B*n*8,cudaMemcpyHostToDevice));
    GemmParams p{}; p.A1=dA;p.B1=dB;p.C=dC;p.sA1=p.sB1=p.sC=p.sA2=p.sB2=n;p.D=D;p.rows=D;p.batch=B;p.epi=EPI_NONE;
    for(int dual=0; dual<2; dual++){
      p.A2=dual?dA2:nullptr; p.B2=dual?dB2:nullptr;
      if(D==144){
        run<48,48,16,24,24,3,false,false>(p,"base");
        run<48,48,24,24,24,3,false,false>(p,"bk24");
        run<48,48,48,24,24,2,false,false>(p,"bk48s2");
        run<48,48,48,24,24,3,false,false>(p,"bk48s3");
        run<48,48,24,24,24,4,false,false>(p,"bk24s4");
        run<48,48,8,24,24,4,false,false>(p,"bk8 ");
        run<48,48,8,24,24,6,false,false>(p,"bk8s6");
        run<72,72,16,24,24,3,false,false>(p,"72x72");
        run<72,48,16,24,24,3,false,false>(p,"72x48");
        run<144,48,16,48,24,3,false,false>(p,"144x48");
        run<48,48,16,24,24,3,false,true>(p,"base");
        run<48,48,16,24,24,3,true,false>(p,"base");
        run<72,72,16,24,24,3,false,true>(p,"72x72");
        run<72,72,16,24,24,3,true,false>(p,"72x72");
      } else {
        run<64,64,16,32,32,3,false,false>(p,"base");
        run<64,64,32,32,32,2,false,false>(p,"bk32s2");
        run<64,64,32,32,32,3,false,false>(p,"bk32s3");
        run<64,64,64,32,32,2,false,false>(p,"bk64s2");
        run<64,64,8,32,32,4,false,false>(p,"bk8 ");
        run<128,64,16,32,32,3,false,false>(p,"128x64");
        run<64,128,16,32,32,3,false,false>(p,"64x128");
        run<64,32,16,32,16,3,false,false>(p,"64x32");
        run<32,32,16,16,16,3,false,false>(p,"32x32");
      }
    }
  }
  return 0;
}
