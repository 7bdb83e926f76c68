// standalone correctness + throughput harness for gemm_dmma.cuh
#include "../opentn_b200/csrc/gemm_dmma.cuh"
#include <vector>
#include <cmath>
#include <cstdlib>
using namespace otn;
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1);}}while(0)

static void cpu_gemm(const double* A,const double* B,double* C,int D,bool ta,bool tb,bool accum){
  for(int i=0;i<D;i++)for(int j=0;j<D;j++){ double s=accum?C[i*D+j]:0; for(int k=0;k<D;k++){ double a=ta?A[k*D+i]:A[i*D+k]; double b=tb?B[j*D+k]:B[k*D+j]; s+=a*b;} C[i*D+j]=s; }
}
template<int BM,int BN,int BK,int WM,int WN,int ST,bool TA,bool TB>
float time_cfg(GemmParams p, int reps){
  int tiles=(p.rows/BM)*(p.D/BN);
  CK((launch_dmma<BM,BN,BK,WM,WN,ST,TA,TB>(p,tiles,0))); CK(cudaDeviceSynchronize());
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0); for(int r=0;r<reps;r++) launch_dmma<BM,BN,BK,WM,WN,ST,TA,TB>(p,tiles,0); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms,e0,e1); return ms/reps;
}
int main(int argc,char**argv){
  // ---------- correctness, D in {256, 64, 48, 16}, batch 3 ----------
  for(int D : {256, 64, 48, 16}){
    int B=3; size_t n=(size_t)D*D;
    std::vector<double> hA1(B*n),hB1(B*n),hA2(B*n),hB2(B*n),hE(n),hC(B*n),ref(B*n);
    srand(1); auto rnd=[&](){return (rand()/(double)RAND_MAX)-0.5;};
    for(auto*v:{&hA1,&hB1,&hA2,&hB2,&hE}) for(auto&x:*v)x=rnd();
    double *dA1,*dB1,*dA2,*dB2,*dE,*dC,*dP;
    for(auto pp:{&dA1,&dB1,&dA2,&dB2,&dC}) CK(cudaMalloc(pp,B*n*8));
    CK(cudaMalloc(&dE,n*8)); CK(cudaMalloc(&dP,B*4096*8));
    CK(cudaMemcpy(dA1,hA1.data(),B*n*8,cudaMemcpyHostToDevice));CK(cudaMemcpy(dB1,hB1.data(),B*n*8,cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dA2,hA2.data(),B*n*8,cudaMemcpyHostToDevice));CK(cudaMemcpy(dB2,hB2.data(),B*n*8,cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dE,hE.data(),n*8,cudaMemcpyHostToDevice));
    for(int mode=0; mode<6; mode++){
      bool ta = (mode%3)==2, tb=(mode%3)==1; bool dual = mode>=3;
      GemmParams p{}; p.A1=dA1;p.B1=dB1;p.A2=dual?dA2:nullptr;p.B2=dual?dB2:nullptr;p.C=dC;p.sA1=p.sB1=p.sA2=p.sB2=p.sC=n;p.D=D;p.row0=0;p.rows=D;p.batch=B;p.active=nullptr;
      p.epi = dual?EPI_RESID:EPI_NONE; p.E=dE; p.sE=0; p.e_index=nullptr; p.partial=dP;
      CK(gemm_launch(p,ta,tb,0)); CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(hC.data(),dC,B*n*8,cudaMemcpyDeviceToHost));
      GemmPlan pl=gemm_plan(D,D,B);
      std::vector<double> hp(B*pl.tiles); CK(cudaMemcpy(hp.data(),dP,hp.size()*8,cudaMemcpyDeviceToHost));
      double maxerr=0, serr=0;
      for(int b=0;b<B;b++){
        cpu_gemm(&hA1[b*n],&hB1[b*n],&ref[b*n],D,ta,tb,false);
        if(dual) cpu_gemm(&hA2[b*n],&hB2[b*n],&ref[b*n],D,ta,tb,true);
        double ss=0; 
        for(size_t i=0;i<n;i++){ double r=ref[b*n+i]; if(dual){ r-=hE[i]; ss+=r*r;} maxerr=fmax(maxerr,fabs(r-hC[b*n+i])); }
        if(dual){ double gs=0; for(int t=0;t<pl.tiles;t++) gs+=hp[b*pl.tiles+t]; serr=fmax(serr,fabs(gs-ss)/ss);}
      }
      printf("D=%d kind=%d mode ta=%d tb=%d dual=%d: max abs err %.3e, partial-sum rel err %.3e\n",D,pl.kind,ta,tb,dual,maxerr,serr);
    }
    for(auto pp:{dA1,dB1,dA2,dB2,dC,dE,dP}) cudaFree(pp);
  }
  // ---------- throughput, D=256, B=1024 ----------
  {
    int D=256,B=1024; size_t n=(size_t)D*D;
    double *dA,*dB,*dC,*dP; CK(cudaMalloc(&dA,B*n*8));CK(cudaMalloc(&dB,B*n*8));CK(cudaMalloc(&dC,B*n*8));CK(cudaMalloc(&dP,B*64*8));
    CK(cudaMemset(dA,0,B*n*8));CK(cudaMemset(dB,0,B*n*8));
    GemmParams p{}; p.A1=dA;p.B1=dB;p.C=dC;p.sA1=p.sB1=p.sC=n;p.D=D;p.rows=D;p.batch=B;p.partial=dP;p.epi=EPI_NONE;
    double fl=2.0*B*D*D*D;
    #define T(BM,BN,BK,WM,WN,ST,TA,TB) { float ms=time_cfg<BM,BN,BK,WM,WN,ST,TA,TB>(p,5); printf("cfg %dx%dx%d warp %dx%d st%d ta%d tb%d: %.3f ms  %.2f TFLOP/s\n",BM,BN,BK,WM,WN,ST,TA,TB,ms,fl/ms/1e9); }
    T(128,128,16,64,32,3,false,false)
    T(128,128,16,32,64,3,false,false)
    T(128,128,16,32,32,3,false,false)
    T(128,128,16,64,32,4,false,false)
    T(128,128,8,64,32,4,false,false)
    T(128,128,32,64,32,2,false,false)
    T(128,64,16,32,32,3,false,false)
    T(128,64,16,64,32,3,false,false)
    T(64,128,16,32,32,3,false,false)
    T(64,64,16,32,32,3,false,false)
    T(64,64,16,32,16,4,false,false)
    T(128,128,16,64,32,3,false,true)
    T(128,128,16,64,32,3,true,false)
    T(128,64,16,32,32,3,false,true)
    T(128,64,16,32,32,3,true,false)
    // dual product
    p.A2=dA;p.B2=dB;p.sA2=p.sB2=n; fl*=2;
    T(128,128,16,64,32,3,false,false)
    T(128,128,16,64,32,3,false,true)
    T(128,128,16,64,32,3,true,false)
  }
  return 0;
}
