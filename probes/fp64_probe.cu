// FP64 peak probe for B200 (sm_100a): DFMA issue rate, DMMA.8x8x4 rate, cuBLAS DGEMM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_probe fp64_probe.cu -lcublas
// Output: one JSON line (written by the caller into profiles/).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
#include <cublas_v2.h>

#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

template<int ILP>
__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double a, double b){
  double acc[ILP];
  #pragma unroll
  for(int i=0;i<ILP;i++) acc[i]=threadIdx.x*1e-3+i;
  for(int it=0;it<iters;it++){
    #pragma unroll
    for(int i=0;i<ILP;i++) acc[i]=fma(acc[i],a,b);
  }
  double s=0;
  #pragma unroll
  for(int i=0;i<ILP;i++) s+=acc[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}

template<int NACC>
__global__ void __launch_bounds__(256) dmma_kernel(double* out, int iters){
  double c[NACC][2];
  #pragma unroll
  for(int i=0;i<NACC;i++){c[i][0]=0;c[i][1]=0;}
  double a=threadIdx.x*1e-6, b=1e-6*(threadIdx.x&7);
  for(int it=0;it<iters;it++){
    #pragma unroll
    for(int i=0;i<NACC;i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};"
                   : "+d"(c[i][0]),"+d"(c[i][1]) : "d"(a),"d"(b));
  }
  double s=0;
  #pragma unroll
  for(int i=0;i<NACC;i++) s+=c[i][0]+c[i][1];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}

static float time_ms(cudaEvent_t e0, cudaEvent_t e1){float ms; CK(cudaEventElapsedTime(&ms,e0,e1)); return ms;}

int main(){
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,0));
  int sms=p.multiProcessorCount;
  double* out; CK(cudaMalloc(&out, sizeof(double)*sms*8*256));
  cudaEvent_t e0,e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  printf("{\"gpu\":\"%s\",\"sms\":%d", p.name, sms);
  // DFMA
  {
    const int ILP=8; int iters=20000; int blocks=sms*4;
    dfma_kernel<ILP><<<blocks,256>>>(out,100,1.0000001,1e-9); CK(cudaDeviceSynchronize());
    float best=1e30f;
    for(int r=0;r<5;r++){ CK(cudaEventRecord(e0)); dfma_kernel<ILP><<<blocks,256>>>(out,iters,1.0000001,1e-9); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); best=std::min(best,time_ms(e0,e1)); }
    double flops=2.0*ILP*(double)iters*256.0*blocks;
    printf(",\"dfma_tflops\":%.3f", flops/best/1e9);
  }
  // DMMA
  {
    const int NACC=8; int iters=20000;
    for(int wps=4; wps<=8; wps*=2){ // CTAs per SM (8 warps each)
      int blocks=sms*wps/4*2; if(wps==4) blocks=sms*2; else blocks=sms*4;
      dmma_kernel<NACC><<<blocks,256>>>(out,100); CK(cudaDeviceSynchronize());
      float best=1e30f;
      for(int r=0;r<5;r++){ CK(cudaEventRecord(e0)); dmma_kernel<NACC><<<blocks,256>>>(out,iters); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); best=std::min(best,time_ms(e0,e1)); }
      double flops=2.0*256.0*NACC*(double)iters*8.0*blocks; // 8x8x4 FMA per warp-instr, 8 warps per CTA
      printf(",\"dmma_tflops_%dcta\":%.3f", blocks/sms, flops/best/1e9);
    }
  }
  // cuBLAS DGEMM / DSYRK
  {
    cublasHandle_t h; cublasCreate(&h);
    for(int N : {4096, 8192, 16384}){
      double *A,*B,*C; size_t bytes=sizeof(double)*(size_t)N*N;
      CK(cudaMalloc(&A,bytes)); CK(cudaMalloc(&B,bytes)); CK(cudaMalloc(&C,bytes));
      CK(cudaMemset(A,0,bytes)); CK(cudaMemset(B,0,bytes)); CK(cudaMemset(C,0,bytes));
      double al=-1.0, be=1.0;
      cublasDgemm(h,CUBLAS_OP_N,CUBLAS_OP_T,N,N,N,&al,A,N,B,N,&be,C,N); CK(cudaDeviceSynchronize());
      float best=1e30f;
      for(int r=0;r<5;r++){ CK(cudaEventRecord(e0)); cublasDgemm(h,CUBLAS_OP_N,CUBLAS_OP_T,N,N,N,&al,A,N,B,N,&be,C,N); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); best=std::min(best,time_ms(e0,e1)); }
      printf(",\"cublas_dgemm_nt_%d_tflops\":%.3f", N, 2.0*N*(double)N*N/best/1e9);
      // sustained: back-to-back for ~2 s
      if(N==8192){
        int reps=(int)(2000.0f/best)+1; CK(cudaEventRecord(e0));
        for(int r=0;r<reps;r++) cublasDgemm(h,CUBLAS_OP_N,CUBLAS_OP_T,N,N,N,&al,A,N,B,N,&be,C,N);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        printf(",\"cublas_dgemm_nt_8192_sustained_tflops\":%.3f", 2.0*N*(double)N*N*reps/time_ms(e0,e1)/1e9);
      }
      best=1e30f;
      for(int r=0;r<3;r++){ CK(cudaEventRecord(e0)); cublasDsyrk(h,CUBLAS_FILL_MODE_LOWER,CUBLAS_OP_N,N,N,&al,A,N,&be,C,N); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); best=std::min(best,time_ms(e0,e1)); }
      printf(",\"cublas_dsyrk_%d_tflops\":%.3f", N, (double)N*N*N/best/1e9);
      cudaFree(A);cudaFree(B);cudaFree(C);
    }
    cublasDestroy(h);
  }
  printf("}\n");
  return 0;
}
