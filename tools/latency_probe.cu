// Microbenchmark: are ld.global.cg (LDG.STRONG.GPU) loads pipelined?  How expensive are the barrier primitives?
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
__device__ __forceinline__ uint4 ldcg16(const void* p){uint4 r;asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];":"=r"(r.x),"=r"(r.y),"=r"(r.z),"=r"(r.w):"l"(p));return r;}
__device__ __forceinline__ uint4 ldplain16(const void* p){uint4 r;asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];":"=r"(r.x),"=r"(r.y),"=r"(r.z),"=r"(r.w):"l"(p));return r;}
__device__ __forceinline__ uint4 ldrelaxed16(const void* p){uint4 r;asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];":"=r"(r.x),"=r"(r.y),"=r"(r.z),"=r"(r.w):"l"(p));return r;}
__device__ __forceinline__ uint4 ldnc16(const void* p){uint4 r;asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];":"=r"(r.x),"=r"(r.y),"=r"(r.z),"=r"(r.w):"l"(p));return r;}
template<int MODE, int N> __global__ void probe(const char* buf, size_t stride, long long* out, unsigned* sink){
  // one warp; lane reads N independent lines
  uint4 v[N]; long long t0=clock64();
  #pragma unroll
  for(int i=0;i<N;++i){ const char* p=buf+((size_t)i*32+threadIdx.x)*stride;
    if(MODE==0)v[i]=ldcg16(p); else if(MODE==1)v[i]=ldplain16(p); else if(MODE==2)v[i]=ldrelaxed16(p); else v[i]=ldnc16(p);}
  unsigned acc=0;
  #pragma unroll
  for(int i=0;i<N;++i)acc+=v[i].x+v[i].y+v[i].z+v[i].w;
  long long t1=clock64();
  if(threadIdx.x==0)out[0]=t1-t0; sink[threadIdx.x]=acc;
}
__global__ void prims(unsigned* ctr, long long* out, unsigned* sink){
  long long t0=clock64(); __threadfence(); long long t1=clock64();
  atomicAdd(ctr,1u); long long t2=clock64();
  asm volatile("red.release.gpu.global.add.u32 [%0], 1;"::"l"(ctr):"memory"); long long t3=clock64();
  unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];":"=r"(v):"l"(ctr):"memory"); long long t4=clock64();
  unsigned w=atomicAdd(ctr,1u); long long t5=clock64(); sink[0]=v+w;
  unsigned long long g0,g1; asm volatile("mov.u64 %0, %%globaltimer;":"=l"(g0)); long long c0=clock64();
  while(clock64()-c0<200000){} asm volatile("mov.u64 %0, %%globaltimer;":"=l"(g1));
  out[0]=t1-t0; out[1]=t2-t1; out[2]=t3-t2; out[3]=t4-t3; out[4]=t5-t4; out[5]=(long long)(g1-g0); out[6]=clock64()-c0;
}
int main(){ char* buf; size_t bytes=(size_t)1<<30; cudaMalloc(&buf,bytes); cudaMemset(buf,1,bytes);
  long long* out; cudaMallocManaged(&out,64*8); unsigned* sink; cudaMalloc(&sink,4096); unsigned* ctr; cudaMalloc(&ctr,4); cudaMemset(ctr,0,4);
  const char* names[4]={"ld.global.cg","ld.global (plain)","ld.relaxed.gpu","ld.global.nc"};
  for(int rep=0;rep<2;++rep){ // rep0: DRAM (cold lines, big stride), rep1: L2-warm (same lines again)
   for(int m=0;m<4;++m){ size_t stride=4096; const char* b=buf+(size_t)m*(64<<20);
    for(int pass=0;pass<2;++pass){
    if(m==0){probe<0,1><<<1,32>>>(b+(1<<20),stride,out,sink);cudaDeviceSynchronize();long long a=out[0];probe<0,8><<<1,32>>>(b,stride,out,sink);cudaDeviceSynchronize();printf("%-18s pass%d 1 load %lld cyc, 8 loads %lld cyc\n",names[m],pass,a,out[0]);}
    if(m==1){probe<1,1><<<1,32>>>(b+(1<<20),stride,out,sink);cudaDeviceSynchronize();long long a=out[0];probe<1,8><<<1,32>>>(b,stride,out,sink);cudaDeviceSynchronize();printf("%-18s pass%d 1 load %lld cyc, 8 loads %lld cyc\n",names[m],pass,a,out[0]);}
    if(m==2){probe<2,1><<<1,32>>>(b+(1<<20),stride,out,sink);cudaDeviceSynchronize();long long a=out[0];probe<2,8><<<1,32>>>(b,stride,out,sink);cudaDeviceSynchronize();printf("%-18s pass%d 1 load %lld cyc, 8 loads %lld cyc\n",names[m],pass,a,out[0]);}
    if(m==3){probe<3,1><<<1,32>>>(b+(1<<20),stride,out,sink);cudaDeviceSynchronize();long long a=out[0];probe<3,8><<<1,32>>>(b,stride,out,sink);cudaDeviceSynchronize();printf("%-18s pass%d 1 load %lld cyc, 8 loads %lld cyc\n",names[m],pass,a,out[0]);}
    }}
   break; }
  prims<<<1,1>>>(ctr,out,sink); cudaDeviceSynchronize(); prims<<<1,1>>>(ctr,out,sink); cudaDeviceSynchronize();
  printf("threadfence %lld, atomicAdd(noret) %lld, red.release %lld, ld.acquire %lld, atomicAdd(ret) %lld cyc; 200000 cyc spin = %lld ns (=> %.0f MHz)\n",out[0],out[1],out[2],out[3],out[4],out[5],out[6]*1000.0/out[5]);
  return 0; }
