// Micro-benchmarks that drive the table-placement decisions in DESIGN.md:
// throughput of shared-memory atomics, global RED (u32/u64) on L2-resident tables,
// random L2 loads, and warp ballots on B200. Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); return 1;}}while(0)

__device__ __forceinline__ uint32_t xs(uint32_t& s){ s^=s<<13; s^=s>>17; s^=s<<5; return s; }

template<int MODE>
__global__ void k_smem(uint32_t* out, int iters, int bins){
  extern __shared__ uint32_t sm[];
  for(int i=threadIdx.x;i<bins;i+=blockDim.x) sm[i]=0;
  __syncthreads();
  uint32_t s = (blockIdx.x*blockDim.x+threadIdx.x)*2654435761u+12345u;
  for(int i=0;i<iters;i++){
    uint32_t r = xs(s);
    uint32_t idx;
    if(MODE==0) idx = r % bins;                 // spread random
    else if(MODE==1) idx = 0;                    // single hot bin
    else if(MODE==2) idx = (r & 3);              // 4 hot bins
    else idx = (threadIdx.x&31) + 32*((r>>5)%(bins/32)); // conflict-free banks
    atomicAdd(&sm[idx],1u);
  }
  __syncthreads();
  uint32_t acc=0; for(int i=threadIdx.x;i<bins;i+=blockDim.x) acc+=sm[i];
  if(acc==0xdeadbeef) out[0]=acc;
}

// 64-bit shared-memory REDs (one word carries two 32-bit counters): spread random / conflict-free 8-byte banks
template<int MODE>
__global__ void k_smem64(uint32_t* out, int iters, int bins){
  extern __shared__ unsigned long long sm64[];
  for(int i=threadIdx.x;i<bins;i+=blockDim.x) sm64[i]=0;
  __syncthreads();
  uint32_t s = (blockIdx.x*blockDim.x+threadIdx.x)*2654435761u+12345u;
  for(int i=0;i<iters;i++){
    uint32_t r = xs(s);
    uint32_t idx = MODE==0 ? r % bins : (threadIdx.x&31) + 32*((r>>5)%(bins/32));
    atomicAdd(&sm64[idx],(1ull<<32)|(r&63));
  }
  __syncthreads();
  unsigned long long acc=0; for(int i=threadIdx.x;i<bins;i+=blockDim.x) acc+=sm64[i];
  if(acc==0xdeadbeef) out[0]=(uint32_t)acc;
}

template<typename T, int MODE>
__global__ void k_glob(T* tab, int iters, uint32_t mask){
  uint32_t s = (blockIdx.x*blockDim.x+threadIdx.x)*2654435761u+777u;
  for(int i=0;i<iters;i++){
    uint32_t r = xs(s);
    uint32_t idx = (MODE==0)? (r & mask) : (MODE==1? 0u : (r&15u));
    atomicAdd(&tab[idx],(T)1);
  }
}

__global__ void k_load(const uint32_t* tab, uint32_t* out, int iters, uint32_t mask){
  uint32_t s = (blockIdx.x*blockDim.x+threadIdx.x)*2654435761u+999u;
  uint32_t acc=0;
  for(int i=0;i<iters;i+=4){
    uint32_t a=xs(s)&mask,b=xs(s)&mask,c=xs(s)&mask,d=xs(s)&mask;
    acc += tab[a]+tab[b]+tab[c]+tab[d];
  }
  if(acc==0xdeadbeef) out[0]=acc;
}
// load then CAS-increment of a 4-bit nibble (sketch update), table mostly unsaturated
__global__ void k_cas(uint32_t* tab, int iters, uint32_t mask){
  uint32_t s = (blockIdx.x*blockDim.x+threadIdx.x)*2654435761u+4242u;
  for(int i=0;i<iters;i++){
    uint32_t r=xs(s); uint32_t w=(r>>3)&mask; uint32_t sh=(r&7)*4;
    uint32_t old=tab[w];
    while(((old>>sh)&15u)!=15u){ uint32_t assumed=old; old=atomicCAS(&tab[w],assumed,assumed+(1u<<sh)); if(old==assumed)break; }
  }
}
__global__ void k_ballot(uint32_t* out, int iters){
  uint32_t s = (blockIdx.x*blockDim.x+threadIdx.x)*2654435761u+31u; uint32_t acc=0;
  for(int i=0;i<iters;i++){ uint32_t r=xs(s); 
    acc+=__popc(__ballot_sync(0xffffffffu,(r&3)==0))+__popc(__ballot_sync(0xffffffffu,(r&3)==1))+__popc(__ballot_sync(0xffffffffu,(r&3)==2))+__popc(__ballot_sync(0xffffffffu,(r&3)==3)); }
  if(acc==0xdeadbeef) out[0]=acc;
}
__global__ void k_match(uint32_t* out, int iters){
  uint32_t s = (blockIdx.x*blockDim.x+threadIdx.x)*2654435761u+31u; uint32_t acc=0;
  for(int i=0;i<iters;i++){ uint32_t r=xs(s); acc+=__match_any_sync(0xffffffffu,r&255); }
  if(acc==0xdeadbeef) out[0]=acc;
}

template<typename F> float timeit(F f){ cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b); f(); cudaDeviceSynchronize(); cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms,a,b); return ms; }

int main(){
  int nsm=148; cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,0)); nsm=p.multiProcessorCount; printf("device %s SMs %d\n",p.name,nsm);
  uint32_t* d; CK(cudaMalloc(&d, 64<<20)); CK(cudaMemset(d,0,64<<20));
  const int iters=4096;
  for(int tpb : {256,1024}){
    int blocks = nsm*(2048/tpb);
    double ops = (double)blocks*tpb*iters;
    float ms;
    ms=timeit([&]{k_smem<0><<<blocks,tpb,4096*4>>>(d,iters,1024);}); printf("smem atomic spread 1024 bins  tpb=%d: %.1f Gops/s\n",tpb,ops/ms/1e6);
    ms=timeit([&]{k_smem<0><<<blocks,tpb,4096*4>>>(d,iters,4096);}); printf("smem atomic spread 4096 bins  tpb=%d: %.1f Gops/s\n",tpb,ops/ms/1e6);
    ms=timeit([&]{k_smem<3><<<blocks,tpb,4096*4>>>(d,iters,4096);}); printf("smem atomic conflict-free     tpb=%d: %.1f Gops/s\n",tpb,ops/ms/1e6);
    ms=timeit([&]{k_smem64<0><<<blocks,tpb,4096*8>>>(d,iters,4096);}); printf("smem atomic u64 spread 4096    tpb=%d: %.1f Gops/s\n",tpb,ops/ms/1e6);
    ms=timeit([&]{k_smem64<1><<<blocks,tpb,4096*8>>>(d,iters,4096);}); printf("smem atomic u64 conflict-free  tpb=%d: %.1f Gops/s\n",tpb,ops/ms/1e6);
    ms=timeit([&]{k_smem<1><<<blocks,tpb,4096*4>>>(d,iters,1024);}); printf("smem atomic 1 hot bin         tpb=%d: %.1f Gops/s\n",tpb,ops/ms/1e6);
    ms=timeit([&]{k_smem<2><<<blocks,tpb,4096*4>>>(d,iters,1024);}); printf("smem atomic 4 hot bins        tpb=%d: %.1f Gops/s\n",tpb,ops/ms/1e6);
    ms=timeit([&]{k_glob<uint32_t,0><<<blocks,tpb>>>(d,iters,65535);}); printf("global RED u32 random 64K bins tpb=%d: %.1f Gops/s\n",tpb,ops/ms/1e6);
    ms=timeit([&]{k_glob<unsigned long long,0><<<blocks,tpb>>>((unsigned long long*)d,iters,65535);}); printf("global RED u64 random 64K bins tpb=%d: %.1f Gops/s\n",tpb,ops/ms/1e6);
    ms=timeit([&]{k_glob<unsigned long long,0><<<blocks,tpb>>>((unsigned long long*)d,iters,32767);}); printf("global RED u64 random 32K bins tpb=%d: %.1f Gops/s\n",tpb,ops/ms/1e6);
    ms=timeit([&]{k_glob<uint32_t,0><<<blocks,tpb>>>(d,iters,1023);}); printf("global RED u32 random 1K bins  tpb=%d: %.1f Gops/s\n",tpb,ops/ms/1e6);
    ms=timeit([&]{k_glob<uint32_t,2><<<blocks,tpb>>>(d,iters/16,0);}); printf("global RED u32 16 hot bins     tpb=%d: %.1f Gops/s\n",tpb,ops/16/ms/1e6);
    ms=timeit([&]{k_glob<uint32_t,1><<<blocks,tpb>>>(d,iters/16,0);}); printf("global RED u32 1 hot bin (ptxas may aggregate) tpb=%d: %.1f Gops/s\n",tpb,ops/16/ms/1e6);
    CK(cudaMemset(d,0,64<<20));
    ms=timeit([&]{k_load<<<blocks,tpb>>>(d,d+(16<<20),iters,(2u<<20)-1);}); printf("random 4B loads 8MB table      tpb=%d: %.1f Gops/s\n",tpb,ops/ms/1e6);
    ms=timeit([&]{k_load<<<blocks,tpb>>>(d,d+(16<<20),iters,(128u<<10)-1);}); printf("random 4B loads 512KB table    tpb=%d: %.1f Gops/s\n",tpb,ops/ms/1e6);
    CK(cudaMemset(d,0,64<<20));
    ms=timeit([&]{k_cas<<<blocks,tpb>>>(d,iters/4,(8u<<20)/4-1);}); printf("nibble load+CAS 8MB (2 launches: unsat->sat) tpb=%d: %.1f Gops/s\n",tpb,ops/4/ms/1e6);
    CK(cudaMemset(d,0xff,64<<20));
    ms=timeit([&]{k_cas<<<blocks,tpb>>>(d,iters,(8u<<20)/4-1);}); printf("nibble probe saturated 8MB     tpb=%d: %.1f Gops/s\n",tpb,ops/ms/1e6);
    CK(cudaMemset(d,0,64<<20));
    ms=timeit([&]{k_ballot<<<blocks,tpb>>>(d,iters);}); printf("4x ballot+popc per iter        tpb=%d: %.1f Giter/s (thread-iters)\n",tpb,ops/ms/1e6);
    ms=timeit([&]{k_match<<<blocks,tpb>>>(d,iters/8);}); printf("match_any per iter             tpb=%d: %.1f Giter/s (thread-iters)\n",tpb,ops/8/ms/1e6);
  }
  CK(cudaDeviceSynchronize());
  return 0;
}
