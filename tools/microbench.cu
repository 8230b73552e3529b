// tools/microbench.cu -- design-decision microbenchmarks for the binning / aggregation kernels (B200).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA %s at %d\n",cudaGetErrorString(e),__LINE__); return 1;}}while(0)

__device__ __forceinline__ uint32_t mix(uint32_t x){ x^=x>>16; x*=0x7feb352dU; x^=x>>15; x*=0x846ca68bU; x^=x>>16; return x; }

template<int MODE> // 0: smem atomicAdd (no return), 1: smem atomicAdd returning, 2: smem atomicCAS-probe style, 3: match_any
__global__ void k_smem(uint32_t* out, int iters, int nbins){
  extern __shared__ uint32_t s[];
  for(int i=threadIdx.x;i<nbins;i+=blockDim.x) s[i]=0;
  __syncthreads();
  uint32_t x = blockIdx.x*blockDim.x+threadIdx.x, acc=0;
  for(int i=0;i<iters;i++){
    x = mix(x+i);
    uint32_t b = x & (nbins-1);
    if(MODE==0) atomicAdd(&s[b],1u);
    else if(MODE==1) acc += atomicAdd(&s[b],1u);
    else if(MODE==2){ uint32_t cur=s[b]; if(cur==0) atomicCAS(&s[b],0u,x|1); acc+=cur; atomicAdd(&s[(b+1)&(nbins-1)],1u);}
    else if(MODE==3){ acc += __match_any_sync(0xffffffffu, b&255); }
  }
  __syncthreads();
  if(acc==0x12345) out[0]=acc;
  if(threadIdx.x==0) out[blockIdx.x+1]=s[0];
}

// global: atomicAdd on nhot counters + 4B scatter write
template<int MODE> // 0: atomics only, 1: atomics + scatter write at returned position, 2: scatter write only (pos hashed)
__global__ void k_glob(uint32_t* cnt, uint32_t* outbuf, size_t n, int nhot, uint32_t cap){
  size_t i0 = (size_t)blockIdx.x*blockDim.x+threadIdx.x, stride=(size_t)gridDim.x*blockDim.x;
  for(size_t i=i0;i<n;i+=stride){
    uint32_t x = mix((uint32_t)i*2654435761u+17);
    uint32_t b = x & (nhot-1);
    if(MODE==0) atomicAdd(&cnt[b],1u);
    else if(MODE==1){ uint32_t p=atomicAdd(&cnt[b],1u); outbuf[(size_t)b*cap + (p%cap)] = x; }
    else { outbuf[(size_t)b*cap + ((x>>13)%cap)] = x; }
  }
}

int main(){
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  uint32_t* d; CK(cudaMalloc(&d, 1<<20));
  int nsm=148; float ms;
  for(int nbins : {1024, 8192, 16384}){
    for(int mode=0; mode<4; mode++){
      int iters=4096, threads=512, blocks=nsm*2;
      size_t smem = nbins*4;
      auto launch=[&](){
        if(mode==0){ cudaFuncSetAttribute(k_smem<0>, cudaFuncAttributeMaxDynamicSharedMemorySize,(int)smem); k_smem<0><<<blocks,threads,smem>>>(d,iters,nbins);}
        if(mode==1){ cudaFuncSetAttribute(k_smem<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,(int)smem); k_smem<1><<<blocks,threads,smem>>>(d,iters,nbins);}
        if(mode==2){ cudaFuncSetAttribute(k_smem<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,(int)smem); k_smem<2><<<blocks,threads,smem>>>(d,iters,nbins);}
        if(mode==3){ cudaFuncSetAttribute(k_smem<3>, cudaFuncAttributeMaxDynamicSharedMemorySize,(int)smem); k_smem<3><<<blocks,threads,smem>>>(d,iters,nbins);}
      };
      launch(); CK(cudaDeviceSynchronize());
      cudaEventRecord(e0); launch(); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms,e0,e1);
      double ops=(double)blocks*threads*iters;
      printf("smem mode=%d nbins=%5d : %.3f ms  %.1f Gops/s\n", mode, nbins, ms, ops/ms/1e6);
    }
  }
  size_t n = (size_t)1<<28; uint32_t cap = 65536;
  for(int nhot : {8192, 16384, 1<<20}){
    uint32_t *cnt,*ob; CK(cudaMalloc(&cnt,(size_t)nhot*4)); 
    uint32_t c2 = nhot > 16384 ? 1024 : cap;
    CK(cudaMalloc(&ob,(size_t)nhot*c2*4));
    for(int mode=0;mode<3;mode++){
      CK(cudaMemset(cnt,0,(size_t)nhot*4));
      auto launch=[&](){
        if(mode==0) k_glob<0><<<nsm*8,256>>>(cnt,ob,n,nhot,c2);
        if(mode==1) k_glob<1><<<nsm*8,256>>>(cnt,ob,n,nhot,c2);
        if(mode==2) k_glob<2><<<nsm*8,256>>>(cnt,ob,n,nhot,c2);
      };
      launch(); CK(cudaDeviceSynchronize());
      cudaEventRecord(e0); launch(); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms,e0,e1);
      printf("glob mode=%d nhot=%7d : %.3f ms  %.1f Gops/s\n", mode, nhot, ms, (double)n/ms/1e6);
    }
    cudaFree(cnt); cudaFree(ob);
  }
  return 0;
}
