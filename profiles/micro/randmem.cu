// Microbenchmark: random 64-byte entry access patterns over a 1 GiB table (B200), to bound hash-table designs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o randmem randmem.cu && ./randmem
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t h32(uint32_t x){x^=x>>16;x*=0x7feb352du;x^=x>>15;x*=0x846ca68bu;x^=x>>16;return x;}
struct __align__(64) E { unsigned long long w[8]; };
__global__ void k_read64(const E*t,uint32_t mask,uint32_t n,unsigned long long*out){
  uint32_t i=blockIdx.x*blockDim.x+threadIdx.x; if(i>=n)return;
  const uint4*p=reinterpret_cast<const uint4*>(t+(h32(i)&mask));
  uint4 a=p[0],b=p[1],c=p[2];
  if((a.x^b.y^c.z)==0x12345u) out[0]=1;
}
__global__ void k_read32(const E*t,uint32_t mask,uint32_t n,unsigned long long*out){
  uint32_t i=blockIdx.x*blockDim.x+threadIdx.x; if(i>=n)return;
  const uint4*p=reinterpret_cast<const uint4*>(t+(h32(i)&mask));
  uint4 a=p[0],b=p[1];
  if((a.x^b.y)==0x12345u) out[0]=1;
}
__global__ void k_red5(E*t,uint32_t mask,uint32_t n){
  uint32_t i=blockIdx.x*blockDim.x+threadIdx.x; if(i>=n)return;
  E*e=t+(h32(i)&mask);
  #pragma unroll
  for(int k=1;k<6;++k) atomicAdd(&e->w[k],(unsigned long long)i);
}
__global__ void k_red3_32B(E*t,uint32_t mask,uint32_t n){   // all ops inside one 32 B sector
  uint32_t i=blockIdx.x*blockDim.x+threadIdx.x; if(i>=n)return;
  E*e=t+(h32(i)&mask);
  #pragma unroll
  for(int k=1;k<4;++k) atomicAdd(&e->w[k],(unsigned long long)i);
}
__global__ void k_cas_red5(E*t,uint32_t mask,uint32_t n){
  uint32_t i=blockIdx.x*blockDim.x+threadIdx.x; if(i>=n)return;
  E*e=t+(h32(i)&mask);
  unsigned long long k=*reinterpret_cast<volatile unsigned long long*>(&e->w[0]);
  if(k==0) k=atomicCAS(&e->w[0],0ull,(unsigned long long)i+1);
  #pragma unroll
  for(int j=1;j<6;++j) atomicAdd(&e->w[j],(unsigned long long)i+k);
}
__global__ void k_store64(E*t,uint32_t mask,uint32_t n){
  uint32_t i=blockIdx.x*blockDim.x+threadIdx.x; if(i>=n)return;
  uint4*p=reinterpret_cast<uint4*>(t+(h32(i)&mask));
  p[0]=make_uint4(i,0,0,0);p[1]=make_uint4(0,0,0,0);p[2]=make_uint4(0,0,0,0);p[3]=make_uint4(0,0,0,0);
}
__global__ void k_rw64(E*t,uint32_t mask,uint32_t n,unsigned long long*out){   // read entry then clear it (extract pattern)
  uint32_t i=blockIdx.x*blockDim.x+threadIdx.x; if(i>=n)return;
  uint4*p=reinterpret_cast<uint4*>(t+(h32(i)&mask));
  uint4 a=p[0],b=p[1],c=p[2];
  if((a.x^b.y^c.z)==0x12345u) out[0]=1;
  p[0]=make_uint4(0,0,0,0);p[1]=make_uint4(0,0,0,0);p[2]=make_uint4(0,0,0,0);
}
int main(){
  const uint32_t cap=1u<<24, n=7000000; E*t; unsigned long long*out;
  cudaMalloc(&t,(size_t)cap*64); cudaMalloc(&out,8); cudaMemset(t,0,(size_t)cap*64);
  cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b);
  auto run=[&](const char*name,auto f,double bytes_per){
    f(); cudaDeviceSynchronize(); cudaMemset(t,0,(size_t)cap*64); cudaDeviceSynchronize();
    cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms,a,b);
    printf("%-14s %8.3f ms  %7.1f Maccess/ms  useful %6.1f GB/s  (%s)\n",name,ms,n/ms/1e3,n*bytes_per/ms/1e6,cudaGetErrorString(cudaGetLastError()));
  };
  int g=(n+255)/256;
  for(uint32_t c: {1u<<24, 1u<<20}){
    uint32_t mask=c-1; printf("table entries %u (%.0f MB)\n",c,c*64.0/1e6);
    run("read64",[&]{k_read64<<<g,256>>>(t,mask,n,out);},64);
    run("read32",[&]{k_read32<<<g,256>>>(t,mask,n,out);},32);
    run("store64",[&]{k_store64<<<g,256>>>(t,mask,n);},64);
    run("read+clear64",[&]{k_rw64<<<g,256>>>(t,mask,n,out);},128);
    run("red5(64B)",[&]{k_red5<<<g,256>>>(t,mask,n);},40);
    run("red3(32B)",[&]{k_red3_32B<<<g,256>>>(t,mask,n);},24);
    run("cas+red5",[&]{k_cas_red5<<<g,256>>>(t,mask,n);},48);
  }
  return 0;
}
