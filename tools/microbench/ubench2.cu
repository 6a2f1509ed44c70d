// What does an in-place 19-slot read-modify-write sweep need to reach the HBM peak on B200? Occupancy (warps/SM), block size,
// persistent vs one-tile blocks, and register double-buffering are varied independently; the residency cap is dynamic shared memory.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("err %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} }while(0)
template<typename T, int Q> __global__ void k_rmw(T* base, size_t slot_elems, size_t n_vec){
  size_t i = (size_t)blockIdx.x*blockDim.x+threadIdx.x; if(i>=n_vec) return;
  T v[Q];
  #pragma unroll
  for(int q=0;q<Q;q++) v[q] = base[q*slot_elems+i];
  #pragma unroll
  for(int q=0;q<Q;q++){ unsigned* p = reinterpret_cast<unsigned*>(&v[q]); p[0] += 1u; }
  #pragma unroll
  for(int q=0;q<Q;q++) base[q*slot_elems+i] = v[q];
}
// persistent: grid = resident blocks, each thread strides through the vectors
template<typename T, int Q> __global__ void k_rmw_persist(T* base, size_t slot_elems, size_t n_vec){
  const size_t stride = (size_t)gridDim.x*blockDim.x;
  for(size_t i=(size_t)blockIdx.x*blockDim.x+threadIdx.x; i<n_vec; i+=stride) {
    T v[Q];
    #pragma unroll
    for(int q=0;q<Q;q++) v[q] = base[q*slot_elems+i];
    #pragma unroll
    for(int q=0;q<Q;q++){ unsigned* p = reinterpret_cast<unsigned*>(&v[q]); p[0] += 1u; }
    #pragma unroll
    for(int q=0;q<Q;q++) base[q*slot_elems+i] = v[q];
  }
}
// persistent, contiguous share per block (block b walks its own range): neighbouring iterations of one block are adjacent in memory
template<typename T, int Q> __global__ void k_rmw_persist_contig(T* base, size_t slot_elems, size_t n_vec){
  const size_t per = (n_vec+gridDim.x-1)/gridDim.x, i0 = per*blockIdx.x, i1 = i0+per<n_vec ? i0+per : n_vec;
  for(size_t i=i0+threadIdx.x; i<i1; i+=blockDim.x) {
    T v[Q];
    #pragma unroll
    for(int q=0;q<Q;q++) v[q] = base[q*slot_elems+i];
    #pragma unroll
    for(int q=0;q<Q;q++){ unsigned* p = reinterpret_cast<unsigned*>(&v[q]); p[0] += 1u; }
    #pragma unroll
    for(int q=0;q<Q;q++) base[q*slot_elems+i] = v[q];
  }
}
template<class F> float timeit(F f, int reps=4){ cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1); f(); cudaDeviceSynchronize(); float best=1e30f; for(int r=0;r<reps;r++){ cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms,e0,e1); if(ms<best) best=ms; } return best; }
template<class K> void cap(K k, int smem){ CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); }
int main(){
  const size_t cells = 512ull*512*512;
  void* buf; CK(cudaMalloc(&buf, cells*19*2)); CK(cudaMemset(buf, 0, cells*19*2));
  const size_t n = cells*2/8; // fp16, 8 B per thread and slot
  const double bytes = (double)cells*19*2*2;
  printf("threads/block x blocks/SM = warps/SM : one-tile blocks | persistent strided | persistent contiguous   (GB/s)\n");
  for(int threads : {64, 128, 256, 512}) for(int warps : {8, 16, 24, 32, 40, 48, 64}) {
    const int bps = warps*32/threads; if(bps<1 || bps>32 || bps*threads>2048) continue;
    const int smem = 227*1024/bps-1024;
    cap(k_rmw<uint2,19>, smem); cap(k_rmw_persist<uint2,19>, smem); cap(k_rmw_persist_contig<uint2,19>, smem);
    const unsigned grid1 = (unsigned)((n+threads-1)/threads), gridp = 148u*bps;
    float t1 = timeit([&]{ k_rmw<uint2,19><<<grid1,threads,smem>>>((uint2*)buf, n, n); });
    float t2 = timeit([&]{ k_rmw_persist<uint2,19><<<gridp,threads,smem>>>((uint2*)buf, n, n); });
    float t3 = timeit([&]{ k_rmw_persist_contig<uint2,19><<<gridp,threads,smem>>>((uint2*)buf, n, n); });
    printf("%3d x %2d = %2d warps : %5.0f | %5.0f | %5.0f\n", threads, bps, warps, bytes/t1*1e-6, bytes/t2*1e-6, bytes/t3*1e-6);
  }
  return 0;
}
