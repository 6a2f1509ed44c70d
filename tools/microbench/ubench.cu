// Microbenchmarks that decide kernel design on B200: (1) issue throughput of FFMA vs packed FFMA2 (f32x2),
// (2) HBM bandwidth of an in-place read-modify-write sweep with the LBM access shape (19 slots, SoA) at 4/8/16 B per thread.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("err %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} }while(0)
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c){ u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
template<int ILP> __global__ void k_ffma(float* out, float a, float b, int iters){
  float acc[ILP]; for(int i=0;i<ILP;i++) acc[i] = threadIdx.x*0.001f+i;
  for(int it=0; it<iters; it++){
    #pragma unroll
    for(int i=0;i<ILP;i++) acc[i] = fmaf(acc[i], a, b);
  }
  float s=0; for(int i=0;i<ILP;i++) s+=acc[i]; out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template<int ILP> __global__ void k_ffma2(float* out, float a, float b, int iters){
  u64 acc[ILP]; u64 A, B;
  asm("mov.b64 %0, {%1,%1};" : "=l"(A) : "f"(a)); asm("mov.b64 %0, {%1,%1};" : "=l"(B) : "f"(b));
  for(int i=0;i<ILP;i++){ float v = threadIdx.x*0.001f+i; asm("mov.b64 %0, {%1,%1};" : "=l"(acc[i]) : "f"(v)); }
  for(int it=0; it<iters; it++){
    #pragma unroll
    for(int i=0;i<ILP;i++) acc[i] = fma2(acc[i], A, B);
  }
  float s=0; for(int i=0;i<ILP;i++){ float x,y; asm("mov.b64 {%0,%1}, %2;" : "=f"(x), "=f"(y) : "l"(acc[i])); s+=x+y; } out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
// mixed: FFMA interleaved with integer ALU ops (LOP3/IADD) to see dual-pipe issue
template<int ILP> __global__ void k_mix(float* out, float a, float b, int iters){
  float acc[ILP]; unsigned x[ILP]; for(int i=0;i<ILP;i++){ acc[i] = threadIdx.x*0.001f+i; x[i]=threadIdx.x+i; }
  for(int it=0; it<iters; it++){
    #pragma unroll
    for(int i=0;i<ILP;i++){ acc[i] = fmaf(acc[i], a, b); x[i] = (x[i]^0x9e3779b9u)+(x[i]>>3); }
  }
  float s=0; for(int i=0;i<ILP;i++) s+=acc[i]+(float)x[i]; out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
// in-place RMW sweep: thread owns V bytes of each of Q slots (SoA), loads all, adds 1, stores back
template<typename T, int Q> __global__ void __launch_bounds__(128) k_rmw(T* base, size_t slot_elems, size_t n_vec){
  size_t i = (size_t)blockIdx.x*blockDim.x+threadIdx.x; if(i>=n_vec) return;
  T v[Q];
  #pragma unroll
  for(int q=0;q<Q;q++) v[q] = base[q*slot_elems+i];
  #pragma unroll
  for(int q=0;q<Q;q++){ unsigned* p = reinterpret_cast<unsigned*>(&v[q]); p[0] += 1u; }
  #pragma unroll
  for(int q=0;q<Q;q++) base[q*slot_elems+i] = v[q];
}
template<class F> float timeit(F f, int reps=5){ cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1); f(); cudaDeviceSynchronize(); float best=1e30f; for(int r=0;r<reps;r++){ cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms,e0,e1); if(ms<best) best=ms; } return best; }
int main(){
  float* out; CK(cudaMalloc(&out, 148*64*1024*sizeof(float)));
  const int iters=4096;
  for(int warps_per_sm : {4, 8, 16, 32}) {
    int threads=128, blocks=148*warps_per_sm/4;
    float t1 = timeit([&]{ k_ffma<8><<<blocks,threads>>>(out,1.0001f,0.5f,iters); });
    float t2 = timeit([&]{ k_ffma2<8><<<blocks,threads>>>(out,1.0001f,0.5f,iters); });
    float t3 = timeit([&]{ k_mix<8><<<blocks,threads>>>(out,1.0001f,0.5f,iters); });
    double n = (double)blocks*threads/32*iters*8; // warp-instr
    printf("warps/SM=%2d  FFMA: %.1f Gwarp-instr/s (%.1f TFLOP/s)  FFMA2: %.1f Gwarp-instr/s (%.1f TFLOP/s)  mix(FFMA+3 ALU): %.1f Gwarp-instr/s total\n", warps_per_sm,
      n/t1*1e-6, n*64/t1*1e-9, n/t2*1e-6, n*128/t2*1e-9, n*4/t3*1e-6);
  }
  // RMW bandwidth: 19 slots of 2^27 2-byte elements (512^3 FP16) and 4-byte (FP32)
  const size_t cells = 512ull*512*512;
  void* buf; CK(cudaMalloc(&buf, cells*19*4)); CK(cudaMemset(buf, 0, cells*19*4));
  { size_t n=cells*2/4;  float t = timeit([&]{ k_rmw<unsigned,19><<<(unsigned)((n+127)/128),128>>>((unsigned*)buf, n, n); }); printf("RMW fp16 19 slots,  4 B/thread/slot: %.3f ms  %.0f GB/s\n", t, cells*19*2*2/t*1e-6); }
  { size_t n=cells*2/8;  float t = timeit([&]{ k_rmw<uint2,19><<<(unsigned)((n+127)/128),128>>>((uint2*)buf, n, n); }); printf("RMW fp16 19 slots,  8 B/thread/slot: %.3f ms  %.0f GB/s\n", t, cells*19*2*2/t*1e-6); }
  { size_t n=cells*2/16; float t = timeit([&]{ k_rmw<uint4,19><<<(unsigned)((n+127)/128),128>>>((uint4*)buf, n, n); }); printf("RMW fp16 19 slots, 16 B/thread/slot: %.3f ms  %.0f GB/s\n", t, cells*19*2*2/t*1e-6); }
  { size_t n=cells*4/4;  float t = timeit([&]{ k_rmw<unsigned,19><<<(unsigned)((n+127)/128),128>>>((unsigned*)buf, n, n); }); printf("RMW fp32 19 slots,  4 B/thread/slot: %.3f ms  %.0f GB/s\n", t, cells*19*4*2/t*1e-6); }
  { size_t n=cells*4/8;  float t = timeit([&]{ k_rmw<uint2,19><<<(unsigned)((n+127)/128),128>>>((uint2*)buf, n, n); }); printf("RMW fp32 19 slots,  8 B/thread/slot: %.3f ms  %.0f GB/s\n", t, cells*19*4*2/t*1e-6); }
  { size_t n=cells*4/16; float t = timeit([&]{ k_rmw<uint4,19><<<(unsigned)((n+127)/128),128>>>((uint4*)buf, n, n); }); printf("RMW fp32 19 slots, 16 B/thread/slot: %.3f ms  %.0f GB/s\n", t, cells*19*4*2/t*1e-6); }
  // the same sweep with occupancy capped through dynamic shared memory: bandwidth as a function of the bytes in flight per SM
  // (blocks/SM x 128 threads x 19 slots x V bytes) -- tells how deep the stream_collide pipeline has to be
  for(int bps : {1, 2, 3, 4, 6, 8, 12, 16}) {
    const int smem = (int)(227*1024/bps) - 1024;
    CK(cudaFuncSetAttribute(k_rmw<uint2,19>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute(k_rmw<uint4,19>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    size_t n8=cells*2/8, n16=cells*4/16;
    float t8 = timeit([&]{ k_rmw<uint2,19><<<(unsigned)((n8+127)/128),128,smem>>>((uint2*)buf, n8, n8); });
    float t16 = timeit([&]{ k_rmw<uint4,19><<<(unsigned)((n16+127)/128),128,smem>>>((uint4*)buf, n16, n16); });
    printf("blocks/SM=%2d  fp16 8 B/thread: %5.0f GB/s (%3d KB in flight/SM)   fp32 16 B/thread: %5.0f GB/s (%3d KB in flight/SM)\n", bps,
      cells*19*2*2/t8*1e-6, bps*128*19*8/1024, cells*19*4*2/t16*1e-6, bps*128*19*16/1024);
  }
  // plain copy for reference
  { size_t n=cells*19*2/16/2; uint4* a=(uint4*)buf; uint4* b=a+n; float t = timeit([&]{ cudaMemcpyAsync(b,a,n*16,cudaMemcpyDeviceToDevice); }); printf("cudaMemcpy D2D %.1f GB: %.0f GB/s (read+write)\n", n*16e-9, n*16*2/t*1e-6); }
  return 0;
}
