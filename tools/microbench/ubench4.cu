// How much HBM bandwidth does the bulk-copy stream_collide pipeline reach as a function of the bulk loads in flight per SM?
// Skeleton of the kernel: in-place 19-slot RMW sweep, tile = 128 threads x VB bytes per slot (VB 8: 16-bit DDFs, 16: FP32), moved
// by cp.async.bulk with mbarrier completion. A block has G compute groups of 128 threads that share ONE ring of S stages: tile k of
// the block goes to group k%G and stage k%S, so S-G stages are loading while G are being worked on (G independent blocks with 2
// private stages each -- the round-1 kernel -- is the special case S=2G). `work` dependent FMAs per loaded word stand in for the collision.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("err %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} }while(0)
constexpr int Q = 19;
__device__ __forceinline__ uint32_t saddr(const void* p){ return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int n){ asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(saddr(b)), "r"(n)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes){ asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(saddr(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity){
  asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" :: "r"(saddr(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* b){
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" :: "r"(saddr(dst)), "l"(src), "r"(bytes), "r"(saddr(b)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes){
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(dst), "r"(saddr(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void group_sync(int g){ asm volatile("bar.sync %0, 128;" :: "r"(g+1) : "memory"); }

template<int VB, int G, int S> __global__ void __launch_bounds__(G*128) k_ring(char* base, size_t slot_bytes, size_t ntiles, int work){
  constexpr int ROW = 128*VB, W = VB/4;
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  unsigned char* ring = smem+128;
  const int tid = threadIdx.x, g = tid/128, t = tid%128, warp = t/32, lane = tid%32;
  if(tid==0) { for(int s=0;s<S;s++) mbar_init(full+s, 4); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  const size_t t0 = ntiles*blockIdx.x/gridDim.x, t1 = ntiles*(blockIdx.x+1)/gridDim.x;
  const long n = (long)(t1-t0);
  auto load = [&](long k) { // my warp's share of the copies of tile k (lane 0 only)
    const int s = (int)(k%S); int cnt = 0;
    for(int q=warp;q<Q;q+=4) cnt++;
    mbar_expect(full+s, cnt*ROW);
    for(int q=warp;q<Q;q+=4) bulk_g2s(ring+(size_t)(s*Q+q)*ROW, base+q*slot_bytes+(t0+k)*ROW, ROW, full+s);
  };
  if(lane==0) for(long j=g; j<S && j<n; j+=G) load(j); // prologue: group j%G loads tile j
  for(long k=g; k<n; k+=G) {
    const int s = (int)(k%S);
    if(k>=S && k-G<S) mbar_wait(full+s, (uint32_t)(((k-S)/S)&1)); // early tiles: the stage's previous fill (a prologue load of another group) must have completed first
    mbar_wait(full+s, (uint32_t)((k/S)&1));
    uint32_t v[Q][W];
    const uint32_t* mine = reinterpret_cast<const uint32_t*>(ring+(size_t)s*Q*ROW)+t*W;
    #pragma unroll
    for(int q=0;q<Q;q++) { if constexpr(W==4) { uint4 x = *reinterpret_cast<const uint4*>(mine+q*128*W); v[q][0]=x.x; v[q][1]=x.y; v[q][2]=x.z; v[q][3]=x.w; } else { uint2 x = *reinterpret_cast<const uint2*>(mine+q*128*W); v[q][0]=x.x; v[q][1]=x.y; } }
    group_sync(g);
    const long kp = k-G; // the tile this group finished before: its stage is refilled now that its stores have had time to read it
    if(lane==0 && kp>=0 && kp+S<n) { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); load(kp+S); }
    float acc[4] = { 1.0f, 2.0f, 3.0f, 4.0f };
    for(int it=0; it<work; it++) {
      #pragma unroll
      for(int j=0;j<4;j++) acc[j] = fmaf(acc[j], 1.0001f, 0.5f);
    }
    v[0][0] += 1u; if(acc[0]+acc[1]+acc[2]+acc[3]==0.123f) v[1][0] += 1u;
    uint32_t* out = reinterpret_cast<uint32_t*>(ring+(size_t)s*Q*ROW)+t*W;
    #pragma unroll
    for(int q=0;q<Q;q++) { if constexpr(W==4) *reinterpret_cast<uint4*>(out+q*128*W) = make_uint4(v[q][0], v[q][1], v[q][2], v[q][3]); else *reinterpret_cast<uint2*>(out+q*128*W) = make_uint2(v[q][0], v[q][1]); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    group_sync(g);
    if(lane==0) {
      for(int q=warp;q<Q;q+=4) bulk_s2g(base+q*slot_bytes+(t0+k)*ROW, ring+(size_t)(s*Q+q)*ROW, ROW);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    if(k+G>=n && lane==0 && k+S<n) { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); load(k+S); } // last tile of this group: nobody defers for it
  }
  if(lane==0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// Mixed load paths: slots q<QT move by bulk copies (in-place in shared memory, bulk stores), slots q>=QT by per-thread cp.async
// (LDGSTS, 16/8 bytes) into the same ring and leave straight from registers (STG). One group per block, S=2 (the round-1 shape).
// Question: do the copy engine and the LSU path have separate limits on the bytes in flight, so that sharing the slots between them
// gets closer to the HBM peak than either alone?
template<int VB, int QT> __global__ void __launch_bounds__(128) k_mix(char* base, size_t slot_bytes, size_t ntiles, int work){
  constexpr int ROW = 128*VB, W = VB/4, S = 2;
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  unsigned char* ring = smem+128;
  const int t = threadIdx.x, warp = t/32, lane = t%32;
  if(t==0) { for(int s=0;s<S;s++) mbar_init(full+s, 4); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  const size_t t0 = ntiles*blockIdx.x/gridDim.x, t1 = ntiles*(blockIdx.x+1)/gridDim.x;
  const long n = (long)(t1-t0);
  auto load = [&](long k) {
    const int s = (int)(k%S);
    if(QT>0 && lane==0) {
      int cnt = 0; for(int q=warp;q<QT;q+=4) cnt++;
      mbar_expect(full+s, cnt*ROW);
      for(int q=warp;q<QT;q+=4) bulk_g2s(ring+(size_t)(s*Q+q)*ROW, base+q*slot_bytes+(t0+k)*ROW, ROW, full+s);
    }
    #pragma unroll
    for(int q=QT;q<Q;q++) {
      if constexpr(VB==16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(saddr(ring+(size_t)(s*Q+q)*ROW+t*VB)), "l"(base+q*slot_bytes+(t0+k)*ROW+t*VB) : "memory");
      else asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(saddr(ring+(size_t)(s*Q+q)*ROW+t*VB)), "l"(base+q*slot_bytes+(t0+k)*ROW+t*VB) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  if(n>0) load(0);
  for(long k=0; k<n; k++) {
    const int s = (int)(k%S);
    if(k+1<n) { if(QT>0 && lane==0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); load(k+1); } else asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    if(QT>0) mbar_wait(full+s, (uint32_t)((k/S)&1));
    uint32_t v[Q][W];
    const uint32_t* mine = reinterpret_cast<const uint32_t*>(ring+(size_t)s*Q*ROW)+t*W;
    #pragma unroll
    for(int q=0;q<Q;q++) { if constexpr(W==4) { uint4 x = *reinterpret_cast<const uint4*>(mine+q*128*W); v[q][0]=x.x; v[q][1]=x.y; v[q][2]=x.z; v[q][3]=x.w; } else { uint2 x = *reinterpret_cast<const uint2*>(mine+q*128*W); v[q][0]=x.x; v[q][1]=x.y; } }
    if(QT>0) __syncthreads();
    float acc[4] = { 1.0f, 2.0f, 3.0f, 4.0f };
    for(int it=0; it<work; it++) {
      #pragma unroll
      for(int j=0;j<4;j++) acc[j] = fmaf(acc[j], 1.0001f, 0.5f);
    }
    v[0][0] += 1u; if(acc[0]+acc[1]+acc[2]+acc[3]==0.123f) v[1][0] += 1u;
    uint32_t* out = reinterpret_cast<uint32_t*>(ring+(size_t)s*Q*ROW)+t*W;
    #pragma unroll
    for(int q=0;q<QT;q++) { if constexpr(W==4) *reinterpret_cast<uint4*>(out+q*128*W) = make_uint4(v[q][0], v[q][1], v[q][2], v[q][3]); else *reinterpret_cast<uint2*>(out+q*128*W) = make_uint2(v[q][0], v[q][1]); }
    #pragma unroll
    for(int q=QT;q<Q;q++) { char* g = base+q*slot_bytes+(t0+k)*ROW+t*VB; if constexpr(W==4) *reinterpret_cast<uint4*>(g) = make_uint4(v[q][0], v[q][1], v[q][2], v[q][3]); else *reinterpret_cast<uint2*>(g) = make_uint2(v[q][0], v[q][1]); }
    if(QT>0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
      if(lane==0) {
        for(int q=warp;q<QT;q+=4) bulk_s2g(base+q*slot_bytes+(t0+k)*ROW, ring+(size_t)(s*Q+q)*ROW, ROW);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
  }
  if(lane==0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
template<class F> float timeit(F f, int reps=4){ cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1); f(); CK(cudaDeviceSynchronize()); float best=1e30f; for(int r=0;r<reps;r++){ cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms,e0,e1); if(ms<best) best=ms; } CK(cudaGetLastError()); return best; }
template<int VB, int G, int S> void run(char* buf, size_t cells, int bps, int work){
  constexpr int ROW = 128*VB;
  const size_t slot_bytes = cells*(VB/4), ntiles = slot_bytes/ROW;
  const int need = 128+S*Q*ROW;
  if((size_t)need*bps>227u*1024u) { return; }
  const int smem = 227*1024/bps-1024 < need ? need : 227*1024/bps-1024;
  CK(cudaFuncSetAttribute(k_ring<VB,G,S>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_ring<VB,G,S>, G*128, smem));
  float t = timeit([&]{ k_ring<VB,G,S><<<148*bps, G*128, smem>>>(buf, slot_bytes, ntiles, work); });
  printf("VB %2d  G %d  S %2d  blocks/SM %d (occ %d)  warps/SM %2d  ring %3d KB/SM  loading %2d stages/SM = %3d KB  work %4d : %5.0f GB/s\n", VB, G, S, bps, occ, bps*G*4, bps*S*Q*ROW/1024,
    bps*(S-G), bps*(S-G)*Q*ROW/1024, work, (double)slot_bytes*Q*2/t*1e-6);
}
template<int VB, int QT> void run_mix(char* buf, size_t cells, int bps, int work){
  constexpr int ROW = 128*VB;
  const size_t slot_bytes = cells*(VB/4), ntiles = slot_bytes/ROW;
  const int need = 128+2*Q*ROW;
  if((size_t)need*bps>227u*1024u) return;
  const int smem = 227*1024/bps-1024 < need ? need : 227*1024/bps-1024;
  CK(cudaFuncSetAttribute(k_mix<VB,QT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  float t = timeit([&]{ k_mix<VB,QT><<<148*bps, 128, smem>>>(buf, slot_bytes, ntiles, work); });
  printf("MIX VB %2d  %2d slots by bulk copy + %2d by cp.async/STG  blocks/SM %d  work %4d : %5.0f GB/s\n", VB, QT, Q-QT, bps, work, (double)slot_bytes*Q*2/t*1e-6);
}
int main(int argc, char** argv){
  const size_t cells = 512ull*512*512;
  char* buf; CK(cudaMalloc(&buf, cells*4*Q)); CK(cudaMemset(buf, 0, cells*4*Q));
  for(int work : {0, 200}) {
    printf("---- mixed load paths, FP32-sized tiles, work %d\n", work);
    run_mix<16,19>(buf, cells, 2, work); run_mix<16,14>(buf, cells, 2, work); run_mix<16,10>(buf, cells, 2, work); run_mix<16,5>(buf, cells, 2, work); run_mix<16,0>(buf, cells, 2, work);
    printf("---- mixed load paths, 16-bit-sized tiles, work %d\n", work);
    run_mix<8,19>(buf, cells, 4, work); run_mix<8,14>(buf, cells, 4, work); run_mix<8,10>(buf, cells, 4, work); run_mix<8,5>(buf, cells, 4, work); run_mix<8,0>(buf, cells, 4, work);
  }
  for(int work : {0, 100, 200, 300}) {
    printf("---- FP32-sized tiles (2 KB per slot row), work %d\n", work);
    run<16,1,2>(buf, cells, 2, work); // round-1 kernel shape
    run<16,2,4>(buf, cells, 1, work);
    run<16,2,5>(buf, cells, 1, work);
    run<16,1,5>(buf, cells, 1, work);
    run<16,3,5>(buf, cells, 1, work);
    printf("---- 16-bit-sized tiles (1 KB per slot row), work %d\n", work);
    run<8,1,2>(buf, cells, 4, work);  // round-1 kernel shape
    run<8,2,5>(buf, cells, 2, work);
    run<8,4,11>(buf, cells, 1, work);
    run<8,3,11>(buf, cells, 1, work);
    run<8,4,10>(buf, cells, 1, work);
    run<8,2,4>(buf, cells, 2, work);
    run<8,1,3>(buf, cells, 3, work);
    run<8,3,7>(buf, cells, 1, work);
  }
  unsigned h[4]; CK(cudaMemcpy(h, buf, 16, cudaMemcpyDeviceToHost)); printf("check: word0=%u\n", h[0]);
  return 0;
}
