#include <cstdio>
#include <cstdint>
#include "/root/repo/fluidx3d_b200/csrc/lbm_core.cuh"
using namespace fx3d;
template<int Q, int COLL, bool VF> __global__ void k(unsigned long long n, float S, unsigned long long* out) {
  unsigned long long s = ((unsigned long long)blockIdx.x*blockDim.x+threadIdx.x)*0x9E3779B97F4A7C15ull+0x1234567ull;
  auto next = [&]() { s ^= s<<13; s ^= s>>7; s ^= s<<17; return (uint32_t)(s>>16); };
  auto rnd = [&]() { return ((float)(next()&0xFFFFFF)/16777216.0f-0.5f); };
  unsigned long long c[8] = {0,0,0,0,0,0,0,0};
  const float inv = 1.0f/S;
  for(unsigned long long k=0;k<n;k++){
    float fa[Q], fb[Q]; F2 f2[Q];
    for(int i=0;i<Q;i++){ fa[i] = 0.02f*rnd()*S; fb[i] = 0.02f*rnd()*S; if(S!=1.0f){ fa[i]=rintf(fa[i]); fb[i]=rintf(fb[i]); } }
    static_for<0,Q,1>([&](auto I){ f2[I] = make_f2(fa[I], fb[I]); });
    float ra,uxa,uya,uza, rb,uxb,uyb,uzb; F2 r2,ux2,uy2,uz2;
    moments<Q,float>(fa, S, inv, ra,uxa,uya,uza); moments<Q,float>(fb, S, inv, rb,uxb,uyb,uzb); moments<Q,F2>(f2, S, inv, r2,ux2,uy2,uz2);
    c[0] += (f2_lo(r2)!=ra)+(f2_hi(r2)!=rb);
    c[1] += (f2_lo(ux2)!=uxa)+(f2_hi(ux2)!=uxb)+(f2_lo(uy2)!=uya)+(f2_hi(uy2)!=uyb)+(f2_lo(uz2)!=uza)+(f2_hi(uz2)!=uzb);
    F2 m2 = momentum<Q,0,F2>(f2); c[2] += (f2_lo(m2)!=momentum<Q,0,float>(fa))+(f2_hi(m2)!=momentum<Q,0,float>(fb));
    float fqa[Q], fqb[Q]; F2 fq2[Q];
    equilibrium<Q,float>(ra,uxa,uya,uza,S,fqa); equilibrium<Q,float>(rb,uxb,uyb,uzb,S,fqb); equilibrium<Q,F2>(make_f2(ra,rb),make_f2(uxa,uxb),make_f2(uya,uyb),make_f2(uza,uzb),S,fq2);
    static_for<0,Q,1>([&](auto I){ c[3] += (__float_as_uint(f2_lo(fq2[I]))!=__float_as_uint(fqa[I]))+(__float_as_uint(f2_hi(fq2[I]))!=__float_as_uint(fqb[I])); });
    float o1,o2,o3,o4; F2 p1,p2,p3,p4;
    collide_cell<Q,COLL,VF,float>(fa,S,inv,false,false,1.0f,0.f,0.f,0.f, 1e-4f,-2e-4f,3e-4f, 1.7f, o1,o2,o3,o4);
    collide_cell<Q,COLL,VF,float>(fb,S,inv,false,false,1.0f,0.f,0.f,0.f, 1e-4f,-2e-4f,3e-4f, 1.7f, o1,o2,o3,o4);
    collide_cell<Q,COLL,VF,F2>(f2,S,inv,false,false,vsplat<F2>(1.0f),vsplat<F2>(0.f),vsplat<F2>(0.f),vsplat<F2>(0.f), 1e-4f,-2e-4f,3e-4f, 1.7f, p1,p2,p3,p4);
    static_for<0,Q,1>([&](auto I){ c[4] += (__float_as_uint(f2_lo(f2[I]))!=__float_as_uint(fa[I]))+(__float_as_uint(f2_hi(f2[I]))!=__float_as_uint(fb[I])); });
  }
  for(int i=0;i<8;i++) if(c[i]) atomicAdd(out+i, c[i]);
}
template<int Q,int COLL,bool VF> void run(const char* name, float S){ unsigned long long* d; cudaMalloc(&d, 64); cudaMemset(d,0,64); k<Q,COLL,VF><<<148,128>>>(200, S, d); unsigned long long h[8]; cudaMemcpy(h,d,64,cudaMemcpyDeviceToHost);
 printf("%s S=%g: rho %llu  u %llu  momentum %llu  feq %llu  collide %llu   (%s)\n", name, S, h[0],h[1],h[2],h[3],h[4], cudaGetErrorString(cudaGetLastError())); cudaFree(d); }
int main(){ run<19,0,false>("q19 srt",1.0f); run<19,0,false>("q19 srt",32768.0f); run<19,1,true>("q19 trt vf",1.0f); run<27,1,true>("q27 trt vf",32768.0f); run<27,0,true>("q27 srt vf",1.0f); return 0; }
