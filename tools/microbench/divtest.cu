#include <cstdio>
#include <cstdint>
#include "/root/repo/fluidx3d_b200/csrc/lbm_core.cuh"
using namespace fx3d;
__global__ void k(unsigned long long n, unsigned long long* out) {
  unsigned long long s = ((unsigned long long)blockIdx.x*blockDim.x+threadIdx.x)*0x9E3779B97F4A7C15ull+0x1234567ull;
  auto next = [&]() { s ^= s<<13; s ^= s>>7; s ^= s<<17; return (uint32_t)(s>>16); };
  unsigned long long c[8] = {0,0,0,0,0,0,0,0};
  for(unsigned long long k=0;k<n;k++){
    float b = __uint_as_float((next()&0x007FFFFFu)|((126u+(next()&1u))<<23));
    float a0 = __uint_as_float((next()&0x807FFFFFu)|((100u+(next()%30u))<<23));
    float a1 = __uint_as_float((next()&0x807FFFFFu)|((100u+(next()%30u))<<23));
    float a2 = __uint_as_float((next()&0x807FFFFFu)|((100u+(next()%30u))<<23));
    float e0=a0/b, e1=a1/b, e2=a2/b;
    float q0,q1,q2; vdiv3(a0,a1,a2,b,q0,q1,q2);
    c[0] += (q0!=e0)+(q1!=e1)+(q2!=e2);
    F2 p0,p1,p2; vdiv3(make_f2(a0,a1), make_f2(a1,a2), make_f2(a2,a0), make_f2(b,b), p0,p1,p2);
    c[1] += (f2_lo(p0)!=e0)+(f2_lo(p1)!=e1)+(f2_lo(p2)!=e2);
    c[2] += (f2_hi(p0)!=e1)+(f2_hi(p1)!=e2)+(f2_hi(p2)!=e0);
    // fusedness of FFMA2
    float x=a0, y=a1, z=-x*y; // exact residual test
    F2 r = vfma(make_f2(x,b), make_f2(y,a2), make_f2(z,a1));
    c[3] += (f2_lo(r)!=fmaf(x,y,z)) + (f2_hi(r)!=fmaf(b,a2,a1));
    F2 ad = vadd(make_f2(a0,b), make_f2(a1,a2)); c[4] += (f2_lo(ad)!=a0+a1)+(f2_hi(ad)!=b+a2);
    F2 mu = vmul(make_f2(a0,b), make_f2(a1,a2)); c[5] += (f2_lo(mu)!=a0*a1)+(f2_hi(mu)!=b*a2);
    F2 su = vsub(make_f2(a0,b), make_f2(a1,a2)); c[6] += (f2_lo(su)!=a0-a1)+(f2_hi(su)!=b-a2);
    c[7] += (vdiv1(0.5f,b)!=0.5f/b);
  }
  for(int i=0;i<8;i++) if(c[i]) atomicAdd(out+i, c[i]);
}
int main(){ unsigned long long* d; cudaMalloc(&d, 64); cudaMemset(d,0,64); k<<<148*4,256>>>(2000, d); unsigned long long h[8]; cudaMemcpy(h,d,64,cudaMemcpyDeviceToHost);
 printf("samples %llu: scalar3 %llu  F2lo %llu  F2hi %llu  ffma2 %llu add2 %llu mul2 %llu sub2 %llu div1 %llu\n", 148ull*4*256*2000, h[0],h[1],h[2],h[3],h[4],h[5],h[6],h[7]); return 0; }
