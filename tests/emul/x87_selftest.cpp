#define B200_KERNELS_ON_CPU
#include "cuda_on_cpu.h"
#include "../../ansel_b200/csrc/x87.cuh"
#include <cstdio>
#include <cstdlib>
#include <random>
static const long double S3 = 1.7320508075688772935274463415058723669L, S12 = 3.4641016151377545870548926830117447339L;
static uint32_t b(float f){ uint32_t u; memcpy(&u,&f,4); return u; }
int main(){
  std::mt19937_64 rng(7); long bad=0, n=0;
  auto rnd=[&](){ uint32_t u=(uint32_t)rng(); int k=rng()%10; if(k==0) u&=0x807fffffu; /* denormal */ if(k==1) u=(u&0x80000000u)|0x3f800000u|(u&0x7fffff); if(k==2) u=0; if(k==3) u=0x80000000u;
     float f; memcpy(&f,&u,4); if(!std::isfinite(f)) f=1.5f; return f; };
  for(long it=0; it<20000000; it++){
    float f=rnd(), t=rnd(), c=rnd();
    if(it%3==0){ // make t and c/S12 nearly cancel
      c = f; t = -(float)(f/3.4641016f); if(it%6==0) t = std::nextafterf(t, 0.0f); }
    volatile long double m=(long double)f*S3; float w1=(float)m, g1=x87::mul_sqrt3(f);
    volatile long double q=(long double)c/S12; volatile long double s=(long double)t+q; float w2=(float)s, g2=x87::add_div_sqrt12(t,c,1);
    volatile long double s2=(long double)t-q; float w3=(float)s2, g3=x87::add_div_sqrt12(t,c,-1);
    n+=3;
    if(b(w1)!=b(g1)){ if(bad<10) printf("mul f=%a want %a got %a\n",f,w1,g1); bad++; }
    if(b(w2)!=b(g2)){ if(bad<10) printf("add t=%a c=%a want %a got %a\n",t,c,w2,g2); bad++; }
    if(b(w3)!=b(g3)){ if(bad<10) printf("sub t=%a c=%a want %a got %a\n",t,c,w3,g3); bad++; }
  }
  printf("%ld operations, %ld mismatches\n", n, bad); return bad!=0; }
