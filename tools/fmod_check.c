// Host brute force of the one-FMA fmod(x, 2 pi) used by the NCO phase wrap (csrc/channel.cu fmod_two_pi) against the C
// library: 2e8 arguments in [0, 1e9) including the neighbourhoods of multiples of 2 pi; must report 0 mismatching.
// build: gcc -O2 -ffp-contract=off -o fmod_check tools/fmod_check.c -lm
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
static double f2(double x){ const double y=2.0*3.14159265358979323846; if(x>=0.0&&x<1e9){ const double k=floor(x*(1.0/y)); const double r0=fma(-k,y,x), rm=fma(-(k-1.0),y,x), rp=fma(-(k+1.0),y,x); return (r0<0.0)?rm:((r0>=y)?rp:r0);} return fmod(x,y);}
int main(){ const double y=2.0*3.14159265358979323846; uint64_t s=88172645463325252ULL; long bad=0,n=0;
 for(long i=0;i<200000000;i++){ s^=s<<13; s^=s>>7; s^=s<<17; double u=(s>>11)*(1.0/9007199254740992.0); double x;
  int m=i&3; if(m==0) x=u*70.0; else if(m==1) x=u*1e9; else if(m==2){ double k=floor(u*1e8); x=k*y; int64_t b; memcpy(&b,&x,8); b+=(int)((s>>3)&7)-3; memcpy(&x,&b,8);} else x=u*1e4;
  if(x<0) continue; double a=f2(x), b=fmod(x,y); n++; if(memcmp(&a,&b,8)) { if(bad<5) printf("x=%.17g a=%.17g b=%.17g\n",x,a,b); bad++; } }
 printf("%ld cases, %ld mismatching\n",n,bad); return bad!=0; }
