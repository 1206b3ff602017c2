// libm_rounding_probe.c -- how often the C library's atan2f / cosf / sinf (what Eigen's computeDirect and the oracle call) differ
// from the correctly rounded value (what the CUDA path computes via float64) over the eigen-solver's argument range.
// Build and run: gcc -O2 -o /tmp/probe tests/libm_rounding_probe.c -lm && /tmp/probe   
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
int main(){ srand(1); long n=5000000, da=0, dc=0, ds=0, dall=0;
 for(long i=0;i<n;i++){ float y=(float)rand()/RAND_MAX*0.3f; float x=((float)rand()/RAND_MAX-0.5f)*0.2f;
  float t1=atan2f(y,x)*(1.0f/3.0f); float t2=(float)atan2((double)y,(double)x)*(1.0f/3.0f);
  if(t1!=t2) da++;
  float c1=cosf(t2), c2=(float)cos((double)t2); if(c1!=c2) dc++;
  float s1=sinf(t2), s2=(float)sin((double)t2); if(s1!=s2) ds++;
  if (t1!=t2 || cosf(t1)!=c2 || sinf(t1)!=s2) dall++; }
 printf("atan2f vs rounded double: %.4f%%  cosf: %.4f%%  sinf: %.4f%%  any: %.4f%%\n",100.0*da/n,100.0*dc/n,100.0*ds/n,100.0*dall/n); return 0; }
