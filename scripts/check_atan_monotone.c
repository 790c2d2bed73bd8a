#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <omp.h>
#include "/root/repo/include/hg_defined_math.h"
static inline float u2f(uint32_t u){float f; memcpy(&f,&u,4); return f;}
int main(){
  long viol=0; uint32_t firstv=0;
  // all positive finite floats: 0x00000000 .. 0x7f7fffff ; check f(u) <= f(u+1)
  #pragma omp parallel for reduction(+:viol) schedule(static)
  for (int64_t u=0; u<0x7f7fffff; u++){
    float a=hg_atanf(u2f((uint32_t)u)), b=hg_atanf(u2f((uint32_t)u+1));
    if (b<a){ viol++; }
  }
  printf("violations %ld\n", viol);
  if(viol){ int c=0; for (int64_t u=0; u<0x7f7fffff && c<20; u++){ float a=hg_atanf(u2f(u)), b=hg_atanf(u2f(u+1)); if(b<a){ printf("x=%.9g f=%.9g next=%.9g\n", u2f(u), a, b); c++; } } }
  // division by sqrt2 monotone check is trivially true (IEEE)
  return 0;
}
