#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <math.h>
static inline int d2i_rz(double f){ if (f!=f) return 0; if (f>=2147483648.0) return 2147483647; if (f<=-2147483648.0) return -2147483647-1; return (int)f; }
/* tests/table_index_check.c — TEST TOOL: the translated kernels' branch-free LOG/EXP index (fxt_table_index in
 * fx8010-emulator-core_b200/csrc/fx8010_translate.inc) against the reference's static_cast<int>((x - x_min) / step) with the ledger's U6 clamp
 * (as the interpreter kernels and the oracle compute it), over binary32 bit patterns: usage  table_index_check [stride]  (1 = all 2^32). */
#include <stdlib.h>
int main(int argc, char** argv){
  long bad=0; const long long stride = argc > 1 ? atoll(argv[1]) : 1;
  #pragma omp parallel for reduction(+:bad) schedule(static)
  for (long long u=0; u<(1LL<<32); u += stride){
    uint32_t b=(uint32_t)u; float a; memcpy(&a,&b,4);
    double xd=(double)a; int idx0; double di0; unsigned f0=0;
    if (fabsf(a)<=1.0f){ volatile double q=(xd+1.0)*31.5; double t=floor(q)+4503599627370496.0; di0=t-4503599627370496.0; uint64_t w; memcpy(&w,&t,8); idx0=(int)(uint32_t)w; }
    else { volatile double q=(xd+1.0)*31.5; int i=d2i_rz(q); i = i<0?0:(i>63?63:i); idx0=(q<2147483648.0)?i:0; di0=(double)idx0; f0=2; }
    double xc=fmin(fmax(xd,-1.0),1.0); volatile double q=(xc+1.0)*31.5; double t=floor(q)+4503599627370496.0;
    int over=(a>=68174088.0f); double di1= over?0.0:t-4503599627370496.0; uint64_t w; memcpy(&w,&t,8); int idx1= over?0:(int)(uint32_t)w; unsigned f1=(fabsf(a)<=1.0f)?0u:2u;
    if (idx0!=idx1 || di0!=di1 || f0!=f1) { bad++; if (bad<5) printf("mismatch a=%g (0x%08x): %d %g %u vs %d %g %u\n", a,b,idx0,di0,f0,idx1,di1,f1); }
  }
  printf("mismatches: %ld\n", bad); return bad!=0;
}
