#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <stdint.h>
int main(){ uint64_t s=88172645463325252ull; long bad=0, tot=0;
 for(int den=1; den<=140000; den++){ double d=(double)den, r=1.0/d;
   for(int t=0;t<400;t++){ s^=s<<13; s^=s>>7; s^=s<<17; double a=(double)(s>>11)/9007199254740992.0*(den*0.7)+ (s&1? 1e-9:0.0);
     if(t&1) a=-a*1e-3; if((t&7)==3) a*=1e-6;
     double q=a*r; double rem=fma(-q,d,a); double q1=fma(rem,r,q); double ref=a/d; tot++; if(q1!=ref){ if(bad<5) printf("den %d a %.17g q1 %.17g ref %.17g\n",den,a,q1,ref); bad++; } } }
 printf("mismatches %ld of %ld\n",bad,tot); return 0; }
