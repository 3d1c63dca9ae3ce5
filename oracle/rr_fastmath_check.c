// Test infrastructure (not part of the product): CPU restatement of the PTX fast paths of
// rils_rols_b200/csrc/rr_sweep_core.cuh (sin, cos, exp, log) — the same IEEE operations in the same order
// (fma = one rounding) — measured against glibc's long-double functions. Exit status 0 iff every result is
// within 1 ulp of the correctly rounded value on the fast range. The reference evaluates these functions with
// libm (node.cpp:38-50 through Eigen's array sin/cos/log/exp); the tolerance of the path is 1e-9 relative.
// Build: gcc -O2 -mfma -ffp-contract=off rr_fastmath_check.c -lm   (tests/test_cpu_oracle.py runs it)
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
static inline uint64_t d2u(double x){uint64_t u;memcpy(&u,&x,8);return u;}
static inline double u2d(uint64_t u){double x;memcpy(&x,&u,8);return x;}
static const double MAGIC=6755399441055744.0;
// ---- sin / cos: |x| < 65536
static const double TWO_OVER_PI=0.6366197723675814;
static const double PIO2_H=1.5707963267948966, PIO2_M=6.123233995736766e-17, PIO2_L=-1.4973849048591698e-33;
static const double S1=-1.66666666666666324348e-01,S2=8.33333333332248946124e-03,S3=-1.98412698298579493134e-04,
 S4=2.75573137070700676789e-06,S5=-2.50507602534068634195e-08,S6=1.58969099521155010221e-10;
static const double C1=4.16666666666666019037e-02,C2=-1.38888888888741095749e-03,C3=2.48015872894767294178e-05,
 C4=-2.75573143513906633035e-07,C5=2.08757232129817482790e-09,C6=-1.13596475577881948265e-11;
static double sincos_fast(double x,int is_cos){
  double kd=fma(x,TWO_OVER_PI,MAGIC);
  int32_t k=(int32_t)(uint32_t)d2u(kd);
  double kf=kd-MAGIC;
  double r=fma(-kf,PIO2_H,x); r=fma(-kf,PIO2_M,r); r=fma(-kf,PIO2_L,r);
  k+=is_cos;
  int odd=k&1;
  double z=r*r;
  double m=odd?1.0:r;
  double a=z*m;
  // coefficients: sin: [S1..S6,0], cos: [-0.5,C1..C6]
  double c0=odd?-0.5:S1,c1=odd?C1:S2,c2=odd?C2:S3,c3=odd?C3:S4,c4=odd?C4:S5,c5=odd?C5:S6,c6=odd?C6:0.0;
  double p=c6; p=fma(p,z,c5); p=fma(p,z,c4); p=fma(p,z,c3); p=fma(p,z,c2); p=fma(p,z,c1); p=fma(p,z,c0);
  double res=fma(a,p,m);
  uint64_t u=d2u(res); u^=((uint64_t)(k&2))<<62; return u2d(u);
}
// ---- exp: |x| < 700
static const double LOG2E=1.4426950408889634, LN2_H=0.6931471805599453, LN2_L=2.3190468138462996e-17;
static double exp_fast(double x){
  double kd=fma(x,LOG2E,MAGIC);
  int32_t k=(int32_t)(uint32_t)d2u(kd);
  double kf=kd-MAGIC;
  double r=fma(-kf,LN2_H,x); r=fma(-kf,LN2_L,r);
  static const double c[]={1.6059043836821613e-10,2.08767569878681e-09,2.505210838544172e-08,2.755731922398589e-07,
    2.7557319223985893e-06,2.48015873015873e-05,0.0001984126984126984,0.001388888888888889,0.008333333333333333,
    0.041666666666666664,0.16666666666666666,0.5,1.0};
  double p=c[0]; for(int i=1;i<13;i++) p=fma(p,r,c[i]);
  p=fma(p,r,1.0);
  uint64_t u=d2u(p); u+=((uint64_t)(int64_t)k)<<52; return u2d(u);
}
// ---- log: x positive normal
static const double Lg1=6.666666666666735130e-01,Lg2=3.999999999940941908e-01,Lg3=2.857142874366239149e-01,Lg4=2.222219843214978396e-01,
 Lg5=1.818357216161805012e-01,Lg6=1.531383769920937332e-01,Lg7=1.479819860511658591e-01;
static const double ln2_hi=6.93147180369123816490e-01,ln2_lo=1.90821492927058770002e-10;
static double rcp_seed(double b){ double r=1.0/b; uint64_t u=d2u(r)&0xffffffff00000000ull; return u2d(u);} // ~ MUFU.RCP64H
static double log_fast(double x){
  uint64_t ux=d2u(x); uint32_t hx=(uint32_t)(ux>>32);
  int32_t k=(int32_t)(hx>>20)-1023; hx&=0x000fffff;
  uint32_t i=(hx+0x95f64)&0x100000;
  uint32_t mh=hx|(i^0x3ff00000); k+=(int32_t)(i>>20);
  double m=u2d(((uint64_t)mh<<32)|(ux&0xffffffffull));
  double f=m-1.0;
  double d=2.0+f;
  double rr=rcp_seed(d); double e=fma(-d,rr,1.0); e=fma(e,e,e); rr=fma(rr,e,rr); e=fma(-d,rr,1.0); rr=fma(rr,e,rr);
  double s=f*rr;
  double dk=(double)k;
  double z=s*s, w=z*z;
  double t1=w*fma(w,fma(w,Lg6,Lg4),Lg2);
  double t2=z*fma(w,fma(w,fma(w,Lg7,Lg5),Lg3),Lg1);
  double R=t2+t1;
  double hfsq=0.5*f*f;
  // k*ln2_hi - ((hfsq - (s*(hfsq+R) + k*ln2_lo)) - f)
  double q=fma(s,hfsq+R,dk*ln2_lo);
  return dk*ln2_hi-((hfsq-q)-f);
}
static double ulp_err(double got,double ref){ if(got==ref) return 0; if(isnan(got)||isnan(ref)) return (isnan(got)&&isnan(ref))?0:1e9;
  double u=nextafter(fabs(ref),INFINITY)-fabs(ref); long double r=(long double)ref; (void)r; return fabs(got-ref)/u; }
int main(){
  srand48(12345); double mx[4]={0,0,0,0}; long bad[4]={0,0,0,0}; const long N=4000000;
  for(long it=0;it<N;it++){
    double u=drand48(); int mode=it&7; double x;
    if(mode<3) x=(u-0.5)*20.0; else if(mode<5) x=(u-0.5)*2000.0; else if(mode<6) x=(u-0.5)*131000.0; else if(mode<7) x=(u-0.5)*1e-3; else x=ldexp(u-0.5,(int)(drand48()*60)-50);
    if(fabs(x)<65536&&fabs(x)>=7.450580596923828e-09){
      double e=ulp_err(sincos_fast(x,0),sinl((long double)x)); if(e>mx[0])mx[0]=e; if(e>1.0)bad[0]++;
      e=ulp_err(sincos_fast(x,1),cosl((long double)x)); if(e>mx[1])mx[1]=e; if(e>1.0)bad[1]++;
    }
    double xe=(mode<5)?x*0.699:(u-0.5)*1400*0.999;
    if(fabs(xe)<700){ double e=ulp_err(exp_fast(xe),expl((long double)xe)); if(e>mx[2])mx[2]=e; if(e>1.0)bad[2]++; }
    double xl=(mode<3)?fabs(x)+1e-300:(mode<5? 1.0+(u-0.5)*0.9 : ldexp(0.5+u,(int)(drand48()*2040)-1020));
    if(xl>=2.3e-308&&isfinite(xl)){ double e=ulp_err(log_fast(xl),logl((long double)xl)); if(e>mx[3])mx[3]=e; if(e>1.0)bad[3]++; }
  }
  printf("max ulp err: sin %.3f cos %.3f exp %.3f log %.3f ; >1ulp counts %ld %ld %ld %ld of %ld\n",mx[0],mx[1],mx[2],mx[3],bad[0],bad[1],bad[2],bad[3],N);
  // glibc for comparison at the same points is ~0.5 ulp
  printf("spot: sin(1e-300)=%g sin(-0.0)=%g exp(0)=%.17g log(1)=%g log(e)=%.17g\n",sincos_fast(1e-300,0),sincos_fast(-0.0,0),exp_fast(0),log_fast(1.0),log_fast(2.718281828459045));
  return (mx[0]<=1.0&&mx[1]<=1.0&&mx[2]<=1.0&&mx[3]<=1.0)?0:1;
}
