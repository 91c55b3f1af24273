/* graph_tile_emulation.c -- single-threaded CPU emulation of the SCHEDULE of prox_graph3_tile_kernel (csrc/prox.cu): three diagonally
 * shifted 32x32 tilings, windows entirely inside a tile updated by up to GT_INNER nine-colour sweeps per visit, (xi, tot) state in
 * fp32, dead band NOISE x max|operand| on the applied updates and 8 dead bands on the stop test.  It was written in round 2 to debug
 * the kernel's convergence WITHOUT GPU time: the x = u - tot formulation drifted, the stop test at 1e-6 lambda sat
 * below 1 ulp and ended in a limit cycle, and both were fixed here first.  Test infrastructure only (like oracle/): nothing in the
 * product builds or loads it.
 *   gcc -O2 -DNOISE=2.4e-7f -o graph_tile_emulation graph_tile_emulation.c -lm
 *   ./graph_tile_emulation rows cols lambda tol max_outer u.bin v.bin      (u.bin: float32 [cols][rows] of one frame; prints the change
 *                                                                          per outer iteration and the tile-sweep count to stderr)
 * scripts/emul/run_emulation.py compares it with the oracle's sequential sweeps on the first prox call of a WaterSurface solve. */
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <string.h>
#define GT_T 32
#define GT_PITCH 33
#ifndef GT_INNER
#define GT_INNER 48
#endif
static long work=0;
#define GT_SHIFT 11
static int imax(int a,int b){return a>b?a:b;} static int imin(int a,int b){return a<b?a:b;}
static void cswap(float*a,float*b){float hi=fmaxf(*a,*b),lo=fminf(*a,*b);*a=hi;*b=lo;}
static float clip9(const float*a_in,float z){float u[9];for(int i=0;i<9;i++)u[i]=a_in[i];for(int i=0;i<8;i++)for(int j=0;j<8-i;j++)cswap(&u[j],&u[j+1]);
 const float inv[9]={1.f,0.5f,1.f/3.f,0.25f,0.2f,1.f/6.f,1.f/7.f,0.125f,1.f/9.f};float cs=0,theta=0;for(int k=0;k<9;k++){cs+=u[k];float t=(cs-z)*inv[k];if(u[k]>t)theta=t;}return theta;}
static void geom(int ctr,int rows,int cols,int wi,int wj,int*i0,int*j0,int*hh,int*ww){*i0=ctr?imax(wi-1,0):wi;*j0=ctr?imax(wj-1,0):wj;
 *hh=ctr?imin(wi+1,rows-1)-*i0+1:imin(3,rows-1-wi);*ww=ctr?imin(wj+1,cols-1)-*j0+1:imin(3,cols-1-wj);}
static int inside(int lo,int hi,int off){return (lo-off+GT_T)/GT_T==(hi-off+GT_T)/GT_T;}
int main(int argc,char**argv){
 int rows=atoi(argv[1]),cols=atoi(argv[2]);float lam=atof(argv[3]),tol=atof(argv[4]);int max_outer=atoi(argv[5]);
 int m=rows*cols;float*U=malloc(4*m),*V=malloc(4*m),*TG=calloc(m,4);FILE*f=fopen(argv[6],"rb");fread(U,4,m,f);fclose(f);
 int nwi=rows-imin(3,rows)+1,nwj=cols-imin(3,cols)+1;long nw=(long)nwi*nwj;float*xig=calloc(nw*9,4);
 for(long i=0;i<nw*9;i++)xig[i]=NAN; /* poison: must never be read before written */
 static float xs[GT_T*GT_PITCH],ts[GT_T*GT_PITCH],xis[GT_T*GT_PITCH*9],rads[GT_T*GT_PITCH];
 int outer=0;int ctr=0;
 while(outer<max_outer){float mych=0;
  for(int ph=0;ph<3;ph++){int off=ph*GT_SHIFT;int ntr=(imax(rows-off,0)+GT_T-1)/GT_T+(off>0),ntc=(imax(cols-off,0)+GT_T-1)/GT_T+(off>0);
   for(int tt=0;tt<ntr*ntc;tt++){int tc=tt/ntr,tr=tt-tc*ntr;int rs=off-(off>0?GT_T:0)+tr*GT_T,cs=off-(off>0?GT_T:0)+tc*GT_T;
    int R0=imax(rs,0),R1=imin(rs+GT_T,rows),C0=imax(cs,0),C1=imin(cs+GT_T,cols);if(R1<=R0||C1<=C0)continue;int TR=R1-R0,TC=C1-C0;
    int first=(outer==0&&ph==0);
    for(int idx=0;idx<TR*TC;idx++){int c=idx/TR,r=idx-c*TR;xs[c*GT_PITCH+r]=U[(C0+c)*rows+R0+r];ts[c*GT_PITCH+r]=first?0.f:TG[(C0+c)*rows+R0+r];}
    for(int idx=0;idx<GT_T*GT_T;idx++){int b=idx/GT_T,aa=idx-b*GT_T;int wi=R0+aa,wj=C0+b;float rad=-1;
     if(aa<TR&&b<TC&&wi<nwi&&wj<nwj){int i0,j0,hh,ww;geom(ctr,rows,cols,wi,wj,&i0,&j0,&hh,&ww);
      if(hh>0&&ww>0&&i0>=R0&&i0+hh<=R1&&j0>=C0&&j0+ww<=C1)rad=lam;}
     rads[b*GT_PITCH+aa]=rad;}
    for(int idx=0;idx<TC*TR*9;idx++){int b=idx/(TR*9),rem=idx-b*(TR*9),aa=rem/9,e=rem-aa*9;
     if(rads[b*GT_PITCH+aa]>=0){float v=0;int have=outer>0;
      if(!have&&ph>0){int i0,j0,hh,ww;geom(ctr,rows,cols,R0+aa,C0+b,&i0,&j0,&hh,&ww);for(int q=0;q<ph;q++)have=have||(inside(i0,i0+hh-1,q*GT_SHIFT)&&inside(j0,j0+ww-1,q*GT_SHIFT));}
      if(have)v=xig[((long)(C0+b)*nwi+R0+aa)*9+e];
      xis[(b*GT_PITCH+aa)*9+e]=v;}}
    float tile_ch=0;
    for(int sw=0;sw<GT_INNER;sw++){float ch=0;work++;
     for(int col=0;col<9;col++){int ci=col%3,cj=col/3;int a0=((ci-R0)%3+3)%3,b0=((cj-C0)%3+3)%3;int na=(TR-a0+2)/3,nb=(TC-b0+2)/3;
      for(int q=0;q<na*nb;q++){int qb=q/na;int aa=a0+3*(q-qb*na),b=b0+3*qb;float radius=rads[b*GT_PITCH+aa];if(radius<0)continue;
       int i0,j0,hh,ww;geom(ctr,rows,cols,R0+aa,C0+b,&i0,&j0,&hh,&ww);float*xw=xis+(b*GT_PITCH+aa)*9;int po=(j0-C0)*GT_PITCH+(i0-R0);
       float r[9],ar[9],xo[9];float sabs=0,wscale=0;
       for(int c=0;c<3;c++)for(int dr=0;dr<3;dr++){int e=c*3+dr;float val=0,x0=0;if(dr<hh&&c<ww){float uu=xs[po+c*GT_PITCH+dr];x0=xw[e];val=uu-ts[po+c*GT_PITCH+dr]+x0;wscale=fmaxf(wscale,fmaxf(fabsf(uu),fmaxf(fabsf(x0),fabsf(ts[po+c*GT_PITCH+dr]))));}r[e]=val;ar[e]=fabsf(val);xo[e]=x0;sabs+=fabsf(val);}
       float noise=NOISE*wscale;float theta=0;if(sabs>radius)theta=clip9(ar,radius);
       float xn9[9];float dmax=0;
       for(int e=0;e<9;e++){xn9[e]=copysignf(fmaxf(ar[e]-theta,0.f),r[e]);dmax=fmaxf(dmax,fabsf(xn9[e]-xo[e]));}
       if(dmax>noise){for(int c=0;c<3;c++)for(int dr=0;dr<3;dr++){int e=c*3+dr;if(dr<hh&&c<ww){float dlt=xn9[e]-xo[e];xw[e]=xn9[e];ts[po+c*GT_PITCH+dr]+=dlt;}}if(dmax>8.f*noise)ch=fmaxf(ch,dmax);}
      }}
     tile_ch=fmaxf(tile_ch,ch);if(!(ch>tol))break;}
    mych=fmaxf(mych,tile_ch);
    for(int idx=0;idx<TC*TR*9;idx++){int b=idx/(TR*9),rem=idx-b*(TR*9),aa=rem/9,e=rem-aa*9;if(rads[b*GT_PITCH+aa]>=0)xig[((long)(C0+b)*nwi+R0+aa)*9+e]=xis[(b*GT_PITCH+aa)*9+e];}
    for(int idx=0;idx<TR*TC;idx++){int c=idx/TR,r=idx-c*TR;TG[(C0+c)*rows+R0+r]=ts[c*GT_PITCH+r];}
   }}
  outer++; if(outer<6||outer%100==0)fprintf(stderr,"outer %d change %g\n",outer,mych);
  if(mych<=tol)break;}
 fprintf(stderr,"tile-sweeps %ld outer iterations %d\n",work,outer);
 for(int p=0;p<m;p++)V[p]=U[p]-TG[p];
 f=fopen(argv[7],"wb");fwrite(V,4,m,f);fclose(f);return 0;}
