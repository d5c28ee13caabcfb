// micro-benchmark: what bounds the FE scatter phase?  (scratch, not product)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <random>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("%s: %s\n",#x,cudaGetErrorString(e)); exit(1);} }while(0)
constexpr int W=640,H=480; constexpr long long A=(long long)W*H; constexpr int N=1000000;

template<int MODE> __global__ void __launch_bounds__(256) k(const uint2* __restrict__ ev, const double4* __restrict__ lut, const double* __restrict__ dt,
    float4* __restrict__ quad, float* __restrict__ sink, int n, double ox, double oy, double oz){
  long long chunk=(n+gridDim.x-1)/gridDim.x, b=blockIdx.x*chunk, e=min(b+chunk,(long long)n);
  float acc=0.f;
  for(long long i=b+threadIdx.x;i<e;i+=256){
    uint2 r=__ldg(ev+i);
    int ex=r.x&0xffff, ey=r.x>>16;
    if(MODE==0){ acc+= (float)ex; continue; }                       // event stream only
    double d=__ldg(dt+r.y);
    const double2* lp=reinterpret_cast<const double2*>(lut+(ey*W+ex));
    double2 bxy=__ldg(lp); double bz=__ldg(reinterpret_cast<const double*>(lp+1));
    if(MODE==1){ acc+=(float)(bxy.x+bz+d); continue; }               // + LUT/dt loads
    double dlx=ox*d,dly=oy*d,dlz=oz*d;
    double px3=bxy.x+(dly*bz-dlz*bxy.y), py3=bxy.y+(dlz*bxy.x-dlx*bz), pz3=bz+(dlx*bxy.y-dly*bxy.x);
    double inv=1.0/pz3; double px=588.0*(px3*inv)+339.8, py=593.9*(py3*inv)+242.4;
    int xx=(int)px, yy=(int)py;
    bool in = (1<=xx&&xx<W-2&&1<=yy&&yy<H-2);
    float dx=(float)(px-xx), dy=(float)(py-yy);
    if(MODE==2){ acc+=dx+dy+(float)in; continue; }                   // + f64 geometry
    if(in){
      float4 v=make_float4((1.f-dx)*(1.f-dy),dx*(1.f-dy),(1.f-dx)*dy,dx*dy);
      if(MODE==3) atomicAdd(quad+(long long)yy*W+xx, v);             // + v4 red
      if(MODE==4) atomicAdd(&quad[(long long)yy*W+xx].x, v.x);       // scalar red instead
      if(MODE==5) quad[(long long)yy*W+xx]=v;                        // plain store instead
    }
  }
  if(MODE<=2 && acc==123.456f) sink[0]=acc;
}
int main(){
  std::mt19937 rng(1); std::vector<uint2> ev(N), evs(N);
  // events concentrated on 20000 landmarks like the bench packet
  std::vector<int> lx(20000), ly(20000);
  for(int i=0;i<20000;i++){lx[i]=rng()%W; ly[i]=rng()%H;}
  for(int i=0;i<N;i++){ int l=rng()%20000; int x=std::min(W-1,std::max(0,lx[l]+(int)(rng()%7)-3)), y=std::min(H-1,std::max(0,ly[l]+(int)(rng()%7)-3)); ev[i]=make_uint2(x|(y<<16), i/100);}  
  evs=ev; std::sort(evs.begin(),evs.end(),[](uint2 a,uint2 b){int ax=a.x&0xffff,ay=a.x>>16,bx=b.x&0xffff,by=b.x>>16; int ta=(ay/32)*20+ax/32,tb=(by/32)*20+bx/32; return ta<tb;});
  std::vector<double4> lut(A); for(long long i=0;i<A;i++){int x=i%W,y=i/W; lut[i]=make_double4((x-339.8)/588.0,(y-242.4)/593.9,1.0,0);}  
  std::vector<double> dt(N/100); for(int i=0;i<N/100;i++) dt[i]=-0.025+0.05*i/(N/100);
  uint2 *d_ev,*d_evs; double4* d_lut; double* d_dt; float4* d_q; float* d_s;
  CK(cudaMalloc(&d_ev,N*8)); CK(cudaMalloc(&d_evs,N*8)); CK(cudaMalloc(&d_lut,A*32)); CK(cudaMalloc(&d_dt,N/100*8)); CK(cudaMalloc(&d_q,A*16)); CK(cudaMalloc(&d_s,16));
  CK(cudaMemcpy(d_ev,ev.data(),N*8,cudaMemcpyHostToDevice)); CK(cudaMemcpy(d_evs,evs.data(),N*8,cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_lut,lut.data(),A*32,cudaMemcpyHostToDevice)); CK(cudaMemcpy(d_dt,dt.data(),N/100*8,cudaMemcpyHostToDevice)); CK(cudaMemset(d_q,0,A*16));
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const char* names[]={"events only","+LUT/dt loads","+f64 geometry","+red.v4","scalar red","plain store"};
  for(int grid : {444, 1184, 2368}) for(int sorted=0;sorted<2;sorted++){
    const uint2* p = sorted? d_evs: d_ev;
    printf("grid %d %s:", grid, sorted?"binned":"time-order");
    for(int mode=0;mode<6;mode++){
      float best=1e9;
      for(int rep=0;rep<6;rep++){
        cudaEventRecord(e0);
        switch(mode){
          case 0: k<0><<<grid,256>>>(p,d_lut,d_dt,d_q,d_s,N,0.8,-1.0,2.4); break;
          case 1: k<1><<<grid,256>>>(p,d_lut,d_dt,d_q,d_s,N,0.8,-1.0,2.4); break;
          case 2: k<2><<<grid,256>>>(p,d_lut,d_dt,d_q,d_s,N,0.8,-1.0,2.4); break;
          case 3: k<3><<<grid,256>>>(p,d_lut,d_dt,d_q,d_s,N,0.8,-1.0,2.4); break;
          case 4: k<4><<<grid,256>>>(p,d_lut,d_dt,d_q,d_s,N,0.8,-1.0,2.4); break;
          case 5: k<5><<<grid,256>>>(p,d_lut,d_dt,d_q,d_s,N,0.8,-1.0,2.4); break;
        }
        cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms,e0,e1); best=std::min(best,ms);
      }
      printf("  %s %.1fus", names[mode], best*1e3);
    }
    printf("\n");
  }
  CK(cudaDeviceSynchronize());
  return 0;
}
