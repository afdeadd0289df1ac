#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <vector>
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
template<int BW,int BH>
__global__ void kt(const __grid_constant__ CUtensorMap tm, float* out, int x, int y, int z){
  extern __shared__ __align__(128) float buf[];
  unsigned long long &bar = *reinterpret_cast<unsigned long long*>(buf + BW*BH);
  unsigned bar_a=(unsigned)__cvta_generic_to_shared(&bar), dst=(unsigned)__cvta_generic_to_shared(buf);
  if(threadIdx.x==0){ asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;"::"r"(bar_a)); asm volatile("fence.mbarrier_init.release.cluster;"); }
  __syncthreads();
  if(threadIdx.x==0){ asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"::"r"(bar_a),"r"(BW*BH*4));
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"::"r"(dst),"l"(&tm),"r"(x),"r"(y),"r"(z),"r"(bar_a):"memory"); }
  unsigned ok=0; while(!ok){ asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p; }":"=r"(ok):"r"(bar_a),"r"(0):"memory"); }
  for(int i=threadIdx.x;i<BW*BH;i+=blockDim.x) out[i]=buf[i];
}
template<int BW,int BH> int run(EncodeTiledFn enc, float* d, int W,int H,int pitch,int N, size_t istride, float* dout, int x,int y,int z){
  CUtensorMap tm; cuuint64_t dims[3]={(cuuint64_t)W,(cuuint64_t)H,(cuuint64_t)N}; cuuint64_t str[2]={(cuuint64_t)pitch*4,(cuuint64_t)istride*4};
  cuuint32_t box[3]={BW,BH,1}, es[3]={1,1,1};
  CUresult r=enc(&tm,CU_TENSOR_MAP_DATA_TYPE_FLOAT32,3,d,dims,str,box,es,CU_TENSOR_MAP_INTERLEAVE_NONE,CU_TENSOR_MAP_SWIZZLE_NONE,CU_TENSOR_MAP_L2_PROMOTION_L2_128B,CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode BW=%d BH=%d -> %d\n",BW,BH,(int)r); if(r) return 1;
  size_t smem=BW*BH*4+64; cudaFuncSetAttribute(kt<BW,BH>,cudaFuncAttributeMaxDynamicSharedMemorySize,(int)smem);
  kt<BW,BH><<<1,128,smem>>>(tm,dout,x,y,z); cudaError_t e=cudaDeviceSynchronize(); printf("  run -> %s\n",cudaGetErrorString(e));
  if(e) return 1; std::vector<float> h(BW*BH); cudaMemcpy(h.data(),dout,BW*BH*4,cudaMemcpyDeviceToHost); printf("  out[0..3]=%g %g %g %g  out[last]=%g\n",h[0],h[1],h[2],h[3],h[BW*BH-1]); return 0; }
int main(){ void*p=0; cudaDriverEntryPointQueryResult q; cudaFree(0); cudaGetDriverEntryPoint("cuTensorMapEncodeTiled",&p,cudaEnableDefault,&q); EncodeTiledFn enc=(EncodeTiledFn)p; printf("enc=%p q=%d\n",p,(int)q);
  int W=333,H=251,pitch=336,N=2; size_t istride=(size_t)pitch*H+64; float* d; cudaMalloc(&d,istride*N*4); std::vector<float> h(istride*N); for(size_t i=0;i<h.size();i++) h[i]=(float)(i%1000); cudaMemcpy(d,h.data(),h.size()*4,cudaMemcpyHostToDevice);
  float* dout; cudaMalloc(&dout,256*256*4);
  run<64,8>(enc,d,W,H,pitch,N,istride,dout,0,0,0);
  run<64,8>(enc,d,W,H,pitch,N,istride,dout,-3,-2,1);
  run<144,74>(enc,d,W,H,pitch,N,istride,dout,-6,-6,0);
  run<144,74>(enc,d,W,H,pitch,N,istride,dout,-6,114,0);
  run<144,64>(enc,d,W,H,pitch,N,istride,dout,-6,114,0);
  return 0; }
