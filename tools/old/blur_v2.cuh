// hesaff_b200/csrc/blur_tma.cu -- K1 v2: separable Gaussian blur + det-of-Hessian epilogue with the input
// tile staged by TMA (cp.async.bulk.tensor.3d -> UTMALDG) and the column pass on packed f32x2 math (FFMA2).
//
// Replaces gaussianBlur / cv::GaussianBlur (helpers.cpp:283-295), hessianResponse (pyramid.cpp:63-114) and
// halfImage (helpers.cpp:331-339).  Arithmetic and operation order are identical to k_blur in pyramid.cu (and so to
// OpenCV's): row pass = left-to-right FMA chain stored as fp32, column pass = centre*k0 then (above+below) FMA'd
// outwards; the packed instructions round each lane separately (bit-identical results).
//
// Tile: 128 x 54 outputs per CTA (+1 px ring for the Hessian), 320 threads, up to 3 CTAs per SM.
//   1. one thread arms an mbarrier and issues ONE 3-D TMA box load {x, y, image} of (~132+2R) x (56+2R) floats;
//      out-of-image elements arrive as zeros;
//   2. CTAs that touch the image border rewrite those elements with the clamped (BORDER_REPLICATE) value;
//   3. row pass, 4 outputs / thread from LDS.128 loads; 4. column pass, 2 columns x 4 rows / thread on f32x2;
//   5. epilogue: float4 stores of L and of the Hessian response R (+ the decimated next-octave seed).
#include <cuda.h>
#include "../../hesaff_b200/csrc/common.cuh"

namespace blurv2 {
constexpr int TW = 128, TH = 54;          // outputs written per tile
constexpr int OH = 56;                    // output rows computed per tile (TH + ring, multiple of 4)
constexpr int THREADS = 320;
constexpr int MAXN = 21;
// TMA needs the box's first column 16-byte aligned: the box starts PADL = roundup(R+1, 4) columns left of the
// tile, so the computed output columns start PADO = PADL - R (1..4) columns left of it (>= the 1-px Hessian ring).
template <int R> struct Cfg {
   static constexpr int PADL = ((R + 1) + 3) & ~3;
   static constexpr int PADO = PADL - R;
   static constexpr int OW = ((PADO + TW + 1) + 3) & ~3;     // output columns computed per tile
   static constexpr int BW = ((OW + 2 * R) + 3) & ~3;        // TMA box width (floats)
   static constexpr int IH = OH + 2 * R;
};
}

struct BlurV2Args {
   float *dstL, *dstR, *half;
   unsigned long long img_stride;
   int W, H, pitch;
   int hW, hH, hpitch;
   float norm2;
};

typedef unsigned long long u64v2;
__device__ __forceinline__ u64v2 v2f_fma(u64v2 a, float b, u64v2 c)
{
   u64v2 d, bb;
   asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
   asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(bb), "l"(c));
   return d;
}
__device__ __forceinline__ u64v2 v2f_mul(u64v2 a, float b)
{
   u64v2 d, bb;
   asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
   asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(bb));
   return d;
}
__device__ __forceinline__ u64v2 v2f_add(u64v2 a, u64v2 b)
{
   u64v2 d;
   asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
   return d;
}

template <int N>
__device__ __forceinline__ void row_taps_v2(const float (&in)[N + 3], const float *__restrict__ k, float (&out)[4])
{
#pragma unroll
   for (int j = 0; j < 4; j++) {
      if (N == 1) {
         out[j] = in[j] * k[0];
      } else if (N == 3) {
         out[j] = __fmaf_rn(in[j + 1], k[1], (in[j] + in[j + 2]) * k[2]);
      } else if (N == 5) {
         float acc = (in[j + 1] + in[j + 3]) * k[3];
         acc = __fmaf_rn(in[j + 2], k[2], acc);
         out[j] = __fmaf_rn(in[j] + in[j + 4], k[4], acc);
      } else {
         float acc = in[j] * k[0];
#pragma unroll
         for (int i = 1; i < N; i++) acc = __fmaf_rn(in[j + i], k[i], acc);
         out[j] = acc;
      }
   }
}

template <int N>
__global__ void __launch_bounds__(blurv2::THREADS) k_blur_v2(const __grid_constant__ CUtensorMap tmap, BlurV2Args a, Taps taps)
{
   using namespace blurv2;
   constexpr int R = N / 2;
   constexpr int IH = Cfg<R>::IH, BW = Cfg<R>::BW, OW = Cfg<R>::OW, PADL = Cfg<R>::PADL, PADO = Cfg<R>::PADO;
   extern __shared__ __align__(128) float smem[];
   float *sIN = smem;                 // IH x BW   (TMA destination)
   float *sMID = smem + IH * BW;      // IH x OW
   float *sOUT = smem;                // OH x OW, aliases sIN after the row pass
   float *sk = sMID + IH * OW;        // N taps (padded to 32 floats)
   unsigned long long &mbar = *reinterpret_cast<unsigned long long *>(sk + 32);

   const int tid = threadIdx.x;
   const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
   const int gx0 = x0 - PADL, gy0 = y0 - 1 - R;
   const unsigned bar = (unsigned)__cvta_generic_to_shared(&mbar);
   if (tid == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
      asm volatile("fence.mbarrier_init.release.cluster;");
   }
   if (tid < N) sk[tid] = taps.k[tid];
   __syncthreads();
   if (tid == 0) {
      const unsigned dst = (unsigned)__cvta_generic_to_shared(sIN);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((unsigned)(IH * BW * sizeof(float))));
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                   ::"r"(dst), "l"(&tmap), "r"(gx0), "r"(gy0), "r"((int)blockIdx.z), "r"(bar)
                   : "memory");
   }
   {
      unsigned ok = 0;
      while (!ok)
         asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                      : "=r"(ok) : "r"(bar), "r"(0) : "memory");
   }

   // ---- BORDER_REPLICATE fix-up: TMA zero-fills outside the image; only border CTAs pay for this -------------
   if (gx0 < 0 || gy0 < 0 || gx0 + BW > a.W || gy0 + IH > a.H) {
      for (int t = tid; t < IH * BW; t += THREADS) {
         const int iy = t / BW, ix = t - iy * BW;
         const int gy = gy0 + iy, gx = gx0 + ix;
         const int cy = min(max(gy, 0), a.H - 1), cx = min(max(gx, 0), a.W - 1);
         if (cy != gy || cx != gx) {
            // the clamped source lies inside the box whenever the tile contains an image pixel; sources are
            // in-image elements, which this loop never writes
            const int sy = min(cy - gy0, IH - 1), sx = min(cx - gx0, BW - 1);
            sIN[t] = sIN[sy * BW + sx];
         }
      }
      __syncthreads();
   }

   // ---- row pass: MID[my][ox] = sum_i IN[my][ox+i] k[i] -------------------------------------------------------
   for (int t = tid; t < IH * (OW / 4); t += THREADS) {
      const int my = t / (OW / 4), g = t - my * (OW / 4);
      float in[N + 3];
      const float *p = sIN + my * BW + 4 * g;
#pragma unroll
      for (int i = 0; i < (N + 3 + 3) / 4; i++) {
         const float4 v = *reinterpret_cast<const float4 *>(p + 4 * i);
         if (4 * i + 0 < N + 3) in[4 * i + 0] = v.x;
         if (4 * i + 1 < N + 3) in[4 * i + 1] = v.y;
         if (4 * i + 2 < N + 3) in[4 * i + 2] = v.z;
         if (4 * i + 3 < N + 3) in[4 * i + 3] = v.w;
      }
      float o[4];
      row_taps_v2<N>(in, sk, o);
      *reinterpret_cast<float4 *>(sMID + my * OW + 4 * g) = make_float4(o[0], o[1], o[2], o[3]);
   }
   __syncthreads();

   // ---- column pass on column pairs (f32x2): OUT[oy][ox] = MID[oy+R][ox] k[R] + sum_i (MID[oy+R-i]+MID[oy+R+i]) k[R+i]
   for (int t = tid; t < (OW / 2) * (OH / 4); t += THREADS) {
      const int gy = t / (OW / 2), cp = t - gy * (OW / 2);
      const u64v2 *col = reinterpret_cast<const u64v2 *>(sMID + (4 * gy) * OW + 2 * cp);
      u64v2 m[N + 3];
#pragma unroll
      for (int i = 0; i < N + 3; i++) m[i] = col[i * (OW / 2)];
      u64v2 *dst = reinterpret_cast<u64v2 *>(sOUT + (4 * gy) * OW + 2 * cp);
#pragma unroll
      for (int j = 0; j < 4; j++) {
         u64v2 acc = v2f_mul(m[j + R], sk[R]);
#pragma unroll
         for (int i = 1; i <= R; i++) acc = v2f_fma(v2f_add(m[j + R - i], m[j + R + i]), sk[R + i], acc);
         dst[j * (OW / 2)] = acc;
      }
   }
   __syncthreads();

   // ---- write L, the Hessian response R (pyramid.cpp:96-101) and the decimated plane ------------------------------
   const size_t ioff = (size_t)blockIdx.z * a.img_stride;
   float *__restrict__ dL = a.dstL + ioff;
   float *__restrict__ dR = a.dstR ? a.dstR + ioff : nullptr;
   float *__restrict__ dH = a.half ? a.half + ioff : nullptr;
   const bool interior = x0 > 0 && y0 > 0 && x0 + TW < a.W && y0 + TH < a.H;   // no output on the image border
   for (int t = tid; t < TH * (TW / 4); t += THREADS) {
      const int ty = t / (TW / 4), g = t - ty * (TW / 4);
      const int gy = y0 + ty, gx = x0 + 4 * g;
      if (gy >= a.H || gx >= a.W) continue;
      const float *c = sOUT + (ty + 1) * OW + 4 * g + PADO;   // OUT(ty+1, 4g+PADO) = pixel (gy, gx)
      const float *u = c - OW, *d = c + OW;
      float cc[6], uu[6], dd[6];
#pragma unroll
      for (int j = 0; j < 6; j++) { cc[j] = c[j - 1]; uu[j] = u[j - 1]; dd[j] = d[j - 1]; }
      *reinterpret_cast<float4 *>(dL + (size_t)gy * a.pitch + gx) = make_float4(cc[1], cc[2], cc[3], cc[4]);
      if (dR) {
         float r[4];
#pragma unroll
         for (int j = 0; j < 4; j++) {
            const float v11 = uu[j], v12 = uu[j + 1], v13 = uu[j + 2];
            const float v21 = cc[j], v22 = cc[j + 1], v23 = cc[j + 2];
            const float v31 = dd[j], v32 = dd[j + 1], v33 = dd[j + 2];
            const float t2 = 2 * v22;
            const float Lxx = (v21 - t2 + v23);
            const float Lyy = (v12 - t2 + v32);
            const float Lxy = (v13 - v11 + v31 - v33) / 4.0f;
            r[j] = (Lxx * Lyy - Lxy * Lxy) * a.norm2;
         }
         if (!interior) {
#pragma unroll
            for (int j = 0; j < 4; j++) {
               const int x = gx + j;
               if (gy == 0 || gy == a.H - 1 || x == 0 || x >= a.W - 1) r[j] = 0.f;
            }
         }
         *reinterpret_cast<float4 *>(dR + (size_t)gy * a.pitch + gx) = make_float4(r[0], r[1], r[2], r[3]);
      }
      if (dH && (gy & 1) == 0) {   // halfImage: out(r,c) = in(2r,2c), size rows/2 x cols/2 (helpers.cpp:333-337)
         const int hy = gy >> 1, hx = gx >> 1;
         if (hy < a.hH) {
            if (hx + 1 < a.hW) *reinterpret_cast<float2 *>(dH + (size_t)hy * a.hpitch + hx) = make_float2(cc[1], cc[3]);
            else if (hx < a.hW) dH[(size_t)hy * a.hpitch + hx] = cc[1];
         }
      }
   }
}

// ---- host side ---------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFnV2)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFnV2 get_encode_v2()
{
   static EncodeTiledFnV2 fn = nullptr;
   static bool tried = false;
   if (!tried) {
      tried = true;
      void *p = nullptr;
      cudaDriverEntryPointQueryResult qres;
      if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
          qres == cudaDriverEntryPointSuccess)
         fn = (EncodeTiledFnV2)p;
   }
   return fn;
}

template <int N>
static int launch_v2_n(const float *src, const BlurV2Args &a, const Taps &taps, int n, cudaStream_t st)
{
   using namespace blurv2;
   constexpr int R = N / 2;
   constexpr int IH = Cfg<R>::IH, BW = Cfg<R>::BW, OW = Cfg<R>::OW;
   EncodeTiledFnV2 enc = get_encode_v2();
   if (!enc) return -1;
   CUtensorMap tm;
   const cuuint64_t dims[3] = {(cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)n};
   const cuuint64_t strides[2] = {(cuuint64_t)a.pitch * sizeof(float), (cuuint64_t)a.img_stride * sizeof(float)};
   const cuuint32_t box[3] = {BW, IH, 1};
   const cuuint32_t estr[3] = {1, 1, 1};
   if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void *)src, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return -1;
   const size_t smem = sizeof(float) * (size_t)(IH * BW + IH * OW + 32) + 16;
   cudaFuncSetAttribute(k_blur_v2<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
   dim3 grid((a.W + TW - 1) / TW, (a.H + TH - 1) / TH, n);
   k_blur_v2<N><<<grid, THREADS, smem, st>>>(tm, a, taps);
   return 0;
}

// Returns 0 when the TMA kernel was launched, -1 when this shape/tap count is not covered (caller falls back).
int ha_launch_blur_v2(const float *src, float *dstL, float *dstR, float *half, int W, int H, int pitch, int hW, int hH,
                       int hpitch, unsigned long long img_stride, float norm, const Taps &taps, int n, cudaStream_t st)
{
   if (taps.n > blurv2::MAXN || (pitch & 3) || (img_stride & 3) || ((uintptr_t)src & 15)) return -1;
   BlurV2Args a;
   a.dstL = dstL; a.dstR = dstR; a.half = half; a.img_stride = img_stride;
   a.W = W; a.H = H; a.pitch = pitch; a.hW = hW; a.hH = hH; a.hpitch = hpitch;
   a.norm2 = norm * norm;   // pyramid.cpp:76
   switch (taps.n) {
#define HA_CASE(N) case N: return launch_v2_n<N>(src, a, taps, n, st);
      HA_CASE(1) HA_CASE(3) HA_CASE(5) HA_CASE(7) HA_CASE(9) HA_CASE(11) HA_CASE(13) HA_CASE(15) HA_CASE(17) HA_CASE(19)
      HA_CASE(21)
#undef HA_CASE
   }
   return -1;
}
