// hesaff_b200/csrc/blur_tma.cu -- K1 v3: separable Gaussian blur + det-of-Hessian epilogue, input tile staged by
// TMA (cp.async.bulk.tensor.3d -> UTMALDG), both filter passes on packed f32x2 math (FFMA2 / FADD2 / FMUL2).
//
// Replaces gaussianBlur / cv::GaussianBlur (helpers.cpp:283-295), hessianResponse (pyramid.cpp:63-114) and
// halfImage (helpers.cpp:331-339).  Arithmetic and operation order are identical to k_blur in pyramid.cu (and so to
// OpenCV's): row pass = left-to-right FMA chain stored as fp32, column pass = centre*k0 then (above+below) FMA'd
// outwards; the packed instructions round each half separately, so results are bit-identical.
//
// v2 was bound by the shared-memory pipe (~0.8 wavefronts per pixel incl. bank conflicts, ncu r1b).  v3 is laid out
// around that resource:
//   * tile = 120 x TH outputs; the CTA computes CW = 128 columns [x0-4, x0+124) so that every phase maps one warp to
//     one row segment: lane <-> one float4 (row pass, epilogue) or one f32x2 column pair (column pass).  No warp
//     straddles two rows => no bank conflicts, no integer divisions for indexing;
//   * the filter taps are read from the kernel-parameter constant bank, not from shared memory;
//   * row pass: 4 outputs / lane as two f32x2 accumulators;
//   * column pass: 2 columns x 8 rows / lane (N+7 LDS.64 for 16 outputs instead of N+3 for 8);
//   * epilogue: 3 output rows / warp item, one LDS.128 per row and lane, the +-1 column neighbours by warp shuffle.
// Tried and dropped, each measured slower on B200 with tools/blur_bench.cu: persistent CTAs with the next tile's TMA
// load prefetched under the column pass (needs a third tile buffer => fewer resident CTAs, which costs more than the
// exposed load latency); a TMA store of the L tile (UTMASTG faults on negative start coordinates, the 128-wide box
// writes 7 % more); odd column shifts; one polling thread + CTA barrier instead of all warps polling the mbarrier; the
// Hessian on f32x2 pairs (the register-pair assembly MOVs eat the gain; note ptxas 12.9 contracts mul.rn.f32x2 +
// sub.rn.f32x2 into FFMA2 even under -fmad=false); a scalar row pass.
// 1920 = 16 x 120, so the 1080p pyramid tiles without waste in x.
#include <cuda.h>
#include <stdlib.h>
#include <algorithm>
#include "common.cuh"

namespace blur3 {
constexpr int TW = 120;            // output columns per tile
constexpr int CW = 128;            // computed columns per tile (one float4 per lane)
constexpr int MAXN = 21;
// The TMA box must start at a column that is a multiple of 4 (a 16-byte aligned global address; a misaligned start
// faults -- measured), i.e. RP = roundup4(R) columns left of x0-4.  SHIFT moves the computed columns S = RP - R to the
// left instead of skipping S leading floats in every row-pass load, when that saves a whole LDS.128 per lane and row;
// the column pass then stores its result S columns to the right, so the blurred tile is aligned again.  Only S = 2 is
// used: an odd shift would split the column pass's STS.64 into two conflicting STS.32 (measured: slower).
template <int N, int OH, bool SHIFT> struct Cfg {
   static constexpr int R = N / 2;
   static constexpr int RP = (R + 3) & ~3;                   // columns loaded left of the computed ones
   static constexpr int D0 = RP - R;
   static constexpr int S = (SHIFT && D0 == 2 && (N + 3 + 3) / 4 < (D0 + N + 3 + 3) / 4) ? 2 : 0;
   static constexpr int D = D0 - S;                          // leading floats skipped by the row pass
   static constexpr int NLD = (D + N + 3 + 3) / 4;           // float4 loads per lane in the row pass
   static constexpr int BW = 124 + 4 * NLD;                  // TMA box width (floats)
   static constexpr int TH = OH - 2;                         // output rows per tile
   static constexpr int IH = OH + 2 * R;                     // input rows per tile
   static constexpr size_t SMEM = sizeof(float) * (size_t)(IH * BW + IH * CW) + 16;
};
}

struct Blur3Args {
   float *dstL, *dstR, *half;
   unsigned long long img_stride;
   int W, H, pitch;
   int hW, hH, hpitch;
   float norm2;
};

typedef unsigned long long u64;
__device__ __forceinline__ u64 f2_pack(float lo, float hi)
{
   u64 d;
   asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
   return d;
}
__device__ __forceinline__ u64 f2_fma(u64 a, float b, u64 c)
{
   u64 d, bb;
   asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
   asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(bb), "l"(c));
   return d;
}
__device__ __forceinline__ u64 f2_mul(u64 a, float b)
{
   u64 d, bb;
   asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
   asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(bb));
   return d;
}
__device__ __forceinline__ u64 f2_add(u64 a, u64 b)
{
   u64 d;
   asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
   return d;
}

// Row filter of the output pair (o, o+1) from v[o .. o+N]: same operation order per element as row_taps in pyramid.cu.
template <int N, int NV>
__device__ __forceinline__ u64 row_pair(const float (&v)[NV], int o, const float *__restrict__ k)
{
#define HA_P(i) f2_pack(v[o + (i)], v[o + (i) + 1])
   if (N == 1) return f2_mul(HA_P(0), k[0]);
   if (N == 3) return f2_fma(HA_P(1), k[1], f2_mul(f2_add(HA_P(0), HA_P(2)), k[2]));
   if (N == 5) {
      u64 acc = f2_mul(f2_add(HA_P(1), HA_P(3)), k[3]);
      acc = f2_fma(HA_P(2), k[2], acc);
      return f2_fma(f2_add(HA_P(0), HA_P(4)), k[4], acc);
   }
   u64 acc = f2_mul(HA_P(0), k[0]);
#pragma unroll
   for (int i = 1; i < N; i++) acc = f2_fma(HA_P(i), k[i], acc);
   return acc;
#undef HA_P
}

template <int N, int OH, int NT, int MINB, bool SHIFT, bool SHFL_EDGES>
__global__ void __launch_bounds__(NT, MINB)
k_blur_tma(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ Blur3Args a, const __grid_constant__ Taps taps)
{
   using namespace blur3;
   typedef Cfg<N, OH, SHIFT> C;
   constexpr int R = C::R, RP = C::RP, D = C::D, S = C::S, NLD = C::NLD, BW = C::BW, TH = C::TH, IH = C::IH;
   constexpr int THREADS = NT, NWARPS = NT / 32;
   extern __shared__ __align__(128) float smem[];
   float *sIN = smem;                 // IH x BW   (TMA destination)
   float *sMID = smem + IH * BW;      // IH x CW   row-filtered
   float *sOUT = smem;                // OH x CW   blurred; aliases sIN once the row pass is done
   unsigned long long &mbar = *reinterpret_cast<unsigned long long *>(sMID + IH * CW);

   const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
   const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
   const int gx0 = x0 - 4 - RP, gy0 = y0 - 1 - R;   // image coordinates of sIN[0][0]
   const unsigned bar = (unsigned)__cvta_generic_to_shared(&mbar);
   if (tid == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
      asm volatile("fence.mbarrier_init.release.cluster;");
      const unsigned dst = (unsigned)__cvta_generic_to_shared(sIN);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((unsigned)(IH * BW * sizeof(float))));
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                   ::"r"(dst), "l"(&tmap), "r"(gx0), "r"(gy0), "r"((int)blockIdx.z), "r"(bar)
                   : "memory");
   }
   __syncthreads();
   {
      unsigned ok = 0;
      while (!ok)
         asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                      : "=r"(ok) : "r"(bar), "r"(0) : "memory");
   }

   // ---- BORDER_REPLICATE fix-up: TMA zero-fills outside the image.  Only the out-of-image elements are rewritten:
   // first the columns left / right of the image on the in-image rows, then whole rows above / below. ------------
   {
      const int lc = max(0, -gx0);                  // columns ix < lc are left of the image
      const int rc0 = min(BW, a.W - gx0);           // columns ix >= rc0 are right of it
      const int tr = max(0, -gy0);                  // rows iy < tr are above the image
      const int br0 = min(IH, a.H - gy0);           // rows iy >= br0 are below it
      if (lc > 0 || rc0 < BW) {
         const int ncols = lc + (BW - rc0);
         for (int t = tid; t < (br0 - tr) * ncols; t += THREADS) {
            const int q = t % ncols, iy = tr + t / ncols;
            const int ix = q < lc ? q : rc0 + (q - lc);
            const int sx = q < lc ? lc : rc0 - 1;
            sIN[iy * BW + ix] = sIN[iy * BW + sx];
         }
         __syncthreads();
      }
      if (tr > 0 || br0 < IH) {
         const int nrows = tr + (IH - br0);
         for (int t = tid; t < nrows * (BW / 4); t += THREADS) {
            const int q = t / (BW / 4), ix = t - q * (BW / 4);
            const int iy = q < tr ? q : br0 + (q - tr);
            const int sy = q < tr ? tr : br0 - 1;
            reinterpret_cast<float4 *>(sIN + iy * BW)[ix] = reinterpret_cast<const float4 *>(sIN + sy * BW)[ix];
         }
         __syncthreads();
      }
   }

   // ---- row pass: MID[my][c] = sum_i IN[my][c + D + i] k[i]; one warp per row, lane <-> columns 4*lane .. 4*lane+3 ----
   for (int my = warp; my < IH; my += NWARPS) {
      const float *p = sIN + my * BW + 4 * lane;
      float v[4 * NLD];
      // When only the last float of the first quad / the first float of the last quad is needed (D = 3: 11 and 19 taps),
      // it is the neighbouring lane's: one shuffle instead of a 4-way bank-conflicting scalar load (4 wavefronts).
      constexpr int LASTF = D + N + 2;                                   // last float index read
      constexpr bool HEAD1 = SHFL_EDGES && D == 3;
      constexpr bool TAIL1 = SHFL_EDGES && (LASTF % 4 == 0) && (LASTF / 4 == NLD - 1) && NLD >= 3;
#pragma unroll
      for (int i = (HEAD1 ? 1 : 0); i < (TAIL1 ? NLD - 1 : NLD); i++) {
         const float4 q = *reinterpret_cast<const float4 *>(p + 4 * i);
         v[4 * i + 0] = q.x; v[4 * i + 1] = q.y; v[4 * i + 2] = q.z; v[4 * i + 3] = q.w;
      }
      if (HEAD1) {
         v[0] = v[1] = v[2] = 0.f;
         v[3] = __shfl_up_sync(0xffffffffu, v[7], 1);
         if (lane == 0) v[3] = p[3];
      }
      if (TAIL1) {
         v[LASTF] = __shfl_down_sync(0xffffffffu, v[LASTF - 4], 1);
         v[LASTF + 1] = v[LASTF + 2] = v[LASTF + 3] = 0.f;
         if (lane == 31) v[LASTF] = p[LASTF];
      }
      const u64 o01 = row_pair<N>(v, D, taps.k);
      const u64 o23 = row_pair<N>(v, D + 2, taps.k);
      *reinterpret_cast<ulonglong2 *>(sMID + my * CW + 4 * lane) = make_ulonglong2(o01, o23);
   }
   __syncthreads();

   // ---- column pass: OUT[oy][c] = MID[oy+R][c] k[R] + sum_i (MID[oy+R-i][c] + MID[oy+R+i][c]) k[R+i];
   //      one warp item = 8 rows x 32 column pairs, lane <-> one pair ------------------------------------------------
   for (int item = warp; item < 2 * (OH / 8); item += NWARPS) {
      const int b = item >> 1, cp = ((item & 1) << 5) + lane;
      const u64 *col = reinterpret_cast<const u64 *>(sMID + (8 * b) * CW) + cp;
      u64 m[N + 7];
#pragma unroll
      for (int i = 0; i < N + 7; i++) m[i] = col[i * (CW / 2)];
      u64 *dst = reinterpret_cast<u64 *>(sOUT + (8 * b) * CW - S) + cp;     // MID column c is OUT column c - S (S even)
#pragma unroll
      for (int j = 0; j < 8; j++) {
         u64 acc = f2_mul(m[j + R], taps.k[R]);
#pragma unroll
         for (int i = 1; i <= R; i++) acc = f2_fma(f2_add(m[j + R - i], m[j + R + i]), taps.k[R + i], acc);
         if (S == 0 || 2 * cp >= S) dst[j * (CW / 2)] = acc;
      }
   }
   __syncthreads();

   // ---- write L, the Hessian response R (pyramid.cpp:96-101) and the decimated plane.  OUT row oy <-> image row
   //      y0 - 1 + oy, OUT column c <-> image column x0 - 4 + c.  One warp item = 3 output rows; lane q = 1..30 owns the
   //      output float4 at x0 + 4(q-1); lanes 0 and 31 only feed the shuffles.  The Hessian runs on f32x2 pairs of
   //      neighbouring pixels: (x,y) and (z,w) of a lane's float4 are register pairs as loaded, the pairs shifted by one
   //      column (left|x), (y|z), (w|right) are assembled once per row and serve three output rows. -------------------------
   const size_t ioff = (size_t)blockIdx.z * a.img_stride;
   float *__restrict__ dL = a.dstL + ioff;
   float *__restrict__ dR = a.dstR ? a.dstR + ioff : nullptr;
   float *__restrict__ dH = a.half ? a.half + ioff : nullptr;
   const bool interior = x0 > 0 && y0 > 0 && x0 + TW < a.W && y0 + TH < a.H;   // no output on the image border
   const int gx = x0 + 4 * (lane - 1);
   const bool owner = lane >= 1 && lane <= 30 && gx < a.W;
   for (int item = warp; item < (TH + 2) / 3; item += NWARPS) {
      const int ty0 = 3 * item;
      float4 q[5];
      float lf[5], rt[5];
#pragma unroll
      for (int r = 0; r < 5; r++) {
         const int oy = min(ty0 + r, OH - 1);
         q[r] = *reinterpret_cast<const float4 *>(sOUT + oy * CW + 4 * lane);
         lf[r] = __shfl_up_sync(0xffffffffu, q[r].w, 1);
         rt[r] = __shfl_down_sync(0xffffffffu, q[r].x, 1);
      }
      if (!owner) continue;
#pragma unroll
      for (int rr = 0; rr < 3; rr++) {
         const int ty = ty0 + rr, gy = y0 + ty;
         if (ty >= TH || gy >= a.H) break;
         const float4 c4 = q[rr + 1];
         *reinterpret_cast<float4 *>(dL + (size_t)gy * a.pitch + gx) = c4;
         if (dR) {
            const float uu[6] = {lf[rr], q[rr].x, q[rr].y, q[rr].z, q[rr].w, rt[rr]};
            const float cc[6] = {lf[rr + 1], c4.x, c4.y, c4.z, c4.w, rt[rr + 1]};
            const float dd[6] = {lf[rr + 2], q[rr + 2].x, q[rr + 2].y, q[rr + 2].z, q[rr + 2].w, rt[rr + 2]};
            float r[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
               const float v11 = uu[j], v12 = uu[j + 1], v13 = uu[j + 2];
               const float v21 = cc[j], v22 = cc[j + 1], v23 = cc[j + 2];
               const float v31 = dd[j], v32 = dd[j + 1], v33 = dd[j + 2];
               // v - 2*v22 in one rounding: 2*v22 is exact, so fma(-2, v22, v) == (v - 2*v22) of pyramid.cpp:96-97
               const float Lxx = __fmaf_rn(-2.0f, v22, v21) + v23;
               const float Lyy = __fmaf_rn(-2.0f, v22, v12) + v32;
               const float Lxy = (v13 - v11 + v31 - v33) * 0.25f;
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
               if (hx + 1 < a.hW) *reinterpret_cast<float2 *>(dH + (size_t)hy * a.hpitch + hx) = make_float2(c4.x, c4.z);
               else if (hx < a.hW) dH[(size_t)hy * a.hpitch + hx] = c4.x;
            }
         }
      }
   }
}

// ---- host side ---------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode()
{
   static EncodeTiledFn fn = nullptr;
   static bool tried = false;
   if (!tried) {
      tried = true;
      void *p = nullptr;
      cudaDriverEntryPointQueryResult qres;
      if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
          qres == cudaDriverEntryPointSuccess)
         fn = (EncodeTiledFn)p;
   }
   return fn;
}

template <int N, int OH, int NT, int MINB, bool SHIFT, bool SHFL_EDGES>
static int launch_tma_n(const float *src, const Blur3Args &a, const Taps &taps, int n, cudaStream_t st)
{
   using namespace blur3;
   typedef Cfg<N, OH, SHIFT> C;
   EncodeTiledFn enc = get_encode();
   if (!enc) return -1;
   CUtensorMap tm;
   const cuuint64_t dims[3] = {(cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)n};
   const cuuint64_t strides[2] = {(cuuint64_t)a.pitch * sizeof(float), (cuuint64_t)a.img_stride * sizeof(float)};
   const cuuint32_t box[3] = {(cuuint32_t)C::BW, (cuuint32_t)C::IH, 1};
   const cuuint32_t estr[3] = {1, 1, 1};
   if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void *)src, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return -1;
   // the attribute is per device (a process may hold contexts on several GPUs): cache it per device index
   static bool attr_set[64] = {};   // one array per instantiation
   int dev = 0;
   if (cudaGetDevice(&dev) != cudaSuccess) return -1;
   if (dev < 0 || dev >= 64 || !attr_set[dev]) {
      if (cudaFuncSetAttribute(k_blur_tma<N, OH, NT, MINB, SHIFT, SHFL_EDGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM) != cudaSuccess)
         return -1;
      if (dev >= 0 && dev < 64) attr_set[dev] = true;
   }
   dim3 grid((a.W + TW - 1) / TW, (a.H + C::TH - 1) / C::TH, n);
   k_blur_tma<N, OH, NT, MINB, SHIFT, SHFL_EDGES><<<grid, NT, C::SMEM, st>>>(tm, a, taps);
   return cudaGetLastError() == cudaSuccess ? 0 : -1;   // a failed launch falls back to k_blur (pyramid.cu)
}

#ifndef HA_BLUR_VARIANTS
#define HA_BLUR_VARIANTS 0
#endif

// Returns 0 when the TMA kernel was launched, -1 when this shape/tap count is not covered (caller falls back).
// variant: 0 = default configuration; bench builds (HA_BLUR_VARIANTS) accept oh | (threads/32 << 8) | (min CTAs/SM << 16) |
// (shift << 25) | (shuffled edge floats << 26).
int ha_launch_blur_tma(const float *src, float *dstL, float *dstR, float *half, int W, int H, int pitch, int hW, int hH,
                       int hpitch, unsigned long long img_stride, float norm, const Taps &taps, int n, cudaStream_t st,
                       int variant)
{
   if (taps.n > blur3::MAXN || (pitch & 3) || (img_stride & 3) || ((uintptr_t)src & 15)) return -1;
   Blur3Args a;
   a.dstL = dstL; a.dstR = dstR; a.half = half; a.img_stride = img_stride;
   a.W = W; a.H = H; a.pitch = pitch; a.hW = hW; a.hH = hH; a.hpitch = hpitch;
   a.norm2 = norm * norm;   // pyramid.cpp:76
#if HA_BLUR_VARIANTS
   if (variant) {
#define HA_V(N, OHV, NW, MINB, SH, SE) \
      if (taps.n == N && variant == (OHV | (NW << 8) | (MINB << 16) | (SH << 25) | (SE << 26))) \
         return launch_tma_n<N, OHV, NW * 32, MINB, SH != 0, SE != 0>(src, a, taps, n, st);
#define HA_VN(N) HA_V(N, 40, 8, 4, 1, 0) HA_V(N, 40, 8, 4, 1, 1) HA_V(N, 48, 8, 3, 1, 1) HA_V(N, 56, 8, 3, 1, 1) HA_V(N, 56, 8, 2, 1, 1) HA_V(N, 56, 8, 2, 1, 0)
      HA_VN(9) HA_VN(11) HA_VN(13) HA_VN(15)
#undef HA_VN
#undef HA_V
      return -1;
   }
#endif
   (void)variant;
   // measured on B200 (tools/blur_bench.cu, 32 x 1080p): up to 13 taps 38-row tiles with 4 CTAs/SM are fastest, from 15
   // taps on the halo makes 54-row tiles (2 CTAs/SM, more registers) win
   switch (taps.n) {
#define HA_CASE(N) case N: return launch_tma_n<N, 40, 256, 4, true, true>(src, a, taps, n, st);
#define HA_CASE_TALL(N) case N: return launch_tma_n<N, 56, 256, 2, true, true>(src, a, taps, n, st);
      HA_CASE(1) HA_CASE(3) HA_CASE(5) HA_CASE(7) HA_CASE(9) HA_CASE(11) HA_CASE(13)
      HA_CASE_TALL(15) HA_CASE_TALL(17) HA_CASE_TALL(19) HA_CASE_TALL(21)
#undef HA_CASE
#undef HA_CASE_TALL
   }
   return -1;
}
