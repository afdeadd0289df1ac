// hesaff_b200/csrc/pyramid.cu -- scale-space pyramid, det-of-Hessian response, 3x3x3 extrema,
// sub-pixel/scale localisation and the order-preserving candidate compaction.
//
// Replaces (reference file:line):
//   gaussianBlur / gaussianBlurInplace -> cv::GaussianBlur   helpers.cpp:283-295
//   HessianDetector::hessianResponse                         pyramid.cpp:63-114
//   halfImage                                                helpers.cpp:331-339
//   HessianDetector::findLevelKeypoints, isMax, isMin        pyramid.cpp:206-222, 39-61
//   HessianDetector::localizeKeypoint, solveLinear3x3        pyramid.cpp:122-204, helpers.cpp:46-88
//   getHessianPointType                                      pyramid.cpp:24-37
//   octaveMap dedup                                          pyramid.cpp:189-193,226
#include <stdlib.h>
#include <algorithm>
#include "common.cuh"

// =================================================================================================
// input conversion (hesaff.cpp:138-148: gray = (B+G+R)/3.0f, exact for gray input)
// =================================================================================================
__global__ void k_convert_u8(const uint8_t *__restrict__ src, size_t row_pitch, size_t img_stride, float *__restrict__ arena,
                             int W, int H, int pitch, int pitch8, unsigned long long img_off, unsigned long long img8_off,
                             unsigned long long arena_stride)
{
   // 4 pixels per thread: the float gray image (hesaff.cpp:138-148: B = G = R for a gray file, (3v)/3.0f = v) and a
   // 16-byte-pitched u8 copy, the exact source of the affine patch sampling (describe.cu)
   const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
   const int y = blockIdx.y;
   const int n = blockIdx.z;
   if (x >= pitch8) return;
   const uint8_t *s = src + (size_t)n * img_stride + (size_t)y * row_pitch + x;
   uchar4 b = make_uchar4(0, 0, 0, 0);
   if (x < W) b.x = s[0];
   if (x + 1 < W) b.y = s[1];
   if (x + 2 < W) b.z = s[2];
   if (x + 3 < W) b.w = s[3];
   float *base = arena + (size_t)n * arena_stride;
   *reinterpret_cast<uchar4 *>(reinterpret_cast<unsigned char *>(base + img8_off) + (size_t)y * pitch8 + x) = b;
   if (x < pitch)
      *reinterpret_cast<float4 *>(base + img_off + (size_t)y * pitch + x) = make_float4((float)b.x, (float)b.y, (float)b.z, (float)b.w);   // pitch is a multiple of 4 floats
}

__global__ void k_convert_f32(const float *__restrict__ src, size_t row_pitch_bytes, size_t img_stride_bytes,
                              float *__restrict__ dst, int W, int H, int pitch, unsigned long long arena_stride)
{
   const int x = blockIdx.x * blockDim.x + threadIdx.x;
   const int y = blockIdx.y;
   const int n = blockIdx.z;
   if (x >= W) return;
   const float *s = (const float *)((const char *)src + (size_t)n * img_stride_bytes + (size_t)y * row_pitch_bytes);
   dst[(size_t)n * arena_stride + (size_t)y * pitch + x] = s[x];
}

// 3-channel interleaved 8-bit input: gray = (float(c0) + c1 + c2) / 3.0f, the expression of hesaff.cpp:145 (the sum of
// three bytes is exact in fp32, so the channel order does not matter; the division is IEEE, -prec-div=true)
__global__ void k_convert_rgb8(const uint8_t *__restrict__ src, size_t row_pitch, size_t img_stride, float *__restrict__ dst,
                               int W, int H, int pitch, unsigned long long arena_stride)
{
   const int x = blockIdx.x * blockDim.x + threadIdx.x;
   const int y = blockIdx.y;
   const int n = blockIdx.z;
   if (x >= W) return;
   const uint8_t *s = src + (size_t)n * img_stride + (size_t)y * row_pitch + 3 * (size_t)x;
   dst[(size_t)n * arena_stride + (size_t)y * pitch + x] = ((float)s[0] + (float)s[1] + (float)s[2]) / 3.0f;
}

void ha_launch_convert_rgb8(const uint8_t *src, size_t row_pitch, size_t img_stride, float *dst, const Geom &g, int n,
                            cudaStream_t st, LaunchCounter &lc)
{
   dim3 grid((g.W + 255) / 256, g.H, n);
   k_convert_rgb8<<<grid, 256, 0, st>>>(src, row_pitch, img_stride, dst, g.W, g.H, g.pitch[0], g.arena_stride);
   lc.n++;
}

void ha_launch_convert_u8(const uint8_t *src, size_t row_pitch, size_t img_stride, float *arena, const Geom &g, int n,
                          cudaStream_t st, LaunchCounter &lc)
{
   dim3 grid((g.pitch8 + 4 * 128 - 1) / (4 * 128), g.H, n);
   k_convert_u8<<<grid, 128, 0, st>>>(src, row_pitch, img_stride, arena, g.W, g.H, g.pitch[0], g.pitch8, g.img_off, g.img8_off,
                                      g.arena_stride);
   lc.n++;
}

void ha_launch_convert_f32(const float *src, size_t row_pitch, size_t img_stride, float *dst, const Geom &g, int n,
                           cudaStream_t st, LaunchCounter &lc)
{
   dim3 grid((g.W + 255) / 256, g.H, n);
   k_convert_f32<<<grid, 256, 0, st>>>(src, row_pitch, img_stride, dst, g.W, g.H, g.pitch[0], g.arena_stride);
   lc.n++;
}

// =================================================================================================
// K1: separable Gaussian blur + det-of-Hessian epilogue (+ decimated copy for the next octave)
//
// Operation order = OpenCV's (see oracle/shim/cv_shim.cpp): row pass first, stored as fp32, then the
// column pass.  Row: N>=7 left-to-right FMA chain; N==5, N==3 the small-kernel forms.  Column:
// centre*k0 then (above+below) FMA'd outwards.  BORDER_REPLICATE by clamping the tile load.
//
// Tile: 128 x 30 outputs per CTA (+1 px ring for the Hessian), 256 threads, register-tiled 4 outputs
// per thread in both passes so that shared-memory traffic is ~(N+3)/4 loads per output.
// =================================================================================================
namespace blurcfg {
constexpr int TW = 128, TH = 30;          // outputs written per tile
constexpr int OW = 132, OH = 32;          // outputs computed per tile (ring + padding to x4)
constexpr int THREADS = 256;
}

struct BlurArgs {
   const float *src;
   float *dstL, *dstR, *half;
   unsigned long long img_stride;
   int W, H, pitch;
   int hW, hH, hpitch;
   float norm2;
};

template <int N>
__device__ __forceinline__ void row_taps(const float (&in)[N + 3], const float *__restrict__ k, float (&out)[4])
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
__global__ void __launch_bounds__(blurcfg::THREADS) k_blur(BlurArgs a, Taps taps)
{
   using namespace blurcfg;
   constexpr int R = N / 2;
   constexpr int IH = OH + 2 * R;
   constexpr int IW = ((OW + 2 * R) + 3) & ~3;
   extern __shared__ __align__(16) float smem[];
   float *sIN = smem;                 // IH x IW
   float *sMID = smem + IH * IW;      // IH x OW
   float *sOUT = smem;                // OH x OW, aliases sIN after the row pass
   __shared__ float sk[N];

   const int tid = threadIdx.x;
   const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
   const size_t ioff = (size_t)blockIdx.z * a.img_stride;
   const float *__restrict__ src = a.src + ioff;
   if (tid < N) sk[tid] = taps.k[tid];

   // ---- load the tile with replicate clamping --------------------------------------------------
   const int gx0 = x0 - 1 - R, gy0 = y0 - 1 - R;
   for (int t = tid; t < IH * IW; t += THREADS) {
      const int iy = t / IW, ix = t - iy * IW;
      int gy = gy0 + iy, gx = gx0 + ix;
      gy = min(max(gy, 0), a.H - 1);
      gx = min(max(gx, 0), a.W - 1);
      sIN[t] = __ldg(src + (size_t)gy * a.pitch + gx);
   }
   __syncthreads();

   // ---- row pass: MID[my][ox] = sum_i IN[my][ox+i] k[i] -------------------------------------------
   for (int t = tid; t < IH * (OW / 4); t += THREADS) {
      const int my = t / (OW / 4), g = t - my * (OW / 4);
      float in[N + 3];
      const float *p = sIN + my * IW + 4 * g;
#pragma unroll
      for (int i = 0; i < (N + 3 + 3) / 4; i++) {
         const float4 v = *reinterpret_cast<const float4 *>(p + 4 * i);
         if (4 * i + 0 < N + 3) in[4 * i + 0] = v.x;
         if (4 * i + 1 < N + 3) in[4 * i + 1] = v.y;
         if (4 * i + 2 < N + 3) in[4 * i + 2] = v.z;
         if (4 * i + 3 < N + 3) in[4 * i + 3] = v.w;
      }
      float o[4];
      row_taps<N>(in, sk, o);
      *reinterpret_cast<float4 *>(sMID + my * OW + 4 * g) = make_float4(o[0], o[1], o[2], o[3]);
   }
   __syncthreads();

   // ---- column pass: OUT[oy][ox] = MID[oy+R][ox] k[R] + sum_i (MID[oy+R-i]+MID[oy+R+i]) k[R+i] -----
   for (int t = tid; t < OW * (OH / 4); t += THREADS) {
      const int gy = t / OW, ox = t - gy * OW;
      float m[N + 3];
#pragma unroll
      for (int i = 0; i < N + 3; i++) m[i] = sMID[(4 * gy + i) * OW + ox];
#pragma unroll
      for (int j = 0; j < 4; j++) {
         float acc = m[j + R] * sk[R];
#pragma unroll
         for (int i = 1; i <= R; i++) acc = __fmaf_rn(m[j + R - i] + m[j + R + i], sk[R + i], acc);
         sOUT[(4 * gy + j) * OW + ox] = acc;
      }
   }
   __syncthreads();

   // ---- write L, the Hessian response R (pyramid.cpp:96-101) and the decimated plane -----------------
   float *__restrict__ dL = a.dstL + ioff;
   float *__restrict__ dR = a.dstR ? a.dstR + ioff : nullptr;
   float *__restrict__ dH = a.half ? a.half + ioff : nullptr;
   for (int t = tid; t < TH * (TW / 4); t += THREADS) {
      const int ty = t / (TW / 4), g = t - ty * (TW / 4);
      const int gy = y0 + ty, gx = x0 + 4 * g;
      if (gy >= a.H || gx >= a.W) continue;
      const float *c = sOUT + (ty + 1) * OW + 4 * g + 1;   // OUT(ty+1, 4g+1) = pixel (gy, gx)
      float v[4] = {c[0], c[1], c[2], c[3]};
      *reinterpret_cast<float4 *>(dL + (size_t)gy * a.pitch + gx) = make_float4(v[0], v[1], v[2], v[3]);
      if (dR) {
         float r[4];
         const float *u = c - OW, *d = c + OW;
#pragma unroll
         for (int j = 0; j < 4; j++) {
            const int x = gx + j;
            if (gy == 0 || gy == a.H - 1 || x == 0 || x >= a.W - 1) { r[j] = 0.f; continue; }
            const float v11 = u[j - 1], v12 = u[j], v13 = u[j + 1];
            const float v21 = c[j - 1], v22 = c[j], v23 = c[j + 1];
            const float v31 = d[j - 1], v32 = d[j], v33 = d[j + 1];
            const float Lxx = (v21 - 2 * v22 + v23);
            const float Lyy = (v12 - 2 * v22 + v32);
            const float Lxy = (v13 - v11 + v31 - v33) / 4.0f;
            r[j] = (Lxx * Lyy - Lxy * Lxy) * a.norm2;
         }
         *reinterpret_cast<float4 *>(dR + (size_t)gy * a.pitch + gx) = make_float4(r[0], r[1], r[2], r[3]);
      }
      if (dH && (gy & 1) == 0) {   // halfImage: out(r,c) = in(2r,2c), size rows/2 x cols/2 (helpers.cpp:333-337)
         const int hy = gy >> 1, hx = gx >> 1;
         if (hy < a.hH) {
            if (hx < a.hW) dH[(size_t)hy * a.hpitch + hx] = v[0];
            if (hx + 1 < a.hW) dH[(size_t)hy * a.hpitch + hx + 1] = v[2];
         }
      }
   }
}

template <int N>
static int launch_blur_n(const BlurArgs &a, const Taps &taps, int n, cudaStream_t st)
{
   using namespace blurcfg;
   constexpr int R = N / 2;
   constexpr int IH = OH + 2 * R;
   constexpr int IW = ((OW + 2 * R) + 3) & ~3;
   const size_t smem = sizeof(float) * (size_t)(IH * IW + IH * OW);
   cudaFuncSetAttribute(k_blur<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
   dim3 grid((a.W + TW - 1) / TW, (a.H + TH - 1) / TH, n);
   k_blur<N><<<grid, THREADS, smem, st>>>(a, taps);
   return 0;
}

// ---- any number of taps (sigma is a free parameter of the reference, helpers.cpp:283-289: number_of_scales = 1 or a large
// initial_sigma need more than HA_MAX_TAPS): one thread per pixel, taps from global memory, the same operation order as the
// tiled kernels for n >= 7 (row: left-to-right FMA chain; column: centre first, then (above + below) outwards).  The row
// pass writes into the response plane, which the Hessian pass overwrites afterwards.
__global__ void k_blur_row_generic(const float *__restrict__ src, float *__restrict__ tmp, int W, int H, int pitch,
                                   unsigned long long img_stride, const float *__restrict__ k, int n)
{
   const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
   if (x >= W) return;
   const size_t ioff = (size_t)blockIdx.z * img_stride + (size_t)y * pitch;
   const float *row = src + ioff;
   const int R = n >> 1;
   float acc = row[max(x - R, 0)] * k[0];
   for (int i = 1; i < n; i++) acc = __fmaf_rn(row[min(max(x - R + i, 0), W - 1)], k[i], acc);
   tmp[ioff + x] = acc;
}

__global__ void k_blur_col_generic(const float *__restrict__ tmp, float *__restrict__ dst, int W, int H, int pitch,
                                   unsigned long long img_stride, const float *__restrict__ k, int n)
{
   const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
   if (x >= W) return;
   const size_t ioff = (size_t)blockIdx.z * img_stride;
   const float *col = tmp + ioff + x;
   const int R = n >> 1;
   float acc = col[(size_t)y * pitch] * k[R];
   for (int i = 1; i <= R; i++)
      acc = __fmaf_rn(col[(size_t)max(y - i, 0) * pitch] + col[(size_t)min(y + i, H - 1) * pitch], k[R + i], acc);
   dst[ioff + (size_t)y * pitch + x] = acc;
}

// halfImage, helpers.cpp:331-339
__global__ void k_half(const float *__restrict__ src, float *__restrict__ dst, int pitch, int hW, int hH, int hpitch,
                       unsigned long long img_stride)
{
   const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
   if (x >= hW || y >= hH) return;
   const size_t ioff = (size_t)blockIdx.z * img_stride;
   dst[ioff + (size_t)y * hpitch + x] = src[ioff + (size_t)(2 * y) * pitch + 2 * x];
}

static int launch_blur_generic(const float *src, float *dstL, float *dstR, float *half, int W, int H, int pitch, int hW, int hH,
                               int hpitch, unsigned long long img_stride, float norm, const Taps &taps, int n, cudaStream_t st,
                               LaunchCounter &lc)
{
   if (!taps.dk || !dstR) return -1;
   dim3 grid((W + 127) / 128, H, n);
   k_blur_row_generic<<<grid, 128, 0, st>>>(src, dstR, W, H, pitch, img_stride, taps.dk, taps.n);
   k_blur_col_generic<<<grid, 128, 0, st>>>(dstR, dstL, W, H, pitch, img_stride, taps.dk, taps.n);
   lc.n += 2;
   ha_launch_hessian(dstL, dstR, W, H, pitch, img_stride, norm, n, st, lc);
   if (half) {
      dim3 hg((hW + 127) / 128, hH, n);
      k_half<<<hg, 128, 0, st>>>(dstL, half, pitch, hW, hH, hpitch, img_stride);
      lc.n++;
   }
   return 0;
}

int ha_launch_blur(const float *src, float *dstL, float *dstR, float *half, int W, int H, int pitch, int hW, int hH,
                   int hpitch, unsigned long long img_stride, float norm, const Taps &taps, int n, cudaStream_t st,
                   LaunchCounter &lc)
{
   if (taps.n > HA_MAX_TAPS)
      return launch_blur_generic(src, dstL, dstR, half, W, H, pitch, hW, hH, hpitch, img_stride, norm, taps, n, st, lc);
   static const bool use_tma = getenv("HESAFF_NO_TMA") == nullptr;
   if (use_tma && ha_launch_blur_tma(src, dstL, dstR, half, W, H, pitch, hW, hH, hpitch, img_stride, norm, taps, n, st) == 0) {
      lc.n++;
      return 0;
   }
   BlurArgs a;
   a.src = src; a.dstL = dstL; a.dstR = dstR; a.half = half; a.img_stride = img_stride;
   a.W = W; a.H = H; a.pitch = pitch; a.hW = hW; a.hH = hH; a.hpitch = hpitch;
   a.norm2 = norm * norm;   // pyramid.cpp:76
   lc.n++;
   switch (taps.n) {
#define HA_CASE(N) case N: return launch_blur_n<N>(a, taps, n, st);
      HA_CASE(1) HA_CASE(3) HA_CASE(5) HA_CASE(7) HA_CASE(9) HA_CASE(11) HA_CASE(13) HA_CASE(15) HA_CASE(17)
      HA_CASE(19) HA_CASE(21) HA_CASE(23) HA_CASE(25) HA_CASE(27) HA_CASE(29) HA_CASE(31) HA_CASE(33)
#undef HA_CASE
   }
   lc.n--;
   return -1;
}

// response of a plane that no blur kernel produced (L[0] of octaves >= 1)
__global__ void k_hessian(const float *__restrict__ src, float *__restrict__ dst, int W, int H, int pitch,
                          unsigned long long img_stride, float norm2)
{
   const int x = blockIdx.x * blockDim.x + threadIdx.x;
   const int y = blockIdx.y * blockDim.y + threadIdx.y;
   if (x >= W || y >= H) return;
   const size_t ioff = (size_t)blockIdx.z * img_stride;
   const float *s = src + ioff;
   float r = 0.f;
   if (x > 0 && y > 0 && x < W - 1 && y < H - 1) {
      const float *u = s + (size_t)(y - 1) * pitch + x, *c = u + pitch, *d = c + pitch;
      const float v11 = u[-1], v12 = u[0], v13 = u[1];
      const float v21 = c[-1], v22 = c[0], v23 = c[1];
      const float v31 = d[-1], v32 = d[0], v33 = d[1];
      const float Lxx = (v21 - 2 * v22 + v23);
      const float Lyy = (v12 - 2 * v22 + v32);
      const float Lxy = (v13 - v11 + v31 - v33) / 4.0f;
      r = (Lxx * Lyy - Lxy * Lxy) * norm2;
   }
   dst[ioff + (size_t)y * pitch + x] = r;
}

void ha_launch_hessian(const float *src, float *dst, int W, int H, int pitch, unsigned long long img_stride, float norm,
                       int n, cudaStream_t st, LaunchCounter &lc)
{
   dim3 block(32, 8), grid((W + 31) / 32, (H + 7) / 8, n);
   k_hessian<<<grid, block, 0, st>>>(src, dst, W, H, pitch, img_stride, norm * norm);
   lc.n++;
}

// =================================================================================================
// K2a: 3x3x3 extrema -> candidate bitmask (findLevelKeypoints + isMax/isMin, pyramid.cpp:206-222,39-61)
// One warp = 32 consecutive pixels of a row = one mask word (ballot).  Every word is written.
// =================================================================================================
// Up to NMS_MAXL consecutive levels per launch: the 3x3 spatial max / min of every response plane is computed once and
// shared by the (up to three) levels whose 3x3x3 neighbourhood contains it, and every plane is read once per launch
// instead of once per level (S = 3: 5 plane reads instead of 9, ~2.6x fewer instructions).
#define NMS_MAXL 3
struct NmsArgs {
   const float *plane[NMS_MAXL + 2];     // R[l0-1] .. R[l0+NL]
   uint32_t *mask[NMS_MAXL];             // candidate bitmask of levels l0 .. l0+NL-1
   unsigned long long img_stride;
   int W, H, pitch, border, wpr;
   float posThr, negThr;
   unsigned long long mask_stride;
};

// Each warp owns a strip of 32 columns (= one mask word per row and level) and marches down NMS_ROWS rows keeping, per
// response plane, the horizontal 3-max / 3-min of the two previous rows in registers.
#define NMS_ROWS 32
#define NMS_WARPS 4

template <int NL>
__global__ void __launch_bounds__(NMS_WARPS * 32) k_nms(NmsArgs a)
{
   constexpr int NP = NL + 2;
   const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
   const int wcol = blockIdx.x * NMS_WARPS + wid;
   if (wcol >= a.wpr) return;
   const int c = wcol * 32 + lane;
   const int cc = min(c, a.W - 1);                       // clamped column for the loads of out-of-image lanes
   // the strip's outer neighbour columns: lane 0 fetches column c-1, lane 31 column c+1 (one extra load instruction)
   const int ce = lane == 0 ? max(cc - 1, 0) : min(cc + 1, a.W - 1);
   const bool edge_lane = lane == 0 || lane == 31;
   const int r0 = blockIdx.y * NMS_ROWS;
   const int r1 = min(r0 + NMS_ROWS, a.H);
   const size_t ioff = (size_t)blockIdx.z * a.img_stride;
   const size_t moff = (size_t)blockIdx.z * a.mask_stride + wcol;
   const bool col_ok = c >= a.border && c < a.W - a.border;

   float hmax[NP][2], hmin[NP][2], vcur[NL][2];
#pragma unroll
   for (int p = 0; p < NP; p++) { hmax[p][0] = hmax[p][1] = 0.f; hmin[p][0] = hmin[p][1] = 0.f; }
#pragma unroll
   for (int l = 0; l < NL; l++) vcur[l][0] = vcur[l][1] = 0.f;

   // software pipeline: the loads of row rr+1 are in flight while row rr is reduced
   float nv[NP], ne[NP];
   {
      const int rl = min(max(r0 - 1, 0), a.H - 1);
#pragma unroll
      for (int p = 0; p < NP; p++) {
         const float *row = a.plane[p] + ioff + (size_t)rl * a.pitch;
         nv[p] = __ldg(row + cc);
         ne[p] = edge_lane ? __ldg(row + ce) : 0.f;
      }
   }
   for (int rr = r0 - 1; rr <= r1; rr++) {
      float v[NP], e[NP];
#pragma unroll
      for (int p = 0; p < NP; p++) { v[p] = nv[p]; e[p] = ne[p]; }
      if (rr < r1) {
         const int rl = min(max(rr + 1, 0), a.H - 1);
#pragma unroll
         for (int p = 0; p < NP; p++) {
            const float *row = a.plane[p] + ioff + (size_t)rl * a.pitch;
            nv[p] = __ldg(row + cc);
            ne[p] = edge_lane ? __ldg(row + ce) : 0.f;
         }
      }
      float nmax[NP], nmin[NP];
#pragma unroll
      for (int p = 0; p < NP; p++) {
         float l = __shfl_up_sync(0xffffffffu, v[p], 1);
         float r = __shfl_down_sync(0xffffffffu, v[p], 1);
         if (lane == 0) l = e[p];
         if (lane == 31) r = e[p];
         nmax[p] = fmaxf(fmaxf(l, r), v[p]);
         nmin[p] = fminf(fminf(l, r), v[p]);
      }
      if (rr >= r0 + 1) {
         const int ro = rr - 1;   // output row: rows ro-1 (age 0), ro (age 1), ro+1 (new) are available
         float pmax[NP], pmin[NP];   // 3x3 max / min of every plane around (ro, c)
#pragma unroll
         for (int p = 0; p < NP; p++) {
            pmax[p] = fmaxf(fmaxf(hmax[p][0], hmax[p][1]), nmax[p]);
            pmin[p] = fminf(fminf(hmin[p][0], hmin[p][1]), nmin[p]);
         }
         const bool pos_ok = col_ok && ro >= a.border && ro < a.H - a.border;
#pragma unroll
         for (int l = 0; l < NL; l++) {
            const float M = fmaxf(fmaxf(pmax[l], pmax[l + 1]), pmax[l + 2]);
            const float m = fminf(fminf(pmin[l], pmin[l + 1]), pmin[l + 2]);
            const float val = vcur[l][1];
            // findLevelKeypoints (pyramid.cpp:215-217): val > positiveThreshold and no neighbour above it (ties pass),
            // or val < negativeThreshold and no neighbour below it
            const bool cand = pos_ok && ((val > a.posThr && !(M > val)) || (val < a.negThr && !(m < val)));
            const unsigned word = __ballot_sync(0xffffffffu, cand);
            if (lane == 0) a.mask[l][moff + (size_t)ro * a.wpr] = word;
         }
      }
#pragma unroll
      for (int p = 0; p < NP; p++) { hmax[p][0] = hmax[p][1]; hmax[p][1] = nmax[p]; hmin[p][0] = hmin[p][1]; hmin[p][1] = nmin[p]; }
#pragma unroll
      for (int l = 0; l < NL; l++) { vcur[l][0] = vcur[l][1]; vcur[l][1] = v[l + 1]; }
   }
}

void ha_launch_nms(const float *arena, const Geom &g, const Geom *, uint32_t *mask, int n, cudaStream_t st, LaunchCounter &lc)
{
   for (int o = 0; o < g.nOct; o++)
      for (int l0 = 1; l0 <= g.S; l0 += NMS_MAXL) {
         const int nl = std::min(NMS_MAXL, g.S - l0 + 1);
         NmsArgs a;
         for (int p = 0; p < nl + 2; p++) a.plane[p] = arena + g.R_off[o][l0 - 1 + p];
         for (int l = 0; l < nl; l++) a.mask[l] = mask + g.mask_off[o][l0 + l];
         a.img_stride = g.arena_stride;
         a.W = g.w[o]; a.H = g.h[o]; a.pitch = g.pitch[o]; a.border = g.border; a.wpr = g.wpr[o];
         a.posThr = g.positiveThreshold; a.negThr = g.negativeThreshold;
         a.mask_stride = g.mask_stride;
         dim3 grid((g.wpr[o] + NMS_WARPS - 1) / NMS_WARPS, (g.h[o] + NMS_ROWS - 1) / NMS_ROWS, n);
         if (nl == 3) k_nms<3><<<grid, NMS_WARPS * 32, 0, st>>>(a);
         else if (nl == 2) k_nms<2><<<grid, NMS_WARPS * 32, 0, st>>>(a);
         else k_nms<1><<<grid, NMS_WARPS * 32, 0, st>>>(a);
         lc.n++;
      }
}

// =================================================================================================
// exclusive prefix sums (deterministic three-kernel scan; 2048 items per block)
// =================================================================================================
#define SCAN_T 256
#define SCAN_I 8
#define SCAN_B (SCAN_T * SCAN_I)

struct PopcIn {
   const uint32_t *w;
   size_t n;
   __device__ uint32_t operator()(size_t i) const { return i < n ? (uint32_t)__popc(w[i]) : 0u; }
};
struct FlagIn {
   const unsigned char *f;
   unsigned char bit;
   const uint32_t *count;
   size_t cap;
   __device__ uint32_t operator()(size_t i) const
   {
      const size_t n = min((size_t)*count, cap);
      return (i < n && (f[i] & bit)) ? 1u : 0u;
   }
};

struct U32In {
   const uint32_t *v;
   size_t n;
   __device__ uint32_t operator()(size_t i) const { return i < n ? v[i] : 0u; }
};

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t &total)
{
   __shared__ uint32_t wsum[SCAN_T / 32];
   const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
   uint32_t inc = v;
#pragma unroll
   for (int d = 1; d < 32; d <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += t;
   }
   if (lane == 31) wsum[wid] = inc;
   __syncthreads();
   uint32_t base = 0, tot = 0;
#pragma unroll
   for (int i = 0; i < SCAN_T / 32; i++) {
      if (i < wid) base += wsum[i];
      tot += wsum[i];
   }
   total = tot;
   __syncthreads();
   return base + inc - v;
}

template <typename In>
__global__ void __launch_bounds__(SCAN_T) k_scan_sums(In in, uint32_t *bsum)
{
   const size_t base = (size_t)blockIdx.x * SCAN_B + (size_t)threadIdx.x * SCAN_I;
   uint32_t s = 0;
#pragma unroll
   for (int i = 0; i < SCAN_I; i++) s += in(base + i);
   uint32_t tot;
   block_exclusive_scan(s, tot);
   if (threadIdx.x == 0) bsum[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(SCAN_T) k_scan_mid(uint32_t *bsum, size_t nb, uint32_t *total_out)
{
   __shared__ uint32_t carry_s;
   if (threadIdx.x == 0) carry_s = 0;
   __syncthreads();
   for (size_t b0 = 0; b0 < nb; b0 += SCAN_T) {
      const size_t i = b0 + threadIdx.x;
      const uint32_t v = i < nb ? bsum[i] : 0u;
      uint32_t tot;
      const uint32_t ex = block_exclusive_scan(v, tot);
      const uint32_t carry = carry_s;
      if (i < nb) bsum[i] = carry + ex;
      __syncthreads();
      if (threadIdx.x == 0) carry_s = carry + tot;
      __syncthreads();
   }
   if (threadIdx.x == 0) *total_out = carry_s;
}

template <typename In>
__global__ void __launch_bounds__(SCAN_T) k_scan_final(In in, const uint32_t *bsum, uint32_t *out, size_t n)
{
   const size_t base = (size_t)blockIdx.x * SCAN_B + (size_t)threadIdx.x * SCAN_I;
   uint32_t v[SCAN_I], s = 0;
#pragma unroll
   for (int i = 0; i < SCAN_I; i++) { v[i] = in(base + i); s += v[i]; }
   uint32_t tot;
   uint32_t run = block_exclusive_scan(s, tot) + bsum[blockIdx.x];
#pragma unroll
   for (int i = 0; i < SCAN_I; i++) {
      if (base + i < n) out[base + i] = run;
      run += v[i];
   }
}

size_t ha_scan_tmp_elems(size_t n) { return (n + SCAN_B - 1) / SCAN_B + 1; }

template <typename In>
static void scan_generic(In in, size_t n, uint32_t *out, uint32_t *tmp, cudaStream_t st, LaunchCounter &lc)
{
   const size_t nb = (n + SCAN_B - 1) / SCAN_B;
   if (nb == 0) { cudaMemsetAsync(out, 0, sizeof(uint32_t), st); return; }
   k_scan_sums<In><<<(unsigned)nb, SCAN_T, 0, st>>>(in, tmp);
   k_scan_mid<<<1, SCAN_T, 0, st>>>(tmp, nb, out + n);
   k_scan_final<In><<<(unsigned)nb, SCAN_T, 0, st>>>(in, tmp, out, n);
   lc.n += 3;
}

void ha_launch_scan_popc(const uint32_t *words, size_t nwords, uint32_t *out, uint32_t *tmp, cudaStream_t st, LaunchCounter &lc)
{
   PopcIn in{words, nwords};
   scan_generic(in, nwords, out, tmp, st, lc);
}

void ha_launch_scan_u32(const uint32_t *vals, size_t n, uint32_t *out, uint32_t *tmp, cudaStream_t st, LaunchCounter &lc)
{
   U32In in{vals, n};
   scan_generic(in, n, out, tmp, st, lc);
}

void ha_launch_scan_flags(const unsigned char *flags, unsigned char bit, const uint32_t *count_ptr, size_t cap, uint32_t *out,
                          uint32_t *tmp, cudaStream_t st, LaunchCounter &lc)
{
   FlagIn in{flags, bit, count_ptr, cap};
   scan_generic(in, cap, out, tmp, st, lc);
}

// =================================================================================================
// K2b: bitmask -> ordered candidate keys.  Word order == (image, octave, level, row, col) == the order
// in which the reference visits extrema, so candidate index order is reference order.
// =================================================================================================
__global__ void __launch_bounds__(256) k_expand(const uint32_t *__restrict__ mask, const uint32_t *__restrict__ woff,
                                                 const Geom *__restrict__ g, size_t nwords, Cand cand, uint32_t cap,
                                                 int *overflow)
{
   const size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (w >= nwords) return;
   uint32_t m = mask[w];
   if (!m) return;
   const int img = (int)(w / g->mask_stride);
   const unsigned long long rem = w - (unsigned long long)img * g->mask_stride;
   int o = 0;
   while (o + 1 < g->nOct && rem >= g->mask_oct_off[o + 1]) o++;
   const unsigned long long ro = rem - g->mask_oct_off[o];
   const unsigned long long per_level = (unsigned long long)g->h[o] * g->wpr[o];
   const int lvl = 1 + (int)(ro / per_level);
   const unsigned long long rl = ro - (unsigned long long)(lvl - 1) * per_level;
   const int r = (int)(rl / g->wpr[o]);
   const int wc = (int)(rl - (unsigned long long)r * g->wpr[o]);
   uint32_t off = woff[w];
   while (m) {
      const int b = __ffs(m) - 1;
      m &= m - 1;
      if (off < cap) cand.key[off] = ha_key(img, o, lvl, r, wc * 32 + b);
      else *overflow = 1;
      off++;
   }
}

void ha_launch_expand(const uint32_t *mask, const uint32_t *woff, const Geom *dg, size_t nwords, Cand cand, uint32_t cap,
                      int *overflow, cudaStream_t st, LaunchCounter &lc)
{
   if (!nwords) return;
   k_expand<<<(unsigned)((nwords + 255) / 256), 256, 0, st>>>(mask, woff, dg, nwords, cand, cap, overflow);
   lc.n++;
}

// =================================================================================================
// K2c: localizeKeypoint (pyramid.cpp:122-204) -- one thread per candidate.
// =================================================================================================
__device__ __forceinline__ void swapf(float &a, float &b) { const float t = a; a = b; b = t; }

// solveLinear3x3, helpers.cpp:46-88 (value swaps, no singularity check)
__device__ __forceinline__ void solve_linear_3x3(float *A, float *b)
{
   int i = 0, pr = 0;
   float vp = fabsf(A[0]);
   const float tmp = fabsf(A[3]);
   if (tmp > vp) { pr = 3; i = 1; vp = tmp; }
   if (fabsf(A[6]) > vp) { pr = 6; i = 2; }
   if (pr != 0) { swapf(A[pr], A[0]); swapf(A[pr + 1], A[1]); swapf(A[pr + 2], A[2]); swapf(b[i], b[0]); }
   vp = A[3] / A[0]; A[4] -= vp * A[1]; A[5] -= vp * A[2]; b[1] -= vp * b[0];
   vp = A[6] / A[0]; A[7] -= vp * A[1]; A[8] -= vp * A[2]; b[2] -= vp * b[0];
   if (fabsf(A[4]) < fabsf(A[7])) { swapf(A[7], A[4]); swapf(A[8], A[5]); swapf(b[2], b[1]); }
   vp = A[7] / A[4];
   A[8] -= vp * A[5];
   b[2] -= vp * b[1];
   b[2] = (b[2]) / A[8];
   b[1] = (b[1] - A[5] * b[2]) / A[4];
   b[0] = (b[0] - A[2] * b[2] - A[1] * b[1]) / A[0];
}

__global__ void __launch_bounds__(128) k_localize(const float *__restrict__ arena, const Geom *__restrict__ g, Cand cand,
                                                   const uint32_t *__restrict__ count, uint32_t cap, uint32_t *map)
{
   const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
   const uint32_t n = min(*count, cap);
   if (i >= n) return;
   int img, o, lvl, r, c;
   ha_unkey(cand.key[i], img, o, lvl, r, c);
   const int cols = g->w[o], rows = g->h[o], pitch = g->pitch[o];
   const float *base = arena + (size_t)img * g->arena_stride;
   const float *low = base + g->R_off[o][lvl - 1];
   const float *cur = base + g->R_off[o][lvl];
   const float *high = base + g->R_off[o][lvl + 1];
#define AT(p, rr, cc) ((p)[(size_t)(rr) * pitch + (cc)])
   float b[3] = {0.f, 0.f, 0.f};
   float val = 0.f;
   int nr = r, nc = c;
   bool alive = true;
   unsigned char flags = 0;
   for (int iter = 0; iter < 5 && alive; iter++) {
      r = nr; c = nc;
      const float dxx = AT(cur, r, c - 1) - 2.0f * AT(cur, r, c) + AT(cur, r, c + 1);
      const float dyy = AT(cur, r - 1, c) - 2.0f * AT(cur, r, c) + AT(cur, r + 1, c);
      const float dss = AT(low, r, c) - 2.0f * AT(cur, r, c) + AT(high, r, c);
      const float dxy = 0.25f * (AT(cur, r + 1, c + 1) - AT(cur, r + 1, c - 1) - AT(cur, r - 1, c + 1) + AT(cur, r - 1, c - 1));
      if (iter == 0) {
         const float edgeScore = (dxx + dyy) * (dxx + dyy) / (dxx * dyy - dxy * dxy);
         if (edgeScore >= g->edgeScoreThreshold || edgeScore < 0) { alive = false; break; }
      }
      const float dxs = 0.25f * (AT(high, r, c + 1) - AT(high, r, c - 1) - AT(low, r, c + 1) + AT(low, r, c - 1));
      const float dys = 0.25f * (AT(high, r + 1, c) - AT(high, r - 1, c) - AT(low, r + 1, c) + AT(low, r - 1, c));
      float A[9];
      A[0] = dxx; A[1] = dxy; A[2] = dxs;
      A[3] = dxy; A[4] = dyy; A[5] = dys;
      A[6] = dxs; A[7] = dys; A[8] = dss;
      const float dx = 0.5f * (AT(cur, r, c + 1) - AT(cur, r, c - 1));
      const float dy = 0.5f * (AT(cur, r + 1, c) - AT(cur, r - 1, c));
      const float ds = 0.5f * (AT(high, r, c) - AT(low, r, c));
      b[0] = -dx; b[1] = -dy; b[2] = -ds;
      solve_linear_3x3(A, b);
      if (isnan(b[0]) || isnan(b[1]) || isnan(b[2])) { alive = false; break; }
      val = AT(cur, r, c) + 0.5f * (dx * b[0] + dy * b[1] + ds * b[2]);
      // MAX_SUBPIXEL_SHIFT is the double literal 0.6; POINT_SAFETY_BORDER 3 (pyramid.cpp:117-120,174-177)
      if ((double)b[0] > 0.6) { if (c < cols - 3) nc++; else { alive = false; break; } }
      if ((double)b[1] > 0.6) { if (r < rows - 3) nr++; else { alive = false; break; } }
      if ((double)b[0] < -0.6) { if (c > 3) nc--; else { alive = false; break; } }
      if ((double)b[1] < -0.6) { if (r > 3) nr--; else { alive = false; break; } }
      if (nr == r && nc == c) break;
   }
   if (alive && !(fabsf(b[0]) > 1.5f || fabsf(b[1]) > 1.5f || fabsf(b[2]) > 1.5f || fabsf(val) < g->finalThreshold)) {
      // scale = curScale * pow(2.0f, b[2]/numberOfScales) (pyramid.cpp:196); 2^t evaluated in double and
      // rounded once, which agrees with glibc's (double-internal) powf
      const float t = b[2] / g->S;
      const float scale = g->sigma[lvl] * (float)exp2((double)t);
      const float pd = (float)(1 << o);
      // getHessianPointType on blur = L[lvl] at the final (r,c)  (pyramid.cpp:199, 24-37)
      int type;
      if (val < 0) type = 2;
      else {
         const float *bl = base + g->L_off[o][lvl] + (size_t)r * pitch + c;
         const float Lxx = (bl[-1] - 2 * bl[0] + bl[1]);
         type = Lxx < 0 ? 0 : 1;
      }
      cand.x[i] = pd * (c + b[0]);
      cand.y[i] = pd * (r + b[1]);
      cand.s[i] = pd * scale;
      cand.response[i] = val;
      cand.type[i] = (unsigned char)type;
      const int cell = r * cols + c;
      cand.cell[i] = cell;
      flags = HA_F_PASS;
      // octaveMap: the first candidate in reference order to reach a cell wins it
      atomicMin(map + (size_t)img * g->map_stride + g->map_off[o] + cell, i);
   }
   cand.flags[i] = flags;
#undef AT
}

void ha_launch_localize(const float *arena, const Geom *dg, Cand cand, const uint32_t *count, uint32_t cap, uint32_t *map,
                        cudaStream_t st, LaunchCounter &lc)
{
   k_localize<<<(cap + 127) / 128, 128, 0, st>>>(arena, dg, cand, count, cap, map);
   lc.n++;
}
