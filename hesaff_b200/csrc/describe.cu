// hesaff_b200/csrc/describe.cu -- affine patch normalisation + SIFT for the source-patch bins whose P x P patch
// lives in shared memory (TINY .. MEDIUM, P <= 95), one CTA per keypoint.
//
// Replaces (reference file:line):
//   AffineShape::normalizeAffine                                 affine.cpp:102-144
//   interpolate                                                  helpers.cpp:209-244
//   per-patch gaussianBlurInplace -> cv::GaussianBlur            helpers.cpp:291-295
//   SIFTDescriptor::computeSiftDescriptor, samplePatch, sample   siftdesc.cpp:51-140
//   photometricallyNormalize                                     helpers.cpp:246-281
//
// Per keypoint (CTA of NT threads):
//   0. the work item (x, y, s, A, blur taps) was prefetched into shared memory while the previous keypoint was
//      processed (3-deep pipeline: queue ticket -> list entry -> parameters), so the loop head waits for nothing;
//   1. u8 source: the axis-aligned bounding box of the affine footprint is copied from the 16-byte aligned u8 image
//      into shared memory with coalesced 16-byte loads (one round trip to L2 instead of four scattered loads per
//      sample); the P x P samples are then taken from shared memory.  A gray 8-bit image is integer valued, so the
//      u8 copy is exact.  fp32 / colour input samples the float image directly;
//   2. separable per-patch blur, register tiled, the replicated borders written by the producing pass;
//   3. axis-aligned resampling to 41 x 41;
//   4. SIFT on the 1245-pixel mask disc; the gradient pass leaves, per pixel, the two orientation-bin contributions
//      and the accumulator slot, so the histogram is one 8-byte read-modify-write per (pixel, cell).
#include <algorithm>
#include <stdlib.h>
#include "common.cuh"
#include "describe.cuh"

// =================================================================================================
// per-patch blur in shared memory, register tiled
// =================================================================================================
// Row strides are multiples of 4 floats, so the row pass moves float4s: PS = roundup4(P + 2R + 3) for S, PT =
// roundup4(P) for T.
// S : P rows, stride PS; S[y*PS + R + x] = sample (y, x); the R columns to the left and R+3 to the right hold the
//     replicated edge value (BORDER_REPLICATE; written by the sampling pass), so the taps need no clamping.
// T : P + 2R + 3 rows of PT; T[(R + y)*PT + x] = row-filtered value; the rows above / below replicate the edge rows
//     (written here by the threads that produce the edge rows).
// out: the blurred patch, stride P, written over S.
template <int N, int NIN>
__device__ __forceinline__ void ha_patch_row_taps(const float (&in)[NIN], const float (&k)[N], float (&out)[4])
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

template <int N, int NT>
__device__ void ha_patch_blur(float *__restrict__ S, float *__restrict__ T, int P, const float *__restrict__ kh)
{
   constexpr int R = N / 2;
   const int PS = (P + 2 * R + 3 + 3) & ~3, PT = (P + 3) & ~3;
   const int tid = threadIdx.x;
   float k[N];
#pragma unroll
   for (int i = 0; i < N; i++) k[i] = kh[i < R ? R - i : i - R];
   const int G = (P + 3) >> 2;
   const uint32_t MG = ha_div_magic(G), MP = ha_div_magic(P);
   // row pass, 4 outputs per thread from (N + 3 + 3) / 4 float4 loads; outputs past column P-1 land in T's padding.
   // The first / last row is also written to the R rows above / the R+3 rows below (BORDER_REPLICATE of the column pass)
   for (int t = tid; t < P * G; t += NT) {
      const int y = ha_div22(t, MG), x0 = (t - y * G) << 2;
      const float4 *p = reinterpret_cast<const float4 *>(S + y * PS + x0);
      constexpr int NQ = (N + 3 + 3) / 4;
      float in[4 * NQ];
#pragma unroll
      for (int i = 0; i < NQ; i++) {
         const float4 q = p[i];
         in[4 * i] = q.x; in[4 * i + 1] = q.y; in[4 * i + 2] = q.z; in[4 * i + 3] = q.w;
      }
      float o[4];
      ha_patch_row_taps<N>(in, k, o);
      const float4 o4 = make_float4(o[0], o[1], o[2], o[3]);
      float *d = T + (R + y) * PT + x0;
      *reinterpret_cast<float4 *>(d) = o4;
      if (y == 0) {
#pragma unroll
         for (int q = 1; q <= R; q++) *reinterpret_cast<float4 *>(d - q * PT) = o4;
      }
      if (y == P - 1) {
#pragma unroll
         for (int q = 1; q <= R + 3; q++) *reinterpret_cast<float4 *>(d + q * PT) = o4;
      }
   }
   __syncthreads();
   // column pass, 4 outputs per thread: centre*k[R], then (above+below) FMA'd outwards
   for (int t = tid; t < G * P; t += NT) {
      const int gy = ha_div22(t, MP), x = t - gy * P, y0 = gy << 2;
      float m[N + 3];
#pragma unroll
      for (int i = 0; i < N + 3; i++) m[i] = T[(y0 + i) * PT + x];
#pragma unroll
      for (int j = 0; j < 4; j++) {
         float acc = m[j + R] * k[R];
#pragma unroll
         for (int i = 1; i <= R; i++) acc = __fmaf_rn(m[j + R - i] + m[j + R + i], k[R + i], acc);
         if (y0 + j < P) S[(y0 + j) * P + x] = acc;
      }
   }
   __syncthreads();
}

// Row pass of the per-patch blur at position x of a replicate-padded row (row[-R..P-1+R] valid), generic n >= 7
__device__ __forceinline__ float ha_padded_row_blur(const float *__restrict__ row, int x, int n, int R,
                                                    const float *__restrict__ kh /* k[R..n-1] */)
{
   const float *p = row + x - R;
   float acc = p[0] * kh[R];
   int i = 1;
   for (; i <= R; i++) acc = __fmaf_rn(p[i], kh[R - i], acc);
   for (; i < n; i++) acc = __fmaf_rn(p[i], kh[i - R], acc);
   return acc;
}

// generic (any n) fallback with the same buffers (parameter sets whose taps no instantiation covers)
template <int NT>
__device__ void ha_patch_blur_generic(float *__restrict__ S, float *__restrict__ T, int P, int n, const float *__restrict__ kh)
{
   const int R = n >> 1, PS = (P + 2 * R + 3 + 3) & ~3, tid = threadIdx.x;
   const float invP = 1.0f / (float)P;
   for (int t = tid; t < P * P; t += NT) {
      const int y = ha_fast_div(t, invP), x = t - y * P;
      const float *row = S + y * PS + R;
      float v;
      if (n == 5) {
         float acc = (row[x - 1] + row[x + 1]) * kh[1];
         acc = __fmaf_rn(row[x], kh[0], acc);
         v = __fmaf_rn(row[x - 2] + row[x + 2], kh[2], acc);
      } else if (n == 3) v = __fmaf_rn(row[x], kh[0], (row[x - 1] + row[x + 1]) * kh[1]);
      else if (n == 1) v = row[x] * kh[0];
      else v = ha_padded_row_blur(row, x, n, R, kh);
      T[(R + y) * P + x] = v;
   }
   __syncthreads();
   for (int t = tid; t < P * P; t += NT) {
      const int y = ha_fast_div(t, invP), x = t - y * P;
      float acc = T[(R + y) * P + x] * kh[0];
      for (int q = 1; q <= R; q++) {
         const int ya = max(y - q, 0), yb = min(y + q, P - 1);
         acc = __fmaf_rn(T[(R + ya) * P + x] + T[(R + yb) * P + x], kh[q], acc);
      }
      S[t] = acc;   // S's padded content is dead after the row pass (barrier above); T is only read here
   }
   __syncthreads();
}

// =================================================================================================
// the kernel
// =================================================================================================
// One work item, as the prefetching thread leaves it in shared memory
struct DescItem {
   float x, y, s, a11, a21, a22;
   int i;            // candidate index, < 0: the queue is empty
   int img;
};

// shared-memory plan of a bin (floats unless stated), see ha_describe_smem_bytes
template <int BIN> struct DescPlan {
   static constexpr int MAXP = BIN == 3 ? HA_BIN_TINY_MAXP : (BIN == 0 ? HA_BIN_SMALL_MAXP : (BIN == 4 ? HA_BIN_MID_MAXP : (BIN == 5 ? HA_BIN_MID2_MAXP : HA_BIN_MEDIUM_MAXP)));
   static constexpr int MAXR = BIN == 3 ? 4 : (BIN == 0 ? 5 : (BIN == 4 ? 7 : (BIN == 5 ? 8 : 10)));   // taps / 2 at MAXP
   static constexpr int PS = (MAXP + 2 * MAXR + 3 + 3) & ~3, PT = (MAXP + 3) & ~3;
   static constexpr int SA0 = MAXP * PS;                                  // S
   static constexpr int SB0 = (MAXP + 2 * MAXR + 3) * PT;                 // T
   static constexpr int V01 = 2 * HA_PATCH_PX + 2;                        // float2 per patch pixel
   static constexpr int SA = ((SA0 > V01 ? SA0 : V01) + 3) & ~3;          // region A: S, blurred patch, then v01
   static constexpr int ACC = 2 * 8 * 128;                                // float2 [8][128]
   static constexpr int SB1 = SB0 > ACC ? SB0 : ACC;
   static constexpr int SB = ((SB1 > HA_PATCH_PX + 3 ? SB1 : HA_PATCH_PX + 3) + 3) & ~3;   // region B: box, T, patch, then acc
   static constexpr int TABF = 4 * ((MAXP + 3) & ~3) + ((MAXP + 3) & ~3) + 3 * 44;   // ctab (float4), rtab, rs_f, rs_i, rs_r
   static constexpr int SC = (TABF + 3) & ~3;                             // region C: sampling tables
   static constexpr int KN = 16;                                          // taps k[R..n-1] (R <= 10), [15] = n
};

template <int NT> struct DescHead {
   DescItem par[2];
   float kern[2][16];
   float red[2 * (NT / 32)];
   uint32_t enext;
   int pad[3];
};

// list entry of the shared-memory bins: candidate index | m << 26, m = (P0 - 1) / 2 <= 47 (k_affine packs it, so that
// the blur taps can be fetched together with the keypoint's parameters)
#define HA_LIST_IDX(e) ((int)((e) & 0x3ffffffu))
#define HA_LIST_M(e) ((int)((uint32_t)(e) >> 26))
#define HA_LIST_NONE 0xffffffffu

// resident CTAs per SM the register budget is set for (shared memory allows as many)
#ifndef DESC_MINB_TINY
#define DESC_MINB_TINY 8
#endif
#define DESC_MINB(BIN) ((BIN) == 3 ? DESC_MINB_TINY : ((BIN) == 0 ? 7 : ((BIN) == 4 ? 4 : ((BIN) == 5 ? 3 : 2))))

// what the kernel needs of the geometry, passed by value: kernel parameters live in the constant bank and are used as
// instruction operands, so they cost no registers over the keypoint loop (read from the Geom in memory they cost eleven)
struct DescGeom {
   int cols, rows, pitch, pitch8;
   float mrSize;
   unsigned long long arena_stride, img_off, img8_off;
};

template <int BIN, int NT, bool U8>
__global__ void __launch_bounds__(NT, DESC_MINB(BIN)) k_describe(const float *__restrict__ arena, const DescGeom g, Tables tb,
                                                 Cand cand, const int *__restrict__ list, const int *__restrict__ list_n,
                                                 int *work_counter, float *patch_dump, int dump_normalized,
                                                 const uint32_t *__restrict__ dump_index, int no_stage)
{
   typedef DescPlan<BIN> PL;
   extern __shared__ __align__(16) unsigned char dsm[];
   DescHead<NT> &sh = *reinterpret_cast<DescHead<NT> *>(dsm);
   float *regA = reinterpret_cast<float *>(dsm + ((sizeof(DescHead<NT>) + 15) & ~(size_t)15));
   float *regB = regA + PL::SA;
   float *regC = regB + PL::SB;
   // region C while sampling
   constexpr int CP = (PL::MAXP + 3) & ~3;
   float4 *ctab = reinterpret_cast<float4 *>(regC);          // per patch column: source column (int bits), fx, 1 - fx, i*a21
   float *rtab = regC + 4 * CP;                              // per patch row: y + j*a22
   float *rs_f = rtab + CP;                                  // resampling table: fractional part per output index
   int *rs_i = reinterpret_cast<int *>(rs_f + 44);           //                   integer part
   int *rs_r = rs_i + 44;                                    //                   integer part times the row stride
   // SIFT
   float2 *v01 = reinterpret_cast<float2 *>(regA);
   float2 *acc = reinterpret_cast<float2 *>(regB);
   float *patch = regB;

   const int tid = threadIdx.x;
   const int nwork = *list_n;
   const int cols = g.cols, rows = g.rows, pitch = g.pitch, pitch8 = g.pitch8;
   const float mrSize = g.mrSize;
   const unsigned long long arena_stride = g.arena_stride, img_off = g.img_off, img8_off = g.img8_off;

   // prefetch pipeline state (thread 0): ticket of item k+2, list entry of item k+1 (HA_LIST_NONE: none)
   int pf_w = -1;
   uint32_t pf_e = HA_LIST_NONE;
   for (int k = -3;; k++) {
      __syncthreads();
      const int cur = k & 1;
      DescItem it;
      it.i = -1;
      if (k >= 0) {
         it = sh.par[cur];
         if (it.i < 0) break;
      }
      // ---- issue the prefetches of the next items; their results are used at the end of this iteration -------------
      int r_w = -1;
      uint32_t r_e = HA_LIST_NONE;
      DescItem r_it;
      r_it.i = -1;
      if (tid == 0) {
         r_w = atomicAdd(work_counter, 1);                                  // ticket of item k+3
         if (pf_w >= 0 && pf_w < nwork) r_e = (uint32_t)__ldg(list + pf_w);   // list entry of item k+2
         if (pf_e != HA_LIST_NONE) {                                        // parameters of item k+1
            const int i1 = HA_LIST_IDX(pf_e);
            const float4 A = cand.A[i1];
            r_it.x = cand.x[i1]; r_it.y = cand.y[i1]; r_it.s = cand.s[i1];
            r_it.a11 = A.x; r_it.a21 = A.z; r_it.a22 = A.w;
            r_it.img = (int)(cand.key[i1] >> 48);
            r_it.i = i1;
         }
      }
      float r_tap = 0.f;
      bool committed = false;
      {
         const uint32_t en = k >= -1 ? sh.enext : HA_LIST_NONE;              // written at the end of iteration k-1
         if (en != HA_LIST_NONE && tid < 16) r_tap = __ldg(tb.pk16 + HA_LIST_M(en) * 16 + tid);
      }

      if (k >= 0) {
         const int i = it.i;
         const float x = it.x, y = it.y, s = it.s;
         const float a11 = it.a11, a21 = it.a21, a22 = it.a22;
         const float a12 = 0.f;     // rectifyAffineTransformationUpIsUp (helpers.cpp:90-97) makes a12 exactly 0
         const float *__restrict__ kern = sh.kern[cur];
         // normalizeAffine, affine.cpp:102-144
         const float mrScale = ceilf(s * mrSize);
         const int P0 = 2 * (int)(mrScale) + 1;
         const float its = (float)P0 / (float)HA_PATCH;
         const int P = P0 + 2, half = P >> 1;
         bool rejected = false;
         const bool oversampled = !((double)its > 0.4);
         if (!oversampled) {
            // interpolate() reports "touches boundary" if any of the P*P samples is outside; positions are
            // monotone in i and j, so the four corners decide
            if (!ha_sample_inside(cols, rows, x, y, a11, a12, a21, a22, -half, -half) ||
                !ha_sample_inside(cols, rows, x, y, a11, a12, a21, a22, half, -half) ||
                !ha_sample_inside(cols, rows, x, y, a11, a12, a21, a22, -half, half) ||
                !ha_sample_inside(cols, rows, x, y, a11, a12, a21, a22, half, half))
               rejected = true;
         }
#if defined(HA_ABL) && HA_ABL == 3
         if (!rejected) { if (tid == 0) cand.flags[i] |= HA_F_DESC; rejected = true; }
#endif
         if (!rejected && !oversampled) {
            const int n = (int)kern[15], R = n >> 1;
            const int PS = (P + 2 * R + 3 + 3) & ~3;
            float *S = regA, *T = regB;
            // ---- tables: resampling positions of interpolate(smoothed, P>>1, P>>1, its, 0, 0, its, patch): c0 + k*its
            // for k = -20..20, split into floor + fraction; per-column / per-row terms of the affine sampling ----------
            const float c0f = (float)half;
            for (int t = tid; t < HA_PATCH; t += NT) {
               const float w = c0f + (t - (HA_PATCH >> 1)) * its;
               const int wi2 = (int)floorf(w);
               rs_i[t] = wi2;
               rs_r[t] = wi2 * P;
               rs_f[t] = w - wi2;
            }
            // bounding box of the footprint in the source image (sample positions are monotone in i and j)
            int bx0 = 0, by0 = 0, bw = 0, brows = 0;
            bool staged = false;
            if (U8) {
               const float xl = x + (float)(-half) * a11, xr = x + (float)half * a11;       // rx = x + j*0.f = x
               const float y00 = (y + (float)(-half) * a22) + (float)(-half) * a21, y01 = (y + (float)(-half) * a22) + (float)half * a21;
               const float y10 = (y + (float)half * a22) + (float)(-half) * a21, y11 = (y + (float)half * a22) + (float)half * a21;
               const int xmin = (int)floorf(fminf(xl, xr)), xmax = (int)floorf(fmaxf(xl, xr)) + 1;
               by0 = (int)floorf(fminf(fminf(y00, y01), fminf(y10, y11)));
               const int ymax = (int)floorf(fmaxf(fmaxf(y00, y01), fmaxf(y10, y11))) + 1;
               bx0 = xmin & ~15;
               bw = ((xmax - bx0 + 1) + 15) & ~15;
               brows = ymax - by0 + 1;
               staged = bw * brows <= PL::SB * 4 && !no_stage;
            }
            for (int t = tid; t < P; t += NT) {
               const int ii = t - half;
               const float wx = x + (float)ii * a11;
               const float fl = floorf(wx);
               const float fx = wx - fl;
               ctab[t] = make_float4(__int_as_float((int)fl - (staged ? bx0 : 0)), fx, 1.0f - fx, (float)ii * a21);
               rtab[t] = y + (float)ii * a22;
            }
            const uint32_t MP = ha_div_magic(P);
            if (U8) {
               const unsigned char *__restrict__ im8 = reinterpret_cast<const unsigned char *>(arena + (size_t)it.img * arena_stride + img8_off);
               unsigned char *box = reinterpret_cast<unsigned char *>(regB);
               if (staged) {
                  const int wq = bw >> 4;
                  const uint32_t Mwq = ha_div_magic(wq);
                  const unsigned char *src = im8 + (size_t)by0 * pitch8 + bx0;
                  for (int t = tid; t < brows * wq; t += NT) {
                     const int r = ha_div22(t, Mwq), q = t - r * wq;
                     *reinterpret_cast<uint4 *>(box + r * bw + 16 * q) = __ldg(reinterpret_cast<const uint4 *>(src + (size_t)r * pitch8 + 16 * q));
                  }
               }
               __syncthreads();
               if (staged) {
                  const int off0 = -by0 * bw;      // (the column table already holds source column - bx0)
                  for (int t = tid; t < P * P; t += NT) {
                     const int jj = ha_div22(t, MP), xx = t - jj * P;
                     const float4 c = ctab[xx];
                     float wy = rtab[jj] + c.w;
                     const float fy = floorf(wy);
                     wy -= fy;
                     const unsigned char *p = box + (off0 + (int)fy * bw + __float_as_int(c.x));
                     const float v = (1.0f - wy) * (c.z * (float)p[0] + c.y * (float)p[1]) + (wy) * (c.z * (float)p[bw] + c.y * (float)p[bw + 1]);
                     float *d = S + jj * PS + R + xx;
                     *d = v;
                  }
               } else {
                  for (int t = tid; t < P * P; t += NT) {
                     const int jj = ha_div22(t, MP), xx = t - jj * P;
                     const float4 c = ctab[xx];
                     float wy = rtab[jj] + c.w;
                     const float fy = floorf(wy);
                     wy -= fy;
                     const unsigned char *p = im8 + ((int)fy * pitch8 + __float_as_int(c.x));
                     const float v = (1.0f - wy) * (c.z * (float)__ldg(p) + c.y * (float)__ldg(p + 1)) +
                                     (wy) * (c.z * (float)__ldg(p + pitch8) + c.y * (float)__ldg(p + pitch8 + 1));
                     float *d = S + jj * PS + R + xx;
                     *d = v;
                  }
               }
            } else {
               const float *__restrict__ im = arena + (size_t)it.img * arena_stride + img_off;
               __syncthreads();
               for (int t = tid; t < P * P; t += NT) {
                  const int jj = ha_div22(t, MP), xx = t - jj * P;
                  const float4 c = ctab[xx];
                  float wy = rtab[jj] + c.w;
                  const float fy = floorf(wy);
                  wy -= fy;
                  const float *p = im + ((int)fy * pitch + __float_as_int(c.x));
                  const float v = (1.0f - wy) * (c.z * __ldg(p) + c.y * __ldg(p + 1)) + (wy) * (c.z * __ldg(p + pitch) + c.y * __ldg(p + pitch + 1));
                  float *d = S + jj * PS + R + xx;
                  *d = v;
               }
            }
            __syncthreads();
            {  // replicate the edge columns: R to the left, R+3 to the right (BORDER_REPLICATE of the row pass)
               const int W2 = 2 * R + 3;
               const uint32_t MW2 = ha_div_magic(W2);
               for (int t = tid; t < P * W2; t += NT) {
                  const int yy = ha_div22(t, MW2), q = t - yy * W2;
                  float *row = S + yy * PS;
                  if (q < R) row[q] = row[R];
                  else row[P + q] = row[R + P - 1];        // columns R+P .. R+P+R+2
               }
            }
            __syncthreads();
#if defined(HA_ABL) && HA_ABL == 2
            if (tid == 0) cand.flags[i] |= HA_F_DESC;
            rejected = true;
#else
            // gaussianBlurInplace(smoothed, 1.5f*its): row pass then column pass, replicate border
            switch (n) {
               // taps n = odd(6*sigma + 1), sigma = 1.5*(P-2)/41: at most 9 in the TINY bin (P <= 39), 11 in SMALL (P <= 47),
               // 15 in MID (P <= 63), 17 in MID2 (P <= 79), 21 in MEDIUM (P <= 95); the instantiations a bin cannot reach would only cost it registers
#define HA_PB(N) case N: if (N <= 2 * PL::MAXR + 1) { ha_patch_blur<N, NT>(S, T, P, kern); break; }
               HA_PB(5) HA_PB(7) HA_PB(9) HA_PB(11) HA_PB(13) HA_PB(15) HA_PB(17) HA_PB(19) HA_PB(21)
#undef HA_PB
               default: ha_patch_blur_generic<NT>(S, T, P, n, kern);
            }
#endif
         }
         if (!rejected) {   // uniform across the CTA
            // The prefetched item has arrived by now (the sampling and the blur ran since its loads were issued): park it
            // in shared memory here, so that its registers are free during the SIFT stage.  Nobody reads the other
            // parameter slot, the other tap row or sh.enext before the top of the next iteration.  (The direct-sampling branch
            // has had no barrier since the top of this iteration, where every thread reads sh.enext.)
            if (oversampled) __syncthreads();
            if (tid == 0) {
               sh.par[cur ^ 1] = r_it;
               sh.enext = r_e;
               pf_e = r_e;
               pf_w = r_w;
            }
            if (tid < 16) sh.kern[cur ^ 1][tid] = r_tap;
            committed = true;
#if defined(HA_ABL) && HA_ABL == 1
            if (tid == 0) cand.flags[i] |= HA_F_DESC;
#else
            float *dump = patch_dump ? patch_dump + (size_t)dump_index[i] * HA_PATCH_PX : nullptr;
            float *dump_raw = dump_normalized ? nullptr : dump, *dump_norm = dump_normalized ? dump : nullptr;
            unsigned char *desc = cand.desc + (size_t)i * 128;
            if (!oversampled) {
               // interpolate(smoothed, P>>1, P>>1, its, 0, 0, its, patch) (affine.cpp:131): axis aligned, tables above
               const float *S = regA;
               ha_sift_describe<NT>([&](int jj, int ii) {
                  const float *p = S + rs_r[jj] + rs_i[ii];
                  return ha_bilinear(p[0], p[1], p[P], p[P + 1], rs_f[ii], rs_f[jj]);
               }, sh.red, patch, v01, acc, tb, desc, dump_raw, dump_norm);
            } else {
               // lots of oversampling: sample the 41x41 patch directly (affine.cpp:135-142); zeros outside the image
               const float *__restrict__ im = arena + (size_t)it.img * arena_stride + img_off;
               const float b11 = a11 * its, b12 = a12 * its, b21 = a21 * its, b22 = a22 * its;
               ha_sift_describe<NT>([&](int jj, int ii) {
                  const int j = jj - (HA_PATCH >> 1), ic = ii - (HA_PATCH >> 1);
                  const float rx = x + j * b12, ry = y + j * b22;
                  float wx = rx + ic * b11, wy = ry + ic * b21;
                  const int xi = (int)floorf(wx), yi = (int)floorf(wy);
                  float v = 0.f;
                  if (xi >= 0 && yi >= 0 && xi < cols - 1 && yi < rows - 1) {
                     wx -= xi; wy -= yi;
                     const float *p = im + (size_t)yi * pitch + xi;
                     v = ha_bilinear(p[0], p[1], p[pitch], p[pitch + 1], wx, wy);
                  }
                  return v;
               }, sh.red, patch, v01, acc, tb, desc, dump_raw, dump_norm);
            }
            if (tid == 0) cand.flags[i] |= HA_F_DESC;
#endif
         }
      }
      if (!committed) {         // prologue iterations and rejected keypoints (uniform)
         __syncthreads();      // every thread has read sh.enext / sh.par[cur] before they are overwritten below
         if (tid == 0) {
            sh.par[cur ^ 1] = r_it;
            sh.enext = r_e;
            pf_e = r_e;
            pf_w = r_w;
         }
         if (tid < 16) sh.kern[cur ^ 1][tid] = r_tap;
      }
   }
}

// dynamic shared memory of a bin's kernel
template <int BIN, int NT> static constexpr int desc_smem()
{
   return (int)(((sizeof(DescHead<NT>) + 15) & ~(size_t)15) + sizeof(float) * (DescPlan<BIN>::SA + DescPlan<BIN>::SB + DescPlan<BIN>::SC));
}

int ha_describe_smem_bytes(int bin)
{
   static_assert(DESC_MINB_TINY * (desc_smem<3, 128>() + 1024) <= 227 * 1024, "TINY: shared memory of DESC_MINB_TINY CTAs per SM");
   switch (bin) {
      case 3: return desc_smem<3, 128>();
      case 0: return desc_smem<0, 128>();
      case 4: return desc_smem<4, 256>();
      case 5: return desc_smem<5, 256>();
      case 1: return desc_smem<1, 256>();
   }
   return 0;
}

// HESAFF_NO_STAGE=1 (tests): sample the u8 source directly instead of from the shared-memory box -- the fallback every
// kernel takes when a footprint's bounding box does not fit; records must not change
int ha_no_stage()
{
   static const int v = getenv("HESAFF_NO_STAGE") && atoi(getenv("HESAFF_NO_STAGE")) > 0;
   return v;
}

struct DescLaunch {
   const float *arena; DescGeom dg; Tables tb; Cand cand; Bins bins; int *work;
   float *patch_dump; int dump_normalized; const uint32_t *dump_index; int src_u8;
};

template <int BIN, int NT>
static void launch_desc(const DescLaunch &a, int ctas_per_sm, cudaStream_t st)
{
   constexpr int smem = desc_smem<BIN, NT>();
   if (a.src_u8) {
      cudaFuncSetAttribute(k_describe<BIN, NT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      k_describe<BIN, NT, true><<<148 * ctas_per_sm, NT, smem, st>>>(a.arena, a.dg, a.tb, a.cand, a.bins.list[BIN], a.bins.count + BIN,
                                                                     a.work + BIN, a.patch_dump, a.dump_normalized, a.dump_index, ha_no_stage());
   } else {
      cudaFuncSetAttribute(k_describe<BIN, NT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      k_describe<BIN, NT, false><<<148 * ctas_per_sm, NT, smem, st>>>(a.arena, a.dg, a.tb, a.cand, a.bins.list[BIN], a.bins.count + BIN,
                                                                      a.work + BIN, a.patch_dump, a.dump_normalized, a.dump_index, ha_no_stage());
   }
}

// Launch plan of the describe stage: "<main stream>;<aux stream>", each a comma-separated list of <bin letter><CTAs per SM>
// with T = TINY, S = SMALL, D = MID, E = MID2, M = MEDIUM, L = LARGE.  Every launch of a bin pulls from that bin's work queue, so a kernel
// that starts late simply helps with what is left, and one that finds its queue empty exits at once.
static const char *describe_plan()
{
   static const char *e = getenv("HESAFF_PLAN");
   return e ? e : "T8,S5,D3,E2,M1;L2,M1";
}

void ha_launch_describe(const float *arena, const Geom *dg, const Geom &hg, Tables tb, Cand cand, Bins bins, int *work_counters,
                        float *scratch, size_t scratch_per_cta, int large_ctas, int maxP, int src_u8, float *patch_dump,
                        int dump_normalized, const uint32_t *dump_index, cudaStream_t st, LaunchCounter &lc,
                        cudaStream_t aux, cudaEvent_t ev_fork, cudaEvent_t ev_join)
{
   const DescGeom geo{hg.W, hg.H, hg.pitch[0], hg.pitch8, hg.mrSize, hg.arena_stride, hg.img_off, hg.img8_off};
   const DescLaunch a{arena, geo, tb, cand, bins, work_counters, patch_dump, dump_normalized, dump_index, src_u8};
   const int per_sm = 227 * 1024;
   int large_slot = 0;                     // LARGE launches running side by side need their own scratch planes
   unsigned seen = 0;                      // bins the plan has launched
   auto run = [&](const char *p, const char *end, cudaStream_t s) {
      while (p < end) {
         const char bin = *p++;
         int n = 0;
         while (p < end && *p >= '0' && *p <= '9') n = n * 10 + (*p++ - '0');
         if (p < end && *p == ',') p++;
         if (n <= 0) continue;
         if (bin == 'T') launch_desc<3, 128>(a, std::min(n, per_sm / (ha_describe_smem_bytes(3) + 1024)), s);
         else if (bin == 'S') launch_desc<0, 128>(a, std::min(n, per_sm / (ha_describe_smem_bytes(0) + 1024)), s);
         else if (bin == 'D') launch_desc<4, 256>(a, std::min(n, per_sm / (ha_describe_smem_bytes(4) + 1024)), s);
         else if (bin == 'E') launch_desc<5, 256>(a, std::min(n, per_sm / (ha_describe_smem_bytes(5) + 1024)), s);
         else if (bin == 'M') launch_desc<1, 256>(a, std::min(n, per_sm / (ha_describe_smem_bytes(1) + 1024)), s);
         else if (bin == 'L') {
            const int avail = std::min(large_ctas / 148, ha_describe_large_max_ctas_per_sm(maxP)) - large_slot;
            if (avail <= 0) continue;
            n = std::min(n, avail);
            ha_launch_describe_large(arena, dg, tb, cand, bins, work_counters, scratch + (size_t)large_slot * 148 * scratch_per_cta,
                                     scratch_per_cta, n, maxP, src_u8, patch_dump, dump_normalized, dump_index, s);
            large_slot += n;
         } else continue;
         seen |= 1u << (bin - 'A');
         lc.n++;
      }
   };
   const char *plan = describe_plan();
   const char *sep = plan;
   while (*sep && *sep != ';') sep++;
   const char *end = sep;
   while (*end) end++;
   if (aux != nullptr && *sep == ';') {
      cudaEventRecord(ev_fork, st);
      cudaStreamWaitEvent(aux, ev_fork, 0);
      run(sep + 1, end, aux);
      run(plan, sep, st);
      cudaEventRecord(ev_join, aux);
      cudaStreamWaitEvent(st, ev_join, 0);
   } else {
      // one stream: the auxiliary list first (the long bins), then the main list
      if (*sep == ';') run(sep + 1, end, st);
      run(plan, sep, st);
   }
   // a plan that leaves a bin out must not drop its keypoints
   for (const char *b = "TSDEML"; *b; b++)
      if (!(seen & (1u << (*b - 'A')))) {
         const char one[3] = {*b, '1', 0};
         run(one, one + 2, st);
      }
}
