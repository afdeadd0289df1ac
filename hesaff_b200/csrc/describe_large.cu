// hesaff_b200/csrc/describe_large.cu -- affine patch normalisation + SIFT for large source patches (P > 95), one CTA
// per keypoint, several CTAs per SM.
//
// Replaces (reference file:line): AffineShape::normalizeAffine affine.cpp:102-144 (its > 0.4 branch), interpolate
// helpers.cpp:209-244, the per-patch gaussianBlurInplace helpers.cpp:291-295, and the SIFT tail of describe.cuh.
//
// The final 41x41 resampling of normalizeAffine is axis aligned (affine.cpp:131: interpolate(smoothed, c, c, its, 0, 0,
// its)), so it reads at most 82 distinct columns and 82 distinct rows of the blurred P x P patch.  Exactly those are
// computed: the row pass of the blur at the <= 82 needed columns of every row, the column pass at the <= 82 needed rows
// (same operation order per output as cv::GaussianBlur, so the patch is bit-identical to the reference's).
//
//   sampling : bands of source rows, two patch rows per work item (one column-table entry, eight loads in flight);
//              the band is stored as float2 = (row 2r, row 2r+1), replicate-padded left and right;
//   row pass : one item = two adjacent needed columns of a row pair on packed f32x2 math (FFMA2): 2 shared-memory
//              loads and 2 packed FMAs per tap for 4 outputs; result T[(R + P + R)][82] goes to an L2-resident
//              scratch plane of the CTA;
//   col pass : one item = two adjacent needed rows x two adjacent columns, packed; -> B[82][82] in shared memory;
//   tail     : 41x41 bilinear resampling of B, SIFT.
#include <algorithm>
#include <stdlib.h>
#include "common.cuh"
#include "describe.cuh"

#define LG_NT 256
#ifndef LG_MINB
#define LG_MINB 3
#endif
#define LG_B (82 * 82)
#define LG_TS 84             // row stride of the row-filtered scratch plane T (floats; rows 16-byte aligned)

struct LargeHead {
   float red[2 * (LG_NT / 32)];
   float rs_f[44];
   int rs_i[44];
   int work;
   int pad[3];
};

// one bilinear sample of interpolate() (helpers.cpp:235-236) from the u8 / float source
template <bool U8>
__device__ __forceinline__ float lg_sample(const void *__restrict__ im, int spitch, int off, float fx, float fx1, float wy)
{
   float p00, p01, p10, p11;
   if (U8) {
      const unsigned char *p = reinterpret_cast<const unsigned char *>(im) + off;
      p00 = (float)__ldg(p); p01 = (float)__ldg(p + 1); p10 = (float)__ldg(p + spitch); p11 = (float)__ldg(p + spitch + 1);
   } else {
      const float *p = reinterpret_cast<const float *>(im) + off;
      p00 = __ldg(p); p01 = __ldg(p + 1); p10 = __ldg(p + spitch); p11 = __ldg(p + spitch + 1);
   }
   return (1.0f - wy) * (fx1 * p00 + fx * p01) + (wy) * (fx1 * p10 + fx * p11);
}

template <bool U8>
__global__ void __launch_bounds__(LG_NT, LG_MINB) k_describe_large(const float *__restrict__ arena, const Geom *__restrict__ g, Tables tb,
                                                            Cand cand, const int *__restrict__ list, const int *__restrict__ list_n,
                                                            int *work_counter, float *scratch, size_t scratch_per_cta,
                                                            float *patch_dump, int dump_normalized,
                                                            const uint32_t *__restrict__ dump_index, int rowbuf_floats, int no_stage)
{
   constexpr int NT = LG_NT;
   extern __shared__ __align__(16) unsigned char dsm[];
   LargeHead &sh = *reinterpret_cast<LargeHead *>(dsm);
   float2 *kh2 = reinterpret_cast<float2 *>(dsm + ((sizeof(LargeHead) + 15) & ~(size_t)15));   // (k, k) for k[R..n-1]
   float *B = reinterpret_cast<float *>(kh2 + HA_MAX_PATCH_R + 1);                              // [82][82]; first the column table
   float *rowbuf = B + LG_B + 4;                                                                // band of row pairs; later patch / acc
   float4 *ctab = reinterpret_cast<float4 *>(B);
   float2 *v01 = reinterpret_cast<float2 *>(B);
   float *patch = rowbuf;
   float2 *acc = reinterpret_cast<float2 *>(rowbuf);
   const int tid = threadIdx.x;
   const int nwork = *list_n;
   const int cols = g->W, rows = g->H;
   const int spitch = U8 ? g->pitch8 : g->pitch[0];
   float *__restrict__ T = scratch + (size_t)blockIdx.x * scratch_per_cta;

   for (;;) {
      __syncthreads();
      if (tid == 0) sh.work = atomicAdd(work_counter, 1);
      __syncthreads();
      const int wi = sh.work;
      if (wi >= nwork) break;
      const int i = list[wi];
      const int img = (int)(cand.key[i] >> 48);
      const float x = cand.x[i], y = cand.y[i], s = cand.s[i];
      const float4 A = cand.A[i];
      const float a11 = A.x, a12 = A.y, a21 = A.z, a22 = A.w;
      const void *__restrict__ im = U8 ? (const void *)(reinterpret_cast<const unsigned char *>(arena + (size_t)img * g->arena_stride + g->img8_off))
                                       : (const void *)(arena + (size_t)img * g->arena_stride + g->img_off);
      // normalizeAffine, affine.cpp:102-144 (this bin: its > 0.4 always)
      const float mrScale = ceilf(s * g->mrSize);
      const int P0 = 2 * (int)(mrScale) + 1;
      const float its = (float)P0 / (float)HA_PATCH;
      const int P = P0 + 2, half = P >> 1;
      // interpolate() reports "touches boundary" if any of the P*P samples is outside; positions are monotone in i and
      // j, so the four corners decide
      if (!ha_sample_inside(cols, rows, x, y, a11, a12, a21, a22, -half, -half) ||
          !ha_sample_inside(cols, rows, x, y, a11, a12, a21, a22, half, -half) ||
          !ha_sample_inside(cols, rows, x, y, a11, a12, a21, a22, -half, half) ||
          !ha_sample_inside(cols, rows, x, y, a11, a12, a21, a22, half, half))
         continue;      // uniform across the CTA
      const int m = (P0 - 1) >> 1;
      const int n = tb.pk_n[m], R = n >> 1;
      const float *__restrict__ kg = tb.pk + tb.pk_off[m];
      for (int t = tid; t <= R; t += NT) { const float k = kg[t]; kh2[t] = make_float2(k, k); }
      // resampling table of interpolate(smoothed, P>>1, P>>1, its, 0, 0, its, patch): position c0 + k*its, k = -20..20
      const float c0f = (float)half;
      for (int t = tid; t < HA_PATCH; t += NT) {
         const float w = c0f + (t - (HA_PATCH >> 1)) * its;
         const int wi2 = (int)floorf(w);
         sh.rs_i[t] = wi2;
         sh.rs_f[t] = w - wi2;
      }
      // per patch column: source column, horizontal weights, skew term i*a21 (a12 = 0.f exactly, helpers.cpp:90-97, so
      // wx = (x + j*0.f) + i*a11 = x + i*a11 depends on the column only); kept in B, which is idle until the column pass
      const bool use_ctab = 4 * P <= LG_B && a12 == 0.f;
      if (use_ctab)
         for (int t = tid; t < P; t += NT) {
            const int ii = t - half;
            const float wx = x + (float)ii * a11, fl = floorf(wx), fx = wx - fl;
            ctab[t] = make_float4(__int_as_float((int)fl), fx, 1.0f - fx, (float)ii * a21);
         }
      const int RS2 = (P + 2 * R + 2) | 1;                            // float2 per padded row pair: R + P + R (+1 read past the last tap) + 1, odd
      // row pairs per band: 16 if they fit, else a power of two (the row pass maps lanes to row pairs)
      const int G2fit = max(1, min(16, rowbuf_floats / (2 * RS2)));
      const int lgLR = min(4, 31 - __clz(G2fit)), LR = 1 << lgLR, LJ = 32 >> lgLR;   // LR <= 16: two lanes share a 16-byte store
      const int G2 = LR;
      float2 *rb2 = reinterpret_cast<float2 *>(rowbuf);
      // source box of the sampling: the part of B behind the column table (B is idle until the column pass)
      unsigned char *box = reinterpret_cast<unsigned char *>(B + ((4 * P + 3) & ~3));
      const int box_cap = (use_ctab && !no_stage) ? (LG_B - ((4 * P + 3) & ~3)) * 4 : 0;
      __syncthreads();
      for (int r0 = 0; r0 < P; r0 += 2 * G2) {
         const int nrp = min(G2, (P - r0 + 1) >> 1);
         const int nrb = (nrp + LR - 1) >> lgLR, njb = (HA_PATCH + LJ - 1) / LJ;
         // ---- sampling: item = (row pair, column).  u8 source: the band is cut into column chunks whose source bounding
         // box fits the idle part of B; the box is copied with coalesced 16-byte loads and the samples come from shared
         // memory (the direct form gathers four bytes per sample from up to 32 different cache lines per warp load: 78 % of
         // the kernel's L1 requests and its largest stall) -----------------------------------------------------------------
         const int jlo = r0, jhi = min(r0 + 2 * nrp - 1, P - 1);
         int CW = P;                                       // columns per chunk
         if (U8 && use_ctab) {
            for (;;) {
               // widest / tallest box of a chunk of CW columns (positions are monotone in i and j)
               const float wxspan = (float)(CW - 1) * a11;
               const int bwmax = (((int)wxspan + 3 + 15 + 15) & ~15);
               const float hy = (float)(jhi - jlo) * a22 + (float)(CW - 1) * fabsf(a21);
               const int brmax = (int)hy + 4;
               if (bwmax * brmax <= box_cap || CW <= 32) break;
               CW = (CW + 1) >> 1;
            }
         }
         for (int c0 = 0; c0 < P; c0 += CW) {
            const int c1 = min(c0 + CW, P) - 1;
            bool staged = false;
            int bx0 = 0, by0 = 0, bw = 0;
            if (U8 && use_ctab) {
               const float xl = x + (float)(c0 - half) * a11, xr = x + (float)(c1 - half) * a11;
               const float ya = y + (float)(jlo - half) * a22, yb = y + (float)(jhi - half) * a22;
               const float sl = (float)(c0 - half) * a21, sr = (float)(c1 - half) * a21;
               const float y00 = ya + sl, y01 = ya + sr, y10 = yb + sl, y11 = yb + sr;
               const int xmin = (int)floorf(fminf(xl, xr)), xmax = (int)floorf(fmaxf(xl, xr)) + 1;
               by0 = (int)floorf(fminf(fminf(y00, y01), fminf(y10, y11)));
               const int ymax = (int)floorf(fmaxf(fmaxf(y00, y01), fmaxf(y10, y11))) + 1;
               bx0 = xmin & ~15;
               bw = ((xmax - bx0 + 1) + 15) & ~15;
               const int brows = ymax - by0 + 1;
               staged = bw * brows <= box_cap;
               if (staged) {
                  const int wq = bw >> 4;
                  const float invwq = 1.0f / (float)wq;
                  const unsigned char *src = reinterpret_cast<const unsigned char *>(im) + (size_t)by0 * spitch + bx0;
                  for (int t = tid; t < brows * wq; t += NT) {
                     const int r = ha_fast_div(t, invwq), q = t - r * wq;
                     *reinterpret_cast<uint4 *>(box + r * bw + 16 * q) = __ldg(reinterpret_cast<const uint4 *>(src + (size_t)r * spitch + 16 * q));
                  }
               }
               __syncthreads();
            }
            const int ncol = c1 - c0 + 1;
            const float invC = 1.0f / (float)ncol;
            if (staged) {
               const int off0 = -(by0 * bw + bx0);
               for (int t = tid; t < nrp * ncol; t += NT) {
                  const int rp = ha_fast_div(t, invC), xx = c0 + (t - rp * ncol);
                  const int ja = r0 + 2 * rp, jb = min(ja + 1, P - 1);
                  const float4 c = ctab[xx];
                  float wya = (y + (float)(ja - half) * a22) + c.w, wyb = (y + (float)(jb - half) * a22) + c.w;
                  const float fya = floorf(wya), fyb = floorf(wyb);
                  wya -= fya; wyb -= fyb;
                  const int xi = __float_as_int(c.x);
                  const unsigned char *pa = box + (off0 + (int)fya * bw + xi), *pb = box + (off0 + (int)fyb * bw + xi);
                  const float va = (1.0f - wya) * (c.z * (float)pa[0] + c.y * (float)pa[1]) + (wya) * (c.z * (float)pa[bw] + c.y * (float)pa[bw + 1]);
                  const float vb = (1.0f - wyb) * (c.z * (float)pb[0] + c.y * (float)pb[1]) + (wyb) * (c.z * (float)pb[bw] + c.y * (float)pb[bw + 1]);
                  rb2[rp * RS2 + R + xx] = make_float2(va, vb);
               }
            } else {
               for (int t = tid; t < nrp * ncol; t += NT) {
                  const int rp = ha_fast_div(t, invC), xx = c0 + (t - rp * ncol);
                  const int ja = r0 + 2 * rp, jb = min(ja + 1, P - 1);
                  float4 c;
                  float rxa = 0.f, rxb = 0.f;
                  if (use_ctab) c = ctab[xx];
                  else {
                     // general A (never produced by k_affine; kept for completeness): per-sample column terms
                     const int ii = xx - half;
                     rxa = x + (float)(ja - half) * a12; rxb = x + (float)(jb - half) * a12;
                     const float wx = rxa + (float)ii * a11, fl = floorf(wx), fx = wx - fl;
                     c = make_float4(__int_as_float((int)fl), fx, 1.0f - fx, (float)ii * a21);
                  }
                  float wya = (y + (float)(ja - half) * a22) + c.w, wyb = (y + (float)(jb - half) * a22) + c.w;
                  const float fya = floorf(wya), fyb = floorf(wyb);
                  wya -= fya; wyb -= fyb;
                  float va, vb;
                  if (use_ctab || rxa == rxb) {
                     const int xi = __float_as_int(c.x);
                     va = lg_sample<U8>(im, spitch, (int)fya * spitch + xi, c.y, c.z, wya);
                     vb = lg_sample<U8>(im, spitch, (int)fyb * spitch + xi, c.y, c.z, wyb);
                  } else {
                     const int ii = xx - half;
                     const float wxb = rxb + (float)ii * a11, flb = floorf(wxb), fxb = wxb - flb;
                     va = lg_sample<U8>(im, spitch, (int)fya * spitch + __float_as_int(c.x), c.y, c.z, wya);
                     vb = lg_sample<U8>(im, spitch, (int)fyb * spitch + (int)flb, fxb, 1.0f - fxb, wyb);
                  }
                  rb2[rp * RS2 + R + xx] = make_float2(va, vb);
               }
            }
            if (c1 + 1 < P) __syncthreads();     // the next chunk's box overwrites this one
         }
         __syncthreads();
         for (int t = tid; t < nrp * (2 * R + 1); t += NT) {   // replicate R columns left, R + 1 right (BORDER_REPLICATE)
            const int rp = t / (2 * R + 1), q = t - rp * (2 * R + 1);
            float2 *row = rb2 + rp * RS2;
            if (q < R) row[q] = row[R];
            else row[P + q] = row[R + P - 1];                   // columns R+P .. R+P+R
         }
         __syncthreads();
         // ---- row pass: item = (row pair, pair of adjacent needed columns x0, x0+1); the tap windows overlap in all but
         // one sample; chain order of every output is the reference's (left to right).  Lanes run over ROW PAIRS (LR of
         // them, RS2 odd: every 8-byte access of a half-warp in its own bank pair) and LJ = 32 / LR column pairs; the tap
         // is the same for every lane (one broadcast read). ---------------------------------------------------------------
         for (int u = tid >> 5; u < nrb * njb; u += NT / 32) {
            const int ub = u / nrb;
            const int rp_ = (u - ub * nrb) * LR + (tid & (LR - 1)), jx_ = ub * LJ + ((tid & 31) >> lgLR);
            const bool live = rp_ < nrp && jx_ < HA_PATCH;
            const int rp = min(rp_, nrp - 1), jx = min(jx_, HA_PATCH - 1);
            const ha_f2 *p = reinterpret_cast<const ha_f2 *>(rb2 + rp * RS2 + sh.rs_i[jx]);   // tap 0 of column x0 = padded column x0
            const ha_f2 *kc = reinterpret_cast<const ha_f2 *>(kh2) + R;                       // k(i) = kh[|i - R|]
            ha_f2 e = p[1];
            ha_f2 c = *kc;
            ha_f2 a = ha_f2_mul(p[0], c), b = ha_f2_mul(e, c);
            p += 2;
            int q = R;
#pragma unroll 4
            for (; q > 0; q--) {                                // rising half: kh[R - i]
               c = *--kc;
               a = ha_f2_fma(e, c, a);
               e = *p++;
               b = ha_f2_fma(e, c, b);
            }
#pragma unroll 4
            for (q = R; q > 0; q--) {                           // falling half: kh[i - R]
               c = *++kc;
               a = ha_f2_fma(e, c, a);
               e = *p++;
               b = ha_f2_fma(e, c, b);
            }
            // fa = (row 2rp, row 2rp+1) of column x0, fb of column x0+1.  The lane LR further on holds the next column pair
            // of the same rows: even column pairs take over row 2rp, odd ones row 2rp+1, and store 16 bytes each
            const float2 fa = ha_f2_unpack(a), fb = ha_f2_unpack(b);
            const bool oddj = (jx_ & 1) != 0;
            const float rx = __shfl_xor_sync(0xffffffffu, oddj ? fa.x : fa.y, LR);
            const float ry = __shfl_xor_sync(0xffffffffu, oddj ? fb.x : fb.y, LR);
            if (!live) continue;
            const int ja = r0 + 2 * rp;
            if (jx_ == HA_PATCH - 1) {             // the last column pair has no partner
               float *d = T + (size_t)(R + ja) * LG_TS + 2 * jx;
               const float2 oa = make_float2(fa.x, fb.x);
               *reinterpret_cast<float2 *>(d) = oa;
               if (ja + 1 < P) *reinterpret_cast<float2 *>(d + LG_TS) = make_float2(fa.y, fb.y);
               // BORDER_REPLICATE of the column pass: R copies of the first / last filtered row
               if (ja == 0) for (int k = 1; k <= R; k++) *reinterpret_cast<float2 *>(d - (size_t)k * LG_TS) = oa;
               if (ja == P - 1) for (int k = 1; k <= R; k++) *reinterpret_cast<float2 *>(d + (size_t)k * LG_TS) = oa;
            } else if (!oddj) {
               float *d = T + (size_t)(R + ja) * LG_TS + 2 * jx;
               const float4 o = make_float4(fa.x, fb.x, rx, ry);
               *reinterpret_cast<float4 *>(d) = o;
               if (ja == 0) for (int k = 1; k <= R; k++) *reinterpret_cast<float4 *>(d - (size_t)k * LG_TS) = o;
               if (ja == P - 1) for (int k = 1; k <= R; k++) *reinterpret_cast<float4 *>(d + (size_t)k * LG_TS) = o;
            } else if (ja + 1 < P) {
               *reinterpret_cast<float4 *>(T + (size_t)(R + ja + 1) * LG_TS + 2 * (jx - 1)) = make_float4(rx, ry, fa.y, fb.y);
            }
         }
         __syncthreads();
      }
      // ---- column pass: item = (pair of adjacent needed rows yy, yy+1) x (pair of adjacent columns); every loaded sample
      // serves both rows.  centre*k0, then (above + below) FMA'd outwards, as the reference. -----------------------------
      for (int t = tid; t < HA_PATCH * HA_PATCH; t += NT) {
         const int jy = t / HA_PATCH, q2 = t - jy * HA_PATCH;
         const ha_f2 *base = reinterpret_cast<const ha_f2 *>(T + (size_t)(R + sh.rs_i[jy]) * LG_TS + 2 * q2);
         const ha_f2 *kc = reinterpret_cast<const ha_f2 *>(kh2);
         ha_f2 am = base[0], bm = base[LG_TS / 2];                        // T[yy - (k-1)], T[yy + 1 + (k-1)]
         ha_f2 acc0 = ha_f2_mul(am, kc[0]), acc1 = ha_f2_mul(bm, kc[0]);
         const ha_f2 *up = base, *dn = base + LG_TS / 2;
#pragma unroll 4
         for (int k = 1; k <= R; k++) {
            up -= LG_TS / 2; dn += LG_TS / 2;
            const ha_f2 ak = *up, bk = *dn, w = kc[k];
            acc0 = ha_f2_fma(ha_f2_add(ak, bm), w, acc0);         // row yy  : T[yy-k] + T[yy+k]
            acc1 = ha_f2_fma(ha_f2_add(am, bk), w, acc1);         // row yy+1: T[yy+1-k] + T[yy+1+k]
            am = ak; bm = bk;
         }
         *reinterpret_cast<ha_f2 *>(B + (2 * jy) * 82 + 2 * q2) = acc0;
         *reinterpret_cast<ha_f2 *>(B + (2 * jy + 1) * 82 + 2 * q2) = acc1;
      }
      __syncthreads();
      {
         float *dump = patch_dump ? patch_dump + (size_t)dump_index[i] * HA_PATCH_PX : nullptr;
         // interpolate(smoothed, P>>1, P>>1, its, 0, 0, its, patch) (affine.cpp:131) from the 82 x 82 blurred grid
         ha_sift_describe<NT>([&](int jj, int ii) {
            const float *p = B + (2 * jj) * 82 + 2 * ii;
            return ha_bilinear(p[0], p[1], p[82], p[83], sh.rs_f[ii], sh.rs_f[jj]);
         }, sh.red, patch, v01, acc, tb, cand.desc + (size_t)i * 128, dump_normalized ? nullptr : dump, dump_normalized ? dump : nullptr);
      }
      if (tid == 0) cand.flags[i] |= HA_F_DESC;
   }
}

// One padded row pair = R + P + R + 2 float2, where R = taps/2 of the per-patch blur (sigma = 1.5*P0/41,
// helpers.cpp:293).  The band buffer holds at least one row pair of the widest possible patch and 40 KB otherwise.
static int large_row_stride(int maxP)
{
   const float sigma = 1.5f * ((float)maxP / (float)HA_PATCH);
   int n = (int)(2.0 * 3.0 * sigma + 1.0);
   if (n % 2 == 0) n++;
   return (maxP + 2 * (n / 2) + 2) | 1;
}
// 40 KB: 16 row pairs per band up to P ~ 260, and 76 KB per CTA = two CTAs per SM beside three TINY CTAs of the main stream
// (measured on the describe stage: 24 KB / 39 KB at three CTAs per SM +4 % / +2.5 %, 48 KB +4 %, 56 KB +8 %, a 128-register
// build +8 %, a layout with the 82 x 82 grid aliased onto a 50 KB band buffer +2.5 %)
#ifndef LG_ROWBUF
#define LG_ROWBUF 10240
#endif
static int large_rowbuf_floats(int maxP) { return std::max(LG_ROWBUF, 2 * large_row_stride(maxP) + 8); }
// rows of the row-filtered scratch plane T per CTA: R + P + R
size_t ha_describe_scratch_floats(int maxP) { return (size_t)(large_row_stride(maxP) + 2) * LG_TS; }

static int large_smem_bytes(int maxP)
{
   return (int)(((sizeof(LargeHead) + 15) & ~(size_t)15) + sizeof(float2) * (HA_MAX_PATCH_R + 1) +
                sizeof(float) * (LG_B + 4 + (size_t)large_rowbuf_floats(maxP)));
}

int ha_describe_large_max_ctas_per_sm(int maxP) { return std::max(1, std::min(LG_MINB, 227 * 1024 / (large_smem_bytes(maxP) + 1024))); }

void ha_launch_describe_large(const float *arena, const Geom *dg, Tables tb, Cand cand, Bins bins, int *work_counters,
                              float *scratch, size_t scratch_per_cta, int ctas_per_sm, int maxP, int src_u8, float *patch_dump,
                              int dump_normalized, const uint32_t *dump_index, cudaStream_t st)
{
   const int smem = large_smem_bytes(maxP);
   if (src_u8) {
      cudaFuncSetAttribute(k_describe_large<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);   // grows with maxP
      k_describe_large<true><<<148 * ctas_per_sm, LG_NT, smem, st>>>(arena, dg, tb, cand, bins.list[2], bins.count + 2, work_counters + 2, scratch,
                                                                   scratch_per_cta, patch_dump, dump_normalized, dump_index,
                                                                   large_rowbuf_floats(maxP), ha_no_stage());
   } else {
      cudaFuncSetAttribute(k_describe_large<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      k_describe_large<false><<<148 * ctas_per_sm, LG_NT, smem, st>>>(arena, dg, tb, cand, bins.list[2], bins.count + 2, work_counters + 2, scratch,
                                                                    scratch_per_cta, patch_dump, dump_normalized, dump_index,
                                                                    large_rowbuf_floats(maxP), ha_no_stage());
   }
}
