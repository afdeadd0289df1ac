/* oracle/hesaff_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C, single-threaded CPU restatement of the reference's detect -> affine -> describe path
 * (perdoch/hesaff @ /root/reference), written from the behaviour of the reference, each function
 * citing the reference lines it follows.  It is the checker the CUDA path is compared with; only
 * tests/, __graft_entry__.smoke() and bench.py's CPU-baseline leg may load it.
 *
 * PARITY PINNED: tests/test_oracle_pin.py requires this file to agree BIT FOR BIT, stage by stage
 * and end to end, with oracle/_ref/libhesaff_ref.so (the reference's own sources compiled unmodified
 * with its Makefile flags against oracle/shim), and tests/golden/ holds outputs of that reference
 * build.  The one third-party piece, cv::GaussianBlur, is restated in gaussian_blur() below and is
 * pinned against real OpenCV 4.13 in tests/test_oracle_blur.py (see oracle/shim/cv_shim.cpp).
 *
 * Build: gcc -O3 -std=gnu99 -mfma -ffp-contract=off (FMAs only where written, as in the shim).
 * Floating-point expression order, float/double promotions and int truncations follow the reference
 * exactly; do not "simplify" them.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "oracle_api.h"

#define FMAF(a, b, c) __builtin_fmaf((a), (b), (c))

const char *orc_name(void) { return "port (oracle/hesaff_oracle.c)"; }

void orc_free(void *p) { free(p); }

void orc_default_params(orc_params *o)
{
   /* hesaff.cpp:28-35; pyramid.h:34-39; affine.h:39-44 */
   o->threshold = 16.0f / 3.0f;
   o->max_iter = 16;
   o->desc_factor = 3.0f * sqrtf(3.0f);
   o->patch_size = 41;
   o->number_of_scales = 3;
   o->initial_sigma = 1.6f;
   o->edge_eigenvalue_ratio = 10.0f;
   o->border = 5;
   o->convergence_threshold = 0.05f;
   o->smm_window_size = 19;
   o->max_octaves = 0;
}

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* ------------------------------------------------------------------------------------------------
 * cv::GaussianBlur as called from helpers.cpp:283-295 (ksize from sigma, sigmaX=sigmaY, BORDER_REPLICATE)
 * OpenCV 4.13 operation order, see oracle/shim/cv_shim.cpp.
 * ---------------------------------------------------------------------------------------------- */
static int blur_size(float sigma)
{
   /* helpers.cpp:286,293 */
   int size = (int)(2.0 * 3.0 * sigma + 1.0);
   if (size % 2 == 0) size++;
   return size;
}

static void gauss_kernel(int n, double sigma, float *k)
{
   const int R = (n - 1) / 2;
   const double scale2X = -0.5 / (sigma * sigma);
   double v[512], sum = 0;
   for (int i = 0; i < R; i++) { double x = (double)(i - R); v[i] = exp(scale2X * x * x); sum += v[i]; }
   v[R] = 1.0;
   sum = sum * 2 + 1.0;
   const double m = 1.0 / sum;
   for (int i = 0; i <= R; i++) k[i] = k[n - 1 - i] = (float)(v[i] * m);
}

static void gaussian_blur(const float *src, int h, int w, float sigma, float *dst)
{
   const int n = blur_size(sigma), R = n / 2;
   float k[1024];
   gauss_kernel(n, (double)sigma, k);
   float *mid = (float *)malloc(sizeof(float) * (size_t)h * w);
   float *p0 = (float *)malloc(sizeof(float) * (size_t)(w + 2 * R));
   float *p = p0 + R;
   for (int y = 0; y < h; y++) {
      const float *s = src + (size_t)y * w;
      float *d = mid + (size_t)y * w;
      for (int x = -R; x < w + R; x++) p[x] = s[clampi(x, 0, w - 1)];
      if (n == 1) {
         for (int x = 0; x < w; x++) d[x] = p[x] * k[0];
      } else if (n == 3) {
         for (int x = 0; x < w; x++) d[x] = FMAF(p[x], k[1], (p[x - 1] + p[x + 1]) * k[2]);
      } else if (n == 5) {
         for (int x = 0; x < w; x++) {
            float acc = (p[x - 1] + p[x + 1]) * k[3];
            acc = FMAF(p[x], k[2], acc);
            d[x] = FMAF(p[x - 2] + p[x + 2], k[4], acc);
         }
      } else {
         for (int x = 0; x < w; x++) {
            float acc = p[x - R] * k[0];
            for (int i = 1; i < n; i++) acc = FMAF(p[x - R + i], k[i], acc);
            d[x] = acc;
         }
      }
   }
   for (int y = 0; y < h; y++) {
      float *d = dst + (size_t)y * w;
      const float *c = mid + (size_t)y * w;
      for (int x = 0; x < w; x++) d[x] = c[x] * k[R];
      for (int i = 1; i <= R; i++) {
         const float *a = mid + (size_t)clampi(y - i, 0, h - 1) * w;
         const float *b = mid + (size_t)clampi(y + i, 0, h - 1) * w;
         const float ki = k[R + i];
         for (int x = 0; x < w; x++) d[x] = FMAF(a[x] + b[x], ki, d[x]);
      }
   }
   free(p0);
   free(mid);
}

void orc_gaussian_blur(const float *src, int h, int w, float sigma, float *dst)
{
   float *tmp = (float *)malloc(sizeof(float) * (size_t)h * w);
   gaussian_blur(src, h, w, sigma, tmp);
   memcpy(dst, tmp, sizeof(float) * (size_t)h * w);
   free(tmp);
}

/* ------------------------------------------------------------------------------------------------
 * HessianDetector::hessianResponse, pyramid.cpp:63-114
 * ---------------------------------------------------------------------------------------------- */
static void hessian_response(const float *in, int rows, int cols, float norm, float *out)
{
   const float norm2 = norm * norm;                                  /* :76 */
   memset(out, 0, sizeof(float) * (size_t)rows * cols);              /* border: never read, see header */
   for (int r = 1; r < rows - 1; r++) {
      const float *a = in + (size_t)(r - 1) * cols, *b = in + (size_t)r * cols, *c = in + (size_t)(r + 1) * cols;
      float *o = out + (size_t)r * cols;
      for (int x = 1; x < cols - 1; x++) {
         const float v11 = a[x - 1], v12 = a[x], v13 = a[x + 1];
         const float v21 = b[x - 1], v22 = b[x], v23 = b[x + 1];
         const float v31 = c[x - 1], v32 = c[x], v33 = c[x + 1];
         const float Lxx = (v21 - 2 * v22 + v23);                     /* :96 */
         const float Lyy = (v12 - 2 * v22 + v32);                     /* :97 */
         const float Lxy = (v13 - v11 + v31 - v33) / 4.0f;            /* :98 */
         o[x] = (Lxx * Lyy - Lxy * Lxy) * norm2;                      /* :101 */
      }
   }
}

void orc_hessian_response(const float *src, int h, int w, float norm, float *dst) { hessian_response(src, h, w, norm, dst); }

/* ------------------------------------------------------------------------------------------------
 * helpers.cpp numeric helpers
 * ---------------------------------------------------------------------------------------------- */
static void swapf(float *a, float *b) { float t = *a; *a = *b; *b = t; }   /* helpers.cpp:40-44 (value swap) */

/* helpers.cpp:46-88 */
static void solve_linear_3x3(float *A, float *b)
{
   int i = 0;
   float *pr = A;
   float vp = fabsf(A[0]);
   float tmp = fabsf(A[3]);
   if (tmp > vp) { pr = A + 3; i = 1; vp = tmp; }
   if (fabsf(A[6]) > vp) { pr = A + 6; i = 2; }
   if (pr != A) { swapf(pr, A); swapf(pr + 1, A + 1); swapf(pr + 2, A + 2); swapf(b + i, b); }
   vp = A[3] / A[0]; A[4] -= vp * A[1]; A[5] -= vp * A[2]; b[1] -= vp * b[0];
   vp = A[6] / A[0]; A[7] -= vp * A[1]; A[8] -= vp * A[2]; b[2] -= vp * b[0];
   if (fabsf(A[4]) < fabsf(A[7])) { swapf(A + 7, A + 4); swapf(A + 8, A + 5); swapf(b + 2, b + 1); }
   vp = A[7] / A[4];
   A[8] -= vp * A[5];
   b[2] -= vp * b[1];
   b[2] = (b[2]) / A[8];
   b[1] = (b[1] - A[5] * b[2]) / A[4];
   b[0] = (b[0] - A[2] * b[2] - A[1] * b[1]) / A[0];
}

/* helpers.cpp:90-97 */
static void rectify_up_is_up(float *a11, float *a12, float *a21, float *a22)
{
   double a = *a11, b = *a12, c = *a21, d = *a22;
   double det = sqrt(fabs(a * d - b * c));
   double b2a2 = sqrt(b * b + a * a);
   *a11 = (float)(b2a2 / det);
   *a12 = 0;
   *a21 = (float)((d * b + c * a) / (b2a2 * det));
   *a22 = (float)(det / b2a2);
}
void orc_rectify(float *A) { rectify_up_is_up(A, A + 1, A + 2, A + 3); }

/* helpers.cpp:104-129 */
static void compute_gauss_mask(float *mask, int size)
{
   int halfSize = size >> 1;
   float scale = (float)halfSize / 3.0f;
   float scale2 = -2.0f * scale * scale;
   float *tmp = (float *)malloc(sizeof(float) * (halfSize + 1));
   for (int i = 0; i <= halfSize; i++) tmp[i] = expf(((float)(i * i) / scale2));
   int endSize = (int)(ceilf(scale * 5.0f) - halfSize);
   for (int i = 1; i < endSize; i++) tmp[halfSize - i] += expf(((float)((i + halfSize) * (i + halfSize)) / scale2));
   for (int i = 0; i <= halfSize; i++)
      for (int j = 0; j <= halfSize; j++) {
         float v = tmp[i] * tmp[j];
         mask[(i + halfSize) * size + (-j + halfSize)] = v;
         mask[(-i + halfSize) * size + (j + halfSize)] = v;
         mask[(i + halfSize) * size + (j + halfSize)] = v;
         mask[(-i + halfSize) * size + (-j + halfSize)] = v;
      }
   free(tmp);
}

/* helpers.cpp:131-147 */
static void compute_circular_gauss_mask(float *mask, int size)
{
   int halfSize = size >> 1;
   float r2 = (float)(halfSize * halfSize);
   float sigma2 = 0.9f * r2;
   float *mp = mask;
   for (int i = 0; i < size; i++)
      for (int j = 0; j < size; j++) {
         float disq = (float)((i - halfSize) * (i - halfSize) + (j - halfSize) * (j - halfSize));
         *mp++ = (disq < r2) ? expf(-disq / sigma2) : 0;
      }
}

/* helpers.cpp:149-175 */
static void inv_sqrt(float *a, float *b, float *c, float *l1, float *l2)
{
   double t, r;
   if (*b != 0) {
      r = (double)(*c - *a) / (2 * *b);
      if (r >= 0) t = 1.0 / (r + sqrt(1 + r * r)); else t = -1.0 / (-r + sqrt(1 + r * r));
      r = 1.0 / sqrt(1 + t * t);
      t = t * r;
   } else {
      r = 1;
      t = 0;
   }
   double x, z, d;
   x = 1.0 / sqrt(r * r * *a - 2 * r * t * *b + t * t * *c);
   z = 1.0 / sqrt(t * t * *a + 2 * r * t * *b + r * r * *c);
   d = sqrt(x * z);
   x /= d; z /= d;
   if (x < z) { *l1 = (float)z; *l2 = (float)x; } else { *l1 = (float)x; *l2 = (float)z; }
   *a = (float)(r * r * x + t * t * z);
   *b = (float)(-r * t * x + t * r * z);
   *c = (float)(t * t * x + r * r * z);
}

/* helpers.cpp:177-188 */
static int get_eigenvalues(float a, float b, float c, float d, float *l1, float *l2)
{
   float trace = a + d;
   float delta1 = (trace * trace - 4 * (a * d - b * c));
   if (delta1 < 0) return 0;
   float delta = sqrtf(delta1);
   *l1 = (trace + delta) / 2.0f;
   *l2 = (trace - delta) / 2.0f;
   return 1;
}

/* helpers.cpp:191-207 */
static int interpolate_check_borders(int imrows, int imcols, float ofsx, float ofsy, float a11, float a12, float a21,
                                     float a22, int rescols, int resrows)
{
   const int width = imcols - 2;
   const int height = imrows - 2;
   const int halfWidth = rescols >> 1;
   const int halfHeight = resrows >> 1;
   float x[4]; x[0] = -halfWidth; x[1] = -halfWidth; x[2] = +halfWidth; x[3] = +halfWidth;
   float y[4]; y[0] = -halfHeight; y[1] = +halfHeight; y[2] = -halfHeight; y[3] = +halfHeight;
   for (int i = 0; i < 4; i++) {
      float imx = ofsx + x[i] * a11 + y[i] * a12;
      float imy = ofsy + x[i] * a21 + y[i] * a22;
      if (floorf(imx) <= 0 || floorf(imy) <= 0 || ceilf(imx) >= width || ceilf(imy) >= height) return 1;
   }
   return 0;
}

/* helpers.cpp:209-244 */
static int interpolate(const float *im, int imrows, int imcols, float ofsx, float ofsy, float a11, float a12, float a21,
                       float a22, float *res, int resrows, int rescols)
{
   int ret = 0;
   const int width = imcols - 1;
   const int height = imrows - 1;
   const int halfWidth = rescols >> 1;
   const int halfHeight = resrows >> 1;
   float *out = res;
   for (int j = -halfHeight; j <= halfHeight; ++j) {
      const float rx = ofsx + j * a12;
      const float ry = ofsy + j * a22;
      for (int i = -halfWidth; i <= halfWidth; ++i) {
         float wx = rx + i * a11;
         float wy = ry + i * a21;
         const int x = (int)floorf(wx);
         const int y = (int)floorf(wy);
         if (x >= 0 && y >= 0 && x < width && y < height) {
            wx -= x; wy -= y;
            const float *p0 = im + (size_t)y * imcols + x, *p1 = p0 + imcols;
            *out++ = (1.0f - wy) * ((1.0f - wx) * p0[0] + wx * p0[1]) + (wy) * ((1.0f - wx) * p1[0] + wx * p1[1]);
         } else {
            *out++ = 0;
            ret = 1;
         }
      }
   }
   return ret;
}

/* helpers.cpp:246-281 */
static void photometrically_normalize(float *image, const float *mask, int n)
{
   float sum = 0, gsum = 0;
   for (int i = 0; i < n; i++) if (mask[i] > 0) { sum += image[i]; gsum++; }
   sum = sum / gsum;
   float var = 0;
   for (int i = 0; i < n; i++) if (mask[i] > 0) var += (sum - image[i]) * (sum - image[i]);
   var = sqrtf(var / gsum);
   if (var < 0.0001) return;
   float fac = 50.0f / var;
   for (int i = 0; i < n; i++) {
      image[i] = 128 + fac * (image[i] - sum);
      if (image[i] > 255) image[i] = 255;
      if (image[i] < 0) image[i] = 0;
   }
}

/* helpers.cpp:331-339 */
static void half_image(const float *in, int rows, int cols, float *out)
{
   const int nr = rows / 2, nc = cols / 2;
   for (int r = 0; r < nr; r++)
      for (int c = 0; c < nc; c++) out[(size_t)r * nc + c] = in[(size_t)(2 * r) * cols + 2 * c];
}

/* ------------------------------------------------------------------------------------------------
 * AffineShape, affine.cpp
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
   orc_params par;
   float *smm_mask, *img, *fx, *fy;        /* smmWindowSize^2, affine.h:63-75 */
   float *patch;                           /* patchSize^2 */
   float *workspace; size_t workspace_len; /* affine.cpp:119-124 */
   /* SIFTDescriptor state, siftdesc.h:40-49 */
   float *sift_mask, *grad, *ori;
   int *bin0, *bin1; float *w0, *w1;
   float vec[128];
} shape_ctx;

/* affine.cpp:14-33 */
static void compute_gradient(const float *img, int height, int width, float *gradx, float *grady)
{
   for (int r = 0; r < height; ++r)
      for (int c = 0; c < width; ++c) {
         float xgrad, ygrad;
         if (c == 0) xgrad = img[r * width + c + 1] - img[r * width + c];
         else if (c == width - 1) xgrad = img[r * width + c] - img[r * width + c - 1];
         else xgrad = img[r * width + c + 1] - img[r * width + c - 1];
         if (r == 0) ygrad = img[(r + 1) * width + c] - img[r * width + c];
         else if (r == height - 1) ygrad = img[r * width + c] - img[(r - 1) * width + c];
         else ygrad = img[(r + 1) * width + c] - img[(r - 1) * width + c];
         gradx[r * width + c] = xgrad;
         grady[r * width + c] = ygrad;
      }
}

/* affine.cpp:35-100. Returns 1 and fills U/iters on convergence. */
static int find_affine_shape(shape_ctx *cx, const float *blur, int rows, int cols, float x, float y, float s,
                             float pixelDistance, float *U, int *iters)
{
   const orc_params *par = &cx->par;
   float eigen_ratio_act = 0.0f, eigen_ratio_bef = 0.0f;
   float u11 = 1.0f, u12 = 0.0f, u21 = 0.0f, u22 = 1.0f, l1 = 1.0f, l2 = 1.0f;
   float lx = x / pixelDistance, ly = y / pixelDistance;
   float ratio = s / (par->initial_sigma * pixelDistance);
   const int ws = par->smm_window_size;
   const int maskPixels = ws * ws;
   for (int l = 0; l < par->max_iter; l++) {
      interpolate(blur, rows, cols, lx, ly, u11 * ratio, u12 * ratio, u21 * ratio, u22 * ratio, cx->img, ws, ws);
      float a = 0, b = 0, c = 0;
      compute_gradient(cx->img, ws, ws, cx->fx, cx->fy);
      for (int i = 0; i < maskPixels; ++i) {
         const float v = cx->smm_mask[i];
         const float gxx = cx->fx[i];
         const float gyy = cx->fy[i];
         const float gxy = gxx * gyy;
         a += gxx * gxx * v;
         b += gxy * v;
         c += gyy * gyy * v;
      }
      a /= maskPixels; b /= maskPixels; c /= maskPixels;
      inv_sqrt(&a, &b, &c, &l1, &l2);
      eigen_ratio_bef = eigen_ratio_act;
      eigen_ratio_act = 1 - l2 / l1;
      float u11t = u11, u12t = u12;
      u11 = a * u11t + b * u21; u12 = a * u12t + b * u22;
      u21 = b * u11t + c * u21; u22 = b * u12t + c * u22;
      if (!get_eigenvalues(u11, u12, u21, u22, &l1, &l2)) break;
      if ((l1 / l2 > 6) || (l2 / l1 > 6)) break;
      if (eigen_ratio_act < par->convergence_threshold && eigen_ratio_bef < par->convergence_threshold) {
         U[0] = u11; U[1] = u12; U[2] = u21; U[3] = u22; *iters = l;
         return 1;
      }
   }
   return 0;
}

/* affine.cpp:102-144. Returns 1 if the patch is rejected. */
static int normalize_affine(shape_ctx *cx, const float *img, int rows, int cols, float x, float y, float s, float a11,
                            float a12, float a21, float a22)
{
   const orc_params *par = &cx->par;
   const int ps = par->patch_size;
   float mrScale = ceilf(s * par->desc_factor);
   int patchImageSize = 2 * (int)(mrScale) + 1;
   float imageToPatchScale = (float)(patchImageSize) / (float)(ps);
   if (interpolate_check_borders(rows, cols, x, y, a11 * imageToPatchScale, a12 * imageToPatchScale,
                                 a21 * imageToPatchScale, a22 * imageToPatchScale, ps, ps))
      return 1;
   if (imageToPatchScale > 0.4) {
      patchImageSize += 2;
      size_t need = (size_t)patchImageSize * patchImageSize;
      if (need > cx->workspace_len) { free(cx->workspace); cx->workspace = (float *)malloc(sizeof(float) * need); cx->workspace_len = need; }
      float *smoothed = cx->workspace;
      if (!interpolate(img, rows, cols, x, y, a11, a12, a21, a22, smoothed, patchImageSize, patchImageSize)) {
         gaussian_blur(smoothed, patchImageSize, patchImageSize, 1.5f * imageToPatchScale, smoothed);
         interpolate(smoothed, patchImageSize, patchImageSize, (float)(patchImageSize >> 1), (float)(patchImageSize >> 1),
                     imageToPatchScale, 0, 0, imageToPatchScale, cx->patch, ps, ps);
      } else
         return 1;
   } else {
      a11 *= imageToPatchScale; a12 *= imageToPatchScale;
      a21 *= imageToPatchScale; a22 *= imageToPatchScale;
      interpolate(img, rows, cols, x, y, a11, a12, a21, a22, cx->patch, ps, ps);
   }
   return 0;
}

/* ------------------------------------------------------------------------------------------------
 * SIFTDescriptor, siftdesc.cpp
 * ---------------------------------------------------------------------------------------------- */
#define SPATIAL_BINS 4       /* siftdesc.h:27 */
#define ORIENTATION_BINS 8   /* siftdesc.h:28 */
#define MAX_BIN_VALUE 0.2f   /* siftdesc.h:29 */

/* siftdesc.cpp:18-49 */
static void precompute_bins_and_weights(shape_ctx *cx)
{
   const int ps = cx->par.patch_size;
   int halfSize = ps >> 1;
   float step = (float)(SPATIAL_BINS + 1) / (2 * halfSize);
   for (int i = 0; i < ps; i++) {
      float x = step * i;
      int xi = (int)(x);
      cx->bin0[i] = xi - 1;
      cx->bin1[i] = xi;
      cx->w1[i] = x - xi;
      cx->w0[i] = 1.0f - cx->w1[i];
      if (cx->bin0[i] < 0) { cx->bin0[i] = 0; cx->w0[i] = 0; }
      if (cx->bin0[i] >= SPATIAL_BINS) { cx->bin0[i] = SPATIAL_BINS - 1; cx->w0[i] = 0; }
      if (cx->bin1[i] < 0) { cx->bin1[i] = 0; cx->w1[i] = 0; }
      if (cx->bin1[i] >= SPATIAL_BINS) { cx->bin1[i] = SPATIAL_BINS - 1; cx->w1[i] = 0; }
      cx->bin0[i] *= ORIENTATION_BINS;
      cx->bin1[i] *= ORIENTATION_BINS;
   }
}

/* siftdesc.cpp:51-81 */
static void sample_patch(shape_ctx *cx)
{
   const int ps = cx->par.patch_size;
   float *vec = cx->vec;
   for (int r = 0; r < ps; ++r) {
      const int br0 = SPATIAL_BINS * cx->bin0[r]; const float wr0 = cx->w0[r];
      const int br1 = SPATIAL_BINS * cx->bin1[r]; const float wr1 = cx->w1[r];
      for (int c = 0; c < ps; ++c) {
         float val = cx->sift_mask[r * ps + c] * cx->grad[r * ps + c];
         const int bc0 = cx->bin0[c]; const float wc0 = cx->w0[c] * val;
         const int bc1 = cx->bin1[c]; const float wc1 = cx->w1[c] * val;
         const float o = (float)((float)(ORIENTATION_BINS) * (cx->ori[r * ps + c] + 2 * M_PI) / (2 * M_PI));
         int bo0 = (int)o;
         const float wo1 = o - bo0;
         bo0 %= ORIENTATION_BINS;
         int bo1 = (bo0 + 1) % ORIENTATION_BINS;
         const float wo0 = 1.0f - wo1;
         val = wr0 * wc0; if (val > 0) { vec[br0 + bc0 + bo0] += val * wo0; vec[br0 + bc0 + bo1] += val * wo1; }
         val = wr0 * wc1; if (val > 0) { vec[br0 + bc1 + bo0] += val * wo0; vec[br0 + bc1 + bo1] += val * wo1; }
         val = wr1 * wc0; if (val > 0) { vec[br1 + bc0 + bo0] += val * wo0; vec[br1 + bc0 + bo1] += val * wo1; }
         val = wr1 * wc1; if (val > 0) { vec[br1 + bc1 + bo0] += val * wo0; vec[br1 + bc1 + bo1] += val * wo1; }
      }
   }
}

/* siftdesc.cpp:83-96 */
static float sift_normalize(float *vec)
{
   float vectlen = 0.0f;
   for (int i = 0; i < 128; i++) { const float val = vec[i]; vectlen += val * val; }
   vectlen = sqrtf(vectlen);
   const float fac = (float)(1.0f / vectlen);
   for (int i = 0; i < 128; i++) vec[i] *= fac;
   return vectlen;
}

/* siftdesc.cpp:98-113 */
static void sift_sample(shape_ctx *cx)
{
   float *vec = cx->vec;
   for (int i = 0; i < 128; i++) vec[i] = 0;
   sample_patch(cx);
   sift_normalize(vec);
   int changed = 0;
   for (int i = 0; i < 128; i++) if (vec[i] > MAX_BIN_VALUE) { vec[i] = MAX_BIN_VALUE; changed = 1; }
   if (changed) sift_normalize(vec);
   for (int i = 0; i < 128; i++) {
      int b = (int)(512.0f * vec[i]);
      if (b > 255) b = 255;
      vec[i] = (float)(b);
   }
}

/* siftdesc.cpp:115-140 */
static void compute_sift_descriptor(shape_ctx *cx, float *patch)
{
   const int width = cx->par.patch_size, height = cx->par.patch_size;
   photometrically_normalize(patch, cx->sift_mask, width * height);
   for (int r = 0; r < height; ++r)
      for (int c = 0; c < width; ++c) {
         float xgrad, ygrad;
         if (c == 0) xgrad = patch[r * width + c + 1] - patch[r * width + c];
         else if (c == width - 1) xgrad = patch[r * width + c] - patch[r * width + c - 1];
         else xgrad = patch[r * width + c + 1] - patch[r * width + c - 1];
         if (r == 0) ygrad = patch[(r + 1) * width + c] - patch[r * width + c];
         else if (r == height - 1) ygrad = patch[r * width + c] - patch[(r - 1) * width + c];
         else ygrad = patch[(r + 1) * width + c] - patch[(r - 1) * width + c];
         cx->grad[r * width + c] = sqrtf(xgrad * xgrad + ygrad * ygrad);
         cx->ori[r * width + c] = atan2f(ygrad, xgrad);
      }
   sift_sample(cx);
}

static void ctx_init(shape_ctx *cx, const orc_params *p)
{
   memset(cx, 0, sizeof(*cx));
   cx->par = *p;
   const int ws = p->smm_window_size, ps = p->patch_size;
   cx->smm_mask = (float *)malloc(sizeof(float) * ws * ws);
   cx->img = (float *)malloc(sizeof(float) * ws * ws);
   cx->fx = (float *)calloc(ws * ws, sizeof(float));
   cx->fy = (float *)calloc(ws * ws, sizeof(float));
   cx->patch = (float *)malloc(sizeof(float) * ps * ps);
   cx->sift_mask = (float *)malloc(sizeof(float) * ps * ps);
   cx->grad = (float *)malloc(sizeof(float) * ps * ps);
   cx->ori = (float *)malloc(sizeof(float) * ps * ps);
   cx->bin0 = (int *)malloc(sizeof(int) * ps); cx->bin1 = (int *)malloc(sizeof(int) * ps);
   cx->w0 = (float *)malloc(sizeof(float) * ps); cx->w1 = (float *)malloc(sizeof(float) * ps);
   compute_gauss_mask(cx->smm_mask, ws);
   compute_circular_gauss_mask(cx->sift_mask, ps);
   precompute_bins_and_weights(cx);
}

static void ctx_free(shape_ctx *cx)
{
   free(cx->smm_mask); free(cx->img); free(cx->fx); free(cx->fy); free(cx->patch); free(cx->workspace);
   free(cx->sift_mask); free(cx->grad); free(cx->ori); free(cx->bin0); free(cx->bin1); free(cx->w0); free(cx->w1);
}

/* ------------------------------------------------------------------------------------------------
 * HessianDetector, pyramid.cpp
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
   orc_params par;
   float edgeScoreThreshold, finalThreshold, positiveThreshold, negativeThreshold;   /* pyramid.h:57-64 */
   int rows, cols;
   const float *low, *cur, *high, *blur, *prevBlur;
   unsigned char *octaveMap;
   /* glue (hesaff.cpp:50-105) */
   const float *image; int irows, icols;
   shape_ctx shape;
   orc_detection *dets; int ndets, cap;
} detector;

/* pyramid.cpp:24-37 */
static int hessian_point_type(const float *ptr, float value)
{
   if (value < 0) return 2;
   float Lxx = (ptr[-1] - 2 * ptr[0] + ptr[1]);
   return Lxx < 0 ? 0 : 1;
}

/* pyramid.cpp:39-61 */
static int is_max(float val, const float *pix, int cols, int row, int col)
{
   for (int r = row - 1; r <= row + 1; r++)
      for (int c = col - 1; c <= col + 1; c++) if (pix[(size_t)r * cols + c] > val) return 0;
   return 1;
}
static int is_min(float val, const float *pix, int cols, int row, int col)
{
   for (int r = row - 1; r <= row + 1; r++)
      for (int c = col - 1; c <= col + 1; c++) if (pix[(size_t)r * cols + c] < val) return 0;
   return 1;
}

/* hesaff.cpp:66-105: the two callbacks, flattened */
static void on_keypoint(detector *D, float x, float y, float s, float pixelDistance, int type, float response)
{
   if (D->ndets == D->cap) { D->cap = D->cap ? 2 * D->cap : 1024; D->dets = (orc_detection *)realloc(D->dets, sizeof(orc_detection) * D->cap); }
   orc_detection *d = &D->dets[D->ndets++];
   memset(d, 0, sizeof(*d));
   d->x = x; d->y = y; d->s = s; d->pd = pixelDistance; d->type = type; d->response = response;
   float U[4]; int iters;
   if (!find_affine_shape(&D->shape, D->prevBlur, D->rows, D->cols, x, y, s, pixelDistance, U, &iters)) return;
   d->affine_ok = 1; d->u11 = U[0]; d->u12 = U[1]; d->u21 = U[2]; d->u22 = U[3]; d->iters = iters;
   float a11 = U[0], a12 = U[1], a21 = U[2], a22 = U[3];
   rectify_up_is_up(&a11, &a12, &a21, &a22);
   d->a11 = a11; d->a12 = a12; d->a21 = a21; d->a22 = a22;
   if (!normalize_affine(&D->shape, D->image, D->irows, D->icols, x, y, s, a11, a12, a21, a22)) {
      compute_sift_descriptor(&D->shape, D->shape.patch);
      d->described = 1;
      for (int i = 0; i < 128; i++) d->desc[i] = (unsigned char)D->shape.vec[i];
   }
}

#define AT(p, r, c) ((p)[(size_t)(r) * cols + (c)])

/* pyramid.cpp:122-204 */
static void localize_keypoint(detector *D, int r, int c, float curScale, float pixelDistance)
{
   const int cols = D->cols, rows = D->rows;
   const float *cur = D->cur, *low = D->low, *high = D->high;
   float b[3] = {0, 0, 0};
   float val = 0;
   int nr = r, nc = c;
   for (int iter = 0; iter < 5; iter++) {
      r = nr; c = nc;
      float dxx = AT(cur, r, c - 1) - 2.0f * AT(cur, r, c) + AT(cur, r, c + 1);
      float dyy = AT(cur, r - 1, c) - 2.0f * AT(cur, r, c) + AT(cur, r + 1, c);
      float dss = AT(low, r, c) - 2.0f * AT(cur, r, c) + AT(high, r, c);
      float dxy = 0.25f * (AT(cur, r + 1, c + 1) - AT(cur, r + 1, c - 1) - AT(cur, r - 1, c + 1) + AT(cur, r - 1, c - 1));
      if (0 == iter) {
         float edgeScore = (dxx + dyy) * (dxx + dyy) / (dxx * dyy - dxy * dxy);
         if (edgeScore >= D->edgeScoreThreshold || edgeScore < 0) return;
      }
      float dxs = 0.25f * (AT(high, r, c + 1) - AT(high, r, c - 1) - AT(low, r, c + 1) + AT(low, r, c - 1));
      float dys = 0.25f * (AT(high, r + 1, c) - AT(high, r - 1, c) - AT(low, r + 1, c) + AT(low, r - 1, c));
      float A[9];
      A[0] = dxx; A[1] = dxy; A[2] = dxs;
      A[3] = dxy; A[4] = dyy; A[5] = dys;
      A[6] = dxs; A[7] = dys; A[8] = dss;
      float dx = 0.5f * (AT(cur, r, c + 1) - AT(cur, r, c - 1));
      float dy = 0.5f * (AT(cur, r + 1, c) - AT(cur, r - 1, c));
      float ds = 0.5f * (AT(high, r, c) - AT(low, r, c));
      b[0] = -dx; b[1] = -dy; b[2] = -ds;
      solve_linear_3x3(A, b);
      if (isnan(b[0]) || isnan(b[1]) || isnan(b[2])) return;
      val = AT(cur, r, c) + 0.5f * (dx * b[0] + dy * b[1] + ds * b[2]);
      /* MAX_SUBPIXEL_SHIFT is the double literal 0.6 (pyramid.cpp:117) */
      if (b[0] > 0.6) { if (c < cols - 3) nc++; else return; }
      if (b[1] > 0.6) { if (r < rows - 3) nr++; else return; }
      if (b[0] < -0.6) { if (c > 3) nc--; else return; }
      if (b[1] < -0.6) { if (r > 3) nr--; else return; }
      if (nr == r && nc == c) break;
   }
   if (fabsf(b[0]) > 1.5 || fabsf(b[1]) > 1.5 || fabsf(b[2]) > 1.5 || fabsf(val) < D->finalThreshold ||
       D->octaveMap[(size_t)r * cols + c] > 0)
      return;
   D->octaveMap[(size_t)r * cols + c] = 1;
   float scale = curScale * powf(2.0f, b[2] / D->par.number_of_scales);
   int type = hessian_point_type(D->blur + (size_t)r * cols + c, val);
   on_keypoint(D, pixelDistance * (c + b[0]), pixelDistance * (r + b[1]), pixelDistance * scale, pixelDistance, type, val);
}

/* pyramid.cpp:206-222 */
static void find_level_keypoints(detector *D, float curScale, float pixelDistance)
{
   const int rows = D->rows, cols = D->cols, border = D->par.border;
   for (int r = border; r < (rows - border); r++)
      for (int c = border; c < (cols - border); c++) {
         const float val = AT(D->cur, r, c);
         if ((val > D->positiveThreshold && (is_max(val, D->cur, cols, r, c) && is_max(val, D->low, cols, r, c) && is_max(val, D->high, cols, r, c))) ||
             (val < D->negativeThreshold && (is_min(val, D->cur, cols, r, c) && is_min(val, D->low, cols, r, c) && is_min(val, D->high, cols, r, c))))
            localize_keypoint(D, r, c, curScale, pixelDistance);
      }
}

/* pyramid.cpp:224-259. L/R (optional) receive copies of all S+2 planes. Returns the next octave's first level. */
static float *detect_octave(detector *D, const float *firstLevel, int rows, int cols, float pixelDistance, int detect,
                            float *Lout, float *Rout)
{
   const int S = D->par.number_of_scales;
   const size_t n = (size_t)rows * cols;
   D->rows = rows; D->cols = cols;
   if (detect) D->octaveMap = (unsigned char *)calloc(n, 1);
   float sigmaStep = powf(2.0f, 1.0f / (float)S);
   float curSigma = D->par.initial_sigma;
   float **L = (float **)calloc(S + 2, sizeof(float *)), **R = (float **)calloc(S + 2, sizeof(float *));
   L[0] = (float *)malloc(sizeof(float) * n);
   memcpy(L[0], firstLevel, sizeof(float) * n);
   R[0] = (float *)malloc(sizeof(float) * n);
   hessian_response(L[0], rows, cols, curSigma * curSigma, R[0]);
   float *next = NULL;
   for (int i = 1; i < S + 2; i++) {
      float sigma = curSigma * sqrtf(sigmaStep * sigmaStep - 1.0f);
      L[i] = (float *)malloc(sizeof(float) * n);
      gaussian_blur(L[i - 1], rows, cols, sigma, L[i]);
      sigma = curSigma * sigmaStep;
      R[i] = (float *)malloc(sizeof(float) * n);
      hessian_response(L[i], rows, cols, sigma * sigma, R[i]);
      if (i >= 2 && detect) {
         D->low = R[i - 2]; D->cur = R[i - 1]; D->high = R[i];
         D->blur = L[i - 1]; D->prevBlur = L[i - 2];
         find_level_keypoints(D, curSigma, pixelDistance);
      }
      if (i == S) { next = (float *)malloc(sizeof(float) * (size_t)(rows / 2) * (cols / 2) + 4); half_image(L[i], rows, cols, next); }
      curSigma *= sigmaStep;
   }
   for (int i = 0; i < S + 2; i++) {
      if (Lout) memcpy(Lout + n * i, L[i], sizeof(float) * n);
      if (Rout) memcpy(Rout + n * i, R[i], sizeof(float) * n);
      free(L[i]); free(R[i]);
   }
   free(L); free(R);
   if (detect) { free(D->octaveMap); D->octaveMap = NULL; }
   return next;
}

static void detector_init(detector *D, const orc_params *p)
{
   memset(D, 0, sizeof(*D));
   D->par = *p;
   /* pyramid.h:57-64 */
   D->edgeScoreThreshold = (p->edge_eigenvalue_ratio + 1.0f) * (p->edge_eigenvalue_ratio + 1.0f) / p->edge_eigenvalue_ratio;
   D->finalThreshold = p->threshold * p->threshold;
   D->positiveThreshold = (float)(0.8 * D->finalThreshold);
   D->negativeThreshold = -D->positiveThreshold;
   ctx_init(&D->shape, p);
}

static void first_level(const float *image, int h, int w, const orc_params *p, float *dst)
{
   /* pyramid.cpp:263,273-280 (upscaleInputImage == 0) */
   float curSigma = 0.5f;
   if (p->initial_sigma > curSigma) {
      float sigma = sqrtf(p->initial_sigma * p->initial_sigma - curSigma * curSigma);
      gaussian_blur(image, h, w, sigma, dst);
   } else
      memcpy(dst, image, sizeof(float) * (size_t)h * w);
}

void orc_first_level(const float *image, int h, int w, const orc_params *p, float *dst) { first_level(image, h, w, p, dst); }

/* pyramid.cpp:261-292 */
int orc_detect(const float *image, int h, int w, const orc_params *p, orc_detection **out)
{
   detector D;
   detector_init(&D, p);
   D.image = image; D.irows = h; D.icols = w;
   float pixelDistance = 1.0f;
   float *firstLevel = (float *)malloc(sizeof(float) * (size_t)h * w + 4);
   first_level(image, h, w, p, firstLevel);
   int rows = h, cols = w, octave = 0;
   int minSize = 2 * p->border + 2;
   while (rows > minSize && cols > minSize) {
      if (p->max_octaves > 0 && octave >= p->max_octaves) break;
      float *next = detect_octave(&D, firstLevel, rows, cols, pixelDistance, 1, NULL, NULL);
      pixelDistance *= 2.0;
      free(firstLevel);
      firstLevel = next;
      rows /= 2; cols /= 2; octave++;
   }
   free(firstLevel);
   ctx_free(&D.shape);
   if (!D.dets) D.dets = (orc_detection *)malloc(sizeof(orc_detection));
   *out = D.dets;
   return D.ndets;
}

void orc_octave_planes(const float *first, int h, int w, const orc_params *p, float *L, float *R, float *next)
{
   detector D;
   detector_init(&D, p);
   float *nx = detect_octave(&D, first, h, w, 1.0f, 0, L, R);
   if (next && nx) memcpy(next, nx, sizeof(float) * (size_t)(h / 2) * (w / 2));
   free(nx);
   ctx_free(&D.shape);
}

int orc_find_affine_shape(const float *blur, int h, int w, const orc_params *p, float x, float y, float s, float pd,
                          float *U, int *iters)
{
   shape_ctx cx;
   ctx_init(&cx, p);
   int ok = find_affine_shape(&cx, blur, h, w, x, y, s, pd, U, iters);
   ctx_free(&cx);
   return ok;
}

int orc_normalize_affine(const float *img, int h, int w, const orc_params *p, float x, float y, float s, float a11,
                         float a12, float a21, float a22, float *patch)
{
   shape_ctx cx;
   ctx_init(&cx, p);
   int rej = normalize_affine(&cx, img, h, w, x, y, s, a11, a12, a21, a22);
   if (!rej) memcpy(patch, cx.patch, sizeof(float) * p->patch_size * p->patch_size);
   ctx_free(&cx);
   return rej;
}

void orc_sift(float *patch, const orc_params *p, unsigned char *desc)
{
   shape_ctx cx;
   ctx_init(&cx, p);
   compute_sift_descriptor(&cx, patch);
   for (int i = 0; i < 128; i++) desc[i] = (unsigned char)cx.vec[i];
   ctx_free(&cx);
}
