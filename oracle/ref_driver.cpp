// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Implements oracle_api.h on top of the reference's OWN objects (pyramid.cpp, affine.cpp,
// siftdesc.cpp, helpers.cpp compiled unmodified from /root/reference by oracle/Makefile).
// The glue class mirrors AffineHessianDetector (hesaff.cpp:50-105) but records every stage.
#include "pyramid.h"
#include "helpers.h"
#include "affine.h"
#include "siftdesc.h"
#include "oracle_api.h"

using namespace cv;

namespace {

void splitParams(const orc_params &o, PyramidParams &p, AffineShapeParams &ap, SIFTDescriptorParams &sp)
{
   // hesaff.cpp:150-163, plus the struct members the CLI leaves at their defaults
   p.threshold = o.threshold;
   p.numberOfScales = o.number_of_scales;
   p.initialSigma = o.initial_sigma;
   p.edgeEigenValueRatio = o.edge_eigenvalue_ratio;
   p.border = o.border;
   ap.maxIterations = o.max_iter;
   ap.patchSize = o.patch_size;
   ap.mrSize = o.desc_factor;
   ap.convergenceThreshold = o.convergence_threshold;
   ap.smmWindowSize = o.smm_window_size;
   ap.initialSigma = o.initial_sigma;
   sp.patchSize = o.patch_size;
}

struct RefDetector : public HessianDetector, AffineShape, HessianKeypointCallback, AffineShapeCallback {
   const Mat image;
   SIFTDescriptor sift;
   std::vector<orc_detection> dets;
   RefDetector(const Mat &image, const PyramidParams &par, const AffineShapeParams &ap, const SIFTDescriptorParams &sp)
      : HessianDetector(par), AffineShape(ap), image(image), sift(sp)
   {
      setHessianKeypointCallback(this);
      setAffineShapeCallback(this);
   }
   void onHessianKeypointDetected(const Mat &blur, float x, float y, float s, float pixelDistance, int type, float response)
   {
      orc_detection d;
      memset(&d, 0, sizeof(d));
      d.x = x; d.y = y; d.s = s; d.pd = pixelDistance; d.type = type; d.response = response;
      dets.push_back(d);
      findAffineShape(blur, x, y, s, pixelDistance, type, response);
   }
   void onAffineShapeFound(const Mat &, float x, float y, float s, float, float a11, float a12, float a21, float a22,
                           int, float, int iters)
   {
      orc_detection &d = dets.back();
      d.affine_ok = 1; d.u11 = a11; d.u12 = a12; d.u21 = a21; d.u22 = a22; d.iters = iters;
      rectifyAffineTransformationUpIsUp(a11, a12, a21, a22);
      d.a11 = a11; d.a12 = a12; d.a21 = a21; d.a22 = a22;
      if (!normalizeAffine(image, x, y, s, a11, a12, a21, a22)) {
         sift.computeSiftDescriptor(this->patch);
         d.described = 1;
         for (int i = 0; i < 128; i++) d.desc[i] = (unsigned char)sift.vec[i];
      }
   }
   Mat response(const Mat &m, float norm) { return hessianResponse(m, norm); }
   bool runAffine(const Mat &blur, float x, float y, float s, float pd) { return findAffineShape(blur, x, y, s, pd, 0, 0.f); }
};

Mat wrap(const float *p, int h, int w) { return Mat(h, w, CV_32FC1, (void *)p); }
void copyOut(const Mat &m, float *dst) { for (int r = 0; r < m.rows; r++) memcpy(dst + (size_t)r * m.cols, m.ptr<float>(r), sizeof(float) * m.cols); }

} // namespace

extern "C" {

const char *orc_name(void) { return "reference (perdoch/hesaff sources + oracle/shim)"; }

void orc_default_params(orc_params *o)
{
   // hesaff.cpp:28-35 and the struct constructors
   PyramidParams p; AffineShapeParams ap;
   o->threshold = 16.0f / 3.0f; o->max_iter = 16; o->desc_factor = 3.0f * sqrtf(3.0f); o->patch_size = 41;
   o->number_of_scales = p.numberOfScales; o->initial_sigma = p.initialSigma;
   o->edge_eigenvalue_ratio = p.edgeEigenValueRatio; o->border = p.border;
   o->convergence_threshold = ap.convergenceThreshold; o->smm_window_size = ap.smmWindowSize;
   o->max_octaves = 0;
}

void orc_free(void *p) { free(p); }

int orc_detect(const float *image, int h, int w, const orc_params *o, orc_detection **out)
{
   PyramidParams p; AffineShapeParams ap; SIFTDescriptorParams sp;
   splitParams(*o, p, ap, sp);
   Mat img = wrap(image, h, w).clone();
   RefDetector det(img, p, ap, sp);
   det.detectPyramidKeypoints(img);
   // the reference has no octave cap; octaves are independent, so dropping the later ones is equivalent
   std::vector<orc_detection> keep;
   float maxpd = o->max_octaves > 0 ? (float)(1 << (o->max_octaves - 1)) : 1e30f;
   for (size_t i = 0; i < det.dets.size(); i++) if (det.dets[i].pd <= maxpd) keep.push_back(det.dets[i]);
   *out = (orc_detection *)malloc(sizeof(orc_detection) * (keep.size() + 1));
   if (!keep.empty()) memcpy(*out, &keep[0], sizeof(orc_detection) * keep.size());
   return (int)keep.size();
}

void orc_gaussian_blur(const float *src, int h, int w, float sigma, float *dst)
{
   Mat r = gaussianBlur(wrap(src, h, w), sigma);
   copyOut(r, dst);
}

void orc_hessian_response(const float *src, int h, int w, float norm, float *dst)
{
   PyramidParams p; AffineShapeParams ap; SIFTDescriptorParams sp;
   Mat img = wrap(src, h, w);
   RefDetector det(img, p, ap, sp);
   Mat r = det.response(img, norm);
   for (int y = 0; y < h; y++)
      for (int x = 0; x < w; x++)
         dst[(size_t)y * w + x] = (y == 0 || x == 0 || y == h - 1 || x == w - 1) ? 0.f : r.at<float>(y, x);
}

void orc_first_level(const float *image, int h, int w, const orc_params *o, float *dst)
{
   // pyramid.cpp:263,273-280
   Mat first = wrap(image, h, w).clone();
   float curSigma = 0.5f;
   if (o->initial_sigma > curSigma) {
      float sigma = sqrt(o->initial_sigma * o->initial_sigma - curSigma * curSigma);
      gaussianBlurInplace(first, sigma);
   }
   copyOut(first, dst);
}

void orc_octave_planes(const float *first_level, int h, int w, const orc_params *o, float *L, float *R, float *next)
{
   // same schedule as detectOctaveKeypoints (pyramid.cpp:227-257), built from the reference's functions
   const int S = o->number_of_scales;
   const size_t n = (size_t)h * w;
   float sigmaStep = pow(2.0f, 1.0f / (float)S);
   float curSigma = o->initial_sigma;
   Mat blur = wrap(first_level, h, w).clone();
   copyOut(blur, L);
   orc_hessian_response(L, h, w, curSigma * curSigma, R);
   for (int i = 1; i < S + 2; i++) {
      float sigma = curSigma * sqrt(sigmaStep * sigmaStep - 1.0f);
      Mat nextBlur = gaussianBlur(blur, sigma);
      sigma = curSigma * sigmaStep;
      copyOut(nextBlur, L + n * i);
      orc_hessian_response(L + n * i, h, w, sigma * sigma, R + n * i);
      if (i == S && next) { Mat half = halfImage(nextBlur); copyOut(half, next); }
      blur = nextBlur;
      curSigma *= sigmaStep;
   }
}

struct ShapeGrabber : public AffineShapeCallback {
   float U[4]; int iters; bool got;
   ShapeGrabber() : iters(0), got(false) {}
   void onAffineShapeFound(const Mat &, float, float, float, float, float a11, float a12, float a21, float a22, int, float, int it)
   { U[0] = a11; U[1] = a12; U[2] = a21; U[3] = a22; iters = it; got = true; }
};

int orc_find_affine_shape(const float *blur, int h, int w, const orc_params *o, float x, float y, float s, float pd,
                          float *U, int *iters)
{
   PyramidParams p; AffineShapeParams ap; SIFTDescriptorParams sp;
   splitParams(*o, p, ap, sp);
   AffineShape shape(ap);
   ShapeGrabber g;
   shape.setAffineShapeCallback(&g);
   bool ok = shape.findAffineShape(wrap(blur, h, w), x, y, s, pd, 0, 0.f);
   if (ok) { memcpy(U, g.U, sizeof(g.U)); *iters = g.iters; }
   return ok ? 1 : 0;
}

void orc_rectify(float *A) { rectifyAffineTransformationUpIsUp(A); }

int orc_normalize_affine(const float *img, int h, int w, const orc_params *o, float x, float y, float s,
                         float a11, float a12, float a21, float a22, float *patch)
{
   PyramidParams p; AffineShapeParams ap; SIFTDescriptorParams sp;
   splitParams(*o, p, ap, sp);
   AffineShape shape(ap);
   bool rejected = shape.normalizeAffine(wrap(img, h, w), x, y, s, a11, a12, a21, a22);
   if (!rejected) copyOut(shape.patch, patch);
   return rejected ? 1 : 0;
}

void orc_sift(float *patch, const orc_params *o, unsigned char *desc)
{
   PyramidParams p; AffineShapeParams ap; SIFTDescriptorParams sp;
   splitParams(*o, p, ap, sp);
   SIFTDescriptor sift(sp);
   Mat m = wrap(patch, o->patch_size, o->patch_size);
   sift.computeSiftDescriptor(m);
   for (int i = 0; i < 128; i++) desc[i] = (unsigned char)sift.vec[i];
}

} // extern "C"
