/* oracle/oracle_api.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * One C API, two implementations, so that every parity test can be run against either:
 *   libhesaff_oracle.so      hesaff_oracle.c : our plain-C restatement of the reference path
 *   _ref/libhesaff_ref.so    ref_driver.cpp  : the reference's own sources (compiled unmodified
 *                                              from /root/reference against oracle/shim/cv.h)
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline leg may load these.
 */
#ifndef HESAFF_ORACLE_API_H
#define HESAFF_ORACLE_API_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_params {
   /* HessianAffineParams, hesaff.cpp:21-36 */
   float threshold;
   int max_iter;
   float desc_factor;
   int patch_size;
   /* defaults the reference CLI never changes: pyramid.h:34-39, affine.h:39-44 */
   int number_of_scales;
   float initial_sigma;
   float edge_eigenvalue_ratio;
   int border;
   float convergence_threshold;
   int smm_window_size;
   /* not in the reference (pyramid.cpp:283-284 has no cap): 0 = all octaves */
   int max_octaves;
} orc_params;

typedef struct orc_detection {
   /* HessianKeypointCallback::onHessianKeypointDetected, pyramid.h:43-47 */
   float x, y, s, pd;
   int type;
   float response;
   /* AffineShapeCallback::onAffineShapeFound, affine.h:48-58 (before rectification) */
   int affine_ok;
   float u11, u12, u21, u22;
   int iters;
   /* Keypoint, hesaff.cpp:41-48 (after rectifyAffineTransformationUpIsUp + normalizeAffine + SIFT) */
   int described;
   float a11, a12, a21, a22;
   unsigned char desc[128];
} orc_detection;

const char *orc_name(void);
void orc_default_params(orc_params *p);
void orc_free(void *p);

/* detectPyramidKeypoints + callbacks (pyramid.cpp:261-292, hesaff.cpp:66-105).
 * Returns the number of detections (g_numberOfPoints); *out is malloc'ed, in reference order. */
int orc_detect(const float *image, int h, int w, const orc_params *p, orc_detection **out);

/* helpers.cpp:283-295 (size from sigma, BORDER_REPLICATE). dst may equal src. */
void orc_gaussian_blur(const float *src, int h, int w, float sigma, float *dst);
/* pyramid.cpp:63-114; the 1-px border the reference leaves uninitialised is written as 0. */
void orc_hessian_response(const float *src, int h, int w, float norm, float *dst);
/* The blur/response planes of one octave, pyramid.cpp:224-259: L and R hold S+2 planes of h*w;
 * next is the first level of the next octave ((h/2)*(w/2)), helpers.cpp:331-339. */
void orc_octave_planes(const float *first_level, int h, int w, const orc_params *p, float *L, float *R, float *next);
/* first blur of the input image, pyramid.cpp:273-280 */
void orc_first_level(const float *image, int h, int w, const orc_params *p, float *dst);
/* affine.cpp:35-100: returns 1 on convergence and fills U (u11,u12,u21,u22) and iters */
int orc_find_affine_shape(const float *blur, int h, int w, const orc_params *p, float x, float y, float s, float pd,
                          float *U, int *iters);
/* helpers.cpp:90-97 */
void orc_rectify(float *A);
/* affine.cpp:102-144: returns 1 if rejected (touches boundary), else 0 and fills patch (patch_size^2) */
int orc_normalize_affine(const float *img, int h, int w, const orc_params *p, float x, float y, float s,
                         float a11, float a12, float a21, float a22, float *patch);
/* siftdesc.cpp:115-140: patch is photometrically normalised in place; desc = quantised vec */
void orc_sift(float *patch, const orc_params *p, unsigned char *desc);

#ifdef __cplusplus
}
#endif
#endif
