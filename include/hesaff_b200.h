/* include/hesaff_b200.h -- C-ABI of the B200-native Hessian-Affine + SIFT hot path.
 *
 * Drop-in boundary for perdoch/hesaff's detect -> affine -> describe path.  The reference has no FFI;
 * its de-facto boundary is the class interface driven by main():
 *     HessianAffineParams                                   hesaff.cpp:21-36
 *     HessianDetector::detectPyramidKeypoints(const Mat&)   pyramid.h:73   (pyramid.cpp:261-292)
 *       -> HessianKeypointCallback::onHessianKeypointDetected   pyramid.h:43-47
 *       -> AffineShape::findAffineShape / normalizeAffine       affine.h:82,85
 *       -> SIFTDescriptor::computeSiftDescriptor                siftdesc.h:51
 *     struct Keypoint / vector<Keypoint> keys                hesaff.cpp:41-48,54
 *     AffineHessianDetector::exportKeypoints                 hesaff.cpp:107-130
 * The per-keypoint callbacks serialise the pipeline, so this ABI is batch-in / keypoints-out: one call
 * takes N gray images and yields, per image, the same Keypoint records in the same order.
 *
 * Plain C: pointers, sizes, ints.  No C++/torch types.  All functions return HESAFF_OK (0) or a negative
 * hesaff_status; hesaff_last_error() gives the text.  A context is single-owner (one host thread, one
 * GPU); several contexts may run concurrently on different GPUs.  There is NO CPU fallback: every entry
 * point fails with HESAFF_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef HESAFF_B200_H
#define HESAFF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HESAFF_B200_ABI_VERSION 1

typedef enum hesaff_status {
   HESAFF_OK = 0,
   HESAFF_ERR_INVALID = -1,   /* bad argument / unsupported parameter value */
   HESAFF_ERR_CUDA = -2,      /* CUDA runtime/driver error, or no usable device */
   HESAFF_ERR_CAPACITY = -3,  /* a fixed-capacity candidate/keypoint buffer overflowed (never silent) */
   HESAFF_ERR_STATE = -4      /* results requested before a successful hesaff_detect */
} hesaff_status;

/* HessianAffineParams (hesaff.cpp:21-36) followed by the members of PyramidParams (pyramid.h:18-41) and
 * AffineShapeParams (affine.h:17-46) that the reference CLI leaves at their constructor defaults, and
 * max_octaves, which the reference lacks (pyramid.cpp:283-284 runs until the image is <= 12 px). */
typedef struct hesaff_params {
   float threshold;               /* 16/3  -> PyramidParams.threshold          hesaff.cpp:30,155 */
   int max_iter;                  /* 16    -> AffineShapeParams.maxIterations  hesaff.cpp:31,158 */
   float desc_factor;             /* 3*sqrt(3) -> AffineShapeParams.mrSize     hesaff.cpp:32,160 */
   int patch_size;                /* 41 (only 41 is supported)                 hesaff.cpp:33,159,163 */
   int verbose;                   /* never read by the reference               hesaff.cpp:27,34 */
   int number_of_scales;          /* 3     pyramid.h:35 */
   float initial_sigma;           /* 1.6   pyramid.h:36, affine.h:40 */
   float edge_eigenvalue_ratio;   /* 10    pyramid.h:38 */
   int border;                    /* 5     pyramid.h:39 */
   float convergence_threshold;   /* 0.05  affine.h:41 */
   int smm_window_size;           /* 19 (only 19 is supported)  affine.h:43 */
   int max_octaves;               /* 0 = reference behaviour (all octaves) */
} hesaff_params;

/* struct Keypoint, hesaff.cpp:41-48: 164 bytes, same member order. */
typedef struct hesaff_keypoint {
   float x, y, s;
   float a11, a12, a21, a22;
   float response;
   int type;                      /* HessianDetector::HESSIAN_DARK=0 / BRIGHT=1 / SADDLE=2, pyramid.h:51-55 */
   unsigned char desc[128];
} hesaff_keypoint;

/* One record per detection (= per onHessianKeypointDetected call), for stage-level parity tests. */
typedef struct hesaff_detection {
   float x, y, s, pd;             /* pyramid.cpp:203 */
   int type;
   float response;
   int affine_ok;                 /* findAffineShape converged, affine.cpp:92-97 */
   float u11, u12, u21, u22;      /* shape before rectification */
   int iters;
   int described;                 /* normalizeAffine succeeded, hesaff.cpp:82 */
   float a11, a12, a21, a22;      /* after rectifyAffineTransformationUpIsUp, helpers.cpp:90-97 */
   unsigned char desc[128];
} hesaff_detection;

typedef struct hesaff_ctx hesaff_ctx;

/* ---- lifetime -------------------------------------------------------------------------------- */
int hesaff_abi_version(void);
const char *hesaff_last_error(void);

/* Fills *p with the reference defaults (HessianAffineParams(), PyramidParams(), AffineShapeParams()). */
int hesaff_params_default(hesaff_params *p);

/* Creates a context on CUDA device `device` able to process batches of images up to max_width x
 * max_height.  `max_batch` images are resident at once (larger batches are processed in chunks);
 * 0 = choose from free device memory.  `max_candidates_per_image` sizes the candidate/keypoint pool
 * (pool = chunk images x this); 0 = width*height/10.  Replaces the AffineHessianDetector constructor
 * (hesaff.cpp:56-64), which precomputes the masks/bin tables (affine.h:63-75, siftdesc.h:40-49). */
int hesaff_create(hesaff_ctx **out, const hesaff_params *p, int device, int max_width, int max_height,
                  int max_batch, int max_candidates_per_image);
int hesaff_destroy(hesaff_ctx *ctx);

/* ---- the hot path ---------------------------------------------------------------------------- */
/* detectPyramidKeypoints + both callbacks for n images of width x height (hesaff.cpp:166-167).
 * `images` holds n planes, plane i at images + i*image_stride_bytes, rows row_pitch_bytes apart.
 *   hesaff_detect_u8   : 8-bit gray (what main() builds from a gray PGM: (B+G+R)/3.0f is exact, hesaff.cpp:138-148)
 *   hesaff_detect_f32  : float gray (for colour input converted by the caller with the same expression)
 *   hesaff_detect_rgb8 : 8-bit, 3 interleaved channels (what imread hands to main(), hesaff.cpp:137); the gray
 *                        conversion (float(c0)+c1+c2)/3.0f of hesaff.cpp:145 runs on the GPU
 * `on_device` != 0 means `images` is a device pointer on the context's GPU; otherwise it is host memory
 * (pinned memory makes the upload asynchronous).  `stream` is a cudaStream_t; NULL selects the context's own
 * non-blocking stream -- it does NOT mean the CUDA default stream (pass cudaStreamLegacy / cudaStreamPerThread for
 * those).  With NULL and a device-resident input the call is ordered after the work already queued on the legacy
 * default stream; input produced on any other stream must be passed together with that stream.  The call returns
 * after the work is enqueued AND the per-image counts are known on the host.
 * Results stay on the device until the next hesaff_detect_* call on this context. */
int hesaff_detect_u8(hesaff_ctx *ctx, const uint8_t *images, int n, int width, int height, size_t row_pitch_bytes,
                     size_t image_stride_bytes, int on_device, void *stream);
int hesaff_detect_f32(hesaff_ctx *ctx, const float *images, int n, int width, int height, size_t row_pitch_bytes,
                      size_t image_stride_bytes, int on_device, void *stream);
int hesaff_detect_rgb8(hesaff_ctx *ctx, const uint8_t *images, int n, int width, int height, size_t row_pitch_bytes,
                       size_t image_stride_bytes, int on_device, void *stream);

/* Image files as they are on disk (replaces cv::imread + the conversion loop, hesaff.cpp:137-148): `files[i]` points to the
 * bytes of a binary PNM file (P5 gray or P6 colour, maxval 255) in host memory, `file_bytes[i]` is its length.  Only the
 * header is parsed on the host; the pixel payload is uploaded untouched and the gray conversion runs on the GPU.  All files
 * of one call must have the same width, height and type.  hesaff_pnm_info parses one header (any output may be NULL). */
int hesaff_pnm_info(const void *file, size_t bytes, int *width, int *height, int *channels, size_t *data_offset);
int hesaff_detect_pnm(hesaff_ctx *ctx, const void *const *files, const size_t *file_bytes, int n, void *stream);

/* The consumer of the records (the step after the path, README:49-53: matching SIFT descriptors): for every query record the
 * nearest database record by squared L2 distance over the 128 descriptor bytes (exact integer arithmetic; ties go to the
 * lower index), that distance and the distance of the second nearest (Lowe's ratio test; may be NULL).  All pointers are
 * DEVICE pointers on `device` -- e.g. hesaff_result_keypoints_device, or the records of every rank after the variable-size
 * all-gather.  With n_db == 0 every index is -1 and the distances are 0xffffffff.  Synchronises `stream` before it returns. */
int hesaff_match_descriptors(int device, const hesaff_keypoint *d_query, size_t n_query, const hesaff_keypoint *d_db,
                             size_t n_db, int32_t *d_best_index, uint32_t *d_best_dist2, uint32_t *d_second_dist2, void *stream);

/* ---- results of the last detect call ---------------------------------------------------------- */
/* Per image: detections (g_numberOfPoints, hesaff.cpp:68) and described keypoints (g_numberOfAffinePoints /
 * keys.size(), hesaff.cpp:103).  Either pointer may be NULL.  Arrays of n ints. */
int hesaff_result_counts(hesaff_ctx *ctx, int *n_detected, int *n_described);
/* Total described keypoints over the batch (= sum of n_described). */
int64_t hesaff_result_total(hesaff_ctx *ctx);
/* Copies all Keypoint records to host memory, image-major, reference order within an image (octave,
 * level, row, column of the initial extremum: the order keys.push_back sees).  Image i's records start
 * at the exclusive prefix sum of n_described.  capacity in records; HESAFF_ERR_CAPACITY if too small. */
int hesaff_result_keypoints(hesaff_ctx *ctx, hesaff_keypoint *out, size_t capacity);
/* Optional streamed output: when `out` is set (pinned host memory makes it asynchronous), every chunk's records are
 * copied to out[...] as soon as the chunk is finished, overlapping the work of the following chunks; a later
 * hesaff_result_keypoints(ctx, out, ...) on the same pointer then costs nothing.  NULL disables. */
int hesaff_set_host_output(hesaff_ctx *ctx, hesaff_keypoint *out, size_t capacity);
/* Same records, device pointer (valid until the next detect call). */
int hesaff_result_keypoints_device(hesaff_ctx *ctx, const hesaff_keypoint **out);
/* exportKeypoints maths (hesaff.cpp:115-125): per keypoint (u, v, a, b, c) with
 * a(x-u)^2 + 2b(x-u)(y-v) + c(y-v)^2 = 1, i.e. E = (A A^T)^-1 / (mrSize*s)^2.  5 floats per keypoint. */
int hesaff_result_ellipses(hesaff_ctx *ctx, float *out_uvabc, size_t capacity);
/* Every detection with its per-stage results (test/diagnostic use), image-major, reference order.
 * n_total receives the number of records (= sum of n_detected). */
int hesaff_result_detections(hesaff_ctx *ctx, hesaff_detection *out, size_t capacity, int64_t *n_total);

/* ---- stage access for parity tests (device -> host copies of intermediate planes) -------------- */
/* Geometry of the pyramid built by the last detect call. */
int hesaff_debug_geometry(hesaff_ctx *ctx, int *n_octaves, int *n_levels /* S+2 */);
int hesaff_debug_octave_size(hesaff_ctx *ctx, int octave, int *width, int *height);
/* kind: 0 = blur plane L[level], 1 = response plane R[level]; out receives width*height floats (dense).
 * Only valid when the batch fit in one chunk. */
int hesaff_debug_plane(hesaff_ctx *ctx, int image, int octave, int level, int kind, float *out);
/* The 41x41 affine-normalised patch (after photometric normalisation when `normalized` != 0) of every
 * described keypoint, recomputed by a diagnostic launch; out receives total*1681 floats. */
int hesaff_debug_patches(hesaff_ctx *ctx, int normalized, float *out, size_t capacity_patches);

/* Kernel launches issued by this context since creation / since the last reset (for bench.py's gpu_launches). */
int64_t hesaff_launch_count(hesaff_ctx *ctx, int reset);
/* Device time (ms, CUDA events on the launch stream) of the stages of the last detect call:
 * [0] upload+convert, [1] pyramid (blur+response), [2] nms+scan+localize+dedup, [3] affine shape,
 * [4] patch+SIFT, [5] compaction/export.  Requires hesaff_set_profiling(ctx, 1) before the call. */
int hesaff_set_profiling(hesaff_ctx *ctx, int enable);
int hesaff_stage_times_ms(hesaff_ctx *ctx, float *out6);
/* Sum of the CUDA-event durations of the individual blur+response launches (K1) of the last profiled call, and their
 * number: the per-launch figure bench.py's roofline uses. */
int hesaff_blur_time_ms(hesaff_ctx *ctx, float *total_ms, int *launches);

/* Text export of one image's keypoints in the reference's file format (hesaff.cpp:107-130,
 * README:27-44): "128\n<count>\n" then "x y a b c d1..d128" per line with ostream default
 * formatting (6 significant digits).  Writes the file; returns the number of keypoints or <0. */
int hesaff_write_sift_file(const char *path, const hesaff_keypoint *kps, size_t n, float desc_factor);

/* The same file content for image `image` of the last detect call, formatted ON THE GPU (one warp per keypoint;
 * float -> text is the exact "%g" conversion of ostream, so the bytes equal hesaff_write_sift_file's).  At GPU
 * detection rates the host's ostream loop of exportKeypoints (hesaff.cpp:107-130) would dominate the CLI.
 *   hesaff_result_sift_text: header + lines into `out` (host memory); out == NULL only reports the size in *nbytes.
 *   hesaff_export_sift_file: writes the file; returns the number of keypoints or <0. */
int hesaff_result_sift_text(hesaff_ctx *ctx, int image, char *out, size_t capacity, size_t *nbytes);
int hesaff_export_sift_file(hesaff_ctx *ctx, int image, const char *path);
/* Binary sidecar for consumers that do not want text: "HESAFFB1", u32 record size (164), u32 0, u64 count, then
 * the hesaff_keypoint records.  read: out == NULL only reports the count in *n. */
int hesaff_write_keypoints_binary(const char *path, const hesaff_keypoint *kps, size_t n);
int hesaff_read_keypoints_binary(const char *path, hesaff_keypoint *out, size_t capacity, size_t *n);
/* Diagnostic: the device float formatter on n floats; out = n slots of 16 bytes, NUL padded. */
int hesaff_debug_format_floats(hesaff_ctx *ctx, const float *in, size_t n, char *out);

#ifdef __cplusplus
}
#endif
#endif /* HESAFF_B200_H */
