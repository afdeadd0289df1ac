// hesaff_b200/csrc/common.cuh -- shared declarations of the sm_100a Hessian-Affine + SIFT path.
//
// Everything in csrc/ is compiled with -fmad=false: the reference is built without FMA contraction
// (Makefile:2 has no -march), and several discrete decisions (keypoint counts) flip if a*b+c is fused.
// FMAs appear only where written (__fmaf_rn), reproducing the operation order of cv::GaussianBlur.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/hesaff_b200.h"

#define HA_MAX_OCT 16
#define HA_MAX_LVL 14          // S+2, so S <= 12
#define HA_MAX_TAPS 33         // pyramid blur taps passed by value to the kernels
#define HA_PATCH 41            // SIFT patch side (siftdesc.h:30, affine.h:42)
#define HA_PATCH_PX (HA_PATCH * HA_PATCH)
#define HA_SIFT_ND 1245         // pixels with a non-zero SIFT mask weight: (r-20)^2 + (c-20)^2 < 400 (helpers.cpp:131-147)
#define HA_SIFT_NN 1361         // those pixels and their 4-neighbours = every patch pixel the descriptor can depend on
#define HA_SMM 19              // SMM window side (affine.h:43)
#define HA_SMM_PX (HA_SMM * HA_SMM)

#define HA_MAX_PATCH_R 1039    // largest half-width of a per-patch blur kernel (shared-memory table of the LARGE bin): sqrt(W*H) <~ 9200

#ifndef HA_BIN_TINY_MAXP
#define HA_BIN_TINY_MAXP 39    // patch+SIFT kernel bins by source-patch side P
#endif
#define HA_BIN_SMALL_MAXP 47
#define HA_BIN_MID_MAXP 63
#define HA_BIN_MID2_MAXP 79
#define HA_BIN_MEDIUM_MAXP 95

struct Taps {
   int n;
   float k[HA_MAX_TAPS];      // the taps when n <= HA_MAX_TAPS (passed to the kernels by value)
   const float *dk;           // device copy of all n taps; the generic kernels use it when n > HA_MAX_TAPS
};

// Geometry of the pyramid of one image and the offsets of every plane inside the per-image arena.
// Lives in device memory (one copy per context); kernels take a pointer.
struct Geom {
   int W, H;                 // input size
   int S;                    // numberOfScales
   int nOct;                 // octaves processed
   int border;               // PyramidParams.border
   int w[HA_MAX_OCT], h[HA_MAX_OCT], pitch[HA_MAX_OCT];   // pitch in floats, multiple of 4
   unsigned long long img_off;                             // float image (original), pitch[0]
   unsigned long long img8_off;                            // u8 copy of a gray 8-bit input (floats from the arena base), row pitch pitch8
   int pitch8;                                             // bytes, multiple of 16
   unsigned long long L_off[HA_MAX_OCT][HA_MAX_LVL];       // blur planes, floats from the image's arena base
   unsigned long long R_off[HA_MAX_OCT][HA_MAX_LVL];       // response planes
   unsigned long long arena_stride;                        // floats between consecutive images
   // candidate bitmask: one bit per pixel of levels 1..S of every octave
   int wpr[HA_MAX_OCT];                                    // 32-bit words per row
   unsigned long long mask_off[HA_MAX_OCT][HA_MAX_LVL];    // words from the image's mask base (index by level 1..S)
   unsigned long long mask_oct_off[HA_MAX_OCT + 1];        // start of each octave's words
   unsigned long long mask_stride;                         // words per image
   // dedup map (octaveMap, pyramid.cpp:189-193,226): one u32 per pixel per octave
   unsigned long long map_off[HA_MAX_OCT];
   unsigned long long map_stride;
   // scale-space constants computed on the host with the reference's float operations
   float sigma[HA_MAX_LVL];      // level sigma: curSigma when findLevelKeypoints(curSigma) runs for that level
   float finalThreshold, positiveThreshold, negativeThreshold, edgeScoreThreshold;   // pyramid.h:57-64
   float initialSigma, mrSize, convergenceThreshold;
   int maxIterations;
};

// Candidate key: img(16) | octave(4) | level(4) | row(20) | col(20)
__host__ __device__ inline unsigned long long ha_key(int img, int o, int lvl, int r, int c)
{
   return ((unsigned long long)img << 48) | ((unsigned long long)o << 44) | ((unsigned long long)lvl << 40) |
          ((unsigned long long)r << 20) | (unsigned long long)c;
}
__host__ __device__ inline void ha_unkey(unsigned long long k, int &img, int &o, int &lvl, int &r, int &c)
{
   img = (int)(k >> 48); o = (int)((k >> 44) & 15); lvl = (int)((k >> 40) & 15);
   r = (int)((k >> 20) & 0xFFFFF); c = (int)(k & 0xFFFFF);
}

// Per-candidate state, structure of arrays, indexed by candidate number (= reference order).
struct Cand {
   unsigned long long *key;
   // after localize
   float *x, *y, *s, *response;
   int *cell;                 // final r*w+c inside the octave (dedup cell)
   unsigned char *type;
   unsigned char *flags;      // HA_F_*
   // after affine
   float4 *U;                 // u11,u12,u21,u22
   float4 *A;                 // a11,a12,a21,a22 (rectified)
   int *iters;
   // after patch+SIFT
   unsigned char *desc;       // 128 per candidate
};
#define HA_F_PASS 1           // passed every localisation test except the octaveMap one
#define HA_F_DET 2            // is a detection (won its octaveMap cell)
#define HA_F_AFFINE 4         // affine shape converged
#define HA_F_DESC 8           // described (normalizeAffine succeeded)

// Work lists for the patch+SIFT kernel, binned by source patch side.
struct Bins {
   int *list[6];              // [0] SMALL, [1] MEDIUM, [2] LARGE, [3] TINY, [4] MID, [5] MID2
   int *count;                // [6]
};

// Precomputed constant tables (host, glibc libm => bit-identical to the oracle), in device memory.
struct Tables {
   const float *smm_mask;     // 19x19, computeGaussMask helpers.cpp:104-129
   const float *sift_mask;    // 41x41, computeCircularGaussMask helpers.cpp:131-147
   // The mask is zero outside a disc that never touches the patch border, so the SIFT passes run over lists:
   const uint2 *sift_disc;    // [HA_SIFT_ND] {patch index r*41+c, mask weight bits} of the pixels inside the disc, raster order
   const uint32_t *sift_out;  // [41*41 - HA_SIFT_ND] patch indices outside the disc
   const uint32_t *sift_need; // [HA_SIFT_NN] index | needed << 14 | in-disc << 15 | row << 16 | col << 24 of the disc pixels and their 4-neighbours
   const uint32_t *sift_all;  // [41*41] the same packing for every pixel (patch dumps)
   // per-patch blur kernels indexed by m = (P0-1)/2 (P0 = 2*int(mrScale)+1): taps n and offset of the
   // R+1 half kernel k[R..n-1] in `pk`
   const int *pk_n;
   const int *pk_off;
   const float *pk;
   int pk_count;
   const float *pk16;         // [48][16] the same half kernels for m <= 47 (R <= 10) at a fixed stride, [m][15] = n
};

// -------------------------------------------------------------------------------------------------
// device helpers shared by the kernels
// -------------------------------------------------------------------------------------------------
// interpolate()'s bilinear expression, helpers.cpp:235-236
__device__ __forceinline__ float ha_bilinear(float p00, float p01, float p10, float p11, float wx, float wy)
{
   return (1.0f - wy) * ((1.0f - wx) * p00 + wx * p01) + (wy) * ((1.0f - wx) * p10 + wx * p11);
}

__device__ __forceinline__ float ha_warp_sum(float v)
{
   v += __shfl_xor_sync(0xffffffffu, v, 16);
   v += __shfl_xor_sync(0xffffffffu, v, 8);
   v += __shfl_xor_sync(0xffffffffu, v, 4);
   v += __shfl_xor_sync(0xffffffffu, v, 2);
   v += __shfl_xor_sync(0xffffffffu, v, 1);
   return v;
}

// host-side launch wrappers (pyramid.cu, keypoints.cu)
struct LaunchCounter {
   long long n;
};

void ha_launch_convert_u8(const uint8_t *src, size_t row_pitch, size_t img_stride, float *arena, const Geom &g, int n,
                          cudaStream_t st, LaunchCounter &lc);
void ha_launch_convert_rgb8(const uint8_t *src, size_t row_pitch, size_t img_stride, float *dst, const Geom &g, int n,
                            cudaStream_t st, LaunchCounter &lc);
void ha_launch_convert_f32(const float *src, size_t row_pitch, size_t img_stride, float *dst, const Geom &g, int n,
                           cudaStream_t st, LaunchCounter &lc);
// one blur level: src plane -> dstL (+ response dstR, + decimated copy `half`), replicate border
int ha_launch_blur(const float *src, float *dstL, float *dstR, float *half, int W, int H, int pitch, int hW, int hH,
                   int hpitch, unsigned long long img_stride, float norm, const Taps &taps, int n, cudaStream_t st,
                   LaunchCounter &lc);
int ha_launch_blur_tma(const float *src, float *dstL, float *dstR, float *half, int W, int H, int pitch, int hW, int hH,
                       int hpitch, unsigned long long img_stride, float norm, const Taps &taps, int n, cudaStream_t st,
                       int variant = 0);
void ha_launch_hessian(const float *src, float *dst, int W, int H, int pitch, unsigned long long img_stride, float norm,
                       int n, cudaStream_t st, LaunchCounter &lc);
void ha_launch_nms(const float *arena, const Geom &g, const Geom *dg, uint32_t *mask, int n, cudaStream_t st,
                   LaunchCounter &lc);
// exclusive scan of popc(words) -> out[0..nwords], out[nwords] = total; tmp holds block sums
void ha_launch_scan_popc(const uint32_t *words, size_t nwords, uint32_t *out, uint32_t *tmp, cudaStream_t st,
                         LaunchCounter &lc);
void ha_launch_scan_flags(const unsigned char *flags, unsigned char bit, const uint32_t *count_ptr, size_t cap,
                          uint32_t *out, uint32_t *tmp, cudaStream_t st, LaunchCounter &lc);
void ha_launch_scan_u32(const uint32_t *vals, size_t n, uint32_t *out, uint32_t *tmp, cudaStream_t st, LaunchCounter &lc);
size_t ha_scan_tmp_elems(size_t n);
// export.cu: text lines of the .hesaff.sift file (length pass / write pass), and the float formatter on its own
void ha_launch_sift_text(bool write, const hesaff_keypoint *keys, const float *ell, uint32_t n, uint32_t *len,
                         const uint32_t *off, char *text, int *bad, cudaStream_t st, LaunchCounter &lc);
void ha_launch_format_floats(const float *in, size_t n, char *out, cudaStream_t st);
void ha_launch_expand(const uint32_t *mask, const uint32_t *woff, const Geom *dg, size_t nwords, Cand cand, uint32_t cap,
                      int *overflow, cudaStream_t st, LaunchCounter &lc);
void ha_launch_localize(const float *arena, const Geom *dg, Cand cand, const uint32_t *count, uint32_t cap,
                        uint32_t *map, cudaStream_t st, LaunchCounter &lc);
void ha_launch_affine(const float *arena, const Geom *dg, Tables tb, Cand cand, const uint32_t *count, uint32_t cap,
                      const uint32_t *map, int *n_det, Bins bins, int *work_counter, cudaStream_t st, LaunchCounter &lc);
void ha_launch_describe(const float *arena, const Geom *dg, const Geom &hg, Tables tb, Cand cand, Bins bins, int *work_counters,
                        float *scratch, size_t scratch_per_cta, int large_ctas, int maxP, int src_u8, float *patch_dump,
                        int dump_normalized, const uint32_t *dump_index, cudaStream_t st, LaunchCounter &lc,
                        cudaStream_t aux = nullptr, cudaEvent_t ev_fork = nullptr, cudaEvent_t ev_join = nullptr);
void ha_launch_compact(Cand cand, const uint32_t *count, uint32_t cap, const uint32_t *desc_off, const Geom *dg,
                       hesaff_keypoint *out, float *ellipses, int *n_desc, const uint32_t *out_base, uint32_t keys_cap,
                       int *overflow, cudaStream_t st, LaunchCounter &lc);
void ha_launch_export_detections(Cand cand, const uint32_t *count, uint32_t cap, const uint32_t *det_off,
                                 const Geom *dg, hesaff_detection *out, cudaStream_t st, LaunchCounter &lc);
int ha_describe_smem_bytes(int bin);
void ha_launch_match(const hesaff_keypoint *query, uint32_t nq, const hesaff_keypoint *db, uint32_t ndb, int32_t *best_index,
                     uint32_t *best_d2, uint32_t *second_d2, cudaStream_t st);
size_t ha_describe_scratch_floats(int maxP);
