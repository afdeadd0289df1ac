// hesaff_b200/csrc/api.cu -- the C-ABI of include/hesaff_b200.h: context, device-memory plan, launch
// schedule of the detect -> affine -> describe path (detectPyramidKeypoints, pyramid.cpp:261-292, with the
// two callbacks of hesaff.cpp:66-105 turned into batch stages).
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include <algorithm>
#include <fstream>
#include <sstream>
#include "common.cuh"
#include <nvtx3/nvToolsExt.h>   // header-only; ranges cost nothing unless a profiler is attached

static thread_local std::string g_err;
static int fail(int code, const std::string &msg)
{
   g_err = msg;
   return code;
}
#define CK(call)                                                                                           \
   do {                                                                                                    \
      cudaError_t e_ = (call);                                                                             \
      if (e_ != cudaSuccess)                                                                               \
         return fail(HESAFF_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));                  \
   } while (0)

// Per-chunk working set.  Two lanes (stream + buffers each) let the upload / pyramid / affine kernels of chunk k+1
// run under the describe kernels of chunk k, and the D2H of finished records run under both.
struct Lane {
   cudaStream_t stream, aux;
   cudaEvent_t ev_fork, ev_join;
   float *arena;
   uint8_t *stage_u8;
   uint32_t *mask;
   uint32_t *woff;             // scan of popc(mask): nwords+1
   uint32_t *scan_tmp;
   uint32_t *map;
   Cand cand;
   uint32_t *det_off, *desc_off;   // cand_cap+1 each
   Bins bins;
   int *counters;              // [0] affine work, [1..6] describe work (one queue per bin)
   float *scratch;
   cudaEvent_t ev[8];
   std::vector<cudaEvent_t> blur_ev;   // profiling: event pairs around every k_blur launch
   cudaEvent_t done;           // recorded after the chunk's k_add_total
   cudaEvent_t front_done;     // recorded after the chunk's affine-shape kernel (end of the front end)
   uint32_t *h_total;          // pinned: [0] described keypoints of the last chunk on this lane, [1] base offset
   int pending_chunk;          // chunk index whose results are not yet copied to the host output (-1: none)
};

struct hesaff_ctx {
   hesaff_params par;
   int device;
   int max_w, max_h;
   int chunk;                  // images resident at once per lane
   int n_lanes;
   uint32_t cand_cap;          // candidate pool of one chunk
   int max_cand_per_image;
   cudaStream_t stream;        // own stream (control / results)
   cudaStream_t copy_stream;   // D2H of finished chunks
   cudaEvent_t ev_start;
   Lane lane[2];
   Geom geom;                  // host copy of the current geometry (W,H of the last call)
   Geom *d_geom;
   Taps taps0;                 // first blur (pyramid.cpp:278-279)
   Taps taps[HA_MAX_LVL];      // incremental blurs, level 1..S+1
   float norm[HA_MAX_LVL];     // sigma_l^2 passed to hessianResponse
   float lvl_sigma[HA_MAX_LVL];   // curSigma of each level (findLevelKeypoints(curSigma), pyramid.cpp:248)
   size_t mask_words_cap, map_elems_cap, scan_tmp_elems, stage_bytes;
   size_t scratch_per_cta; int large_ctas; int maxP;
   Tables tables;
   std::vector<void *> table_allocs;
   // results of the last call
   int n_images;
   int *d_ndet, *d_ndesc; size_t counts_cap;
   int *d_overflow;
   std::vector<int> h_ndet, h_ndesc;
   hesaff_keypoint *d_keys; float *d_ell; size_t keys_cap;
   uint32_t *d_out_base;       // running total of described keypoints over chunks
   hesaff_detection *d_dets; size_t dets_cap;
   char *d_text; size_t text_cap;            // GPU-formatted .hesaff.sift lines of one image
   uint32_t *d_tlen, *d_toff; size_t tlen_cap;
   int *d_bad;
   hesaff_keypoint *host_out; size_t host_out_cap; bool host_out_filled;   // optional streamed host output
   int64_t total_desc, total_det;
   int last_chunks;
   int last_u8;                // the last call's input was gray u8 (the u8 image copy in the arena is valid)
   bool have_result;
   LaunchCounter lc;
   // profiling (forces single-lane, serialised execution so that stage times are clean)
   bool profiling;
   float stage_ms[6];
   float blur_ms;              // sum over the k_blur launches of the last call (profiling mode)
   int blur_launches;
};

extern "C" int hesaff_abi_version(void) { return HESAFF_B200_ABI_VERSION; }
extern "C" const char *hesaff_last_error(void) { return g_err.c_str(); }

extern "C" int hesaff_params_default(hesaff_params *p)
{
   if (!p) return fail(HESAFF_ERR_INVALID, "params is NULL");
   p->threshold = 16.0f / 3.0f;              // hesaff.cpp:30
   p->max_iter = 16;                         // hesaff.cpp:31
   p->desc_factor = 3.0f * sqrtf(3.0f);      // hesaff.cpp:32
   p->patch_size = 41;                       // hesaff.cpp:33
   p->verbose = 0;                           // hesaff.cpp:34
   p->number_of_scales = 3;                  // pyramid.h:35
   p->initial_sigma = 1.6f;                  // pyramid.h:36
   p->edge_eigenvalue_ratio = 10.0f;         // pyramid.h:38
   p->border = 5;                            // pyramid.h:39
   p->convergence_threshold = 0.05f;         // affine.h:41
   p->smm_window_size = 19;                  // affine.h:43
   p->max_octaves = 0;
   return HESAFF_OK;
}

// ---- host-side constant tables (glibc libm, same expressions as the reference) ---------------------
static int blur_size(float sigma)
{
   int size = (int)(2.0 * 3.0 * sigma + 1.0);   // helpers.cpp:286
   if (size % 2 == 0) size++;
   return size;
}

// cv::getGaussianKernel(n, sigma, CV_32F) as OpenCV 4.x evaluates it (pinned in tests/test_oracle_blur.py)
static void gauss_kernel(int n, double sigma, std::vector<float> &k)
{
   const int R = (n - 1) / 2;
   const double scale2X = -0.5 / (sigma * sigma);
   std::vector<double> v(R + 1);
   double sum = 0;
   for (int i = 0; i < R; i++) { double x = (double)(i - R); v[i] = exp(scale2X * x * x); sum += v[i]; }
   v[R] = 1.0;
   sum = sum * 2 + 1.0;
   const double m = 1.0 / sum;
   k.resize(n);
   for (int i = 0; i <= R; i++) k[i] = k[n - 1 - i] = (float)(v[i] * m);
}

// computeGaussMask, helpers.cpp:104-129
static void smm_mask(float *mask, int size)
{
   int halfSize = size >> 1;
   float scale = float(halfSize) / 3.0f;
   float scale2 = -2.0f * scale * scale;
   std::vector<float> tmp(halfSize + 1);
   for (int i = 0; i <= halfSize; i++) tmp[i] = expf((float(i * i) / scale2));
   int endSize = int(ceilf(scale * 5.0f) - halfSize);
   for (int i = 1; i < endSize; i++) tmp[halfSize - i] += expf((float((i + halfSize) * (i + halfSize)) / scale2));
   for (int i = 0; i <= halfSize; i++)
      for (int j = 0; j <= halfSize; j++) {
         float v = tmp[i] * tmp[j];
         mask[(i + halfSize) * size + (-j + halfSize)] = v;
         mask[(-i + halfSize) * size + (j + halfSize)] = v;
         mask[(i + halfSize) * size + (j + halfSize)] = v;
         mask[(-i + halfSize) * size + (-j + halfSize)] = v;
      }
}

// computeCircularGaussMask, helpers.cpp:131-147
static void sift_mask(float *mask, int size)
{
   int halfSize = size >> 1;
   float r2 = float(halfSize * halfSize);
   float sigma2 = 0.9f * r2;
   float *mp = mask;
   for (int i = 0; i < size; i++)
      for (int j = 0; j < size; j++) {
         float disq = float((i - halfSize) * (i - halfSize) + (j - halfSize) * (j - halfSize));
         *mp++ = (disq < r2) ? expf(-disq / sigma2) : 0;
      }
}

// Reorders a pixel list so that every aligned group of 32 entries (= the pixels one warp touches in one round) holds 32
// different patch indices mod 32, lane l and lane l + 16 in bank pairs 16 apart: the scattered shared-memory accesses of
// the SIFT passes (patch[idx +- 1], patch[idx +- 41], float2 v01[idx]) are then free of bank conflicts.  The passes are
// order independent per pixel.
template <typename T, typename F> static void bank_order(std::vector<T> &v, F index_of)
{
   std::vector<std::vector<T>> bucket(32);
   for (const T &e : v) bucket[index_of(e) & 31].push_back(e);
   for (auto &b : bucket) std::reverse(b.begin(), b.end());   // pop_back takes them in raster order
   std::vector<T> out;
   out.reserve(v.size());
   size_t left = v.size();
   while (left) {
      // the last group may have to take two of a residue; until then, one per residue, fullest buckets first
      std::vector<int> order(32);
      for (int i = 0; i < 32; i++) order[i] = i;
      std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return bucket[a].size() > bucket[b].size(); });
      std::vector<int> lane_res(32, -1);
      size_t take = std::min<size_t>(32, left);
      std::vector<int> chosen;
      for (int i = 0; i < 32 && chosen.size() < take; i++)
         if (!bucket[order[i]].empty()) chosen.push_back(order[i]);
      std::sort(chosen.begin(), chosen.end());
      // residue r goes to lane r when free (lanes 0..15 <-> residues 0..15, lanes 16..31 <-> 16..31)
      std::vector<T> group;
      std::vector<char> used(32, 0);
      std::vector<int> lane_of(chosen.size(), -1);
      for (size_t k = 0; k < chosen.size(); k++)
         if ((size_t)chosen[k] < take && !used[chosen[k]]) { lane_of[k] = chosen[k]; used[chosen[k]] = 1; }
      for (size_t k = 0; k < chosen.size(); k++)
         if (lane_of[k] < 0)
            for (size_t l = 0; l < take; l++)
               if (!used[l]) { lane_of[k] = (int)l; used[l] = 1; break; }
      group.resize(take);
      for (size_t k = 0; k < chosen.size(); k++) { group[lane_of[k]] = bucket[chosen[k]].back(); bucket[chosen[k]].pop_back(); }
      // fewer non-empty buckets than slots: fill the free lanes from the fullest buckets
      for (size_t l = 0; l < take; l++)
         if (!used[l]) {
            int best = 0;
            for (int i = 1; i < 32; i++) if (bucket[i].size() > bucket[best].size()) best = i;
            group[l] = bucket[best].back();
            bucket[best].pop_back();
         }
      for (const T &e : group) out.push_back(e);
      left -= group.size();
   }
   v.swap(out);
}

template <typename T> static int upload(hesaff_ctx *c, const std::vector<T> &h, const T **d)
{
   void *p = nullptr;
   CK(cudaMalloc(&p, std::max<size_t>(sizeof(T) * h.size(), 16)));
   c->table_allocs.push_back(p);
   CK(cudaMemcpy(p, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice));
   *d = (const T *)p;
   return HESAFF_OK;
}

static int make_taps(hesaff_ctx *c, float sigma, Taps &t)
{
   const int n = blur_size(sigma);
   std::vector<float> k;
   gauss_kernel(n, (double)sigma, k);
   t.n = n;
   t.dk = nullptr;
   memset(t.k, 0, sizeof(t.k));
   if (n <= HA_MAX_TAPS) {
      for (int i = 0; i < n; i++) t.k[i] = k[i];
      return HESAFF_OK;
   }
   // more taps than the tiled kernels take by value: the generic kernels read them from device memory
   return upload(c, k, &t.dk);
}

static int build_tables(hesaff_ctx *c)
{
   std::vector<float> m19(HA_SMM_PX), m41(HA_PATCH_PX);
   smm_mask(m19.data(), HA_SMM);
   sift_mask(m41.data(), HA_PATCH);
   int rc;
   if ((rc = upload(c, m19, &c->tables.smm_mask))) return rc;
   if ((rc = upload(c, m41, &c->tables.sift_mask))) return rc;
   {  // pixel lists of the SIFT passes (see Tables)
      std::vector<uint2> disc;
      std::vector<uint32_t> outside, need, all;
      std::vector<char> needed(HA_PATCH_PX, 0);
      for (int r = 0; r < HA_PATCH; r++)
         for (int q = 0; q < HA_PATCH; q++) {
            const int t = r * HA_PATCH + q;
            if (m41[t] > 0) {
               uint32_t bits;
               memcpy(&bits, &m41[t], 4);
               disc.push_back(make_uint2((unsigned)t, bits));
               if (r < 1 || q < 1 || r > HA_PATCH - 2 || q > HA_PATCH - 2) return fail(HESAFF_ERR_INVALID, "SIFT mask touches the patch border");
               needed[t] = needed[t - 1] = needed[t + 1] = needed[t - HA_PATCH] = needed[t + HA_PATCH] = 1;
            } else outside.push_back((uint32_t)t);
         }
      for (int t = 0; t < HA_PATCH_PX; t++) {
         // patch index | needed << 14 | inside the disc << 15 | row << 16 | column << 24
         const uint32_t w = (uint32_t)t | (needed[t] ? 0x4000u : 0u) | (m41[t] > 0 ? 0x8000u : 0u) |
                            ((uint32_t)(t / HA_PATCH) << 16) | ((uint32_t)(t % HA_PATCH) << 24);
         all.push_back(w);
         if (needed[t]) need.push_back(w);
      }
      if (disc.size() != HA_SIFT_ND || need.size() != HA_SIFT_NN) return fail(HESAFF_ERR_INVALID, "SIFT mask geometry differs from HA_SIFT_ND / HA_SIFT_NN");
      bank_order(disc, [](const uint2 &e) { return e.x; });
      bank_order(outside, [](uint32_t e) { return e; });
      // (the need / all lists stay in raster order: their pass reads neighbouring taps of the blurred patch)
      if ((rc = upload(c, disc, &c->tables.sift_disc))) return rc;
      if ((rc = upload(c, outside, &c->tables.sift_out))) return rc;
      if ((rc = upload(c, need, &c->tables.sift_need))) return rc;
      if ((rc = upload(c, all, &c->tables.sift_all))) return rc;
   }
   // per-patch blur kernels, indexed by m = (P0-1)/2.  interpolateCheckBorders (helpers.cpp:191-207) only accepts a
   // keypoint whose 41x41 footprint (+-20*its*A, det A = 1) lies inside the image: 40*its*a11 < W and 40*its*a22 < H,
   // hence its < sqrt(W*H)/40 and P0 = 41*its < 1.025*sqrt(W*H).
   const int maxP0 = c->maxP;
   const int count = maxP0 / 2 + 2;
   std::vector<int> pn(count, 1), poff(count, 0);
   std::vector<float> pk;
   for (int m = 0; m < count; m++) {
      const int P0 = 2 * m + 1;
      const float its = float(P0) / float(HA_PATCH);     // affine.cpp:109
      const float sigma = 1.5f * its;                     // affine.cpp:129
      const int n = blur_size(sigma);
      std::vector<float> k;
      gauss_kernel(n, (double)sigma, k);
      pn[m] = n;
      poff[m] = (int)pk.size();
      for (int i = n / 2; i < n; i++) pk.push_back(k[i]);
      if (n / 2 > HA_MAX_PATCH_R) return fail(HESAFF_ERR_INVALID, "image too large: the per-patch blur would need more than 2079 taps (sqrt(W*H) must stay below ~9200)");
   }
   c->tables.pk_count = count;
   {  // fixed-stride copy for the shared-memory bins (P0 <= 93: m <= 46, R <= 10)
      std::vector<float> pk16(48 * 16, 0.f);
      for (int m = 0; m < 48 && m < count; m++) {
         const int R = pn[m] / 2;
         if (R > 14) break;
         for (int i = 0; i <= R; i++) pk16[m * 16 + i] = pk[poff[m] + i];
         pk16[m * 16 + 15] = (float)pn[m];
      }
      if ((rc = upload(c, pk16, &c->tables.pk16))) return rc;
   }
   if ((rc = upload(c, pn, &c->tables.pk_n))) return rc;
   if ((rc = upload(c, poff, &c->tables.pk_off))) return rc;
   if ((rc = upload(c, pk, &c->tables.pk))) return rc;
   return HESAFF_OK;
}

// ---- geometry / memory plan ------------------------------------------------------------------------
static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static int plan_geometry(const hesaff_params &p, int W, int H, Geom &g)
{
   memset(&g, 0, sizeof(g));
   g.W = W; g.H = H; g.S = p.number_of_scales; g.border = p.border;
   // octave loop of detectPyramidKeypoints, pyramid.cpp:283-291
   const int minSize = 2 * p.border + 2;
   int rows = H, cols = W, o = 0;
   while (rows > minSize && cols > minSize) {
      if (p.max_octaves > 0 && o >= p.max_octaves) break;
      if (o >= HA_MAX_OCT) break;
      g.w[o] = cols; g.h[o] = rows; g.pitch[o] = (int)align_up(cols, 4);
      rows /= 2; cols /= 2; o++;       // halfImage, helpers.cpp:333
   }
   g.nOct = o;
   size_t off = 0;
   g.img_off = off;
   off += align_up((size_t)H * align_up(W, 4), 64);
   if (g.nOct == 0) g.pitch[0] = (int)align_up(W, 4);
   g.pitch8 = (int)align_up(W, 16);
   g.img8_off = off;
   off += align_up(((size_t)H * g.pitch8 + 3) / 4, 64);
   for (o = 0; o < g.nOct; o++)
      for (int l = 0; l < g.S + 2; l++) {
         g.L_off[o][l] = off; off += align_up((size_t)g.h[o] * g.pitch[o], 64);
         g.R_off[o][l] = off; off += align_up((size_t)g.h[o] * g.pitch[o], 64);
      }
   g.arena_stride = off;
   size_t mo = 0, po = 0;
   for (o = 0; o < g.nOct; o++) {
      g.wpr[o] = (g.w[o] + 31) / 32;
      g.mask_oct_off[o] = mo;
      for (int l = 1; l <= g.S; l++) { g.mask_off[o][l] = mo; mo += (size_t)g.h[o] * g.wpr[o]; }
      g.map_off[o] = po; po += (size_t)g.h[o] * g.w[o];
   }
   g.mask_oct_off[g.nOct] = mo;
   g.mask_stride = mo;
   g.map_stride = po;
   // thresholds, pyramid.h:57-64
   g.edgeScoreThreshold = (p.edge_eigenvalue_ratio + 1.0f) * (p.edge_eigenvalue_ratio + 1.0f) / p.edge_eigenvalue_ratio;
   g.finalThreshold = p.threshold * p.threshold;
   g.positiveThreshold = (float)(0.8 * g.finalThreshold);
   g.negativeThreshold = -g.positiveThreshold;
   g.initialSigma = p.initial_sigma; g.mrSize = p.desc_factor; g.convergenceThreshold = p.convergence_threshold;
   g.maxIterations = p.max_iter;
   return HESAFF_OK;
}

static size_t per_image_bytes(const Geom &g, int max_cand)
{
   // arena + host-input staging (up to 4 bytes per pixel) + candidate mask and its scan + octaveMap + candidate pool (SoA state,
   // descriptors, work lists) + the output records of the call (hesaff_keypoint + ellipse for half the candidates)
   return g.arena_stride * 4 + (size_t)g.W * g.H * 4 + g.mask_stride * 8 + g.map_stride * 4 + (size_t)max_cand * 260 +
          (size_t)(max_cand / 2) * (sizeof(hesaff_keypoint) + 20);
}

template <typename T> static int dmalloc(T **p, size_t n)
{
   void *q = nullptr;
   CK(cudaMalloc(&q, std::max<size_t>(n * sizeof(T), 256)));
   *p = (T *)q;
   return HESAFF_OK;
}

static int alloc_lane(hesaff_ctx *c, Lane &L, const Geom &g)
{
   int rc;
   L = Lane();
   L.pending_chunk = -1;
   CK(cudaStreamCreateWithFlags(&L.stream, cudaStreamNonBlocking));
   CK(cudaStreamCreateWithFlags(&L.aux, cudaStreamNonBlocking));
   CK(cudaEventCreateWithFlags(&L.ev_fork, cudaEventDisableTiming));
   CK(cudaEventCreateWithFlags(&L.ev_join, cudaEventDisableTiming));
   for (int i = 0; i < 8; i++) CK(cudaEventCreate(&L.ev[i]));
   CK(cudaEventCreateWithFlags(&L.done, cudaEventDisableTiming));
   CK(cudaEventCreateWithFlags(&L.front_done, cudaEventDisableTiming));
   CK(cudaMallocHost((void **)&L.h_total, 4 * sizeof(uint32_t)));
   const int chunk = c->chunk;
   if ((rc = dmalloc(&L.arena, g.arena_stride * chunk))) return rc;
   if ((rc = dmalloc(&L.stage_u8, c->stage_bytes))) return rc;
   if ((rc = dmalloc(&L.mask, c->mask_words_cap + 1))) return rc;
   if ((rc = dmalloc(&L.woff, c->mask_words_cap + 2))) return rc;
   if ((rc = dmalloc(&L.scan_tmp, c->scan_tmp_elems))) return rc;
   if ((rc = dmalloc(&L.map, c->map_elems_cap + 1))) return rc;
   const size_t cc = c->cand_cap;
   if ((rc = dmalloc(&L.cand.key, cc))) return rc;
   if ((rc = dmalloc(&L.cand.x, cc)) || (rc = dmalloc(&L.cand.y, cc)) || (rc = dmalloc(&L.cand.s, cc)) ||
       (rc = dmalloc(&L.cand.response, cc)) || (rc = dmalloc(&L.cand.cell, cc)) || (rc = dmalloc(&L.cand.type, cc)) ||
       (rc = dmalloc(&L.cand.flags, cc)) || (rc = dmalloc(&L.cand.U, cc)) || (rc = dmalloc(&L.cand.A, cc)) ||
       (rc = dmalloc(&L.cand.iters, cc)) || (rc = dmalloc(&L.cand.desc, cc * 128)))
      return rc;
   if ((rc = dmalloc(&L.det_off, cc + 2)) || (rc = dmalloc(&L.desc_off, cc + 2))) return rc;
   for (int b = 0; b < 6; b++)
      if ((rc = dmalloc(&L.bins.list[b], cc))) return rc;
   if ((rc = dmalloc(&L.bins.count, 8))) return rc;
   if ((rc = dmalloc(&L.counters, 8))) return rc;
   if ((rc = dmalloc(&L.scratch, c->scratch_per_cta * c->large_ctas))) return rc;
   return HESAFF_OK;
}

static void free_lane(Lane &L)
{
   if (L.stream) cudaStreamSynchronize(L.stream);
   void *ptrs[] = {L.arena, L.stage_u8, L.mask, L.woff, L.scan_tmp, L.map, L.cand.key, L.cand.x, L.cand.y, L.cand.s,
                   L.cand.response, L.cand.cell, L.cand.type, L.cand.flags, L.cand.U, L.cand.A, L.cand.iters, L.cand.desc,
                   L.det_off, L.desc_off, L.bins.list[0], L.bins.list[1], L.bins.list[2], L.bins.list[3], L.bins.list[4], L.bins.list[5], L.bins.count, L.counters, L.scratch};
   for (void *p : ptrs) if (p) cudaFree(p);
   for (int i = 0; i < 8; i++) if (L.ev[i]) cudaEventDestroy(L.ev[i]);
   if (L.done) cudaEventDestroy(L.done);
   if (L.front_done) cudaEventDestroy(L.front_done);
   if (L.ev_fork) cudaEventDestroy(L.ev_fork);
   if (L.ev_join) cudaEventDestroy(L.ev_join);
   if (L.aux) cudaStreamDestroy(L.aux);
   if (L.h_total) cudaFreeHost(L.h_total);
   if (L.stream) cudaStreamDestroy(L.stream);
   for (cudaEvent_t e : L.blur_ev) cudaEventDestroy(e);
   L = Lane();
}

extern "C" int hesaff_create(hesaff_ctx **out, const hesaff_params *p, int device, int max_width, int max_height,
                             int max_batch, int max_candidates_per_image)
{
   if (!out || !p) return fail(HESAFF_ERR_INVALID, "NULL argument");
   *out = nullptr;
   if (p->patch_size != HA_PATCH) return fail(HESAFF_ERR_INVALID, "only patch_size == 41 is supported");
   if (p->smm_window_size != HA_SMM) return fail(HESAFF_ERR_INVALID, "only smm_window_size == 19 is supported");
   if (p->number_of_scales < 1 || p->number_of_scales + 2 > HA_MAX_LVL) return fail(HESAFF_ERR_INVALID, "number_of_scales out of range [1,12]");
   if (p->border < 2) return fail(HESAFF_ERR_INVALID, "border must be >= 2 (pyramid.cpp:208)");
   if (p->max_iter < 1) return fail(HESAFF_ERR_INVALID, "max_iter must be >= 1");
   if (max_width < 1 || max_height < 1 || max_width >= (1 << 20) || max_height > 65535)
      return fail(HESAFF_ERR_INVALID, "bad max size (width < 2^20, height <= 65535: rows index the y dimension of the launch grids)");
   int ndev = 0;
   if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
      return fail(HESAFF_ERR_CUDA, "no CUDA device: this library has no CPU fallback");
   if (device < 0 || device >= ndev) return fail(HESAFF_ERR_INVALID, "bad device index");
   CK(cudaSetDevice(device));
   cudaDeviceProp prop;
   CK(cudaGetDeviceProperties(&prop, device));
   if (prop.major < 10) return fail(HESAFF_ERR_CUDA, "device is not sm_100 class (kernels are compiled for sm_100a only)");

   hesaff_ctx *c = new hesaff_ctx();
   c->par = *p; c->device = device; c->max_w = max_width; c->max_h = max_height;
   c->have_result = false; c->profiling = false; c->lc.n = 0;
   c->d_dets = nullptr; c->dets_cap = 0; c->d_text = nullptr; c->text_cap = 0; c->d_tlen = c->d_toff = nullptr; c->tlen_cap = 0;
   c->d_bad = nullptr; c->d_keys = nullptr; c->d_ell = nullptr; c->keys_cap = 0;
   c->d_ndet = c->d_ndesc = nullptr; c->counts_cap = 0; c->d_overflow = nullptr; c->d_out_base = nullptr; c->d_geom = nullptr;
   c->host_out = nullptr; c->host_out_cap = 0; c->host_out_filled = false;
   c->stream = c->copy_stream = nullptr; c->ev_start = nullptr; c->n_lanes = 0;
   memset(c->stage_ms, 0, sizeof(c->stage_ms));
   *out = c;   // so that a failed create can still be destroyed

   // scale-space constants with the reference's float operations (pyramid.cpp:227-240,257,278)
   {
      float curSigma0 = 0.5f;
      c->taps0.n = 0;
      if (p->initial_sigma > curSigma0) {
         float sigma = sqrtf(p->initial_sigma * p->initial_sigma - curSigma0 * curSigma0);
         { int rc = make_taps(c, sigma, c->taps0); if (rc) return rc; }
      }
      const int S = p->number_of_scales;
      float sigmaStep = powf(2.0f, 1.0f / (float)S);
      float curSigma = p->initial_sigma;
      c->norm[0] = curSigma * curSigma;
      c->lvl_sigma[0] = curSigma;
      for (int i = 1; i < S + 2; i++) {
         float sigma = curSigma * sqrtf(sigmaStep * sigmaStep - 1.0f);
         { int rc = make_taps(c, sigma, c->taps[i]); if (rc) return rc; }
         sigma = curSigma * sigmaStep;
         c->norm[i] = sigma * sigma;
         curSigma *= sigmaStep;
         c->lvl_sigma[i] = curSigma;
      }
   }
   Geom g;
   plan_geometry(*p, max_width, max_height, g);
   c->max_cand_per_image = max_candidates_per_image > 0 ? max_candidates_per_image
                                                        : std::max(4096, (int)(((size_t)max_width * max_height) / 10));
   size_t free_b = 0, total_b = 0;
   CK(cudaMemGetInfo(&free_b, &total_b));
   const size_t pib = per_image_bytes(g, c->max_cand_per_image);
   // two lanes of `chunk` images each; a batch that fits one chunk (max_batch given and small) still gets 2 lanes of
   // that size so that consecutive calls need no reallocation
   int chunk = max_batch;
   size_t chunk_cap = 128;   // large chunks amortise the launch-bound small octaves (pyramid stage: 50 -> 40 ms per 1024 x 1080p)
   if (const char *e = getenv("HESAFF_CHUNK")) chunk_cap = (size_t)std::max(1, atoi(e));   // tuning knob (bench experiments)
   if (chunk <= 0) chunk = (int)std::min<size_t>(chunk_cap, std::max<size_t>(1, (size_t)(free_b * 0.45) / (2 * pib)));
   // work-list entries carry the candidate index in 26 bits (describe.cu): at most 2^26 candidates per chunk
   if ((size_t)c->max_cand_per_image > (1ull << 26)) return fail(HESAFF_ERR_INVALID, "candidate pool of one image exceeds 2^26");
   chunk = (int)std::min<size_t>((size_t)chunk, (1ull << 26) / (size_t)c->max_cand_per_image);
   chunk = std::min(chunk, 65535);    // the image index is the z dimension of the launch grids and 16 bits of the candidate key
   c->n_lanes = 2;
   if ((size_t)chunk * pib * 2 > free_b * 0.85) c->n_lanes = 1;
   if ((size_t)chunk * pib * c->n_lanes > free_b * 0.9) return fail(HESAFF_ERR_CUDA, "not enough device memory for max_batch images of this size");
   c->chunk = chunk;

   c->cand_cap = (uint32_t)((size_t)chunk * c->max_cand_per_image);

   CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
   CK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
   CK(cudaEventCreateWithFlags(&c->ev_start, cudaEventDisableTiming));
   int rc;
   c->maxP = (int)(1.025 * sqrt((double)max_width * (double)max_height)) + 10;   // largest source-patch side P = P0 + 2
   if ((rc = build_tables(c))) return rc;
   if ((rc = dmalloc(&c->d_geom, 1))) return rc;
   if ((rc = dmalloc(&c->d_out_base, 2)) || (rc = dmalloc(&c->d_overflow, 2))) return rc;
   c->stage_bytes = (size_t)max_width * max_height * 4 * chunk;   // u8 or f32 host input staging
   c->mask_words_cap = g.mask_stride * chunk;
   c->scan_tmp_elems = std::max(ha_scan_tmp_elems(c->mask_words_cap), ha_scan_tmp_elems(c->cand_cap)) + 8;
   c->map_elems_cap = g.map_stride * chunk;
   c->large_ctas = 148 * 4;
   c->scratch_per_cta = align_up(ha_describe_scratch_floats(c->maxP), 64);
   for (int l = 0; l < c->n_lanes; l++)
      if ((rc = alloc_lane(c, c->lane[l], g))) return rc;
   return HESAFF_OK;
}

extern "C" int hesaff_destroy(hesaff_ctx *c)
{
   if (!c) return HESAFF_OK;
   cudaSetDevice(c->device);
   if (c->stream) cudaStreamSynchronize(c->stream);
   if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
   for (int l = 0; l < 2; l++) free_lane(c->lane[l]);
   void *ptrs[] = {c->d_geom, c->d_out_base, c->d_overflow, c->d_ndet, c->d_ndesc, c->d_keys, c->d_ell, c->d_dets,
                   c->d_text, c->d_tlen, c->d_toff, c->d_bad};
   for (void *p : ptrs) if (p) cudaFree(p);
   for (void *p : c->table_allocs) cudaFree(p);
   if (c->ev_start) cudaEventDestroy(c->ev_start);
   if (c->stream) cudaStreamDestroy(c->stream);
   if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
   delete c;
   return HESAFF_OK;
}

// Streamed host output: when set (pinned memory recommended), every chunk's Keypoint records are copied to
// `out` as soon as the chunk is finished, overlapping the next chunks; hesaff_result_keypoints on the same pointer is
// then free.  NULL disables.
extern "C" int hesaff_set_host_output(hesaff_ctx *c, hesaff_keypoint *out, size_t capacity)
{
   if (!c) return fail(HESAFF_ERR_INVALID, "ctx is NULL");
   c->host_out = out; c->host_out_cap = out ? capacity : 0; c->host_out_filled = false;
   return HESAFF_OK;
}

// ---- the hot path ------------------------------------------------------------------------------------
__global__ void k_add_total(uint32_t *base, const uint32_t *add, uint32_t *report)
{
   report[1] = *base;     // base offset of this chunk's records
   report[0] = *add;      // number of records of this chunk
   *base += *add;
}

// copies the records of the chunk that last ran on lane L to the streamed host output (if any)
static int flush_lane_output(hesaff_ctx *c, Lane &L)
{
   if (L.pending_chunk < 0) return HESAFF_OK;
   L.pending_chunk = -1;
   CK(cudaEventSynchronize(L.done));
   if (!c->host_out) return HESAFF_OK;
   const size_t cnt = L.h_total[0], base = L.h_total[1];
   if (base + cnt > c->host_out_cap) return fail(HESAFF_ERR_CAPACITY, "streamed host output buffer too small");
   if (cnt) CK(cudaMemcpyAsync(c->host_out + base, c->d_keys + base, cnt * sizeof(hesaff_keypoint), cudaMemcpyDeviceToHost, c->copy_stream));
   return HESAFF_OK;
}

enum InFmt { IN_U8 = 0, IN_F32 = 1, IN_RGB8 = 2 };

// NVTX ranges mark the host-side enqueue of every stage of a chunk (nsys / ncu --nvtx correlate them with the kernels);
// the guard keeps the range stack balanced on the error returns
struct NvtxStages {
   int depth = 0;
   explicit NvtxStages(const char *outer) { nvtxRangePushA(outer); depth = 1; }
   void next(const char *name) { if (depth > 1) nvtxRangePop(); else depth = 2; nvtxRangePushA(name); }
   ~NvtxStages() { while (depth-- > 0) nvtxRangePop(); }
};

// `ptrs` (optional, host input only): image i starts at ptrs[i] instead of images + i * img_stride
static int detect_impl(hesaff_ctx *c, const void *images, InFmt fmt, int n, int W, int H, size_t row_pitch,
                       size_t img_stride, int on_device, void *stream_, const void *const *ptrs = nullptr)
{
   if (!c) return fail(HESAFF_ERR_INVALID, "ctx is NULL");
   if ((!images && !ptrs) || n < 0 || W < 1 || H < 1) return fail(HESAFF_ERR_INVALID, "bad image arguments");
   if (ptrs && on_device) return fail(HESAFF_ERR_INVALID, "per-image pointers are for host input");
   if (W > c->max_w || H > c->max_h) return fail(HESAFF_ERR_INVALID, "image larger than the context was created for");
   const size_t esz = fmt == IN_U8 ? 1 : (fmt == IN_RGB8 ? 3 : 4);
   if (row_pitch < (size_t)W * esz || img_stride < row_pitch * (size_t)(H - 1) + (size_t)W * esz)
      return fail(HESAFF_ERR_INVALID, "row pitch / image stride too small");
   CK(cudaSetDevice(c->device));
   cudaStream_t ust = stream_ ? (cudaStream_t)stream_ : c->stream;
   if (on_device && !stream_) {
      // stream == NULL means the context's own (non-blocking) stream, NOT the default stream: order it after whatever
      // the caller queued on the legacy default stream (the usual producer of a device-resident input)
      CK(cudaEventRecord(c->ev_start, cudaStreamLegacy));
      CK(cudaStreamWaitEvent(ust, c->ev_start, 0));
   }
   c->have_result = false;
   c->host_out_filled = false;

   Geom &g = c->geom;
   plan_geometry(c->par, W, H, g);
   for (int l = 0; l < g.S + 2; l++) g.sigma[l] = c->lvl_sigma[l];
   CK(cudaMemcpyAsync(c->d_geom, &g, sizeof(Geom), cudaMemcpyHostToDevice, ust));

   // output buffers for the whole batch
   if ((size_t)n > c->counts_cap || !c->d_ndet) {
      if (c->d_ndet) cudaFree(c->d_ndet);
      if (c->d_ndesc) cudaFree(c->d_ndesc);
      c->d_ndet = c->d_ndesc = nullptr;
      int rc;
      if ((rc = dmalloc(&c->d_ndet, (size_t)n + 1)) || (rc = dmalloc(&c->d_ndesc, (size_t)n + 1))) return rc;
      c->counts_cap = n;
   }
   const size_t want_keys = std::max<size_t>(1024, (size_t)n * (size_t)(c->max_cand_per_image / 2));
   if (want_keys > c->keys_cap) {
      if (c->d_keys) cudaFree(c->d_keys);
      if (c->d_ell) cudaFree(c->d_ell);
      c->d_keys = nullptr; c->d_ell = nullptr;
      int rc;
      if ((rc = dmalloc(&c->d_keys, want_keys)) || (rc = dmalloc(&c->d_ell, want_keys * 5))) return rc;
      c->keys_cap = want_keys;
   }
   CK(cudaMemsetAsync(c->d_ndet, 0, sizeof(int) * ((size_t)n + 1), ust));
   CK(cudaMemsetAsync(c->d_ndesc, 0, sizeof(int) * ((size_t)n + 1), ust));
   CK(cudaMemsetAsync(c->d_out_base, 0, sizeof(uint32_t) * 2, ust));
   CK(cudaMemsetAsync(c->d_overflow, 0, sizeof(int) * 2, ust));
   CK(cudaEventRecord(c->ev_start, ust));
   memset(c->stage_ms, 0, sizeof(c->stage_ms));
   c->blur_ms = 0.f; c->blur_launches = 0;
   c->n_images = n;
   c->last_chunks = 0;
   c->last_u8 = fmt == IN_U8;
   const int S = g.S;
   const int lanes = c->profiling ? 1 : c->n_lanes;
   for (int l = 0; l < c->n_lanes; l++) {
      c->lane[l].pending_chunk = -1;
      CK(cudaStreamWaitEvent(c->lane[l].stream, c->ev_start, 0));
   }
   cudaEvent_t prev_done = nullptr, prev_front = nullptr;
   // (HESAFF_OVERLAP=0 turns this off) the front end (pyramid .. affine shape) of chunk k+1 may run under the describe stage of chunk k
   // (front ends stay in order among themselves, and so do the describe + compaction stages)
   static const bool overlap_front = !getenv("HESAFF_OVERLAP") || atoi(getenv("HESAFF_OVERLAP")) > 0;
   const bool overlap = overlap_front && !c->profiling;

   for (int start = 0, k = 0; start < n; start += c->chunk, k++) {
      const int cn = std::min(c->chunk, n - start);
      Lane &L = c->lane[k % lanes];
      cudaStream_t st = L.stream;
      // the lane's previous chunk must be finished before its buffers are reused by the host-side staging copy and
      // before its records can be streamed out
      { int rc = flush_lane_output(c, L); if (rc) return rc; }
      c->last_chunks++;
      if (c->profiling) cudaEventRecord(L.ev[0], st);
      NvtxStages nvtx("hesaff:chunk");
      nvtx.next("hesaff:upload+convert");
      // ---- stage 0: upload + gray float image (hesaff.cpp:138-148) ---------------------------------
      const char *src = ptrs ? nullptr : (const char *)images + (size_t)start * img_stride;
      const void *dsrc = src;
      size_t d_row_pitch = row_pitch, d_img_stride = img_stride;
      if (!on_device) {
         d_row_pitch = (size_t)W * esz; d_img_stride = d_row_pitch * H;
         if (ptrs) {
            for (int i = 0; i < cn; i++) {
               const char *pi = (const char *)ptrs[start + i];
               if (row_pitch == d_row_pitch)
                  CK(cudaMemcpyAsync(L.stage_u8 + (size_t)i * d_img_stride, pi, d_img_stride, cudaMemcpyHostToDevice, st));
               else
                  CK(cudaMemcpy2DAsync(L.stage_u8 + (size_t)i * d_img_stride, d_row_pitch, pi, row_pitch, d_row_pitch, H,
                                       cudaMemcpyHostToDevice, st));
            }
         } else if (row_pitch == d_row_pitch && img_stride == d_img_stride)
            CK(cudaMemcpyAsync(L.stage_u8, src, d_img_stride * cn, cudaMemcpyHostToDevice, st));
         else
            for (int i = 0; i < cn; i++)
               CK(cudaMemcpy2DAsync(L.stage_u8 + (size_t)i * d_img_stride, d_row_pitch, src + (size_t)i * img_stride,
                                    row_pitch, d_row_pitch, H, cudaMemcpyHostToDevice, st));
         dsrc = L.stage_u8;
      }
      // Compute of consecutive chunks is serialised (the describe stage already fills the SMs with its own
      // side-by-side kernels; a second lane's kernels only disturb that packing -- measured); what the second lane
      // buys is the upload of chunk k+1 and the download of chunk k-1 running under the compute of chunk k.
      if (overlap) { if (prev_front) CK(cudaStreamWaitEvent(st, prev_front, 0)); }
      else if (prev_done) CK(cudaStreamWaitEvent(st, prev_done, 0));
      float *img_plane = L.arena + g.img_off;
      if (fmt == IN_U8) ha_launch_convert_u8((const uint8_t *)dsrc, d_row_pitch, d_img_stride, L.arena, g, cn, st, c->lc);
      else if (fmt == IN_RGB8) ha_launch_convert_rgb8((const uint8_t *)dsrc, d_row_pitch, d_img_stride, img_plane, g, cn, st, c->lc);
      else ha_launch_convert_f32((const float *)dsrc, d_row_pitch, d_img_stride, img_plane, g, cn, st, c->lc);
      if (c->profiling) cudaEventRecord(L.ev[1], st);
      nvtx.next("hesaff:pyramid");

      // ---- stage 1: pyramid (pyramid.cpp:261-292, 224-259) ------------------------------------------
      size_t bev = 0;
      auto blur_mark = [&]() {   // profiling: an event before and after every blur launch
         if (!c->profiling) return;
         if (bev >= L.blur_ev.size()) { cudaEvent_t e; cudaEventCreate(&e); L.blur_ev.push_back(e); }
         cudaEventRecord(L.blur_ev[bev++], st);
      };
      if (g.nOct > 0) {
         blur_mark();
         float *L00 = L.arena + g.L_off[0][0], *R00 = L.arena + g.R_off[0][0];
         if (c->taps0.n > 0) {
            if (ha_launch_blur(img_plane, L00, R00, nullptr, g.w[0], g.h[0], g.pitch[0], 0, 0, 0, g.arena_stride, c->norm[0],
                               c->taps0, cn, st, c->lc))
               return fail(HESAFF_ERR_INVALID, "unsupported blur size");
         } else {
            Taps id; id.n = 1; id.dk = nullptr; memset(id.k, 0, sizeof(id.k)); id.k[0] = 1.0f;   // firstLevel = image.clone()
            ha_launch_blur(img_plane, L00, R00, nullptr, g.w[0], g.h[0], g.pitch[0], 0, 0, 0, g.arena_stride, c->norm[0], id,
                           cn, st, c->lc);
         }
         blur_mark();
      }
      for (int o = 0; o < g.nOct; o++) {
         if (o > 0)   // response of the decimated first level (cur = hessianResponse(blur, sigma0^2), pyramid.cpp:230)
            ha_launch_hessian(L.arena + g.L_off[o][0], L.arena + g.R_off[o][0], g.w[o], g.h[o], g.pitch[o], g.arena_stride,
                              c->norm[0], cn, st, c->lc);
         for (int i = 1; i < S + 2; i++) {
            const bool seed_next = (i == S) && (o + 1 < g.nOct);   // halfImage(nextBlur) at i == numberOfScales
            float *half = seed_next ? L.arena + g.L_off[o + 1][0] : nullptr;
            blur_mark();
            if (ha_launch_blur(L.arena + g.L_off[o][i - 1], L.arena + g.L_off[o][i], L.arena + g.R_off[o][i], half, g.w[o],
                               g.h[o], g.pitch[o], seed_next ? g.w[o + 1] : 0, seed_next ? g.h[o + 1] : 0,
                               seed_next ? g.pitch[o + 1] : 0, g.arena_stride, c->norm[i], c->taps[i], cn, st, c->lc))
               return fail(HESAFF_ERR_INVALID, "unsupported blur size");
            blur_mark();
         }
      }
      if (c->profiling) cudaEventRecord(L.ev[2], st);
      nvtx.next("hesaff:nms+localize");

      // ---- stage 2: extrema, ordered compaction, localisation, dedup ------------------------------------
      const size_t nwords = g.mask_stride * cn;
      ha_launch_nms(L.arena, g, c->d_geom, L.mask, cn, st, c->lc);
      ha_launch_scan_popc(L.mask, nwords, L.woff, L.scan_tmp, st, c->lc);
      const uint32_t *d_count = L.woff + nwords;   // number of candidates in this chunk
      ha_launch_expand(L.mask, L.woff, c->d_geom, nwords, L.cand, c->cand_cap, c->d_overflow, st, c->lc);
      CK(cudaMemsetAsync(L.map, 0xFF, sizeof(uint32_t) * g.map_stride * cn, st));
      ha_launch_localize(L.arena, c->d_geom, L.cand, d_count, c->cand_cap, L.map, st, c->lc);
      if (c->profiling) cudaEventRecord(L.ev[3], st);
      nvtx.next("hesaff:affine-shape");

      // ---- stage 3: affine shape ---------------------------------------------------------------------
      CK(cudaMemsetAsync(L.counters, 0, sizeof(int) * 7, st));
      CK(cudaMemsetAsync(L.bins.count, 0, sizeof(int) * 6, st));
      ha_launch_affine(L.arena, c->d_geom, c->tables, L.cand, d_count, c->cand_cap, L.map, c->d_ndet + start, L.bins,
                       L.counters, st, c->lc);
      if (c->profiling) cudaEventRecord(L.ev[4], st);
      if (overlap) {
         CK(cudaEventRecord(L.front_done, st));
         prev_front = L.front_done;
         if (prev_done) CK(cudaStreamWaitEvent(st, prev_done, 0));
      }

      nvtx.next("hesaff:patch+sift");
      // ---- stage 4: patch normalisation + SIFT ----------------------------------------------------------
      ha_launch_describe(L.arena, c->d_geom, c->geom, c->tables, L.cand, L.bins, L.counters + 1, L.scratch, c->scratch_per_cta,
                         c->large_ctas, c->maxP, fmt == IN_U8, nullptr, 0, nullptr, st, c->lc, L.aux, L.ev_fork, L.ev_join);
      if (c->profiling) cudaEventRecord(L.ev[5], st);

      nvtx.next("hesaff:compact");
      // ---- stage 5: ordered compaction into Keypoint records -------------------------------------------
      ha_launch_scan_flags(L.cand.flags, HA_F_DESC, d_count, c->cand_cap, L.desc_off, L.scan_tmp, st, c->lc);
      // records of this chunk start at the running total of the previous chunks (device-side base); chunk order is
      // guaranteed by the wait above
      ha_launch_compact(L.cand, d_count, c->cand_cap, L.desc_off, c->d_geom, c->d_keys, c->d_ell, c->d_ndesc + start,
                        c->d_out_base, (uint32_t)std::min<size_t>(c->keys_cap, 0xFFFFFFFFu), c->d_overflow, st, c->lc);
      uint32_t *d_report = nullptr;
      CK(cudaHostGetDevicePointer((void **)&d_report, L.h_total, 0));
      k_add_total<<<1, 1, 0, st>>>(c->d_out_base, L.desc_off + c->cand_cap, d_report);
      c->lc.n++;
      CK(cudaEventRecord(L.done, st));
      prev_done = L.done;
      L.pending_chunk = k;

      if (c->profiling) {
         cudaEventRecord(L.ev[6], st);
         CK(cudaEventSynchronize(L.ev[6]));
         for (int q = 0; q < 6; q++) {
            float ms = 0;
            cudaEventElapsedTime(&ms, L.ev[q], L.ev[q + 1]);
            c->stage_ms[q] += ms;
         }
         for (size_t q = 0; q + 1 < bev; q += 2) {
            float ms = 0;
            cudaEventElapsedTime(&ms, L.blur_ev[q], L.blur_ev[q + 1]);
            c->blur_ms += ms;
            c->blur_launches++;
         }
      }
   }
   // drain: stream out the last chunks, then counts to the host
   for (int l = 0; l < c->n_lanes; l++) { int rc = flush_lane_output(c, c->lane[l]); if (rc) return rc; }
   for (int l = 0; l < c->n_lanes; l++) CK(cudaStreamSynchronize(c->lane[l].stream));
   c->h_ndet.assign(n, 0); c->h_ndesc.assign(n, 0);
   int overflow = 0;
   uint32_t total = 0;
   if (n > 0) {
      CK(cudaMemcpyAsync(c->h_ndet.data(), c->d_ndet, sizeof(int) * n, cudaMemcpyDeviceToHost, ust));
      CK(cudaMemcpyAsync(c->h_ndesc.data(), c->d_ndesc, sizeof(int) * n, cudaMemcpyDeviceToHost, ust));
   }
   CK(cudaMemcpyAsync(&overflow, c->d_overflow, sizeof(int), cudaMemcpyDeviceToHost, ust));
   CK(cudaMemcpyAsync(&total, c->d_out_base, sizeof(uint32_t), cudaMemcpyDeviceToHost, ust));
   CK(cudaStreamSynchronize(ust));
   CK(cudaStreamSynchronize(c->copy_stream));
   CK(cudaGetLastError());
   if (overflow) return fail(HESAFF_ERR_CAPACITY, "candidate pool / keypoint buffer overflow; raise max_candidates_per_image");
   if (total > c->keys_cap) return fail(HESAFF_ERR_CAPACITY, "keypoint output buffer overflow; raise max_candidates_per_image");
   c->total_desc = total;
   c->total_det = 0;
   for (int i = 0; i < n; i++) c->total_det += c->h_ndet[i];
   c->have_result = true;
   c->host_out_filled = c->host_out != nullptr;
   return HESAFF_OK;
}

extern "C" int hesaff_detect_u8(hesaff_ctx *ctx, const uint8_t *images, int n, int width, int height, size_t row_pitch_bytes,
                                size_t image_stride_bytes, int on_device, void *stream)
{
   return detect_impl(ctx, images, IN_U8, n, width, height, row_pitch_bytes, image_stride_bytes, on_device, stream);
}

extern "C" int hesaff_detect_f32(hesaff_ctx *ctx, const float *images, int n, int width, int height, size_t row_pitch_bytes,
                                 size_t image_stride_bytes, int on_device, void *stream)
{
   return detect_impl(ctx, images, IN_F32, n, width, height, row_pitch_bytes, image_stride_bytes, on_device, stream);
}

extern "C" int hesaff_detect_rgb8(hesaff_ctx *ctx, const uint8_t *images, int n, int width, int height, size_t row_pitch_bytes,
                                  size_t image_stride_bytes, int on_device, void *stream)
{
   return detect_impl(ctx, images, IN_RGB8, n, width, height, row_pitch_bytes, image_stride_bytes, on_device, stream);
}

// ---- binary PNM files: header on the host, pixels on the device (the step before the path, hesaff.cpp:137-148) ---------
// Parses the header of a binary PNM ("P5" gray / "P6" colour, maxval 255; '#' comments allowed) held in memory.
extern "C" int hesaff_pnm_info(const void *file, size_t bytes, int *width, int *height, int *channels, size_t *data_offset)
{
   const unsigned char *p = (const unsigned char *)file;
   if (!p || bytes < 7 || p[0] != 'P' || (p[1] != '5' && p[1] != '6')) return fail(HESAFF_ERR_INVALID, "not a binary PNM (P5/P6) file");
   size_t pos = 2;
   long vals[3] = {0, 0, 0};
   for (int got = 0; got < 3;) {
      if (pos >= bytes) return fail(HESAFF_ERR_INVALID, "truncated PNM header");
      const unsigned char ch = p[pos];
      if (ch == '#') { while (pos < bytes && p[pos] != '\n') pos++; continue; }
      if (ch == ' ' || ch == '\t' || ch == '\n' || ch == '\r' || ch == '\v' || ch == '\f') { pos++; continue; }
      if (ch < '0' || ch > '9') return fail(HESAFF_ERR_INVALID, "bad PNM header");
      long v = 0;
      while (pos < bytes && p[pos] >= '0' && p[pos] <= '9') { v = v * 10 + (p[pos] - '0'); pos++; if (v > (1l << 30)) return fail(HESAFF_ERR_INVALID, "bad PNM header"); }
      vals[got++] = v;
   }
   pos++;   // the single whitespace byte after maxval
   const int ch = p[1] == '5' ? 1 : 3;
   if (vals[0] <= 0 || vals[1] <= 0 || vals[2] != 255) return fail(HESAFF_ERR_INVALID, "PNM: only maxval 255 is supported");
   if (pos > bytes || bytes - pos < (size_t)vals[0] * (size_t)vals[1] * (size_t)ch) return fail(HESAFF_ERR_INVALID, "truncated PNM pixel data");
   if (width) *width = (int)vals[0];
   if (height) *height = (int)vals[1];
   if (channels) *channels = ch;
   if (data_offset) *data_offset = pos;
   return HESAFF_OK;
}

extern "C" int hesaff_detect_pnm(hesaff_ctx *ctx, const void *const *files, const size_t *file_bytes, int n, void *stream)
{
   if (!ctx || !files || !file_bytes || n < 1) return fail(HESAFF_ERR_INVALID, "bad arguments");
   std::vector<const void *> pix((size_t)n);
   int W = 0, H = 0, CH = 0;
   for (int i = 0; i < n; i++) {
      int w, h, ch;
      size_t off;
      const int rc = hesaff_pnm_info(files[i], file_bytes[i], &w, &h, &ch, &off);
      if (rc) return rc;
      if (i == 0) { W = w; H = h; CH = ch; }
      else if (w != W || h != H || ch != CH) return fail(HESAFF_ERR_INVALID, "the files of one call must have the same size and type");
      pix[i] = (const char *)files[i] + off;
   }
   // the pixel payload is handed to the copy engine as it lies in the file; P6 is converted to gray on the GPU
   return detect_impl(ctx, nullptr, CH == 1 ? IN_U8 : IN_RGB8, n, W, H, (size_t)W * CH, (size_t)W * CH * H, 0, stream, pix.data());
}

// ---- consumer of the records: descriptor matching (match.cu) ----------------------------------------------------------
extern "C" int hesaff_match_descriptors(int device, const hesaff_keypoint *d_query, size_t n_query, const hesaff_keypoint *d_db,
                                        size_t n_db, int32_t *d_best_index, uint32_t *d_best_dist2, uint32_t *d_second_dist2,
                                        void *stream)
{
   if ((n_query && !d_query) || (n_db && !d_db) || !d_best_index || !d_best_dist2) return fail(HESAFF_ERR_INVALID, "NULL argument");
   if (n_query > 0xffffffffull || n_db > 0x7fffffffull) return fail(HESAFF_ERR_INVALID, "too many records");
   int ndev = 0;
   if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(HESAFF_ERR_CUDA, "no CUDA device: this library has no CPU fallback");
   if (device < 0 || device >= ndev) return fail(HESAFF_ERR_INVALID, "bad device index");
   CK(cudaSetDevice(device));
   ha_launch_match(d_query, (uint32_t)n_query, d_db, (uint32_t)n_db, d_best_index, d_best_dist2, d_second_dist2, (cudaStream_t)stream);
   CK(cudaGetLastError());
   CK(cudaStreamSynchronize((cudaStream_t)stream));
   return HESAFF_OK;
}

// ---- results -------------------------------------------------------------------------------------------
#define NEED_RESULT(c)                                                                           \
   do {                                                                                          \
      if (!(c)) return fail(HESAFF_ERR_INVALID, "ctx is NULL");                                  \
      if (!(c)->have_result) return fail(HESAFF_ERR_STATE, "no successful hesaff_detect_* yet"); \
      CK(cudaSetDevice((c)->device));                                                            \
   } while (0)

extern "C" int hesaff_result_counts(hesaff_ctx *c, int *n_detected, int *n_described)
{
   NEED_RESULT(c);
   if (n_detected) memcpy(n_detected, c->h_ndet.data(), sizeof(int) * c->n_images);
   if (n_described) memcpy(n_described, c->h_ndesc.data(), sizeof(int) * c->n_images);
   return HESAFF_OK;
}

extern "C" int64_t hesaff_result_total(hesaff_ctx *c)
{
   if (!c || !c->have_result) return fail(HESAFF_ERR_STATE, "no result");
   return c->total_desc;
}

extern "C" int hesaff_result_keypoints(hesaff_ctx *c, hesaff_keypoint *out, size_t capacity)
{
   NEED_RESULT(c);
   if ((size_t)c->total_desc > capacity) return fail(HESAFF_ERR_CAPACITY, "output capacity too small");
   if (out == c->host_out && c->host_out_filled) return HESAFF_OK;   // already streamed during the detect call
   if (c->total_desc) CK(cudaMemcpy(out, c->d_keys, sizeof(hesaff_keypoint) * c->total_desc, cudaMemcpyDeviceToHost));
   return HESAFF_OK;
}

extern "C" int hesaff_result_keypoints_device(hesaff_ctx *c, const hesaff_keypoint **out)
{
   NEED_RESULT(c);
   *out = c->d_keys;
   return HESAFF_OK;
}

extern "C" int hesaff_result_ellipses(hesaff_ctx *c, float *out, size_t capacity)
{
   NEED_RESULT(c);
   if ((size_t)c->total_desc > capacity) return fail(HESAFF_ERR_CAPACITY, "output capacity too small");
   if (c->total_desc) CK(cudaMemcpy(out, c->d_ell, sizeof(float) * 5 * c->total_desc, cudaMemcpyDeviceToHost));
   return HESAFF_OK;
}

extern "C" int hesaff_result_detections(hesaff_ctx *c, hesaff_detection *out, size_t capacity, int64_t *n_total)
{
   NEED_RESULT(c);
   if (c->last_chunks != 1) return fail(HESAFF_ERR_STATE, "per-detection records are kept for single-chunk calls only");
   if (n_total) *n_total = c->total_det;
   if (!out) return HESAFF_OK;
   if ((size_t)c->total_det > capacity) return fail(HESAFF_ERR_CAPACITY, "output capacity too small");
   if ((size_t)c->total_det > c->dets_cap) {
      if (c->d_dets) cudaFree(c->d_dets);
      c->d_dets = nullptr;
      int rc;
      if ((rc = dmalloc(&c->d_dets, (size_t)c->total_det + 16))) return rc;
      c->dets_cap = c->total_det + 16;
   }
   Lane &L = c->lane[0];
   const size_t nwords = c->geom.mask_stride * c->n_images;
   const uint32_t *d_count = L.woff + nwords;
   ha_launch_scan_flags(L.cand.flags, HA_F_DET, d_count, c->cand_cap, L.det_off, L.scan_tmp, L.stream, c->lc);
   ha_launch_export_detections(L.cand, d_count, c->cand_cap, L.det_off, c->d_geom, c->d_dets, L.stream, c->lc);
   CK(cudaStreamSynchronize(L.stream));
   if (c->total_det) CK(cudaMemcpy(out, c->d_dets, sizeof(hesaff_detection) * c->total_det, cudaMemcpyDeviceToHost));
   return HESAFF_OK;
}

// ---- stage access -------------------------------------------------------------------------------------------
extern "C" int hesaff_debug_geometry(hesaff_ctx *c, int *n_octaves, int *n_levels)
{
   NEED_RESULT(c);
   if (n_octaves) *n_octaves = c->geom.nOct;
   if (n_levels) *n_levels = c->geom.S + 2;
   return HESAFF_OK;
}

extern "C" int hesaff_debug_octave_size(hesaff_ctx *c, int octave, int *width, int *height)
{
   NEED_RESULT(c);
   if (octave < 0 || octave >= c->geom.nOct) return fail(HESAFF_ERR_INVALID, "bad octave");
   if (width) *width = c->geom.w[octave];
   if (height) *height = c->geom.h[octave];
   return HESAFF_OK;
}

extern "C" int hesaff_debug_plane(hesaff_ctx *c, int image, int octave, int level, int kind, float *out)
{
   NEED_RESULT(c);
   if (c->last_chunks != 1) return fail(HESAFF_ERR_STATE, "planes are kept for single-chunk calls only");
   const Geom &g = c->geom;
   if (image < 0 || image >= c->n_images || octave < 0 || octave >= g.nOct || level < 0 || level >= g.S + 2)
      return fail(HESAFF_ERR_INVALID, "bad plane index");
   const float *p = c->lane[0].arena + (size_t)image * g.arena_stride + (kind ? g.R_off[octave][level] : g.L_off[octave][level]);
   CK(cudaMemcpy2D(out, sizeof(float) * g.w[octave], p, sizeof(float) * g.pitch[octave], sizeof(float) * g.w[octave],
                   g.h[octave], cudaMemcpyDeviceToHost));
   return HESAFF_OK;
}

extern "C" int hesaff_debug_patches(hesaff_ctx *c, int normalized, float *out, size_t capacity_patches)
{
   NEED_RESULT(c);
   if (c->last_chunks != 1) return fail(HESAFF_ERR_STATE, "patches can be recomputed for single-chunk calls only");
   if ((size_t)c->total_desc > capacity_patches) return fail(HESAFF_ERR_CAPACITY, "output capacity too small");
   if (!c->total_desc) return HESAFF_OK;
   float *d = nullptr;
   int rc;
   if ((rc = dmalloc(&d, (size_t)c->total_desc * HA_PATCH_PX))) return rc;
   Lane &L = c->lane[0];
   CK(cudaMemsetAsync(L.counters + 1, 0, sizeof(int) * 6, L.stream));
   ha_launch_describe(L.arena, c->d_geom, c->geom, c->tables, L.cand, L.bins, L.counters + 1, L.scratch, c->scratch_per_cta,
                      c->large_ctas, c->maxP, c->last_u8, d, normalized, L.desc_off, L.stream, c->lc);
   CK(cudaStreamSynchronize(L.stream));
   cudaError_t e = cudaMemcpy(out, d, sizeof(float) * HA_PATCH_PX * c->total_desc, cudaMemcpyDeviceToHost);
   cudaFree(d);
   CK(e);
   return HESAFF_OK;
}

extern "C" int64_t hesaff_launch_count(hesaff_ctx *c, int reset)
{
   if (!c) return 0;
   const int64_t v = c->lc.n;
   if (reset) c->lc.n = 0;
   return v;
}

extern "C" int hesaff_set_profiling(hesaff_ctx *c, int enable)
{
   if (!c) return fail(HESAFF_ERR_INVALID, "ctx is NULL");
   c->profiling = enable != 0;
   return HESAFF_OK;
}

extern "C" int hesaff_stage_times_ms(hesaff_ctx *c, float *out6)
{
   if (!c || !out6) return fail(HESAFF_ERR_INVALID, "NULL argument");
   memcpy(out6, c->stage_ms, sizeof(c->stage_ms));
   return HESAFF_OK;
}

extern "C" int hesaff_blur_time_ms(hesaff_ctx *c, float *total_ms, int *launches)
{
   if (!c) return fail(HESAFF_ERR_INVALID, "NULL argument");
   if (total_ms) *total_ms = c->blur_ms;
   if (launches) *launches = c->blur_launches;
   return HESAFF_OK;
}

// exportKeypoints, hesaff.cpp:107-130 (text formatting on the host; ostream defaults = 6 significant digits)
static void format_sift_lines_host(std::ostream &out, const hesaff_keypoint *kps, size_t n, float desc_factor)
{
   for (size_t i = 0; i < n; i++) {
      const hesaff_keypoint &k = kps[i];
      const double sc = (double)(desc_factor * k.s);
      const double a = k.a11, b = k.a12, c = k.a21, d = k.a22;
      const double p = a * a + b * b, q = a * c + b * d, r = c * c + d * d;
      const double det = p * r - q * q, isc2 = 1.0 / (sc * sc);
      out << k.x << " " << k.y << " " << (float)(r / det * isc2) << " " << (float)(-q / det * isc2) << " "
          << (float)(p / det * isc2);
      for (size_t j = 0; j < 128; j++) out << " " << int(k.desc[j]);
      out << "\n";
   }
}

extern "C" int hesaff_write_sift_file(const char *path, const hesaff_keypoint *kps, size_t n, float desc_factor)
{
   if (!path || (!kps && n)) return fail(HESAFF_ERR_INVALID, "NULL argument");
   std::ofstream out(path);
   if (!out) return fail(HESAFF_ERR_INVALID, std::string("cannot open ") + path);
   out << 128 << "\n" << n << "\n";
   format_sift_lines_host(out, kps, n, desc_factor);
   out.flush();
   if (!out) return fail(HESAFF_ERR_INVALID, std::string("write failed: ") + path);
   return (int)n;
}

// ---- GPU-side text export -------------------------------------------------------------------
// first record and count of image `image` in the image-major output
static void image_range(const hesaff_ctx *c, int image, size_t &first, size_t &n)
{
   first = 0;
   for (int i = 0; i < image; i++) first += (size_t)c->h_ndesc[i];
   n = (size_t)c->h_ndesc[image];
}

extern "C" int hesaff_result_sift_text(hesaff_ctx *c, int image, char *out, size_t capacity, size_t *nbytes)
{
   NEED_RESULT(c);
   if (image < 0 || image >= c->n_images) return fail(HESAFF_ERR_INVALID, "bad image index");
   size_t first, n;
   image_range(c, image, first, n);
   char header[64];
   const int hl = snprintf(header, sizeof(header), "128\n%zu\n", n);
   cudaStream_t st = c->stream;
   int rc;
   if (n + 1 > c->tlen_cap) {
      if (c->d_tlen) cudaFree(c->d_tlen);
      if (c->d_toff) cudaFree(c->d_toff);
      c->d_tlen = c->d_toff = nullptr;
      c->tlen_cap = 0;
      if ((rc = dmalloc(&c->d_tlen, n + 1)) || (rc = dmalloc(&c->d_toff, n + 2))) return rc;
      c->tlen_cap = n + 1;
   }
   if (!c->d_bad && (rc = dmalloc(&c->d_bad, 1))) return rc;
   uint32_t body = 0;
   int bad = 0;
   if (n) {
      if (ha_scan_tmp_elems(n) > c->scan_tmp_elems) return fail(HESAFF_ERR_CAPACITY, "scan scratch too small for this image");
      CK(cudaMemsetAsync(c->d_bad, 0, sizeof(int), st));
      ha_launch_sift_text(false, c->d_keys + first, c->d_ell + first * 5, (uint32_t)n, c->d_tlen, nullptr, nullptr, c->d_bad, st, c->lc);
      ha_launch_scan_u32(c->d_tlen, n, c->d_toff, c->lane[0].scan_tmp, st, c->lc);
      CK(cudaMemcpyAsync(&body, c->d_toff + n, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
      CK(cudaMemcpyAsync(&bad, c->d_bad, sizeof(int), cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
   }
   if (bad) {
      // a value outside the device formatter's range (inf / nan / >= 2^63): format this image on the host
      std::vector<hesaff_keypoint> k(n);
      CK(cudaMemcpy(k.data(), c->d_keys + first, sizeof(hesaff_keypoint) * n, cudaMemcpyDeviceToHost));
      std::ostringstream os;
      format_sift_lines_host(os, k.data(), n, c->par.desc_factor);
      const std::string s = os.str();
      if (nbytes) *nbytes = (size_t)hl + s.size();
      if (!out) return HESAFF_OK;
      if ((size_t)hl + s.size() > capacity) return fail(HESAFF_ERR_CAPACITY, "text buffer too small");
      memcpy(out, header, hl);
      memcpy(out + hl, s.data(), s.size());
      return HESAFF_OK;
   }
   if (nbytes) *nbytes = (size_t)hl + body;
   if (!out) return HESAFF_OK;
   if ((size_t)hl + body > capacity) return fail(HESAFF_ERR_CAPACITY, "text buffer too small");
   memcpy(out, header, hl);
   if (n) {
      if ((size_t)body > c->text_cap) {
         if (c->d_text) cudaFree(c->d_text);
         c->d_text = nullptr;
         c->text_cap = 0;
         if ((rc = dmalloc(&c->d_text, (size_t)body + 64))) return rc;
         c->text_cap = (size_t)body + 64;
      }
      ha_launch_sift_text(true, c->d_keys + first, c->d_ell + first * 5, (uint32_t)n, nullptr, c->d_toff, c->d_text, c->d_bad, st, c->lc);
      CK(cudaMemcpyAsync(out + hl, c->d_text, body, cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
   }
   return HESAFF_OK;
}

extern "C" int hesaff_export_sift_file(hesaff_ctx *c, int image, const char *path)
{
   NEED_RESULT(c);
   if (!path) return fail(HESAFF_ERR_INVALID, "NULL path");
   size_t nb = 0;
   int rc = hesaff_result_sift_text(c, image, nullptr, 0, &nb);
   if (rc) return rc;
   std::vector<char> buf(nb);
   if ((rc = hesaff_result_sift_text(c, image, buf.data(), buf.size(), &nb))) return rc;
   FILE *f = fopen(path, "wb");
   if (!f) return fail(HESAFF_ERR_INVALID, std::string("cannot open ") + path);
   const size_t w = fwrite(buf.data(), 1, nb, f);
   if (fclose(f) != 0 || w != nb) return fail(HESAFF_ERR_INVALID, std::string("write failed: ") + path);
   return c->h_ndesc[image];
}

// Binary sidecar: "HESAFFB1", u32 record size (164), u32 reserved, u64 count, then the Keypoint records.
extern "C" int hesaff_write_keypoints_binary(const char *path, const hesaff_keypoint *kps, size_t n)
{
   if (!path || (!kps && n)) return fail(HESAFF_ERR_INVALID, "NULL argument");
   FILE *f = fopen(path, "wb");
   if (!f) return fail(HESAFF_ERR_INVALID, std::string("cannot open ") + path);
   const uint32_t hdr[2] = {(uint32_t)sizeof(hesaff_keypoint), 0u};
   const uint64_t cnt = n;
   bool ok = fwrite("HESAFFB1", 1, 8, f) == 8 && fwrite(hdr, sizeof(hdr), 1, f) == 1 && fwrite(&cnt, sizeof(cnt), 1, f) == 1;
   ok = ok && (n == 0 || fwrite(kps, sizeof(hesaff_keypoint), n, f) == n);
   if (fclose(f) != 0 || !ok) return fail(HESAFF_ERR_INVALID, std::string("write failed: ") + path);
   return HESAFF_OK;
}

extern "C" int hesaff_read_keypoints_binary(const char *path, hesaff_keypoint *out, size_t capacity, size_t *n)
{
   if (!path || !n) return fail(HESAFF_ERR_INVALID, "NULL argument");
   FILE *f = fopen(path, "rb");
   if (!f) return fail(HESAFF_ERR_INVALID, std::string("cannot open ") + path);
   char magic[8];
   uint32_t hdr[2];
   uint64_t cnt = 0;
   bool ok = fread(magic, 1, 8, f) == 8 && !memcmp(magic, "HESAFFB1", 8) && fread(hdr, sizeof(hdr), 1, f) == 1 &&
             hdr[0] == sizeof(hesaff_keypoint) && fread(&cnt, sizeof(cnt), 1, f) == 1;
   if (!ok) { fclose(f); return fail(HESAFF_ERR_INVALID, std::string("not a hesaff binary keypoint file: ") + path); }
   *n = (size_t)cnt;
   if (!out) { fclose(f); return HESAFF_OK; }
   if (cnt > capacity) { fclose(f); return fail(HESAFF_ERR_CAPACITY, "output capacity too small"); }
   ok = cnt == 0 || fread(out, sizeof(hesaff_keypoint), cnt, f) == cnt;
   fclose(f);
   if (!ok) return fail(HESAFF_ERR_INVALID, std::string("truncated file: ") + path);
   return HESAFF_OK;
}

// Diagnostic: the device float formatter on its own; out receives n slots of 16 bytes, NUL padded ("?" = not covered).
extern "C" int hesaff_debug_format_floats(hesaff_ctx *c, const float *in, size_t n, char *out)
{
   if (!c || !in || !out) return fail(HESAFF_ERR_INVALID, "NULL argument");
   CK(cudaSetDevice(c->device));
   if (!n) return HESAFF_OK;
   float *d_in = nullptr;
   char *d_out = nullptr;
   int rc;
   if ((rc = dmalloc(&d_in, n))) return rc;
   if ((rc = dmalloc(&d_out, n * 16))) { cudaFree(d_in); return rc; }
   cudaError_t e = cudaMemcpy(d_in, in, sizeof(float) * n, cudaMemcpyHostToDevice);
   if (e == cudaSuccess) {
      ha_launch_format_floats(d_in, n, d_out, c->stream);
      e = cudaStreamSynchronize(c->stream);
   }
   if (e == cudaSuccess) e = cudaMemcpy(out, d_out, n * 16, cudaMemcpyDeviceToHost);
   cudaFree(d_in);
   cudaFree(d_out);
   CK(e);
   return HESAFF_OK;
}
