// tools/blur_bench.cu -- standalone micro-benchmark + bit-exactness check of the pyramid blur kernels.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false -prec-div=true -prec-sqrt=true -ftz=false \
//        -lineinfo -DHA_BLUR_VARIANTS=1 -o tools/blur_bench tools/blur_bench.cu -lcuda
// Reference = k_blur<N> of pyramid.cu (no TMA, scalar math; pinned bit-exact against the oracle by the GPU parity tests).
// Every variant must reproduce its L, R and decimated planes bit for bit.  Timing: CUDA events, inputs/outputs of one
// launch (32 x 1080p: 265 MB read, 531 MB written) exceed the 126 MB L2.
#define HA_BLUR_VARIANTS 1
#include <stdio.h>
#include <string.h>
#include <math.h>
#include <vector>
#include <string>
#include "../hesaff_b200/csrc/pyramid.cu"
#include "../hesaff_b200/csrc/blur_tma.cu"

#define CKB(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

static void gauss(int n, double sigma, Taps &t)
{
   const int R = (n - 1) / 2;
   std::vector<double> v(R + 1);
   double sum = 0;
   for (int i = 0; i < R; i++) { double x = i - R; v[i] = exp(-0.5 / (sigma * sigma) * x * x); sum += v[i]; }
   v[R] = 1; sum = 2 * sum + 1;
   t.n = n; t.dk = nullptr; memset(t.k, 0, sizeof(t.k));
   for (int i = 0; i <= R; i++) t.k[i] = t.k[n - 1 - i] = (float)(v[i] / sum);
}

struct Planes { float *src, *L, *R, *half; };

static size_t cmp(const float *a, const float *b, size_t n, const char *what, int W, int pitch, int H, size_t istride, int nimg)
{
   // compare only the in-image columns (the padding up to the pitch is never read by anybody)
   size_t bad = 0;
   for (int im = 0; im < nimg; im++)
      for (int y = 0; y < H; y++)
         for (int x = 0; x < W; x++) {
            const size_t i = (size_t)im * istride + (size_t)y * pitch + x;
            if (memcmp(a + i, b + i, 4)) { if (bad < 3) printf("   %s mismatch img %d (%d,%d): %.9g vs %.9g\n", what, im, x, y, a[i], b[i]); bad++; }
         }
   return bad;
}

int main(int argc, char **argv)
{
   int W = 1920, H = 1080, n = 32, iters = 20;
   if (argc > 2) { W = atoi(argv[1]); H = atoi(argv[2]); }
   if (argc > 3) n = atoi(argv[3]);
   const char *filter = argc > 4 ? argv[4] : nullptr;   // only variants whose name contains this
   const int only_n = argc > 5 ? atoi(argv[5]) : 0;
   const int pitch = (W + 3) & ~3, hW = W / 2, hH = H / 2, hpitch = (hW + 3) & ~3;
   const size_t istride = ((size_t)H * pitch + 63) & ~(size_t)63, hstride = istride;   // half plane lives at the same image stride
   const size_t total = istride * n;
   printf("# blur_bench %dx%d x %d images, pitch %d\n", W, H, n, pitch);
   std::vector<float> h(total);
   unsigned s = 12345;
   for (size_t i = 0; i < total; i++) { s = s * 1664525u + 1013904223u; h[i] = (float)((s >> 13) % 256) + ((s >> 9) & 15) * 0.0625f; }
   float *src, *Lr, *Rr, *Hr, *L, *Rp, *Hp;
   CKB(cudaMalloc(&src, total * 4)); CKB(cudaMalloc(&Lr, total * 4)); CKB(cudaMalloc(&Rr, total * 4)); CKB(cudaMalloc(&Hr, total * 4));
   CKB(cudaMalloc(&L, total * 4)); CKB(cudaMalloc(&Rp, total * 4)); CKB(cudaMalloc(&Hp, total * 4));
   CKB(cudaMemcpy(src, h.data(), total * 4, cudaMemcpyHostToDevice));
   std::vector<float> hLr(total), hRr(total), hHr(total), hL(total), hR(total), hH2(total);
   cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
   LaunchCounter lc{0};
   const int taps_list[] = {5, 7, 9, 11, 13, 15};
   const double sig_list[] = {0.75, 1.0, 1.226274, 1.5198685, 1.946588, 2.452547};
   const double alg_bytes = 12.0 * W * H * n + 4.0 * hW * hH * n;
   for (int ti = 0; ti < 6; ti++) {
      const int N = taps_list[ti];
      if (only_n && N != only_n) continue;
      Taps t; gauss(N, sig_list[ti], t);
      const float norm = 2.56f;
      // reference: non-TMA k_blur (HESAFF_NO_TMA path), called directly
      {
         BlurArgs a; a.src = src; a.dstL = Lr; a.dstR = Rr; a.half = Hr; a.img_stride = istride; a.W = W; a.H = H; a.pitch = pitch;
         a.hW = hW; a.hH = hH; a.hpitch = hpitch; a.norm2 = norm * norm;
         CKB(cudaMemset(Lr, 0, total * 4)); CKB(cudaMemset(Rr, 0, total * 4)); CKB(cudaMemset(Hr, 0, total * 4));
         switch (N) {
            case 5: launch_blur_n<5>(a, t, n, 0); break; case 7: launch_blur_n<7>(a, t, n, 0); break;
            case 9: launch_blur_n<9>(a, t, n, 0); break; case 11: launch_blur_n<11>(a, t, n, 0); break;
            case 13: launch_blur_n<13>(a, t, n, 0); break; case 15: launch_blur_n<15>(a, t, n, 0); break;
         }
         CKB(cudaDeviceSynchronize());
         CKB(cudaMemcpy(hLr.data(), Lr, total * 4, cudaMemcpyDeviceToHost)); CKB(cudaMemcpy(hRr.data(), Rr, total * 4, cudaMemcpyDeviceToHost));
         CKB(cudaMemcpy(hHr.data(), Hr, total * 4, cudaMemcpyDeviceToHost));
      }
      struct V { const char *name; int kind; int variant; };
#define VV(oh, nw, minb, sh, se) (oh | (nw << 8) | (minb << 16) | (sh << 25) | (se << 26))
      std::vector<V> vs = {{"v1 k_blur (no TMA)", 1, 0}, {"v3 default", 3, 0},
                           {"v3 oh40 x4 noshfl", 3, VV(40, 8, 4, 1, 0)}, {"v3 oh40 x4", 3, VV(40, 8, 4, 1, 1)},
                           {"v3 oh48 x3", 3, VV(48, 8, 3, 1, 1)}, {"v3 oh56 x3", 3, VV(56, 8, 3, 1, 1)}, {"v3 oh56 x2", 3, VV(56, 8, 2, 1, 1)},
                           {"v3 oh56 x2 noshfl", 3, VV(56, 8, 2, 1, 0)}};
      for (const V &v : vs) {
         if (filter && !strstr(v.name, filter)) continue;
         auto run = [&]() -> int {
            if (v.kind == 1) {
               BlurArgs a; a.src = src; a.dstL = L; a.dstR = Rp; a.half = Hp; a.img_stride = istride; a.W = W; a.H = H; a.pitch = pitch;
               a.hW = hW; a.hH = hH; a.hpitch = hpitch; a.norm2 = norm * norm;
               switch (N) {
                  case 5: return launch_blur_n<5>(a, t, n, 0); case 7: return launch_blur_n<7>(a, t, n, 0);
                  case 9: return launch_blur_n<9>(a, t, n, 0); case 11: return launch_blur_n<11>(a, t, n, 0);
                  case 13: return launch_blur_n<13>(a, t, n, 0); case 15: return launch_blur_n<15>(a, t, n, 0);
               }
               return -1;
            }
            return ha_launch_blur_tma(src, L, Rp, Hp, W, H, pitch, hW, hH, hpitch, istride, norm, t, n, 0, v.variant);
         };
         CKB(cudaMemset(L, 0xFF, total * 4)); CKB(cudaMemset(Rp, 0xFF, total * 4)); CKB(cudaMemset(Hp, 0, total * 4));
         if (run() != 0) { continue; }
         cudaError_t e = cudaDeviceSynchronize();
         if (e != cudaSuccess) { printf("N=%2d %-26s FAILED: %s\n", N, v.name, cudaGetErrorString(e)); cudaGetLastError(); if (e == cudaErrorIllegalAddress || e == cudaErrorLaunchFailure) return 1; continue; }
         CKB(cudaMemcpy(hL.data(), L, total * 4, cudaMemcpyDeviceToHost)); CKB(cudaMemcpy(hR.data(), Rp, total * 4, cudaMemcpyDeviceToHost));
         CKB(cudaMemcpy(hH2.data(), Hp, total * 4, cudaMemcpyDeviceToHost));
         size_t bad = cmp(hL.data(), hLr.data(), total, "L", W, pitch, H, istride, n) + cmp(hR.data(), hRr.data(), total, "R", W, pitch, H, istride, n) +
                      cmp(hH2.data(), hHr.data(), total, "half", hW, hpitch, hH, hstride, n);
         for (int i = 0; i < 3; i++) run();
         CKB(cudaDeviceSynchronize());
         cudaEventRecord(e0);
         for (int i = 0; i < iters; i++) run();
         cudaEventRecord(e1);
         CKB(cudaDeviceSynchronize());
         float ms; cudaEventElapsedTime(&ms, e0, e1);
         const double us = ms * 1e3 / iters;
         printf("N=%2d %-26s %8.1f us  %7.1f GB/s algorithmic  mismatches %zu %s\n", N, v.name, us, alg_bytes / us / 1e3, bad, bad ? "  <<<<< NOT BIT-EXACT" : "");
         fflush(stdout);
      }
   }
   return 0;
}
