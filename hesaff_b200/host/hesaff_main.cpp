// hesaff_b200/host/hesaff_main.cpp -- C++ host, drop-in for the reference CLI `hesaff <image>`
// (hesaff.cpp:133-180): same HessianAffineParams, same stdout line, same <image>.hesaff.sift output format
// (README:27-44), calling the sm_100a CUDA path through the C-ABI of include/hesaff_b200.h.
//
//   hesaff image.pgm|image.ppm [--threshold T] [--scales S] [--max-octaves K] [--device D] [image2 ...]
//
// No OpenCV: PNM (P5/P6, maxval 255) is read here; P6 is converted with the reference's expression
// gray = (float(B) + G + R) / 3.0f (hesaff.cpp:144).  There is no CPU fallback: without a usable B200 the
// program reports the library's error and exits non-zero.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>
#include "../../include/hesaff_b200.h"

using namespace std;

// hesaff.cpp:21-36
struct HessianAffineParams
{
   float threshold;
   int   max_iter;
   float desc_factor;
   int   patch_size;
   bool  verbose;
   HessianAffineParams()
      {
         threshold = 16.0f/3.0f;
         max_iter = 16;
         desc_factor = 3.0f*sqrt(3.0f);
         patch_size = 41;
         verbose = false;
      }
};

static double wallTime()
{
   struct timespec ts;
   clock_gettime(CLOCK_MONOTONIC, &ts);
   return (double)ts.tv_sec + (double)ts.tv_nsec / 1.0e9;
}

// The whole file as it is on disk; the library parses the PNM header and the pixels go to the GPU untouched
static bool readFile(const char *path, vector<char> &bytes)
{
   ifstream f(path, ios::binary | ios::ate);
   if (!f) return false;
   const streamsize n = f.tellg();
   if (n <= 0) return false;
   bytes.resize((size_t)n);
   f.seekg(0);
   f.read(bytes.data(), n);
   return (streamsize)f.gcount() == n;
}

int main(int argc, char **argv)
{
   HessianAffineParams par;
   hesaff_params p;
   hesaff_params_default(&p);
   int device = 0;
   vector<const char *> files;
   for (int i = 1; i < argc; i++) {
      if (!strcmp(argv[i], "--threshold") && i + 1 < argc) par.threshold = (float)atof(argv[++i]);
      else if (!strcmp(argv[i], "--scales") && i + 1 < argc) p.number_of_scales = atoi(argv[++i]);
      else if (!strcmp(argv[i], "--max-octaves") && i + 1 < argc) p.max_octaves = atoi(argv[++i]);
      else if (!strcmp(argv[i], "--device") && i + 1 < argc) device = atoi(argv[++i]);
      else files.push_back(argv[i]);
   }
   if (files.empty()) {
      printf("\nUsage: hesaff image_name.ppm\nDetects Hessian Affine points and describes them using SIFT descriptor.\nThe detector assumes that the vertical orientation is preserved.\n\n");
      return 0;
   }
   // copy params (hesaff.cpp:150-163)
   p.threshold = par.threshold;
   p.max_iter = par.max_iter;
   p.patch_size = par.patch_size;
   p.desc_factor = par.desc_factor;
   p.verbose = par.verbose;

   int status = 0;
   for (size_t fi = 0; fi < files.size(); fi++) {
      int w = 0, h = 0;
      vector<char> file;
      vector<hesaff_keypoint> keys;
      int nDetected = 0, nAffine = 0;
      double t1 = 0, elapsed = 0;
      bool written = false;
      if (readFile(files[fi], file) && hesaff_pnm_info(file.data(), file.size(), &w, &h, 0, 0) == HESAFF_OK) {
         hesaff_ctx *ctx = 0;
         int rc = hesaff_create(&ctx, &p, device, w, h, 1, 0);
         if (rc == HESAFF_OK) {
            // imread + the gray conversion (float(c0)+c1+c2)/3.0f of hesaff.cpp:137-148: header on the host, pixels on the GPU
            const void *fp = file.data();
            const size_t fb = file.size();
            t1 = wallTime();
            rc = hesaff_detect_pnm(ctx, &fp, &fb, 1, 0);
            if (rc == HESAFF_OK) {
               rc = hesaff_result_counts(ctx, &nDetected, &nAffine);
               keys.resize((size_t)hesaff_result_total(ctx));
               if (rc == HESAFF_OK && !keys.empty()) rc = hesaff_result_keypoints(ctx, keys.data(), keys.size());
               elapsed = wallTime() - t1;
            }
            // exportKeypoints (hesaff.cpp:169-173), with the text formatted on the GPU; the reference times only the
            // detector (hesaff.cpp:166-168), and so does the line below
            if (rc == HESAFF_OK) {
               const string out = string(files[fi]) + ".hesaff.sift";
               if (hesaff_export_sift_file(ctx, 0, out.c_str()) >= 0) written = true;
               else rc = HESAFF_ERR_INVALID;
               if (rc == HESAFF_OK && getenv("HESAFF_BINARY_SIDECAR"))
                  rc = hesaff_write_keypoints_binary((string(files[fi]) + ".hesaff.bin").c_str(), keys.data(), keys.size());
            }
         }
         if (rc != HESAFF_OK) {
            fprintf(stderr, "hesaff_b200: %s\n", hesaff_last_error());
            if (ctx) hesaff_destroy(ctx);
            return 2;
         }
         hesaff_destroy(ctx);
      }
      // an unreadable file behaves like the reference: empty image -> "128\n0\n", exit code 0
      cout << "Detected " << nDetected << " keypoints and " << nAffine << " affine shapes in " << elapsed << " sec." << endl;
      string out = string(files[fi]) + ".hesaff.sift";
      if (!written && hesaff_write_sift_file(out.c_str(), keys.data(), keys.size(), par.desc_factor) < 0) {
         fprintf(stderr, "hesaff_b200: %s\n", hesaff_last_error());
         status = 1;
      }
   }
   return status;
}
