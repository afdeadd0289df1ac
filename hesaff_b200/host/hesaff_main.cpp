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
#include <thread>
#include <algorithm>
#include <dlfcn.h>
#include <cuda_runtime_api.h>
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

// One image through the C-ABI on `device`; mirrors main() of hesaff.cpp:133-180 for one file.  Returns 0, or 2 on a library error.
struct FileResult {
   int nDetected = 0, nAffine = 0;
   double elapsed = 0;
   int status = 0;
   string error;
};

static FileResult processFile(const char *path, int device, const hesaff_params &p, const HessianAffineParams &par)
{
   FileResult r;
   int w = 0, h = 0;
   vector<char> file;
   vector<hesaff_keypoint> keys;
   bool written = false;
   if (readFile(path, file) && hesaff_pnm_info(file.data(), file.size(), &w, &h, 0, 0) == HESAFF_OK) {
      hesaff_ctx *ctx = 0;
      int rc = hesaff_create(&ctx, &p, device, w, h, 1, 0);
      if (rc == HESAFF_OK) {
         // imread + the gray conversion (float(c0)+c1+c2)/3.0f of hesaff.cpp:137-148: header on the host, pixels on the GPU
         const void *fp = file.data();
         const size_t fb = file.size();
         const double t1 = wallTime();
         rc = hesaff_detect_pnm(ctx, &fp, &fb, 1, 0);
         if (rc == HESAFF_OK) {
            rc = hesaff_result_counts(ctx, &r.nDetected, &r.nAffine);
            keys.resize((size_t)hesaff_result_total(ctx));
            if (rc == HESAFF_OK && !keys.empty()) rc = hesaff_result_keypoints(ctx, keys.data(), keys.size());
            r.elapsed = wallTime() - t1;
         }
         // exportKeypoints (hesaff.cpp:169-173), with the text formatted on the GPU; the reference times only the
         // detector (hesaff.cpp:166-168), and so does the line main() prints
         if (rc == HESAFF_OK) {
            const string out = string(path) + ".hesaff.sift";
            if (hesaff_export_sift_file(ctx, 0, out.c_str()) >= 0) written = true;
            else rc = HESAFF_ERR_INVALID;
            if (rc == HESAFF_OK && getenv("HESAFF_BINARY_SIDECAR"))
               rc = hesaff_write_keypoints_binary((string(path) + ".hesaff.bin").c_str(), keys.data(), keys.size());
         }
      }
      if (rc != HESAFF_OK) {
         r.error = hesaff_last_error();
         r.status = 2;
      }
      if (ctx) hesaff_destroy(ctx);
      if (r.status) return r;
   }
   // an unreadable file behaves like the reference: empty image -> "128\n0\n", exit code 0
   if (!written && hesaff_write_sift_file((string(path) + ".hesaff.sift").c_str(), keys.data(), keys.size(), par.desc_factor) < 0) {
      r.error = hesaff_last_error();
      r.status = 1;
   }
   return r;
}

// ---- several GPUs of one node: images are independent, so the files are split into contiguous blocks, one
// host thread + one context per GPU, and the only exchange is ONE ncclAllGather of the per-image {detected, described}
// counts (single process, one communicator per device).  NCCL is loaded at run time, so the single-GPU tool does not need it.
typedef struct ncclComm *ncclComm_t;
struct NcclApi {
   void *lib = nullptr;
   int (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
   int (*CommDestroy)(ncclComm_t) = nullptr;
   int (*GroupStart)() = nullptr;
   int (*GroupEnd)() = nullptr;
   int (*AllGather)(const void *, void *, size_t, int, ncclComm_t, void *) = nullptr;
   const char *(*GetErrorString)(int) = nullptr;
   bool load()
   {
      lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
      if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
      if (!lib) return false;
      CommInitAll = (int (*)(ncclComm_t *, int, const int *))dlsym(lib, "ncclCommInitAll");
      CommDestroy = (int (*)(ncclComm_t))dlsym(lib, "ncclCommDestroy");
      GroupStart = (int (*)())dlsym(lib, "ncclGroupStart");
      GroupEnd = (int (*)())dlsym(lib, "ncclGroupEnd");
      AllGather = (int (*)(const void *, void *, size_t, int, ncclComm_t, void *))dlsym(lib, "ncclAllGather");
      GetErrorString = (const char *(*)(int))dlsym(lib, "ncclGetErrorString");
      return CommInitAll && CommDestroy && GroupStart && GroupEnd && AllGather;
   }
};

// counts[g] = this GPU's block of {detected, described} pairs -> every GPU (and the host) gets all blocks, in file order
static bool allGatherCounts(int G, const vector<vector<int>> &counts, size_t maxBlock, vector<int> &all, string &err)
{
   NcclApi nccl;
   if (!nccl.load()) { err = "libnccl.so.2 not found"; return false; }
   vector<int> devs(G);
   for (int g = 0; g < G; g++) devs[g] = g;
   vector<ncclComm_t> comms(G);
   int rc = nccl.CommInitAll(comms.data(), G, devs.data());
   if (rc) { err = string("ncclCommInitAll: ") + (nccl.GetErrorString ? nccl.GetErrorString(rc) : "error"); return false; }
   const size_t n = 2 * maxBlock;             // int32 per rank (blocks padded to the largest)
   vector<int *> send(G), recv(G);
   vector<cudaStream_t> st(G);
   bool ok = true;
   for (int g = 0; g < G && ok; g++) {
      ok = cudaSetDevice(g) == cudaSuccess && cudaStreamCreate(&st[g]) == cudaSuccess &&
           cudaMalloc((void **)&send[g], n * sizeof(int)) == cudaSuccess && cudaMalloc((void **)&recv[g], n * G * sizeof(int)) == cudaSuccess;
      if (!ok) break;
      vector<int> pad(n, 0);
      std::copy(counts[g].begin(), counts[g].end(), pad.begin());
      ok = cudaMemcpy(send[g], pad.data(), n * sizeof(int), cudaMemcpyHostToDevice) == cudaSuccess;
   }
   if (ok) {
      nccl.GroupStart();
      for (int g = 0; g < G; g++) {
         cudaSetDevice(g);
         rc = nccl.AllGather(send[g], recv[g], n, /* ncclInt32 */ 2, comms[g], st[g]);
         if (rc) ok = false;
      }
      rc = nccl.GroupEnd();
      if (rc) ok = false;
      for (int g = 0; g < G; g++) { cudaSetDevice(g); if (cudaStreamSynchronize(st[g]) != cudaSuccess) ok = false; }
   }
   if (ok) {
      // every rank holds the same table; rank G-1's copy is the one read back (a cheap check that the data really travelled)
      all.resize(n * G);
      cudaSetDevice(G - 1);
      ok = cudaMemcpy(all.data(), recv[G - 1], n * G * sizeof(int), cudaMemcpyDeviceToHost) == cudaSuccess;
   }
   if (!ok) err = "NCCL all-gather of the keypoint counts failed";
   for (int g = 0; g < G; g++) {
      cudaSetDevice(g);
      if (send[g]) cudaFree(send[g]);
      if (recv[g]) cudaFree(recv[g]);
      if (st[g]) cudaStreamDestroy(st[g]);
      nccl.CommDestroy(comms[g]);
   }
   return ok;
}

int main(int argc, char **argv)
{
   HessianAffineParams par;
   hesaff_params p;
   hesaff_params_default(&p);
   int device = 0, gpus = 1;
   vector<const char *> files;
   for (int i = 1; i < argc; i++) {
      if (!strcmp(argv[i], "--threshold") && i + 1 < argc) par.threshold = (float)atof(argv[++i]);
      else if (!strcmp(argv[i], "--scales") && i + 1 < argc) p.number_of_scales = atoi(argv[++i]);
      else if (!strcmp(argv[i], "--max-octaves") && i + 1 < argc) p.max_octaves = atoi(argv[++i]);
      else if (!strcmp(argv[i], "--device") && i + 1 < argc) device = atoi(argv[++i]);
      else if (!strcmp(argv[i], "--gpus") && i + 1 < argc) gpus = atoi(argv[++i]);
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
   if (gpus <= 1) {
      for (size_t fi = 0; fi < files.size(); fi++) {
         const FileResult r = processFile(files[fi], device, p, par);
         if (r.status == 2) { fprintf(stderr, "hesaff_b200: %s\n", r.error.c_str()); return 2; }
         cout << "Detected " << r.nDetected << " keypoints and " << r.nAffine << " affine shapes in " << r.elapsed << " sec." << endl;
         if (r.status) { fprintf(stderr, "hesaff_b200: %s\n", r.error.c_str()); status = 1; }
      }
      return status;
   }

   // ---- image-parallel over `gpus` devices --------------------------------------------------------------------------
   const int G = gpus;
   const size_t n = files.size();
   vector<size_t> start(G + 1, 0);
   for (int g = 0; g < G; g++) start[g + 1] = start[g] + n / G + ((size_t)g < n % G ? 1 : 0);     // contiguous blocks
   vector<vector<FileResult>> res(G);
   vector<std::thread> th;
   for (int g = 0; g < G; g++)
      th.emplace_back([&, g]() {
         for (size_t fi = start[g]; fi < start[g + 1]; fi++) res[g].push_back(processFile(files[fi], g, p, par));
      });
   for (auto &t : th) t.join();
   size_t maxBlock = 0;
   vector<vector<int>> counts(G);
   for (int g = 0; g < G; g++) {
      maxBlock = std::max(maxBlock, res[g].size());
      for (const FileResult &r : res[g]) {
         if (r.status == 2) { fprintf(stderr, "hesaff_b200: %s\n", r.error.c_str()); return 2; }
         if (r.status) { fprintf(stderr, "hesaff_b200: %s\n", r.error.c_str()); status = 1; }
         counts[g].push_back(r.nDetected);
         counts[g].push_back(r.nAffine);
      }
   }
   vector<int> all;
   string err;
   if (!allGatherCounts(G, counts, maxBlock, all, err)) { fprintf(stderr, "hesaff_b200: %s\n", err.c_str()); return 2; }
   long long totalDet = 0, totalAff = 0;
   for (int g = 0; g < G; g++)
      for (size_t k = 0; k < res[g].size(); k++) {
         const int nd = all[(size_t)g * 2 * maxBlock + 2 * k], na = all[(size_t)g * 2 * maxBlock + 2 * k + 1];
         cout << "Detected " << nd << " keypoints and " << na << " affine shapes in " << res[g][k].elapsed << " sec." << endl;
         totalDet += nd; totalAff += na;
      }
   cout << "Total over " << n << " images on " << G << " GPUs: " << totalDet << " keypoints and " << totalAff
        << " affine shapes (counts all-gathered with NCCL)." << endl;
   return status;
}
