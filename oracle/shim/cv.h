// oracle/shim/cv.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A minimal stand-in for the legacy OpenCV umbrella header <cv.h> so that the
// reference sources under /root/reference (pyramid.cpp, affine.cpp, siftdesc.cpp,
// helpers.cpp, hesaff.cpp) compile UNMODIFIED, in place, without OpenCV.
// Only the cv:: surface those five files touch is provided (grep of every use:
// SURVEY.md section 8(c)).  The one piece of arithmetic on the hot path that
// lives in OpenCV, cv::GaussianBlur (helpers.cpp:287,294), is restated in
// cv_shim.cpp and pinned bit-for-bit against the real OpenCV 4.13 (cv2) in
// tests/test_oracle_blur.py.
#ifndef HESAFF_ORACLE_CV_SHIM_H
#define HESAFF_ORACLE_CV_SHIM_H

#include <cassert>
#include <cmath>
#include <math.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#define CV_8U 0
#define CV_32F 5
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << 3))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)

namespace cv {

typedef unsigned char uchar;

struct Scalar {
   double val[4];
   Scalar(double v0 = 0) { val[0] = v0; val[1] = val[2] = val[3] = 0; }
};

struct Size {
   int width, height;
   Size(int w = 0, int h = 0) : width(w), height(h) {}
};

enum { BORDER_REPLICATE = 1 };

class Mat {
public:
   int rows, cols;
   size_t step;   // bytes per row
   uchar *data;

   Mat() : rows(0), cols(0), step(0), data(0), type_(0), refcount_(0) {}
   Mat(int r, int c, int type) { create(r, c, type); }
   Mat(int r, int c, int type, const Scalar &s) { create(r, c, type); *this = s; }
   // external storage, not owned (affine.cpp:124)
   Mat(int r, int c, int type, void *ext)
      : rows(r), cols(c), step((size_t)c * elemSize(type)), data((uchar *)ext), type_(type), refcount_(0) {}
   Mat(const Mat &m)
      : rows(m.rows), cols(m.cols), step(m.step), data(m.data), type_(m.type_), refcount_(m.refcount_)
   {
      if (refcount_) ++*refcount_;
   }
   ~Mat() { release(); }
   Mat &operator=(const Mat &m)
   {
      if (this != &m) {
         if (m.refcount_) ++*m.refcount_;
         release();
         rows = m.rows; cols = m.cols; step = m.step; data = m.data; type_ = m.type_; refcount_ = m.refcount_;
      }
      return *this;
   }
   Mat &operator=(const Scalar &s)
   {
      if ((type_ & 7) == CV_32F) {
         float v = (float)s.val[0];
         for (int r = 0; r < rows; r++) { float *p = ptr<float>(r); for (int c = 0; c < cols * channels(); c++) p[c] = v; }
      } else {
         uchar v = (uchar)s.val[0];
         for (int r = 0; r < rows; r++) { uchar *p = ptr<uchar>(r); for (int c = 0; c < cols * channels(); c++) p[c] = v; }
      }
      return *this;
   }
   int type() const { return type_; }
   int channels() const { return (type_ >> 3) + 1; }
   bool empty() const { return data == 0 || rows * cols == 0; }
   Mat clone() const
   {
      Mat m;
      if (data) { m.create(rows, cols, type_); for (int r = 0; r < rows; r++) memcpy(m.data + r * m.step, data + r * step, m.step); }
      return m;
   }
   template <typename T> T *ptr(int r = 0) { return (T *)(data + (size_t)r * step); }
   template <typename T> const T *ptr(int r = 0) const { return (const T *)(data + (size_t)r * step); }
   template <typename T> T &at(int r, int c) { return ((T *)(data + (size_t)r * step))[c]; }
   template <typename T> const T &at(int r, int c) const { return ((const T *)(data + (size_t)r * step))[c]; }

   static Mat zeros(int r, int c, int type)
   {
      Mat m(r, c, type);
      if (m.data) memset(m.data, 0, (size_t)r * m.step);
      return m;
   }
   // column vector -> diagonal matrix (hesaff.cpp:123)
   static Mat diag(const Mat &d)
   {
      int n = d.rows * d.cols;
      Mat m = zeros(n, n, d.type());
      for (int i = 0; i < n; i++) m.at<float>(i, i) = ((const float *)d.data)[i];
      return m;
   }
   Mat t() const
   {
      Mat m(cols, rows, type_);
      for (int r = 0; r < rows; r++) for (int c = 0; c < cols; c++) m.at<float>(c, r) = at<float>(r, c);
      return m;
   }
   void create(int r, int c, int type)
   {
      rows = r; cols = c; type_ = type; step = (size_t)c * elemSize(type);
      size_t bytes = (size_t)r * step;
      if (bytes == 0) { data = 0; refcount_ = 0; return; }
      data = (uchar *)malloc(bytes);   // like cv::Mat: uninitialised storage
      refcount_ = new int(1);
   }
   static size_t elemSize(int type) { return (size_t)(((type & 7) == CV_32F) ? 4 : 1) * ((type >> 3) + 1); }

private:
   void release()
   {
      if (refcount_ && --*refcount_ == 0) { free(data); delete refcount_; }
      data = 0; refcount_ = 0;
   }
   int type_;
   int *refcount_;
};

inline Mat operator*(const Mat &a, const Mat &b)
{
   Mat m = Mat::zeros(a.rows, b.cols, a.type());
   for (int r = 0; r < a.rows; r++)
      for (int c = 0; c < b.cols; c++) {
         float s = 0;
         for (int k = 0; k < a.cols; k++) s += a.at<float>(r, k) * b.at<float>(k, c);
         m.at<float>(r, c) = s;
      }
   return m;
}

template <typename T> class Mat_ : public Mat {
public:
   Mat_(int r, int c) : Mat(r, c, CV_32FC1) {}
};

template <typename T> class MatCommaInitializer_ {
public:
   MatCommaInitializer_(const Mat_<T> &m) : m_(m), i_(0) {}
   MatCommaInitializer_<T> &operator,(T v) { put(v); return *this; }
   void put(T v) { ((T *)m_.data)[i_++] = v; }
   operator Mat() const { return m_; }
private:
   Mat m_;
   int i_;
};
template <typename T> inline MatCommaInitializer_<T> operator<<(const Mat_<T> &m, T v)
{
   MatCommaInitializer_<T> ci(m);
   ci.put(v);
   return ci;
}

// 2x2 only (hesaff.cpp:117): A = u * diag(w) * vt, w descending, w >= 0.
class SVD {
public:
   enum { FULL_UV = 4 };
   Mat u, w, vt;
   SVD(const Mat &A, int flags = 0);
};

void GaussianBlur(const Mat &src, Mat &dst, Size ksize, double sigmaX, double sigmaY = 0, int borderType = BORDER_REPLICATE);

// PNM (P5/P6, maxval 255) reader returning 8-bit 3-channel BGR like cv::imread's default.
Mat imread(const std::string &path);

} // namespace cv

#endif
