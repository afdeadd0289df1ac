// hesaff_b200/csrc/export.cu -- GPU-side text formatting of the .hesaff.sift file.
//
// Replaces the formatting loop of AffineHessianDetector::exportKeypoints (hesaff.cpp:107-130): one line per keypoint,
//     x y a b c d1 ... d128 \n
// where the five floats go through `ostream << float` with default flags, i.e. printf("%g") with precision 6 on the
// float promoted to double, and the descriptor bytes are printed as ints.  At GPU detection rates the host's ostream
// loop (about 400 bytes and 133 conversions per keypoint) dominates the run time of the drop-in CLI, so the lines are
// produced on the device: one warp per keypoint, a length pass, an exclusive scan, a write pass.
//
// fmt_g6 reproduces glibc's correctly rounded conversion exactly (round-half-even on the exact binary value) with a
// small fixed-size big integer, for every finite float below 2^63; anything else (inf, nan, >= 2^63) raises a flag and
// the caller formats that image on the host instead.
#include "common.cuh"

namespace {

// 192-bit unsigned integer, little-endian 32-bit limbs; loops are fully unrolled so it lives in registers
struct Big {
   uint32_t w[6];
};

__device__ __forceinline__ void big_mul_small(Big &b, uint32_t k)
{
   unsigned long long carry = 0;
#pragma unroll
   for (int i = 0; i < 6; i++) {
      const unsigned long long t = (unsigned long long)b.w[i] * k + carry;
      b.w[i] = (uint32_t)t;
      carry = t >> 32;
   }
}

__device__ __forceinline__ uint32_t big_bit(const Big &b, int pos)   // pos >= 0; bits beyond 191 are 0
{
   uint32_t v = 0;
#pragma unroll
   for (int i = 0; i < 6; i++)
      if ((pos >> 5) == i) v = b.w[i];
   return (v >> (pos & 31)) & 1u;
}

// bits [pos, pos+32) of b
__device__ __forceinline__ uint32_t big_extract32(const Big &b, int pos)
{
   const int li = pos >> 5, sh = pos & 31;
   uint32_t lo = 0, hi = 0;
#pragma unroll
   for (int i = 0; i < 6; i++) {
      if (li == i) lo = b.w[i];
      if (li + 1 == i) hi = b.w[i];
   }
   return sh ? (lo >> sh) | (hi << (32 - sh)) : lo;
}

// any bit below pos set?
__device__ __forceinline__ bool big_any_below(const Big &b, int pos)
{
   bool any = false;
#pragma unroll
   for (int i = 0; i < 6; i++) {
      if (32 * (i + 1) <= pos) any |= b.w[i] != 0;
      else if (32 * i < pos) any |= (b.w[i] & ((1u << (pos - 32 * i)) - 1u)) != 0;
   }
   return any;
}

// round-half-even(m * 2^q * 10^s) for 0 <= s <= 55 (m < 2^24); the result is known to be < 2^31
__device__ uint32_t scaled_round_pos(uint32_t m, int q, int s)
{
   Big n;
   n.w[0] = m; n.w[1] = n.w[2] = n.w[3] = n.w[4] = n.w[5] = 0;
   int left = s;
   while (left >= 13) { big_mul_small(n, 1220703125u); left -= 13; }   // 5^13
   uint32_t p = 1;
   for (int i = 0; i < left; i++) p *= 5u;
   big_mul_small(n, p);
   const int sh = q + s;                           // value = n * 2^sh
   if (sh >= 0) return n.w[0] << sh;               // an integer below 10^7: no rounding
   const int r = -sh;
   if (r > 192) return 0;
   uint32_t d = r < 192 ? big_extract32(n, r) : 0u;
   const bool half = big_bit(n, r - 1) != 0;
   const bool sticky = big_any_below(n, r - 1);
   if (half && (sticky || (d & 1u))) d++;
   return d;
}

// round-half-even(m * 2^q / 10^t) for t >= 1 and m * 2^q < 2^63
__device__ uint32_t scaled_round_neg(uint32_t m, int q, int t)
{
   unsigned long long num = q >= 0 ? (unsigned long long)m << q : (unsigned long long)m >> (-q);   // exact: values >= 10^6 have q > -24... see caller
   unsigned long long den = 1;
   for (int i = 0; i < t; i++) den *= 10ull;
   unsigned long long d = num / den;
   const unsigned long long rem = num - d * den;
   if (2 * rem > den || (2 * rem == den && (d & 1ull))) d++;
   return (uint32_t)d;
}

// printf("%g") of (double)f, precision 6.  Writes at most 13 characters; returns the length, or -1 for values the
// device path does not cover (inf, nan, |f| >= 2^63).
__device__ int fmt_g6(float f, char *o)
{
   uint32_t bits = __float_as_uint(f);
   int n = 0;
   if (bits >> 31) o[n++] = '-';
   bits &= 0x7fffffffu;
   if (bits == 0) { o[n++] = '0'; return n; }
   if (bits >= 0x5f000000u) return -1;              // >= 2^63, inf, nan
   const int ex = (int)(bits >> 23);
   uint32_t m = bits & 0x7fffffu;
   int q;
   if (ex == 0) q = -149; else { m |= 0x800000u; q = ex - 150; }      // |f| = m * 2^q
   const int e2 = q + 31 - __clz(m);                                   // floor(log2 |f|)
   int e10 = (int)floorf((float)e2 * 0.30103f);                        // floor(log10 2^e2): true exponent or one below
   if (e10 < -46) e10 = -46;
   uint32_t d = 0;
   for (int attempt = 0; attempt < 3; attempt++) {
      const int s = 5 - e10;
      // for s < 0 the value is >= 10^6 > 2^19, so q >= -4 and the right shift of m in scaled_round_neg drops nothing
      // only when q >= 0; keep it exact by folding the negative q into the divisor instead
      if (s >= 0) d = scaled_round_pos(m, q, s);
      else if (q >= 0) d = scaled_round_neg(m, q, -s);
      else {                                                            // m * 2^q / 10^t = m / (10^t * 2^-q), q in [-4, -1]
         unsigned long long den = 1ull << (-q);
         for (int i = 0; i < -s; i++) den *= 10ull;
         unsigned long long dd = (unsigned long long)m / den;
         const unsigned long long rem = (unsigned long long)m - dd * den;
         if (2 * rem > den || (2 * rem == den && (dd & 1ull))) dd++;
         d = (uint32_t)dd;
      }
      if (d >= 1000000u) { e10++; continue; }       // estimate one too low, or 999999.5 rounded up to 10^6
      if (d < 100000u) { e10--; continue; }
      break;
   }
   char dig[6];
#pragma unroll
   for (int i = 5; i >= 0; i--) { dig[i] = (char)('0' + d % 10u); d /= 10u; }
   int nd = 6;
   while (nd > 1 && dig[nd - 1] == '0') nd--;                          // %g removes trailing zeros
   if (e10 < -4 || e10 >= 6) {                                         // scientific: d[.ddddd]e+XX
      o[n++] = dig[0];
      if (nd > 1) {
         o[n++] = '.';
         for (int i = 1; i < nd; i++) o[n++] = dig[i];
      }
      o[n++] = 'e';
      int e = e10;
      if (e < 0) { o[n++] = '-'; e = -e; } else o[n++] = '+';
      if (e >= 100) { o[n++] = (char)('0' + e / 100); e %= 100; }
      o[n++] = (char)('0' + e / 10);
      o[n++] = (char)('0' + e % 10);
   } else if (e10 >= 0) {                                              // ddd[.ddd]
      for (int i = 0; i <= e10; i++) o[n++] = i < nd ? dig[i] : '0';
      if (nd > e10 + 1) {
         o[n++] = '.';
         for (int i = e10 + 1; i < nd; i++) o[n++] = dig[i];
      }
   } else {                                                            // 0.000ddd
      o[n++] = '0'; o[n++] = '.';
      for (int i = 0; i < -e10 - 1; i++) o[n++] = '0';
      for (int i = 0; i < nd; i++) o[n++] = dig[i];
   }
   return n;
}

__device__ __forceinline__ int fmt_u8(unsigned v, char *o)   // " ddd"
{
   int n = 0;
   o[n++] = ' ';
   if (v >= 100) { o[n++] = (char)('0' + v / 100); v %= 100; o[n++] = (char)('0' + v / 10); o[n++] = (char)('0' + v % 10); }
   else if (v >= 10) { o[n++] = (char)('0' + v / 10); o[n++] = (char)('0' + v % 10); }
   else o[n++] = (char)('0' + v);
   return n;
}

}   // namespace

// One warp per keypoint.  Lanes 0..4 format x, y, a, b, c (a leading space on all but x); every lane formats four
// descriptor bytes; a warp prefix sum places the pieces.  WRITE = false: only the line length is stored.
template <bool WRITE>
__global__ void __launch_bounds__(128) k_sift_text(const hesaff_keypoint *__restrict__ keys, const float *__restrict__ ell,
                                                   uint32_t n, uint32_t *__restrict__ len, const uint32_t *__restrict__ off,
                                                   char *__restrict__ text, int *bad)
{
   const int lane = threadIdx.x & 31;
   const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
   for (uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += nwarps) {
      char f[16], dd[16];
      int fl = 0;
      if (lane < 5) {
         if (lane > 0) f[fl++] = ' ';
         const int l = fmt_g6(ell[(size_t)i * 5 + lane], f + fl);     // (u, v, a, b, c): u = x, v = y
         if (l < 0) { *bad = 1; } else fl += l;
      }
      int dl = 0;
      {
         const uint32_t v4 = reinterpret_cast<const uint32_t *>(keys[i].desc)[lane];
#pragma unroll
         for (int j = 0; j < 4; j++) dl += fmt_u8((v4 >> (8 * j)) & 255u, dd + dl);
      }
      // exclusive prefix sums over the warp: floats first (lanes 0..4), then the descriptor pieces
      int fo = fl, dn = dl;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
         const int a = __shfl_up_sync(0xffffffffu, fo, d), b = __shfl_up_sync(0xffffffffu, dn, d);
         if (lane >= d) { fo += a; dn += b; }
      }
      const int ftot = __shfl_sync(0xffffffffu, fo, 31), dtot = __shfl_sync(0xffffffffu, dn, 31);
      if (!WRITE) {
         if (lane == 0) len[i] = (uint32_t)(ftot + dtot + 1);
      } else {
         char *line = text + off[i];
         char *p = line + (fo - fl);
         for (int k = 0; k < fl; k++) p[k] = f[k];
         p = line + ftot + (dn - dl);
         for (int k = 0; k < dl; k++) p[k] = dd[k];
         if (lane == 31) line[ftot + dtot] = '\n';
      }
   }
}

void ha_launch_sift_text(bool write, const hesaff_keypoint *keys, const float *ell, uint32_t n, uint32_t *len,
                         const uint32_t *off, char *text, int *bad, cudaStream_t st, LaunchCounter &lc)
{
   if (!n) return;
   const unsigned blocks = (unsigned)std::min<size_t>(((size_t)n * 32 + 127) / 128, 148 * 16);
   if (write) k_sift_text<true><<<blocks, 128, 0, st>>>(keys, ell, n, len, off, text, bad);
   else k_sift_text<false><<<blocks, 128, 0, st>>>(keys, ell, n, len, off, text, bad);
   lc.n++;
}

// diagnostic: formats n floats into 16-byte NUL-padded slots (tests compare with the host's "%g")
__global__ void k_format_floats(const float *__restrict__ in, size_t n, char *__restrict__ out)
{
   const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n) return;
   char b[16];
   int l = fmt_g6(in[i], b);
   if (l < 0) { b[0] = '?'; l = 1; }
   for (int k = 0; k < 16; k++) out[i * 16 + k] = k < l ? b[k] : '\0';
}

void ha_launch_format_floats(const float *in, size_t n, char *out, cudaStream_t st)
{
   if (!n) return;
   k_format_floats<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(in, n, out);
}
