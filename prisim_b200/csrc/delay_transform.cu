// Windowed delay transform along frequency, batched over rows (baselines of one snapshot).
// Replaces interferometry.py:8114-8134 and delay_spectrum.py:1305-1327 of the reference:
//   X = fftshift(ifft(pad_right(x * bp * wts, npad))) * (nchan + npad) * df,
//   then DSP.downsampler(X, 1+pad) = linear interpolation at positions arange(0, n, 1+pad).
//
// Three hand-written kernels (no cuFFT):
//   k_delay_fft_r8 register-resident radix-8 Stockham inverse FFT (fp64), global -> registers ->
//                 (smem between passes) -> global with the bp*wts multiply, scale and fftshift
//                 fused; power-of-two lengths 64..2048 without interpolation.  When 1+pad is an
//                 integer m and nchan is even, decimating the m*nchan-point padded transform by m
//                 is exactly the nchan-point transform, so that one is computed (default pad=1).
//   k_delay_fft   radix-2 in-smem inverse FFT with linear-interpolated decimation on store: the
//                 remaining power-of-two cases (non-integer decimation, N < 64, N >= 4096) and
//                 lengths 3 * 2^k (pad = 0.5): three interleaved 2^k-point transforms + radix-3 combine.
//   k_delay_dft   direct evaluation of just the output samples needed (any length, any pad):
//                 O(nout * nchan) per row, twiddles from an exact integer-indexed table.
// HBM-bound: algorithmic bytes per row = nchan*(16 + 8 + 8) read + nout*16 written.
#include "common.cuh"
#include <cmath>
#include <cstdlib>

namespace {

constexpr int FFT_THREADS = 256;

struct DelayParams {
  const double2* x;          // [nrows, nchan] or null
  const double* bp;          // null -> 1
  const double* wts;         // null -> 1
  long long bp_stride, wts_stride;
  double2* out;              // [nrows, nout]
  const double2* twiddle;    // exp(+2 pi i j / nfft), j < nfft
  int nrows, nchan, nfft, log2n, nout;
  int radix3;                // nfft = 3 * 2^log2n (k_delay_fft only): three interleaved 2^log2n-point transforms + a radix-3 combine
  int shift;                 // (nfft+1)/2: fftshift(X)[j] = X[(j + shift) % nfft]
  double scale;              // df  ( = (1/nfft) * nfft * df )
  double factor;             // decimation step in units of the nfft grid (1 = none)
};

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

__device__ __forceinline__ double2 load_in(const DelayParams& P, int row, int n) {
  if (n >= P.nchan) return make_double2(0.0, 0.0);                 // zero padding (:8124)
  double w = 1.0;
  if (P.bp) w *= P.bp[(size_t)row * P.bp_stride + n];
  if (P.wts) w *= P.wts[(size_t)row * P.wts_stride + n];
  if (!P.x) return make_double2(w, 0.0);                           // lag_kernel (:8127)
  double2 v = P.x[(size_t)row * P.nchan + n];
  return make_double2(v.x * w, v.y * w);
}

__global__ void __launch_bounds__(FFT_THREADS) k_delay_fft(const DelayParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double2* buf = reinterpret_cast<double2*>(smem_raw);
  const int row = blockIdx.x;
  const int N = P.nfft;
  const int NSUB = P.radix3 ? 3 : 1;            // interleaved sub-sequences x[NSUB m + r], each M = N / NSUB = 2^log2n points
  const int M = N / NSUB;
  // decimation in time: sub-sequence r goes to buf[r M ...] in bit-reversed order
  for (int n = threadIdx.x; n < N; n += FFT_THREADS) {
    const int r = n % NSUB, m = n / NSUB;
    const int br = P.log2n > 0 ? (int)(__brev((unsigned)m) >> (32 - P.log2n)) : 0;
    buf[r * M + br] = load_in(P, row, n);
  }
  __syncthreads();
  for (int st = 1; st <= P.log2n; ++st) {
    const int half = 1 << (st - 1);
    const int tstride = NSUB * (M >> st);        // exp(+2 pi i j / 2^st) = twiddle[j * N / 2^st]
    for (int i = threadIdx.x; i < N / 2; i += FFT_THREADS) {
      const int sub = i / (M / 2), ii = i - sub * (M / 2);
      int j = ii & (half - 1);
      int base = sub * M + ((ii - j) << 1) + j;
      double2 w = P.twiddle[j * tstride];
      double2 a = buf[base], bb = cmul(buf[base + half], w);
      buf[base] = make_double2(a.x + bb.x, a.y + bb.y);
      buf[base + half] = make_double2(a.x - bb.x, a.y - bb.y);
    }
    __syncthreads();
  }
  if (P.radix3) {
    // X[k + s M] = Y0[k] + w3^s W^k Y1[k] + w3^(2s) W^(2k) Y2[k],  W = exp(+2 pi i / N), w3 = exp(+2 pi i / 3); in place per k
    const double hs = 0.86602540378443864676;    // sin(2 pi / 3)
    for (int k = threadIdx.x; k < M; k += FFT_THREADS) {
      const double2 a = buf[k];
      const double2 b = cmul(buf[M + k], P.twiddle[k]);
      const double2 c = cmul(buf[2 * M + k], P.twiddle[(2 * k) % N]);
      const double2 sum = make_double2(b.x + c.x, b.y + c.y), dif = make_double2(b.x - c.x, b.y - c.y);
      const double2 re = make_double2(a.x - 0.5 * sum.x, a.y - 0.5 * sum.y);           // a + (b + c) cos(2 pi/3)
      const double2 im = make_double2(-hs * dif.y, hs * dif.x);                         // i sin(2 pi/3) (b - c)
      buf[k] = make_double2(a.x + sum.x, a.y + sum.y);
      buf[M + k] = make_double2(re.x + im.x, re.y + im.y);
      buf[2 * M + k] = make_double2(re.x - im.x, re.y - im.y);
    }
    __syncthreads();
  }
  // scale, fftshift, linear-interpolated decimation (:8131-8134)
  for (int i = threadIdx.x; i < P.nout; i += FFT_THREADS) {
    double pos = i * P.factor;
    int j0 = (int)floor(pos);
    double fr = pos - j0;
    double2 a = buf[(j0 + P.shift) % N];
    double2 o;
    if (fr > 0.0) {
      if (j0 + 1 < N) {
        double2 c = buf[(j0 + 1 + P.shift) % N];
        o = make_double2(a.x + (c.x - a.x) * fr, a.y + (c.y - a.y) * fr);
      } else {
        o = make_double2(nan(""), nan(""));                        // outside the interpolation range
      }
    } else {
      o = a;
    }
    P.out[(size_t)row * P.nout + i] = make_double2(o.x * P.scale, o.y * P.scale);
  }
}


// ---------------------------------------------------------------------------------------------
// k_delay_fft_r8: register-resident mixed-radix (8,8,...,{8,4,2}) Stockham autosort inverse FFT.
// The first pass reads global memory directly (fused bp*wts multiply and zero padding), the last
// pass writes global memory directly (fused scale + fftshift); only the passes in between go
// through shared memory (one padded row per transform, in place: read, barrier, write, barrier --
// 5 barriers for N=1024 instead of the 10 of the radix-2 kernel, and half the shared memory of a
// ping-pong scheme, so 6 CTAs = 12 rows are resident per SM).  One row = N/8 threads; a CTA carries 256/(N/8) rows (>= 1).
// Used for 64 <= N <= 2048 when no interpolation is needed (factor == 1).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 cmuli(double2 a) { return make_double2(-a.y, a.x); }          // * (+i)

template <int R> __device__ __forceinline__ void ifft_small(double2 (&v)[R]);
template <> __device__ __forceinline__ void ifft_small<2>(double2 (&v)[2]) {
  double2 a = v[0], b = v[1];
  v[0] = cadd(a, b); v[1] = csub(a, b);
}
template <> __device__ __forceinline__ void ifft_small<4>(double2 (&v)[4]) {
  double2 s0 = cadd(v[0], v[2]), s1 = csub(v[0], v[2]), s2 = cadd(v[1], v[3]), s3 = cmuli(csub(v[1], v[3]));
  v[0] = cadd(s0, s2); v[1] = cadd(s1, s3); v[2] = csub(s0, s2); v[3] = csub(s1, s3);
}
template <> __device__ __forceinline__ void ifft_small<8>(double2 (&v)[8]) {
  const double h = 0.70710678118654752440;
  double2 a[4], b[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { a[i] = cadd(v[i], v[i + 4]); b[i] = csub(v[i], v[i + 4]); }
  b[1] = make_double2(h * (b[1].x - b[1].y), h * (b[1].x + b[1].y));        // * exp(+i pi/4)
  b[2] = cmuli(b[2]);                                                       // * exp(+i pi/2)
  b[3] = make_double2(-h * (b[3].x + b[3].y), h * (b[3].x - b[3].y));       // * exp(+3 i pi/4)
  ifft_small<4>(a);
  ifft_small<4>(b);
#pragma unroll
  for (int i = 0; i < 4; ++i) { v[2 * i] = a[i]; v[2 * i + 1] = b[i]; }
}

// w^1 .. w^(R-1) from ONE table load: the kernel is bound by L1/TEX requests (21 twiddle loads per thread were a
// quarter of them) while the FP64 pipe idles at 16 %, so the powers are formed by 6 complex multiplications
// (squares where possible; error <= 3 ulp)
template <int R> struct TwiddlePowers {
  double2 w[R];
  __device__ __forceinline__ explicit TwiddlePowers(const double2 w1) {
    w[0] = make_double2(1.0, 0.0);
    if (R > 1) w[1] = w1;
    if (R > 2) w[2] = cmul(w1, w1);
    if (R > 3) w[3] = cmul(w[2], w1);
    if (R > 4) w[4] = cmul(w[2], w[2]);
    if (R > 5) w[5] = cmul(w[4], w1);
    if (R > 6) w[6] = cmul(w[3], w[3]);
    if (R > 7) w[7] = cmul(w[6], w1);
  }
};

// branch-free loader (all loads of a butterfly in flight together): bp / wts always valid pointers
// (the host substitutes a broadcast "ones" row), index clamped for the zero-padded tail
template <bool HAS_X>
__device__ __forceinline__ double2 load_in_nb(const DelayParams& P, int row, int n) {
  const int nc = n < P.nchan ? n : P.nchan - 1;
  double w = P.bp[(size_t)row * P.bp_stride + nc];
  if (P.wts) w *= P.wts[(size_t)row * P.wts_stride + nc];                 // uniform predicate; null = already folded into bp
  double2 v = make_double2(1.0, 0.0);
  if (HAS_X) v = P.x[(size_t)row * P.nchan + nc];
  const double m = n < P.nchan ? w : 0.0;
  return make_double2(v.x * m, v.y * m);
}

__device__ __forceinline__ int padi(int i) { return i + (i >> 3); }   // 16 B of padding per 128 B: the stride-8 stores of the first pass hit distinct banks

// One Stockham pass of radix R for the butterflies j = t, t+T, ... of one row.  `first` reads global
// memory (window multiply, zero padding), `last` writes global memory (scale, fftshift); passes in
// between work IN PLACE on one padded shared-memory row: all reads, a CTA barrier, then all writes
// (middle passes are radix 8 with exactly one butterfly per thread, so the inputs sit in registers).
template <int R, bool HAS_X>
__device__ __forceinline__ void stockham_pass(const DelayParams& P, int row, bool active, int t, int T, int Ns, bool first,
                                              bool last, double2* __restrict__ buf) {
  const int N = P.nfft, NR = N / R;
  if (last) {                                   // smem (or global, single-pass transforms) -> global
    if (active)
      for (int j = t; j < NR; j += T) {
        double2 v[R];
#pragma unroll
        for (int r = 0; r < R; ++r) v[r] = first ? load_in_nb<HAS_X>(P, row, j + r * NR) : buf[padi(j + r * NR)];
        const int k = j & (Ns - 1);
        if (Ns > 1) {
          const TwiddlePowers<R> tw(__ldg(&P.twiddle[(k * (N / (Ns * R))) & (N - 1)]));
#pragma unroll
          for (int r = 1; r < R; ++r) v[r] = cmul(v[r], tw.w[r]);
        }
        ifft_small<R>(v);
        const int j0 = (j - k) * R + k;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          int i = j0 + r * Ns - P.shift; if (i < 0) i += N;                  // fftshift: out[i] = X[(i + shift) % N]
          P.out[(size_t)row * P.nout + i] = make_double2(v[r].x * P.scale, v[r].y * P.scale);
        }
      }
    return;
  }
  // first / middle pass: one butterfly per thread (T == NR)
  double2 v[R];
  const int j = t;
  if (active) {
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = first ? load_in_nb<HAS_X>(P, row, j + r * NR) : buf[padi(j + r * NR)];
  }
  if (!first) __syncthreads();                  // everyone has read this pass's inputs
  if (active) {
    const int k = j & (Ns - 1);
    if (Ns > 1) {
      const TwiddlePowers<R> tw(__ldg(&P.twiddle[(k * (N / (Ns * R))) & (N - 1)]));
#pragma unroll
      for (int r = 1; r < R; ++r) v[r] = cmul(v[r], tw.w[r]);
    }
    ifft_small<R>(v);
    const int j0 = (j - k) * R + k;
#pragma unroll
    for (int r = 0; r < R; ++r) buf[padi(j0 + r * Ns)] = v[r];
  }
  __syncthreads();                              // outputs visible to the next pass
}

#ifndef PB_DT_MINBLOCKS     // resident CTAs per SM asked of ptxas: 2 -> 128 registers, no spills (0.67 ms at C2 size); 3 / 4 -> 80 / 64
#define PB_DT_MINBLOCKS 2   // registers with 600-850 bytes of spills, 1.35 / 1.32 ms
#endif
template <bool HAS_X>
__global__ void __launch_bounds__(256, PB_DT_MINBLOCKS) k_delay_fft_r8(const DelayParams P, int T, int rows_per_cta) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int N = P.nfft;
  const int lrow = threadIdx.x / T, t = threadIdx.x % T;
  const int row = blockIdx.x * rows_per_cta + lrow;
  const bool active = lrow < rows_per_cta && row < P.nrows;
  double2* buf = reinterpret_cast<double2*>(smem_raw) + (size_t)lrow * (N + (N >> 3));
  int Ns = 1, rem = P.log2n;
  bool first = true;
  while (rem > 0) {
    const int lr = rem >= 3 ? 3 : rem;                                     // radix 8, then 4 or 2
    const bool last = rem == lr;
    if (lr == 3) stockham_pass<8, HAS_X>(P, row, active, t, T, Ns, first, last, buf);
    else if (lr == 2) stockham_pass<4, HAS_X>(P, row, active, t, T, Ns, first, last, buf);
    else stockham_pass<2, HAS_X>(P, row, active, t, T, Ns, first, last, buf);
    Ns <<= lr; rem -= lr; first = false;
  }
}


// ---------------------------------------------------------------------------------------------
// k_delay_fft_w32: one WARP per 1024-point row, two passes of radix 32 (1024 = 32 x 32).
// k_delay_fft_r8 is bound by L1/TEX requests (profiles/delay_fft_r8_r01_ncu.txt: 72 % of peak, HBM at 46 %): with 8
// points per thread and the 8.8.8.2 factorisation every point crosses the L1/TEX pipe 8 times (global load, three
// shared-memory round trips, global store) plus the weight and twiddle loads.  Here a lane holds 32 points:
//   pass 1  lane a loads x[a + 32 b] * w (b = 0..31; every b is one coalesced 512-byte run), does the 32-point
//           inverse FFT over b in registers, multiplies by W^(a c) (powers of ONE table entry, four interleaved
//           chains) and writes row a of a 32 x 33 shared-memory tile;
//   pass 2  lane c reads column c (conflict-free: 16-byte elements, row stride 33), does the 32-point inverse FFT
//           over a in registers and stores X[c + 32 d] (scale and fftshift fused; every d is a coalesced run).
// 4 L1/TEX crossings per point instead of 8, one __syncwarp instead of 5 CTA barriers, no idle lanes in a radix-2
// tail.  ~190 registers per thread: 8 warps per SM, each with 32 x 16 B loads in flight per lane (128 KB per SM).
// The 32-point transform is 4 x 8: four-point transforms over q of v[p + 8 q], twiddles W32^(p c4), eight-point
// transforms over p; output index c4 + 4 c8.
// ---------------------------------------------------------------------------------------------
__constant__ double W32C[22] = {1.00000000000000000000e+00, 9.80785280403230430579e-01, 9.23879532511286738483e-01, 8.31469612302545235671e-01, 7.07106781186547572737e-01, 5.55570233019602288671e-01, 3.82683432365089837290e-01, 1.95090322016128331351e-01, 6.12323399573676603587e-17, -1.95090322016128192573e-01, -3.82683432365089726268e-01, -5.55570233019601955604e-01, -7.07106781186547461715e-01, -8.31469612302545346694e-01, -9.23879532511286738483e-01, -9.80785280403230430579e-01, -1.00000000000000000000e+00, -9.80785280403230430579e-01, -9.23879532511286849505e-01, -8.31469612302545457716e-01, -7.07106781186547683760e-01, -5.55570233019602177649e-01};
__constant__ double W32S[22] = {0.00000000000000000000e+00, 1.95090322016128248084e-01, 3.82683432365089781779e-01, 5.55570233019602177649e-01, 7.07106781186547461715e-01, 8.31469612302545235671e-01, 9.23879532511286738483e-01, 9.80785280403230430579e-01, 1.00000000000000000000e+00, 9.80785280403230430579e-01, 9.23879532511286738483e-01, 8.31469612302545457716e-01, 7.07106781186547572737e-01, 5.55570233019602177649e-01, 3.82683432365089892802e-01, 1.95090322016128608906e-01, 1.22464679914735320717e-16, -1.95090322016128359106e-01, -3.82683432365089670757e-01, -5.55570233019601955604e-01, -7.07106781186547461715e-01, -8.31469612302545235671e-01};

__device__ __forceinline__ void ifft32(double2 (&v)[32]) {
#pragma unroll
  for (int p = 0; p < 8; ++p) {
    double2 t[4] = {v[p], v[p + 8], v[p + 16], v[p + 24]};
    ifft_small<4>(t);
    v[p] = t[0];
#pragma unroll
    for (int c4 = 1; c4 < 4; ++c4)
      v[p + 8 * c4] = p == 0 ? t[c4] : cmul(t[c4], make_double2(W32C[p * c4], W32S[p * c4]));
  }
  double2 o[32];
#pragma unroll
  for (int c4 = 0; c4 < 4; ++c4) {
    double2 t[8];
#pragma unroll
    for (int pp = 0; pp < 8; ++pp) t[pp] = v[pp + 8 * c4];
    ifft_small<8>(t);
#pragma unroll
    for (int c8 = 0; c8 < 8; ++c8) o[c4 + 4 * c8] = t[c8];
  }
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = o[i];
}

constexpr int W32_WARPS = 4;                  // rows per CTA
constexpr int W32_STRIDE = 33;                // double2 elements per shared-memory tile row

template <bool HAS_X>
__global__ void __launch_bounds__(32 * W32_WARPS, 2) k_delay_fft_w32(const DelayParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double2* tile = reinterpret_cast<double2*>(smem_raw) + (size_t)warp * 32 * W32_STRIDE;
  const int row = blockIdx.x * W32_WARPS + warp;
  if (row >= P.nrows) return;                  // warp-uniform; no CTA barrier below
  // all 32 loads of the lane in flight together (the kernel is bound by memory latency: 8 warps per SM); the window is
  // applied afterwards from the (L1-resident) weight row(s), zero padding by a zero weight
  double2 v[32];
  if (HAS_X) {
    const double2* xrow = P.x + (size_t)row * P.nchan + lane;
#pragma unroll
    for (int b = 0; b < 32; ++b) v[b] = lane + 32 * b < P.nchan ? xrow[32 * b] : make_double2(0.0, 0.0);
  }
  {
    const double* brow = P.bp + (size_t)row * P.bp_stride + lane;
    const double* wrow = P.wts ? P.wts + (size_t)row * P.wts_stride + lane : nullptr;
#pragma unroll
    for (int b = 0; b < 32; ++b) {
      double m = 0.0;
      if (lane + 32 * b < P.nchan) { m = brow[32 * b]; if (wrow) m *= wrow[32 * b]; }
      v[b] = HAS_X ? make_double2(v[b].x * m, v[b].y * m) : make_double2(m, 0.0);
    }
  }
  ifft32(v);                                   // v[c] = sum_b x[a + 32 b] W32^(b c)
  {
    // twiddles W^(a c), W = exp(2 pi i / 1024): four chains c = c0, c0 + 4, ... stepped by W^(4 a)
    const double2 w1 = __ldg(&P.twiddle[lane]);
    const double2 w2 = cmul(w1, w1);
    const double2 w4 = cmul(w2, w2);
    double2 ch[4] = {make_double2(1.0, 0.0), w1, w2, cmul(w2, w1)};
#pragma unroll
    for (int c = 0; c < 32; ++c) {
      if (c >= 4) ch[c & 3] = cmul(ch[c & 3], w4);
      tile[lane * W32_STRIDE + c] = c == 0 ? v[0] : cmul(v[c], ch[c & 3]);
    }
  }
  __syncwarp();
#pragma unroll
  for (int a = 0; a < 32; ++a) v[a] = tile[a * W32_STRIDE + lane];
  ifft32(v);                                   // v[d] = X[c + 32 d], c = lane
  double2* out = P.out + (size_t)row * P.nout;
#pragma unroll
  for (int d = 0; d < 32; ++d) {
    const int i = (lane + 32 * d - P.shift) & 1023;              // fftshift: out[i] = X[(i + shift) mod N]
    out[i] = make_double2(v[d].x * P.scale, v[d].y * P.scale);
  }
}

// direct evaluation of the needed bins; blockDim.x threads share one row staged in smem
__global__ void __launch_bounds__(FFT_THREADS) k_delay_dft(const DelayParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double2* xin = reinterpret_cast<double2*>(smem_raw);             // [nchan]
  const int row = blockIdx.x;
  const int N = P.nfft;
  for (int n = threadIdx.x; n < P.nchan; n += FFT_THREADS) xin[n] = load_in(P, row, n);
  __syncthreads();
  auto bin = [&](int kk) {                                         // X[kk] without the 1/N
    double re = 0.0, im = 0.0;
    long long idx = 0;
    for (int n = 0; n < P.nchan; ++n) {
      double2 w = P.twiddle[idx];
      double2 v = xin[n];
      re += v.x * w.x - v.y * w.y;
      im += v.x * w.y + v.y * w.x;
      idx += kk; if (idx >= N) idx -= N;
    }
    return make_double2(re, im);
  };
  for (int i = threadIdx.x; i < P.nout; i += FFT_THREADS) {
    double pos = i * P.factor;
    int j0 = (int)floor(pos);
    double fr = pos - j0;
    double2 a = bin((j0 + P.shift) % N);
    double2 o = a;
    if (fr > 0.0) {
      if (j0 + 1 < N) {
        double2 c = bin((j0 + 1 + P.shift) % N);
        o = make_double2(a.x + (c.x - a.x) * fr, a.y + (c.y - a.y) * fr);
      } else {
        o = make_double2(nan(""), nan(""));
      }
    }
    P.out[(size_t)row * P.nout + i] = make_double2(o.x * P.scale, o.y * P.scale);
  }
}

__global__ void k_mul_rows(const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ out, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] * b[i];
}

__global__ void k_fill_ones(double* p, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = 1.0;
}

__global__ void k_twiddle(double2* tw, int n) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  // exact argument reduction: 2j/n in half-turns, j < n
  double sn, cs;
  sincospi(2.0 * (double)j / (double)n, &sn, &cs);
  tw[j] = make_double2(cs, sn);
}

struct Plan { int nfft, nout; double factor; bool pow2; int log2n; bool three_pow2; };

Plan make_plan(int nchan, double pad, int downsample) {
  Plan p;
  if (pad < 0.0) pad = 0.0;                                        // interferometry.py:8090-8093
  int npad = (int)(nchan * pad);                                   // :8123
  int ntot = nchan + npad;
  double factor = (pad > 0.0 && downsample) ? 1.0 + pad : 1.0;
  p.nout = (int)ceil(ntot / factor - 1e-12);                       // len(arange(0, ntot, factor))
  p.nfft = ntot; p.factor = factor;
  double m = round(factor);
  if (fabs(factor - m) < 1e-12 && (long long)m * nchan == ntot && (nchan % 2 == 0 || m == 1.0)) {
    p.nfft = nchan; p.factor = 1.0; p.nout = nchan;                // exact shortcut, see header comment
  }
  p.pow2 = (p.nfft & (p.nfft - 1)) == 0 && p.nfft >= 2 && p.nfft <= 8192;
  p.log2n = 0;
  while ((1 << p.log2n) < p.nfft) ++p.log2n;
  // 3 * 2^k (e.g. pad = 0.5 on a power-of-two band): three interleaved 2^k-point transforms and a radix-3 combine
  p.three_pow2 = false;
  if (!p.pow2 && p.nfft % 3 == 0) {
    const int m = p.nfft / 3;
    if ((m & (m - 1)) == 0 && m >= 1 && p.nfft <= 12288) {
      p.three_pow2 = true;
      p.log2n = 0;
      while ((1 << p.log2n) < m) ++p.log2n;
    }
  }
  return p;
}

}  // namespace

extern "C" {

int pb200_delay_nout(int nchan, double pad, int downsample) {
  if (nchan <= 0) return 0;
  return make_plan(nchan, pad, downsample).nout;
}

int pb200_delay_transform(pb200_ctx* ctx, const void* d_x, const double* d_bp, long long bp_row_stride,
                          const double* d_wts, long long wts_row_stride, int nrows, int nchan, double df,
                          double pad, int downsample, void* d_out, void* stream_) {
  if (!ctx) return PB200_EINVAL;
  if (nrows <= 0 || nchan <= 0 || !d_out || df <= 0.0) return pb_fail(ctx, PB200_EINVAL, "pb200_delay_transform: bad arguments");
  if (!d_x && !d_bp && !d_wts) return pb_fail(ctx, PB200_EINVAL, "pb200_delay_transform: nothing to transform");
  cudaStream_t stream = (cudaStream_t)stream_;
  PbDeviceGuard guard(ctx->device);
  Plan pl = make_plan(nchan, pad, downsample);
  if (ctx->twiddle_n != pl.nfft) {
    PB_CUDA(ctx, cudaStreamSynchronize(stream));
    if (ctx->twiddle) cudaFree(ctx->twiddle);
    ctx->twiddle = nullptr; ctx->twiddle_n = 0;
    PB_CUDA(ctx, cudaMalloc(&ctx->twiddle, sizeof(double2) * (size_t)pl.nfft));
    k_twiddle<<<pb_div_up(pl.nfft, 256), 256, 0, stream>>>((double2*)ctx->twiddle, pl.nfft);
    PB_CHECK_LAUNCH(ctx, "k_twiddle");
    ctx->twiddle_n = pl.nfft;
  }
  DelayParams P;
  P.x = (const double2*)d_x; P.bp = d_bp; P.wts = d_wts; P.bp_stride = bp_row_stride; P.wts_stride = wts_row_stride;
  P.out = (double2*)d_out; P.twiddle = (const double2*)ctx->twiddle;
  P.nrows = nrows; P.nchan = nchan; P.nfft = pl.nfft; P.log2n = pl.log2n; P.nout = pl.nout;
  P.shift = (pl.nfft + 1) / 2;
  P.scale = df; P.factor = pl.factor;
  P.radix3 = pl.three_pow2 ? 1 : 0;
  if (pl.pow2 && pl.factor == 1.0 && pl.nfft >= 64 && pl.nfft <= 2048) {
    const int T = pl.nfft / 8;                         // one radix-8 butterfly per thread and pass
    const int rows_per_cta = 256 / T;
    const size_t smem = sizeof(double2) * (size_t)(pl.nfft + (pl.nfft >> 3)) * rows_per_cta;
    // one weight vector where possible (every load is an L1/TEX request, the unit this kernel is bound by): a missing
    // factor is dropped, two broadcast rows are multiplied once into scratch; only per-row bp AND wts keep two loads
    {
      void* tmp;
      int rc1 = pb_scratch(ctx, 6, sizeof(double) * (size_t)nchan, &tmp);
      if (rc1) return rc1;
      if (!P.bp && !P.wts) {
        k_fill_ones<<<pb_div_up(nchan, 256), 256, 0, stream>>>((double*)tmp, nchan);
        PB_CHECK_LAUNCH(ctx, "k_fill_ones");
        P.bp = (const double*)tmp; P.bp_stride = 0;
      } else if (!P.bp) {
        P.bp = P.wts; P.bp_stride = P.wts_stride; P.wts = nullptr;
      } else if (P.wts && P.bp_stride == 0 && P.wts_stride == 0) {
        k_mul_rows<<<pb_div_up(nchan, 256), 256, 0, stream>>>(P.bp, P.wts, (double*)tmp, nchan);
        PB_CHECK_LAUNCH(ctx, "k_mul_rows");
        P.bp = (const double*)tmp; P.wts = nullptr;
      }
    }
    if (pl.nfft == 1024 && !ctx->dt_force_r8) {       // one warp per row, 32 x 32 (see k_delay_fft_w32)
      const size_t smem32 = sizeof(double2) * 32 * W32_STRIDE * W32_WARPS;
      if (P.x) {
        PB_CUDA(ctx, cudaFuncSetAttribute(k_delay_fft_w32<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem32));
        k_delay_fft_w32<true><<<pb_div_up(nrows, W32_WARPS), 32 * W32_WARPS, smem32, stream>>>(P);
      } else {
        PB_CUDA(ctx, cudaFuncSetAttribute(k_delay_fft_w32<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem32));
        k_delay_fft_w32<false><<<pb_div_up(nrows, W32_WARPS), 32 * W32_WARPS, smem32, stream>>>(P);
      }
      PB_CHECK_LAUNCH(ctx, "k_delay_fft_w32");
      return PB200_OK;
    }
    if (P.x) {
      PB_CUDA(ctx, cudaFuncSetAttribute(k_delay_fft_r8<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k_delay_fft_r8<true><<<pb_div_up(nrows, rows_per_cta), 256, smem, stream>>>(P, T, rows_per_cta);
    } else {
      PB_CUDA(ctx, cudaFuncSetAttribute(k_delay_fft_r8<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k_delay_fft_r8<false><<<pb_div_up(nrows, rows_per_cta), 256, smem, stream>>>(P, T, rows_per_cta);
    }
    PB_CHECK_LAUNCH(ctx, "k_delay_fft_r8");
  } else if (pl.pow2 || pl.three_pow2) {
    size_t smem = sizeof(double2) * (size_t)pl.nfft;
    PB_CUDA(ctx, cudaFuncSetAttribute(k_delay_fft, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_delay_fft<<<nrows, FFT_THREADS, smem, stream>>>(P);
    PB_CHECK_LAUNCH(ctx, "k_delay_fft");
  } else {
    size_t smem = sizeof(double2) * (size_t)nchan;
    if (smem > 200 * 1024) return pb_fail(ctx, PB200_EUNSUPPORTED, "pb200_delay_transform: nchan too large for the generic path");
    PB_CUDA(ctx, cudaFuncSetAttribute(k_delay_dft, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_delay_dft<<<nrows, FFT_THREADS, smem, stream>>>(P);
    PB_CHECK_LAUNCH(ctx, "k_delay_dft");
  }
  return PB200_OK;
}

}  // extern "C"
