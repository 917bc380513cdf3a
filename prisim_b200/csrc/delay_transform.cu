// Windowed delay transform along frequency, batched over rows (baselines of one snapshot).
// Replaces interferometry.py:8114-8134 and delay_spectrum.py:1305-1327 of the reference:
//   X = fftshift(ifft(pad_right(x * bp * wts, npad))) * (nchan + npad) * df,
//   then DSP.downsampler(X, 1+pad) = linear interpolation at positions arange(0, n, 1+pad).
//
// Two hand-written kernels (no cuFFT):
//   k_delay_fft   radix-2 decimation-in-time inverse FFT of one row in shared memory (fp64),
//                 fused with the bp*wts multiply on load and with scale + fftshift + decimation on
//                 store.  Used when the transform length is a power of two <= 8192.  When 1+pad
//                 is an integer m and nchan is even, decimating the m*nchan-point padded
//                 transform by m is exactly the nchan-point transform, so that one is computed.
//   k_delay_dft   direct evaluation of just the output samples needed (any length, any pad):
//                 O(nout * nchan) per row, twiddles from an exact integer-indexed table.
// HBM-bound: algorithmic bytes per row = nchan*(16 + 8 + 8) read + nout*16 written.
#include "common.cuh"
#include <cmath>

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
  // bit-reversed load
  for (int n = threadIdx.x; n < N; n += FFT_THREADS) {
    int r = (int)(__brev((unsigned)n) >> (32 - P.log2n));
    buf[r] = load_in(P, row, n);
  }
  __syncthreads();
  for (int st = 1; st <= P.log2n; ++st) {
    const int half = 1 << (st - 1);
    const int tstride = N >> st;
    for (int i = threadIdx.x; i < N / 2; i += FFT_THREADS) {
      int j = i & (half - 1);
      int base = ((i - j) << 1) + j;
      double2 w = P.twiddle[j * tstride];
      double2 a = buf[base], bb = cmul(buf[base + half], w);
      buf[base] = make_double2(a.x + bb.x, a.y + bb.y);
      buf[base + half] = make_double2(a.x - bb.x, a.y - bb.y);
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

__global__ void k_twiddle(double2* tw, int n) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  // exact argument reduction: 2j/n in half-turns, j < n
  double sn, cs;
  sincospi(2.0 * (double)j / (double)n, &sn, &cs);
  tw[j] = make_double2(cs, sn);
}

struct Plan { int nfft, nout; double factor; bool pow2; int log2n; };

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
  PB_CUDA(ctx, cudaSetDevice(ctx->device));
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
  if (pl.pow2) {
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
