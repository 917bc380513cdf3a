// Gridded (HEALPix) primary beam: bilinear interpolation on the sphere of a log10 power beam map at
// the culled source directions, per channel, plus the per-channel maximum used to renormalise it.
// Replaces the external-beam step of the reference, scripts/run_prisim.py:1897-1908:
//     theta_phi = (pi/2 - alt, az)
//     interp_logbeam = OPS.healpix_interp_along_axis(log10(external_beam), theta_phi, beam_freqs -> chans)
//     interp_logbeam -= max(0, nanmax(interp_logbeam, axis=0));   pbeam = 10**interp_logbeam
// healpix_interp_along_axis (un-vendored astroutils) = healpy.get_interp_val (bilinear in the RING scheme: two
// pixels in each of the two rings bracketing the colatitude) followed by a 1-D interpolation in frequency.
// Both steps are linear in the map, so they commute: the host resamples the map to the observing channels once
// (prisim_b200.primary_beams.HealpixBeam) and the device does the per-snapshot gather.
//
// Layout: map [npix][nchan] (channel fastest), so the four pixels of a source are four contiguous rows and a
// CTA's loads are fully coalesced.  HBM-bound gather: algorithmic bytes per (source, channel) = 4 pixels x
// sizeof(map element) read + 8 B written (+ 8 B re-read by the column maximum).
#include "common.cuh"

namespace {

constexpr double PI = 3.14159265358979323846;

struct RingInfo { long long startpix; int ringpix; double theta; bool shifted; };

__device__ __forceinline__ RingInfo ring_info(int ring, int nside) {
  const long long npix = 12LL * nside * nside, ncap = 2LL * nside * (nside - 1);
  const double fact2 = 4.0 / (double)npix, fact1 = 2.0 * nside * fact2;
  const int northring = ring > 2 * nside ? 4 * nside - ring : ring;
  RingInfo R;
  if (northring < nside) {
    const double tmp = (double)northring * northring * fact2;
    R.theta = atan2(sqrt(tmp * (2.0 - tmp)), 1.0 - tmp);
    R.ringpix = 4 * northring;
    R.shifted = true;
    R.startpix = 2LL * northring * (northring - 1);
  } else {
    R.theta = acos((2.0 * nside - northring) * fact1);
    R.ringpix = 4 * nside;
    R.shifted = ((northring - nside) & 1) == 0;
    R.startpix = ncap + (long long)(northring - nside) * R.ringpix;
  }
  if (northring != ring) {          // southern hemisphere: mirror
    R.theta = PI - R.theta;
    R.startpix = npix - R.startpix - R.ringpix;
  }
  return R;
}

__device__ __forceinline__ int ring_above(double z, int nside) {
  const double az = fabs(z);
  if (az <= 2.0 / 3.0) return (int)(nside * (2.0 - 1.5 * z));
  const int iring = (int)(nside * sqrt(3.0 * (1.0 - az)));
  return z > 0 ? iring : 4 * nside - iring - 1;
}

// the four pixels and weights of healpy.get_interp_weights (RING scheme)
__device__ void interp_weights(double theta, double phi, int nside, long long pix[4], double wgt[4]) {
  const long long npix = 12LL * nside * nside;
  const double z = cos(theta);
  const int ir1 = ring_above(z, nside), ir2 = ir1 + 1;
  double theta1 = 0.0, theta2 = 0.0;
  if (ir1 > 0) {
    const RingInfo R = ring_info(ir1, nside);
    theta1 = R.theta;
    const double dphi = 2.0 * PI / R.ringpix, sh = R.shifted ? 0.5 : 0.0;
    const double tmp = phi / dphi - sh;
    int i1 = tmp < 0 ? (int)tmp - 1 : (int)tmp;
    const double w1 = (phi - (i1 + sh) * dphi) / dphi;
    int i2 = i1 + 1;
    if (i1 < 0) i1 += R.ringpix;
    if (i2 >= R.ringpix) i2 -= R.ringpix;
    pix[0] = R.startpix + i1; pix[1] = R.startpix + i2;
    wgt[0] = 1.0 - w1; wgt[1] = w1;
  }
  if (ir2 < 4 * nside) {
    const RingInfo R = ring_info(ir2, nside);
    theta2 = R.theta;
    const double dphi = 2.0 * PI / R.ringpix, sh = R.shifted ? 0.5 : 0.0;
    const double tmp = phi / dphi - sh;
    int i1 = tmp < 0 ? (int)tmp - 1 : (int)tmp;
    const double w1 = (phi - (i1 + sh) * dphi) / dphi;
    int i2 = i1 + 1;
    if (i1 < 0) i1 += R.ringpix;
    if (i2 >= R.ringpix) i2 -= R.ringpix;
    pix[2] = R.startpix + i1; pix[3] = R.startpix + i2;
    wgt[2] = 1.0 - w1; wgt[3] = w1;
  }
  if (ir1 == 0) {                   // north of the first ring: blend with the four polar pixels
    const double wtheta = theta / theta2;
    wgt[2] *= wtheta; wgt[3] *= wtheta;
    const double fac = (1.0 - wtheta) * 0.25;
    wgt[0] = fac; wgt[1] = fac; wgt[2] += fac; wgt[3] += fac;
    pix[0] = (pix[2] + 2) & 3; pix[1] = (pix[3] + 2) & 3;
  } else if (ir2 == 4 * nside) {    // south of the last ring
    const double wtheta = (theta - theta1) / (PI - theta1);
    wgt[0] *= (1.0 - wtheta); wgt[1] *= (1.0 - wtheta);
    const double fac = wtheta * 0.25;
    wgt[0] += fac; wgt[1] += fac; wgt[2] = fac; wgt[3] = fac;
    pix[2] = ((pix[0] + 2) & 3) + npix - 4; pix[3] = ((pix[1] + 2) & 3) + npix - 4;
  } else {
    const double wtheta = (theta - theta1) / (theta2 - theta1);
    wgt[0] *= (1.0 - wtheta); wgt[1] *= (1.0 - wtheta);
    wgt[2] *= wtheta; wgt[3] *= wtheta;
  }
}

// one warp per source: lane 0 derives the 4 pixels + weights, the warp then streams the 4 map rows
template <typename MAP>
__global__ void __launch_bounds__(256) k_healpix_gather(const MAP* __restrict__ map, int nside, const double* __restrict__ dircos,
                                                        int nsrc, int nchan, double* __restrict__ logbeam) {
  const int lane = threadIdx.x & 31;
  const int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (s >= nsrc) return;
  long long pix[4] = {0, 0, 0, 0}; double wgt[4] = {0, 0, 0, 0};
  if (lane == 0) {
    const double l = dircos[3 * (size_t)s], m = dircos[3 * (size_t)s + 1], n = dircos[3 * (size_t)s + 2];
    const double theta = acos(fmin(1.0, fmax(-1.0, n)));             // pi/2 - alt
    double phi = atan2(l, m);                                        // az, North through East
    if (phi < 0) phi += 2.0 * PI;
    interp_weights(theta, phi, nside, pix, wgt);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) { pix[i] = __shfl_sync(0xffffffffu, pix[i], 0); wgt[i] = __shfl_sync(0xffffffffu, wgt[i], 0); }
  const MAP* r0 = map + (size_t)pix[0] * nchan; const MAP* r1 = map + (size_t)pix[1] * nchan;
  const MAP* r2 = map + (size_t)pix[2] * nchan; const MAP* r3 = map + (size_t)pix[3] * nchan;
  double* out = logbeam + (size_t)s * nchan;
  for (int f = lane; f < nchan; f += 32)
    out[f] = wgt[0] * (double)r0[f] + wgt[1] * (double)r1[f] + wgt[2] * (double)r2[f] + wgt[3] * (double)r3[f];
}

// colmax[f] = max(0, max_s logbeam[s,f])   (run_prisim.py:1904-1905); colmax zeroed by the caller, non-negative doubles
// order like their bit patterns so the cross-CTA reduction is an integer atomicMax
__global__ void __launch_bounds__(256) k_colmax(const double* __restrict__ logbeam, int nsrc, int nchan, double* __restrict__ colmax) {
  __shared__ double red[8][33];
  const int f = blockIdx.x * 32 + threadIdx.x;                       // blockDim = (32, 8)
  double m = 0.0;
  if (f < nchan)
    for (int s = blockIdx.y * 8 + threadIdx.y; s < nsrc; s += 8 * gridDim.y) m = fmax(m, logbeam[(size_t)s * nchan + f]);   // fmax drops NaNs (nanmax)
  red[threadIdx.y][threadIdx.x] = m;
  __syncthreads();
  if (threadIdx.y == 0 && f < nchan) {
    for (int i = 1; i < 8; ++i) m = fmax(m, red[i][threadIdx.x]);
    atomicMax(reinterpret_cast<unsigned long long*>(colmax) + f, (unsigned long long)__double_as_longlong(m));
  }
}

}  // namespace

extern "C" int pb200_healpix_beam(pb200_ctx* ctx, const void* d_map, int map_dtype, int nside, const double* d_dircos,
                                  int nsrc, int nchan, double* d_logbeam, double* d_colmax, void* stream_) {
  if (!ctx) return PB200_EINVAL;
  if (!d_map || nside <= 0 || (nside & (nside - 1)) || nsrc < 0 || nchan <= 0 || !d_logbeam || !d_colmax)
    return pb_fail(ctx, PB200_EINVAL, "pb200_healpix_beam: bad arguments (nside must be a power of two)");
  if (map_dtype != PB200_AMP_F32 && map_dtype != PB200_AMP_F64)
    return pb_fail(ctx, PB200_EINVAL, "pb200_healpix_beam: map_dtype must be PB200_AMP_F32 or PB200_AMP_F64");
  cudaStream_t stream = (cudaStream_t)stream_;
  PbDeviceGuard guard(ctx->device);
  if (nsrc > 0) {
    if (!d_dircos) return pb_fail(ctx, PB200_EINVAL, "pb200_healpix_beam: null d_dircos");
    if (map_dtype == PB200_AMP_F64) k_healpix_gather<double><<<pb_div_up(nsrc, 8), 256, 0, stream>>>((const double*)d_map, nside, d_dircos, nsrc, nchan, d_logbeam);
    else k_healpix_gather<float><<<pb_div_up(nsrc, 8), 256, 0, stream>>>((const float*)d_map, nside, d_dircos, nsrc, nchan, d_logbeam);
    PB_CHECK_LAUNCH(ctx, "k_healpix_gather");
  }
  PB_CUDA(ctx, cudaMemsetAsync(d_colmax, 0, sizeof(double) * (size_t)nchan, stream));
  const int ysplit = nsrc > 0 ? (pb_div_up(nsrc, 8) < 4 * ctx->sm_count ? pb_div_up(nsrc, 8) : 4 * ctx->sm_count) : 1;
  k_colmax<<<dim3(pb_div_up(nchan, 32), ysplit), dim3(32, 8), 0, stream>>>(d_logbeam, nsrc, nchan, d_colmax);
  PB_CHECK_LAUNCH(ctx, "k_colmax");
  return PB200_OK;
}
