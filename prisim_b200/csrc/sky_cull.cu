// Sky coordinates -> ENU direction cosines, horizon / ROI cull and order-preserving compaction.
// Replaces interferometry.py:6174-6219 of the reference (hadec2altaz :6177, alt >= 90 - roi_radius
// :6216, skypos_altaz[m2] :6219); the spherical trigonometry restates astroutils
// GEOM.hadec2altaz / altaz2dircos (un-vendored dependency).
//
// Three small fp64 kernels (O(nsrc0) work, HBM-bound, negligible next to the phase sum):
//   k_sky_flags   : dircos + keep flag per source, per-block keep counts
//   k_scan_counts : exclusive scan of the block counts (single block)
//   k_compact     : intra-block ballot scan + scatter, preserving catalogue order so that
//                   d_index equals the reference's ascending m2 (obs_catalog_indices :6377)
#include "common.cuh"

namespace {

constexpr int CULL_THREADS = 256;
constexpr double DEG2RAD = 0.017453292519943295769;

__device__ __forceinline__ void sky_to_dircos(const double* __restrict__ skypos, int s, int coords,
                                              double sinlat, double coslat, double& l, double& m,
                                              double& n, double& alt_deg) {
  if (coords == PB200_SKY_DIRCOS) {
    l = skypos[3 * s]; m = skypos[3 * s + 1]; n = skypos[3 * s + 2];
    alt_deg = asin(fmin(1.0, fmax(-1.0, n))) / DEG2RAD;
  } else if (coords == PB200_SKY_ALTAZ) {
    double alt = skypos[2 * s], az = skypos[2 * s + 1];
    double sa, ca, sz, cz;
    sincos(alt * DEG2RAD, &sa, &ca);
    sincos(az * DEG2RAD, &sz, &cz);
    l = ca * sz; m = ca * cz; n = sa;          // East, North, Up; az from North through East
    alt_deg = alt;
  } else {                                     // HA (west positive), Dec at the given latitude
    double sh, ch, sd, cd;
    sincos(skypos[2 * s] * DEG2RAD, &sh, &ch);
    sincos(skypos[2 * s + 1] * DEG2RAD, &sd, &cd);
    l = -cd * sh;
    m = sd * coslat - cd * ch * sinlat;
    n = sd * sinlat + cd * ch * coslat;
    alt_deg = asin(fmin(1.0, fmax(-1.0, n))) / DEG2RAD;
  }
}

__global__ void __launch_bounds__(CULL_THREADS)
k_sky_flags(const double* __restrict__ skypos, int nsrc0, int coords, double sinlat, double coslat,
            double alt_min_deg, int use_center, double cx, double cy, double cz, double cos_radius,
            double* __restrict__ tmp_dircos, unsigned char* __restrict__ flags,
            int* __restrict__ block_counts) {
  int s = blockIdx.x * CULL_THREADS + threadIdx.x;
  int keep = 0;
  if (s < nsrc0) {
    double l, m, n, alt;
    sky_to_dircos(skypos, s, coords, sinlat, coslat, l, m, n, alt);
    tmp_dircos[3 * (size_t)s] = l; tmp_dircos[3 * (size_t)s + 1] = m; tmp_dircos[3 * (size_t)s + 2] = n;
    if (use_center) keep = (l * cx + m * cy + n * cz >= cos_radius) ? 1 : 0;   // spherematch about the pointing centre, :6213
    else keep = (alt >= alt_min_deg) ? 1 : 0;                                  // interferometry.py:6216
    flags[s] = (unsigned char)keep;
  }
  int total = __syncthreads_count(keep);
  if (threadIdx.x == 0) block_counts[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024)
k_scan_counts(int* __restrict__ block_counts, int nblocks, int* __restrict__ total_out) {
  __shared__ int warp_sums[32];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nblocks; base += 1024) {
    int i = base + threadIdx.x;
    int v = (i < nblocks) ? block_counts[i] : 0;
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += t;
    }
    if (lane == 31) warp_sums[w] = incl;
    __syncthreads();
    if (w == 0) {
      int ws = warp_sums[lane];
      int wi = ws;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, wi, d);
        if (lane >= d) wi += t;
      }
      warp_sums[lane] = wi - ws;               // exclusive warp offsets
    }
    __syncthreads();
    int excl = carry + warp_sums[w] + incl - v;
    if (i < nblocks) block_counts[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total_out = carry;
}

__global__ void __launch_bounds__(CULL_THREADS)
k_compact(const double* __restrict__ tmp_dircos, const unsigned char* __restrict__ flags, int nsrc0,
          const int* __restrict__ block_offsets, double* __restrict__ dircos, int32_t* __restrict__ index) {
  __shared__ int warp_off[CULL_THREADS / 32];
  int s = blockIdx.x * CULL_THREADS + threadIdx.x;
  int keep = (s < nsrc0) ? flags[s] : 0;
  unsigned ballot = __ballot_sync(0xffffffffu, keep);
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) warp_off[w] = __popc(ballot);
  __syncthreads();
  int off = block_offsets[blockIdx.x];
  for (int i = 0; i < w; ++i) off += warp_off[i];
  off += __popc(ballot & ((1u << lane) - 1u));
  if (keep) {
    dircos[3 * (size_t)off] = tmp_dircos[3 * (size_t)s];
    dircos[3 * (size_t)off + 1] = tmp_dircos[3 * (size_t)s + 1];
    dircos[3 * (size_t)off + 2] = tmp_dircos[3 * (size_t)s + 2];
    index[off] = s;
  }
}

}  // namespace

extern "C" int pb200_sky_cull(pb200_ctx* ctx, const double* d_skypos, int nsrc0, int coords,
                              double latitude_deg, double roi_radius_deg, const double* h_roi_center,
                              double* d_dircos, int32_t* d_index, int* h_nsrc, void* stream_) {
  if (!ctx) return PB200_EINVAL;
  if (!h_nsrc || nsrc0 < 0) return pb_fail(ctx, PB200_EINVAL, "pb200_sky_cull: bad arguments");
  if (coords < PB200_SKY_ALTAZ || coords > PB200_SKY_DIRCOS)
    return pb_fail(ctx, PB200_EINVAL, "pb200_sky_cull: coords must be PB200_SKY_ALTAZ/HADEC/DIRCOS");
  *h_nsrc = 0;
  if (nsrc0 == 0) return PB200_OK;
  if (!d_skypos || !d_dircos || !d_index) return pb_fail(ctx, PB200_EINVAL, "pb200_sky_cull: null device pointer");
  cudaStream_t stream = (cudaStream_t)stream_;
  PbDeviceGuard guard(ctx->device);
  int nblocks = pb_div_up(nsrc0, CULL_THREADS);
  size_t b_dircos = sizeof(double) * 3 * (size_t)nsrc0;
  size_t b_flags = ((size_t)nsrc0 + 255) & ~(size_t)255;
  size_t b_counts = sizeof(int) * ((size_t)nblocks + 1);
  void* scratch;
  int rc = pb_scratch(ctx, 0, b_dircos + b_flags + b_counts + 256, &scratch);
  if (rc) return rc;
  double* tmp_dircos = (double*)scratch;
  unsigned char* flags = (unsigned char*)scratch + b_dircos;
  int* counts = (int*)((unsigned char*)scratch + b_dircos + b_flags);
  int* total = counts + nblocks;
  double lat = latitude_deg * DEG2RAD;
  double cx = 0, cy = 0, cz = 1;
  if (h_roi_center) { cx = h_roi_center[0]; cy = h_roi_center[1]; cz = h_roi_center[2]; }
  k_sky_flags<<<nblocks, CULL_THREADS, 0, stream>>>(d_skypos, nsrc0, coords, sin(lat), cos(lat),
                                                    90.0 - roi_radius_deg, h_roi_center != nullptr, cx, cy, cz,
                                                    cos(roi_radius_deg * DEG2RAD), tmp_dircos, flags, counts);
  PB_CHECK_LAUNCH(ctx, "k_sky_flags");
  k_scan_counts<<<1, 1024, 0, stream>>>(counts, nblocks, total);
  PB_CHECK_LAUNCH(ctx, "k_scan_counts");
  k_compact<<<nblocks, CULL_THREADS, 0, stream>>>(tmp_dircos, flags, nsrc0, counts, d_dircos, d_index);
  PB_CHECK_LAUNCH(ctx, "k_compact");
  PB_CUDA(ctx, cudaMemcpyAsync(h_nsrc, total, sizeof(int), cudaMemcpyDeviceToHost, stream));
  PB_CUDA(ctx, cudaStreamSynchronize(stream));
  return PB200_OK;
}
