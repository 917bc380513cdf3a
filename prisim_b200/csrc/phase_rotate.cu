// Re-phasing of visibilities to a new phase centre:  V[b,f] *= exp(-2 pi i f b.(s_old - s_new)/c).
// Replaces the elementwise product at interferometry.py:7871-7881 of the reference
// (InterferometerArray.phase_centering; called through rotate_visibilities at
// scripts/run_prisim.py:2282).  HBM-bound: 32 B moved per element; the phase is formed and range-
// reduced in fp64 (b.ds f/c reaches hundreds of turns) and evaluated with the fp64 sincospi.
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) k_phase_rotate(double2* __restrict__ vis, const double* __restrict__ bl,
                                                      const double* __restrict__ freqs, double dx, double dy, double dz,
                                                      int nbl, int nchan) {
  for (int b = blockIdx.y; b < nbl; b += gridDim.y) {      // grid-stride: gridDim.y is capped at 65535
    const double tau = (bl[3 * (size_t)b] * dx + bl[3 * (size_t)b + 1] * dy + bl[3 * (size_t)b + 2] * dz) / PB_SPEED_OF_LIGHT;
    for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < nchan; f += gridDim.x * blockDim.x) {
      double u = tau * freqs[f];
      u -= rint(u);
      double sn, cs;
      sincospi(2.0 * u, &sn, &cs);
      double2 v = vis[(size_t)b * nchan + f];
      vis[(size_t)b * nchan + f] = make_double2(v.x * cs + v.y * sn, v.y * cs - v.x * sn);   // v * (cs - i sn)
    }
  }
}

}  // namespace

extern "C" int pb200_phase_rotate(pb200_ctx* ctx, void* d_vis, const double* d_bl, int nbl, const double* h_dpos,
                                  const double* h_freqs, int nchan, void* stream_) {
  if (!ctx) return PB200_EINVAL;
  if (!d_vis || !d_bl || !h_dpos || !h_freqs || nbl <= 0 || nchan <= 0)
    return pb_fail(ctx, PB200_EINVAL, "pb200_phase_rotate: bad arguments");
  cudaStream_t stream = (cudaStream_t)stream_;
  PbDeviceGuard guard(ctx->device);
  const double* dfreq;
  int rc = pb_channels_device(ctx, h_freqs, nchan, nchan, stream, &dfreq);
  if (rc) return rc;
  dim3 grid(pb_div_up(nchan, 256), nbl < 65535 ? nbl : 65535);
  k_phase_rotate<<<grid, 256, 0, stream>>>((double2*)d_vis, d_bl, dfreq, h_dpos[0], h_dpos[1], h_dpos[2], nbl, nchan);
  PB_CHECK_LAUNCH(ctx, "k_phase_rotate");
  return PB200_OK;
}
