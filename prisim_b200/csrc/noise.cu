// Thermal noise + gains, fused:  rms, noise = rms/sqrt2 (N + iN), vis = gains*skyvis + noise.
// Replaces interferometry.py:6676-6693 (generate_noise) and :6707-6722 (add_noise); the scalar
// helpers thermalNoiseRMS/generateNoise (:89-329) compute the same expressions.
//
// The reference draws from numpy's global Mersenne Twister (:6693), which makes the result depend
// on call order and on how the work is split; here the deviates come from Philox4x32-10 keyed by
// (seed, global element index), so any sharding of baselines over GPUs reproduces the same noise.
// HBM-bound elementwise kernel: algorithmic bytes per element = 16 (skyvis) + 8 (Tsys) + 8 (A_eff)
// + 8 (eff_Q) read, 8 (rms) + 16 (noise) + 16 (vis) written = 80 B.
#include "common.cuh"

namespace {

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t out[4]) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

struct NoiseParams {
  const double2* skyvis;
  const double* tsys;
  const double* aeff;
  const double* effq;
  const double2* gains;
  double* rms;
  double2* noise;
  double2* vis;
  double scale;            // 2k/sqrt(df t)/Jy  or 1/sqrt(df t)
  int flux_unit_k;
  unsigned long long seed;
  long long elem_offset;   // (snapshot*nbl_total + bl_offset)*nchan
  long long n;
  long long st[6];         // element strides (row, col) of tsys, aeff, effq
  int nchan;
  int add_only;
};

__global__ void __launch_bounds__(256) k_noise(const NoiseParams P) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.n) return;
  if (P.add_only) {                                             // vis = gains*skyvis + noise (:6722)
    double2 v = P.skyvis[i], nzi = P.noise[i];
    if (P.gains) {
      double2 gn = P.gains[i];
      v = make_double2(gn.x * v.x - gn.y * v.y, gn.x * v.y + gn.y * v.x);
    }
    P.vis[i] = make_double2(v.x + nzi.x, v.y + nzi.y);
    return;
  }
  const long long row = i / P.nchan, col = i - row * P.nchan;
  const double tsys = P.tsys[row * P.st[0] + col * P.st[1]];
  const double effq = P.effq[row * P.st[4] + col * P.st[5]];
  // interferometry.py:6687 / :6689
  double rms = P.flux_unit_k ? P.scale * tsys / effq : P.scale * (tsys / P.aeff[row * P.st[2] + col * P.st[3]] / effq);
  unsigned long long g = (unsigned long long)(P.elem_offset + i);
  uint32_t r[4];
  philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), 0u, 0u, (uint32_t)P.seed, (uint32_t)(P.seed >> 32), r);
  // two uniforms in (0,1) with 53 / 32 significant bits, Box-Muller -> two independent N(0,1)
  double u1 = ((double)(((unsigned long long)r[0] << 21) ^ (unsigned long long)(r[1] >> 11)) + 0.5) * (1.0 / 9007199254740992.0);
  double u2 = ((double)(((unsigned long long)r[2] << 21) ^ (unsigned long long)(r[3] >> 11)) + 0.5) * (1.0 / 9007199254740992.0);
  double rad = sqrt(-2.0 * log(u1));
  double sn, cs;
  sincospi(2.0 * u2, &sn, &cs);
  const double a = rms * 0.70710678118654752440;               // rms/sqrt(2) (:6693)
  double2 nz = make_double2(a * rad * cs, a * rad * sn);
  if (P.rms) P.rms[i] = rms;
  if (P.noise) P.noise[i] = nz;
  if (P.vis) {
    double2 v = P.skyvis[i];
    if (P.gains) {                                              // gains * skyvis (:6722)
      double2 gn = P.gains[i];
      v = make_double2(gn.x * v.x - gn.y * v.y, gn.x * v.y + gn.y * v.x);
    }
    P.vis[i] = make_double2(v.x + nz.x, v.y + nz.y);
  }
}

}  // namespace

extern "C" int pb200_noise(pb200_ctx* ctx, const void* d_skyvis, const double* d_tsys, const double* d_aeff,
                           const double* d_effq, const long long* strides, const void* d_gains, int nbl, int nchan,
                           double df, double t_acc, int flux_unit_k, uint64_t seed, int snapshot, int bl_offset,
                           int nbl_total, int add_only, double* d_rms, void* d_noise, void* d_vis, void* stream_) {
  if (!ctx) return PB200_EINVAL;
  if (nbl <= 0 || nchan <= 0) return pb_fail(ctx, PB200_EINVAL, "pb200_noise: bad shape");
  if (add_only) {
    if (!d_skyvis || !d_noise || !d_vis) return pb_fail(ctx, PB200_EINVAL, "pb200_noise: add_only needs skyvis, noise and vis");
  } else if (!d_tsys || !d_effq || !strides || (!flux_unit_k && !d_aeff) || df <= 0.0 || t_acc <= 0.0) {
    return pb_fail(ctx, PB200_EINVAL, "pb200_noise: bad arguments");
  }
  if (d_vis && !d_skyvis) return pb_fail(ctx, PB200_EINVAL, "pb200_noise: d_vis requested without d_skyvis");
  if (nbl_total < bl_offset + nbl) return pb_fail(ctx, PB200_EINVAL, "pb200_noise: shard exceeds nbl_total");
  cudaStream_t stream = (cudaStream_t)stream_;
  PB_CUDA(ctx, cudaSetDevice(ctx->device));
  NoiseParams P;
  P.skyvis = (const double2*)d_skyvis; P.tsys = d_tsys; P.aeff = d_aeff; P.effq = d_effq;
  P.gains = (const double2*)d_gains; P.rms = d_rms; P.noise = (double2*)d_noise; P.vis = (double2*)d_vis;
  P.scale = add_only ? 0.0 : (flux_unit_k ? 1.0 / sqrt(t_acc * df) : 2.0 * PB_BOLTZMANN / sqrt(t_acc * df) / PB_JY);
  for (int i = 0; i < 6; ++i) P.st[i] = strides ? strides[i] : 0;
  P.nchan = nchan; P.add_only = add_only;
  P.flux_unit_k = flux_unit_k;
  P.seed = seed;
  P.elem_offset = ((long long)snapshot * nbl_total + bl_offset) * (long long)nchan;
  P.n = (long long)nbl * nchan;
  k_noise<<<pb_div_up(P.n, 256), 256, 0, stream>>>(P);
  PB_CHECK_LAUNCH(ctx, "k_noise");
  return PB200_OK;
}
