// Thermal noise + gains, fused:  rms, noise = rms/sqrt2 (N + iN), vis = gains*skyvis + noise.
// Replaces interferometry.py:6676-6693 (generate_noise) and :6707-6722 (add_noise); the scalar
// helpers thermalNoiseRMS/generateNoise (:89-329) compute the same expressions.
//
// The reference draws from numpy's global Mersenne Twister (:6693), which makes the result depend
// on call order and on how the work is split; here the deviates come from Philox4x32-10 keyed by
// (seed, global row = snapshot x baseline, channel), so any sharding of baselines over GPUs reproduces the same noise.
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
  long long row_offset;    // snapshot*nbl_total + bl_offset: global row of this shard's first baseline
  long long row_step;      // global rows between consecutive local rows (1 = contiguous block, world size = interleaved shard)
  long long nrows;
  long long st[6];         // element strides (row, col) of tsys, aeff, effq
  int nchan;
  int add_only;
};

// One standard complex normal pair (N + iN) from two 32-bit words: Box-Muller in fp32 (accurate logf / sincospif, on
// the otherwise idle FP32 pipe) -- the deviates carry 24 significant bits and reach 6.7 sigma, the thermal rms and the
// scaling stay fp64.  An fp64 Box-Muller on 53-bit uniforms (the first version of this kernel) made it FP64-issue
// bound at 3.9 TB/s; noise parity with the reference is statistical by construction (different generator).
__device__ __forceinline__ float2 normal_pair(uint32_t w0, uint32_t w1) {
  const float u1 = (float)(((double)w0 + 0.5) * (1.0 / 4294967296.0));      // (0, 1]
  const float u2 = (float)w1 * (1.0f / 4294967296.0f);                       // [0, 1]: sincospi is periodic
  const float rad = sqrtf(-2.0f * logf(u1));
  float sn, cs;
  sincospif(2.0f * u2, &sn, &cs);
  return make_float2(rad * cs, rad * sn);
}

// Thread = the two elements (row, c) and (row, c + H) of one row, H = ceil(nchan / 2): they share ONE generator call,
// Philox(counter = global_row * H + c, key = seed), words 0-1 for the first and 2-3 for the second, and every access
// of a warp is a contiguous run (lanes = consecutive c).  The counter depends on the global row only, so any sharding
// of the baselines (rows) reproduces the same numbers.
__global__ void __launch_bounds__(256) k_noise(const NoiseParams P) {
  const int H = (P.nchan + 1) >> 1;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= P.nrows * (long long)H) return;
  long long row; int c;
  if (P.nrows * (long long)H < (1ll << 31)) { const unsigned u = (unsigned)t; row = u / (unsigned)H; c = (int)(u - (unsigned)row * (unsigned)H); }
  else { row = t / H; c = (int)(t - row * H); }
  const int ne = (c + H < P.nchan) ? 2 : 1;
  if (P.add_only) {                                             // vis = gains*skyvis + noise (:6722)
    for (int e = 0; e < ne; ++e) {
      const long long i = row * P.nchan + c + e * H;
      double2 v = P.skyvis[i], nzi = P.noise[i];
      if (P.gains) {
        double2 gn = P.gains[i];
        v = make_double2(gn.x * v.x - gn.y * v.y, gn.x * v.y + gn.y * v.x);
      }
      P.vis[i] = make_double2(v.x + nzi.x, v.y + nzi.y);
    }
    return;
  }
  uint32_t r[4];
  const unsigned long long ctr = (unsigned long long)(P.row_offset + row * P.row_step) * (unsigned long long)H + (unsigned long long)c;
  philox4x32_10((uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u, (uint32_t)P.seed, (uint32_t)(P.seed >> 32), r);
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    if (e >= ne) break;
    const int col = c + e * H;
    const long long i = row * P.nchan + col;
    const double tsys = P.tsys[row * P.st[0] + col * P.st[1]];
    const double effq = P.effq[row * P.st[4] + col * P.st[5]];
    // interferometry.py:6687 / :6689
    const double rms = P.flux_unit_k ? P.scale * tsys / effq : P.scale * (tsys / P.aeff[row * P.st[2] + col * P.st[3]] / effq);
    const float2 z = e ? normal_pair(r[2], r[3]) : normal_pair(r[0], r[1]);
    const double a = rms * 0.70710678118654752440;               // rms/sqrt(2) (:6693)
    const double2 nz = make_double2(a * (double)z.x, a * (double)z.y);
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
}

}  // namespace

extern "C" int pb200_noise(pb200_ctx* ctx, const void* d_skyvis, const double* d_tsys, const double* d_aeff,
                           const double* d_effq, const long long* strides, const void* d_gains, int nbl, int nchan,
                           double df, double t_acc, int flux_unit_k, uint64_t seed, int snapshot, int bl_offset,
                           int bl_step, int nbl_total, int add_only, double* d_rms, void* d_noise, void* d_vis, void* stream_) {
  if (!ctx) return PB200_EINVAL;
  if (nbl <= 0 || nchan <= 0) return pb_fail(ctx, PB200_EINVAL, "pb200_noise: bad shape");
  if (add_only) {
    if (!d_skyvis || !d_noise || !d_vis) return pb_fail(ctx, PB200_EINVAL, "pb200_noise: add_only needs skyvis, noise and vis");
  } else if (!d_tsys || !d_effq || !strides || (!flux_unit_k && !d_aeff) || df <= 0.0 || t_acc <= 0.0) {
    return pb_fail(ctx, PB200_EINVAL, "pb200_noise: bad arguments");
  }
  if (d_vis && !d_skyvis) return pb_fail(ctx, PB200_EINVAL, "pb200_noise: d_vis requested without d_skyvis");
  if (bl_step < 1) bl_step = 1;
  if (nbl_total < bl_offset + (nbl - 1) * bl_step + 1) return pb_fail(ctx, PB200_EINVAL, "pb200_noise: shard exceeds nbl_total");
  cudaStream_t stream = (cudaStream_t)stream_;
  PbDeviceGuard guard(ctx->device);
  NoiseParams P;
  P.skyvis = (const double2*)d_skyvis; P.tsys = d_tsys; P.aeff = d_aeff; P.effq = d_effq;
  P.gains = (const double2*)d_gains; P.rms = d_rms; P.noise = (double2*)d_noise; P.vis = (double2*)d_vis;
  P.scale = add_only ? 0.0 : (flux_unit_k ? 1.0 / sqrt(t_acc * df) : 2.0 * PB_BOLTZMANN / sqrt(t_acc * df) / PB_JY);
  for (int i = 0; i < 6; ++i) P.st[i] = strides ? strides[i] : 0;
  P.nchan = nchan; P.add_only = add_only;
  P.flux_unit_k = flux_unit_k;
  P.seed = seed;
  P.row_offset = (long long)snapshot * nbl_total + bl_offset;
  P.row_step = bl_step;
  P.nrows = nbl;
  k_noise<<<pb_div_up((long long)nbl * ((nchan + 1) / 2), 256), 256, 0, stream>>>(P);
  PB_CHECK_LAUNCH(ctx, "k_noise");
  return PB200_OK;
}
