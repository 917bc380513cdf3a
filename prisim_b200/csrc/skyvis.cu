// The phase sum  V[b,f] = sum_s amp[s,f] w[s,b,f] exp(-2 pi i f (s_s - s_pc).b / c)  -- the hot kernel.
// Replaces interferometry.py:6155-6165, :6255 (+ baseline_delay_horizon.py:240), :6258-6283 and
// :6332-6340 / :6348-6376 of the reference, which materialise [nsrc,nbl,nchan] complex128 slabs
// and call numpy exp/sum on them.
//
// Design (see DESIGN.md section "K1"):
//   * FP32-FMA-issue bound, not HBM and not tensor cores: 6 FMA-pipe issues per term (complex
//     rotation 2 FMUL + 2 FFMA, accumulate 2 FFMA) is the algorithmic cost.
//   * one thread owns one baseline x KT consecutive channels, accumulators in registers; a warp's
//     32 lanes are 32 baselines of the same channel block, so amplitude reads are shared-memory
//     broadcasts (LDS.128 = 4 channels for the whole warp).
//   * a CTA = WC channel blocks (one PB200_SLAB-channel slab) x WB baseline groups; it streams
//     source tiles (PB200_SRC_TILE rows of the slab + fp64 geometry) through a double-buffered
//     shared-memory ring filled by TMA bulk copies (cp.async.bulk + mbarrier).
//   * phases: tau = s.b/c - tau_pc in fp64; the anchor phase tau*f_k0 is range-reduced in fp64
//     and only the fraction goes to fp32; channels advance by an fp32 complex rotation r =
//     exp(-2 pi i tau df), re-anchored every KA channels (survey section 8d: K <= 32-64 holds 1e-5).
//   * fp32 accumulators are flushed into the fp64 output every FLUSH_TILES source tiles
//     (<= 1024 sources), the output buffer itself being the fp64 accumulator (one owner thread per
//     (b,f), so no atomics and a deterministic sum order).
#include "common.cuh"

namespace {

constexpr int KT = 32;                         // channels per thread
constexpr int WC = PB200_SLAB / KT;            // channel blocks (warps) per slab = 4
constexpr int WB = 4;                          // baseline groups (warps) per CTA
constexpr int BL_PER_CTA = 32 * WB;            // 128
constexpr int NTHREADS = 32 * WC * WB;         // 512
constexpr int T = PB200_SRC_TILE;              // sources per tile
constexpr int FLUSH_TILES = 32;                // fp32 -> fp64 flush cadence (1024 sources)
constexpr int NSTAGE = 2;

struct SkyvisParams {
  const float* amp;        // [nslab][nsrc_pad][SLAB]
  const double* geom;      // [nsrc_pad][4]: l, m, n, taper coefficient q
  const double* bl;        // [nbl][3] metres
  const double* freqs;     // device [nchan_pad] Hz (padded channels repeat the last frequency)
  double* vis;             // [nbl][nchan] complex128
  double pc[3];            // phase-centre dircos
  double f0, df;           // uniform channels: f_k = f0 + k df
  int nsrc_pad, nbl, nchan, nslab;
};

struct __align__(16) Tile {
  float amp[T][PB200_SLAB];    // 16 KB
  double geom[T][4];           // 1 KB
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// TMA bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// fraction of x in [-0.5, 0.5] (round-to-nearest-even magic number; |x| < 2^51)
__device__ __forceinline__ double frac_turns(double x) {
  const double M = 6755399441055744.0;   // 1.5 * 2^52
  double r = __dadd_rn(__dadd_rn(x, M), -M);
  return x - r;
}

template <bool TAPER, bool DIRECT>
__global__ void __launch_bounds__(NTHREADS, 1) k_skyvis(const SkyvisParams P) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Tile* tiles = reinterpret_cast<Tile*>(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + NSTAGE * sizeof(Tile));
  float* sfreq = reinterpret_cast<float*>(smem_raw + NSTAGE * sizeof(Tile) + 64);   // [SLAB] taper: (f/1e8)^2 ; direct: unused
  double* sfreq64 = reinterpret_cast<double*>(smem_raw + NSTAGE * sizeof(Tile) + 64 + PB200_SLAB * sizeof(float));   // [SLAB]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wc = warp % WC, wb = warp / WC;
  const int slab = blockIdx.x;
  const int b = blockIdx.y * BL_PER_CTA + wb * 32 + lane;
  const bool valid = b < P.nbl;
  const int kbase = slab * PB200_SLAB + wc * KT;      // first global channel of this thread
  const int ntiles = P.nsrc_pad / T;

  // baseline in light-seconds (fp64), phase-centre delay
  double bx = 0, by = 0, bz = 0;
  if (valid) {
    bx = P.bl[3 * (size_t)b] / PB_SPEED_OF_LIGHT;
    by = P.bl[3 * (size_t)b + 1] / PB_SPEED_OF_LIGHT;
    bz = P.bl[3 * (size_t)b + 2] / PB_SPEED_OF_LIGHT;
  }
  const double tau_pc = P.pc[0] * bx + P.pc[1] * by + P.pc[2] * bz;     // interferometry.py:6165
  const double blen2 = bx * bx + by * by + bz * bz;                      // (|b|/c)^2, taper
  const double fk0 = P.f0 + (double)kbase * P.df;

  if (tid < PB200_SLAB) {
    double f = P.freqs[slab * PB200_SLAB + tid];
    sfreq64[tid] = f;
    float fs = (float)(f * 1e-8);
    sfreq[tid] = fs * fs;
  }
  if (tid == 0) {
    for (int i = 0; i < NSTAGE; ++i) mbar_init(&full[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const float* amp_slab = P.amp + (size_t)slab * P.nsrc_pad * PB200_SLAB;
  auto issue = [&](int tile, int stage) {
    mbar_expect_tx(&full[stage], (uint32_t)sizeof(Tile));
    tma_bulk_g2s(&tiles[stage].amp[0][0], amp_slab + (size_t)tile * T * PB200_SLAB, sizeof(float) * T * PB200_SLAB, &full[stage]);
    tma_bulk_g2s(&tiles[stage].geom[0][0], P.geom + (size_t)tile * T * 4, sizeof(double) * T * 4, &full[stage]);
  };
  if (tid == 0) {
    issue(0, 0);
    if (ntiles > 1) issue(1, 1);
  }

  float acc_re[KT], acc_im[KT];
#pragma unroll
  for (int k = 0; k < KT; ++k) { acc_re[k] = 0.f; acc_im[k] = 0.f; }

  auto flush = [&]() {
    if (valid) {
      double2* row = reinterpret_cast<double2*>(P.vis) + (size_t)b * P.nchan;
#pragma unroll
      for (int k = 0; k < KT; ++k) {
        int ch = kbase + k;
        if (ch < P.nchan) {
          double2 v = row[ch];
          v.x += (double)acc_re[k]; v.y += (double)acc_im[k];
          row[ch] = v;
        }
      }
    }
#pragma unroll
    for (int k = 0; k < KT; ++k) { acc_re[k] = 0.f; acc_im[k] = 0.f; }
  };

  for (int tile = 0; tile < ntiles; ++tile) {
    const int stage = tile & 1;
    mbar_wait(&full[stage], (tile >> 1) & 1);
    const Tile& tl = tiles[stage];
#pragma unroll 1
    for (int s = 0; s < T; ++s) {
      const double4 g = *reinterpret_cast<const double4*>(&tl.geom[s][0]);
      const double tau_g = g.x * bx + g.y * by + g.z * bz;           // baseline_delay_horizon.py:240
      const double tau = tau_g - tau_pc;                             // interferometry.py:6332
      const float4* arow = reinterpret_cast<const float4*>(&tl.amp[s][wc * KT]);
      float kap = 0.f;
      if (TAPER) {
        // w = exp(-1/2 (u_proj/sigma)^2), u_proj^2 = (|b|^2 - (c tau_g)^2) f^2/c^2 (interferometry.py:6262-6283);
        // g.w = 1/2 * 2 ln2 * d^2 * 1e16 * log2(e)  so that  w = exp2(-g.w * (|b/c|^2 - tau_g^2) * (f/1e8)^2)
        kap = (float)(g.w * fmax(blen2 - tau_g * tau_g, 0.0));
      }
      if (DIRECT) {
#pragma unroll
        for (int k4 = 0; k4 < KT / 4; ++k4) {
          const float4 a4 = arow[k4];
          const float av[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int k = 4 * k4 + j;
            float x = (float)(2.0 * frac_turns(tau * sfreq64[wc * KT + k]));
            float sn, cs;
            sincospif(x, &sn, &cs);
            float a = av[j];
            if (TAPER) a *= exp2f(-kap * sfreq[wc * KT + k]);
            acc_re[k] = fmaf(a, cs, acc_re[k]);
            acc_im[k] = fmaf(-a, sn, acc_im[k]);
          }
        }
      } else {
        // anchor phasor exp(-2 pi i tau f_k0) and per-channel rotation exp(-2 pi i tau df)
        float x0 = (float)(2.0 * frac_turns(tau * fk0));
        float xd = (float)(2.0 * frac_turns(tau * P.df));
        float sn, cs, rsn, rcs;
        sincospif(x0, &sn, &cs);
        sincospif(xd, &rsn, &rcs);
        float pr = cs, pi = -sn;
        const float rr = rcs, ri = -rsn;
#pragma unroll
        for (int k4 = 0; k4 < KT / 4; ++k4) {
          const float4 a4 = arow[k4];
          const float av[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int k = 4 * k4 + j;
            float a = av[j];
            if (TAPER) a *= exp2f(-kap * sfreq[wc * KT + k]);
            acc_re[k] = fmaf(a, pr, acc_re[k]);
            acc_im[k] = fmaf(a, pi, acc_im[k]);
            const float nr = fmaf(-pi, ri, pr * rr);
            const float ni = fmaf(pi, rr, pr * ri);
            pr = nr; pi = ni;
          }
        }
      }
    }
    __syncthreads();                                   // everyone is done with this stage
    if (tid == 0 && tile + NSTAGE < ntiles) issue(tile + NSTAGE, stage);
    if (((tile + 1) % FLUSH_TILES) == 0) flush();
  }
  flush();
}

// geometry staging: [nsrc_pad][4] = (l, m, n, taper coefficient), zero rows for padding
__global__ void k_geom_stage(const double* __restrict__ dircos, const double* __restrict__ fwhm_deg, int nsrc,
                              int nsrc_pad, double* __restrict__ geom) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nsrc_pad) return;
  double4 g = make_double4(0, 0, 0, 0);
  if (s < nsrc) {
    g.x = dircos[3 * (size_t)s]; g.y = dircos[3 * (size_t)s + 1]; g.z = dircos[3 * (size_t)s + 2];
    if (fwhm_deg) {
      double d = 2.0 * sin(0.5 * fwhm_deg[s] * 0.017453292519943295769);      // interferometry.py:6268
      // 1/(2 sigma^2) = ln2 * d^2 (:6270); fold the (f/1e8)^2 scaling and log2(e)
      g.w = log(2.0) * d * d * 1.0e16 * 1.4426950408889634074;
    }
  }
  reinterpret_cast<double4*>(geom)[s] = g;
}

}  // namespace

extern "C" int pb200_skyvis(pb200_ctx* ctx, const double* d_dircos, const float* d_amp, int nsrc,
                            const double* d_bl, int nbl, const double* h_pc, const double* h_freqs, int nchan,
                            const double* d_src_fwhm_deg, void* d_vis, int method, void* stream_) {
  if (!ctx) return PB200_EINVAL;
  if (nsrc < 0 || nbl <= 0 || nchan <= 0 || !d_bl || !h_pc || !h_freqs || !d_vis)
    return pb_fail(ctx, PB200_EINVAL, "pb200_skyvis: bad arguments");
  if (nsrc > 0 && (!d_dircos || !d_amp)) return pb_fail(ctx, PB200_EINVAL, "pb200_skyvis: null source arrays");
  if (method < PB200_SKYVIS_AUTO || method > PB200_SKYVIS_DIRECT)
    return pb_fail(ctx, PB200_EINVAL, "pb200_skyvis: unknown method");
  cudaStream_t stream = (cudaStream_t)stream_;
  PB_CUDA(ctx, cudaSetDevice(ctx->device));
  PB_CUDA(ctx, cudaMemsetAsync(d_vis, 0, sizeof(double) * 2 * (size_t)nbl * nchan, stream));
  if (nsrc == 0) return PB200_OK;                      // empty ROI: zeros (interferometry.py:6378-6382)

  // uniform channel grid?  (reference channels are f0 + k*df, run_prisim.py:900)
  const double df = nchan > 1 ? (h_freqs[nchan - 1] - h_freqs[0]) / (double)(nchan - 1) : 0.0;
  bool uniform = true;
  for (int k = 0; k < nchan; ++k)
    if (fabs(h_freqs[k] - (h_freqs[0] + k * df)) > 1e-4) { uniform = false; break; }   // 1e-4 Hz * 1e-5 s = 1e-9 turn
  bool direct = (method == PB200_SKYVIS_DIRECT) || (method == PB200_SKYVIS_AUTO && !uniform);
  if (method == PB200_SKYVIS_RECURRENCE && !uniform)
    return pb_fail(ctx, PB200_EINVAL, "pb200_skyvis: recurrence kernel needs uniformly spaced channels");

  const int nslab = (nchan + PB200_SLAB - 1) / PB200_SLAB;
  const int nchan_pad = nslab * PB200_SLAB;
  const int nsrc_pad = pb200_nsrc_pad(nsrc);
  void *geom, *dfreq;
  int rc = pb_scratch(ctx, 2, sizeof(double) * 4 * (size_t)nsrc_pad, &geom);
  if (rc) return rc;
  rc = pb_scratch(ctx, 3, sizeof(double) * (size_t)nchan_pad, &dfreq);
  if (rc) return rc;
  {
    // padded channel frequencies (pinned staging is unnecessary: nchan doubles)
    double* tmp = new double[nchan_pad];
    for (int k = 0; k < nchan_pad; ++k) tmp[k] = h_freqs[k < nchan ? k : nchan - 1];
    cudaError_t e = cudaMemcpyAsync(dfreq, tmp, sizeof(double) * nchan_pad, cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    delete[] tmp;
    if (e != cudaSuccess) return pb_fail(ctx, PB200_ECUDA, "freq upload: %s", cudaGetErrorString(e));
  }
  k_geom_stage<<<pb_div_up(nsrc_pad, 256), 256, 0, stream>>>(d_dircos, d_src_fwhm_deg, nsrc, nsrc_pad, (double*)geom);
  PB_CHECK_LAUNCH(ctx, "k_geom_stage");

  SkyvisParams P;
  P.amp = d_amp; P.geom = (const double*)geom; P.bl = d_bl; P.freqs = (const double*)dfreq; P.vis = (double*)d_vis;
  P.pc[0] = h_pc[0]; P.pc[1] = h_pc[1]; P.pc[2] = h_pc[2];
  P.f0 = h_freqs[0]; P.df = df;
  P.nsrc_pad = nsrc_pad; P.nbl = nbl; P.nchan = nchan; P.nslab = nslab;
  const size_t smem = NSTAGE * sizeof(Tile) + 64 + PB200_SLAB * (sizeof(float) + sizeof(double));
  dim3 grid(nslab, pb_div_up(nbl, BL_PER_CTA));
  const bool taper = d_src_fwhm_deg != nullptr;
#define LAUNCH(TP, DR)                                                                                    \
  do {                                                                                                    \
    PB_CUDA(ctx, cudaFuncSetAttribute(k_skyvis<TP, DR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    k_skyvis<TP, DR><<<grid, NTHREADS, smem, stream>>>(P);                                                \
  } while (0)
  if (taper && direct) LAUNCH(true, true);
  else if (taper) LAUNCH(true, false);
  else if (direct) LAUNCH(false, true);
  else LAUNCH(false, false);
#undef LAUNCH
  PB_CHECK_LAUNCH(ctx, "k_skyvis");
  return PB200_OK;
}
