// The phase sum  V[b,f] = sum_s amp[s,f] w[s,b,f] exp(-2 pi i f (s_s - s_pc).b / c)  -- the hot kernel.
// Replaces interferometry.py:6155-6165, :6255 (+ baseline_delay_horizon.py:240), :6258-6283 and
// :6332-6340 / :6348-6376 of the reference, which materialise [nsrc,nbl,nchan] complex128 slabs
// and call numpy exp/sum on them.
//
// Design (DESIGN.md section "K1"):
//   * FP32-FMA-issue bound, not HBM and not tensor cores: the phasor depends on (source, baseline, channel) jointly, so the
//     sum is no GEMM; the cost is the complex rotate + accumulate per term on the FMA pipe.
//   * one thread owns one baseline x KT=32 consecutive channels, 64 fp32 accumulators in registers; a warp's 32 lanes are 32
//     baselines of the same channel block, so amplitude reads are shared-memory broadcasts (one LDS.128 = 4 channels for
//     the whole warp).
//   * a CTA = 16 warps = SPC slabs of PB200_SLAB channels x 16/(4 SPC) baseline groups (default SPC = 4: 512 channels x 32
//     baselines); it streams source tiles (32 rows per slab + fp64 geometry) through a double-buffered shared-memory ring
//     filled by TMA bulk copies (cp.async.bulk + mbarrier).
//   * persistent CTAs, one per SM: whole output tiles in waves, the remainder split along the source axis (stream-K) with
//     deterministic head partials added by k_skyvis_finalize (struct Sched).
//   * per tile the CTA first computes, cooperatively and once per (source, baseline), the delay tau = s.b/c - tau_pc in
//     fp64 and from it the channel rotations r^2 (and r^8) to full fp32 accuracy, the MUFU argument increment of one
//     channel and -- default quarter-block form -- the anchor phase and its per-block step as 32-bit fixed-point turn
//     fractions, parked in shared memory for all channel-block warps of that baseline.
//   * each thread anchors its channel block with exp(-2 pi i tau f_k0) on the MUFU (argument range-reduced in fp64 /
//     fixed point and pre-scaled so that the MUFU's own 1/(2 pi) multiply lands on the reduced turn fraction without a
//     systematic bias), both channels of the first pair directly.
//   * channels advance two per packed FFMA2 (sm_100 f32x2): complex rotation by r^2 (MODE 0), or quarter blocks (MODE 3,
//     default): anchors of the other 8-channel quarters by r^8, inside a quarter one rotation and two three-term steps.
//   * fp32 accumulators are flushed into fp64 running sums (scratch in warp-tile layout, L2 evict-last) every
//     FLUSH_TILES source tiles, after every tile for the brightest sources (sorted first by the caller); no atomics.
#include "common.cuh"
#include <cstdlib>
#include <type_traits>

namespace {

constexpr int KT = 32;                         // channels per thread
constexpr int WCS = PB200_SLAB / KT;           // channel blocks per slab = 4
constexpr int NWARPS = 16;
constexpr int NTHREADS = 32 * NWARPS;          // 512
constexpr int T = PB200_SRC_TILE;              // sources per tile
#ifndef PB_FLUSH_TILES
#define PB_FLUSH_TILES 8
#endif
#ifndef PB_STAGGER
#define PB_STAGGER 4
#endif
#ifndef PB_LIFT_MAX_ANGLE
#define PB_LIFT_MAX_ANGLE 2.0   // rad per two-channel step; tan(1) = 1.56 keeps the shear intermediates below 2.6
#endif
#ifndef PB_SRC_UNROLL
#define PB_SRC_UNROLL 4    // sources in flight per thread in the channel loop (measured: 1 -> 4.34, 4 -> 4.42 Tterms/s at the time)
#endif
#ifndef PB_TWO_ANCHOR      // packed recurrence: both channels of the first pair anchored directly (MUFU) and the two-channel rotation r^2 taken
#define PB_TWO_ANCHOR 2    // accurately from the cooperative stage, instead of p1 = p0 r and r^2 = r r in fp32: 8 FMA-pipe ops and the systematic
#endif                     // drift of a twice-rounded r^2 less.  1: second anchor through its own fp64 reduction; 2: by adding the stage's fp32
                           // argument increment.  Measured on config 2, all 6e7 cells (tools/err_c2.py, profiles/err_variants_r02.txt), with
                           // brightest-first ordering: 0 -> 4.55 Tterms/s, max error 6.4e-6;  1 -> 4.50, 4.4e-6;  2 -> 4.64, 4.4e-6;
                           // 2 with PB_FLUSH_TILES = 8 -> 4.62, 3.0e-6 (the default);  re-anchoring every 16 channels bought nothing (4.31, 4.2e-6)
#ifndef PB_L2_HINTS        // L2 cache policies.  1: the fp64 running sums are read and written evict-last, so that the amplitude stream (0.73 GB per
#define PB_L2_HINTS 1      // wave) does not push the 39 MB of live running sums out to DRAM (measured per config-2 launch: 57 -> 1.5 GB written).
#endif                     // 2: additionally the amplitude stream is loaded evict-first -- WORSE: the CTAs of a wave share every amplitude tile
                           // through L2, and evict-first drops it before the last of them has read it (100 -> 296 GB read, 1.6 % slower).
#ifndef PB_ACC_SCALAR      // 1: the accumulates of the packed rotation loop as scalar FFMA (3 x 32-bit operands) instead of FFMA2 (3 x 64-bit)
#define PB_ACC_SCALAR 0
#endif
#ifndef PB_STAGE_FP64      // 1: the stage's rotations r^2 (and r^8) from ONE fp64 sincospi + two fp64 complex squarings (FP64 pipe, idle otherwise)
#define PB_STAGE_FP64 0    // instead of one fp32 sincospif + first-order correction each (~25 FMA-pipe instructions each)
#endif
#ifndef PB_Q3_INT          // quarter-block form: 1 = the per-source anchor argument from 32-bit fixed-point turn fractions staged per (source,
#define PB_Q3_INT 1        // baseline) -- X0 + block * D in one IMAD, then I2F.F64, DMUL, F2F instead of DMUL, 3 DADD, DMUL, F2F -- and the stage
#endif                     // products packed into two 16-byte records (2 LDS.128 instead of 3 LDS.64 + 1 LDS.32 per source): 100 instead of 124
                           // non-FMA instructions per 4 sources.  Config 2, two A/B rounds in one session: 4.84 / 4.85 vs 4.70 / 4.67 Tterms/s,
                           // max error 3.29e-6 vs 3.30e-6 (an fp32 scale constant instead of the fp64 product: 4.90 but 6.5e-6)
#ifndef PB_ABLATE          // developer ablation switches for tools/variants.sh (0 in the product): 1 = skip the per-tile
#define PB_ABLATE 0        // precompute after tile 0, 2 = skip the anchors, 4 = skip the flushes, 8 = constant amplitudes
#endif
constexpr int FLUSH_TILES = PB_FLUSH_TILES;                // fp32 -> fp64 flush cadence (256 sources); round-2 kernel, brightest sources first: max error at C2 / speed 4.4e-6 / 4.64 @16, 3.0e-6 / 4.62 @8, 2.7e-6 / 4.58 @4
constexpr int NSTAGE = 2;
constexpr int SRC_UNROLL = PB_SRC_UNROLL;
constexpr int STAGGER = PB_STAGGER;                     // source-loop chunks per tile between which the warps of a scheduler take turns precomputing
// CTA shape: SPC slabs (SPC*128 channels) x WB baseline groups, SPC*WCS*WB = 16 warps.  A wider
// channel extent shares each (source, baseline) delay/rotation among more warps (less per-tile
// precompute); a wider baseline extent re-reads the amplitude tile less often.
template <int SPC> struct Shape {
  static constexpr int WC = SPC * WCS;                  // channel blocks per CTA
  static constexpr int WB = NWARPS / WC;                // baseline groups per CTA
  static constexpr int BL = 32 * WB;                    // baselines per CTA
  static constexpr int PRE = T * BL / NTHREADS;         // (source, baseline) pairs per thread per tile
};

// 1 / fl32(1/(2 pi)): multiplying the fp64 turn fraction by this and rounding to fp32 makes the
// FMUL by 0.15915494f that sin.approx/cos.approx prepend (SASS: FMUL R, R, 0.15915494; MUFU.SIN)
// return the fraction itself to fp32 rounding, with no coherent phase-scale bias.
#define PB_INV_RCP2PI_F32 (1.0 / 0.15915493667125701904296875)

// Work decomposition of one launch (persistent CTAs, one per SM).  An output tile = one CTA-sized block of
// baselines x channels summed over all S source tiles.  The first `nwave * ncta` output tiles are done whole, one per
// CTA and wave, all CTAs sweeping the source axis together (the amplitude table is then read from DRAM once per wave
// and shared through L2).  The remaining `ntail` (< ncta) output tiles are split ALONG THE SOURCE AXIS over all CTAs
// (stream-K): CTA c owns the contiguous unit range [c U / ncta, (c+1) U / ncta) of the ntail x S (tile, source tile)
// units, so every SM finishes at the same time whatever nbl is -- without this a launch of 480 tiles on 148 SMs (one
// eighth of HERA-350) runs 4 waves for 3.24 waves of work.  A CTA whose range starts inside an output tile writes that
// partial sum to its own "head" slot; k_skyvis_finalize adds the head slots to the tile's own slot in CTA order, so the
// result is deterministic (no atomics).
struct Sched {
  int gx;                  // output tiles along the channel axis (tile t -> x = t % gx, y = t / gx)
  int ntile;               // output tiles
  int S;                   // source tiles per output tile
  int nwave;               // whole-tile waves
  int ntail;               // output tiles of the stream-K tail
  int ncta;                // persistent CTAs (gridDim.x)
};

struct SkyvisParams {
  Sched sc;
  const void* amp;         // [nslab][nsrc_pad][SLAB] fp32 (fp64 for the fp64 kernel with an fp64 table)
  const double* geom;      // [nsrc_pad][4]: l, m, n, taper coefficient
  const double* bl;        // [nbl][3] metres
  const double* freqs;     // device [nchan_pad] Hz (padded channels repeat the last frequency)
  double* vis;             // [nbl][nchan] complex128, rows vis_stride complex elements apart
  long long vis_stride;
  double2* accum;          // fp64 running sums, warp-tile layout [slot][warp][k][lane] (coalesced flushes); slots
                           // 0..ntile-1 = output tiles, ntile + c = head partial of CTA c
  double pc[3];            // phase-centre dircos
  double f0, df;           // uniform channels: f_k = f0 + k df
  int nsrc_pad, nbl, nchan, nslab;
  int bright_tiles;        // the first bright_tiles source tiles hold the brightest sources: flushed to fp64 after every tile
  int kt, wc, wb;          // finalize: channels per thread, channel blocks and baseline groups per CTA of the launch that filled `accum`
  const unsigned* smax2_bits;   // device: float bits of max_s |s - s_pc|^2 (k_geom_stage)
};

template <int SPC> struct __align__(16) TileIn {   // TMA destination
  float amp[SPC][T][PB200_SLAB];               // SPC x 16 KB
  double geom[T][4];                           // 1 KB
};
template <int SPC> struct __align__(16) TilePre {  // produced by the CTA once per tile
  double tau[T][Shape<SPC>::BL];
  float2 rot[T][Shape<SPC>::BL];
  float xd[T][Shape<SPC>::BL];                     // MUFU argument increment of one channel step (PB_TWO_ANCHOR == 2)
};
template <int SPC> struct __align__(16) TileRot8 { float2 rot8[T][Shape<SPC>::BL]; };   // eight-channel rotation r^8 (MODE 3)
template <int SPC> struct __align__(16) TileQ3 {    // PB_Q3_INT: stage products of the quarter-block form, two 16-byte records per pair
  uint4 xa[T][Shape<SPC>::BL];                       // anchor phase at the CTA's first channel and its step per 32-channel block (2^-32 turn), argument increment (float bits), 0
  float4 rr[T][Shape<SPC>::BL];                      // r^2 (x, y), r^8 (z, w)
};
template <int SPC> struct __align__(16) TileKap { float kap[T][Shape<SPC>::BL]; };   // taper exponent coefficient

// fraction of x in [-0.5, 0.5] (round-to-nearest-even magic number; |x| < 2^51)
__device__ __forceinline__ double frac_turns(double x) {
  const double M = 6755399441055744.0;   // 1.5 * 2^52
  double r = __dadd_rn(__dadd_rn(x, M), -M);
  return x - r;
}

// exp(-2 pi i u) for an fp64 phase u in turns: fp64 range reduction, MUFU evaluation
// acc += p * a on a channel pair: packed, or as two scalar FFMA (PB_ACC_SCALAR)
__device__ __forceinline__ float2 acc_pair(float2 p, float2 a, float2 acc) {
#if PB_ACC_SCALAR
  return make_float2(fmaf(p.x, a.x, acc.x), fmaf(p.y, a.y, acc.y));
#else
  return __ffma2_rn(p, a, acc);
#endif
}
__device__ __forceinline__ float anchor_arg(double u) { return (float)(frac_turns(u) * PB_INV_RCP2PI_F32); }
__device__ __forceinline__ float2 mufu_phasor(float x) { return make_float2(__cosf(x), -__sinf(x)); }
__device__ __forceinline__ float2 anchor_phasor(double u) { return mufu_phasor(anchor_arg(u)); }

// exp(-2 pi i d) to full fp32 accuracy: accurate sincospif on the fp32 half-turn argument plus
// the first-order correction for the part of the fp64 argument the fp32 rounding dropped
__device__ __forceinline__ float2 rotation_phasor(double d_turns) {
  const double x = 2.0 * frac_turns(d_turns);          // half-turns in [-1, 1]
  const float x32 = (float)x;
  const float e = (float)(x - (double)x32) * 3.14159265358979f;
  float sn, cs;
  sincospif(x32, &sn, &cs);
  return make_float2(fmaf(-e, sn, cs), -fmaf(e, cs, sn));
}

struct Geometry {
  double bx, by, bz, tau_pc, blen2;
};

__device__ __forceinline__ Geometry load_baseline(const SkyvisParams& P, int b, bool valid) {
  Geometry G = {0, 0, 0, 0, 0};
  if (valid) {   // baseline in light-seconds
    G.bx = P.bl[3 * (size_t)b] / PB_SPEED_OF_LIGHT;
    G.by = P.bl[3 * (size_t)b + 1] / PB_SPEED_OF_LIGHT;
    G.bz = P.bl[3 * (size_t)b + 2] / PB_SPEED_OF_LIGHT;
  }
  G.tau_pc = P.pc[0] * G.bx + P.pc[1] * G.by + P.pc[2] * G.bz;     // interferometry.py:6165
  G.blen2 = G.bx * G.bx + G.by * G.by + G.bz * G.bz;                // (|b|/c)^2, taper
  return G;
}

// One piece of work of a persistent CTA: output tile `tile`, source tiles [s0, s1), partial sums into `slot`.
struct Segment {
  int tile, s0, s1;
  size_t slot;
};

// Iterates the segments of CTA `cta` (see Sched): whole tiles first, then its share of the stream-K tail.
struct SegmentIter {
  const Sched sc;
  const int cta;
  int wave;
  long long u, u1;
  __device__ SegmentIter(const Sched& sc_, int cta_) : sc(sc_), cta(cta_), wave(0) {
    const long long U = (long long)sc.ntail * sc.S;
    u = U * cta / sc.ncta;
    u1 = U * (cta + 1) / sc.ncta;
  }
  __device__ bool next(Segment& sg) {
    if (wave < sc.nwave) {
      sg.tile = wave * sc.ncta + cta; sg.s0 = 0; sg.s1 = sc.S; sg.slot = (size_t)sg.tile;
      ++wave;
      return true;
    }
    if (u >= u1) return false;
    const int q = (int)(u / sc.S);
    sg.s0 = (int)(u - (long long)q * sc.S);
    sg.s1 = (int)min((long long)sc.S, sg.s0 + (u1 - u));
    sg.tile = sc.nwave * sc.ncta + q;
    sg.slot = sg.s0 == 0 ? (size_t)sg.tile : (size_t)sc.ntile + cta;
    u += sg.s1 - sg.s0;
    return true;
  }
};

// Move the fp32 partial sums of one thread into its fp64 running sums and clear them.  The running sums live in a
// scratch buffer laid out [slot][warp][k][lane] so that every load/store of a warp is one contiguous 512-byte run
// (the [nbl][nchan] output layout would put the 32 lanes 16 KB apart); k_skyvis_finalize transposes the scratch into
// the output once at the end.  `first`: the slot holds nothing yet -- store instead of read-modify-write (no memset
// of the scratch, one read less).
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ double2 ld_keep(const double2* p, uint64_t pol) {
  double2 v;
  asm volatile("ld.global.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(pol) : "memory");
  return v;
}
__device__ __forceinline__ void st_keep(double2* p, double2 v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(p), "d"(v.x), "d"(v.y), "l"(pol) : "memory");
}

// `last`: the segment is complete -- its sums are only read again by k_skyvis_finalize, so they are written evict-first
// (a line that stayed tagged evict-last after its tile was done would squat in L2 for the rest of the launch).
__device__ __forceinline__ void flush_acc(const SkyvisParams& P, size_t slot, bool first, float2 (&acc_re)[KT / 2],
                                          float2 (&acc_im)[KT / 2], bool last = false) {
  double2* base = P.accum + ((slot * NWARPS + (threadIdx.x >> 5)) * KT) * 32 + (threadIdx.x & 31);
  const uint64_t keep = PB_L2_HINTS ? (last ? l2_policy_evict_first() : l2_policy_evict_last()) : 0;
#pragma unroll
  for (int k = 0; k < KT; ++k) {
    double2 v = first ? make_double2(0.0, 0.0) : (PB_L2_HINTS ? ld_keep(base + k * 32, keep) : base[k * 32]);
    v.x += (double)((k & 1) ? acc_re[k >> 1].y : acc_re[k >> 1].x);
    v.y += (double)((k & 1) ? acc_im[k >> 1].y : acc_im[k >> 1].x);
    if (PB_L2_HINTS) st_keep(base + k * 32, v, keep);
    else base[k * 32] = v;
  }
#pragma unroll
  for (int k = 0; k < KT / 2; ++k) { acc_re[k] = make_float2(0.f, 0.f); acc_im[k] = make_float2(0.f, 0.f); }
}

// scratch [slot][warp][k][lane] -> vis[b][ch]: one CTA per warp tile, transposed through shared memory.  Adds the head
// partials of the CTAs whose stream-K range starts inside this output tile, in CTA order.
template <int KTV>
__global__ void __launch_bounds__(256) k_skyvis_finalize(const SkyvisParams P) {
  __shared__ double2 tile[KTV][33];
  const Sched sc = P.sc;
  const int nw = P.wb * P.wc;                           // warps per output tile
  const size_t t = blockIdx.x;                          // (output tile * nw + warp)
  const int warp = (int)(t % nw);
  const int ot = (int)(t / nw);
  const int gx = ot % sc.gx, gy = ot / sc.gx;
  const int wb = warp % P.wb, wc = warp / P.wb;
  const size_t wtile = (size_t)KTV * 32;                // double2 per warp tile
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  // head partials: CTAs c with  q S < floor(c U / ncta) < (q+1) S,  q = index of this tile in the tail
  int c_lo = 0, c_hi = 0;
  const int q = ot - sc.nwave * sc.ncta;
  const long long U = (long long)sc.ntail * sc.S;
  if (q >= 0) {
    const long long lo = (long long)q * sc.S, hi = lo + sc.S;
    c_lo = (int)(lo * sc.ncta / U);
    while (c_lo < sc.ncta && U * c_lo / sc.ncta <= lo) ++c_lo;
    c_hi = c_lo;
    while (c_hi < sc.ncta && U * c_hi / sc.ncta < hi) ++c_hi;
  }
  for (int k = ty; k < KTV; k += 8) {
    double2 v = P.accum[t * wtile + k * 32 + tx];
    for (int c = c_lo; c < c_hi; ++c) {
      if (U * (c + 1) / sc.ncta == U * c / sc.ncta) continue;      // empty range: no head partial was written
      const double2 h = P.accum[((size_t)(sc.ntile + c) * nw + warp) * wtile + k * 32 + tx];
      v.x += h.x; v.y += h.y;
    }
    tile[k][tx] = v;
  }
  __syncthreads();
  const int b0 = gy * 32 * P.wb + wb * 32;
  double2* vis = reinterpret_cast<double2*>(P.vis);
  for (int r = ty; r < 32; r += 8) {
    const int b = b0 + r;
#pragma unroll
    for (int kk = 0; kk < KTV; kk += 32) {
      const int k = kk + tx;
      const int ch = (gx * P.wc + wc) * KTV + k;
      if (k < KTV && b < P.nbl && ch < P.nchan) vis[(size_t)b * P.vis_stride + ch] = tile[k][r];
    }
  }
}

// =================================================================================================
// Recurrence kernel (uniform channel grid)
// =================================================================================================
// MODE: phasor stepping of the channel loop -- 0 complex rotation by r^2 (4 ops per two channels), 1 lifted rotation
// (3 shears) on CTA rows of short baselines, 2 three-term recurrence in 16-channel half blocks, 3 "quarter blocks": one MUFU
// anchor pair per source, the anchors of the other three 8-channel quarters by exact rotations r^8, and inside a quarter one
// rotation by r^2 plus two three-term steps (76 instead of 92 packed instructions per source).  TAPER implies MODE 0.
template <int SPC, bool PACKED, bool TAPER, int MODE>
__global__ void __launch_bounds__(NTHREADS, 1) k_skyvis(const SkyvisParams P) {
  static_assert(!TAPER || MODE == 0, "the taper folds into the complex rotation only");
  static_assert(MODE != 1 || PACKED, "the lifted rotation is packed only");
  using S = Shape<SPC>;
  constexpr int WB = S::WB, WC = S::WC;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  TileIn<SPC>* tin = reinterpret_cast<TileIn<SPC>*>(smem_raw);
  constexpr bool Q3I = MODE == 3 && PB_Q3_INT;
  TilePre<SPC>* tpre = reinterpret_cast<TilePre<SPC>*>(smem_raw + NSTAGE * sizeof(TileIn<SPC>));
  TileQ3<SPC>* tq3 = reinterpret_cast<TileQ3<SPC>*>(smem_raw + NSTAGE * sizeof(TileIn<SPC>));      // Q3I: instead of tpre / trot8
  unsigned char* tail = smem_raw + NSTAGE * (sizeof(TileIn<SPC>) + (Q3I ? sizeof(TileQ3<SPC>) : sizeof(TilePre<SPC>)));
  uint64_t* full = reinterpret_cast<uint64_t*>(tail);
  float* sfreq2 = reinterpret_cast<float*>(tail + 64);                     // [SPC*SLAB] (f/1e8)^2, taper only
  TileKap<SPC>* tkap = reinterpret_cast<TileKap<SPC>*>(tail + 64 + SPC * PB200_SLAB * sizeof(float));   // [NSTAGE], taper only
  TileRot8<SPC>* trot8 = reinterpret_cast<TileRot8<SPC>*>(tail + 64 + SPC * PB200_SLAB * sizeof(float));  // [NSTAGE], MODE 3 only (never with the taper)

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wb = warp % WB, wc = warp / WB;           // tid % BL == wb*32 + lane: producer and consumer baseline coincide
  const int sl = wc / WCS, wcs = wc % WCS;            // slab within the CTA, channel block within the slab
  const int bcol = wb * 32 + lane;
  const double df = P.df;
  if (tid == 0) {
    for (int i = 0; i < NSTAGE; ++i) mbar_init(&full[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  float2 acc_re[KT / 2], acc_im[KT / 2];
#pragma unroll
  for (int k = 0; k < KT / 2; ++k) { acc_re[k] = make_float2(0.f, 0.f); acc_im[k] = make_float2(0.f, 0.f); }

  uint32_t fill = 0;                                  // source tiles this CTA has consumed so far: stage = fill & 1, mbarrier parity = (fill >> 1) & 1
  SegmentIter seg_iter(P.sc, (int)blockIdx.x);
  Segment sg;
#pragma unroll 1
  while (seg_iter.next(sg)) {
  const int tile_x = sg.tile % P.sc.gx, tile_y = sg.tile / P.sc.gx;
  const int slab = tile_x * SPC + sl;
  const int b = tile_y * S::BL + bcol;
  const bool valid = b < P.nbl;
  const int kbase = slab * PB200_SLAB + wcs * KT;     // first global channel of this thread
  const int ntiles = sg.s1 - sg.s0;
  const Geometry G = load_baseline(P, b, valid);
  const double fk0 = P.f0 + (double)kbase * P.df;
  const double fcta0 = P.f0 + (double)(tile_x * SPC * PB200_SLAB) * P.df;   // first channel of the CTA tile (Q3I)

  if (TAPER && tid < SPC * PB200_SLAB) {
    const int ch = tile_x * SPC * PB200_SLAB + tid;
    const float fs = (float)(P.freqs[ch < P.nslab * PB200_SLAB ? ch : P.nslab * PB200_SLAB - 1] * 1e-8);
    sfreq2[tid] = fs * fs;
  }
  // taper recurrence constants (uniform grid): F_k = F0 + k dF in units of 1e8 Hz
  const float tF0 = (float)(fk0 * 1e-8), tdF = (float)(df * 1e-8);

  // slabs of this output tile that exist (the last tile along x may be short when nslab % SPC != 0)
  const int nsl = min(SPC, P.nslab - tile_x * SPC);
  auto issue = [&](int i, int stage) {                 // source tile sg.s0 + i of this segment
    const size_t row0 = (size_t)(sg.s0 + i) * T;
    mbar_expect_tx(&full[stage], (uint32_t)(nsl * sizeof(float) * T * PB200_SLAB + sizeof(double) * T * 4));
    const uint64_t stream_pol = PB_L2_HINTS == 2 ? l2_policy_evict_first() : 0;
    for (int j = 0; j < nsl; ++j) {
      const float* src = (const float*)P.amp + ((size_t)(tile_x * SPC + j) * P.nsrc_pad + row0) * PB200_SLAB;
      if (PB_L2_HINTS == 2) tma_bulk_g2s_hint(&tin[stage].amp[j][0][0], src, sizeof(float) * T * PB200_SLAB, &full[stage], stream_pol);
      else tma_bulk_g2s(&tin[stage].amp[j][0][0], src, sizeof(float) * T * PB200_SLAB, &full[stage]);
    }
    tma_bulk_g2s(&tin[stage].geom[0][0], P.geom + row0 * 4, sizeof(double) * T * 4, &full[stage]);
  };
  if (tid == 0) {
    issue(0, fill & 1);
    if (ntiles > 1) issue(1, (fill + 1) & 1);
  }
  const bool live = sl < nsl;                          // warps of a missing slab only help with the precompute

  // Lifted rotation (3 FFMA2 instead of 2 FMUL2 + 2 FFMA2 per two-channel step): a rotation by phi is the shear
  // product  x += t y;  y += s x;  x += t y  with t = -tan(phi/2), s = sin(phi).  Determinant 1 by construction and
  // more accurate than the complex multiply while |phi| <= PB_LIFT_MAX_ANGLE (measured in tools/lift_accuracy.py);
  // tan blows up towards pi, so a CTA row uses it only if every (source, baseline) pair of the row stays below the
  // limit:  |phi| = 4 pi df |tau|,  |tau| <= |b|/c max_s |s - s_pc|.  The decision is CTA-uniform.
  constexpr bool three_term = MODE == 2;
  constexpr bool two_anchor = PACKED && MODE != 2 && PB_TWO_ANCHOR;      // the stage stores r^2, the loop anchors channels k0 and k0 + 1
  bool lift_row = false;
  if (MODE == 1) {
    const float smax2 = __uint_as_float(*P.smax2_bits);
    const double lim = PB_LIFT_MAX_ANGLE / (4.0 * 3.14159265358979323846 * fabs(df));
    lift_row = !__syncthreads_or(valid && G.blen2 * (double)smax2 > lim * lim);
  }

  // cooperative per-tile stage: tau and the channel rotation for S::PRE sources of this
  // thread's own baseline (sources s = wc, wc+WC, ...)
  auto precompute = [&](int tile) {
    const int stage = (fill + tile) & 1;
    mbar_wait(&full[stage], ((fill + tile) >> 1) & 1);
#pragma unroll 2
    for (int j = 0; j < S::PRE; ++j) {
      const int s = wc + WC * j;
      const double4 g = *reinterpret_cast<const double4*>(&tin[stage].geom[s][0]);
      const double tau_g = g.x * G.bx + g.y * G.by + g.z * G.bz;            // baseline_delay_horizon.py:240
      const double tau = tau_g - G.tau_pc;                                  // interferometry.py:6332
      if constexpr (Q3I) {
        const int X0 = (int)__double2ll_rn(frac_turns(tau * fcta0) * 4294967296.0);           // wraps at +-1/2 turn, as the phase does
        const int D = (int)__double2ll_rn(frac_turns(tau * (32.0 * df)) * 4294967296.0);
        tq3[stage].xa[s][bcol] = make_uint4((unsigned)X0, (unsigned)D, __float_as_uint(anchor_arg(tau * df)), 0u);
        const float2 r2 = rotation_phasor(2.0 * tau * df), r8 = rotation_phasor(8.0 * tau * df);
        tq3[stage].rr[s][bcol] = make_float4(r2.x, r2.y, r8.x, r8.y);
        continue;
      }
      tpre[stage].tau[s][bcol] = tau;
      float2 rp;
      if (PB_STAGE_FP64 && MODE == 3) {
        double sn, cs;
        sincospi(2.0 * frac_turns(2.0 * tau * df) , &sn, &cs);                       // exp(-2 pi i 2 tau df) = (cs, -sn)
        rp = make_float2((float)cs, (float)(-sn));
        double ar = cs, ai = -sn;
#pragma unroll
        for (int i = 0; i < 2; ++i) { const double t = fma(ar, ar, -ai * ai); ai = 2.0 * ar * ai; ar = t; }
        trot8[stage].rot8[s][bcol] = make_float2((float)ar, (float)ai);
      } else {
        rp = rotation_phasor((three_term || two_anchor) ? 2.0 * tau * df : tau * df);   // packed rows: the two-channel rotation r^2
        if (MODE == 3) trot8[stage].rot8[s][bcol] = rotation_phasor(8.0 * tau * df);
      }
      // lifted rows: shear coefficients of the two-channel step, t = -tan(phi) and s = sin(2 phi) for r = e^{i phi}
      if (lift_row) rp = two_anchor ? make_float2(-__fdiv_rn(rp.y, 1.0f + rp.x), rp.y) : make_float2(-__fdiv_rn(rp.y, rp.x), 2.0f * rp.x * rp.y);
      tpre[stage].rot[s][bcol] = rp;
      if (PB_TWO_ANCHOR == 2 && two_anchor) tpre[stage].xd[s][bcol] = anchor_arg(tau * df);
      if (TAPER) {
        // w = exp(-1/2 (u_proj/sigma)^2), u_proj^2 = (|b|^2 - (c tau_g)^2) f^2/c^2 (interferometry.py:6262-6283);
        // g.w = ln2 d^2 1e16 log2(e)  so that  w = exp2(-g.w (|b/c|^2 - tau_g^2) (f/1e8)^2); sqrt argument clamped at 0
        tkap[stage].kap[s][bcol] = (float)(g.w * fmax(G.blen2 - tau_g * tau_g, 0.0));
      }
    }
  };

  precompute(0);
  __syncthreads();

  auto run_tile = [&](int tile, auto lift_c) {
    constexpr bool LIFT = decltype(lift_c)::value;
    const int stage = (fill + tile) & 1;
    const TileIn<SPC>& ti = tin[stage];
    const TilePre<SPC>& tp = tpre[stage];
    // software pipeline over sources: the anchor(s) of source s+1 are evaluated while the channel loop of s runs
    // anchors of channels k0 and k0 + 1 of source s: exp(-2 pi i tau f) with the product reduced in fp64; with
    // PB_TWO_ANCHOR == 2 the second one adds the per-channel argument increment from the stage in fp32 (|sum| <= 2 pi)
    auto anchors2 = [&](int s, float2& p, float2& q) {
      const float x = anchor_arg(tp.tau[s][bcol] * fk0);
      p = mufu_phasor(x);
      if (PB_TWO_ANCHOR == 2 && two_anchor && !LIFT) q = mufu_phasor(x + tp.xd[s][bcol]);
      else if (LIFT || two_anchor) q = anchor_phasor(tp.tau[s][bcol] * (fk0 + df));
      else q = make_float2(0.f, 0.f);
    };
    float2 p_next, q_next;
    anchors2(0, p_next, q_next);
    float2 r_next = tp.rot[0][bcol];
    // the next tile's precompute (fp64 / XU / ALU work, no FFMA2) is staggered over the four warps of a
    // scheduler (warp >> 2 = index within the scheduler): at any time at most one of them is off the FMA
    // pipe and the other three keep it fed, instead of all 16 warps leaving it idle together
#pragma unroll 1
    for (int chunk = 0; chunk < STAGGER; ++chunk) {
    if (chunk == (warp >> 2) % STAGGER && tile + 1 < ntiles && !((PB_ABLATE & 1) && tile > 0)) precompute(tile + 1);
#pragma unroll SRC_UNROLL
    for (int s = chunk * (T / STAGGER); s < (chunk + 1) * (T / STAGGER); ++s) {
      const float2 p0 = p_next, q0 = q_next, r = r_next;
      const int sn = (s + 1 < T) ? s + 1 : s;
      if (PB_ABLATE & 2) { p_next = tp.rot[sn][bcol ^ 1]; q_next = tp.rot[sn][bcol ^ 2]; }
      else anchors2(sn, p_next, q_next);
      r_next = tp.rot[sn][bcol];
      const float4* arow = reinterpret_cast<const float4*>(&ti.amp[sl][s][wcs * KT]);
      float kap = 0.f;
      if (TAPER) kap = tkap[stage].kap[s][bcol];

      if (LIFT) {
        // both channels of the pair are anchored directly (the XU and FP64 pipes are idle under the FFMA2
        // stream), then stepped two channels at a time by three packed shears; r = (t, s)
        float2 PR = make_float2(p0.x, q0.x), PI = make_float2(p0.y, q0.y);
        const float2 TT = make_float2(r.x, r.x), SS = make_float2(r.y, r.y);
#pragma unroll
        for (int k4 = 0; k4 < KT / 4; ++k4) {
          const float4 a4 = (PB_ABLATE & 8) ? make_float4(kap + 1.f, kap + 2.f, kap + 3.f, kap + 4.f) : arow[k4];
          const float2 A0 = make_float2(a4.x, a4.y), A1 = make_float2(a4.z, a4.w);
          acc_re[2 * k4] = __ffma2_rn(PR, A0, acc_re[2 * k4]);
          acc_im[2 * k4] = __ffma2_rn(PI, A0, acc_im[2 * k4]);
          float2 x1 = __ffma2_rn(PI, TT, PR);
          PI = __ffma2_rn(x1, SS, PI);
          PR = __ffma2_rn(PI, TT, x1);
          acc_re[2 * k4 + 1] = __ffma2_rn(PR, A1, acc_re[2 * k4 + 1]);
          acc_im[2 * k4 + 1] = __ffma2_rn(PI, A1, acc_im[2 * k4 + 1]);
          if (k4 + 1 < KT / 4) {
            x1 = __ffma2_rn(PI, TT, PR);
            PI = __ffma2_rn(x1, SS, PI);
            PR = __ffma2_rn(PI, TT, x1);
          }
        }
      } else if (PACKED) {
        // two channels per packed register: P = (p_k, p_{k+1}), stepped by r^2
        const float p1r = two_anchor ? q0.x : fmaf(-p0.y, r.y, p0.x * r.x), p1i = two_anchor ? q0.y : fmaf(p0.y, r.x, p0.x * r.y);
        const float r2r = two_anchor ? r.x : fmaf(-r.y, r.y, r.x * r.x), r2i = two_anchor ? r.y : 2.0f * r.x * r.y;
        float2 PR = make_float2(p0.x, p1r), PI = make_float2(p0.y, p1i);
        float2 RR = make_float2(r2r, r2r), RI = make_float2(r2i, r2i);
        float2 HM1 = make_float2(0.f, 0.f);
        if (TAPER) {
          // Gaussian taper w_k = exp2(-kap F_k^2) by recurrence instead of one MUFU per term:
          //   w_{k+2} = w_k G_k,  G_k = exp2(-kap (4 F_k dF + 4 dF^2)),  G_{k+2} = G_k H,  H = exp2(-8 kap dF^2)
          // folded into the phasor (q = p w) and the rotation (R_k = r^2 G_k, R_{k+2} = R_k + R_k (H - 1));
          // H - 1 is carried separately because H rounds to 1 - O(1e-6) in fp32.
          const float F1 = tF0 + tdF;
          const float w0 = exp2f(-kap * tF0 * tF0), w1 = exp2f(-kap * F1 * F1);
          const float c4 = 4.0f * kap * tdF;
          const float g0 = exp2f(-c4 * (tF0 + tdF)), g1 = exp2f(-c4 * (F1 + tdF));
          const float y = -8.0f * kap * tdF * tdF * 0.69314718056f;
          const float hm1 = fmaf(0.5f * y, y, y);
          PR.x *= w0; PR.y *= w1; PI.x *= w0; PI.y *= w1;
          RR = make_float2(r2r * g0, r2r * g1); RI = make_float2(r2i * g0, r2i * g1);
          HM1 = make_float2(hm1, hm1);
        }
        float2 NRI = make_float2(-RI.x, -RI.y);
#pragma unroll
        for (int k4 = 0; k4 < KT / 4; ++k4) {
          const float4 a4 = (PB_ABLATE & 8) ? make_float4(kap + 1.f, kap + 2.f, kap + 3.f, kap + 4.f) : arow[k4];
          const float2 A0 = make_float2(a4.x, a4.y), A1 = make_float2(a4.z, a4.w);
          // operand order chosen for the register-reuse cache: every packed instruction reads at
          // most two fresh 64-bit operands (PR / PI / A stay in the same operand slot across
          // consecutive instructions), which keeps FFMA2 at its 2-cycle pipe rate
          float2 t1 = __fmul2_rn(PR, RR), t2 = __fmul2_rn(PR, RI);
          acc_re[2 * k4] = acc_pair(PR, A0, acc_re[2 * k4]);
          acc_im[2 * k4] = acc_pair(PI, A0, acc_im[2 * k4]);
          float2 nr = __ffma2_rn(PI, NRI, t1);
          float2 ni = __ffma2_rn(PI, RR, t2);
          PR = nr; PI = ni;
          if (TAPER) { RR = __ffma2_rn(RR, HM1, RR); RI = __ffma2_rn(RI, HM1, RI); NRI = make_float2(-RI.x, -RI.y); }
          if (k4 + 1 < KT / 4) {
            t1 = __fmul2_rn(PR, RR); t2 = __fmul2_rn(PR, RI);
          }
          acc_re[2 * k4 + 1] = acc_pair(PR, A1, acc_re[2 * k4 + 1]);
          acc_im[2 * k4 + 1] = acc_pair(PI, A1, acc_im[2 * k4 + 1]);
          if (k4 + 1 < KT / 4) {
            nr = __ffma2_rn(PI, NRI, t1);
            ni = __ffma2_rn(PI, RR, t2);
            PR = nr; PI = ni;
            if (TAPER) { RR = __ffma2_rn(RR, HM1, RR); RI = __ffma2_rn(RI, HM1, RI); NRI = make_float2(-RI.x, -RI.y); }
          }
        }
      } else {
        float pr = p0.x, pi = p0.y;
        const float rr = r.x, ri = r.y;
#pragma unroll
        for (int k4 = 0; k4 < KT / 4; ++k4) {
          const float4 a4 = arow[k4];
          const float av[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int k = 4 * k4 + j;
            float a = av[j];
            if (TAPER) a *= exp2f(-kap * sfreq2[wc * KT + k]);
            if (k & 1) { acc_re[k >> 1].y = fmaf(a, pr, acc_re[k >> 1].y); acc_im[k >> 1].y = fmaf(a, pi, acc_im[k >> 1].y); }
            else       { acc_re[k >> 1].x = fmaf(a, pr, acc_re[k >> 1].x); acc_im[k >> 1].x = fmaf(a, pi, acc_im[k >> 1].x); }
            const float nr = fmaf(-pi, ri, pr * rr);
            const float ni = fmaf(pi, rr, pr * ri);
            pr = nr; pi = ni;
          }
        }
      }
    }
    }
  };

  // Three-term recurrence (method PB200_SKYVIS_RECURRENCE_3TERM): Z_{j+1} = 2 cos(2 phi) Z_j - Z_{j-1} on packed channel
  // pairs Z_j = (z_2j, z_2j+1): ONE FFMA2 per component and step and no intermediate results, where the complex
  // rotation needs two and writes two temporaries -- the loop is bound by register-file result bandwidth (DESIGN.md
  // K1).  The recurrence error grows like (steps^2 / 2) * ulp(2 cos), so a 32-channel block is split into two
  // 16-channel halves of 8 steps: Z_0 from two MUFU anchors (free: XU/FP64 pipes are idle), Z_1 = Z_0 r^2 by one
  // complex rotation, Z_2..Z_7 by the recurrence.
  auto run_tile3 = [&](int tile) {
    // PACKED: FFMA2/FMUL2 on (even, odd) channel pairs; otherwise the same arithmetic as scalar FFMA on the two halves
    auto fma2 = [](float2 a, float2 b, float2 c) { return PACKED ? __ffma2_rn(a, b, c) : make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); };
    auto mul2 = [](float2 a, float2 b) { return PACKED ? __fmul2_rn(a, b) : make_float2(a.x * b.x, a.y * b.y); };
    const int stage = (fill + tile) & 1;
    const TileIn<SPC>& ti = tin[stage];
    const TilePre<SPC>& tp = tpre[stage];
    constexpr int H = KT / 2;                            // channels per half block
    auto anchors = [&](int s, float2 (&an)[4]) {
      const double tau = tp.tau[s][bcol];
      const double d = frac_turns(tau * df);
      const double f0 = frac_turns(tau * fk0), fh = frac_turns(f0 + (double)H * d);
      const float x0 = (float)(f0 * PB_INV_RCP2PI_F32), x1 = (float)((f0 + d) * PB_INV_RCP2PI_F32);
      const float x2 = (float)(fh * PB_INV_RCP2PI_F32), x3 = (float)((fh + d) * PB_INV_RCP2PI_F32);
      an[0] = make_float2(__cosf(x0), -__sinf(x0)); an[1] = make_float2(__cosf(x1), -__sinf(x1));
      an[2] = make_float2(__cosf(x2), -__sinf(x2)); an[3] = make_float2(__cosf(x3), -__sinf(x3));
    };
    float2 an_next[4];
    anchors(0, an_next);
    float2 r_next = tp.rot[0][bcol];
#pragma unroll 1
    for (int chunk = 0; chunk < STAGGER; ++chunk) {
    if (chunk == (warp >> 2) % STAGGER && tile + 1 < ntiles) precompute(tile + 1);
#pragma unroll 2
    for (int s = chunk * (T / STAGGER); s < (chunk + 1) * (T / STAGGER); ++s) {
      float2 an[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) an[i] = an_next[i];
      const float2 r = r_next;
      const int sn = (s + 1 < T) ? s + 1 : s;
      if (PB_ABLATE & 2) { an_next[0] = tp.rot[sn][bcol ^ 1]; an_next[1] = tp.rot[sn][bcol ^ 2]; an_next[2] = tp.rot[sn][bcol ^ 3]; an_next[3] = tp.rot[sn][bcol ^ 4]; }
      else anchors(sn, an_next);
      r_next = tp.rot[sn][bcol];
      const float4* arow = reinterpret_cast<const float4*>(&ti.amp[sl][s][wcs * KT]);
      const float2 RR = make_float2(r.x, r.x), RI = make_float2(r.y, r.y), NRI = make_float2(-r.y, -r.y);
      const float2 CC = make_float2(2.0f * r.x, 2.0f * r.x);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float2 PRm = make_float2(an[2 * h].x, an[2 * h + 1].x), PIm = make_float2(an[2 * h].y, an[2 * h + 1].y);   // Z_0
        const float2 t1 = mul2(PRm, RR), t2 = mul2(PRm, RI);
        float2 PRc = fma2(PIm, NRI, t1), PIc = fma2(PIm, RR, t2);                                       // Z_1
#pragma unroll
        for (int k4 = 0; k4 < H / 4; ++k4) {
          const float4 a4 = (PB_ABLATE & 8) ? make_float4(r.x + 1.f, r.x + 2.f, r.y + 3.f, r.y + 4.f) : arow[h * (H / 4) + k4];
          const float2 A0 = make_float2(a4.x, a4.y), A1 = make_float2(a4.z, a4.w);
          const int j = h * (H / 2) + 2 * k4;             // accumulator (channel pair) index of A0
          if (k4 > 0) {                                   // advance two steps: (m, c) <- (c, C c - m), twice
            const float2 PRn = fma2(CC, PRc, make_float2(-PRm.x, -PRm.y));
            const float2 PIn = fma2(CC, PIc, make_float2(-PIm.x, -PIm.y));
            PRm = fma2(CC, PRn, make_float2(-PRc.x, -PRc.y));
            PIm = fma2(CC, PIn, make_float2(-PIc.x, -PIc.y));
            // now (PRn, PIn) = Z_2k4 and (PRm, PIm) = Z_2k4+1: rename so that m < c again
            const float2 tr = PRm, tii = PIm;
            PRm = PRn; PIm = PIn; PRc = tr; PIc = tii;
          }
          acc_re[j] = fma2(PRm, A0, acc_re[j]);
          acc_im[j] = fma2(PIm, A0, acc_im[j]);
          acc_re[j + 1] = fma2(PRc, A1, acc_re[j + 1]);
          acc_im[j + 1] = fma2(PIc, A1, acc_im[j + 1]);
        }
      }
    }
    }
  };

  // MODE 3: quarter blocks.  Per source ONE pair of MUFU anchors (channels k0, k0 + 1; as in mode 0), the anchor pairs of the
  // other three 8-channel quarters by rotating with the accurately rounded r^8 of the stage, and inside a quarter
  //   P1 = P0 r^2 (complex, 4 packed instructions),  P2 = 2 cos(2 phi) P1 - P0,  P3 = 2 cos(2 phi) P2 - P1  (2 each):
  // 8 + 8 accumulate = 16 packed instructions per quarter + 4 for the next quarter's anchor = 76 per source instead of the 92 of
  // the pure rotation, with recurrence runs of two steps (error growth of the three-term form is steps^2 / 2 ulp).
  auto run_tile4 = [&](int tile) {
    const int stage = (fill + tile) & 1;
    const TileIn<SPC>& ti = tin[stage];
    const TilePre<SPC>& tp = tpre[stage];
    const TileRot8<SPC>& t8 = trot8[stage];
    const TileQ3<SPC>& tq = tq3[stage];
    auto anchors2 = [&](int s, float2& p, float2& q) {
      if constexpr (Q3I) {
        const uint4 a = tq.xa[s][bcol];
        // the scale to MUFU units stays an fp64 product rounded once (an fp32 constant would bias every phase by its rounding error;
        // measured: max error 6.5e-6 instead of 3.3e-6 on config 2)
        const float x = (float)((double)(int)(a.x + (unsigned)wc * a.y) * (PB_INV_RCP2PI_F32 / 4294967296.0));
        p = mufu_phasor(x);
        q = mufu_phasor(x + __uint_as_float(a.z));
      } else {
        const float x = anchor_arg(tp.tau[s][bcol] * fk0);
        p = mufu_phasor(x);
        q = mufu_phasor(x + tp.xd[s][bcol]);
      }
    };
    auto rotations = [&](int s, float2& r, float2& r8) {
      if constexpr (Q3I) {
        const float4 v = tq.rr[s][bcol];
        r = make_float2(v.x, v.y); r8 = make_float2(v.z, v.w);
      } else {
        r = tp.rot[s][bcol]; r8 = t8.rot8[s][bcol];
      }
    };
    float2 p_next, q_next;
    anchors2(0, p_next, q_next);
    float2 r_next, r8_next;
    rotations(0, r_next, r8_next);
#pragma unroll 1
    for (int chunk = 0; chunk < STAGGER; ++chunk) {
    if (chunk == (warp >> 2) % STAGGER && tile + 1 < ntiles) precompute(tile + 1);
#pragma unroll SRC_UNROLL
    for (int s = chunk * (T / STAGGER); s < (chunk + 1) * (T / STAGGER); ++s) {
      const float2 p0 = p_next, q0 = q_next, r = r_next, r8 = r8_next;
      const int sn = (s + 1 < T) ? s + 1 : s;
      anchors2(sn, p_next, q_next);
      rotations(sn, r_next, r8_next);
      const float4* arow = reinterpret_cast<const float4*>(&ti.amp[sl][s][wcs * KT]);
      const float2 RR = make_float2(r.x, r.x), RI = make_float2(r.y, r.y), NRI = make_float2(-r.y, -r.y);
      const float2 CC = make_float2(2.0f * r.x, 2.0f * r.x);
      const float2 ER = make_float2(r8.x, r8.x), EI = make_float2(r8.y, r8.y), NEI = make_float2(-r8.y, -r8.y);
      float2 PR = make_float2(p0.x, q0.x), PI = make_float2(p0.y, q0.y);          // pair 0 of quarter 0
#pragma unroll
      for (int qd = 0; qd < KT / 8; ++qd) {
        const float4 a4 = arow[2 * qd], b4 = arow[2 * qd + 1];
        const float2 A0 = make_float2(a4.x, a4.y), A1 = make_float2(a4.z, a4.w), A2 = make_float2(b4.x, b4.y), A3 = make_float2(b4.z, b4.w);
        const int j = 4 * qd;
        float2 QR = PR, QI = PI;
        if (qd + 1 < KT / 8) {                               // anchor pair of the next quarter: this one rotated by r^8
          const float2 u1 = __fmul2_rn(PR, ER), u2 = __fmul2_rn(PR, EI);
          QR = __ffma2_rn(PI, NEI, u1);
          QI = __ffma2_rn(PI, ER, u2);
        }
        const float2 t1 = __fmul2_rn(PR, RR), t2 = __fmul2_rn(PR, RI);
        acc_re[j] = __ffma2_rn(PR, A0, acc_re[j]);
        acc_im[j] = __ffma2_rn(PI, A0, acc_im[j]);
        const float2 P1R = __ffma2_rn(PI, NRI, t1), P1I = __ffma2_rn(PI, RR, t2);      // pair 1 = pair 0 x r^2
        acc_re[j + 1] = __ffma2_rn(P1R, A1, acc_re[j + 1]);
        acc_im[j + 1] = __ffma2_rn(P1I, A1, acc_im[j + 1]);
        const float2 P2R = __ffma2_rn(CC, P1R, make_float2(-PR.x, -PR.y)), P2I = __ffma2_rn(CC, P1I, make_float2(-PI.x, -PI.y));
        acc_re[j + 2] = __ffma2_rn(P2R, A2, acc_re[j + 2]);
        acc_im[j + 2] = __ffma2_rn(P2I, A2, acc_im[j + 2]);
        const float2 P3R = __ffma2_rn(CC, P2R, make_float2(-P1R.x, -P1R.y)), P3I = __ffma2_rn(CC, P2I, make_float2(-P1I.x, -P1I.y));
        acc_re[j + 3] = __ffma2_rn(P3R, A3, acc_re[j + 3]);
        acc_im[j + 3] = __ffma2_rn(P3I, A3, acc_im[j + 3]);
        PR = QR; PI = QI;
      }
    }
    }
  };

  bool fresh = true;                                   // nothing of this segment is in its slot yet
  for (int tile = 0; tile < ntiles; ++tile) {
    if (!live && tile + 1 < ntiles) precompute(tile + 1);
    if (live) {
      if constexpr (MODE == 3) run_tile4(tile);
      else if constexpr (MODE == 2) run_tile3(tile);
      else if constexpr (MODE == 1) { if (lift_row) run_tile(tile, std::true_type()); else run_tile(tile, std::false_type()); }
      else run_tile(tile, std::false_type());
    }
    __syncthreads();                                   // tile consumed, next tile's tau/rot visible
    if (tid == 0 && tile + NSTAGE < ntiles) issue(tile + NSTAGE, (fill + tile) & 1);
    // flushes are staggered over the warps of a scheduler like the precompute (accumulators are thread-private,
    // so a warp may flush at any tile boundary): one warp waits on its global read-modify-write, three keep going
    // The fp32 partial sums of the brightest sources (sorted first by the caller, P.bright_tiles) go to fp64 after every
    // tile: an fp32 add rounds at ulp(|partial sum|), which the few bright sources would otherwise set for everybody else.
    const bool bright = sg.s0 + tile < P.bright_tiles;
    if (live && (bright || ((tile + 1 + (warp >> 2) * (FLUSH_TILES / 4)) % FLUSH_TILES) == 0) && !(PB_ABLATE & 4)) {
      flush_acc(P, sg.slot, fresh, acc_re, acc_im);
      fresh = false;
    }
  }
  if (live) flush_acc(P, sg.slot, fresh, acc_re, acc_im, true);
  fill += (uint32_t)ntiles;
  }   // segments
}

// =================================================================================================
// Pair form of the quarter-block recurrence: one thread = TWO baselines x 16 channels
// =================================================================================================
// Same arithmetic per (baseline, channel) as MODE 3 of k_skyvis, other ownership: a lane owns baselines `lane` and `lane + 32`
// of a 64-baseline CTA row and 16 consecutive channels (two 8-channel quarters), so that every broadcast LDS.128 of four
// amplitudes feeds two baselines (4 instead of 8 amplitude loads per 32 terms; the round-1 ablation priced the amplitude
// loads at 15 % of the loop).  The price is one anchor (fp64 range reduction + 4 MUFU) per 16 instead of per 32 terms, on pipes
// that are otherwise idle.  CTA = 16 warps = 16 channel blocks (2 slabs, 256 channels) x 64 baselines; the cooperative stage
// is shared by 16 channel-block warps as in the 4-slab shape of k_skyvis.  Scratch layout: each thread writes its two
// baselines as two "virtual warps" (2 * warp + half) of 16 channels x 32 lanes, so k_skyvis_finalize<16> with wc = 16,
// wb = 2 transposes it like any other launch.
#ifndef PB_PAIR_UNROLL
#define PB_PAIR_UNROLL 1
#endif
#ifndef PB_PAIR_UNROLL
#define PB_PAIR_UNROLL 1
#endif
constexpr int PAIR_UNROLL = PB_PAIR_UNROLL;
constexpr int KTP = 16;                       // channels per thread
constexpr int SPCP = 2;                       // slabs per CTA
constexpr int WCP = SPCP * PB200_SLAB / KTP;  // 16 channel blocks = warps
constexpr int BLP = 64;                       // baselines per CTA
static_assert(WCP == NWARPS, "pair form: one warp per channel block");
struct __align__(16) TilePreP {
  double tau[T][BLP];
  float2 rot[T][BLP];                         // r^2
  float2 rot8[T][BLP];                        // r^8
  float xd[T][BLP];                           // MUFU argument increment of one channel
};

__global__ void __launch_bounds__(NTHREADS, 1) k_skyvis_pair(const SkyvisParams P) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  TileIn<SPCP>* tin = reinterpret_cast<TileIn<SPCP>*>(smem_raw);
  TilePreP* tpre = reinterpret_cast<TilePreP*>(smem_raw + NSTAGE * sizeof(TileIn<SPCP>));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + NSTAGE * (sizeof(TileIn<SPCP>) + sizeof(TilePreP)));

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int sl = warp / (WCP / SPCP), wcs = warp % (WCP / SPCP);   // slab within the CTA, 16-channel block within the slab
  const int scol = tid & (BLP - 1), sgrp = tid / BLP;              // stage: baseline column and source group (8 groups x 4 sources)
  const double df = P.df;
  if (tid == 0) {
    for (int i = 0; i < NSTAGE; ++i) mbar_init(&full[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  float2 acc_re[2][KTP / 2], acc_im[2][KTP / 2];
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int k = 0; k < KTP / 2; ++k) { acc_re[h][k] = make_float2(0.f, 0.f); acc_im[h][k] = make_float2(0.f, 0.f); }

  // fp32 partial sums -> fp64 running sums of the segment's slot (see flush_acc), one virtual warp per baseline half
  auto flush = [&](size_t slot, bool first, bool last) {
    const uint64_t keep = PB_L2_HINTS ? (last ? l2_policy_evict_first() : l2_policy_evict_last()) : 0;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      double2* base = P.accum + ((slot * (2 * NWARPS) + 2 * warp + h) * KTP) * 32 + lane;
#pragma unroll
      for (int k = 0; k < KTP; ++k) {
        double2 v = first ? make_double2(0.0, 0.0) : (PB_L2_HINTS ? ld_keep(base + k * 32, keep) : base[k * 32]);
        v.x += (double)((k & 1) ? acc_re[h][k >> 1].y : acc_re[h][k >> 1].x);
        v.y += (double)((k & 1) ? acc_im[h][k >> 1].y : acc_im[h][k >> 1].x);
        if (PB_L2_HINTS) st_keep(base + k * 32, v, keep);
        else base[k * 32] = v;
      }
#pragma unroll
      for (int k = 0; k < KTP / 2; ++k) { acc_re[h][k] = make_float2(0.f, 0.f); acc_im[h][k] = make_float2(0.f, 0.f); }
    }
  };

  uint32_t fill = 0;
  SegmentIter seg_iter(P.sc, (int)blockIdx.x);
  Segment sg;
#pragma unroll 1
  while (seg_iter.next(sg)) {
    const int tile_x = sg.tile % P.sc.gx, tile_y = sg.tile / P.sc.gx;
    const int kbase = (tile_x * SPCP + sl) * PB200_SLAB + wcs * KTP;     // first global channel of this thread
    const int ntiles = sg.s1 - sg.s0;
    const int bs = tile_y * BLP + scol;                                    // the baseline this thread stages
    const Geometry G = load_baseline(P, bs, bs < P.nbl);
    const double fk0 = P.f0 + (double)kbase * P.df;
    auto issue = [&](int i, int stage) {
      const size_t row0 = (size_t)(sg.s0 + i) * T;
      mbar_expect_tx(&full[stage], (uint32_t)(SPCP * sizeof(float) * T * PB200_SLAB + sizeof(double) * T * 4));
      for (int j = 0; j < SPCP; ++j)
        tma_bulk_g2s(&tin[stage].amp[j][0][0], (const float*)P.amp + ((size_t)(tile_x * SPCP + j) * P.nsrc_pad + row0) * PB200_SLAB,
                     sizeof(float) * T * PB200_SLAB, &full[stage]);
      tma_bulk_g2s(&tin[stage].geom[0][0], P.geom + row0 * 4, sizeof(double) * T * 4, &full[stage]);
    };
    if (tid == 0) {
      issue(0, fill & 1);
      if (ntiles > 1) issue(1, (fill + 1) & 1);
    }
    // cooperative per-tile stage: tau, r^2, r^8 and the one-channel argument increment of 4 sources of this thread's column
    auto precompute = [&](int tile) {
      const int stage = (fill + tile) & 1;
      mbar_wait(&full[stage], ((fill + tile) >> 1) & 1);
#pragma unroll 2
      for (int j = 0; j < T * BLP / NTHREADS; ++j) {
        const int s = sgrp + (NTHREADS / BLP) * j;
        const double4 g = *reinterpret_cast<const double4*>(&tin[stage].geom[s][0]);
        const double tau = g.x * G.bx + g.y * G.by + g.z * G.bz - G.tau_pc;     // baseline_delay_horizon.py:240, interferometry.py:6332
        tpre[stage].tau[s][scol] = tau;
        tpre[stage].rot[s][scol] = rotation_phasor(2.0 * tau * df);
        tpre[stage].rot8[s][scol] = rotation_phasor(8.0 * tau * df);
        tpre[stage].xd[s][scol] = anchor_arg(tau * df);
      }
    };
    precompute(0);
    __syncthreads();

    bool fresh = true;
    for (int tile = 0; tile < ntiles; ++tile) {
      const int stage = (fill + tile) & 1;
      const TileIn<SPCP>& ti = tin[stage];
      const TilePreP& tp = tpre[stage];
      // unit u = 2 s + h: (source s, baseline half h); the anchors / rotations of unit u + 1 are fetched while unit u runs
      auto fetch = [&](int u, float& x, float& xd) {           // the fp64 part of the anchors of unit u
        const int s = u >> 1, col = lane + 32 * (u & 1);
        x = anchor_arg(tp.tau[s][col] * fk0);
        xd = tp.xd[s][col];
      };
      float x_n, xd_n;
      fetch(0, x_n, xd_n);
#pragma unroll 1
      for (int chunk = 0; chunk < STAGGER; ++chunk) {
        if (chunk == (warp >> 2) % STAGGER && tile + 1 < ntiles) precompute(tile + 1);
#pragma unroll PAIR_UNROLL
        for (int s = chunk * (T / STAGGER); s < (chunk + 1) * (T / STAGGER); ++s) {
          const float4* arow = reinterpret_cast<const float4*>(&ti.amp[sl][s][wcs * KTP]);
          const float4 a0 = arow[0], a1 = arow[1], a2 = arow[2], a3 = arow[3];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const float2 p0 = mufu_phasor(x_n), q0 = mufu_phasor(x_n + xd_n);
            const float2 r = tp.rot[s][lane + 32 * h], r8 = tp.rot8[s][lane + 32 * h];
            fetch(min(2 * s + h + 1, 2 * T - 1), x_n, xd_n);
            const float2 RR = make_float2(r.x, r.x), RI = make_float2(r.y, r.y), NRI = make_float2(-r.y, -r.y);
            const float2 CC = make_float2(2.0f * r.x, 2.0f * r.x);
            const float2 ER = make_float2(r8.x, r8.x), EI = make_float2(r8.y, r8.y), NEI = make_float2(-r8.y, -r8.y);
            float2 PR = make_float2(p0.x, q0.x), PI = make_float2(p0.y, q0.y);          // pair 0 of quarter 0
#pragma unroll
            for (int qd = 0; qd < KTP / 8; ++qd) {
              const float4 a4 = qd == 0 ? a0 : a2, b4 = qd == 0 ? a1 : a3;
              const float2 A0 = make_float2(a4.x, a4.y), A1 = make_float2(a4.z, a4.w), A2 = make_float2(b4.x, b4.y), A3 = make_float2(b4.z, b4.w);
              const int j = 4 * qd;
              float2 QR = PR, QI = PI;
              if (qd + 1 < KTP / 8) {                            // anchor pair of the next quarter: this one rotated by r^8
                const float2 u1 = __fmul2_rn(PR, ER), u2 = __fmul2_rn(PR, EI);
                QR = __ffma2_rn(PI, NEI, u1);
                QI = __ffma2_rn(PI, ER, u2);
              }
              const float2 t1 = __fmul2_rn(PR, RR), t2 = __fmul2_rn(PR, RI);
              acc_re[h][j] = __ffma2_rn(PR, A0, acc_re[h][j]);
              acc_im[h][j] = __ffma2_rn(PI, A0, acc_im[h][j]);
              const float2 P1R = __ffma2_rn(PI, NRI, t1), P1I = __ffma2_rn(PI, RR, t2);      // pair 1 = pair 0 x r^2
              acc_re[h][j + 1] = __ffma2_rn(P1R, A1, acc_re[h][j + 1]);
              acc_im[h][j + 1] = __ffma2_rn(P1I, A1, acc_im[h][j + 1]);
              const float2 P2R = __ffma2_rn(CC, P1R, make_float2(-PR.x, -PR.y)), P2I = __ffma2_rn(CC, P1I, make_float2(-PI.x, -PI.y));
              acc_re[h][j + 2] = __ffma2_rn(P2R, A2, acc_re[h][j + 2]);
              acc_im[h][j + 2] = __ffma2_rn(P2I, A2, acc_im[h][j + 2]);
              const float2 P3R = __ffma2_rn(CC, P2R, make_float2(-P1R.x, -P1R.y)), P3I = __ffma2_rn(CC, P2I, make_float2(-P1I.x, -P1I.y));
              acc_re[h][j + 3] = __ffma2_rn(P3R, A3, acc_re[h][j + 3]);
              acc_im[h][j + 3] = __ffma2_rn(P3I, A3, acc_im[h][j + 3]);
              PR = QR; PI = QI;
            }
          }
        }
      }
      __syncthreads();                                   // tile consumed, next tile's stage visible
      if (tid == 0 && tile + NSTAGE < ntiles) issue(tile + NSTAGE, stage);
      const bool bright = sg.s0 + tile < P.bright_tiles;
      if (bright || ((tile + 1 + (warp >> 2) * (FLUSH_TILES / 4)) % FLUSH_TILES) == 0) {
        flush(sg.slot, fresh, false);
        fresh = false;
      }
    }
    flush(sg.slot, fresh, true);
    fill += (uint32_t)ntiles;
  }
}

// =================================================================================================
// Direct kernel: arbitrary channel frequencies, one accurate sincospi per term (no recurrence)
// =================================================================================================
template <bool TAPER>
__global__ void __launch_bounds__(NTHREADS, 1) k_skyvis_direct(const SkyvisParams P) {
  using S = Shape<1>;
  constexpr int WB = S::WB;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  TileIn<1>* tin = reinterpret_cast<TileIn<1>*>(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + NSTAGE * sizeof(TileIn<1>));
  float* sfreq2 = reinterpret_cast<float*>(smem_raw + NSTAGE * sizeof(TileIn<1>) + 64);
  double* sfreq64 = reinterpret_cast<double*>(smem_raw + NSTAGE * sizeof(TileIn<1>) + 64 + PB200_SLAB * sizeof(float));

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wb = warp % WB, wc = warp / WB;
  if (tid == 0) {
    for (int i = 0; i < NSTAGE; ++i) mbar_init(&full[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  float2 acc_re[KT / 2], acc_im[KT / 2];
#pragma unroll
  for (int k = 0; k < KT / 2; ++k) { acc_re[k] = make_float2(0.f, 0.f); acc_im[k] = make_float2(0.f, 0.f); }
  uint32_t fill = 0;
  SegmentIter seg_iter(P.sc, (int)blockIdx.x);
  Segment sg;
#pragma unroll 1
  while (seg_iter.next(sg)) {
    const int slab = sg.tile % P.sc.gx;
    const int b = (sg.tile / P.sc.gx) * S::BL + wb * 32 + lane;
    const bool valid = b < P.nbl;
    const int ntiles = sg.s1 - sg.s0;
    const Geometry G = load_baseline(P, b, valid);
    __syncthreads();                                   // previous segment done with sfreq*, barriers initialised
    if (tid < PB200_SLAB) {
      const double f = P.freqs[slab * PB200_SLAB + tid];
      sfreq64[tid] = f;
      const float fs = (float)(f * 1e-8);
      sfreq2[tid] = fs * fs;
    }
    __syncthreads();
    const float* amp_slab = (const float*)P.amp + (size_t)slab * P.nsrc_pad * PB200_SLAB;
    auto issue = [&](int i, int stage) {
      const size_t row0 = (size_t)(sg.s0 + i) * T;
      mbar_expect_tx(&full[stage], (uint32_t)sizeof(TileIn<1>));
      tma_bulk_g2s(&tin[stage].amp[0][0][0], amp_slab + row0 * PB200_SLAB, sizeof(float) * T * PB200_SLAB, &full[stage]);
      tma_bulk_g2s(&tin[stage].geom[0][0], P.geom + row0 * 4, sizeof(double) * T * 4, &full[stage]);
    };
    if (tid == 0) {
      issue(0, fill & 1);
      if (ntiles > 1) issue(1, (fill + 1) & 1);
    }
    bool fresh = true;
    for (int tile = 0; tile < ntiles; ++tile) {
      const int stage = (fill + tile) & 1;
      mbar_wait(&full[stage], ((fill + tile) >> 1) & 1);
      const TileIn<1>& ti = tin[stage];
#pragma unroll 1
      for (int s = 0; s < T; ++s) {
        const double4 g = *reinterpret_cast<const double4*>(&ti.geom[s][0]);
        const double tau_g = g.x * G.bx + g.y * G.by + g.z * G.bz;
        const double tau = tau_g - G.tau_pc;
        const float4* arow = reinterpret_cast<const float4*>(&ti.amp[0][s][wc * KT]);
        float kap = 0.f;
        if (TAPER) kap = (float)(g.w * fmax(G.blen2 - tau_g * tau_g, 0.0));
#pragma unroll
        for (int k4 = 0; k4 < KT / 4; ++k4) {
          const float4 a4 = arow[k4];
          const float av[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int k = 4 * k4 + j;
            const float x = (float)(2.0 * frac_turns(tau * sfreq64[wc * KT + k]));   // interferometry.py:6332
            float sn, cs;
            sincospif(x, &sn, &cs);
            float a = av[j];
            if (TAPER) a *= exp2f(-kap * sfreq2[wc * KT + k]);
            if (k & 1) { acc_re[k >> 1].y = fmaf(a, cs, acc_re[k >> 1].y); acc_im[k >> 1].y = fmaf(-a, sn, acc_im[k >> 1].y); }
            else       { acc_re[k >> 1].x = fmaf(a, cs, acc_re[k >> 1].x); acc_im[k >> 1].x = fmaf(-a, sn, acc_im[k >> 1].x); }   // :6340
          }
        }
      }
      __syncthreads();
      if (tid == 0 && tile + NSTAGE < ntiles) issue(tile + NSTAGE, stage);
      if (sg.s0 + tile < P.bright_tiles || ((tile + 1) % FLUSH_TILES) == 0) { flush_acc(P, sg.slot, fresh, acc_re, acc_im); fresh = false; }
    }
    flush_acc(P, sg.slot, fresh, acc_re, acc_im, true);
    fill += (uint32_t)ntiles;
  }
}


// =================================================================================================
// fp64 kernel: same sum, every product and sum in double.  For skies whose visibilities are the small
// residue of a strongly cancelling sum (smooth diffuse emission on resolved baselines:
// |V_b| << sqrt(sum a^2)), where fp32 products cannot reach 1e-5 rms(V_b).  B200's FP64 pipe runs at
// half the FP32 rate (measured 58.8 DFMA lanes/clk/SM), so this costs ~2-3x, not ~60x.
//   * thread = one baseline x 16 channels (32 fp64 accumulators); CTA = 8 channel blocks (one slab) x 2
//     baseline groups; persistent CTAs over the same (whole tiles + stream-K tail) schedule as k_skyvis.
//   * cooperative stage per sub-tile of TS sources, ONE thread per (source, baseline): delay tau, the slab
//     anchor exp(-2 pi i tau f_slab0) and the channel rotation r = exp(-2 pi i tau df) with the fp64 sincospi,
//     then the anchors of all 8 channel blocks by complex powering (r^16 by four squarings, seven complex
//     multiplications) -- 3.5 DFMA per block anchor instead of one sincospi (~35 fp64 operations) per thread
//     and source -- and, for extended sources, the taper weight and its per-channel ratio at every block
//     start by a second-order chain from 6 transcendentals per (source, baseline) instead of 2 exp2 per thread
//     and source.  Both are parked in shared memory (double-buffered; the stage of sub-tile i+1 overlaps the
//     channel loops of sub-tile i).
//   * channel loop: three-term recurrence z_{k+1} = 2 cos(phi) z_k - z_{k-1} for the unit phasor (2 DFMA per
//     term; error growth steps^2/2 ulp = 1e-14 over 16 channels), accumulate 2 DFMA per term; with the taper
//     w_{k+1} = w_k g_k, g_{k+1} = g_k + g_k (h - 1) and a w first: 7 DFMA per term (was 8 with the taper folded
//     into a complex rotation).
// =================================================================================================
#ifndef PB_F64_WB
#define PB_F64_WB 2                           // baseline warps per CTA: 2 = one 512-thread CTA per SM; 1 = 256-thread CTAs, two per SM (measured: no gain)
#endif
#ifndef PB_F64_ABLATE
#define PB_F64_ABLATE 0                       // timing experiments only (wrong results): 1 = no stage after the first, 2 = no channel loop
#endif
constexpr int KT64 = 16;
constexpr int WC64 = PB200_SLAB / KT64;       // 8 channel blocks = one slab
constexpr int WB64 = PB_F64_WB;
constexpr int NW64 = WC64 * WB64;             // warps per CTA
constexpr int NT64 = 32 * NW64;
constexpr int BL64 = 32 * WB64;               // baselines per CTA
constexpr int T64 = 16 * WB64;                // source rows per TMA tile (half the shared memory of T = 32 for the two-CTA shape)
constexpr int CTAS64 = 2 / WB64;              // resident CTAs per SM
static_assert(WB64 == 1 || WB64 == 2, "fp64 CTA shape");
template <bool TAPER> struct Sub64 { static constexpr int TS = TAPER ? 4 : 8; };   // sources per cooperative sub-tile
template <int TS> struct __align__(16) Pre64 {
  double2 anc[TS][WC64][BL64];                // exp(-2 pi i tau f) at the first channel of every 16-channel block
  double2 rot[TS][BL64];                      // exp(-2 pi i tau df)
};
template <int TS> struct __align__(16) Pre64Taper {
  double w[TS][WC64][BL64];                   // taper weight at the first channel of every block
  double g[TS][WC64][BL64];                   // w_{k+1} / w_k there
  double hm1[TS][BL64];                       // g_{k+1} / g_k - 1 (the same for every channel)
};
template <typename AMP> struct __align__(16) TileIn64 {
  AMP amp[T64][PB200_SLAB];
  double geom[T64][4];
};

template <typename AMP, bool TAPER>
__global__ void __launch_bounds__(NT64, CTAS64) k_skyvis_fp64(const SkyvisParams P) {
  constexpr int TS = Sub64<TAPER>::TS;
  constexpr int NSUB = T64 / TS;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  TileIn64<AMP>* tin = reinterpret_cast<TileIn64<AMP>*>(smem_raw);
  Pre64<TS>* pre = reinterpret_cast<Pre64<TS>*>(smem_raw + NSTAGE * sizeof(TileIn64<AMP>));
  Pre64Taper<TS>* pret = reinterpret_cast<Pre64Taper<TS>*>(smem_raw + NSTAGE * sizeof(TileIn64<AMP>) + 2 * sizeof(Pre64<TS>));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + NSTAGE * sizeof(TileIn64<AMP>) + 2 * sizeof(Pre64<TS>) +
                                               (TAPER ? 2 * sizeof(Pre64Taper<TS>) : 0));

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wb = warp % WB64, wc = warp / WB64;
  const int bcol = wb * 32 + lane;            // == tid % BL64: the stage thread of a pair owns the same baseline as in the channel loop
  const double df = P.df;
  if (tid == 0) {
    for (int i = 0; i < NSTAGE; ++i) mbar_init(&full[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  uint32_t fill = 0;
  SegmentIter seg_iter(P.sc, (int)blockIdx.x);
  Segment sg;
#pragma unroll 1
  while (seg_iter.next(sg)) {
    const int slab = sg.tile % P.sc.gx;
    const int b = (sg.tile / P.sc.gx) * BL64 + bcol;
    const bool valid = b < P.nbl;
    const int kbase = slab * PB200_SLAB + wc * KT64;
    const int ntiles = sg.s1 - sg.s0;
    const Geometry G = load_baseline(P, b, valid);
    const double fs0 = P.f0 + (double)(slab * PB200_SLAB) * P.df;        // first channel of the slab
    const double tF0 = fs0 * 1e-8, tdF = df * 1e-8;
    const AMP* amp_slab = (const AMP*)P.amp + (size_t)slab * P.nsrc_pad * PB200_SLAB;
    auto issue = [&](int i, int stage) {
      const size_t row0 = (size_t)(sg.s0 + i) * T64;
      mbar_expect_tx(&full[stage], (uint32_t)sizeof(TileIn64<AMP>));
      tma_bulk_g2s(&tin[stage].amp[0][0], amp_slab + row0 * PB200_SLAB, sizeof(AMP) * T64 * PB200_SLAB, &full[stage]);
      tma_bulk_g2s(&tin[stage].geom[0][0], P.geom + row0 * 4, sizeof(double) * T64 * 4, &full[stage]);
    };
    if (tid == 0) {
      issue(0, fill & 1);
      if (ntiles > 1) issue(1, (fill + 1) & 1);
    }
    // cooperative stage of sub-tile `sub` (sources sub*TS .. sub*TS+TS-1) of source tile `tile`.  Without the taper TS = 8 and
    // thread (wc, bcol) does source wc; with it TS = 4 and the work of a (source, baseline) pair is split over two threads so
    // that all 16 warps carry a similar load before the barrier: warps wc < 4 the phasor part of source wc, warps wc >= 4 the
    // taper part of source wc - 4.
    auto stage_sub = [&](int tile, int sub) {
      const int st = (fill + tile) & 1;
      if (sub == 0) mbar_wait(&full[st], ((fill + tile) >> 1) & 1);
      if (PB_F64_ABLATE == 1 && (tile | sub) != 0) return;
      const int buf = (tile * NSUB + sub) & 1;
      const int ss = TAPER ? (wc & (TS - 1)) : wc;                           // source of the sub-tile this thread works on
      const bool do_phasor = !TAPER || wc < TS, do_taper = TAPER && wc >= TS;
      const double4 g = *reinterpret_cast<const double4*>(&tin[st].geom[sub * TS + ss][0]);
      const double tau_g = g.x * G.bx + g.y * G.by + g.z * G.bz;            // baseline_delay_horizon.py:240
      if (do_phasor) {
        const double tau = tau_g - G.tau_pc;                                // interferometry.py:6332
        double sn, cs;
        sincospi(2.0 * frac_turns(tau * df), &sn, &cs);
        const double2 r = make_double2(cs, -sn);
        pre[buf].rot[ss][bcol] = r;
        sincospi(2.0 * frac_turns(tau * fs0), &sn, &cs);
        double2 a = make_double2(cs, -sn);
        double2 r16 = r;
#pragma unroll
        for (int i = 0; i < 4; ++i) r16 = make_double2(fma(r16.x, r16.x, -r16.y * r16.y), 2.0 * r16.x * r16.y);
        pre[buf].anc[ss][0][bcol] = a;
#pragma unroll
        for (int j = 1; j < WC64; ++j) {
          a = make_double2(fma(-a.y, r16.y, a.x * r16.x), fma(a.y, r16.x, a.x * r16.y));
          pre[buf].anc[ss][j][bcol] = a;
        }
      }
      if (do_taper) {
        // w_k = exp2(-kap F_k^2), F_k = F0 + k dF in units of 1e8 Hz (interferometry.py:6262-6283; g.w folds ln2 d^2 1e16 log2 e,
        // sqrt argument clamped at 0).  Block starts k = 16 j by the chain  w <- w X, X <- X Y;  g <- g Z.
        const double kap = g.w * fmax(G.blen2 - tau_g * tau_g, 0.0);
        double w = exp2(-kap * tF0 * tF0);
        double gk = exp2(-kap * tdF * (2.0 * tF0 + tdF));
        const double hm1 = expm1(-2.0 * kap * tdF * tdF * 0.69314718055994530942);
        pret[buf].hm1[ss][bcol] = hm1;
        // the block-level ratios from powers of h = g_{k+1} / g_k and g_0 instead of three more exp2:
        //   Z = h^16 (g at the next block start), Y = h^256, X = g_0^16 h^120 (w at the next block start)
        const double h = 1.0 + hm1;
        const double h2 = h * h, h4 = h2 * h2, h8 = h4 * h4, Z = h8 * h8;
        const double h32 = Z * Z, h64 = h32 * h32, h128 = h64 * h64, Y = h128 * h128;
        const double g2 = gk * gk, g4 = g2 * g2, g8 = g4 * g4;
        double X = (g8 * g8) * ((h64 * h32) * (Z * h8));
#pragma unroll
        for (int j = 0; j < WC64; ++j) {
          pret[buf].w[ss][j][bcol] = w;
          pret[buf].g[ss][j][bcol] = gk;
          w *= X; X *= Y; gk *= Z;
        }
      }
    };
    double acc_re[KT64], acc_im[KT64];
#pragma unroll
    for (int k = 0; k < KT64; ++k) { acc_re[k] = 0.0; acc_im[k] = 0.0; }
    stage_sub(0, 0);
    __syncthreads();
#pragma unroll 1
    for (int tile = 0; tile < ntiles; ++tile) {
      const int st = (fill + tile) & 1;
      const TileIn64<AMP>& ti = tin[st];
#pragma unroll 1
      for (int sub = 0; sub < NSUB; ++sub) {
        if (sub + 1 < NSUB) stage_sub(tile, sub + 1);
        else if (tile + 1 < ntiles) stage_sub(tile + 1, 0);
        const int buf = (tile * NSUB + sub) & 1;
#pragma unroll 1
        for (int ss = 0; ss < (PB_F64_ABLATE == 2 ? 0 : TS); ++ss) {
          const double2 z0 = pre[buf].anc[ss][wc][bcol];
          const double2 r = pre[buf].rot[ss][bcol];
          const AMP* arow = &ti.amp[sub * TS + ss][wc * KT64];
          // unit phasors: three-term recurrence z_{k+1} = 2 cos(phi) z_k - z_{k-1}
          double pr = z0.x, pi = z0.y;
          double qr = fma(-pi, r.y, pr * r.x), qi = fma(pi, r.x, pr * r.y);      // z_1 = z_0 r
          const double C = 2.0 * r.x;
          if (TAPER) {
            double w = pret[buf].w[ss][wc][bcol], gk = pret[buf].g[ss][wc][bcol];
            const double hm1 = pret[buf].hm1[ss][bcol];
#pragma unroll
            for (int k = 0; k < KT64; k += 2) {
              if (k > 0) {
                pr = fma(C, qr, -pr); pi = fma(C, qi, -pi);
                qr = fma(C, pr, -qr); qi = fma(C, pi, -qi);
              }
              const double a0 = (double)arow[k] * w;
              w *= gk; gk = fma(gk, hm1, gk);
              const double a1 = (double)arow[k + 1] * w;
              w *= gk; gk = fma(gk, hm1, gk);
              acc_re[k] = fma(a0, pr, acc_re[k]); acc_im[k] = fma(a0, pi, acc_im[k]);
              acc_re[k + 1] = fma(a1, qr, acc_re[k + 1]); acc_im[k + 1] = fma(a1, qi, acc_im[k + 1]);
            }
          } else {
#pragma unroll
            for (int k = 0; k < KT64; k += 2) {
              if (k > 0) {
                pr = fma(C, qr, -pr); pi = fma(C, qi, -pi);                       // z_k     (overwrites z_{k-2})
                qr = fma(C, pr, -qr); qi = fma(C, pi, -qi);                       // z_{k+1} (overwrites z_{k-1})
              }
              const double a0 = (double)arow[k], a1 = (double)arow[k + 1];
              acc_re[k] = fma(a0, pr, acc_re[k]); acc_im[k] = fma(a0, pi, acc_im[k]);
              acc_re[k + 1] = fma(a1, qr, acc_re[k + 1]); acc_im[k + 1] = fma(a1, qi, acc_im[k + 1]);
            }
          }
        }
        __syncthreads();                               // sub-tile consumed, the next one's anchors visible
      }
      if (tid == 0 && tile + NSTAGE < ntiles) issue(tile + NSTAGE, st);
    }
    // partial sums of this segment -> scratch slot (k_skyvis_finalize adds head partials and transposes)
    double2* base = P.accum + ((sg.slot * NW64 + warp) * KT64) * 32 + lane;
#pragma unroll
    for (int k = 0; k < KT64; ++k) base[k * 32] = make_double2(acc_re[k], acc_im[k]);
    fill += (uint32_t)ntiles;
    (void)kbase;
  }
}

// geometry staging: [nsrc_pad][4] = (l, m, n, taper coefficient), zero rows for padding
// also: max_s |s - s_pc|^2 (float bits, atomicMax on the unsigned pattern of a non-negative float; rounded up),
// the bound on |tau|/|b/c| that decides which CTA rows may use the lifted rotation
__global__ void k_geom_stage(const double* __restrict__ dircos, const double* __restrict__ fwhm_deg, int nsrc,
                             int nsrc_pad, double* __restrict__ geom, double pcx, double pcy, double pcz,
                             unsigned* __restrict__ smax2_bits) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nsrc_pad) return;
  double4 g = make_double4(0, 0, 0, 0);
  if (s < nsrc) {
    g.x = dircos[3 * (size_t)s]; g.y = dircos[3 * (size_t)s + 1]; g.z = dircos[3 * (size_t)s + 2];
    const double dx = g.x - pcx, dy = g.y - pcy, dz = g.z - pcz;
    atomicMax(smax2_bits, __float_as_uint(__double2float_ru(dx * dx + dy * dy + dz * dz)));
    if (fwhm_deg) {
      double d = 2.0 * sin(0.5 * fwhm_deg[s] * 0.017453292519943295769);      // interferometry.py:6268
      // 1/(2 sigma^2) = ln2 * d^2 (:6270); fold the (f/1e8)^2 scaling and log2(e)
      g.w = log(2.0) * d * d * 1.0e16 * 1.4426950408889634074;
    }
  }
  reinterpret_cast<double4*>(geom)[s] = g;
}

}  // namespace

extern "C" int pb200_skyvis(pb200_ctx* ctx, const double* d_dircos, const void* d_amp, int amp_dtype, int nsrc,
                            const double* d_bl, int nbl, const double* h_pc, const double* h_freqs, int nchan,
                            const double* d_src_fwhm_deg, int nsrc_bright, void* d_vis, long long vis_row_stride, int method,
                            void* stream_) {
  if (!ctx) return PB200_EINVAL;
  if (nsrc < 0 || nbl <= 0 || nchan <= 0 || !d_bl || !h_pc || !h_freqs || !d_vis)
    return pb_fail(ctx, PB200_EINVAL, "pb200_skyvis: bad arguments");
  if (nsrc > 0 && (!d_dircos || !d_amp)) return pb_fail(ctx, PB200_EINVAL, "pb200_skyvis: null source arrays");
  if (amp_dtype != PB200_AMP_F32 && !(amp_dtype == PB200_AMP_F64 && method == PB200_SKYVIS_FP64))
    return pb_fail(ctx, PB200_EINVAL, "pb200_skyvis: an fp64 amplitude table needs method PB200_SKYVIS_FP64");
  if (method < PB200_SKYVIS_AUTO || method > PB200_SKYVIS_RECURRENCE_PAIR)
    return pb_fail(ctx, PB200_EINVAL, "pb200_skyvis: unknown method");
  cudaStream_t stream = (cudaStream_t)stream_;
  PbDeviceGuard guard(ctx->device);
  if (vis_row_stride == 0) vis_row_stride = nchan;
  if (vis_row_stride < nchan) return pb_fail(ctx, PB200_EINVAL, "pb200_skyvis: vis_row_stride smaller than nchan");
  if (nsrc == 0) {                                     // empty ROI: zeros (interferometry.py:6378-6382)
    PB_CUDA(ctx, cudaMemset2DAsync(d_vis, sizeof(double) * 2 * (size_t)vis_row_stride, 0, sizeof(double) * 2 * (size_t)nchan, nbl, stream));
    return PB200_OK;
  }

  // uniform channel grid?  (reference channels are f0 + k*df, run_prisim.py:900)
  const double df = nchan > 1 ? (h_freqs[nchan - 1] - h_freqs[0]) / (double)(nchan - 1) : 0.0;
  const bool uniform = pb200_channels_uniform(h_freqs, nchan) != 0;
  const bool want_rec = (method == PB200_SKYVIS_RECURRENCE || method == PB200_SKYVIS_RECURRENCE_SCALAR || method == PB200_SKYVIS_FP64 ||
                         method == PB200_SKYVIS_RECURRENCE_LIFT || method == PB200_SKYVIS_RECURRENCE_3TERM || method == PB200_SKYVIS_RECURRENCE_3TERM_SCALAR ||
                         method == PB200_SKYVIS_RECURRENCE_QUARTER || method == PB200_SKYVIS_RECURRENCE_PAIR);
  const bool direct = (method == PB200_SKYVIS_DIRECT) || (method == PB200_SKYVIS_AUTO && !uniform);
  if (want_rec && !uniform)
    return pb_fail(ctx, PB200_EINVAL, "pb200_skyvis: recurrence kernel needs uniformly spaced channels");

  const int nslab = (nchan + PB200_SLAB - 1) / PB200_SLAB;
  const int nchan_pad = nslab * PB200_SLAB;
  const int nsrc_pad = pb200_nsrc_pad(nsrc);
  void* geom;
  int rc = pb_scratch(ctx, 2, sizeof(double) * 4 * (size_t)nsrc_pad + 16, &geom);
  if (rc) return rc;
  unsigned* smax2_bits = reinterpret_cast<unsigned*>((double*)geom + 4 * (size_t)nsrc_pad);
  PB_CUDA(ctx, cudaMemsetAsync(smax2_bits, 0, 16, stream));
  const double* dfreq;                                 // padded channel frequencies, resident per ctx (no copy, no sync after the first call)
  rc = pb_channels_device(ctx, h_freqs, nchan, nchan_pad, stream, &dfreq);
  if (rc) return rc;
  k_geom_stage<<<pb_div_up(nsrc_pad, 256), 256, 0, stream>>>(d_dircos, d_src_fwhm_deg, nsrc, nsrc_pad, (double*)geom,
                                                             h_pc[0], h_pc[1], h_pc[2], smax2_bits);
  PB_CHECK_LAUNCH(ctx, "k_geom_stage");

  SkyvisParams P;
  P.amp = d_amp; P.geom = (const double*)geom; P.bl = d_bl; P.freqs = dfreq; P.vis = (double*)d_vis; P.vis_stride = vis_row_stride;
  P.pc[0] = h_pc[0]; P.pc[1] = h_pc[1]; P.pc[2] = h_pc[2];
  P.f0 = h_freqs[0]; P.df = df;
  P.nsrc_pad = nsrc_pad; P.nbl = nbl; P.nchan = nchan; P.nslab = nslab;
  P.smax2_bits = smax2_bits;
  P.bright_tiles = nsrc_bright > 0 ? (nsrc_bright + T - 1) / T : 0;
  // AUTO on a uniform grid without taper takes the quarter-block form (measured on config 2, same process, two A/B rounds:
  // 4.71 vs 4.65 Tterms/s at 4 slabs per CTA, 4.56 vs 4.54 at 2; max error 3.3e-6 vs 3.0e-6, rms 2.7e-7 vs 3.1e-7);
  // PB200_SKYVIS_RECURRENCE asks for the plain rotation explicitly
  int mode = method == PB200_SKYVIS_RECURRENCE_LIFT ? 1 : (method == PB200_SKYVIS_RECURRENCE_3TERM || method == PB200_SKYVIS_RECURRENCE_3TERM_SCALAR ? 2 :
                   ((method == PB200_SKYVIS_RECURRENCE_QUARTER || (method == PB200_SKYVIS_AUTO && !d_src_fwhm_deg)) ? 3 : 0));
  const bool taper = d_src_fwhm_deg != nullptr;

  // CTA shape: fp64 kernel 1 slab x 64 baselines; direct kernel 1 slab x 128 baselines; recurrence kernel the widest channel
  // extent the slab count allows: the (source, baseline) delay / rotation stage is shared by all channel-block warps of a CTA
  // (round 1 measured 4.10 / 4.34 / 4.31 Tterms/s for 1 / 2 / 4 slabs, the round-2 kernel 4.46 / 4.54 / 4.65)
  const bool fp64 = method == PB200_SKYVIS_FP64;
  const bool pair = method == PB200_SKYVIS_RECURRENCE_PAIR;
  if (pair && (taper || nslab % SPCP != 0))
    return pb_fail(ctx, PB200_EINVAL, "pb200_skyvis: the pair form needs point sources and a multiple of 256 channels");
  int spc = (direct || fp64) ? 1 : (pair ? SPCP : (nslab % 4 == 0 ? 4 : (nslab % 2 == 0 ? 2 : 1)));
  if (!direct && !fp64 && !pair && (ctx->skyvis_spc_env == 1 || ctx->skyvis_spc_env == 2 || ctx->skyvis_spc_env == 4)) spc = ctx->skyvis_spc_env;
  if (mode == 3 && spc == 1) mode = 0;       // the r^8 table of the quarter form does not fit beside 128-baseline tiles (258 KB): plain rotation
  P.kt = fp64 ? KT64 : (pair ? KTP : KT);
  P.wc = fp64 ? WC64 : (pair ? WCP : spc * WCS);
  P.wb = fp64 ? WB64 : (pair ? 2 : NWARPS / P.wc);       // pair form: the two baseline halves of a thread are two virtual warps
  const int nw = P.wb * P.wc;                 // warps per CTA
  const int bl_per_cta = 32 * P.wb;
  // persistent schedule (struct Sched): one CTA per SM (every variant needs > half of an SM's shared memory or registers)
  Sched& sc = P.sc;
  sc.gx = pb_div_up(nslab, spc);
  sc.ntile = sc.gx * pb_div_up(nbl, bl_per_cta);
  sc.S = nsrc_pad / (fp64 ? T64 : T);
  const long long units = (long long)sc.ntile * sc.S;
  const int resident = ctx->sm_count * (fp64 ? CTAS64 : 1);
  sc.ncta = (int)(units < resident ? units : resident);
  sc.nwave = sc.ntile / sc.ncta;
  sc.ntail = sc.ntile % sc.ncta;
  const size_t slot_bytes = (size_t)nw * P.kt * 32 * sizeof(double2);
  void* accum;
  rc = pb_scratch(ctx, 4, ((size_t)sc.ntile + sc.ncta) * slot_bytes, &accum);
  if (rc) return rc;
  P.accum = (double2*)accum;
  const dim3 grid(sc.ncta);
#define LAUNCH_N(KERNEL, SMEM, NT)                                                                        \
  do {                                                                                                    \
    PB_CUDA(ctx, cudaFuncSetAttribute(KERNEL, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SMEM))); \
    if ((NT) != NTHREADS) /* two CTAs per SM need the full shared-memory carve-out */                    \
      PB_CUDA(ctx, cudaFuncSetAttribute(KERNEL, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared)); \
    KERNEL<<<grid, (NT), (SMEM), stream>>>(P);                                                            \
  } while (0)
#define LAUNCH(KERNEL, SMEM) LAUNCH_N(KERNEL, SMEM, NTHREADS)
  if (fp64) {
#define SMEM64(AMP, TP) (NSTAGE * sizeof(TileIn64<AMP>) + 2 * sizeof(Pre64<Sub64<TP>::TS>) + (TP ? 2 * sizeof(Pre64Taper<Sub64<TP>::TS>) : 0) + 64)
    if (amp_dtype == PB200_AMP_F64) {
      if (taper) LAUNCH_N((k_skyvis_fp64<double, true>), SMEM64(double, true), NT64);
      else LAUNCH_N((k_skyvis_fp64<double, false>), SMEM64(double, false), NT64);
    } else {
      if (taper) LAUNCH_N((k_skyvis_fp64<float, true>), SMEM64(float, true), NT64);
      else LAUNCH_N((k_skyvis_fp64<float, false>), SMEM64(float, false), NT64);
    }
#undef SMEM64
    PB_CHECK_LAUNCH(ctx, "k_skyvis_fp64");
    k_skyvis_finalize<KT64><<<(unsigned)sc.ntile * nw, 256, 0, stream>>>(P);
    PB_CHECK_LAUNCH(ctx, "k_skyvis_finalize");
    return PB200_OK;
  }
  if (pair) {
    LAUNCH(k_skyvis_pair, NSTAGE * (sizeof(TileIn<SPCP>) + sizeof(TilePreP)) + 64);
    PB_CHECK_LAUNCH(ctx, "k_skyvis_pair");
    k_skyvis_finalize<KTP><<<(unsigned)sc.ntile * nw, 256, 0, stream>>>(P);
    PB_CHECK_LAUNCH(ctx, "k_skyvis_finalize");
    return PB200_OK;
  }
  const bool packed = method != PB200_SKYVIS_RECURRENCE_SCALAR && method != PB200_SKYVIS_RECURRENCE_3TERM_SCALAR;
#define SMEM_REC(SPC) ((PB_Q3_INT && mode == 3 && !taper) ? NSTAGE * (sizeof(TileIn<SPC>) + sizeof(TileQ3<SPC>)) + 64 + SPC * PB200_SLAB * sizeof(float) : \
                       NSTAGE * (sizeof(TileIn<SPC>) + sizeof(TilePre<SPC>)) + 64 + SPC * PB200_SLAB * sizeof(float) +                   \
                       (taper ? NSTAGE * sizeof(TileKap<SPC>) : 0) + (mode == 3 && !taper ? NSTAGE * sizeof(TileRot8<SPC>) : 0))
#define LAUNCH_REC(SPC)                                                                   \
  do {                                                                                    \
    if (taper) {                                                                          \
      if (packed) LAUNCH((k_skyvis<SPC, true, true, 0>), SMEM_REC(SPC));                  \
      else LAUNCH((k_skyvis<SPC, false, true, 0>), SMEM_REC(SPC));                        \
    } else if (mode == 2) {                                                               \
      if (packed) LAUNCH((k_skyvis<SPC, true, false, 2>), SMEM_REC(SPC));                 \
      else LAUNCH((k_skyvis<SPC, false, false, 2>), SMEM_REC(SPC));                       \
    } else if (mode == 3) {                                                               \
      LAUNCH((k_skyvis<SPC, true, false, 3>), SMEM_REC(SPC));                             \
    } else if (mode == 1) {                                                               \
      LAUNCH((k_skyvis<SPC, true, false, 1>), SMEM_REC(SPC));                             \
    } else {                                                                              \
      if (packed) LAUNCH((k_skyvis<SPC, true, false, 0>), SMEM_REC(SPC));                 \
      else LAUNCH((k_skyvis<SPC, false, false, 0>), SMEM_REC(SPC));                       \
    }                                                                                     \
  } while (0)
  if (direct) {
    const size_t smem = NSTAGE * sizeof(TileIn<1>) + 64 + PB200_SLAB * (sizeof(float) + sizeof(double));
    if (taper) LAUNCH(k_skyvis_direct<true>, smem);
    else LAUNCH(k_skyvis_direct<false>, smem);
  } else if (spc == 4) {
    LAUNCH_REC(4);
  } else if (spc == 2) {
    LAUNCH_REC(2);
  } else {
    LAUNCH_REC(1);
  }
#undef LAUNCH_REC
#undef SMEM_REC
#undef LAUNCH
#undef LAUNCH_N
  PB_CHECK_LAUNCH(ctx, "k_skyvis");
  k_skyvis_finalize<KT><<<(unsigned)sc.ntile * NWARPS, 256, 0, stream>>>(P);
  PB_CHECK_LAUNCH(ctx, "k_skyvis_finalize");
  return PB200_OK;
}

// 1 when the channels are f0 + k df to within 1e-4 Hz (1e-4 Hz x 1e-5 s = 1e-9 turn): the recurrence kernels apply
extern "C" int pb200_channels_uniform(const double* h_freqs, int nchan) {
  if (!h_freqs || nchan <= 0) return 0;
  const double df = nchan > 1 ? (h_freqs[nchan - 1] - h_freqs[0]) / (double)(nchan - 1) : 0.0;
  for (int k = 0; k < nchan; ++k)
    if (fabs(h_freqs[k] - (h_freqs[0] + k * df)) > 1e-4) return 0;
  return 1;
}
