/* prisim_b200 -- C-ABI of the B200-native visibility engine (libprisim_b200.so).
 *
 * PRISim (nithyanandan/PRISim v2.2.1) has no FFI: its seam is the Python method surface of
 * prisim/interferometry.py.  Every entry point below names the reference code it replaces
 * (file:line relative to the reference root).  The Python shim in prisim_b200/ binds these with
 * ctypes (prisim_b200/_lib.py); INTEGRATION.md shows the stub a PRISim maintainer would add.
 *
 * Conventions
 *   - plain C types only; all `d_` pointers are DEVICE pointers owned by the caller; `h_` pointers
 *     are HOST pointers read during the call.
 *   - every function returns 0 on success or a negative PB200_E* code; pb200_last_error(ctx)
 *     gives the message.  Nothing throws or aborts.
 *   - work is enqueued on the CUDA stream passed as `stream` (a cudaStream_t cast to void*;
 *     NULL = legacy default stream); functions do not synchronise unless stated.  Two one-off
 *     exceptions: the first call that sees a new channel grid (h_freqs) uploads it and waits for
 *     that copy (the ctx keeps the last four grids resident, later calls enqueue nothing for it),
 *     and a call that needs more ctx scratch than any before it reallocates (cudaFree/cudaMalloc
 *     synchronise the device).  Steady-state calls with the same shapes are fully asynchronous.
 *   - every entry point runs on the ctx's device and restores the caller's current device.
 *   - one pb200_ctx per device per host thread; a ctx is not re-entrant.
 *   - visibilities are [nbl, nchan] row-major (channel fastest), complex128 = interleaved
 *     (re, im) doubles.
 */
#ifndef PRISIM_B200_H
#define PRISIM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PB200_VERSION 200

#define PB200_OK            0
#define PB200_EINVAL       -1   /* bad argument */
#define PB200_ECUDA        -2   /* CUDA runtime error (message has the CUDA string) */
#define PB200_ENOMEM       -3
#define PB200_EUNSUPPORTED -4

typedef struct pb200_ctx pb200_ctx;

int pb200_version(void);
int pb200_ctx_create(pb200_ctx** out, int device);
void pb200_ctx_destroy(pb200_ctx* ctx);
const char* pb200_last_error(const pb200_ctx* ctx);
/* developer / test knobs: "skyvis_spc" = 0 (automatic) | 1 | 2 | 4 slabs of 128 channels per CTA of the phase-sum kernel;
 * "dt_force_r8" != 0: 1024-point delay transforms through the radix-8 kernel instead of the warp-per-row one (A/B).        */
int pb200_ctx_set_option(pb200_ctx* ctx, const char* name, long long value);
/* number of kernels this ctx has launched since creation (bench.py's "gpu_launches") */
long long pb200_launch_count(const pb200_ctx* ctx);

/* ---------------------------------------------------------------------------------------------
 * Sky coordinates -> direction cosines, horizon / ROI cull, order-preserving compaction.
 * Replaces interferometry.py:6174-6219 (hadec2altaz :6177, altitude cut :6216, selection :6219)
 * and the alt>=0 pre-selection of scripts/run_prisim.py:1872-1876.
 *
 *   coords: PB200_SKY_ALTAZ  d_skypos = [nsrc0,2] (alt, az) degrees
 *           PB200_SKY_HADEC  d_skypos = [nsrc0,2] (HA, Dec) degrees, needs latitude_deg
 *           PB200_SKY_DIRCOS d_skypos = [nsrc0,3] (l, m, n) East-North-Up direction cosines
 *   keep source s  iff  alt_s >= 90 - roi_radius_deg   (roi_center='zenith', reference default
 *   roi_radius = 90), or, when h_roi_center (host, [3] ENU direction cosines) is not NULL, iff the
 *   source lies within roi_radius_deg of that direction (roi_center='pointing_center', :6213)
 *   outputs (capacity nsrc0): d_dircos [nsrc,3] fp64, d_index [nsrc] int32 (ascending, = the
 *   reference's m2 / obs_catalog_indices), *h_nsrc (host; the call synchronises the stream).
 */
#define PB200_SKY_ALTAZ  0
#define PB200_SKY_HADEC  1
#define PB200_SKY_DIRCOS 2
int pb200_sky_cull(pb200_ctx* ctx, const double* d_skypos, int nsrc0, int coords, double latitude_deg,
                   double roi_radius_deg, const double* h_roi_center, double* d_dircos, int32_t* d_index,
                   int* h_nsrc, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Source spectrum x primary beam -> amplitude table.
 * Replaces interferometry.py:6249 (skymodel.generate_spectrum), :6251-6254 (beam, pbfluxes) and
 * the beam-table step of ROI_parameters.append_settings (:4583-4615); beam maths from
 * primary_beams.py:9-441 (wrapper), :517-625 (Airy), :629-730 (Gaussian), :975-1235 (dipole),
 * :812-971 (ground plane), :1239-1478 (analytic n1 x n2 array), :1482-1754 (phased array).
 */
#define PB200_BEAM_DELTA    0   /* primary_beams.py:355-359 */
#define PB200_BEAM_AIRY     1   /* 'hera'/'hirax' presets (:239-247) and shape 'dish' (:369-373) */
#define PB200_BEAM_GAUSSIAN 2   /* shape 'gaussian' (:374-377) */
#define PB200_BEAM_DIPOLE   3   /* 'mwa_dipole'/'paper' presets (:320-349), shape 'dipole' (:360-368) */
#define PB200_BEAM_TABLE    4   /* caller-supplied pbeam[nsrc,nchan] (roi_info['pbeam'], interferometry.py:6189-6202) */
#define PB200_BEAM_LOGTABLE 5   /* d_pbeam holds log10 beam[nsrc,nchan] from pb200_healpix_beam; pb = 10^(log - d_logmax[f])
                                   (scripts/run_prisim.py:1904-1908) */

#define PB200_ARRAY_NONE     0
#define PB200_ARRAY_ANALYTIC 1  /* isotropic_radiators_array_field_pattern, 'mwa' preset without pointing_info (:273-285) */
#define PB200_ARRAY_ELEMENTS 2  /* array_field_pattern with element positions, delays and gains (:287-316, :385-416) */

#define PB200_DIPOLE_GENERAL  0 /* primary_beams.py:1224-1225 (wrapper default, :11-12) */
#define PB200_DIPOLE_SHORT    1 /* :1216-1218 */
#define PB200_DIPOLE_HALFWAVE 2 /* :1220-1222 */

typedef struct pb200_beam_desc {
  int32_t element;            /* PB200_BEAM_* */
  int32_t array_mode;         /* PB200_ARRAY_* : multiplies the element FIELD pattern */
  int32_t dipole_mode;        /* PB200_DIPOLE_* */
  int32_t achromatic;         /* !=0: evaluate at ref_freq_hz and broadcast (interferometry.py:4583-4588) */
  double  size;               /* dish diameter / Gaussian FWHM aperture / dipole length [m] */
  double  pointing[3];        /* element pointing centre, ENU direction cosines (Airy/Gaussian) */
  double  orientation[3];     /* dipole axis, ENU direction cosines (:1201 default east) */
  double  groundplane;        /* height [m] above ground plane; <= 0: none (:418-439) */
  double  ground_scale;       /* modifier['scale'], 0 = no modifier (:955-958) */
  double  ground_max;         /* modifier['max'], <= 0 = no clip (:959-960) */
  double  ref_freq_hz;        /* used when achromatic */
  /* analytic array (:1430-1478) */
  int32_t nax1, nax2;
  double  sep1, sep2, east2ax1_deg;
  double  array_pointing[3];  /* ENU direction cosines of the array pointing centre */
  /* element array (:1595-1754): all arrays are DEVICE pointers of n_elements entries;
     d_delays/d_gains are [n_elements, nrand] (delay jitter / gain jitter realisations already
     applied by the caller, :1655/:1666); pattern = mean over nrand of |E * F|^2 (:317, :416) */
  int32_t n_elements, nrand;
  const double* d_element_locs;   /* [n_elements,3] metres ENU */
  const double* d_delays;         /* seconds */
  const double* d_gains;
  const double* d_logmax;         /* [nchan] per-channel normalisation for PB200_BEAM_LOGTABLE (device) */
} pb200_beam_desc;

/* power-law spectrum (astroutils SkyModel 'func'/'power-law', parameters as built at
 * scripts/run_prisim.py:1629-1636): S = offset + scale (f/f_ref)^index.  Arrays are indexed by
 * the ORIGINAL catalogue index (gathered through d_index).  d_flux_offset may be NULL.
 * d_spectrum != NULL selects a tabulated spectrum [nsrc0, nchan] (fp64) instead.             */
typedef struct pb200_spectrum_desc {
  const double* d_flux_scale;
  const double* d_index;
  const double* d_freq_ref;
  const double* d_flux_offset;
  const double* d_spectrum;
} pb200_spectrum_desc;

/* Amplitude-table layout (consumed by pb200_skyvis): fp32, channel slabs of PB200_SLAB channels,
 * each slab a dense [nsrc_pad, PB200_SLAB] matrix; nsrc_pad = nsrc rounded up to PB200_SRC_TILE,
 * padding rows/channels are zero.  pb200_amp_bytes gives the buffer size.                     */
#define PB200_SLAB     128
#define PB200_SRC_TILE 32
#define PB200_AMP_F32  0      /* default: fp32 table (7.6e-8 effect on point-source skies) */
#define PB200_AMP_F64  1      /* fp64 table for PB200_SKYVIS_FP64 on strongly cancelling (diffuse) skies */
size_t pb200_amp_bytes(int nsrc, int nchan);   /* fp32 table; an fp64 table is twice this */
int pb200_nsrc_pad(int nsrc);

/* d_dircos/d_index from pb200_sky_cull (nsrc entries); h_freqs [nchan] Hz (host);
 * d_pbeam: only for PB200_BEAM_TABLE, fp64 [nsrc, nchan] rows aligned with d_index order.
 * Output: d_amp (layout above).                                                               */
int pb200_amp_table(pb200_ctx* ctx, const double* d_dircos, const int32_t* d_index, int nsrc,
                    const pb200_spectrum_desc* spec, const pb200_beam_desc* beam, const double* d_pbeam,
                    const double* h_freqs, int nchan, int amp_dtype, void* d_amp, void* stream);

/* d_amp_out = d_amp_in with source row s multiplied by d_scale[s * scale_stride] (fp64, device); in place allowed.
 * With d_scale = one column of d_dircos (stride 3) the scaled table fed to pb200_skyvis gives one component of the
 * visibility gradient w.r.t. the baseline vector, gradient_mode='baseline' (interferometry.py:6312-6343).       */
int pb200_amp_scale(pb200_ctx* ctx, const void* d_amp_in, int amp_dtype, int nsrc, int nchan, const double* d_scale,
                    int scale_stride, void* d_amp_out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * The phase sum.  Replaces interferometry.py:6155-6165 (phase-centre delays), :6255 +
 * baseline_delay_horizon.py:240 (geometric delays), :6258-6283 (extended-source taper) and
 * :6332-6340 / :6348-6376 (phase matrix, exp, sum over sources):
 *     V[b,f] = sum_s amp[s,f] w[s,b,f] exp(-2 pi i f ((s_s - s_pc) . b) / c)
 *   d_dircos   [nsrc,3] fp64 source direction cosines (ENU)
 *   d_amp      amplitude table from pb200_amp_table; amp_dtype = PB200_AMP_F32, or PB200_AMP_F64
 *              (accepted by method PB200_SKYVIS_FP64 only)
 *   d_bl       [nbl,3] fp64 baselines, local ENU metres
 *   h_pc       [3] phase-centre direction cosines (host)
 *   h_freqs    [nchan] Hz (host).  Uniformly spaced channels take the recurrence kernel;
 *              anything else takes the direct (sincospi per term) kernel.
 *   d_src_fwhm_deg  NULL, or [nsrc] sqrt(major*minor) FWHM in degrees (:6267) -> taper on
 *   nsrc_bright  0, or: the caller has ordered the sources (d_dircos, amplitude-table rows, d_src_fwhm_deg) so
 *              that the first nsrc_bright are the brightest; their fp32 partial sums are then moved to the fp64
 *              running sums after every 32-source tile instead of every 512 sources, so that the handful of
 *              sources that carry most of sum a^2 do not set the rounding unit of everybody else's additions.
 *              The sum itself does not depend on the order.
 *   d_vis      [nbl,nchan] complex128, overwritten; consecutive baseline rows are vis_row_stride complex elements
 *              apart (0 = nchan, i.e. dense).  A multiple of nchan addresses every n-th row of a larger array: the
 *              interleaved baseline shard of one rank inside the writing rank's buffer (sharding.py).
 *   method     PB200_SKYVIS_AUTO | _RECURRENCE | _DIRECT | _RECURRENCE_SCALAR | _FP64 | _RECURRENCE_LIFT | _RECURRENCE_3TERM | _RECURRENCE_QUARTER | _RECURRENCE_PAIR
 */
#define PB200_SKYVIS_AUTO       0
#define PB200_SKYVIS_RECURRENCE 1
#define PB200_SKYVIS_DIRECT     2
#define PB200_SKYVIS_RECURRENCE_SCALAR 3   /* same algorithm with scalar FFMA instead of packed FFMA2 (A/B measurement) */
#define PB200_SKYVIS_FP64       4           /* recurrence with every product and sum in fp64 (strongly cancelling skies) */
#define PB200_SKYVIS_RECURRENCE_LIFT 5     /* packed recurrence with the 3-op lifted (shear) rotation on CTA rows of short baselines
                                              (A/B measurement: fewer FFMA2 and more accurate, but slower -- the loop is bound by
                                              register-file operand bandwidth, not by the FMA pipe) */
#define PB200_SKYVIS_RECURRENCE_3TERM 6    /* packed three-term recurrence Z_{j+1} = 2cos(2phi) Z_j - Z_{j-1} in 16-channel half blocks (A/B) */
#define PB200_SKYVIS_RECURRENCE_3TERM_SCALAR 7   /* the same with scalar FFMA (A/B) */
#define PB200_SKYVIS_RECURRENCE_QUARTER 8   /* packed "quarter blocks": one MUFU anchor pair per source, the other three 8-channel quarters anchored by
                                              exact r^8 rotations, inside a quarter one r^2 rotation + two three-term steps: 76 instead of 92 packed
                                              instructions per source (the taper keeps the plain rotation) */
#define PB200_SKYVIS_RECURRENCE_PAIR 9      /* the quarter-block arithmetic with one thread owning two baselines x 16 channels, so that every amplitude
                                              load feeds two baselines (A/B; point sources, nchan a multiple of 256) */
int pb200_skyvis(pb200_ctx* ctx, const double* d_dircos, const void* d_amp, int amp_dtype, int nsrc,
                 const double* d_bl, int nbl, const double* h_pc, const double* h_freqs, int nchan,
                 const double* d_src_fwhm_deg, int nsrc_bright, void* d_vis, long long vis_row_stride, int method,
                 void* stream);
/* 1 when h_freqs is f0 + k df to within 1e-4 Hz -- the test pb200_skyvis applies before it takes a recurrence
 * (or the fp64) kernel; the host shim uses the same test to decide whether precision control applies.        */
int pb200_channels_uniform(const double* h_freqs, int nchan);

/* ---------------------------------------------------------------------------------------------
 * Thermal noise + gains.  Replaces interferometry.py:6676-6693 (generate_noise) and :6707-6722
 * (add_noise):  rms = 2 k Tsys / (A_eff eff_Q sqrt(df t_acc)) / Jy  (flux_unit_k != 0: the K
 * form :6689);  noise = rms/sqrt2 (N + iN);  vis = gains*skyvis + noise.
 * One snapshot, logical shape [nbl,nchan].  d_tsys / d_aeff / d_effq are fp64 with element
 * strides (row, col) given in `strides[6]` = {tsys_row, tsys_col, aeff_row, aeff_col, effq_row,
 * effq_col}; a zero stride broadcasts (a [nchan] Tsys is {0,1}, a scalar is {0,0}).
 * Normal deviates come from Philox4x32-10 keyed by seed with counter (snapshot*nbl_total + bl_offset + b*bl_step) *
 * ceil(nchan/2) + (f mod ceil(nchan/2)) -- words 0-1 serve channel f < ceil(nchan/2), words 2-3 channel f +
 * ceil(nchan/2) -- so a result does not depend on how baselines are sharded across GPUs (bl_step = 1: contiguous
 * block starting at bl_offset; bl_step = world size: interleaved shard, local row b is global baseline bl_offset + b*bl_step).  Box-Muller in fp32 (24-bit
 * deviates), rms and scaling in fp64.  d_gains (complex128 [nbl,nchan]) may be NULL (unity).  Any of d_rms /
 * d_noise / d_vis may be NULL to skip that output.
 * add_only != 0: d_noise is an INPUT and only d_vis = gains*skyvis + noise is written (:6722).
 */
int pb200_noise(pb200_ctx* ctx, const void* d_skyvis, const double* d_tsys, const double* d_aeff,
                const double* d_effq, const long long* strides, const void* d_gains, int nbl, int nchan,
                double df, double t_acc, int flux_unit_k, uint64_t seed, int snapshot, int bl_offset, int bl_step,
                int nbl_total, int add_only, double* d_rms, void* d_noise, void* d_vis, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Windowed delay transform.  Replaces interferometry.py:8114-8134 and delay_spectrum.py:1305-1327:
 *   X = fftshift(ifft(pad_right(x * bp * wts, npad))) * (nchan+npad) * df, then linear-interp
 *   decimation by (1+pad) when `downsample`.
 *   d_x [nrows,nchan] complex128 (NULL: transform bp*wts only -> lag_kernel), d_bp / d_wts
 *   [nrows,nchan] fp64 or NULL (=1); wts_row_stride/bp_row_stride = 0 broadcasts one row.
 *   d_out [nrows, nout] complex128 with nout = pb200_delay_nout(nchan, pad, downsample).
 */
int pb200_delay_nout(int nchan, double pad, int downsample);
int pb200_delay_transform(pb200_ctx* ctx, const void* d_x, const double* d_bp, long long bp_row_stride,
                          const double* d_wts, long long wts_row_stride, int nrows, int nchan, double df,
                          double pad, int downsample, void* d_out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Gridded (HEALPix, RING) primary beam.  Replaces the external-beam step of scripts/run_prisim.py:1897-1908
 * (OPS.healpix_interp_along_axis = healpy bilinear interpolation at (theta, phi) = (pi/2 - alt, az) + spectral
 * interpolation, then renormalisation by the per-channel maximum over the ROI).  The two interpolations are
 * linear and commute: d_map is the log10 power beam ALREADY resampled to the observing channels,
 * [npix = 12 nside^2][nchan] (channel fastest), fp32 or fp64 (map_dtype = PB200_AMP_F32/F64).
 *   d_dircos [nsrc,3] culled ENU direction cosines (beam frame: pole = zenith, phi = azimuth)
 *   outputs: d_logbeam [nsrc,nchan] fp64 interpolated log10 beam; d_colmax [nchan] = max(0, max_s logbeam)
 * Feed both to pb200_amp_table with element PB200_BEAM_LOGTABLE (d_pbeam = d_logbeam, d_logmax = d_colmax).
 */
int pb200_healpix_beam(pb200_ctx* ctx, const void* d_map, int map_dtype, int nside, const double* d_dircos, int nsrc,
                       int nchan, double* d_logbeam, double* d_colmax, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Re-phasing to a new phase centre.  Replaces the elementwise product of
 * InterferometerArray.phase_centering, interferometry.py:7869-7881 (called via rotate_visibilities,
 * scripts/run_prisim.py:2282):  V[b,f] *= exp(-2 pi i f b.(s_old - s_new)/c), in place.
 *   d_vis [nbl,nchan] complex128 (in/out), d_bl [nbl,3] metres in the frame of the direction
 *   cosines, h_dpos [3] = s_old - s_new (host), h_freqs [nchan] Hz (host).
 */
int pb200_phase_rotate(pb200_ctx* ctx, void* d_vis, const double* d_bl, int nbl, const double* h_dpos,
                       const double* h_freqs, int nchan, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Multi-GPU gather through peer memory.  Replaces the exchange through _part_N.hdf5 files and the rank-0
 * concatenation of scripts/run_prisim.py:1995, :2233-2276.  The writing rank allocates the full output buffer
 * (pb200_device_alloc) and exports it (pb200_peer_export, 64-byte CUDA IPC handle); every other rank maps it
 * (pb200_peer_open; peer access over NVLink is enabled lazily) and passes its slice as d_vis to pb200_skyvis, whose
 * epilogue then stores the finished visibilities directly into the writing rank's HBM -- no separate collective.
 */
int pb200_device_alloc(pb200_ctx* ctx, size_t bytes, void** d_out);
int pb200_device_free(pb200_ctx* ctx, void* d_ptr);
int pb200_peer_export(pb200_ctx* ctx, void* d_ptr, void* h_handle64);
int pb200_peer_open(pb200_ctx* ctx, const void* h_handle64, void** d_out);
int pb200_peer_close(pb200_ctx* ctx, void* d_ptr);

/* ---------------------------------------------------------------------------------------------
 * Issue-rate microbenchmark (FP32 FMA lanes / clk / SM etc.) used for the measured roofline
 * denominator.  Fills out[0..n) with: [0] FFMA lane-ops/s, [1] FFMA
 * lane-ops/clk/SM, [2] MUFU lane-ops/s, [3] DFMA lane-ops/s, [4] SM clock (Hz) seen.  Synchronises.
 */
int pb200_microbench(pb200_ctx* ctx, double* out, int n);

#ifdef __cplusplus
}
#endif
#endif /* PRISIM_B200_H */
