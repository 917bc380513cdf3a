// Source spectrum x primary beam -> fp32 amplitude table (fused, fp64 arithmetic, one pass).
// Replaces interferometry.py:6249-6254 of the reference (generate_spectrum, primary_beam_generator,
// pbfluxes = pb * fluxes) and the beam-table step of ROI_parameters.append_settings (:4583-4615).
// Pattern formulas restate prisim/primary_beams.py: wrapper :224-441, Airy :609-623, Gaussian
// :716-728, dipole :1205-1235, ground plane :950-966, analytic array :1451-1476, element array
// :1730-1746 (evaluated here in fp64; the reference uses fp32 there).
//
// Mapping: one CTA per (padded) source row, threads stride over channels, so per-source geometry
// is computed once per thread and every store is a coalesced 512-byte run of one slab row.
// Bound: fp64 transcendental throughput (j1 / pow / sincospi), ~1e2 flops per (source, channel);
// algorithmic bytes = 4 B written per (source, channel).  It is O(nsrc*nchan) against the phase
// sum's O(nsrc*nbl*nchan) and stays < 1 % of a snapshot.
#include "common.cuh"

namespace {

constexpr int AMP_THREADS = 128;
constexpr double TWO_PI = 6.283185307179586476925;

struct AmpParams {
  pb200_beam_desc beam;
  pb200_spectrum_desc spec;
  const double* dircos;
  const int32_t* index;
  const double* pbeam;
  const double* freqs;   // device copy [nchan]
  void* amp;
  int nsrc, nsrc_pad, nchan, nslab;
};

__device__ __forceinline__ double airy_field(double sinx, double k, double diameter) {
  // primary_beams.py:614, :618: 2 J1(k D/2 sin x)/(k D/2 sin x), normalised by its value at x = tol
  const double tol = 1e-10;
  double arg = k * 0.5 * diameter * sinx;
  double arg0 = k * 0.5 * diameter * sin(tol);
  return (2.0 * j1(arg) / arg) / (2.0 * j1(arg0) / arg0);
}

template <typename OUT>
__global__ void __launch_bounds__(AMP_THREADS) k_amp_table(const AmpParams P) {
  const int s = blockIdx.x;
  const bool live = s < P.nsrc;
  // ---- per-source geometry (fp64, computed redundantly by every thread of the CTA) ----
  double l = 0, m = 0, n = 1;
  int cat = 0;
  if (live) {
    l = P.dircos[3 * (size_t)s]; m = P.dircos[3 * (size_t)s + 1]; n = P.dircos[3 * (size_t)s + 2];
    cat = P.index ? P.index[s] : s;
  }
  const pb200_beam_desc& B = P.beam;
  // Airy / Gaussian: angle from the pointing centre (primary_beams.py:605-607)
  double sinx = 0.0;
  bool zero_elem = false;
  if (B.element == PB200_BEAM_AIRY || B.element == PB200_BEAM_GAUSSIAN) {
    double dot = l * B.pointing[0] + m * B.pointing[1] + n * B.pointing[2];
    double cx = m * B.pointing[2] - n * B.pointing[1];
    double cy = n * B.pointing[0] - l * B.pointing[2];
    double cz = l * B.pointing[1] - m * B.pointing[0];
    sinx = sqrt(cx * cx + cy * cy + cz * cz);
    zero_elem = (dot <= 0.0) || (n <= 0.0);                        // x >= pi/2 or alt <= 0 (:607)
    if (B.element == PB200_BEAM_AIRY) sinx = fmax(sinx, sin(1e-10));   // small_angle_tol (:611-612)
  }
  // dipole: angle between the dipole axis and the source (:1207-1211)
  double dcos = 0.0, dsin = 1.0;
  bool dip_zero = false;
  if (B.element == PB200_BEAM_DIPOLE) {
    dcos = l * B.orientation[0] + m * B.orientation[1] + n * B.orientation[2];
    double ang = acos(fmin(1.0, fmax(-1.0, dcos)));
    dsin = sin(ang);
    dip_zero = fabs(fabs(dcos) - 1.0) < 1e-10;
  }
  // analytic array: rotated relative direction cosines (:1441-1451)
  double rel1 = 0.0, rel2 = 0.0;
  if (B.array_mode == PB200_ARRAY_ANALYTIC) {
    double se, ce;
    sincos(B.east2ax1_deg * 0.017453292519943295769, &se, &ce);
    double l1 = l * ce + m * se, m1 = -l * se + m * ce;
    double pl = B.array_pointing[0] * ce + B.array_pointing[1] * se;
    double pm = -B.array_pointing[0] * se + B.array_pointing[1] * ce;
    rel1 = l1 - pl; rel2 = m1 - pm;
  }
  // ground plane modifier (:955-963)
  double gmod = 1.0;
  if (B.groundplane > 0.0 && B.ground_scale != 0.0) {
    gmod = B.ground_scale / sqrt(fabs(n));
    if (B.ground_max > 0.0) gmod = fmin(fmax(gmod, 0.0), B.ground_max);
  }
  double fscale = 0, findex = 0, fref = 1, foff = 0;
  if (live && !P.spec.d_spectrum) {
    fscale = P.spec.d_flux_scale[cat];
    findex = P.spec.d_index[cat];
    fref = P.spec.d_freq_ref[cat];
    foff = P.spec.d_flux_offset ? P.spec.d_flux_offset[cat] : 0.0;
  }

  const int nchan_pad = P.nslab * PB200_SLAB;
  for (int f = threadIdx.x; f < nchan_pad; f += AMP_THREADS) {
    OUT out = (OUT)0;
    if (live && f < P.nchan) {
      const double freq = P.freqs[f];
      // ---- spectrum (run_prisim.py:1629-1636 parameters; astroutils power law) ----
      double flux;
      if (P.spec.d_spectrum) flux = P.spec.d_spectrum[(size_t)cat * P.nchan + f];
      else flux = foff + fscale * pow(freq / fref, findex);
      // ---- beam ----
      const double fb = B.achromatic ? B.ref_freq_hz : freq;
      const double k = TWO_PI * fb / PB_SPEED_OF_LIGHT;
      double pb;
      if (B.element == PB200_BEAM_TABLE) {
        pb = P.pbeam[(size_t)s * P.nchan + f];
      } else if (B.element == PB200_BEAM_LOGTABLE) {
        pb = exp10(P.pbeam[(size_t)s * P.nchan + f] - B.d_logmax[f]);        // run_prisim.py:1906-1908
      } else {
        double ep = 1.0;                                            // delta (:355-359)
        if (B.element == PB200_BEAM_AIRY) {
          ep = zero_elem ? 0.0 : airy_field(sinx, k, B.size);       // (:614-616)
        } else if (B.element == PB200_BEAM_GAUSSIAN) {
          double sigma_aprtr = B.size / (2.0 * sqrt(2.0 * log(2.0))) / (PB_SPEED_OF_LIGHT / fb);   // (:717)
          double sigma_dircos = 1.0 / (TWO_PI * sigma_aprtr);       // (:721)
          double q = sinx / sigma_dircos;
          ep = zero_elem ? 0.0 : exp(-0.5 * q * q);                 // (:723-725)
        } else if (B.element == PB200_BEAM_DIPOLE) {
          double kh = k * 0.5 * B.size;                             // (:1205-1206)
          if (B.dipole_mode == PB200_DIPOLE_SHORT) {
            ep = dsin;                                              // (:1217)
          } else {
            double maxp = 1.0;
            if (B.dipole_mode == PB200_DIPOLE_HALFWAVE) {
              ep = cos(0.5 * 3.14159265358979323846 * dcos) / dsin; // (:1221)
            } else {
              maxp = 1.0 - cos(kh);                                 // (:1224)
              ep = (cos(kh * dcos) - cos(kh)) / dsin;               // (:1225)
            }
            if (dip_zero) ep = kh * sin(kh * dcos) * (dsin / dcos); // (:1228) L'Hospital form
            ep /= maxp;                                             // (:1235)
          }
        }
        // ---- array factor (field), power = mean_r |ep * F|^2  (:317, :416) ----
        double pw;
        if (B.array_mode == PB200_ARRAY_ANALYTIC) {
          double lam = PB_SPEED_OF_LIGHT / fb;
          double phi = TWO_PI * B.sep1 * rel1 / lam, psi = TWO_PI * B.sep2 * rel2 / lam;   // (:1460-1461)
          double n1 = (double)B.nax1;
          double t1 = (fabs(phi) < 1e-10) ? cos(0.5 * n1 * phi) / cos(0.5 * phi)
                                          : sin(0.5 * n1 * phi) / sin(0.5 * phi) / n1;     // (:1467-1469)
          double t2 = (fabs(psi) < 1e-10) ? cos(0.5 * n1 * psi) / cos(0.5 * psi)
                                          : sin(0.5 * n1 * psi) / sin(0.5 * psi) / n1;     // (:1471-1473, nax1 sic)
          double fld = ep * t1 * t2;
          pw = fld * fld;
        } else if (B.array_mode == PB200_ARRAY_ELEMENTS) {
          pw = 0.0;
          for (int r = 0; r < B.nrand; ++r) {
            double fr = 0.0, fi = 0.0;
            for (int e = 0; e < B.n_elements; ++e) {
              const double* loc = B.d_element_locs + 3 * e;
              double gd = -(loc[0] * l + loc[1] * m + loc[2] * n) / PB_SPEED_OF_LIGHT;     // (:1730)
              double turns = fb * (gd + B.d_delays[e * B.nrand + r]);                      // (:1737-1742)
              turns -= rint(turns);
              double sn, cs;
              sincospi(2.0 * turns, &sn, &cs);
              double g = B.d_gains ? B.d_gains[e * B.nrand + r] : 1.0;
              fr += g * cs; fi += g * sn;
            }
            fr /= B.n_elements; fi /= B.n_elements;                                        // (:1743)
            pw += (ep * fr) * (ep * fr) + (ep * fi) * (ep * fi);
          }
          pw /= B.nrand;
        } else {
          pw = ep * ep;
        }
        // ---- ground plane (:418-439, :953-966) ----
        if (B.groundplane > 0.0) {
          double gp = (2.0 * sin(k * B.groundplane * n) * gmod) / (2.0 * sin(k * B.groundplane));
          pw *= gp * gp;
        }
        pb = pw;
      }
      out = (OUT)(pb * flux);                                       // pbfluxes (:6254)
    }
    const int slab = f / PB200_SLAB, c = f - slab * PB200_SLAB;
    reinterpret_cast<OUT*>(P.amp)[((size_t)slab * P.nsrc_pad + s) * PB200_SLAB + c] = out;
  }
}

}  // namespace

// out[slab][s][c] = in[slab][s][c] * scale[s * stride]: the amplitude table of one component of the visibility
// gradient w.r.t. the baseline vector (interferometry.py:6343: dircos[s, i] * pbfluxes[s, f])
template <typename T>
__global__ void k_amp_scale(const T* __restrict__ in, const double* __restrict__ scale, int stride, int nsrc,
                            int nsrc_pad, int nslab, T* __restrict__ out) {
  const size_t n4 = (size_t)nslab * nsrc_pad * (PB200_SLAB / 4);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const int s = (int)((i / (PB200_SLAB / 4)) % nsrc_pad);
    const double f = s < nsrc ? scale[(size_t)s * stride] : 0.0;
    if (sizeof(T) == 4) {
      float4 v = reinterpret_cast<const float4*>(in)[i];
      v.x = (float)(v.x * f); v.y = (float)(v.y * f); v.z = (float)(v.z * f); v.w = (float)(v.w * f);
      reinterpret_cast<float4*>(out)[i] = v;
    } else {
      double4 v = reinterpret_cast<const double4*>(in)[i];
      v.x *= f; v.y *= f; v.z *= f; v.w *= f;
      reinterpret_cast<double4*>(out)[i] = v;
    }
  }
}

extern "C" {

int pb200_amp_scale(pb200_ctx* ctx, const void* d_amp_in, int amp_dtype, int nsrc, int nchan, const double* d_scale,
                    int scale_stride, void* d_amp_out, void* stream_) {
  if (!ctx) return PB200_EINVAL;
  if (nsrc <= 0 || nchan <= 0 || !d_amp_in || !d_amp_out || !d_scale || scale_stride <= 0)
    return pb_fail(ctx, PB200_EINVAL, "pb200_amp_scale: bad arguments");
  if (amp_dtype != PB200_AMP_F32 && amp_dtype != PB200_AMP_F64)
    return pb_fail(ctx, PB200_EINVAL, "pb200_amp_scale: amp_dtype must be PB200_AMP_F32 or PB200_AMP_F64");
  cudaStream_t stream = (cudaStream_t)stream_;
  PbDeviceGuard guard(ctx->device);
  const int nslab = (nchan + PB200_SLAB - 1) / PB200_SLAB, nsrc_pad = pb200_nsrc_pad(nsrc);
  const int blocks = ctx->sm_count * 8;
  if (amp_dtype == PB200_AMP_F32)
    k_amp_scale<float><<<blocks, 256, 0, stream>>>((const float*)d_amp_in, d_scale, scale_stride, nsrc, nsrc_pad, nslab, (float*)d_amp_out);
  else
    k_amp_scale<double><<<blocks, 256, 0, stream>>>((const double*)d_amp_in, d_scale, scale_stride, nsrc, nsrc_pad, nslab, (double*)d_amp_out);
  PB_CHECK_LAUNCH(ctx, "k_amp_scale");
  return PB200_OK;
}

int pb200_nsrc_pad(int nsrc) { return ((nsrc + PB200_SRC_TILE - 1) / PB200_SRC_TILE) * PB200_SRC_TILE; }

size_t pb200_amp_bytes(int nsrc, int nchan) {
  size_t nslab = (size_t)((nchan + PB200_SLAB - 1) / PB200_SLAB);
  size_t rows = (size_t)pb200_nsrc_pad(nsrc > 0 ? nsrc : 1);
  return nslab * rows * PB200_SLAB * sizeof(float);
}

int pb200_amp_table(pb200_ctx* ctx, const double* d_dircos, const int32_t* d_index, int nsrc,
                    const pb200_spectrum_desc* spec, const pb200_beam_desc* beam, const double* d_pbeam,
                    const double* h_freqs, int nchan, int amp_dtype, void* d_amp, void* stream_) {
  if (!ctx) return PB200_EINVAL;
  if (nsrc < 0 || nchan <= 0 || !spec || !beam || !h_freqs || !d_amp)
    return pb_fail(ctx, PB200_EINVAL, "pb200_amp_table: bad arguments");
  if (nsrc > 0 && !d_dircos) return pb_fail(ctx, PB200_EINVAL, "pb200_amp_table: null d_dircos");
  if (amp_dtype != PB200_AMP_F32 && amp_dtype != PB200_AMP_F64)
    return pb_fail(ctx, PB200_EINVAL, "pb200_amp_table: amp_dtype must be PB200_AMP_F32 or PB200_AMP_F64");
  if (beam->element < PB200_BEAM_DELTA || beam->element > PB200_BEAM_LOGTABLE)
    return pb_fail(ctx, PB200_EINVAL, "pb200_amp_table: unknown beam element type");
  if ((beam->element == PB200_BEAM_TABLE || beam->element == PB200_BEAM_LOGTABLE) && !d_pbeam)
    return pb_fail(ctx, PB200_EINVAL, "pb200_amp_table: PB200_BEAM_TABLE / LOGTABLE need d_pbeam");
  if (beam->element == PB200_BEAM_LOGTABLE && !beam->d_logmax)
    return pb_fail(ctx, PB200_EINVAL, "pb200_amp_table: PB200_BEAM_LOGTABLE needs d_logmax");
  if (beam->array_mode == PB200_ARRAY_ELEMENTS &&
      (beam->n_elements <= 0 || beam->nrand <= 0 || !beam->d_element_locs || !beam->d_delays))
    return pb_fail(ctx, PB200_EINVAL, "pb200_amp_table: element array needs locations, delays, nrand >= 1");
  if (beam->array_mode == PB200_ARRAY_ANALYTIC && beam->nax1 <= 0)
    return pb_fail(ctx, PB200_EINVAL, "pb200_amp_table: analytic array needs nax1 >= 1");
  if (!spec->d_spectrum && (!spec->d_flux_scale || !spec->d_index || !spec->d_freq_ref))
    return pb_fail(ctx, PB200_EINVAL, "pb200_amp_table: power-law spectrum needs scale, index and freq_ref");
  cudaStream_t stream = (cudaStream_t)stream_;
  PbDeviceGuard guard(ctx->device);
  const double* dfreq;
  int rc = pb_channels_device(ctx, h_freqs, nchan, nchan, stream, &dfreq);
  if (rc) return rc;
  AmpParams P;
  P.beam = *beam; P.spec = *spec;
  P.dircos = d_dircos; P.index = d_index; P.pbeam = d_pbeam; P.freqs = dfreq; P.amp = d_amp;
  P.nsrc = nsrc; P.nsrc_pad = pb200_nsrc_pad(nsrc > 0 ? nsrc : 1); P.nchan = nchan;
  P.nslab = (nchan + PB200_SLAB - 1) / PB200_SLAB;
  if (amp_dtype == PB200_AMP_F64) k_amp_table<double><<<P.nsrc_pad, AMP_THREADS, 0, stream>>>(P);
  else k_amp_table<float><<<P.nsrc_pad, AMP_THREADS, 0, stream>>>(P);
  PB_CHECK_LAUNCH(ctx, "k_amp_table");
  return PB200_OK;
}

}  // extern "C"
