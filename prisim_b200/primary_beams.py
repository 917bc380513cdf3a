"""GPU primary beams behind the reference's ``primary_beam_generator`` interface.

Mirrors prisim/primary_beams.py:9-441 (same argument names and meaning, same errors for bad
input) for the patterns on the hot path: presets 'hera', 'hirax', 'mwa', 'mwa_dipole', 'paper' and
custom shapes 'delta', 'dish', 'gaussian', 'dipole', with optional phased-array factor and ground
plane.  'vla'/'gmrt' (PBCOR polynomials) and 'rect'/'square' (broken in the reference) are out of
scope.  The evaluation itself is the fused CUDA kernel behind
``pb200_amp_table``; this module only translates the telescope dictionary into a
``pb200_beam_desc``.
"""
from __future__ import annotations

import numpy as NP
import scipy.constants as FCNST
import torch

from . import _lib, engine
from . import geometry as GEOM


def _to_dircos(vec, coords):
    vec = NP.asarray(vec, dtype=NP.float64).ravel()
    if coords == "altaz":
        if vec.size != 2:
            raise IndexError("Pointing center in Alt-Az coordinates must contain exactly two elements.")
        return GEOM.altaz2dircos(vec.reshape(1, 2), units="degrees")[0]
    if coords == "dircos":
        if vec.size != 3:
            raise IndexError("Pointing center in direction cosine coordinates must contain exactly three elements.")
        return vec
    raise ValueError("coordinates must be 'altaz' or 'dircos'")


def _dipole_orientation(telescope):
    """primary_beams.py:250-265 / :326-341."""
    if ("orientation" in telescope) and ("ocoords" in telescope):
        return _to_dircos(telescope["orientation"], telescope["ocoords"])
    if ("orientation" not in telescope) and ("ocoords" in telescope):
        if telescope["ocoords"] == "altaz":
            return _to_dircos([0.0, 90.0], "altaz")
        if telescope["ocoords"] == "dircos":
            return NP.asarray([1.0, 0.0, 0.0])
        raise ValueError('key "ocoords" in telescope dictionary contains invalid value')
    if ("orientation" in telescope) and ("ocoords" not in telescope):
        raise KeyError('key "ocoords" in telescope dictionary not specified.')
    return NP.asarray([1.0, 0.0, 0.0])


def _element_array(element_locs, pointing_info, device, rng=None):
    """array_field_pattern's beamformer settings (primary_beams.py:1599-1671) -> device arrays."""
    locs = NP.asarray(element_locs, dtype=NP.float64)
    if locs.ndim != 2 or locs.shape[1] not in (2, 3):
        raise ValueError("antpos must be a 2- or 3-column array")
    if locs.shape[1] == 2:
        locs = NP.hstack((locs, NP.zeros((locs.shape[0], 1))))
    ne = locs.shape[0]
    nrand = 1
    delays = NP.zeros(ne)
    gains = NP.ones(ne)
    if pointing_info is not None:
        nrand = pointing_info.get("nrand", 1)
        if nrand is None:
            nrand = 1
        if not isinstance(nrand, int):
            raise TypeError("nrand must be an integer")
        if nrand < 1:
            raise ValueError("nrand must be positive")
        if "delays" in pointing_info:
            if pointing_info["delays"] is not None:
                if not isinstance(pointing_info["delays"], NP.ndarray):
                    raise TypeError("delays must be a numpy array")
                if pointing_info["delays"].size != ne:
                    raise ValueError("size of delays must be equal to the number of antennas")
                delays = pointing_info["delays"].ravel().astype(NP.float64)
        elif "pointing_center" in pointing_info:
            if "pointing_coords" not in pointing_info:
                raise KeyError("pointing_coords not specified.")
            pc = _to_dircos(pointing_info["pointing_center"], pointing_info["pointing_coords"])
            delays = NP.dot(locs, pc) / FCNST.c            # delay compensation (:1632)
        if pointing_info.get("gains", None) is not None:
            if pointing_info["gains"].size != ne:
                raise ValueError("size of gains must be equal to the number of antennas")
            gains = NP.asarray(pointing_info["gains"], dtype=NP.float64).ravel()
        rng = NP.random if rng is None else rng
        delays = NP.repeat(delays.reshape(ne, 1), nrand, axis=1)
        gains = NP.repeat(gains.reshape(ne, 1), nrand, axis=1)
        if pointing_info.get("delayerr", None) is not None:
            if pointing_info["delayerr"] < 0.0:
                raise ValueError("delayerr must be non-negative")
            delays = delays + pointing_info["delayerr"] * rng.standard_normal((ne, nrand))       # (:1655)
        if pointing_info.get("gainerr", None) is not None:
            if pointing_info["gainerr"] < 0.0:
                raise ValueError("gainerr must be non-negative")
            gains = gains * 10 ** ((pointing_info["gainerr"] / 10.0) * rng.standard_normal((ne, nrand)))   # (:1665-1666)
    else:
        delays = delays.reshape(ne, 1)
        gains = gains.reshape(ne, 1)
    return dict(n_elements=ne, nrand=nrand, d_element_locs=engine._f64(locs, device),
                d_delays=engine._f64(delays, device), d_gains=engine._f64(gains, device))


def beam_desc_from_telescope(telescope, pointing_info=None, pointing_center=None, skyunits="altaz", east2ax1=0.0,
                             short_dipole_approx=False, half_wave_dipole_approx=False, achromatic_freq_hz=None,
                             device=0, rng=None):
    """Translate the reference's telescope / pointing_info dictionaries into a ``pb200_beam_desc``
    following the branch structure of primary_beams.py:224-439."""
    if (telescope is None) or (not isinstance(telescope, dict)):
        raise TypeError("telescope must be specified as a dictionary")
    kw = {}
    dipole_mode = _lib.DIPOLE_GENERAL
    if short_dipole_approx:
        dipole_mode = _lib.DIPOLE_SHORT
    elif half_wave_dipole_approx:
        dipole_mode = _lib.DIPOLE_HALFWAVE
    shape_for_ground = telescope.get("shape", None)
    if "id" in telescope:
        tid = telescope["id"]
        if (tid == "hera") or (tid == "hirax"):                               # :239-247
            kw.update(element=_lib.BEAM_AIRY, size=14.0 if tid == "hera" else 6.0,
                      pointing=_to_dircos(telescope["orientation"], telescope["ocoords"]))
        elif tid == "mwa":                                                    # :248-319
            if skyunits not in ("altaz", "dircos"):
                raise ValueError("skyunits must be in Alt-Az or direction cosine coordinates for MWA.")
            kw.update(element=_lib.BEAM_DIPOLE, size=0.74, orientation=_dipole_orientation(telescope),
                      dipole_mode=dipole_mode)
            if pointing_info is None:                                         # :273-285
                kw.update(array_mode=_lib.ARRAY_ANALYTIC, nax1=4, nax2=4, sep1=1.1, sep2=1.1,
                          east2ax1_deg=east2ax1, array_pointing=(0.0, 0.0, 1.0))
            else:                                                             # :287-316
                if "element_locs" not in telescope:
                    xlocs, ylocs = NP.meshgrid(1.1 * NP.linspace(-1.5, 1.5, 4), 1.1 * NP.linspace(1.5, -1.5, 4))
                    element_locs = NP.hstack((xlocs.reshape(-1, 1), ylocs.reshape(-1, 1), NP.zeros(xlocs.size).reshape(-1, 1)))
                else:
                    element_locs = telescope["element_locs"]
                kw.update(array_mode=_lib.ARRAY_ELEMENTS, **_element_array(element_locs, pointing_info, device, rng))
        elif (tid == "mwa_dipole") or (tid == "paper"):                       # :320-349
            if skyunits not in ("altaz", "dircos"):
                raise ValueError("skyunits must be in Alt-Az or direction cosine coordinates for MWA dipole.")
            kw.update(element=_lib.BEAM_DIPOLE, size=0.74 if tid == "mwa_dipole" else 2.0,
                      orientation=_dipole_orientation(telescope), dipole_mode=dipole_mode)
        elif (tid == "vla") or ("gmrt" in tid):
            raise NotImplementedError("VLA/GMRT PBCOR polynomial beams are outside the hot-path scope")
        else:
            raise ValueError("No presets available for the specified telescope ID. Set custom parameters instead in input parameter telescope.")
    else:                                                                     # :354-416
        shape = telescope.get("shape", "delta")
        pc = (0.0, 0.0, 1.0) if pointing_center is None else _to_dircos(pointing_center, skyunits)
        if shape == "delta":
            kw.update(element=_lib.BEAM_DELTA)
        elif shape == "dipole":
            orient = _to_dircos(telescope["orientation"], telescope["ocoords"])
            kw.update(element=_lib.BEAM_DIPOLE, size=telescope["size"], orientation=orient, dipole_mode=dipole_mode)
        elif shape == "dish":
            kw.update(element=_lib.BEAM_AIRY, size=telescope["size"], pointing=pc)
        elif shape == "gaussian":
            kw.update(element=_lib.BEAM_GAUSSIAN, size=telescope["size"], pointing=pc)
        elif shape in ("rect", "square"):
            raise NotImplementedError("rect/square apertures raise NameError in the reference and are out of scope")
        else:
            raise ValueError('Value in key "shape" of telescope dictionary invalid.')
        if (pointing_info is not None) and ("element_locs" in telescope):     # :385-412
            kw.update(array_mode=_lib.ARRAY_ELEMENTS, **_element_array(telescope["element_locs"], pointing_info, device, rng))
    if "groundplane" in telescope and telescope["groundplane"] is not None:  # :418-439
        if shape_for_ground != "dish":
            kw["groundplane"] = float(telescope["groundplane"])
            mod = telescope.get("ground_modify", None)
            if isinstance(mod, dict):
                kw["ground_scale"] = mod.get("scale", 1.0)
                kw["ground_max"] = mod.get("max", 0.0)
    if achromatic_freq_hz is not None:
        kw.update(achromatic=1, ref_freq_hz=float(achromatic_freq_hz))
    return engine.make_beam_desc(**kw)


class HealpixBeam(object):
    """External gridded beam (HEALPix RING map) prepared for the device gather ``pb200_healpix_beam``.

    The reference (scripts/run_prisim.py:482-524, :1897-1908) interpolates log10(external_beam [npix, nfreq_b])
    bilinearly on the sphere at every source and then along frequency onto the observing channels, per snapshot.
    Both interpolations are linear in the map, so here the map is resampled to the channels ONCE (this constructor,
    scipy interp1d of kind `spec_interp`; `chromatic=False` takes the map nearest to `select_freq` for all channels,
    :1902-1903) and stored on the GPU as [npix, nchan]; per snapshot the device only gathers."""

    def __init__(self, external_beam, beam_freqs_hz, channels_hz, chromatic=True, select_freq=None, spec_interp="cubic",
                 dtype=torch.float32, device=None):
        from scipy import interpolate
        external_beam = NP.asarray(external_beam, dtype=NP.float64)
        if external_beam.ndim != 2:
            raise ValueError("external_beam must be [npix, nfreq]")
        npix = external_beam.shape[0]
        nside = int(round(NP.sqrt(npix / 12.0)))
        if 12 * nside * nside != npix or nside & (nside - 1):
            raise ValueError("external_beam must be a HEALPix map with nside a power of two")
        beam_freqs_hz = NP.asarray(beam_freqs_hz, dtype=NP.float64).ravel()
        channels_hz = NP.asarray(channels_hz, dtype=NP.float64).ravel()
        if spec_interp == "fft":
            raise NotImplementedError("'fft' spectral interpolation of external beams is not supported; use 'cubic' or 'linear'")
        with NP.errstate(divide="ignore"):
            logbeam = NP.log10(external_beam)
        if chromatic:
            if beam_freqs_hz.size != external_beam.shape[1]:
                raise ValueError("beam_freqs_hz must match the second axis of external_beam")
            logmap = interpolate.interp1d(beam_freqs_hz, logbeam, axis=1, kind=spec_interp, bounds_error=False,
                                          fill_value="extrapolate")(channels_hz)
        else:
            nearest = int(NP.argmin(NP.abs(beam_freqs_hz - (channels_hz[channels_hz.size // 2] if select_freq is None else select_freq))))
            logmap = NP.repeat(logbeam[:, nearest].reshape(-1, 1), channels_hz.size, axis=1)
        self.nside, self.nchan = nside, channels_hz.size
        self.device = engine._dev(device)
        self.map = torch.as_tensor(NP.ascontiguousarray(logmap)).to(device="cuda:{0}".format(self.device), dtype=dtype).contiguous()

    def table(self, dircos, nsrc):
        """(log10 beam [nsrc, nchan], per-channel max(0, max)) at culled ENU direction cosines (device tensors)."""
        return engine.healpix_beam(self.map, self.nside, dircos, nsrc, self.nchan)

    def pbeam(self, skypos_altaz_deg):
        """Power beam [nsrc, nchan] as numpy -- what run_prisim puts in roi_info['pbeam'] (:1908)."""
        altaz = NP.asarray(skypos_altaz_deg, dtype=NP.float64).reshape(-1, 2)
        dircos, _ = engine.sky_cull(altaz, "altaz", roi_radius_deg=180.0, device=self.device)
        logbeam, colmax = self.table(dircos, altaz.shape[0])
        return torch.pow(10.0, logbeam - colmax.unsqueeze(0)).cpu().numpy()


def primary_beam_generator(skypos, frequency, telescope, freq_scale="GHz", skyunits="degrees", east2ax1=0.0,
                           pointing_info=None, pointing_center=None, short_dipole_approx=False,
                           half_wave_dipole_approx=False, device=None):
    """Same call as prisim/primary_beams.py:9-12.  Returns the power pattern [nsrc, nchan] as a
    numpy float64 array holding fp32-precision values (the device amplitude table is fp32).
    skyunits must be 'altaz' or 'dircos' (the 'degrees' form only serves the vla/gmrt presets)."""
    try:
        skypos, frequency, telescope
    except NameError:
        raise NameError("Sky positions, frequency and telescope inputs must be specified.")
    scale = {"ghz": 1.0e9, "mhz": 1.0e6, "khz": 1.0e3}.get(str(freq_scale).lower(), 1.0)     # :212-217
    frequency = (NP.asarray(frequency, dtype=NP.float64) * scale).reshape(-1)
    if skyunits not in ("altaz", "dircos"):
        raise ValueError('skyunits must be "altaz" or "dircos" for the beams on the hot path')
    device = engine._dev(device)
    skypos = NP.asarray(skypos, dtype=NP.float64)
    skypos = skypos.reshape(-1, 2 if skyunits == "altaz" else 3)
    beam = beam_desc_from_telescope(telescope, pointing_info=pointing_info, pointing_center=pointing_center,
                                    skyunits=skyunits, east2ax1=east2ax1, short_dipole_approx=short_dipole_approx,
                                    half_wave_dipole_approx=half_wave_dipole_approx, device=device)
    # keep every position (no horizon cull): radius 180 degrees about the zenith
    dircos, index = engine.sky_cull(skypos, skyunits, roi_radius_deg=180.0, device=device)
    nsrc = skypos.shape[0]
    ones = torch.ones(nsrc, dtype=torch.float64, device=dircos.device)
    spec = {"flux_scale": ones, "index": torch.zeros_like(ones), "freq_ref": ones}
    amp = engine.amp_table(dircos, index, nsrc, spec, beam, frequency, device=device)
    return engine.amp_table_to_dense(amp, nsrc, frequency.size).double().cpu().numpy()
